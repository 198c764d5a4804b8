"""CPU: the C-ABI library builds, loads and exports every symbol include/neurecon_b200.h declares;
the host-side mirror of the reference interface behaves (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT, build_neus


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "neurecon_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(nr_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from neurecon_b200 import _lib, build
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    header = _header_symbols()
    assert header, "no symbols parsed from the header"
    for name in header:
        assert hasattr(lib, name), "library lacks %s" % name
    assert set(header) == set(_lib.declared_symbols()), (
        set(header) ^ set(_lib.declared_symbols()))
    assert lib.nr_version() >= 100


def test_error_convention():
    from neurecon_b200 import _lib
    lib = _lib.get_lib()
    # invalid arguments are reported through the return code + nr_last_error, never by throwing
    rc = lib.nr_sample_pdf(None, None, None, 4, 8, 4, 0, 1e-5, None, None, None, None, None)  # R=4, null data
    assert rc == -1
    assert "null" in _lib.last_error()
    with pytest.raises(ValueError):
        _lib.check(rc, "sample_pdf")


def test_cpu_tensors_raise_no_fallback():
    from neurecon_b200.utils import rend_util
    m = build_neus()
    x = torch.zeros(4, 3)
    with torch.no_grad():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            m.implicit_surface.forward(x)
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            rend_util.near_far_from_sphere(x, x)
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            rend_util.sample_pdf(torch.zeros(2, 8), torch.zeros(2, 7), 4, det=True)


def test_unknown_upsample_algo_raises_like_reference():
    from neurecon_b200.models.frameworks import neus
    m = build_neus()
    with pytest.raises(NotImplementedError):
        neus.volume_render(torch.zeros(2, 3), torch.ones(2, 3), m, upsample_algo="nope")


def test_batchify_query_shapes():
    from neurecon_b200.utils.train_util import batchify_query
    x = torch.arange(2 * 5 * 7 * 3, dtype=torch.float32).reshape(2, 5, 7, 3)

    def fn(p):
        return p.sum(-1), {"twice": p * 2}

    s, d = batchify_query(fn, x, chunk=11, dim_batchify=1)
    assert torch.equal(s, x.sum(-1)) and torch.equal(d["twice"], x * 2)
    s0 = batchify_query(lambda p: p[..., 0], x[0], chunk=4, dim_batchify=0)
    assert torch.equal(s0, x[0, ..., 0])
