"""Load the UNMODIFIED reference from /root/reference with import stubs
(SURVEY.md appendix B).  Only usable in the build container; the GPU box has no
/root/reference, so this is imported solely by tests/golden/make_golden.py and
by CPU tests that skip when the reference is absent.  TEST INFRASTRUCTURE ONLY.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("NEURECON_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "models", "frameworks"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _AttrDict(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def load():
    """Return a namespace with the reference's hot-path modules."""
    if not available():
        raise RuntimeError("reference not present at %s" % REF_ROOT)
    _stub("addict", Dict=_AttrDict)
    _stub("imageio")
    sk = _stub("skimage")
    sk.transform = _stub("skimage.transform", rescale=None)
    sk.measure = _stub("skimage.measure")
    mpl = _stub("matplotlib")
    mpl.pyplot = _stub("matplotlib.pyplot")
    _stub("plyfile")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from models.frameworks import neus, volsdf, unisurf
        from models import base, ray_casting
        from utils import rend_util, train_util
    return types.SimpleNamespace(neus=neus, volsdf=volsdf, unisurf=unisurf, base=base,
                                 ray_casting=ray_casting, rend_util=rend_util, train_util=train_util)
