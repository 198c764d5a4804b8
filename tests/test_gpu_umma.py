"""GPU: tcgen05 / TMEM / bulk-copy plumbing of the fused MLP kernel, checked on a plain GEMM."""
import numpy as np
import pytest
import torch

from neurecon_b200 import _lib, umma_pack

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run(K, N, variant=0, seed=0):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(128, K, generator=g)
    B = torch.randn(K, N, generator=g)
    img = umma_pack.pack_a_tiles(A.to(DEV))
    D = torch.full((128, N), float("nan"), device=DEV)
    Bd = B.to(DEV).contiguous()
    lib = _lib.get_devtools()
    _lib.check(lib.nr_selftest_umma(_lib.ptr(img), _lib.ptr(Bd), K, N, _lib.ptr(D), variant, _lib.stream_ptr()), "selftest")
    torch.cuda.synchronize()
    want = A.to(torch.bfloat16).double() @ B.to(torch.bfloat16).double()
    err = ((D.double().cpu() - want).abs().max() / want.abs().max()).item()
    return err


@pytest.mark.parametrize("K,N", [(16, 32), (64, 128), (48, 128), (256, 128), (256, 32), (128, 64)])
def test_umma_gemm_matches_bf16_matmul(K, N):
    err = _run(K, N)
    if not err < 1e-5:
        alts = {v: _run(K, N, v) for v in (1, 2, 3)}
        pytest.fail("production encoding err=%g; variants: %s" % (err, alts))


def _run2(K, N, variant=0, seed=0):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(256, K, generator=g)
    B = torch.randn(K, N, generator=g)
    img = umma_pack.pack_a_tiles(A.to(DEV), dtype=torch.float16)     # k-chunk major, M-tile minor
    D = torch.full((256, N), float("nan"), device=DEV)
    Bd = B.to(DEV).contiguous()
    lib = _lib.get_devtools()
    _lib.check(lib.nr_selftest_umma2(_lib.ptr(img), _lib.ptr(Bd), K, N, _lib.ptr(D), variant, _lib.stream_ptr()), "selftest2")
    torch.cuda.synchronize()
    want = A.to(torch.float16).double() @ B.to(torch.float16).double()
    return ((D.double().cpu() - want).abs().max() / want.abs().max()).item()


@pytest.mark.parametrize("K,N,variant", [(64, 128, 0), (256, 256, 0), (256, 256, 1), (128, 64, 1), (256, 128, 1), (256, 256, 3), (64, 128, 2)])
def test_umma_cta_pair_gemm(K, N, variant):
    """tcgen05.mma.cta_group::2 on a 2-CTA cluster: M = 256 split over the pair, B halves local (variant 0) or written
    into the peer's shared memory through DSMEM (variant bit 0); hand-off by cluster barrier or (variant bit 1) by
    release.cluster arrivals of all warps on the leader's mbarrier, as in the fused pair kernel."""
    assert _run2(K, N, variant) < 1e-5


def test_device_packer_matches_torch_packer():
    """nr_umma_pack_a (one launch, any strides) against the torch composition it replaces, bit for bit: plain, transposed view,
    broadcast row (stride 0), bf16, and the (hi, lo) split image"""
    rs = np.random.RandomState(0)
    for rows, K, n_mt, k_pad in ((256, 256, 2, 256), (217, 39, 2, 64), (39, 256, 1, 256), (3, 130, 1, 192)):
        W = torch.from_numpy(rs.normal(size=(rows, K)).astype(np.float32))
        for dt in (torch.float16, torch.bfloat16):
            want = umma_pack.pack_a_tiles(W, n_mtiles=n_mt, k_pad=k_pad, dtype=dt)
            got = umma_pack.pack_a_tiles(W.to(DEV), n_mtiles=n_mt, k_pad=k_pad, dtype=dt)
            assert got.shape == want.shape and torch.equal(got.cpu(), want)
        Wt = torch.from_numpy(rs.normal(size=(K, rows)).astype(np.float32))
        want = umma_pack.pack_a_tiles(Wt.t().contiguous(), n_mtiles=n_mt, k_pad=k_pad, dtype=torch.float16)
        got = umma_pack.pack_a_tiles(Wt.to(DEV).t(), n_mtiles=n_mt, k_pad=k_pad, dtype=torch.float16)
        assert torch.equal(got.cpu(), want)
        want = umma_pack.pack_a_tiles_split(W, n_mtiles=n_mt, k_pad=k_pad)
        got = umma_pack.pack_a_tiles_split(W.to(DEV), n_mtiles=n_mt, k_pad=k_pad)
        assert got.shape == want.shape and torch.equal(got.cpu(), want)
    row = torch.from_numpy(rs.normal(size=(1, 256)).astype(np.float32))
    want = umma_pack.pack_a_tiles(row.expand(32, 256).contiguous(), n_mtiles=1, k_pad=256, dtype=torch.float16)
    got = umma_pack.pack_a_tiles(row.to(DEV).expand(32, 256), n_mtiles=1, k_pad=256, dtype=torch.float16)
    assert torch.equal(got.cpu(), want)
