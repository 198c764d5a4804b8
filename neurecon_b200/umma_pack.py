"""Host-side packing of MLP weights into the bf16, pre-swizzled shared-memory images the
tcgen05 kernels stream with bulk async copies (layout: csrc/umma.cuh)."""
import numpy as np
import torch

_A_IDX = None


def _a_tile_index():
    """uint16 index of element (r, k) inside a [128 x 64] K-major SWIZZLE_128B tile."""
    global _A_IDX
    if _A_IDX is None:
        r = np.arange(128)[:, None]
        k = np.arange(64)[None, :]
        off = r * 128 + ((((k >> 3) ^ (r & 7)) & 7) << 4) + ((k & 7) << 1)
        _A_IDX = torch.from_numpy((off // 2).astype(np.int64))
    return _A_IDX


def pack_a_tiles(W, n_mtiles=None, k_pad=None):
    """W: [rows, K] float tensor -> int16 tensor [n_mtiles * n_kchunks, 8192] of A tiles ordered
    (mt major, kc minor); rows/K zero-padded to 128 / 64 multiples."""
    rows, K = W.shape
    n_mt = n_mtiles if n_mtiles is not None else (rows + 127) // 128
    kp = k_pad if k_pad is not None else K
    n_kc = (kp + 63) // 64
    full = torch.zeros(n_mt * 128, n_kc * 64, dtype=torch.float32, device=W.device)
    full[:rows, :K] = W.float()
    bf = full.to(torch.bfloat16).view(torch.int16)
    tiles = bf.reshape(n_mt, 128, n_kc, 64).permute(0, 2, 1, 3).reshape(n_mt * n_kc, 128 * 64)
    idx = _a_tile_index().to(W.device).reshape(-1)
    out = torch.empty_like(tiles)
    out[:, idx] = tiles
    return out.contiguous()
