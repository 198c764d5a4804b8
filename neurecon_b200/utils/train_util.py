"""Drop-in for the hot-path part of the reference's ``utils/train_util.py``."""
import torch


def batchify_query(query_fn, *args, chunk, dim_batchify):
    """utils/train_util.py:23-71: flatten ``[(B), N_rays, N_pts, ...]`` to points, call
    ``query_fn`` on slices of ``chunk`` points and restore the ``[(B), N_rays, N_pts, ...]``
    shape.  Kept for callers that chunk explicitly; the fused kernels tile internally, so
    the neurecon_b200 renderers do not need it to bound memory."""
    n_rays = args[0].shape[dim_batchify]
    n_pts = args[0].shape[dim_batchify + 1]
    flat = [a.flatten(dim_batchify, dim_batchify + 1) for a in args]
    total = flat[0].shape[dim_batchify]
    if dim_batchify not in (0, 1, 2):
        raise NotImplementedError
    pieces = []
    for i in range(0, total, chunk):
        sl = [a.narrow(dim_batchify, i, min(chunk, total - i)) for a in flat]
        r = query_fn(*sl)
        pieces.append(r if isinstance(r, tuple) else [r])

    def restore(v):
        return v.reshape([*v.shape[:dim_batchify], n_rays, n_pts, *v.shape[dim_batchify + 1:]])

    out = []
    for entry in zip(*pieces):
        if isinstance(entry[0], dict):
            merged = {}
            for d in entry:
                for k, v in d.items():
                    merged.setdefault(k, []).append(v)
            out.append({k: restore(torch.cat(v, dim=dim_batchify)) for k, v in merged.items()})
        else:
            out.append(restore(torch.cat(entry, dim=dim_batchify)))
    return out[0] if len(out) == 1 else tuple(out)
