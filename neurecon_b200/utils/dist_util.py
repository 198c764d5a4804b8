"""Data-parallel helpers: the only parallelism in the reference is ray sharding for rendering
(``nn.DataParallel(dim=1)``, neus.py:413-414) and DDP's gradient all-reduce in training
(train.py:124,205; utils/dist_util.py:13-39).  One process per GPU, torch.distributed (NCCL over
NVLink on the GPU box, gloo in CPU tests)."""
import os

import torch
import torch.distributed as dist


def get_rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def get_world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def init_env(backend=None):
    """torchrun-style bootstrap (utils/dist_util.py:13-39 of the reference): reads RANK / LOCAL_RANK /
    WORLD_SIZE, binds the GPU, creates the process group.  Returns (rank, local_rank, world_size)."""
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    return rank, local_rank, world


def shard_range(n, rank=None, world=None):
    """Contiguous range [lo, hi) of `n` independent units (rays, grid planes) owned by `rank`;
    the first n % world ranks get one extra unit.  Covers [0, n) exactly, in rank order."""
    rank = get_rank() if rank is None else rank
    world = get_world_size() if world is None else world
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_rays(local, n_total, dim=0):
    """All-gather per-rank ray shards (possibly ragged by one) back into rank order along `dim`."""
    world = get_world_size()
    if world == 1:
        return local
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    longest = max(hi - lo for lo, hi in sizes)
    pad_shape = list(local.shape)
    pad_shape[dim] = longest
    padded = local.new_zeros(pad_shape)
    padded.narrow(dim, 0, local.shape[dim]).copy_(local)
    outs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(outs, padded)
    return torch.cat([o.narrow(dim, 0, hi - lo) for o, (lo, hi) in zip(outs, sizes)], dim=dim)


def allreduce_gradients(parameters, average=True):
    """ONE flat-buffer sum all-reduce of every gradient (3.2 MB for NeuS, 5.6 MB with NeRF++: latency-
    bound, so a single call instead of DDP's buckets), then scatter back.  `average=True` reproduces
    DDP's mean reduction (train.py:124)."""
    world = get_world_size()
    params = [p for p in parameters if p.grad is not None]
    if world == 1 or not params:
        return
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if average:
        flat /= world
    off = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n


def broadcast_parameters(module, src=0):
    """Replicate the (<= 5.6 MB) weights once, e.g. after loading a checkpoint on rank 0."""
    if get_world_size() == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src)
