"""CPU: the N>1 host logic (ray sharding, gather, flat gradient all-reduce) on world_size-2 gloo."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neurecon_b200.utils import dist_util


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 442368, 1000003):
        for world in (1, 2, 3, 4, 8):
            rs = [dist_util.shard_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            sizes = [hi - lo for lo, hi in rs]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_rays):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, lr, w = dist_util.init_env(backend="gloo")
    assert (r, w) == (rank, world)
    # ray sharding + gather: every rank "renders" f(ray) on its range; the gather restores rank order
    rays = torch.arange(n_rays, dtype=torch.float32)[:, None].expand(n_rays, 3)
    lo, hi = dist_util.shard_range(n_rays)
    local = rays[lo:hi] * 2 + 1
    full = dist_util.gather_rays(local, n_rays)
    assert torch.equal(full, rays * 2 + 1)
    # flat gradient all-reduce == mean over ranks of per-shard gradients == gradient of the mean loss
    torch.manual_seed(0)
    lin = torch.nn.Linear(3, 2)
    dist_util.broadcast_parameters(lin)
    x = torch.randn(n_rays, 3, generator=torch.Generator().manual_seed(1))
    lin(x[lo:hi]).square().mean().backward()
    dist_util.allreduce_gradients(lin.parameters())
    ref = torch.nn.Linear(3, 2)
    ref.load_state_dict(lin.state_dict())
    sum(ref(x[a:b]).square().mean() for a, b in (dist_util.shard_range(n_rays, q, world) for q in range(world))).div(world).backward()
    for p, q in zip(lin.parameters(), ref.parameters()):
        assert torch.allclose(p.grad, q.grad, atol=1e-6), (p.grad, q.grad)
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_sharding_and_grad_allreduce():
    mp.spawn(_worker, args=(2, _free_port(), 11), nprocs=2, join=True)
