"""Data-parallel NeuS training step over NCCL (run under torchrun, one rank per GPU):
every rank renders its contiguous share of the rays under autograd, gradients are summed with ONE flat
all-reduce (dist_util.allreduce_gradients, mean like DDP), and rank 0 checks them against the single-GPU gradient
of the mean loss over all rays.  Also times the all-reduce of the 802 491 fp32 gradients (3.2 MB)."""
import os, sys
import torch
import torch.distributed as dist
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import neurecon_b200
from conftest import build_neus, rel_err
from neurecon_b200.models.frameworks import neus
from neurecon_b200.utils import dist_util, synthetic

rank, local_rank, world = dist_util.init_env()
dev = torch.device("cuda", local_rank)
neurecon_b200.set_precision("fp32")
R = 64 * world
o, d = synthetic.make_rays(R, seed=3)
target = torch.rand(R, 3, generator=torch.Generator().manual_seed(1))


def loss_on(model, lo, hi):
    rgb, _, ret = neus.volume_render(o[lo:hi].to(dev), d[lo:hi].to(dev), model, detailed_output=True, perturb=False)
    nn_ = ret["implicit_nablas"].norm(dim=-1)
    return F.l1_loss(rgb, target[lo:hi].to(dev)) + 0.1 * F.mse_loss(nn_, torch.ones_like(nn_))


m = build_neus(seed=1, device=dev)
dist_util.broadcast_parameters(m)
lo, hi = dist_util.shard_range(R)
m.zero_grad()
loss_on(m, lo, hi).backward()
dist_util.allreduce_gradients(m.parameters())
torch.cuda.synchronize()
ok = True
if rank == 0:
    ref = build_neus(seed=1, device=dev)
    ref.zero_grad()
    sum(loss_on(ref, *dist_util.shard_range(R, r, world)) for r in range(world)).div(world).backward()
    worst = max(rel_err(p.grad, q.grad) for p, q in zip(m.parameters(), ref.parameters()))
    print("world %d: data-parallel gradients vs single-GPU mean-loss gradients: max rel err %.2e" % (world, worst))
    ok = worst < 1e-4
# all ranks hold identical gradients
g = torch.cat([p.grad.reshape(-1) for p in m.parameters()])
g0 = g.clone()
dist.broadcast(g0, 0)
same = torch.equal(g, g0)
# time the flat all-reduce
flat = torch.zeros(802491, device=dev)
for _ in range(5):
    dist.all_reduce(flat)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    dist.all_reduce(flat)
e1.record()
torch.cuda.synchronize()
if rank == 0:
    print("identical on all ranks: %s; flat 3.2 MB all-reduce: %.1f us" % (same, e0.elapsed_time(e1) / 50 * 1e3))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if (ok and same) else 1)
