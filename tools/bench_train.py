"""NeuS training step (512 rays: render under autograd + L1 + eikonal + mask BCE, backward) -- ms per step and
rays/s on one GPU, or per rank under torchrun.  Usage: python tools/bench_train.py [rays] [steps] [graph]
``graph``: the whole iteration (render, losses, backward, all-reduce, FusedAdam) replayed as one CUDA graph
(train_util.CapturedStep) instead of ~1 100 eager launches."""
import os, sys, time
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import neurecon_b200
from conftest import build_neus
from neurecon_b200.models.frameworks import neus
from neurecon_b200.utils import synthetic

if os.environ.get("NEURECON_B200_PRECISION"):
    neurecon_b200.set_precision(os.environ["NEURECON_B200_PRECISION"])
from neurecon_b200.utils import dist_util
R = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
# under torchrun: one rank per GPU, R rays per rank, ONE flat gradient all-reduce per step (train.py:124 semantics)
rank, local_rank, world = dist_util.init_env()
dev = torch.device("cuda", local_rank)
m = build_neus(seed=1, device=dev)
GRAPH = len(sys.argv) > 3 and sys.argv[3] == "graph"
if GRAPH:
    from neurecon_b200.utils import train_util
    opt = train_util.FusedAdam(m.parameters(), lr=5e-4, capturable=True)
else:
    opt = torch.optim.Adam(m.parameters(), lr=5e-4)
o, d = synthetic.make_rays(R, seed=3 + rank)
o, d = o.to(dev), d.to(dev)
target = torch.rand(R, 3, device=dev)


ones = torch.ones(R, device=dev)


def step_eager(o, d, target):
    opt.zero_grad(set_to_none=not GRAPH)
    rgb, _, ret = neus.volume_render(o, d, m, detailed_output=True, perturb=True)
    nn_ = ret["implicit_nablas"].norm(dim=-1)
    loss = F.l1_loss(rgb, target) + 0.1 * F.mse_loss(nn_, torch.ones_like(nn_)) \
        + F.binary_cross_entropy(ret["mask_volume"].clamp(1e-3, 1 - 1e-3), ones)
    loss.backward()
    dist_util.allreduce_gradients(m.parameters())
    opt.step()
    return loss


step = step_eager
if GRAPH:
    step = train_util.CapturedStep(step_eager, (o, d, target), optimizer=opt, warmup=3)
else:
    for _ in range(3):
        step(o, d, target)
torch.cuda.synchronize()
if world > 1:
    torch.distributed.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    loss = step(o, d, target)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
if world > 1:
    t = torch.tensor([ms], device=dev)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms = t.item()
    if rank == 0:
        print("[%s tier%s] NeuS data-parallel training step, %d x %d rays: %.2f ms/step (max over ranks), %.0f rays/s"
              % (neurecon_b200.get_precision(), ", one CUDA graph" if GRAPH else "", world, R, ms, world * R / ms * 1e3))
        sys.stdout.flush()
    del step                      # a live CUDA graph holding NCCL kernels must go before the communicator does
    torch.cuda.synchronize()
    torch.distributed.destroy_process_group()
    sys.exit(0)
print("[%s tier%s] NeuS training step, %d rays: %.2f ms/step, %.0f rays/s, loss %.4f, peak mem %.2f GB (algorithmic ~1.85 GFLOP/ray => %.1f TFLOP/s)"
      % (neurecon_b200.get_precision(), ", one CUDA graph" if GRAPH else "", R, ms, R / ms * 1e3, loss.item(), torch.cuda.max_memory_allocated() / 2**30, R * 1.85e9 / (ms * 1e-3) / 1e12))
