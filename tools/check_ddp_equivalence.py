"""torchrun --nproc-per-node 2 tools/check_ddp_equivalence.py: with the SAME batch on every rank, N-rank training
(averaged gradients, SyncBatchNorm statistics, bucketed all-reduce overlapped with the backward pass) must follow the
single-process run: compares losses and parameters after three steps on rank 0.  GPU box only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from scd_resnet_b200 import ops, synthetic
from scd_resnet_b200.centerNetOffset import CenterNetResidual
from scd_resnet_b200.training import TrainEngine

rank = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(rank)
dist.init_process_group("nccl")
x = synthetic.make_tiles(4, seed=0).cuda()
locs, counts = synthetic.make_objects(4, seed=3)
tg = list(ops.render_targets(locs.cuda(), counts.cuda(), with_npos=True))


def run(group):
    m = CenterNetResidual(10).cuda()
    m.load_state_dict(synthetic.make_state_dict(m, 1234))
    m.train()
    eng = TrainEngine(m, process_group=group)
    losses = [eng.train_step(x, tg).clone() for _ in range(3)]
    torch.cuda.synchronize()
    return torch.stack(losses).cpu(), {k: v.detach().float().cpu().clone() for k, v in m.state_dict().items()}


l2, p2 = run(dist.group.WORLD)
if rank == 0:
    l1, p1 = run(None)
    print("losses  N ranks:", l2[:, 0].tolist(), " single:", l1[:, 0].tolist())
    worst = max(((p2[k] - p1[k]).abs().max().item() / (p1[k].abs().max().item() + 1e-12), k) for k in p1 if p1[k].dtype.is_floating_point)
    print("largest relative parameter difference:", worst)
    # identical data on all ranks: sums over ranks = world x the single sums, so the only differences are fp rounding of
    # the all-reduced sums; Adam's sign-like first steps amplify a flipped near-zero gradient to 2 lr
    assert (l2 - l1).abs().max() <= 2e-3 * l1.abs().max(), (l2, l1)
    assert worst[0] < 5e-2, worst
    print("OK")
dist.barrier()
dist.destroy_process_group()
