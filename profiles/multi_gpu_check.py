"""torchrun --nproc-per-node N profiles/multi_gpu_check.py : the copy-engine (symmetric memory) candidate gather against the
NCCL all-gather, bit for bit, and the two-rank global-negatives loss against the single-process decomposition."""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.getcwd())
from brainmagick_b200 import distrib, functional as BF
import brainmagick_b200 as bb

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(100 + rank)
B, F, T = 16, 1024, 360
cand = torch.randn(B, F, T, device=dev)
for trial in range(3):
    g = distrib.CandidateGather(cand, uniform=True)
    used_symm = g.event is not None
    out, off = g.wait()
    ref = torch.empty(world * B, F, T, device=dev)
    dist.all_gather_into_tensor(ref, cand)
    torch.cuda.synchronize()
    assert off == rank * B and torch.equal(out, ref), "symmetric gather differs from NCCL all-gather"
    cand = cand * 1.5 + trial                        # a different payload each step: the buffer reuse barrier is exercised
# loss: every rank holds the same global tensors, takes its shard
torch.manual_seed(7)
est_all = (torch.randn(world * B, F, T, device=dev) * 0.01)
cand_all = torch.randn(world * B, F, T, device=dev)
mine = slice(rank * B, rank * B + B)
clip = bb.ClipLoss(global_negatives=True, uniform_batches=True).to(dev).train()
e = est_all[mine].clone().requires_grad_(True)
c = cand_all[mine].clone()
clip.prefetch_candidates(c)
loss = clip(e, c, torch.ones(B, 1, T, dtype=torch.bool, device=dev))
loss.backward()
ref_e = est_all[mine].clone().requires_grad_(True)
ref_loss = BF.clip_loss(ref_e, cand_all, rank * B)
ref_loss.backward()
torch.cuda.synchronize()
BF.check_tc_status()
assert abs(loss.item() - ref_loss.item()) < 1e-6 and torch.equal(e.grad, ref_e.grad)
if rank == 0:
    print(f"multi-GPU check OK on {world} ranks: symmetric-memory gather used = {used_symm}, loss {loss.item():.6f}")
dist.barrier()
dist.destroy_process_group()
