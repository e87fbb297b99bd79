"""Where does the HOST spend its time enqueuing one training step (cfg2, B=256)?  cProfile over 10 resident steps, top
functions by own time; plus the GPU time of the same steps (CUDA events) to see which side paces the step.
    python profiles/host_profile.py [n_steps]"""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import brainmagick_b200 as bb  # noqa: E402
from brainmagick_b200 import synthetic  # noqa: E402

n_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
cfg = bench.CONFIGS["cfg2"]
B, C, T, F, S = 256, cfg["C"], cfg["T"], cfg["F"], cfg["S"]
torch.manual_seed(2036)
model = bb.SimpleConv(in_channels=dict(meg=C), out_channels=F, n_subjects=S,
                      **{k: (dict(v) if isinstance(v, dict) else v) for k, v in bench.CLIP_CONV.items()}).to(dev).train()
clip = bb.ClipLoss(uniform_batches=True).to(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=3e-4, fused=True)
positions = synthetic.normalised_positions(S, C, n_valid=cfg["n_valid"], seed=7)
meg, feats, subj = (t.to(dev) for t in bench.make_host_batch(cfg, B, 1))
subj_l = subj.tolist()
recs = [synthetic.SyntheticRecording(s, positions[s]) for s in range(S)]
mask = torch.ones(B, 1, T, dtype=torch.bool, device=dev)


def step():
    opt.zero_grad(set_to_none=True)
    batch = synthetic.SyntheticBatch(meg, subj, [recs[s] for s in subj_l])
    est = model(dict(meg=meg), batch)
    loss = clip(est, feats, mask)
    loss.backward()
    opt.step()
    return loss


for _ in range(5):
    step()
torch.cuda.synchronize()
# host time per step with an EMPTY launch queue ahead (sync before every step) vs GPU time of the step
host, gpu = [], []
for _ in range(n_steps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    step()
    e1.record()
    host.append((time.perf_counter() - t0) * 1e3)
    torch.cuda.synchronize()
    gpu.append(e0.elapsed_time(e1))
print(f"host enqueue ms/step: mean {sum(host) / len(host):.2f} min {min(host):.2f};  GPU ms/step (launch-paced when the host is "
      f"slower): mean {sum(gpu) / len(gpu):.2f} min {min(gpu):.2f}")
pr = cProfile.Profile()
pr.enable()
for _ in range(n_steps):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(45)
