"""Diagnostic for the full-width accuracy task: loss trajectories of (a) the CUDA drop-in, (b) the oracle's PyTorch ops run on
the GPU (cuDNN / cuBLAS fp32, TF32 off) and (c) the stored CPU-oracle trajectory, same task / schedule / initial state."""
import json, os, sys
import torch
sys.path.insert(0, os.getcwd())
import brainmagick_b200 as bb
from brainmagick_b200 import synthetic, functional as BF
from oracle import make_accuracy_golden as mg, bm_oracle

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
gold = json.load(open("tests/golden/accuracy_full_width.json"))
cfg, task, sched, p0 = mg.build(gold["spec"])
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else len(sched)
d = task["train"]
# (b) oracle ops on the GPU
p0g = {k: v.cuda() for k, v in p0.items()}
tr = bm_oracle.CpuTrainer(cfg, p0g, lr=gold["spec"]["lr"])
pos_g = task["positions"].cuda()
lb = []
for idx, ban in sched[:nsteps]:
    lb.append(tr.step(d["meg"][idx].cuda(), pos_g, d["subj"][idx].cuda(), d["subj"][idx].cuda(), d["feats"][idx].cuda(), ban.cuda()))
# (a) CUDA drop-in
model = bb.SimpleConv(in_channels=dict(meg=cfg.in_channels), out_channels=cfg.out_channels, hidden=dict(meg=cfg.hidden), depth=cfg.depth,
    dilation_period=5, kernel_size=3, skip=True, subject_layers=True, subject_dim=0, complex_out=True, glu=2, glu_context=1, merger=True,
    initial_linear=cfg.initial_linear, merger_channels=cfg.merger_channels, gelu=True, batch_norm=True, merger_pos_dim=cfg.merger_pos_dim,
    n_subjects=cfg.n_subjects)
model.load_state_dict(p0); model = model.cuda().train()
clip = bb.ClipLoss().cuda().train()
opt = torch.optim.Adam(model.parameters(), lr=gold["spec"]["lr"], betas=(0.9, 0.999))
recs = [synthetic.SyntheticRecording(s, task["positions"][s]) for s in range(cfg.n_subjects)]
mask = torch.ones(32, 1, 360, dtype=torch.bool, device="cuda")
la = []
for idx, ban in sched[:nsteps]:
    meg, feats, subj = d["meg"][idx].cuda(), d["feats"][idx].cuda(), d["subj"][idx].cuda()
    batch = synthetic.SyntheticBatch(meg, subj, [recs[int(s)] for s in d["subj"][idx]])
    model.merger.ban_centre_override = ban
    opt.zero_grad(set_to_none=True)
    loss = clip(model(dict(meg=meg), batch), feats, mask)
    loss.backward()
    opt.step()
    la.append(float(loss.detach()))
BF.check_tc_status()
for i in list(range(0, 12)) + list(range(15, nsteps, 16)):
    print(f"step {i:3d}: cuda {la[i]:8.4f}   oracle-ops-on-gpu {lb[i]:8.4f}   cpu oracle (golden) {gold['losses'][i]:8.4f}")
