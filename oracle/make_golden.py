"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the VERBATIM reference modules
(bm/models/simpleconv.py, bm/models/common.py, bm/losses.py loaded by `oracle/ref_loader.py`) on small seeded
inputs, in this build container (the reference tree does not travel to the GPU box; the fixtures do).

    python oracle/make_golden.py            # rewrites tests/golden/

Each fixture holds: the configuration, the reference state_dict (`p.<key>`), the inputs, and the reference's
outputs: estimate, scores, loss, probabilities, every parameter gradient (`g.<key>`), and BN running stats after
the step (`bn.<key>`).  Everything fp32, produced by the reference's own code path:
    estimate = SimpleConv(...)(dict(meg=meg), batch)            simpleconv.py:198
    loss     = ClipLoss()(estimate, candidates, mask)           losses.py:104
    loss.backward()                                             solver.py:385
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # name: dict(B, Bc, C, T, F, S, hidden, depth, MC, IL, P, train, n_valid per recording)
    "train_small": dict(B=6, Bc=6, C=10, T=24, F=8, S=3, hidden=16, depth=10, MC=12, IL=12, P=32,
                        train=True, n_valid=None, seed=1),
    "eval_small": dict(B=5, Bc=5, C=10, T=24, F=8, S=3, hidden=16, depth=10, MC=12, IL=12, P=32,
                       train=False, n_valid=None, seed=2),
    "train_padded_negs": dict(B=8, Bc=11, C=14, T=37, F=12, S=4, hidden=32, depth=10, MC=20, IL=24, P=72,
                              train=True, n_valid=[14, 9, 6, 14], seed=3),
    "train_depth4": dict(B=4, Bc=4, C=7, T=50, F=5, S=2, hidden=24, depth=4, MC=8, IL=8, P=8,
                         train=True, n_valid=None, seed=4),
}


def run_case(name, c):
    common, simpleconv, losses = ref_loader.load_reference()
    torch.manual_seed(100 + c["seed"])
    kw = ref_loader.clip_conv_kwargs(hidden=c["hidden"], depth=c["depth"], merger_channels=c["MC"],
                                     initial_linear=c["IL"], merger_pos_dim=c["P"])
    model = simpleconv.SimpleConv(in_channels=dict(meg=c["C"]), out_channels=c["F"], n_subjects=c["S"], **kw)
    clip = losses.ClipLoss()
    # make BN affine / running stats non-trivial so that parity actually exercises them
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.add_(0.1 * torch.randn_like(m.weight))
                m.bias.add_(0.1 * torch.randn_like(m.bias))
                m.running_mean.add_(0.1 * torch.randn_like(m.running_mean))
                m.running_var.mul_(1 + 0.2 * torch.rand_like(m.running_var))
    state = {k: v.detach().clone() for k, v in model.state_dict().items()}

    B, Bc, C, T, F, S = c["B"], c["Bc"], c["C"], c["T"], c["F"], c["S"]
    meg = torch.randn(B, C, T).clamp_(-20, 20)
    cand = torch.randn(Bc, F, T)
    subj = torch.randint(0, S, (B,))
    nv = c["n_valid"] or [C] * S
    recs = [ref_loader.FakeRecording(s, C, nv[s], seed=c["seed"]) for s in range(S)]
    for b in range(B):
        meg[b, nv[int(subj[b])]:] = 0
    batch = ref_loader.FakeBatch(meg, subj, [recs[int(s)] for s in subj])
    pos_rec = torch.stack([model.merger.position_getter.get_recording_layout(r).clone() for r in recs])
    full = torch.full((S, C, 2), common.PositionGetter.INVALID)
    full[:, :pos_rec.shape[1]] = pos_rec

    model.train(c["train"])
    clip.train(c["train"])
    torch.manual_seed(9000 + c["seed"])
    ban = torch.rand(2)                       # the ONE rand(2) the forward will draw (common.py:343)
    torch.manual_seed(9000 + c["seed"])
    est = model(dict(meg=meg.clone()), batch)
    mask = torch.ones(B, 1, T, dtype=torch.bool)
    loss = clip(est, cand, mask)
    scores = clip.get_scores(est, cand)
    probs = clip.get_probabilities(est, cand)
    loss.backward()

    out = dict(cfg=np.array([B, Bc, C, T, F, S, c["hidden"], c["depth"], c["MC"], c["IL"], c["P"],
                             int(c["train"])], dtype=np.int64),
               meg=meg.numpy(), candidates=cand.numpy(), subject_index=subj.numpy(),
               rec_positions=full.numpy(), rec_of_sample=subj.numpy(), ban_centre=ban.numpy(),
               estimate=est.detach().numpy(), scores=scores.detach().numpy(), loss=loss.detach().numpy(),
               probs=probs.detach().numpy())
    for k, v in state.items():
        out["p." + k] = v.numpy()
    for k, v in model.named_parameters():
        out["g." + k] = v.grad.numpy() if v.grad is not None else np.zeros(0, np.float32)
    for k, v in model.state_dict().items():
        if "running" in k or "num_batches" in k:
            out["bn." + k] = v.numpy()
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(f"{name}: loss={float(loss):.6f} est.std={float(est.std()):.4f} "
          f"size={os.path.getsize(os.path.join(OUT, name + '.npz')) / 1024:.0f} KiB")


if __name__ == "__main__":
    torch.set_num_threads(1)
    for name, c in CASES.items():
        run_case(name, c)
