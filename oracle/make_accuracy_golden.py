"""TEST INFRASTRUCTURE ONLY -- the REFERENCE side of the accuracy-parity test at the BASELINE widths (SURVEY.md 8(d):
208 sensors, hidden 320, F = 1024, T = 360, 27 subjects), computed once in the build container (CPU, ~15 min) because the
oracle cannot train at this size in seconds on the GPU box:

    python oracle/make_accuracy_golden.py        # writes tests/golden/accuracy_full_width.json

The oracle (oracle/bm_oracle.py, pinned to the verbatim reference by tests/test_oracle_vs_reference.py) trains from a seeded
state_dict on the learnable synthetic retrieval task (oracle/accuracy_task.py) with a fixed batch order and spatial-dropout
centres, then is evaluated with scripts/run_eval_probs.py:237-264 semantics (top-k over all held-out candidates).
tests/test_gpu_zy_accuracy.py::test_top10_accuracy_parity_at_baseline_widths regenerates the same task and schedule from the same
seeds, trains the CUDA drop-in, and compares top-10 / top-1 accuracy and the loss trajectory with the numbers stored here."""
from __future__ import annotations

import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import accuracy_task as at, bm_oracle  # noqa: E402

SPEC = dict(C=208, F=1024, S=27, T=360, n_train=4096, n_eval=1024, batch=32, epochs=2, lr=1e-3, noise=3.0, eval_noise=8.0, latent=8,
            init_seed=3, task_seed=0, sched_seed=1)
OUT = os.path.join(ROOT, "tests", "golden", "accuracy_full_width.json")


def build(spec=SPEC):
    cfg = bm_oracle.Config(in_channels=spec["C"], out_channels=spec["F"], n_subjects=spec["S"])
    task = at.make_task(cfg, n_train=spec["n_train"], n_eval=spec["n_eval"], T=spec["T"], latent=spec["latent"],
                        seed=spec["task_seed"], noise=spec["noise"], eval_noise=spec.get("eval_noise"))
    sched = at.batches(spec["n_train"], spec["batch"], spec["epochs"], seed=spec["sched_seed"])
    p0 = bm_oracle.init_state_dict(cfg, seed=spec["init_seed"])
    return cfg, task, sched, p0


def main():
    t0 = time.time()
    cfg, task, sched, p0 = build()
    print(f"task built in {time.time() - t0:.0f} s; {len(sched)} steps of B={SPEC['batch']}", flush=True)
    tr = bm_oracle.CpuTrainer(cfg, p0, lr=SPEC["lr"])
    d = task["train"]
    losses = []
    for i, (idx, ban) in enumerate(sched):
        losses.append(tr.step(d["meg"][idx], task["positions"], d["subj"][idx], d["subj"][idx], d["feats"][idx], ban))
        if i % 16 == 0:
            print(f"step {i}: loss {losses[-1]:.4f}  ({time.time() - t0:.0f} s)", flush=True)
    params = {k: v.detach() for k, v in tr.p.items()}
    acc10, est = at.eval_oracle(cfg, params, task, k=10)
    acc1, _ = at.eval_oracle(cfg, params, task, k=1)
    out = dict(spec=SPEC, steps=len(sched), losses=losses, top10=acc10, top1=acc1,
               estimate_norm=float(est.double().norm()), seconds=time.time() - t0,
               what="CPU oracle (fp32 torch ops) on the learnable synthetic retrieval task at the BASELINE widths")
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != "losses"}), flush=True)


if __name__ == "__main__":
    main()
