"""TEST INFRASTRUCTURE ONLY -- the learnable synthetic retrieval task of SURVEY.md 8(d) ("Accuracy eval set"):
a shared low-pass latent drives both the MEG (through per-subject sensor mixing, delayed by 150 ms = 18 samples at
120 Hz, bm/solver.py:262-274) and the stimulus features; the encoder must learn to align them.  Used to compare the
top-10 segment-retrieval accuracy (scripts/run_eval_probs.py:237-264 semantics) of the CUDA drop-in and of the CPU
oracle trained from the same initial state on the same batches."""
from __future__ import annotations

import typing as tp

import torch

from . import bm_oracle


def make_task(cfg: bm_oracle.Config, n_train: int, n_eval: int, T: int, latent: int = 8, seed: int = 0, delay: int = 18,
              noise: float = 1.0, eval_noise: tp.Optional[float] = None):
    """`eval_noise`: sensor noise of the held-out segments when it should differ from the training segments' (a harder
    evaluation keeps the top-k accuracy of a well-trained model away from 100 %)."""
    g = torch.Generator().manual_seed(seed)
    N = n_train + n_eval
    C, F, S = cfg.in_channels, cfg.out_channels, cfg.n_subjects
    z = torch.randn(N, latent, T + delay, generator=g)
    kernel = torch.hann_window(9, periodic=False)[None, None] / 4.0
    z = torch.nn.functional.conv1d(z.reshape(-1, 1, T + delay), kernel, padding=4).reshape(N, latent, T + delay)
    A_f = torch.randn(F, latent, generator=g) / latent ** 0.5
    A_m = torch.randn(S, C, latent, generator=g) / latent ** 0.5
    subj = torch.randint(0, S, (N,), generator=g)
    feats = torch.einsum("fl,nlt->nft", A_f, z[:, :, delay:]) + 0.1 * torch.randn(N, F, T, generator=g)
    level = torch.full((N, 1, 1), float(noise))
    if eval_noise is not None:
        level[n_train:] = float(eval_noise)
    meg = torch.einsum("ncl,nlt->nct", A_m[subj], z[:, :, :T]) + level * torch.randn(N, C, T, generator=g)
    meg = meg.clamp_(-20, 20)
    pos = torch.rand(S, C, 2, generator=g)
    for s in range(S):
        pos[s] = bm_oracle.normalise_layout(pos[s])
    sl = lambda a, b: dict(meg=meg[a:b], feats=feats[a:b], subj=subj[a:b])   # noqa: E731
    return dict(train=sl(0, n_train), eval=sl(n_train, N), positions=pos)


def batches(n_train: int, batch: int, epochs: int, seed: int = 1):
    """The fixed batch order and per-step spatial-dropout centres shared by both implementations."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(epochs):
        perm = torch.randperm(n_train, generator=g)
        for i in range(0, n_train - batch + 1, batch):
            out.append((perm[i:i + batch], torch.rand(2, generator=g)))
    return out


def train_oracle(cfg, params, task, schedule, lr=1e-3):
    tr = bm_oracle.CpuTrainer(cfg, params, lr=lr)
    losses = []
    d = task["train"]
    for idx, ban in schedule:
        losses.append(tr.step(d["meg"][idx], task["positions"], d["subj"][idx], d["subj"][idx], d["feats"][idx], ban))
    return {k: v.detach() for k, v in tr.p.items()}, losses


def eval_oracle(cfg, params, task, k=10):
    d = task["eval"]
    est = []
    with torch.no_grad():
        for i in range(0, len(d["meg"]), 256):
            est.append(bm_oracle.simpleconv_forward(params, cfg, d["meg"][i:i + 256], task["positions"], d["subj"][i:i + 256],
                                                    d["subj"][i:i + 256], False))
    est = torch.cat(est)
    return bm_oracle.topk_accuracy(est, d["feats"], torch.arange(len(est)), k=k), est
