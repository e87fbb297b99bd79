"""TEST / BENCH INFRASTRUCTURE ONLY -- the training step of bm/solver.py:297,343-387 around the VERBATIM reference modules
(`bm.models.simpleconv.SimpleConv`, `bm.losses.ClipLoss`, loaded unmodified by oracle/ref_loader.py), on CPU tensors:

    estimate = model(dict(meg=meg), batch)                       solver.py:297
    loss = clip(estimate, features, features_mask)               solver.py:373
    loss += mod.training_penalty for every module that has one   solver.py:376-378
    optimizer.zero_grad(); loss.backward(); optimizer.step()     solver.py:384-387  (Adam lr 3e-4, betas (0.9, 0.999), train.py:119)

The Solver itself cannot be imported here (flashy, dora, hydra, julius, mne absent), hence this ~30-line harness; the
batch / recording objects are the minimal stand-ins of ref_loader (what PositionGetter and SimpleConv.forward touch).
What `bench.py --impl reference` and its `cpu_baseline` leg time (kind "reference").  Never imported by the product.
"""
from __future__ import annotations

import typing as tp

import torch

from . import ref_loader


class VerbatimTrainer:
    def __init__(self, in_channels: int, out_channels: int, n_subjects: int, n_valid: tp.Sequence[int] = (),
                 seed: int = 2036, lr: float = 3e-4):
        common, simpleconv, losses = ref_loader.load_reference()
        torch.manual_seed(seed)                                          # conf/config.yaml:33
        self.model = simpleconv.SimpleConv(in_channels=dict(meg=in_channels), out_channels=out_channels,
                                           n_subjects=n_subjects, **ref_loader.clip_conv_kwargs())
        self.clip = losses.ClipLoss()
        self.model.train()
        self.clip.train()
        self.opt = torch.optim.Adam(list(self.model.parameters()) + list(self.clip.parameters()), lr=lr, betas=(0.9, 0.999))
        self.recordings = [ref_loader.FakeRecording(s, in_channels, n_valid[s % len(n_valid)] if n_valid else None, seed=seed)
                           for s in range(n_subjects)]

    def step(self, meg: torch.Tensor, features: torch.Tensor, subject_index: torch.Tensor) -> float:
        batch = ref_loader.FakeBatch(meg, subject_index, [self.recordings[int(s)] for s in subject_index])
        mask = torch.ones(len(meg), 1, meg.shape[-1], dtype=torch.bool)
        estimate = self.model(dict(meg=meg), batch)
        loss = self.clip(estimate, features, mask)
        for mod in self.model.modules():
            if hasattr(mod, "training_penalty"):
                loss += mod.training_penalty
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return float(loss.detach())
