"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the SimpleConv variants of the paper's ablation table
(grids/nmi/ablation_final.py:42-52 over bm/models/simpleconv.py:85-196, 198-249):

    merger=False            raw sensors go straight to initial_linear               simpleconv.py:106-110, 210-211
    initial_linear=0        no 1x1 conv after the merger                            simpleconv.py:112-120, 213-214
    subject_layers=False    no per-subject linear layer                              simpleconv.py:122-128, 216-217
    subject_dim=64          ScaledEmbedding of the subject appended as channels      simpleconv.py:148-151, 231-233
    glu=0, skip=False       ConvSequence without GLU blocks / residuals              common.py:79-151
    gelu=False              ReLU in the ConvSequence AND in the head                 simpleconv.py:85-90, 187
    complex_out=False       no head: the last conv maps to out_channels, no BN/act   simpleconv.py:190-193
    subsample_meg_channels  only n sensors (drawn with random.Random(1234)) are kept  simpleconv.py:97-102, 200-203
This is step (a) -- the oracle -- of that row: it is pinned against the verbatim reference (`tests/golden/ablation_*.npz`,
made by `oracle/make_golden.py`) so that the CUDA variants can be built against it.  The CUDA SimpleConv still accepts only
the `clip_conv` family.

Only `tests/`, `__graft_entry__.smoke()` and the CPU legs of `bench.py` may import this module.
"""
from __future__ import annotations

import typing as tp

import torch
import torch.nn.functional as F

from . import bm_oracle, deepmel_oracle


class Variant(tp.NamedTuple):
    in_channels: int
    out_channels: int
    n_subjects: int
    hidden: int = 320
    depth: int = 10
    merger: bool = True
    merger_channels: int = 270
    merger_pos_dim: int = 2048
    merger_dropout: float = 0.2
    initial_linear: int = 270
    subject_layers: bool = True
    subject_dim: int = 0
    embedding_scale: float = 1.0
    glu: int = 2
    glu_context: int = 1
    gelu: bool = True
    skip: bool = True
    complex_out: bool = True
    kernel_size: int = 3
    dilation_period: int = 5
    subsample_meg_channels: int = 0

    def sequence_spec(self) -> deepmel_oracle.SequenceSpec:
        c = self.in_channels
        if self.merger:
            c = self.merger_channels
        if self.initial_linear:
            c = self.initial_linear
        c += self.subject_dim
        widths = [c] + [self.hidden] * self.depth
        if not self.complex_out:
            widths[-1] = self.out_channels                     # simpleconv.py:192-193
        return deepmel_oracle.SequenceSpec(
            channels=tuple(widths), kernel=self.kernel_size, dilation_growth=2, dilation_period=self.dilation_period,
            batch_norm=True, skip=self.skip, activation_on_last=self.complex_out, glu=self.glu,
            glu_context=self.glu_context, activation="gelu" if self.gelu else "lrelu", leakiness=0.0)


def forward(p: tp.Dict[str, torch.Tensor], v: Variant, meg, rec_positions, rec_of_sample, subject_index,
            training: bool, ban_centre=None, new_stats=None) -> torch.Tensor:
    """simpleconv.py:198-249 for one variant; parameters under the reference's state_dict names."""
    length = meg.shape[-1]
    x = meg
    if v.subsample_meg_channels:                                                                 # simpleconv.py:97-102,200-203
        import random
        indexes = list(range(v.in_channels))
        random.Random(1234).shuffle(indexes)
        mask = torch.zeros(1, v.in_channels, 1, dtype=x.dtype)
        mask[:, indexes[:v.subsample_meg_channels]] = 1.
        x = x * mask
    if v.merger:
        centre = ban_centre if (training and v.merger_dropout) else None
        w = bm_oracle.attention_weights(rec_positions, p["merger.heads"], centre, v.merger_dropout)
        x = torch.einsum("bct,boc->bot", x, w[rec_of_sample])                                   # common.py:358
    if v.initial_linear:
        x = F.conv1d(x, p["initial_linear.0.weight"], p["initial_linear.0.bias"])
    if v.subject_layers:
        x = torch.einsum("bct,bcd->bdt", x, p["subject_layers.weights"][subject_index])         # common.py:55-58
    if v.subject_dim:
        emb = p["subject_embedding.embedding.weight"][subject_index] * v.embedding_scale        # common.py:41-42
        x = torch.cat([x, emb[:, :, None].expand(-1, -1, length)], dim=1)
    seq = {k[len("encoders.meg."):]: t for k, t in p.items() if k.startswith("encoders.meg.")}
    stats = {} if new_stats is not None else None
    x = deepmel_oracle.conv_sequence(x, seq, v.sequence_spec(), training, new_stats=stats)
    if new_stats is not None:
        new_stats.update({"encoders.meg." + k: t for k, t in stats.items()})
    if v.complex_out:
        q = F.conv1d(x, p["final.0.weight"], p["final.0.bias"])
        q = F.gelu(q) if v.gelu else F.relu(q)                                                   # simpleconv.py:187
        x = F.conv_transpose1d(q, p["final.2.weight"], p["final.2.bias"])
    return x[:, :, :length]
