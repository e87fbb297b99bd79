"""Synthetic stand-ins for what SimpleConv.forward reads from a `SegmentBatch` (bm/dataset.py:209-278) and a
`Recording` (bm/studies/api.py:49-268), so the drop-in can be driven without mne / the dataset pipeline:
`batch.meg`, `batch.subject_index`, `batch._recordings[i].{recording_index, recording_uid, mne_info.ch_names,
study_name()}`, `len(batch)`.  Sensor layouts come from an explicit positions table: `SyntheticInfo.find_layout()`
answers what `mne.find_layout(info)` (the only mne call on the hot path, bm/models/common.py:196) would."""
from __future__ import annotations

import types
import typing as tp

import numpy as np
import torch

INVALID = -0.1


class SyntheticInfo:
    def __init__(self, ch_names, layout_names, layout_pos):
        self.ch_names = list(ch_names)
        self.layout_names = list(layout_names)
        self.layout_pos = np.asarray(layout_pos, dtype=np.float64)

    def find_layout(self):
        """What `mne.find_layout(info)` would return: an object with .names and .pos[n, 4]."""
        return types.SimpleNamespace(names=list(self.layout_names), pos=self.layout_pos)


class SyntheticRecording:
    """positions [C,2] in [0,1] (already min-max normalised: each axis must span exactly [0,1]) with INVALID rows
    for padded / missing sensors."""

    def __init__(self, recording_index: int, positions: torch.Tensor):
        pos = np.asarray(positions, dtype=np.float32)
        names = [f"S{recording_index}_{i}" for i in range(len(pos))]
        valid = ~np.all(pos == np.float32(INVALID), axis=1)
        self.recording_index = int(recording_index)
        self.recording_uid = f"synthetic_{recording_index}"
        lay = np.zeros((int(valid.sum()), 4))
        lay[:, :2] = pos[valid].astype(np.float64)
        self.mne_info = SyntheticInfo(names, [n for n, v in zip(names, valid) if v], lay)

    def study_name(self):
        return "synthetic"


class SyntheticBatch:
    """The part of `SegmentBatch` (bm/dataset.py:209-258) the model, the scaler and the solver's batch preparation touch:
    tensor fields + the per-sample recordings list, `.to`, `.replace`, indexing by a mask / index tensor, `len`."""
    TENSORS = ("meg", "features", "features_mask", "subject_index", "recording_index")

    def __init__(self, meg: torch.Tensor, subject_index: torch.Tensor, recordings: tp.Sequence[SyntheticRecording],
                 features: tp.Optional[torch.Tensor] = None, features_mask: tp.Optional[torch.Tensor] = None,
                 recording_index: tp.Optional[torch.Tensor] = None):
        self.meg = meg
        self.subject_index = subject_index
        self.features = features
        self.features_mask = features_mask
        self._recording_index = recording_index
        self._recordings = list(recordings)

    @property
    def recording_index(self) -> tp.Optional[torch.Tensor]:
        """int64 [B]; derived from the recordings on first use (the training step itself never reads it)."""
        if self._recording_index is None and self._recordings:
            self._recording_index = torch.tensor([r.recording_index for r in self._recordings], dtype=torch.long)
        return self._recording_index

    def _rebuild(self, **fields):
        kw = {name: getattr(self, name) for name in self.TENSORS}
        kw["recordings"] = self._recordings
        kw.update(fields)
        return SyntheticBatch(**kw)

    def to(self, device) -> "SyntheticBatch":
        return self._rebuild(**{n: getattr(self, n).to(device) for n in self.TENSORS if getattr(self, n) is not None})

    def replace(self, **fields) -> "SyntheticBatch":
        if "_recordings" in fields:
            fields["recordings"] = fields.pop("_recordings")
        return self._rebuild(**fields)

    def __getitem__(self, index) -> "SyntheticBatch":
        rows = torch.arange(len(self), device=self.meg.device)[index].tolist()
        picked = {n: getattr(self, n)[index] for n in self.TENSORS if getattr(self, n) is not None}
        return self._rebuild(recordings=[self._recordings[i] for i in rows] if self._recordings else [], **picked)

    def __len__(self):
        return len(self.meg)


def normalised_positions(n_rec: int, n_channels: int, n_valid: tp.Sequence[int] = (), seed: int = 0) -> torch.Tensor:
    """U[0,1]^2 sensor positions per recording whose valid rows span exactly [0,1] on both axes."""
    g = torch.Generator().manual_seed(seed)
    pos = torch.rand(n_rec, n_channels, 2, generator=g)
    for r in range(n_rec):
        nv = n_valid[r % len(n_valid)] if len(n_valid) else n_channels
        p = pos[r, :nv]
        p = (p - p.min(0).values) / (p.max(0).values - p.min(0).values)
        pos[r, :nv] = p
        pos[r, nv:] = INVALID
    return pos


def make_batch(meg, subject_index, rec_positions, rec_of_sample, features=None) -> SyntheticBatch:
    recs = [SyntheticRecording(r, rec_positions[r]) for r in range(len(rec_positions))]
    return SyntheticBatch(meg, subject_index, [recs[int(r)] for r in rec_of_sample], features)
