"""Batch preparation on the GPU: drop-in `BatchScaler` / `ScaleReject` (reference: bm/norm.py).

The reference scales every sample with the RobustScaler of its recording in a Python loop over the batch
(norm.py:255-260), then clamps and looks for out-of-range samples in further passes (norm.py:332-337), and the solver
crops the result (solver.py:262-274).  Here the per-recording constants live in one [R, C] table on the device and one
kernel pass does gather + affine + clamp + peak (+ the crop, in `ScaleReject.prepare`): 4 B read and 4 B written per
kept element.  The arithmetic is the reference's two fp32 roundings, so results are bit-identical.

Surface kept: `BatchScaler(features_builder, n_samples_per_recording, per_channel, device, n_samples_features)`, `.fit`,
`.transform`, `.inverse_transform`, `.meg_scalers`, `.feature_scalers`; `ScaleReject(scaler, limit, exclude_empty_features,
clip)`, `__call__(batch) -> (batch[keep], keep)`, `.rejection_rate`.  A batch is anything with `.meg`, `.features`,
`.features_mask`, `.recording_index`, `.replace(**fields)` and boolean-mask indexing (the reference's SegmentBatch).
`fit` is one-off setup and runs in plain torch; the per-batch path needs CUDA tensors and has no CPU fallback.
"""
from __future__ import annotations

import random
import typing as tp
from collections import OrderedDict, defaultdict

import torch

from ._lib import call, ptr, stream


def _as_nd(x: torch.Tensor) -> torch.Tensor:
    """[B, C, T] -> [B*T, C] (norm.py:23-26)."""
    return x.transpose(1, 2).reshape(-1, x.shape[1])


class Scaler:
    """Constants of one fitted scaler: `center_`, `scale_` (None = identity, the reference's NoOp scalers)."""
    center_: tp.Optional[torch.Tensor] = None
    scale_: tp.Optional[torch.Tensor] = None

    def transform(self, X):
        return X if self.scale_ is None else (X - self.center_.to(X)) / self.scale_.to(X)

    def inverse_transform(self, X):
        return X if self.scale_ is None else (X * self.scale_.to(X)) + self.center_.to(X)


class RobustScaler(Scaler):
    """Median / inter-quantile range per channel (norm.py:46-85); zero ranges (padded channels) become 1."""

    def __init__(self, lowq=0.25, highq=0.75, subsample=1., device="cpu"):
        self.lowq, self.highq, self.subsample, self.device = lowq, highq, subsample, device

    def fit(self, X: torch.Tensor):
        X = X.to(self.device)
        low, mid, high = [], [], []
        for d in range(X.shape[1]):
            col = X[:, d]
            # the uniform draw is made even when subsample == 1, so that the global RNG stream advances exactly as it
            # does in the reference (norm.py:63-64)
            col = col[torch.rand_like(col) < self.subsample].sort()[0]
            n = len(col)
            low.append(col[int(self.lowq * n)])
            mid.append(col[int(0.5 * n)])
            high.append(col[int(self.highq * n)])
        self.center_ = torch.stack(mid).cpu()
        scale = (torch.stack(high) - torch.stack(low)).cpu()
        assert (scale != 0).any()
        scale[scale == 0] = 1           # zero-padded channels (norm.py:73-78)
        self.scale_ = scale
        return self


class StandardScaler(Scaler):
    """Mean / std over the masked entries, per channel or global (norm.py:88-113)."""

    def __init__(self, per_channel):
        self.per_channel = per_channel

    def fit(self, X, mask):
        kept = X[mask.expand_as(X)]
        if self.per_channel:
            kept = kept.reshape(-1, X.shape[1])
        self.center_, self.scale_ = kept.mean(dim=0), kept.std(dim=0)
        return self


class NoOpScaler(Scaler):
    def fit(self, X, mask):
        return self


class NoOpCategoryCountScaler(NoOpScaler):
    """Identity that also counts the categories (norm.py:124-147), for get_categorical_feature_weights."""

    def __init__(self, cardinality):
        self.cardinality = cardinality

    def fit(self, X, mask):
        assert bool((X == X.int()).all()) and X.min().item() == 0 and X.max().item() < self.cardinality
        self.categories_count_ = torch.histc(X[mask], bins=self.cardinality, min=0, max=self.cardinality - 1)
        return self


class BatchScaler:
    def __init__(self, features_builder, n_samples_per_recording=200, per_channel=False, device: str = 'cpu',
                 n_samples_features: tp.Optional[int] = None):
        self.n_samples_per_recording = n_samples_per_recording
        self.n_samples_features = n_samples_features
        self.device = device
        self.meg_scalers: tp.Dict[int, Scaler] = {}
        self.features_builder = features_builder
        self.feature_scalers: tp.Dict[str, Scaler] = OrderedDict()
        for name, feature in features_builder.items():         # norm.py:160-174
            if feature.normalizable:
                self.feature_scalers[name] = StandardScaler(per_channel)
            elif feature.categorical:
                self.feature_scalers[name] = NoOpCategoryCountScaler(feature.cardinality)
            else:
                self.feature_scalers[name] = NoOpScaler()
        self._tables: tp.Dict[tp.Any, tp.Any] = {}

    @classmethod
    def from_reference(cls, scaler) -> "BatchScaler":
        """Adopts the constants of a fitted reference `bm.norm.BatchScaler` (or of a checkpointed one)."""
        new = cls.__new__(cls)
        new.n_samples_per_recording = scaler.n_samples_per_recording
        new.n_samples_features = scaler.n_samples_features
        new.device = scaler.device
        new.features_builder = scaler.features_builder
        new.meg_scalers, new.feature_scalers, new._tables = {}, OrderedDict(), {}
        for key, src in scaler.meg_scalers.items():
            s = Scaler()
            s.center_, s.scale_ = src.center_.clone(), src.scale_.clone()
            new.meg_scalers[key] = s
        for name, src in scaler.feature_scalers.items():
            s = Scaler()
            if hasattr(src, "scale_"):
                s.center_, s.scale_ = src.center_.clone(), src.scale_.clone()
            new.feature_scalers[name] = s
        return new

    # -- one-off setup (norm.py:176-237) ------------------------------------------------------------------
    def fit(self, loaders: tp.Sequence[tp.Iterable]):
        meg_of = defaultdict(list)
        all_features, all_mask = [], []
        for loader in loaders:
            remaining = self.n_samples_per_recording
            for batch in loader:
                remaining -= len(batch.meg)
                index = batch.recording_index[0].item()
                assert (batch.recording_index == index).all()
                meg_of[index].append(batch.meg)
                all_features.append(batch.features)
                all_mask.append(batch.features_mask)
                if remaining <= 0:
                    break
        if self.n_samples_features is not None:
            order = list(range(len(all_features)))
            random.Random(1234).shuffle(order)                  # norm.py:209-211
            remaining, cut = self.n_samples_features, len(order)
            for pos, idx in enumerate(order):
                remaining -= len(all_features[idx])
                if remaining <= 0:
                    cut = pos + 1
                    break
            all_features = [all_features[idx] for idx in order[:cut]]
            all_mask = [all_mask[idx] for idx in order[:cut]]
        features, mask = _as_nd(torch.cat(all_features)), _as_nd(torch.cat(all_mask))
        for index, chunks in meg_of.items():
            assert index not in self.meg_scalers
            self.meg_scalers[index] = RobustScaler(device=self.device).fit(_as_nd(torch.cat(chunks)))
        for name, fscaler in self.feature_scalers.items():
            fscaler.fit(features[:, self.features_builder.get_slice(name)], mask)
            if isinstance(fscaler, StandardScaler):
                assert (fscaler.scale_ > 0).all(), f"feature {name} is constant and cannot be normalized"
        self._tables.clear()

    # -- device tables --------------------------------------------------------------------------------------
    def tables(self, device, n_channels: int):
        """(center [R, C], scale [R, C], slot_of [max recording_index + 3] int32 (see `slots`), feat_center [1, F], feat_scale [1, F])."""
        key = (str(device), n_channels, len(self.meg_scalers))
        hit = self._tables.get(key)
        if hit is not None:
            return hit
        assert self.meg_scalers, "BatchScaler is not fitted"
        ids = sorted(self.meg_scalers)
        center = torch.stack([self.meg_scalers[i].center_.float() for i in ids])
        scale = torch.stack([self.meg_scalers[i].scale_.float() for i in ids])
        assert center.shape[1] == n_channels, f"scalers were fitted on {center.shape[1]} channels, got {n_channels}"
        # slot_of[1 + recording_index]; one sentinel (-1) entry at each end absorbs every out-of-range index after a clamp
        slot_of = torch.full((max(ids) + 3,), -1, dtype=torch.int32)
        slot_of[torch.tensor(ids) + 1] = torch.arange(len(ids), dtype=torch.int32)
        dim = self.features_builder.dimension
        fc, fs = torch.zeros(1, dim), torch.ones(1, dim)
        for name, fscaler in self.feature_scalers.items():
            if fscaler.scale_ is not None:
                sl = self.features_builder.get_slice(name)
                fc[0, sl], fs[0, sl] = fscaler.center_.float(), fscaler.scale_.float()
        out = tuple(t.to(device).contiguous() for t in (center, scale, slot_of, fc, fs))
        self._tables[key] = out
        return out

    def slots(self, recording_index: torch.Tensor, slot_of: torch.Tensor) -> torch.Tensor:
        """Table row of each sample; -1 for a recording the scaler was not fitted on (the kernel then poisons the sample
        with NaN -- the reference raises KeyError there -- without a device-to-host check in the per-batch path)."""
        idx = recording_index.to(slot_of.device).long()
        return slot_of[(idx + 1).clamp_(0, len(slot_of) - 1)].contiguous()

    # -- per batch --------------------------------------------------------------------------------------------
    def _apply(self, batch, inverse: bool, limit=0.0, clip=False, t0=0, crop=0, want_peak=False):
        meg, features = batch.meg, batch.features
        if features.shape[1] != self.features_builder.dimension:
            raise ValueError(f"Invalid channel dim {features.shape[1]} for features, "
                             f"expected {self.features_builder.dimension}")
        B, C, T = meg.shape
        center, scale, slot_of, fc, fs = self.tables(meg.device, C)
        slot = self.slots(batch.recording_index, slot_of)
        meg = meg.contiguous().float()
        features = features.contiguous().float()
        out_meg = torch.empty(B, C, T - crop, device=meg.device, dtype=torch.float32)
        peak = torch.empty(B, device=meg.device, dtype=torch.int32) if want_peak else None
        call("bm_scale_clamp_crop", ptr(meg), ptr(slot), ptr(center), ptr(scale), B, C, T, t0, T - crop, float(limit),
             int(clip), int(inverse), ptr(out_meg), ptr(peak), stream())
        Tf = features.shape[2]
        out_feat = torch.empty(B, features.shape[1], Tf - crop, device=meg.device, dtype=torch.float32)
        call("bm_scale_clamp_crop", ptr(features), None, ptr(fc), ptr(fs), B, features.shape[1], Tf, 0, Tf - crop, 0.0,
             0, int(inverse), ptr(out_feat), None, stream())
        return out_meg, out_feat, peak

    def _transform(self, batch, inverse_transform: bool):
        meg, features, _ = self._apply(batch, inverse_transform)
        return batch.replace(meg=meg, features=features)

    def transform(self, batch):
        return self._transform(batch, inverse_transform=False)

    def inverse_transform(self, batch):
        return self._transform(batch, inverse_transform=True)

    def inverse_transform_feature(self, feature_name, feature_data):
        s = self.feature_scalers[feature_name]
        return feature_data if s.scale_ is None else feature_data * s.scale_.to(feature_data).reshape(-1, 1) \
            + s.center_.to(feature_data).reshape(-1, 1)

    def get_categorical_feature_weights(self, feature_name) -> torch.Tensor:
        """norm.py:289-308: inverse-sqrt-frequency class weights with E[w] = 1."""
        scaler = self.feature_scalers[feature_name]
        assert isinstance(scaler, NoOpCategoryCountScaler)
        probs = scaler.categories_count_ / scaler.categories_count_.sum()
        weights = 1 / torch.sqrt(probs)
        weights[probs == 0] = 0.
        return weights / torch.sqrt(probs).sum()


class ScaleReject:
    """Rescale, then clamp (`clip`) or reject the samples still beyond `limit` (norm.py:311-345)."""

    def __init__(self, scaler: BatchScaler, limit=16, exclude_empty_features=False, clip=False):
        self.scaler = scaler
        self.limit = limit
        self.clip = clip
        self.exclude_empty_features = exclude_empty_features
        self._rejection_count = 0
        self._count = 0

    def _scaled(self, batch, offset_samples: int):
        can_reject = (not self.clip) or self.exclude_empty_features
        meg, features, peak = self.scaler._apply(batch, False, self.limit, self.clip, t0=offset_samples,
                                                 crop=offset_samples, want_peak=can_reject)
        B = meg.shape[0]
        self._count += B
        if not can_reject:
            # clamped values cannot exceed the limit (norm.py:332-335): nothing to test, and no host synchronisation
            return meg, features, torch.ones(B, device=meg.device, dtype=torch.bool), None, B
        keep = torch.empty(B, device=meg.device, dtype=torch.bool)
        rows = torch.empty(B, device=meg.device, dtype=torch.int32)
        n_keep = torch.empty(1, device=meg.device, dtype=torch.int32)
        mask = batch.features_mask.contiguous() if self.exclude_empty_features else None
        assert mask is None or mask.dtype == torch.bool
        call("bm_reject_compact", ptr(peak), ptr(mask), 0 if mask is None else mask[0].numel(), float(self.limit), B,
             ptr(keep), ptr(rows), ptr(n_keep), stream())
        kept = int(n_keep.item())                      # the reference synchronises here too (norm.py:338)
        self._rejection_count += B - kept
        return meg, features, keep, rows, kept

    def __call__(self, batch):
        meg, features, keep, _, kept = self._scaled(batch, 0)
        batch = batch.replace(meg=meg, features=features)
        return (batch if kept == len(keep) else batch[keep]), keep

    def prepare(self, batch, offset_samples: int = 0):
        """`Solver._process_batch` from the scaler to the model input in one pass (solver.py:245-274): returns
        (meg[keep][..., off:], features[keep][..., :-off], features_mask[keep][..., :-off], keep)."""
        meg, features, keep, rows, kept = self._scaled(batch, int(offset_samples))
        mask = batch.features_mask
        if offset_samples:
            mask = mask[..., :-offset_samples]
        if kept < len(keep):
            rows = rows[:kept].contiguous()
            packed = []
            for t in (meg, features):
                out = torch.empty((kept,) + tuple(t.shape[1:]), device=t.device, dtype=t.dtype)
                if kept:
                    call("bm_gather_rows", ptr(t), ptr(rows), kept, t[0].numel(), ptr(out), stream())
                packed.append(out)
            meg, features = packed
            mask = mask[keep]
        return meg, features, mask, keep

    @property
    def rejection_rate(self):
        return self._rejection_count / max(self._count, 1)
