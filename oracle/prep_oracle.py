"""TEST INFRASTRUCTURE ONLY -- CPU restatement of brainmagick's pre-model batch preparation.

What the reference does to every batch between the DataLoader and `model(inputs, batch)`:
    bm/norm.py:81-85, 109-113   RobustScaler / StandardScaler .transform / .inverse_transform:
                                (X - center) / scale   resp.   X * scale + center     (fp32, per channel)
    bm/norm.py:239-275          BatchScaler._transform: meg scaled by the scaler of ITS recording
                                (python loop over samples), features scaled slice by slice
    bm/norm.py:325-341          ScaleReject.__call__: optional clamp to +-limit, reject samples whose
                                max |meg| still exceeds limit (or whose features are empty), return batch[keep], keep
    bm/solver.py:262-274        offset crop: meg[..., off:], features[..., :-off], mask[..., :-off]

Pinned against the verbatim `bm/norm.py` (loaded by `oracle/ref_loader.load_reference_norm`) through
`tests/golden/prep_small.npz` (made by `oracle/make_golden.py`).  Arithmetic is two fp32 roundings per element
(subtract, divide), so the CUDA path is compared BIT-EXACT against this file.

Only `tests/`, `__graft_entry__.smoke()` and the CPU legs of `bench.py` may import this module.
"""
from __future__ import annotations

import typing as tp

import numpy as np


def scale_rows(x: np.ndarray, center: np.ndarray, scale: np.ndarray, inverse: bool = False) -> np.ndarray:
    """x [B, C, T]; center/scale [B, C] (already looked up per sample) or [C].  norm.py:81-85 / 109-113."""
    x = np.asarray(x, dtype=np.float32)
    c = np.asarray(center, dtype=np.float32)[..., None]
    s = np.asarray(scale, dtype=np.float32)[..., None]
    if inverse:
        return (x * s).astype(np.float32) + c
    return ((x - c).astype(np.float32) / s).astype(np.float32)


def batch_transform(meg, recording_index, meg_center, meg_scale, features, feat_center, feat_scale,
                    inverse: bool = False):
    """BatchScaler._transform (norm.py:239-275).  `meg_center/meg_scale`: {recording_index: [C]} tables;
    `feat_center/feat_scale`: [F] vectors (a scalar StandardScaler or a NoOp scaler broadcast over its slice)."""
    rec = [int(r) for r in recording_index]
    c = np.stack([meg_center[r] for r in rec]).astype(np.float32)
    s = np.stack([meg_scale[r] for r in rec]).astype(np.float32)
    return scale_rows(meg, c, s, inverse), scale_rows(features, feat_center, feat_scale, inverse)


def scale_reject(meg, recording_index, meg_center, meg_scale, features, features_mask, feat_center, feat_scale,
                 limit: float = 16.0, clip: bool = False, exclude_empty_features: bool = False):
    """ScaleReject.__call__ (norm.py:325-341) -> (meg[keep], features[keep], features_mask[keep], keep)."""
    meg, features = batch_transform(meg, recording_index, meg_center, meg_scale, features, feat_center, feat_scale)
    if clip:
        meg = np.clip(meg, -np.float32(limit), np.float32(limit))
    peak = np.abs(meg).reshape(len(meg), -1).max(-1)
    reject = peak > limit
    if exclude_empty_features:
        reject |= features_mask.reshape(len(features_mask), -1).sum(-1) == 0
    keep = ~reject
    return meg[keep], features[keep], features_mask[keep], keep


def offset_crop(meg, features, features_mask, offset_samples: int):
    """solver.py:262-274 (`task.offset_meg_ms`): the brain signal is moved to the past by `offset_samples`."""
    if not offset_samples:
        return meg, features, features_mask
    return meg[..., offset_samples:], features[..., :-offset_samples], features_mask[..., :-offset_samples]


def prepare(meg, recording_index, meg_center, meg_scale, features, features_mask, feat_center, feat_scale,
            limit=20.0, clip=True, exclude_empty_features=False, offset_samples=0) -> tp.Dict[str, np.ndarray]:
    """solver.py:243-274 up to the model call: scale + reject, then the offset crop."""
    meg, features, mask, keep = scale_reject(meg, recording_index, meg_center, meg_scale, features, features_mask,
                                             feat_center, feat_scale, limit, clip, exclude_empty_features)
    meg, features, mask = offset_crop(meg, features, mask, offset_samples)
    return dict(meg=np.ascontiguousarray(meg), features=np.ascontiguousarray(features),
                features_mask=np.ascontiguousarray(mask), keep=keep)
