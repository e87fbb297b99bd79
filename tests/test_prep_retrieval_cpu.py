"""CPU checks for the SURVEY 8(f) rows "retrieval evaluation" and "batch preparation":
the oracles against the fixtures produced by the verbatim reference (tests/golden/{prep,retrieval}_small.npz), and the
host-side logic of the product modules that needs no GPU (scaler fitting, device tables, vocabulary bookkeeping)."""
import types

import numpy as np
import pytest
import torch

from conftest import load_arrays as load_golden
from oracle import eval_oracle, prep_oracle
from brainmagick_b200 import norm as bnorm


# ------------------------------------------------------------------------------------------------------------------
# batch preparation
# ------------------------------------------------------------------------------------------------------------------
def _tables(g, tag):
    ids = [int(i) for i in g["rec_ids"]]
    center = {r: g[tag + "meg_center"][i] for i, r in enumerate(ids)}
    scale = {r: g[tag + "meg_scale"][i] for i, r in enumerate(ids)}
    return center, scale, g[tag + "feat_center"], g[tag + "feat_scale"]


@pytest.mark.parametrize("tag", ["pc0.", "pc1."])
def test_prep_oracle_transform_bit_exact(tag):
    g = load_golden("prep_small")
    center, scale, fc, fs = _tables(g, tag)
    meg, feats = prep_oracle.batch_transform(g["meg"], g["recording_index"], center, scale, g["features"], fc, fs)
    assert np.array_equal(meg, g[tag + "transform.meg"])
    assert np.array_equal(feats, g[tag + "transform.features"])
    imeg, ifeats = prep_oracle.batch_transform(meg, g["recording_index"], center, scale, feats, fc, fs, inverse=True)
    assert np.array_equal(imeg, g[tag + "inverse.meg"])
    assert np.array_equal(ifeats, g[tag + "inverse.features"])


@pytest.mark.parametrize("clip", [0, 1])
@pytest.mark.parametrize("excl", [0, 1])
def test_prep_oracle_scale_reject_bit_exact(clip, excl):
    g = load_golden("prep_small")
    for tag in ("pc0.", "pc1."):
        center, scale, fc, fs = _tables(g, tag)
        out = prep_oracle.prepare(g["meg"], g["recording_index"], center, scale, g["features"], g["features_mask"],
                                  fc, fs, limit=float(g["limit"]), clip=bool(clip), exclude_empty_features=bool(excl),
                                  offset_samples=int(g["offset"]))
        k = f"{tag}clip{clip}.excl{excl}."
        assert np.array_equal(out["keep"], g[k + "keep"])
        assert np.array_equal(out["meg"], g[k + "meg"])
        assert np.array_equal(out["features"], g[k + "features"])
        assert np.array_equal(out["features_mask"], g[k + "features_mask"])
    g = load_golden("prep_small")
    assert not g["pc0.clip0.excl0.keep"].all() and g["pc0.clip1.excl0.keep"].all()      # the fixture does reject


class _Builder(dict):
    def __init__(self):
        super().__init__(a=types.SimpleNamespace(normalizable=True, categorical=False, cardinality=0),
                         b=types.SimpleNamespace(normalizable=False, categorical=False, cardinality=0))
        self.dimension = 5

    def get_slice(self, name):
        return dict(a=slice(0, 3), b=slice(3, 5))[name]


@pytest.mark.parametrize("per_channel", [False, True])
def test_scaler_fit_matches_reference(per_channel):
    """RobustScaler / StandardScaler fitting (one-off setup, plain torch) reproduces the verbatim reference's constants."""
    g = load_golden("prep_small")
    tag = "pc1." if per_channel else "pc0."
    for i, r in enumerate(g["rec_ids"]):
        s = bnorm.RobustScaler().fit(torch.from_numpy(g[f"fit.meg.{int(r)}"]))
        assert np.array_equal(s.center_.numpy(), g[tag + "meg_center"][i])
        assert np.array_equal(s.scale_.numpy(), g[tag + "meg_scale"][i])
    feats = bnorm._as_nd(torch.from_numpy(g["fit.features"]))
    mask = bnorm._as_nd(torch.from_numpy(g["fit.features_mask"]))
    s = bnorm.StandardScaler(per_channel).fit(feats[:, :3], mask)
    assert np.array_equal(np.broadcast_to(s.center_.numpy(), (3,)), g[tag + "feat_center"][:3])
    assert np.array_equal(np.broadcast_to(s.scale_.numpy(), (3,)), g[tag + "feat_scale"][:3])


def test_robust_scaler_keeps_rng_stream():
    """The reference draws rand_like per column even at subsample=1; a drop-in must leave the RNG where it leaves it."""
    x = torch.randn(64, 5)
    torch.manual_seed(5)
    bnorm.RobustScaler().fit(x)
    after = torch.rand(3)
    torch.manual_seed(5)
    for d in range(5):
        torch.rand_like(x[:, d])
    assert torch.equal(after, torch.rand(3))


def test_batch_scaler_tables_and_slots():
    g = load_golden("prep_small")
    sc = bnorm.BatchScaler(_Builder(), per_channel=False)
    for i, r in enumerate(g["rec_ids"]):
        s = bnorm.Scaler()
        s.center_, s.scale_ = torch.from_numpy(g["pc0.meg_center"][i]), torch.from_numpy(g["pc0.meg_scale"][i])
        sc.meg_scalers[int(r)] = s
    sc.feature_scalers["a"].center_ = torch.tensor(float(g["pc0.feat_center"][0]))
    sc.feature_scalers["a"].scale_ = torch.tensor(float(g["pc0.feat_scale"][0]))
    center, scale, slot_of, fc, fs = sc.tables("cpu", 6)
    assert center.shape == (3, 6) and slot_of.tolist()[5] == 0 and slot_of.tolist()[10] == 1 and slot_of.tolist()[12] == 2
    assert slot_of.tolist()[0] == -1 and slot_of.tolist()[-1] == -1 and len(slot_of) == 14
    slots = sc.slots(torch.from_numpy(g["recording_index"]), slot_of)
    assert slots.tolist() == [0, 2, 1, 1, 0, 2, 0]
    assert np.array_equal(fc.numpy()[0], g["pc0.feat_center"]) and np.array_equal(fs.numpy()[0], g["pc0.feat_scale"])
    with pytest.raises(AssertionError):
        sc.tables("cpu", 7)


def test_prep_needs_cuda():
    sc = bnorm.BatchScaler(_Builder())
    s = bnorm.Scaler()
    s.center_, s.scale_ = torch.zeros(6), torch.ones(6)
    sc.meg_scalers[0] = s
    batch = types.SimpleNamespace(meg=torch.zeros(2, 6, 8), features=torch.zeros(2, 5, 8),
                                  features_mask=torch.ones(2, 1, 8, dtype=torch.bool),
                                  recording_index=torch.zeros(2, dtype=torch.long))
    with pytest.raises(Exception):       # CPU tensors never reach a CPU implementation: there is none
        sc.transform(batch)


# ------------------------------------------------------------------------------------------------------------------
# retrieval evaluation
# ------------------------------------------------------------------------------------------------------------------
def test_eval_oracle_probs_and_accuracy():
    g = load_golden("retrieval_small")
    preds, trues = torch.from_numpy(g["preds"]), torch.from_numpy(g["trues"])
    probs = eval_oracle.builds_probs(preds, trues, batch_size=7)
    assert np.abs(probs.numpy() - g["probs"]).max() < 1e-6
    labels, targets = torch.from_numpy(g["vocab_labels"]), torch.from_numpy(g["target_labels"])
    acc = [eval_oracle.accuracy_from_probs(probs, targets, labels, k) for k in (1, 5, 10)]
    assert np.allclose(acc, g["acc"])
    window = eval_oracle.crop_window(-0.5, 120.0, -0.45, -0.4)
    assert window == tuple(g["window"])
    pw = eval_oracle.builds_probs(preds, trues, 10, window)
    assert np.abs(pw.numpy() - g["probs_window"]).max() < 1e-6


def test_eval_oracle_wer():
    g = load_golden("retrieval_small")
    t = lambda k: torch.from_numpy(g[k])     # noqa: E731
    res = eval_oracle.wer_ranking(t("wer_estimates"), t("wer_word_hashes"), t("wer_outputs"), t("wer_negatives"),
                                  t("wer_negative_hashes"), int(g["wer_topx"]))
    assert res["wer"] == pytest.approx(float(g["wer"]))
    assert res["wer_vocab"] == pytest.approx(float(g["wer_vocab"]))
    assert res["soft_correct"] == pytest.approx(float(g["wer_soft"]), rel=1e-5)
    assert 0 < float(g["wer"]) < 1            # the fixture discriminates


def test_retrieval_window_expression():
    from brainmagick_b200 import retrieval
    args = types.SimpleNamespace(tmin=-0.5, sample_rate=120.0)
    assert retrieval._window(args, -0.45, -0.4) == (5, 11)      # float truncation, as the reference's expression
    assert retrieval._window(args, None, None) == (None, None)


def test_retrieval_vocabulary_bookkeeping():
    from brainmagick_b200 import retrieval
    shared = torch.tensor([30, 10, 30, 20, 10, 30], dtype=torch.int32)
    words = torch.tensor([10, 99, 30, 5], dtype=torch.int32)
    vocab, order, seg, own = retrieval._vocabulary(shared, words)
    assert vocab.tolist() == [10, 20, 30]
    assert order.tolist() == [1, 4, 3, 0, 2, 5] and seg.tolist() == [0, 2, 3, 6]
    assert own.tolist() == [0, 3, 2, 3]                  # 99 and 5 occur nowhere among the shared negatives -> slot V


def test_batch_scaler_slots_for_unknown_recordings():
    sc = bnorm.BatchScaler(_Builder())
    for r in (4, 9):
        s = bnorm.Scaler()
        s.center_, s.scale_ = torch.zeros(6), torch.ones(6)
        sc.meg_scalers[r] = s
    _, _, slot_of, _, _ = sc.tables("cpu", 6)
    got = sc.slots(torch.tensor([9, 4, 5, 77, -3]), slot_of)
    assert got.tolist() == [1, 0, -1, -1, -1] and got.dtype == torch.int32
