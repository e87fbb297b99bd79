"""Size-independent properties of the retrieval evaluation at the BASELINE feature size (F=1024, T=360: 368 640 values
per candidate), where the CPU oracle would take minutes: a segment must retrieve itself, probabilities are a distribution,
the selection is ordered.  (Batch preparation and DeepMel are checked at full size in their own files.)"""
import pytest
import torch

from brainmagick_b200 import functional as BF
from brainmagick_b200 import retrieval
from brainmagick_b200.losses import ClipLoss

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_self_retrieval_at_baseline_feature_size():
    gen = torch.Generator(device=DEV).manual_seed(5)
    M, N, F, T, k = 384, 256, 1024, 360, 10
    trues = torch.randn(M, F, T, device=DEV, generator=gen)
    preds = trues[:N]                                    # every query IS one of the candidates
    labels = torch.arange(M, device=DEV, dtype=torch.int64) * 11 - 7
    clip = ClipLoss().eval()
    bank = retrieval.CandidateBank(clip, trues)
    assert bank.n == M and bank.n_pad == 512
    scores = bank.scores(clip, preds)
    BF.check_tc_status()
    r = retrieval._topk(scores, M, k, labels, labels[:N], want_soft=True)
    # Cauchy-Schwarz: <c_b, c_o>/||c_o|| <= ||c_b||, with equality only for the segment itself
    assert torch.equal(r["top_idx"][:, 0], torch.arange(N, device=DEV))
    assert (r["hit"] == 0).all()
    assert (r["top_prob"][:, 0] > 0.999).all()
    assert (r["top_prob"][:, :-1] >= r["top_prob"][:, 1:]).all()
    assert (r["top_prob"].sum(1) <= 1 + 1e-5).all()
    assert (r["soft"] - r["top_prob"][:, 0]).abs().max().item() < 1e-6      # distinct labels: soft == p(own column)
    # the diagonal score is the candidate's own norm
    own = scores[torch.arange(N), torch.arange(N)]
    assert torch.allclose(own, trues[:N].reshape(N, -1).norm(dim=1), rtol=1e-4)
    acc = retrieval.retrieval_accuracy(clip, preds, None, labels[:N], labels, topk=(1, 5), batch_size=128, bank=bank)
    assert acc == {1: 1.0, 5: 1.0}


def test_prep_nan_and_unknown_recording_semantics():
    """Tensor.clamp_ propagates NaN (norm.py:333) and a recording without a fitted scaler is an error in the reference
    (KeyError, norm.py:256): here such a sample comes out all-NaN, which Solver._process_batch's finiteness assert catches."""
    import types
    from brainmagick_b200 import norm as bnorm, synthetic

    class Builder(dict):
        dimension = 2

        def __init__(self):
            super().__init__(f=types.SimpleNamespace(normalizable=False, categorical=False, cardinality=0))

        def get_slice(self, name):
            return slice(0, 2)

    sc = bnorm.BatchScaler(Builder())
    for r in (4, 9):
        s = bnorm.Scaler()
        s.center_, s.scale_ = torch.full((3,), 0.5), torch.full((3,), 2.0)
        sc.meg_scalers[r] = s
    meg = torch.randn(3, 3, 40, device=DEV)
    meg[0, 1, 7] = float("nan")
    batch = synthetic.SyntheticBatch(meg, torch.zeros(3, dtype=torch.long, device=DEV), [],
                                     features=torch.randn(3, 2, 40, device=DEV),
                                     features_mask=torch.ones(3, 1, 40, dtype=torch.bool, device=DEV),
                                     recording_index=torch.tensor([4, 9, 5], device=DEV))
    out, feats, mask, keep = bnorm.ScaleReject(sc, limit=20.0, clip=True).prepare(batch, 4)
    assert keep.all() and out.shape == (3, 3, 36)
    assert torch.isnan(out[0, 1, 3]) and torch.isnan(out[0]).sum() == 1
    assert torch.isfinite(out[1]).all()
    assert torch.isnan(out[2]).all()                               # recording 5 was never fitted
    want = ((meg[1] - 0.5) / 2.0).clamp(-20, 20)[:, 4:]
    assert torch.equal(out[1], want)
    assert torch.equal(feats, batch.features[..., :-4])


def _ablation_step(name):
    import numpy as np
    from conftest import load_arrays, rel_err
    from oracle.make_golden import ABLATION_BASE as c, ABLATIONS
    from oracle.ref_loader import clip_conv_kwargs
    import brainmagick_b200 as bb
    from brainmagick_b200 import synthetic
    g = load_arrays(name)
    kw = clip_conv_kwargs(hidden=c["hidden"], depth=c["depth"], merger_channels=c["MC"], initial_linear=c["IL"],
                          merger_pos_dim=c["P"])
    kw.update(ABLATIONS[name])
    model = bb.SimpleConv(in_channels=dict(meg=c["C"]), out_channels=c["F"], n_subjects=c["S"], **kw)
    model.load_state_dict({k[2:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("p.")}, strict=True)
    model = model.to(DEV).train()
    meg = torch.from_numpy(g["meg"]).to(DEV)
    subj = torch.from_numpy(g["subject_index"])
    batch = synthetic.make_batch(meg, subj.to(DEV), torch.from_numpy(g["rec_positions"]), subj)
    if model.merger is not None:
        model.merger.ban_centre_override = torch.from_numpy(g["ban_centre"])
    est = model(dict(meg=meg), batch)
    loss = ClipLoss().to(DEV)(est, torch.from_numpy(g["candidates"]).to(DEV),
                              torch.ones(len(meg), 1, meg.shape[2], dtype=torch.bool, device=DEV))
    loss.backward()
    BF.check_tc_status()
    assert rel_err(est.detach().cpu(), torch.from_numpy(g["estimate"])) < 1e-4
    assert abs(loss.item() - float(g["loss"])) < 1e-4
    scale = max(float(np.linalg.norm(a)) for k, a in g.items() if k.startswith("g.") and k.endswith("weight"))
    for pname, p in model.named_parameters():
        want = torch.from_numpy(g["g." + pname])
        if want.norm() < 1e-5 * scale:
            assert p.grad.abs().max().item() < 1e-4 * scale + 1e-6, pname
        else:
            assert rel_err(p.grad.cpu(), want) < 5e-4, (pname, rel_err(p.grad.cpu(), want))


def test_ablation_reference_configuration():
    """The un-ablated row of the ablation table (a depth-4 clip_conv) against the verbatim-reference fixture."""
    _ablation_step("ablation_reference")


@pytest.mark.parametrize("name", ["ablation_no_glu", "ablation_no_skip", "ablation_relu", "ablation_no_complex_out",
                                  "ablation_no_merger", "ablation_no_initial_linear", "ablation_no_subject_layers",
                                  "ablation_subject_embedding", "ablation_subsample_channels"])
def test_ablation_rows(name):
    """grids/nmi/ablation_final.py:42-52, each against the fixture the verbatim reference produced with the same change."""
    _ablation_step(name)
