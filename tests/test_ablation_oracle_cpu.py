"""the oracle of the SimpleConv ablation variants (grids/nmi/ablation_final.py:42-52) against
fixtures made by the verbatim reference with the same constructor change.  CPU only; the CUDA variants are the next step."""
import numpy as np
import pytest
import torch

from conftest import load_arrays, rel_err
from oracle import ablation_oracle, bm_oracle
from oracle.make_golden import ABLATION_BASE, ABLATIONS


@pytest.mark.parametrize("name", list(ABLATIONS))
def test_ablation_oracle_matches_reference(name):
    g = load_arrays(name)
    c = ABLATION_BASE
    v = ablation_oracle.Variant(in_channels=c["C"], out_channels=c["F"], n_subjects=c["S"], hidden=c["hidden"],
                                depth=c["depth"], merger_channels=c["MC"], initial_linear=c["IL"],
                                merger_pos_dim=c["P"])._replace(**ABLATIONS[name])
    p = {k[2:]: torch.from_numpy(a).clone() for k, a in g.items() if k.startswith("p.")}
    for k, t in p.items():
        if t.is_floating_point() and "running" not in k:
            t.requires_grad_(True)
    B = len(g["meg"])
    subj = torch.from_numpy(g["subject_index"])
    est = ablation_oracle.forward(p, v, torch.from_numpy(g["meg"]), torch.from_numpy(g["rec_positions"]), subj, subj,
                                  training=True, ban_centre=torch.from_numpy(g["ban_centre"]))
    assert est.shape == g["estimate"].shape
    assert rel_err(est.detach(), torch.from_numpy(g["estimate"])) < 2e-6
    loss = bm_oracle.clip_loss(est, torch.from_numpy(g["candidates"]))
    assert abs(loss.item() - float(g["loss"])) < 1e-6
    loss.backward()
    scale = max(float(np.linalg.norm(a)) for k, a in g.items() if k.startswith("g.") and k.endswith("weight"))
    for k, a in g.items():
        if not k.startswith("g.") or a.size == 0:
            continue
        got, want = p[k[2:]].grad, torch.from_numpy(a)
        if want.norm() < 1e-5 * scale:          # biases in front of a training BatchNorm
            assert got.abs().max() < 1e-5 * scale, k
        else:
            assert rel_err(got, want) < 3e-5, (k, rel_err(got, want))


@pytest.mark.parametrize("name", list(ABLATIONS))
def test_ablation_constructor_matches_reference_layout_and_init(name):
    """The drop-in SimpleConv with the same constructor change: same state_dict keys/shapes as the verbatim reference and,
    seeded alike, the same initial parameters (the fixtures perturb the BatchNorm affine parameters only)."""
    import brainmagick_b200 as bb
    from oracle.ref_loader import clip_conv_kwargs
    g = load_arrays(name)
    c = ABLATION_BASE
    kw = clip_conv_kwargs(hidden=c["hidden"], depth=c["depth"], merger_channels=c["MC"], initial_linear=c["IL"],
                          merger_pos_dim=c["P"])
    kw.update(ABLATIONS[name])
    torch.manual_seed(c["seed"])
    model = bb.SimpleConv(in_channels=dict(meg=c["C"]), out_channels=c["F"], n_subjects=c["S"], **kw)
    sd = model.state_dict()
    assert sorted(sd) == sorted(k[2:] for k in g if k.startswith("p."))
    for k, v in sd.items():
        want = g["p." + k]
        assert tuple(v.shape) == want.shape, k
        is_bn_affine = k.startswith("encoders.meg.sequence.") and k.split(".")[-2] == "1" and k.endswith(("weight", "bias"))
        if not is_bn_affine:
            assert np.array_equal(v.numpy(), want), k
    assert (model.merger is None) == (not kw.get("merger", True))
    assert (model.subject_embedding is not None) == bool(kw.get("subject_dim"))
