"""The drop-in modules run on the CPU through `tests/abi_emulator.py` (a torch restatement of each FP32 entry point's
documented contract) against the verbatim-reference fixtures.

  * The fixtures of the default configuration pass on the GPU; that they also pass here validates the emulator's reading of
    the C-ABI contracts AND pins the host-side composition numerically without a GPU.
  * Compositions that have not had their first GPU run (the ablation rows) are held to their fixtures the same way, so what
    remains unverified for them is only kernels that are individually verified elsewhere.
The CUDA kernels themselves are never exercised here; `-m gpu` tests do that."""
import numpy as np
import pytest
import torch

import abi_emulator
from conftest import GOLDEN_CASES, load_arrays, load_golden, rel_err

TOL = 2e-5


def _check_grads(model, ref, tol=5e-5):
    scale = max(float(np.linalg.norm(np.asarray(v))) for k, v in ref.items() if k.endswith("weight") and np.asarray(v).size)
    for name, p in model.named_parameters():
        want = torch.as_tensor(np.asarray(ref[name]))
        if want.numel() == 0:
            continue
        assert p.grad is not None, name
        if want.norm() < 1e-5 * scale:
            assert p.grad.abs().max().item() < 1e-4 * scale + 1e-6, name
        else:
            assert rel_err(p.grad, want) < tol, (name, rel_err(p.grad, want))


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_default_configuration_on_the_emulator(name):
    import brainmagick_b200 as bb
    from brainmagick_b200 import synthetic
    cfg, train, t = load_golden(name)
    model = bb.SimpleConv(
        in_channels=dict(meg=cfg.in_channels), out_channels=cfg.out_channels, hidden=dict(meg=cfg.hidden),
        depth=cfg.depth, dilation_period=cfg.dilation_period, kernel_size=cfg.kernel_size, skip=True,
        subject_layers=True, subject_dim=0, complex_out=True, glu=cfg.glu, glu_context=cfg.glu_context, merger=True,
        initial_linear=cfg.initial_linear, merger_channels=cfg.merger_channels, gelu=True, batch_norm=True,
        merger_pos_dim=cfg.merger_pos_dim, merger_dropout=cfg.merger_dropout, n_subjects=cfg.n_subjects)
    model.load_state_dict({k[2:]: v for k, v in t.items() if k.startswith("p.")}, strict=True)
    model.train(train)
    model.merger.ban_centre_override = t["ban_centre"]
    batch = synthetic.make_batch(t["meg"], t["subject_index"], t["rec_positions"], t["rec_of_sample"])
    clip = bb.ClipLoss().train(train)
    with abi_emulator.emulated():
        est = model(dict(meg=t["meg"]), batch)
        loss = clip(est, t["candidates"], torch.ones(len(t["meg"]), 1, t["meg"].shape[2], dtype=torch.bool))
        loss.backward()
        probs = clip.get_probabilities(est.detach(), t["candidates"])
    assert rel_err(est.detach(), t["estimate"]) < TOL
    assert abs(loss.item() - t["loss"].item()) < 1e-5
    assert rel_err(probs, t["probs"]) < TOL
    _check_grads(model, {k[2:]: v.numpy() for k, v in t.items() if k.startswith("g.")})
    sd = model.state_dict()
    for key, v in t.items():
        if key.startswith("bn."):
            if "num_batches" in key:
                assert int(sd[key[3:]]) == int(v)
            else:
                assert rel_err(sd[key[3:]], v) < TOL, key


def _ablation_names():
    from oracle.make_golden import ABLATIONS
    return list(ABLATIONS)


@pytest.mark.parametrize("name", _ablation_names())
def test_ablation_rows_on_the_emulator(name):
    import brainmagick_b200 as bb
    from brainmagick_b200 import synthetic
    from oracle.make_golden import ABLATION_BASE as c, ABLATIONS
    from oracle.ref_loader import clip_conv_kwargs
    g = load_arrays(name)
    kw = clip_conv_kwargs(hidden=c["hidden"], depth=c["depth"], merger_channels=c["MC"], initial_linear=c["IL"],
                          merger_pos_dim=c["P"])
    kw.update(ABLATIONS[name])
    model = bb.SimpleConv(in_channels=dict(meg=c["C"]), out_channels=c["F"], n_subjects=c["S"], **kw)
    model.load_state_dict({k[2:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("p.")}, strict=True)
    model.train()
    if model.merger is not None:
        model.merger.ban_centre_override = torch.from_numpy(g["ban_centre"])
    meg, subj = torch.from_numpy(g["meg"]), torch.from_numpy(g["subject_index"])
    batch = synthetic.make_batch(meg, subj, torch.from_numpy(g["rec_positions"]), subj)
    with abi_emulator.emulated():
        est = model(dict(meg=meg), batch)
        loss = bb.ClipLoss()(est, torch.from_numpy(g["candidates"]), torch.ones(len(meg), 1, meg.shape[2], dtype=torch.bool))
        loss.backward()
    assert rel_err(est.detach(), torch.from_numpy(g["estimate"])) < TOL
    assert abs(loss.item() - float(g["loss"])) < 1e-5
    _check_grads(model, {k[2:]: v for k, v in g.items() if k.startswith("g.")})


def _deepmel_names():
    from oracle.make_golden import DEEPMEL_CASES
    return list(DEEPMEL_CASES)


@pytest.mark.parametrize("case", _deepmel_names())
def test_deepmel_on_the_emulator(case):
    import brainmagick_b200 as bb
    from oracle.make_golden import DEEPMEL_CASES
    g = load_arrays(case)
    c = DEEPMEL_CASES[case]
    model = bb.DeepMel(n_in_channels=c["n_in"], **c["params"])
    model.load_state_dict({k[2:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("p.")}, strict=True)
    model.train()
    mel = torch.from_numpy(g["mel"]).requires_grad_(True)
    est = torch.from_numpy(g["estimate"]).requires_grad_(True)
    with abi_emulator.emulated():
        cand = model(mel)
        cand.retain_grad()
        loss = bb.ClipLoss()(est, cand, torch.ones(len(est), 1, est.shape[2], dtype=torch.bool))
        loss.backward()
        model.eval()
        with torch.no_grad():
            out_eval = model(mel.detach())
    assert rel_err(cand.detach(), torch.from_numpy(g["candidates"])) < TOL
    assert abs(loss.item() - float(g["loss"])) < 1e-5
    assert rel_err(est.grad, torch.from_numpy(g["g.estimate"])) < 5e-5
    assert rel_err(cand.grad, torch.from_numpy(g["g.candidates"])) < 5e-5
    assert mel.grad is not None and torch.isfinite(mel.grad).all()
    _check_grads(model, {k[2:]: v for k, v in g.items() if k.startswith("g.") and k[2:] not in ("estimate", "candidates")})
    assert rel_err(out_eval, torch.from_numpy(g["candidates_eval"])) < TOL
