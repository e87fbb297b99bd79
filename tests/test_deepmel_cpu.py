"""CPU checks for SURVEY 8(f) row 3 (DeepMel / stand-alone ConvSequence, candidate-side ClipLoss gradient): the oracle
against fixtures produced by the verbatim reference classes, and the drop-in module's construction parity."""
import numpy as np
import pytest
import torch

from conftest import load_arrays, rel_err
from oracle import bm_oracle, deepmel_oracle
from oracle.make_golden import DEEPMEL_CASES


def _spec(case):
    c = DEEPMEL_CASES[case]
    kw = dict(c["params"])
    if "leakiness" not in kw:
        kw["leakiness"] = 0.0
    return deepmel_oracle.deep_mel_spec(c["n_in"], **kw)


@pytest.mark.parametrize("case", list(DEEPMEL_CASES))
def test_deepmel_oracle_matches_reference(case):
    g = load_arrays(case)
    spec = _spec(case)
    p = {k[2:]: torch.from_numpy(v).clone().requires_grad_(v.dtype == np.float32 and "running" not in k)
         for k, v in g.items() if k.startswith("p.")}
    mel = torch.from_numpy(g["mel"])
    est = torch.from_numpy(g["estimate"]).clone().requires_grad_(True)
    stats = {}
    cand = deepmel_oracle.conv_sequence(mel, p, spec, training=True, new_stats=stats)
    cand.retain_grad()
    assert rel_err(cand.detach(), torch.from_numpy(g["candidates"])) < 2e-6
    loss = bm_oracle.clip_loss(est, cand)
    assert abs(float(loss.detach()) - float(g["loss"])) < 1e-6
    loss.backward()
    assert rel_err(est.grad, torch.from_numpy(g["g.estimate"])) < 1e-5
    assert rel_err(cand.grad, torch.from_numpy(g["g.candidates"])) < 1e-5
    for k, v in g.items():
        if k.startswith("g.") and k[2:] in p:
            ref = torch.from_numpy(v)
            if ref.norm() < 1e-6:        # conv biases in front of a training BatchNorm: exactly 0 up to rounding
                assert p[k[2:]].grad.abs().max() < 1e-5
            else:
                assert rel_err(p[k[2:]].grad, ref) < 2e-5, k
    for k, v in stats.items():
        assert rel_err(v, torch.from_numpy(g["bn." + k])) < 1e-6, k
    # eval forward with the updated running statistics
    p_eval = {k: v.detach() for k, v in p.items()}
    p_eval.update({k: v for k, v in stats.items()})
    out = deepmel_oracle.conv_sequence(mel, p_eval, spec, training=False)
    assert rel_err(out, torch.from_numpy(g["candidates_eval"])) < 2e-6


def test_deepmel_module_layout_matches_fixture():
    from brainmagick_b200.features import DeepMel
    for case, c in DEEPMEL_CASES.items():
        g = load_arrays(case)
        torch.manual_seed(c["seed"])
        model = DeepMel(n_in_channels=c["n_in"], **c["params"])
        keys = sorted(k[2:] for k in g if k.startswith("p."))
        assert sorted(model.state_dict().keys()) == keys
        for k, v in model.state_dict().items():
            assert tuple(v.shape) == g["p." + k].shape
        # seeded construction draws the same parameters as the reference (the fixture perturbs BatchNorm only)
        for k, v in model.state_dict().items():
            if k.split('.')[-2] == '0':             # the convolutions (sequence.k.0.*, glus.k.0.*)
                assert np.array_equal(v.numpy(), g["p." + k]), k


def test_convsequence_rejects_what_has_no_kernel():
    from brainmagick_b200.common import ConvSequence
    for kw in (dict(stride=2), dict(stride=1, kernel=5), dict(stride=1, kernel=3, dropout=0.1),
               dict(stride=1, kernel=3, groups=2), dict(stride=1, kernel=3, glu=1, glu_context=2),
               dict(stride=1, kernel=3, activation=torch.nn.Tanh)):
        with pytest.raises(NotImplementedError):
            ConvSequence([4, 4, 4], **kw)
    seq = ConvSequence([4, 4, 4], kernel=3, stride=1)
    with pytest.raises(RuntimeError):        # CPU tensors never reach a CPU implementation
        seq(torch.zeros(1, 4, 8))
