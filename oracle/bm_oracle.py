"""TEST INFRASTRUCTURE ONLY -- CPU restatement of brainmagick's contrastive hot path.

This file is the *oracle* of the repo: a plain, functional, CPU restatement (torch CPU ops; fp32 or fp64
by dtype of the inputs) of exactly what the reference computes in
    bm/models/simpleconv.py:198-249  (SimpleConv.forward at the `clip_conv` configuration)
    bm/models/common.py:45-62, 79-151, 183-271, 312-362 (SubjectLayers, ConvSequence, PositionGetter,
                                                         FourierEmb, ChannelMerger)
    bm/losses.py:77-114             (ClipLoss.get_scores / get_probabilities / forward)
Backward = torch autograd over this restated forward (the reference has no hand-written backward either).

Who may import it: `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` -- as the checker / the CPU baseline, never as the product.  `brainmagick_b200/` never imports it.

Parity pinning: the reference's own tests hold NO golden vectors for this path, so the
restatement is pinned against the verbatim reference modules executed in the build container
(`oracle/make_golden.py` -> `tests/golden/*.npz`, checked by `tests/test_oracle.py`).

Parameters use the reference's `state_dict` key names.
"""
from __future__ import annotations

import math
import typing as tp

import torch
import torch.nn.functional as F

INVALID = -0.1  # common.py:184


# ------------------------------------------------------------------------------------------------
# configuration of the `clip_conv` family
# ------------------------------------------------------------------------------------------------
class Config(tp.NamedTuple):
    """Hyper-parameters of the clip_conv family (conf/model/clip_conv.yaml:6-22, model_defaults/defaults.yaml:35-82)."""
    in_channels: int            # C, sensors (padded)
    out_channels: int           # F
    n_subjects: int             # S
    hidden: int = 320
    depth: int = 10
    kernel_size: int = 3
    dilation_growth: int = 2
    dilation_period: int = 5
    glu: int = 2
    glu_context: int = 1
    merger_channels: int = 270
    merger_pos_dim: int = 2048
    merger_dropout: float = 0.2
    initial_linear: int = 270
    bn_eps: float = 1e-5
    bn_momentum: float = 0.1

    def dilations(self):
        """common.py:108-115: dilation resets to 1 every `dilation_period` layers, else grows."""
        out, d = [], 1
        for k in range(self.depth):
            if self.dilation_period and (k % self.dilation_period) == 0:
                d = 1
            out.append(d)
            d *= self.dilation_growth
        return out

    def has_glu(self, k: int) -> bool:
        """common.py:133: a GLU block follows layer k when (k+1) % glu == 0."""
        return bool(self.glu) and (k + 1) % self.glu == 0


# ------------------------------------------------------------------------------------------------
# A.1  positions -> Fourier embedding -> spatial-attention weights
# ------------------------------------------------------------------------------------------------
def normalise_layout(xy: torch.Tensor) -> torch.Tensor:
    """common.py:215-217: min-max normalise x and y independently over the valid channels."""
    x, y = xy[:, 0], xy[:, 1]
    x = (x - x.min()) / (x.max() - x.min())
    y = (y - y.min()) / (y.max() - y.min())
    return torch.stack([x, y], dim=1)


def fourier_emb(positions: torch.Tensor, dimension: int, margin: float = 0.2) -> torch.Tensor:
    """common.py:254-271.  positions [..., 2] -> [..., dimension]; cos block then sin block; the x
    frequency is the slow index (freqs_x is the column vector)."""
    *O, D = positions.shape
    assert D == 2
    n_freqs = (dimension // 2) ** 0.5
    freqs_y = torch.arange(n_freqs).to(positions)
    freqs_x = freqs_y[:, None]
    width = 1 + 2 * margin
    positions = positions + margin
    p_x = 2 * math.pi * freqs_x / width
    p_y = 2 * math.pi * freqs_y / width
    positions = positions[..., None, None, :]
    loc = (positions[..., 0] * p_x + positions[..., 1] * p_y).view(*O, -1)
    return torch.cat([torch.cos(loc), torch.sin(loc)], dim=-1)


def attention_weights(positions: torch.Tensor, heads: torch.Tensor,
                      ban_centre: tp.Optional[torch.Tensor] = None, radius: float = 0.2) -> torch.Tensor:
    """common.py:337-357.  positions [R,C,2] (one row per recording), heads [O,P] -> softmax weights [R,O,C].

    The reference evaluates this per *sample*; the weights only depend on the recording's layout and on the
    single per-forward ban centre (common.py:343), so one row per recording is the same arithmetic."""
    emb = fourier_emb(positions, heads.shape[1])                                   # [R,C,P]
    offset = torch.zeros(positions.shape[:2], dtype=positions.dtype, device=positions.device)
    offset[(positions.float() == INVALID).all(dim=-1)] = float("-inf")             # common.py:339-340 (fp32 compare)
    if ban_centre is not None:
        banned = (positions - ban_centre.to(positions)).norm(dim=-1) <= radius     # common.py:344-346
        offset[banned] = float("-inf")
    scores = torch.einsum("rcd,od->roc", emb, heads)                               # common.py:355
    scores = scores + offset[:, None]
    return torch.softmax(scores, dim=2)                                            # common.py:357


# ------------------------------------------------------------------------------------------------
# A.2  sensor chain: spatial mix -> initial 1x1 conv -> per-subject 1x1
# ------------------------------------------------------------------------------------------------
def sensor_chain(meg, weights_per_sample, il_weight, il_bias, subj_weights, subject_index):
    """common.py:358 ; simpleconv.py:213-214 ; common.py:55-58 (SubjectLayers contracts its FIRST weight axis)."""
    u = torch.einsum("bct,boc->bot", meg, weights_per_sample)
    v = F.conv1d(u, il_weight, il_bias)
    m = subj_weights[subject_index]                                                 # gather, common.py:57
    return torch.einsum("bct,bcd->bdt", v, m)


# ------------------------------------------------------------------------------------------------
# A.3 / A.4  ConvSequence
# ------------------------------------------------------------------------------------------------
def batch_norm_train(y, gamma, beta, eps):
    """BatchNorm1d in training mode (common.py:118-119): biased batch statistics over (B,T)."""
    mean = y.mean(dim=(0, 2))
    var = y.var(dim=(0, 2), unbiased=False)
    yhat = (y - mean[None, :, None]) / torch.sqrt(var[None, :, None] + eps)
    return yhat * gamma[None, :, None] + beta[None, :, None], mean, var


def conv_sequence(x, p: tp.Dict[str, torch.Tensor], cfg: Config, training: bool, prefix="encoders.meg.",
                  bn_updates: tp.Optional[dict] = None):
    """common.py:142-151 with the layer bodies of common.py:98-140 at the clip_conv settings."""
    for k, d in enumerate(cfg.dilations()):
        old_x = x
        w, b = p[f"{prefix}sequence.{k}.0.weight"], p[f"{prefix}sequence.{k}.0.bias"]
        y = F.conv1d(x, w, b, stride=1, padding=(cfg.kernel_size // 2) * d, dilation=d)   # common.py:112-114
        g, be = p[f"{prefix}sequence.{k}.1.weight"], p[f"{prefix}sequence.{k}.1.bias"]
        rm, rv = p[f"{prefix}sequence.{k}.1.running_mean"], p[f"{prefix}sequence.{k}.1.running_var"]
        if training:
            z, mean, var = batch_norm_train(y, g, be, cfg.bn_eps)
            if bn_updates is not None:
                n = y.shape[0] * y.shape[2]
                bn_updates[f"{prefix}sequence.{k}.1.running_mean"] = \
                    (1 - cfg.bn_momentum) * rm + cfg.bn_momentum * mean.detach()
                bn_updates[f"{prefix}sequence.{k}.1.running_var"] = \
                    (1 - cfg.bn_momentum) * rv + cfg.bn_momentum * var.detach() * n / (n - 1)
        else:
            z = (y - rm[None, :, None]) / torch.sqrt(rv[None, :, None] + cfg.bn_eps) * g[None, :, None] \
                + be[None, :, None]
        x = F.gelu(z)                                                                 # exact erf GELU, common.py:120
        if x.shape == old_x.shape:                                                    # skip, common.py:146-147
            x = x + old_x
        if cfg.has_glu(k):                                                            # common.py:133-138,148-150
            gw, gb = p[f"{prefix}glus.{k}.0.weight"], p[f"{prefix}glus.{k}.0.bias"]
            h = F.conv1d(x, gw, gb, padding=cfg.glu_context)
            x = F.glu(h, dim=1)
    return x


# ------------------------------------------------------------------------------------------------
# A.5  head
# ------------------------------------------------------------------------------------------------
def head(x, p):
    """simpleconv.py:185-189,246-247: Conv1d(H,2H,1) -> GELU -> ConvTranspose1d(2H,F,1) (weight [2H,F,1])."""
    q = F.gelu(F.conv1d(x, p["final.0.weight"], p["final.0.bias"]))
    return F.conv_transpose1d(q, p["final.2.weight"], p["final.2.bias"])


def simpleconv_forward(p: tp.Dict[str, torch.Tensor], cfg: Config, meg, rec_positions, rec_of_sample,
                       subject_index, training: bool, ban_centre=None, bn_updates=None):
    """simpleconv.py:198-249 at clip_conv.  `rec_positions` [R,C,2] are per-recording sensor positions
    (INVALID for padded channels), `rec_of_sample` [B] maps each sample to its row of `rec_positions`."""
    length = meg.shape[-1]
    centre = ban_centre if (training and cfg.merger_dropout) else None
    w = attention_weights(rec_positions, p["merger.heads"], centre, cfg.merger_dropout)
    x = sensor_chain(meg, w[rec_of_sample], p["initial_linear.0.weight"], p["initial_linear.0.bias"],
                     p["subject_layers.weights"], subject_index)
    x = conv_sequence(x, p, cfg, training, bn_updates=bn_updates)
    x = head(x, p)
    return x[:, :, :length]


# ------------------------------------------------------------------------------------------------
# A.6  ClipLoss
# ------------------------------------------------------------------------------------------------
def clip_scores(estimates, candidates):
    """losses.py:91-94 (pool/center/linear/trim all off at the default configuration)."""
    inv_norms = 1 / (1e-8 + candidates.norm(dim=(1, 2), p=2))
    return torch.einsum("bct,oct,o->bo", estimates, candidates, inv_norms)


def clip_probabilities(estimates, candidates):
    """losses.py:97-102."""
    return F.softmax(clip_scores(estimates, candidates), dim=1)


def clip_loss(estimate, candidate, target_offset: int = 0):
    """losses.py:104-114.  `target_offset` is the multi-GPU extension of SURVEY.md 8(e): rank r's rows match
    candidates r*B_loc + arange(B_loc); 0 reproduces the reference exactly."""
    assert estimate.size(0) <= candidate.size(0)
    scores = clip_scores(estimate, candidate)
    target = torch.arange(len(scores), device=scores.device) + target_offset
    return F.cross_entropy(scores, target)


# ------------------------------------------------------------------------------------------------
# training step (restates solver.py:297,373,384-385 without the Solver)
# ------------------------------------------------------------------------------------------------
def training_step(p, cfg, meg, rec_positions, rec_of_sample, subject_index, candidates, ban_centre=None,
                  training=True, all_candidates=None, target_offset=0):
    """forward + loss + backward.  Returns dict(estimate, scores, loss, grads{name: tensor}, bn_updates)."""
    params = {k: (v.detach().clone().requires_grad_(True) if v.is_floating_point() and
                  not k.endswith(("running_mean", "running_var")) else v) for k, v in p.items()}
    bn_updates: dict = {}
    est = simpleconv_forward(params, cfg, meg, rec_positions, rec_of_sample, subject_index, training,
                             ban_centre, bn_updates)
    cands = candidates if all_candidates is None else all_candidates
    scores = clip_scores(est, cands)
    loss = F.cross_entropy(scores, torch.arange(len(scores), device=scores.device) + target_offset)
    names = [k for k, v in params.items() if v.requires_grad]
    grads = torch.autograd.grad(loss, [params[k] for k in names], allow_unused=True)
    return dict(estimate=est.detach(), scores=scores.detach(), loss=loss.detach(),
                grads={k: g for k, g in zip(names, grads)}, bn_updates=bn_updates)


# ------------------------------------------------------------------------------------------------
# retrieval accuracy (scripts/run_eval_probs.py:237-264 semantics on unique candidates)
# ------------------------------------------------------------------------------------------------
def topk_accuracy(estimates, candidates, true_index, k=10, chunk=256, on="probs"):
    """Top-k segment retrieval (scripts/run_eval_probs.py:237-264: `probs.topk(k)` holds the true segment).

    on="probs" is the reference's literal arithmetic: the fp32 softmax of losses.py:97-102, then topk.  When score gaps
    exceed ~87 the softmax underflows to exact zeros and topk's choice among the tied zeros is unspecified (see
    `degenerate_rows`); on="scores" ranks the scores themselves -- the same order wherever the probabilities are distinct."""
    hits = 0
    for i in range(0, len(estimates), chunk):
        vals = clip_scores(estimates[i:i + chunk], candidates)
        if on == "probs":
            vals = torch.softmax(vals, dim=1)
        top = vals.topk(k, dim=1).indices
        hits += (top == true_index[i:i + chunk, None]).any(dim=1).sum().item()
    return hits / len(estimates)


def degenerate_rows(estimates, candidates, k=10, chunk=256) -> int:
    """Rows whose fp32 softmax has fewer than k distinct non-zero probabilities among its k+1 best: there the reference's
    `probs.topk(k)` is decided by tie-breaking, not by the model."""
    n = 0
    for i in range(0, len(estimates), chunk):
        top = clip_probabilities(estimates[i:i + chunk], candidates).topk(k + 1, dim=1).values
        n += int((top[:, 1:] == top[:, :-1]).any(dim=1).sum().item())
    return n


# ------------------------------------------------------------------------------------------------
# synthetic inputs; shared byte-for-byte by the CPU and GPU runs
# ------------------------------------------------------------------------------------------------
def init_state_dict(cfg: Config, seed: int = 0, dtype=torch.float32) -> tp.Dict[str, torch.Tensor]:
    """Random parameters with the reference's shapes and init *scales* (A.7); not RNG-identical to the
    reference constructor (fixtures made from the reference carry its own state_dict)."""
    g = torch.Generator().manual_seed(seed)
    H, MC, IL, P = cfg.hidden, cfg.merger_channels, cfg.initial_linear, cfg.merger_pos_dim

    def uni(shape, fan_in):
        bound = 1 / math.sqrt(fan_in)
        return (torch.rand(shape, generator=g, dtype=dtype) * 2 - 1) * bound

    p = {}
    p["merger.heads"] = torch.randn(MC, P, generator=g, dtype=dtype) / P ** 0.5
    p["initial_linear.0.weight"] = uni((IL, MC, 1), MC)
    p["initial_linear.0.bias"] = uni((IL,), MC)
    p["subject_layers.weights"] = torch.randn(cfg.n_subjects, IL, IL, generator=g, dtype=dtype) / IL ** 0.5
    p["final.0.weight"] = uni((2 * H, H, 1), H)
    p["final.0.bias"] = uni((2 * H,), H)
    p["final.2.weight"] = uni((2 * H, cfg.out_channels, 1), cfg.out_channels)
    p["final.2.bias"] = uni((cfg.out_channels,), cfg.out_channels)
    chin = IL
    for k in range(cfg.depth):
        pre = f"encoders.meg.sequence.{k}."
        p[pre + "0.weight"] = uni((H, chin, cfg.kernel_size), chin * cfg.kernel_size)
        p[pre + "0.bias"] = uni((H,), chin * cfg.kernel_size)
        p[pre + "1.weight"] = 1 + 0.1 * torch.randn(H, generator=g, dtype=dtype)
        p[pre + "1.bias"] = 0.1 * torch.randn(H, generator=g, dtype=dtype)
        p[pre + "1.running_mean"] = 0.1 * torch.randn(H, generator=g, dtype=dtype)
        p[pre + "1.running_var"] = 1 + 0.2 * torch.rand(H, generator=g, dtype=dtype)
        p[pre + "1.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
        if cfg.has_glu(k):
            kk = 1 + 2 * cfg.glu_context
            p[f"encoders.meg.glus.{k}.0.weight"] = uni((2 * H, H, kk), H * kk)
            p[f"encoders.meg.glus.{k}.0.bias"] = uni((2 * H,), H * kk)
        chin = H
    return p


def synthetic_batch(cfg: Config, batch: int, T: int = 360, seed: int = 2036, n_valid: tp.Sequence[int] = (),
                    n_candidates: tp.Optional[int] = None):
    """meg ~ N(0,1).clamp(+-20); candidates ~ N(0,1); subject ~ U{0..S-1}; one recording per subject with
    positions ~ U[0,1]^2 (INVALID on padded channels).  `n_valid` cycles over subjects (mixed studies)."""
    g = torch.Generator().manual_seed(seed)
    C, S = cfg.in_channels, cfg.n_subjects
    meg = torch.randn(batch, C, T, generator=g).clamp_(-20, 20)
    cand = torch.randn(n_candidates or batch, cfg.out_channels, T, generator=g)
    subj = torch.randint(0, S, (batch,), generator=g)
    pos = torch.rand(S, C, 2, generator=g)
    if n_valid:
        for s in range(S):
            nv = n_valid[s % len(n_valid)]
            pos[s, nv:] = INVALID
        for b in range(batch):
            meg[b, n_valid[int(subj[b]) % len(n_valid)]:] = 0            # zero padding, dataset.py:353-354
    ban = torch.rand(2, generator=g)
    return dict(meg=meg, candidates=cand, subject_index=subj, rec_positions=pos,
                rec_of_sample=subj.clone(), ban_centre=ban)


# ------------------------------------------------------------------------------------------------
# CPU baseline: the same training step on the host cores (restates solver.py:297,373,384-387 + train.py:119)
# ------------------------------------------------------------------------------------------------
class CpuTrainer:
    """forward + ClipLoss + backward + Adam(lr=3e-4, betas=(0.9, 0.999)) on CPU tensors, with BatchNorm running
    statistics carried across steps -- what `bench.py --impl reference` and the `cpu_baseline` leg time."""

    def __init__(self, cfg: Config, params: tp.Dict[str, torch.Tensor], lr: float = 3e-4):
        self.cfg = cfg
        self.p = {k: v.clone() for k, v in params.items()}
        self.leaves = {k: v.requires_grad_(True) for k, v in self.p.items()
                       if v.is_floating_point() and "running" not in k}
        self.opt = torch.optim.Adam(list(self.leaves.values()), lr=lr, betas=(0.9, 0.999))

    def step(self, meg, rec_positions, rec_of_sample, subject_index, candidates, ban_centre=None) -> float:
        self.opt.zero_grad(set_to_none=True)
        bn_updates: dict = {}
        est = simpleconv_forward(self.p, self.cfg, meg, rec_positions, rec_of_sample, subject_index, True,
                                 ban_centre, bn_updates)
        loss = clip_loss(est, candidates)
        loss.backward()
        self.opt.step()
        with torch.no_grad():
            for k, v in bn_updates.items():
                self.p[k].copy_(v)
        return float(loss.detach())
