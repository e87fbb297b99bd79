"""autograd.Functions that run the hot path through the C ABI (libbm_b200.so).

`encoder_forward` is SimpleConv.forward at the clip_conv configuration (bm/models/simpleconv.py:198-249) as ONE
autograd node: its backward launches the hand-written gradient kernels in reverse order, so `loss.backward()`
leaves ordinary dense `.grad` tensors on every parameter (what bm/solver.py:384-387 and
flashy.distrib.sync_model expect).  `clip_loss` / `clip_scores` are ClipLoss (bm/losses.py:77-114).
"""
from __future__ import annotations

import dataclasses
import typing as tp

import torch

from . import _lib
from ._lib import call, ptr, stream


@dataclasses.dataclass
class EncoderPlan:
    """Everything `encoder_forward` needs that is not a differentiable tensor."""
    dilations: tp.List[int]
    glu_after: tp.List[bool]
    kernel_size: int
    glu_kernel: int
    training: bool
    bn_eps: float
    bn_momentum: float
    rec_positions: torch.Tensor            # [R,C,2] fp32, rows = recordings present in the batch
    rec_of_sample: torch.Tensor            # [B] int32 -> row of rec_positions
    rec_order: torch.Tensor                # [B] int32 samples sorted by recording
    rec_off: torch.Tensor                  # [R+1] int32 CSR offsets
    subject: torch.Tensor                  # [B] int32
    freq: torch.Tensor                     # [sqrt(P/2)] fp32 Fourier frequencies
    ban_centre: tp.Optional[torch.Tensor]  # [2] fp32 or None
    ban_radius: float
    bn_buffers: tp.List[tp.Tuple[torch.Tensor, torch.Tensor]]   # (running_mean, running_var) per layer
    keep_for_backward: bool = True


def _empty(shape, like, dtype=torch.float32):
    return torch.empty(shape, device=like.device, dtype=dtype)


class _EncoderFn(torch.autograd.Function):
    """inputs: plan, meg, heads, il_w, il_b, subj_w, w0, b0, w2, b2, then per layer k: (conv_w, conv_b, gamma, beta)
    and, appended in layer order, (glu_w, glu_b) for every layer followed by a GLU block."""

    @staticmethod
    def forward(ctx, plan: EncoderPlan, meg, heads, il_w, il_b, subj_w, w0, b0, w2, b2, *layer_params):
        depth = len(plan.dilations)
        conv_p = [layer_params[4 * k:4 * k + 4] for k in range(depth)]
        glu_flat = layer_params[4 * depth:]
        glu_p, gi = {}, 0
        for k in range(depth):
            if plan.glu_after[k]:
                glu_p[k] = glu_flat[2 * gi:2 * gi + 2]
                gi += 1
        st = stream()
        meg = meg.contiguous()
        B, C, T = meg.shape
        R = plan.rec_positions.shape[0]
        O, P = heads.shape
        IL = il_w.shape[0]
        S, _, D = subj_w.shape
        H = conv_p[0][0].shape[0]
        F = w2.shape[1]
        rows = B * T
        save = plan.keep_for_backward

        # K1 attention weights per recording
        emb = _empty((R, C, P), meg)
        att = _empty((R, O, C), meg)
        call("bm_attention_weights_fwd", ptr(plan.rec_positions), ptr(plan.freq), ptr(heads.contiguous()),
             ptr(plan.ban_centre), float(plan.ban_radius), R, C, O, P, ptr(emb), ptr(att), st)
        # K2 sensor chain
        u = _empty((B, T, O), meg)
        v = _empty((B, T, IL), meg)
        x = _empty((B, T, D), meg)
        il_w2 = il_w.reshape(IL, O).contiguous()
        call("bm_sensor_chain_fwd", ptr(meg), ptr(att), ptr(plan.rec_of_sample), ptr(il_w2), ptr(il_b.contiguous()),
             ptr(subj_w.contiguous()), ptr(plan.subject), B, C, T, O, IL, D, ptr(u), ptr(v), ptr(x), st)

        # K3/K4 ConvSequence
        stats = _empty((2 * H,), meg, torch.float64)
        saved_layers = []
        for k in range(depth):
            cw, cb, gamma, beta = conv_p[k]
            cout, cin, kw = cw.shape
            wf = _empty((kw, cin, cout), meg)
            wb = _empty((kw, cout, cin), meg)
            call("bm_conv_weight_prep", ptr(cw.contiguous()), cout, cin, kw, ptr(wf), ptr(wb), st)
            y = _empty((B, T, cout), meg)
            mean = _empty((cout,), meg)
            invstd = _empty((cout,), meg)
            rm, rv = plan.bn_buffers[k]
            if plan.training:
                call("bm_conv1d_fwd", ptr(x), ptr(wf), ptr(cb.contiguous()), B, T, cin, cout, kw, plan.dilations[k],
                     ptr(y), ptr(stats), st)
                call("bm_bn_stats_finalize", ptr(stats), rows, float(plan.bn_eps), float(plan.bn_momentum),
                     ptr(rm), ptr(rv), ptr(mean), ptr(invstd), cout, st)
            else:
                call("bm_conv1d_fwd", ptr(x), ptr(wf), ptr(cb.contiguous()), B, T, cin, cout, kw, plan.dilations[k],
                     ptr(y), None, st)
                call("bm_bn_eval_stats", ptr(rm), ptr(rv), float(plan.bn_eps), ptr(mean), ptr(invstd), cout, st)
            skip = cin == cout
            x_new = _empty((B, T, cout), meg)
            call("bm_bn_gelu_skip_fwd", ptr(y), ptr(mean), ptr(invstd), ptr(gamma.contiguous()),
                 ptr(beta.contiguous()), ptr(x) if skip else None, ptr(x_new), rows, cout, st)
            rec = dict(x_in=x, y=y, mean=mean, invstd=invstd, wb=wb, skip=skip, x_new=x_new)
            x = x_new
            if plan.glu_after[k]:
                gw, gb = glu_p[k]
                g2h, gcin, gkw = gw.shape
                gwf = _empty((gkw, gcin, g2h), meg)
                gwb = _empty((gkw, g2h, gcin), meg)
                call("bm_conv_weight_prep", ptr(gw.contiguous()), g2h, gcin, gkw, ptr(gwf), ptr(gwb), st)
                h = _empty((B, T, g2h), meg) if save else None
                out = _empty((B, T, g2h // 2), meg)
                call("bm_conv1d_glu_fwd", ptr(x), ptr(gwf), ptr(gb.contiguous()), B, T, gcin, g2h // 2, gkw,
                     ptr(h), ptr(out), st)
                rec.update(h=h, gwb=gwb)
                x = out
            saved_layers.append(rec if save else None)

        # K5 head
        w0_2 = w0.reshape(2 * H, H).contiguous()
        w2_2 = w2.reshape(2 * H, F).contiguous()
        h1 = _empty((B, T, 2 * H), meg)
        q = _empty((B, T, 2 * H), meg)
        est = _empty((B, F, T), meg)
        call("bm_head_fwd", ptr(x), ptr(w0_2), ptr(b0.contiguous()), ptr(w2_2), ptr(b2.contiguous()), B, T, H, F,
             ptr(h1), ptr(q), ptr(est), st)

        if save:
            ctx.plan = plan
            ctx.dims = (B, C, T, R, O, P, IL, S, D, H, F)
            ctx.saved = dict(meg=meg, emb=emb, att=att, u=u, v=v, il_w2=il_w2, subj_w=subj_w.contiguous(),
                             layers=saved_layers, x_last=x, w0_2=w0_2, w2_2=w2_2, h1=h1, q=q,
                             conv_p=conv_p, glu_p=glu_p, il_shape=il_w.shape, w0_shape=w0.shape, w2_shape=w2.shape)
        return est

    @staticmethod
    def backward(ctx, dest):
        plan: EncoderPlan = ctx.plan
        s = ctx.saved
        B, C, T, R, O, P, IL, S, D, H, F = ctx.dims
        depth = len(plan.dilations)
        st = stream()
        meg = s["meg"]
        rows = B * T
        dest = dest.contiguous()

        # head
        dq = _empty((B, T, 2 * H), meg)
        g = _empty((B, T, H), meg)
        dw0 = _empty((2 * H, H), meg)
        db0 = _empty((2 * H,), meg)
        dw2 = _empty((2 * H, F), meg)
        db2 = _empty((F,), meg)
        call("bm_head_bwd", ptr(dest), ptr(s["x_last"]), ptr(s["w0_2"]), ptr(s["w2_2"]), ptr(s["h1"]), ptr(s["q"]),
             B, T, H, F, ptr(dq), ptr(g), ptr(dw0), ptr(db0), ptr(dw2), ptr(db2), st)
        del dq

        sums = _empty((2 * H,), meg, torch.float64)
        layer_grads: tp.List[tp.Any] = [None] * depth
        glu_grads = {}
        for k in reversed(range(depth)):
            rec = s["layers"][k]
            cw, cb, gamma, beta = s["conv_p"][k]
            cout, cin, kw = cw.shape
            if plan.glu_after[k]:
                gw, gb = s["glu_p"][k]
                g2h, gcin, gkw = gw.shape
                dh = _empty((B, T, g2h), meg)
                call("bm_glu_bwd", ptr(g), ptr(rec["h"]), rows, g2h // 2, ptr(dh), st)
                dgw = _empty(gw.shape, meg)
                dgb = _empty((g2h,), meg)
                call("bm_conv1d_bwd_weight", ptr(dh), ptr(rec["x_new"]), B, T, gcin, g2h, gkw, 1, ptr(dgw), ptr(dgb), st)
                g = _empty((B, T, gcin), meg)
                call("bm_conv1d_bwd_data", ptr(dh), ptr(rec["gwb"]), None, B, T, gcin, g2h, gkw, 1, ptr(g), st)
                glu_grads[k] = (dgw, dgb)
                del dh
            dy = _empty((B, T, cout), meg)
            dgamma = _empty((cout,), meg)
            dbeta = _empty((cout,), meg)
            call("bm_bn_gelu_skip_bwd", ptr(g), ptr(rec["y"]), ptr(rec["mean"]), ptr(rec["invstd"]),
                 ptr(gamma.contiguous()), ptr(beta.contiguous()), 1 if plan.training else 0, rows, cout, ptr(sums),
                 ptr(dy), ptr(dgamma), ptr(dbeta), st)
            dcw = _empty(cw.shape, meg)
            dcb = _empty((cout,), meg)
            call("bm_conv1d_bwd_weight", ptr(dy), ptr(rec["x_in"]), B, T, cin, cout, kw, plan.dilations[k],
                 ptr(dcw), ptr(dcb), st)
            g_in = _empty((B, T, cin), meg)
            call("bm_conv1d_bwd_data", ptr(dy), ptr(rec["wb"]), ptr(g) if rec["skip"] else None, B, T, cin, cout, kw,
                 plan.dilations[k], ptr(g_in), st)
            g = g_in
            layer_grads[k] = (dcw, dcb, dgamma, dbeta)
            del dy

        # sensor chain + attention
        subj_order = torch.argsort(plan.subject, stable=True).to(torch.int32)
        counts = torch.bincount(plan.subject, minlength=S)
        subj_off = torch.zeros(S + 1, dtype=torch.int32, device=meg.device)
        subj_off[1:] = torch.cumsum(counts, 0).to(torch.int32)
        dv = _empty((B, T, IL), meg)
        du = _empty((B, T, O), meg)
        d_subj = _empty((S, IL, D), meg)
        d_il_w = _empty((IL, O), meg)
        d_il_b = _empty((IL,), meg)
        d_att = _empty((R, O, C), meg)
        call("bm_sensor_chain_bwd", ptr(g), ptr(meg), ptr(s["il_w2"]), ptr(s["subj_w"]), ptr(plan.subject),
             ptr(s["u"]), ptr(s["v"]), ptr(subj_order), ptr(subj_off), ptr(plan.rec_order), ptr(plan.rec_off),
             B, C, T, O, IL, D, S, R, ptr(dv), ptr(du), ptr(d_subj), ptr(d_il_w), ptr(d_il_b), ptr(d_att), st)
        dscores = _empty((R, O, C), meg)
        dheads = _empty((O, P), meg)
        call("bm_attention_weights_bwd", ptr(d_att), ptr(s["att"]), ptr(s["emb"]), R, C, O, P, ptr(dscores),
             ptr(dheads), st)

        grads = [None, None, dheads, d_il_w.reshape(s["il_shape"]), d_il_b, d_subj,
                 dw0.reshape(s["w0_shape"]), db0, dw2.reshape(s["w2_shape"]), db2]
        for k in range(depth):
            grads.extend(layer_grads[k])
        for k in range(depth):
            if plan.glu_after[k]:
                grads.extend(glu_grads[k])
        ctx.saved = None
        return tuple(grads)


def encoder_forward(plan: EncoderPlan, meg, heads, il_w, il_b, subj_w, w0, b0, w2, b2, layer_params):
    return _EncoderFn.apply(plan, meg, heads, il_w, il_b, subj_w, w0, b0, w2, b2, *layer_params)


# ----------------------------------------------------------------------------------------------------
# ClipLoss
# ----------------------------------------------------------------------------------------------------
def clip_scores(estimates: torch.Tensor, candidates: torch.Tensor, want_probs: bool = False):
    """ClipLoss.get_scores / get_probabilities (bm/losses.py:77-102); no autograd."""
    est = estimates.detach().contiguous().float()
    cand = candidates.detach().contiguous().float()
    Bn, Bc = est.shape[0], cand.shape[0]
    KT = est[0].numel()
    assert cand[0].numel() == KT
    ss = _empty((Bc,), est, torch.float64)
    inv = _empty((Bc,), est)
    scores = _empty((Bn, Bc), est)
    probs = _empty((Bn, Bc), est) if want_probs else None
    call("bm_clip_scores", ptr(est), ptr(cand), Bn, Bc, KT, ptr(ss), ptr(inv), ptr(scores), ptr(probs), stream())
    return probs if want_probs else scores


class _ClipLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, estimate, candidate, target_offset: int):
        est = estimate.contiguous()
        cand = candidate.contiguous()
        Bn, Bc = est.shape[0], cand.shape[0]
        KT = est[0].numel()
        ss = _empty((Bc,), est, torch.float64)
        inv = _empty((Bc,), est)
        scores = _empty((Bn, Bc), est)
        probs = _empty((Bn, Bc), est)
        row_loss = _empty((Bn,), est)
        loss = _empty((1,), est)
        call("bm_clip_loss_fwd", ptr(est), ptr(cand), Bn, Bc, KT, int(target_offset), ptr(ss), ptr(inv),
             ptr(scores), ptr(probs), ptr(row_loss), ptr(loss), stream())
        ctx.save_for_backward(probs, inv, cand)
        ctx.meta = (Bn, Bc, KT, int(target_offset), est.shape)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, gout):
        probs, inv, cand = ctx.saved_tensors
        Bn, Bc, KT, off, shape = ctx.meta
        if ctx.needs_input_grad[1]:
            raise NotImplementedError("gradient w.r.t. candidates (feature_model) is a later row of SURVEY 8(f)")
        G = _empty((Bn, Bc), probs)
        dest = _empty(shape, probs)
        gout = gout.reshape(1).contiguous().float()
        call("bm_clip_loss_bwd", ptr(probs), ptr(inv), ptr(cand), ptr(gout), Bn, Bc, KT, off, ptr(G), ptr(dest),
             stream())
        return dest, None, None


def clip_loss(estimate, candidate, target_offset: int = 0):
    return _ClipLossFn.apply(estimate, candidate, target_offset)


def library_loaded() -> bool:
    return _lib._lib is not None
