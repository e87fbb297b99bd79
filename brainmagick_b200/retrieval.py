"""Retrieval evaluation on the GPU.

Mirrors, with the reference's names and argument meaning:
    scripts/run_eval_probs.py:267-307   builds_probs(clip, preds, trues, dset_args, batch_size, tmin, tmax)
    scripts/run_eval_probs.py:237-264   _get_accuracy_from_probs(probs, target_labels, vocab_labels, topk)
    bm/wer.py:80-116                    the ranking loop of get_wer -> wer_ranking(...)
and adds `retrieval_accuracy`, the fused form the first two are used for (run_eval_probs.py:331-362): scores on the
tensor cores, softmax + top-k + label match in one kernel, nothing but the per-query hit ranks leaves the GPU.

What changes against the reference, deliberately:
  * the candidate set is made resident ONCE (`CandidateBank`: flattened, zero-padded to the tensor-core tile, inverse
    norms computed once) instead of being re-normalised by every `get_probabilities` call (losses.py:91);
  * get_wer scores ONE estimate per call (wer.py:99), a GEMV that streams all negatives from HBM per estimate; here the
    shared negatives are scored for a whole batch of estimates at once and each estimate's own true output, which the
    reference writes into the last negative slot (wer.py:93-94), is a row-wise dot product.
Everything numeric runs through the C ABI (`bm_clip_scores`, `bm_retrieval_*`, `bm_rowdot_scaled`); there is no CPU path.
"""
from __future__ import annotations

import typing as tp

import torch

from . import functional as BF
from ._lib import call, ptr, stream

TILE = 256          # candidate-axis granularity of the tensor-core score GEMM (tc_conv2: N tiles of 256 or 320)


def _window(dset_args, tmin, tmax) -> tp.Tuple[tp.Optional[int], tp.Optional[int]]:
    """run_eval_probs.py:279-290 (same float expression, so the same truncation)."""
    lo = None if tmin is None else int((tmin - dset_args.tmin) * dset_args.sample_rate)
    hi = None if tmax is None else int((tmax - dset_args.tmin) * dset_args.sample_rate)
    return lo, hi


def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("brainmagick_b200.retrieval needs a CUDA device (no CPU path)")
    return torch.device("cuda", torch.cuda.current_device())


class CandidateBank:
    """A fixed candidate set [M, F, T] resident on the GPU as `rows` [M_pad, F*T] (zero rows beyond M) + `inv_norms`."""

    def __init__(self, clip, candidates: torch.Tensor, device: tp.Optional[torch.device] = None,
                 chunk_rows: int = 1024):
        device = device or (candidates.device if candidates.is_cuda else _device())
        probe, _ = clip._prepare(candidates[:1], candidates[:1])
        self.n = candidates.shape[0]
        self.row_elems = probe[0].numel()
        pad = self.n >= TILE // 2 and self.row_elems % 32 == 0
        self.n_pad = -(-self.n // TILE) * TILE if pad else self.n
        self.rows = torch.zeros(self.n_pad, self.row_elems, device=device, dtype=torch.float32)
        for i in range(0, self.n, chunk_rows):                      # chunked: a 20k x 1.47 MB set is never held twice
            part, _ = clip._prepare(candidates[i:i + chunk_rows], candidates[i:i + 1])
            self.rows[i:i + part.shape[0]].copy_(part.reshape(part.shape[0], -1), non_blocking=True)
        self.inv_norms = BF.candidate_inv_norms(self.rows)

    def scores(self, clip, estimates: torch.Tensor) -> torch.Tensor:
        """[b, M_pad] scores of a query batch (columns >= M belong to the zero padding)."""
        est, _ = clip._prepare(estimates, estimates[:1])
        est = est.to(self.rows.device, non_blocking=True).reshape(est.shape[0], -1)
        assert est.shape[1] == self.row_elems, "estimates and candidates disagree on F*T"
        return BF.clip_scores_prenormed(est, self.rows, self.inv_norms)


def _i64(t: torch.Tensor, device) -> torch.Tensor:
    return t.to(device=device, dtype=torch.int64).contiguous()


def _topk(vals, n_cols, k, labels=None, targets=None, own_values=None, own_col=0, own_labels=None, is_prob=False,
          want_soft=False, want_stats=False):
    """One `bm_retrieval_topk` launch over the rows of `vals` [Bn, ld]."""
    Bn, ld = vals.shape
    dev = vals.device
    top_idx = torch.empty(Bn, k, device=dev, dtype=torch.int64)
    top_prob = torch.empty(Bn, k, device=dev, dtype=torch.float32)
    match = labels is not None and targets is not None
    hit = torch.empty(Bn, device=dev, dtype=torch.int32) if match else None
    soft = torch.empty(Bn, device=dev, dtype=torch.float32) if (match and want_soft) else None
    row_max = torch.empty(Bn, device=dev, dtype=torch.float32) if want_stats else None
    row_sum = torch.empty(Bn, device=dev, dtype=torch.float32) if want_stats else None
    call("bm_retrieval_topk", ptr(vals), ld, Bn, n_cols, ptr(own_values), own_col, int(is_prob), k, ptr(labels),
         ptr(own_labels), ptr(targets), ptr(top_idx), ptr(top_prob), ptr(hit), ptr(soft), ptr(row_max), ptr(row_sum),
         stream())
    return dict(top_idx=top_idx, top_prob=top_prob, hit=hit, soft=soft, row_max=row_max, row_sum=row_sum)


# ------------------------------------------------------------------------------------------------------------------
# scripts/run_eval_probs.py
# ------------------------------------------------------------------------------------------------------------------
def builds_probs(clip, preds: torch.Tensor, trues: torch.Tensor, dset_args, batch_size: int = 100,
                 tmin: tp.Optional[float] = None, tmax: tp.Optional[float] = None) -> torch.Tensor:
    """Probability of every candidate segment for every prediction: [len(preds), len(trues)], on the host like the
    reference's (run_eval_probs.py:296-307)."""
    lo, hi = _window(dset_args, tmin, tmax)
    preds, trues = preds[..., lo:hi], trues[..., lo:hi]
    bank = CandidateBank(clip, trues)
    probs = torch.zeros(len(preds), len(trues))
    for i in range(0, len(preds), batch_size):
        scores = bank.scores(clip, preds[i:i + batch_size])
        out = torch.empty(scores.shape[0], bank.n, device=scores.device, dtype=torch.float32)
        call("bm_retrieval_probs", ptr(scores), scores.shape[1], scores.shape[0], bank.n, ptr(out), stream())
        probs[i:i + batch_size] = out.cpu()
    return probs


def _get_accuracy_from_probs(probs: torch.Tensor, target_labels: torch.Tensor, vocab_labels: torch.Tensor,
                             topk: int = 10, batch_size: int = 4096) -> float:
    """Top-k accuracy of probability rows [B, V] against labels (run_eval_probs.py:237-264)."""
    assert len(target_labels) == len(probs)
    assert len(vocab_labels) == probs.shape[1]
    dev = probs.device if probs.is_cuda else _device()
    labels = _i64(vocab_labels, dev)
    hits = 0
    for i in range(0, len(probs), batch_size):
        p = probs[i:i + batch_size].to(dev, dtype=torch.float32).contiguous()
        r = _topk(p, p.shape[1], topk, labels, _i64(target_labels[i:i + batch_size], dev), is_prob=True)
        hits += int((r["hit"] >= 0).sum())
    return hits / len(probs)


def retrieval_accuracy(clip, preds: torch.Tensor, trues: torch.Tensor, target_labels: torch.Tensor,
                       vocab_labels: torch.Tensor, topk: tp.Sequence[int] = (1, 5, 10), batch_size: int = 1024,
                       dset_args=None, tmin: tp.Optional[float] = None, tmax: tp.Optional[float] = None,
                       bank: tp.Optional[CandidateBank] = None) -> tp.Dict[int, float]:
    """`_get_accuracy_from_probs(builds_probs(...), ...)` for several k at once (run_eval_probs.py:331-362) without
    materialising the probabilities: {k: accuracy}."""
    assert len(target_labels) == len(preds)
    if tmin is not None or tmax is not None:
        lo, hi = _window(dset_args, tmin, tmax)
        preds = preds[..., lo:hi]
        trues = None if trues is None else trues[..., lo:hi]
    bank = bank or CandidateBank(clip, trues)          # `bank`: candidates already resident (cropped like `preds`)
    assert len(vocab_labels) == bank.n
    dev = bank.rows.device
    labels = _i64(vocab_labels, dev)
    kmax = max(topk)
    ranks = []
    for i in range(0, len(preds), batch_size):
        scores = bank.scores(clip, preds[i:i + batch_size])
        r = _topk(scores, bank.n, kmax, labels, _i64(target_labels[i:i + batch_size], dev))
        ranks.append(r["hit"])
    ranks = torch.cat(ranks).cpu()
    return {int(k): float(((ranks >= 0) & (ranks < k)).float().mean()) for k in topk}


# ------------------------------------------------------------------------------------------------------------------
# bm/wer.py
# ------------------------------------------------------------------------------------------------------------------
def _vocabulary(shared_hashes: torch.Tensor, word_hashes: torch.Tensor):
    """Bookkeeping for the per-word probabilities (wer.py:101-104), done once instead of once per estimate.

    shared_hashes: hashes of the negatives every estimate shares (all but the overwritten last slot).
    Returns (vocab [V] sorted distinct hashes, order [n_shared] the shared columns grouped by word -- column order kept
    inside a word --, seg [V+1] group offsets, own_word [n] index of each estimate's own word in vocab, V if absent)."""
    shared = shared_hashes.to(torch.int64).cpu()
    vocab, inverse = torch.unique(shared, return_inverse=True)
    order = torch.argsort(inverse, stable=True)
    seg = torch.zeros(len(vocab) + 1, dtype=torch.int64)
    seg[1:] = torch.cumsum(torch.bincount(inverse, minlength=len(vocab)), 0)
    wh = word_hashes.to(torch.int64).cpu()
    slot = torch.searchsorted(vocab, wh).clamp_(max=len(vocab) - 1)
    own_word = torch.where(vocab[slot] == wh, slot, torch.full_like(slot, len(vocab)))
    return vocab, order, seg, own_word


def wer_ranking(clip, estimates: torch.Tensor, word_hashes: torch.Tensor, outputs: torch.Tensor,
                negatives: torch.Tensor, negative_hashes: torch.Tensor, topx: int, batch_size: int = 1024,
                wer_random: bool = False) -> tp.Dict[str, float]:
    """The ranking loop of get_wer (wer.py:80-116) for all estimates at once.

    For estimate i the reference ranks it against `negatives` with the LAST slot replaced by `outputs[i]` (hash
    `word_hashes[i]`).  Returns {'wer', 'wer_vocab', 'soft_correct'} (wer.py:117-120 + the soft count of :114-115)."""
    n, n_neg = len(estimates), len(negatives)
    assert n_neg >= 2 and len(outputs) == n and len(word_hashes) == n and len(negative_hashes) == n_neg
    bank = CandidateBank(clip, negatives)
    dev = bank.rows.device
    own_col = n_neg - 1
    neg_hashes = _i64(negative_hashes, dev)

    vocab, order, seg, own_word_all = _vocabulary(negative_hashes[:own_col], word_hashes)
    V = len(vocab)
    perm_d = order.to(dev, dtype=torch.int32)
    seg_d = seg.to(dev, dtype=torch.int32)
    vocab_labels = torch.cat([vocab, vocab.new_zeros(1)]).to(dev)           # column V's label comes from own_labels
    wh_all = word_hashes.to(torch.int64).cpu()

    miss = miss_vocab = 0
    soft_sum = 0.0
    for i in range(0, n, batch_size):
        est = estimates[i:i + batch_size]
        if wer_random:                                                       # wer.py:95-96
            est = torch.randn_like(est)
        out = outputs[i:i + batch_size]
        est_p, out_p = clip._prepare(est, out)
        est_d = est_p.to(dev, non_blocking=True).reshape(len(est), -1).contiguous().float()
        out_d = out_p.to(dev, non_blocking=True).reshape(len(est), -1).contiguous().float()
        wh = wh_all[i:i + batch_size].to(dev)
        own_word = own_word_all[i:i + batch_size].to(dev, dtype=torch.int32)

        scores = BF.clip_scores_prenormed(est_d, bank.rows, bank.inv_norms)
        own = torch.empty(len(est), device=dev, dtype=torch.float32)
        call("bm_rowdot_scaled", ptr(est_d), ptr(out_d), len(est), est_d.shape[1], ptr(own), stream())
        r = _topk(scores, n_neg, topx, neg_hashes, wh, own_values=own, own_col=own_col, own_labels=wh,
                  want_soft=True, want_stats=True)
        vocab_p = torch.empty(len(est), V + 1, device=dev, dtype=torch.float32)
        call("bm_retrieval_vocab_probs", ptr(scores), scores.shape[1], len(est), ptr(own), ptr(r["row_max"]),
             ptr(r["row_sum"]), ptr(perm_d), ptr(seg_d), V, ptr(own_word), ptr(vocab_p), stream())
        rv = _topk(vocab_p, V + 1, topx, vocab_labels, wh, own_col=V, own_labels=wh, is_prob=True)
        miss += int((r["hit"] < 0).sum())
        miss_vocab += int((rv["hit"] < 0).sum())
        soft_sum += float(r["soft"].double().sum())
    return dict(wer=miss / n, wer_vocab=miss_vocab / n, soft_correct=soft_sum / n)
