"""TEST INFRASTRUCTURE ONLY -- CPU restatement of brainmagick's retrieval evaluation.

    scripts/run_eval_probs.py:267-307   builds_probs: probs[i] = softmax_j <pred_i, true_j / ||true_j||>, in query
                                        batches, after an optional [tmin, tmax] crop
    scripts/run_eval_probs.py:237-264   _get_accuracy_from_probs: hit if the target label is among the labels of
                                        the top-k columns of the row; mean over rows
    bm/wer.py:80-116                    the ranking loop of get_wer: for each estimate the LAST negative is replaced by
                                        the estimate's own true output, probabilities over the negatives, the same
                                        probabilities summed per distinct word hash, top-x hit on both, "soft" hit
The scores / probabilities themselves are ClipLoss.get_probabilities (bm/losses.py:77-102), restated in
`oracle/bm_oracle.py` and pinned against the verbatim reference there; `tests/golden/retrieval_small.npz` pins this
file against the verbatim `ClipLoss` driven by the loops above (see `oracle/make_golden.py`).

Only `tests/`, `__graft_entry__.smoke()` and the CPU legs of `bench.py` may import this module.
"""
from __future__ import annotations

import typing as tp

import torch

from . import bm_oracle


def crop_window(dset_tmin: float, sample_rate: float, tmin, tmax) -> tp.Tuple[tp.Optional[int], tp.Optional[int]]:
    """run_eval_probs.py:279-290."""
    lo = None if tmin is None else int((tmin - dset_tmin) * sample_rate)
    hi = None if tmax is None else int((tmax - dset_tmin) * sample_rate)
    return lo, hi


def builds_probs(preds: torch.Tensor, trues: torch.Tensor, batch_size: int = 100, window=(None, None),
                 probabilities=bm_oracle.clip_probabilities) -> torch.Tensor:
    """run_eval_probs.py:267-307 -> probs [len(preds), len(trues)]."""
    lo, hi = window
    preds, trues = preds[..., lo:hi], trues[..., lo:hi]
    rows = [probabilities(preds[i:i + batch_size], trues) for i in range(0, len(preds), batch_size)]
    return torch.cat(rows, dim=0)


def accuracy_from_probs(probs: torch.Tensor, target_labels: torch.Tensor, vocab_labels: torch.Tensor,
                        topk: int = 10) -> float:
    """run_eval_probs.py:237-264."""
    assert len(target_labels) == len(probs) and len(vocab_labels) == probs.shape[1]
    best = probs.topk(topk, dim=1).indices
    hits = (vocab_labels[best] == target_labels[:, None]).any(dim=1)
    return hits.float().mean().item()


def wer_ranking(estimates: torch.Tensor, word_hashes: torch.Tensor, outputs: torch.Tensor,
                negatives: torch.Tensor, negative_hashes: torch.Tensor, topx: int,
                probabilities=bm_oracle.clip_probabilities) -> tp.Dict[str, float]:
    """bm/wer.py:80-116, one estimate at a time like the reference (the last negative slot is overwritten)."""
    negatives = negatives.clone()
    negative_hashes = negative_hashes.clone()
    hit = hit_vocab = soft = 0.0
    for est, wh, out in zip(estimates, word_hashes, outputs):
        negatives[-1] = out
        negative_hashes[-1] = wh
        p = probabilities(est[None], negatives)[0]
        vocab, inverse = torch.unique(negative_hashes, return_inverse=True)
        p_vocab = torch.zeros(len(vocab), dtype=p.dtype).scatter_add_(0, inverse, p)
        best = p.topk(topx).indices
        best_vocab = p_vocab.topk(topx).indices
        hit += float((negative_hashes[best] == wh).any())
        hit_vocab += float((vocab[best_vocab] == wh).any())
        soft += float(p[negative_hashes == wh].sum())
    n = len(estimates)
    return dict(wer=1 - hit / n, wer_vocab=1 - hit_vocab / n, soft_correct=soft / n)
