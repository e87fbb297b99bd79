"""Multi-GPU plumbing (one process per GPU, torch.distributed/NCCL over NVLink).

The path shards by batch row: each rank encodes its rows (BatchNorm statistics stay per-rank,
like the reference which never converts to SyncBatchNorm), then
  * `all_gather_candidates`: the ONE forward exchange -- NCCL all-gather of the candidate block so every rank
    scores its rows against the global batch (new semantic; the reference keeps negatives local, README.md:139-143);
  * `all_gather_candidates_with_grad`: the same exchange when a trainable feature model produced the candidates --
    its backward is the reduce-scatter(sum) of the candidate gradients;
  * `sync_gradients`: the reference's gradient all-reduce(avg) (flashy.distrib.sync_model, bm/solver.py:386),
    done on one flat bucket.
"""
from __future__ import annotations

import typing as tp

import torch
import torch.distributed as dist


def world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def all_gather_candidates(candidate: torch.Tensor) -> tp.Tuple[torch.Tensor, int]:
    """[B_loc, F, T] per rank -> ([W*B_loc, F, T], target offset of this rank's rows).  For candidates that carry no
    gradient (no feature_model); see `all_gather_candidates_with_grad` otherwise."""
    W = world_size()
    if W == 1:
        return candidate, 0
    candidate = candidate.contiguous()
    out = torch.empty((W * candidate.shape[0],) + tuple(candidate.shape[1:]), dtype=candidate.dtype,
                      device=candidate.device)
    dist.all_gather_into_tensor(out, candidate)
    return out, rank() * candidate.shape[0]


class _GatherWithGrad(torch.autograd.Function):
    """All-gather whose backward is the reduce-scatter(sum) SURVEY.md 8(e) asks for when the candidates carry a gradient
    (a trainable feature model, bm/solver.py:304-320): rank r's block of the gathered tensor feeds EVERY rank's loss term,
    so its gradient is the sum over ranks of that block of their gradients.  (Parameter gradients are then averaged by
    `sync_gradients` / flashy's sync_model like all others, which yields the gradient of the mean of the per-rank losses.)"""

    @staticmethod
    def forward(ctx, candidate):
        candidate = candidate.contiguous()
        W = world_size()
        out = torch.empty((W * candidate.shape[0],) + tuple(candidate.shape[1:]), dtype=candidate.dtype,
                          device=candidate.device)
        dist.all_gather_into_tensor(out, candidate)
        ctx.rows = candidate.shape[0]
        return out

    @staticmethod
    def backward(ctx, dgathered):
        dgathered = dgathered.contiguous()
        n, r = ctx.rows, rank()
        if dist.get_backend() == "nccl":
            own = torch.empty((n,) + tuple(dgathered.shape[1:]), dtype=dgathered.dtype, device=dgathered.device)
            dist.reduce_scatter_tensor(own, dgathered, op=dist.ReduceOp.SUM)
            return own
        total = dgathered.clone()              # gloo (CPU tests) has no reduce-scatter
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
        return total[r * n:(r + 1) * n].clone()


def all_gather_candidates_with_grad(candidate: torch.Tensor) -> tp.Tuple[torch.Tensor, int]:
    """`all_gather_candidates` for candidates that require grad: differentiable (reduce-scatter in backward)."""
    if world_size() == 1:
        return candidate, 0
    return _GatherWithGrad.apply(candidate), rank() * candidate.shape[0]


class CandidateGather:
    """The candidate all-gather started EARLY (it does not depend on the encoder): NCCL moves the blocks over NVLink while
    the encoder's forward kernels run; `wait()` joins it on the current stream just before the contrastive matmul."""

    def __init__(self, candidate: torch.Tensor):
        self.source = candidate
        W = world_size()
        self.offset = rank() * candidate.shape[0]
        if W == 1:
            self.out, self.work = candidate, None
            return
        candidate = candidate.contiguous()
        self.out = torch.empty((W * candidate.shape[0],) + tuple(candidate.shape[1:]), dtype=candidate.dtype,
                               device=candidate.device)
        self.work = dist.all_gather_into_tensor(self.out, candidate, async_op=True)

    def wait(self) -> tp.Tuple[torch.Tensor, int]:
        if self.work is not None:
            self.work.wait()          # current stream waits for NCCL's stream; the host does not block
            self.work = None
        return self.out, self.offset


def sync_gradients(params: tp.Iterable[torch.nn.Parameter]) -> None:
    """all-reduce(avg) of every .grad through one flat fp32 bucket."""
    W = world_size()
    if W == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(W)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
