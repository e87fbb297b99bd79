"""CPU, world_size 2, gloo: the N>1 host logic (candidate all-gather + target offsets, gradient all-reduce) and the
multi-GPU oracle of SURVEY.md 8(e): per-rank encoder (local BatchNorm statistics) + global negatives."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from brainmagick_b200 import distrib
        torch.manual_seed(100 + rank)
        cand = torch.randn(3, 4, 5)
        gathered, off = distrib.all_gather_candidates(cand)
        assert gathered.shape == (3 * world, 4, 5) and off == 3 * rank
        assert torch.equal(gathered[off:off + 3], cand)
        # gradient all-reduce(avg) over a flat bucket
        p1, p2 = torch.nn.Parameter(torch.zeros(4, 3)), torch.nn.Parameter(torch.zeros(7))
        p1.grad = torch.full((4, 3), float(rank + 1))
        p2.grad = torch.arange(7.) * (rank + 1)
        distrib.sync_gradients([p1, p2])
        mean_scale = sum(r + 1 for r in range(world)) / world
        assert torch.allclose(p1.grad, torch.full((4, 3), mean_scale))
        assert torch.allclose(p2.grad, torch.arange(7.) * mean_scale)
        early = distrib.CandidateGather(cand)              # the prefetched (async) form used by ClipLoss.prefetch_candidates
        g2, off2 = early.wait()
        assert torch.equal(g2, gathered) and off2 == off
        ret[rank] = (gathered.clone(), off)
    finally:
        dist.destroy_process_group()


def test_gloo_allgather_and_grad_sync():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        g0, off0 = ret[0]
        g1, off1 = ret[1]
        assert torch.equal(g0, g1) and (off0, off1) == (0, 3)


def test_multi_gpu_oracle_decomposes_into_per_rank_losses():
    """global loss = mean over ranks of CE(scores_r[B/W, B], targets r*B/W + arange); with ONE
    rank it equals the reference loss."""
    from oracle import bm_oracle
    torch.manual_seed(0)
    est = torch.randn(8, 5, 12)
    cand = torch.randn(8, 5, 12)
    ref = bm_oracle.clip_loss(est, cand)
    parts = [bm_oracle.clip_loss(est[r * 4:(r + 1) * 4], cand, target_offset=r * 4) for r in range(2)]
    assert abs(float(sum(parts) / 2) - float(ref)) < 1e-6


def _grad_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from brainmagick_b200 import distrib
        from oracle import bm_oracle
        torch.manual_seed(7)                      # every rank builds the SAME global tensors and takes its shard
        est = torch.randn(world * 3, 4, 6)
        cand = torch.randn(world * 3, 4, 6)
        mine = slice(rank * 3, rank * 3 + 3)
        c = cand[mine].clone().requires_grad_(True)
        e = est[mine].clone().requires_grad_(True)
        gathered, off = distrib.all_gather_candidates_with_grad(c)
        assert off == 3 * rank and gathered.requires_grad and torch.equal(gathered.detach(), cand)
        loss = bm_oracle.clip_loss(e, gathered, target_offset=off)       # this rank's term of the global loss
        loss.backward()
        ret[rank] = (float(loss), c.grad.clone(), e.grad.clone())
    finally:
        dist.destroy_process_group()


def test_gather_with_grad_reduce_scatters_candidate_gradients():
    """SURVEY 8(e) with trainable candidates: per-rank losses over all-gathered candidates; each rank's candidate gradient
    must be the sum over ranks of its block == W x the gradient of the global (mean-over-ranks) loss, which the averaging
    gradient all-reduce of the feature model's parameters then turns into the global gradient."""
    from oracle import bm_oracle
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_grad_worker, args=(world, port, ret), nprocs=world, join=True)
        torch.manual_seed(7)
        est = torch.randn(world * 3, 4, 6).requires_grad_(True)
        cand = torch.randn(world * 3, 4, 6).requires_grad_(True)
        parts = [bm_oracle.clip_loss(est[r * 3:(r + 1) * 3], cand, target_offset=r * 3) for r in range(world)]
        total = sum(parts) / world
        total.backward()
        for r in range(world):
            loss_r, dc_r, de_r = ret[r]
            assert abs(loss_r - float(parts[r])) < 1e-6
            assert torch.allclose(dc_r, world * cand.grad[r * 3:(r + 1) * 3], atol=1e-6)
            assert torch.allclose(de_r, world * est.grad[r * 3:(r + 1) * 3], atol=1e-6)


def _module_worker(rank, world, port, ret):
    """ClipLoss(global_negatives=True) itself on two gloo ranks, its CUDA entry points served by the CPU emulator."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import abi_emulator
        import brainmagick_b200 as bb
        torch.manual_seed(11)
        est = torch.randn(world * 3, 4, 6)
        cand = torch.randn(world * 3, 4, 6)
        mine = slice(rank * 3, rank * 3 + 3)
        mask = torch.ones(3, 1, 6, dtype=torch.bool)
        out = {}
        with abi_emulator.emulated():
            clip = bb.ClipLoss(global_negatives=True)
            # (1) constant candidates, gather started early (prefetch) and picked up by forward
            e = est[mine].clone().requires_grad_(True)
            c = cand[mine].clone()
            clip.prefetch_candidates(c)
            assert clip._prefetched is not None
            loss = clip(e, c, mask)
            loss.backward()
            out["plain"] = (float(loss), e.grad.clone())
            # (2) same without the prefetch
            e2 = est[mine].clone().requires_grad_(True)
            loss2 = clip(e2, cand[mine].clone(), mask)
            loss2.backward()
            assert abs(float(loss2) - float(loss)) < 1e-6 and torch.allclose(e2.grad, e.grad, atol=1e-6)
            # (3) trainable candidates: differentiable gather, no prefetch
            e3 = est[mine].clone().requires_grad_(True)
            c3 = cand[mine].clone().requires_grad_(True)
            clip.prefetch_candidates(c3)
            assert clip._prefetched is None
            loss3 = clip(e3, c3, mask)
            loss3.backward()
            out["trainable"] = (float(loss3), e3.grad.clone(), c3.grad.clone())
        ret[rank] = out
    finally:
        dist.destroy_process_group()


def test_cliploss_module_with_global_negatives_on_two_ranks():
    from oracle import bm_oracle
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_module_worker, args=(world, port, ret), nprocs=world, join=True)
        torch.manual_seed(11)
        est = torch.randn(world * 3, 4, 6).requires_grad_(True)
        cand = torch.randn(world * 3, 4, 6).requires_grad_(True)
        parts = [bm_oracle.clip_loss(est[r * 3:(r + 1) * 3], cand, target_offset=r * 3) for r in range(world)]
        (sum(parts) / world).backward()
        for r in range(world):
            loss_r, de_r = ret[r]["plain"]
            assert abs(loss_r - float(parts[r])) < 1e-5
            assert torch.allclose(de_r, world * est.grad[r * 3:(r + 1) * 3], atol=1e-6)
            loss_t, de_t, dc_t = ret[r]["trainable"]
            assert abs(loss_t - float(parts[r])) < 1e-5
            assert torch.allclose(de_t, world * est.grad[r * 3:(r + 1) * 3], atol=1e-6)
            assert torch.allclose(dc_t, world * cand.grad[r * 3:(r + 1) * 3], atol=1e-6)


def _ragged_worker(rank, world, port, ret):
    """Per-rank batch sizes differ (no drop_last, per-rank rejection): rank r holds 3 - r rows."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import abi_emulator
        import brainmagick_b200 as bb
        from brainmagick_b200 import distrib
        counts = [3 - r for r in range(world)]
        starts = [sum(counts[:r]) for r in range(world)]
        torch.manual_seed(13)
        est = torch.randn(sum(counts), 4, 6)
        cand = torch.randn(sum(counts), 4, 6)
        mine = slice(starts[rank], starts[rank] + counts[rank])
        assert distrib.gather_counts(counts[rank], torch.device("cpu")) == counts
        g, off = distrib.all_gather_candidates(cand[mine].clone())
        assert off == starts[rank] and torch.equal(g, cand)
        pre = distrib.CandidateGather(cand[mine].clone())
        g2, off2 = pre.wait()
        assert off2 == off and torch.equal(g2, cand)
        out = {}
        mask = torch.ones(counts[rank], 1, 6, dtype=torch.bool)
        with abi_emulator.emulated():
            clip = bb.ClipLoss(global_negatives=True)
            e = est[mine].clone().requires_grad_(True)
            c = cand[mine].clone().requires_grad_(True)
            loss = clip(e, c, mask)
            loss.backward()
            out = (float(loss), e.grad.clone(), c.grad.clone())
        ret[rank] = out
    finally:
        dist.destroy_process_group()


def test_ragged_per_rank_batches():
    """ADVICE r1 (distrib.py): the gathers must not assume equal per-rank row counts."""
    from oracle import bm_oracle
    world = 2
    counts = [3 - r for r in range(world)]
    starts = [sum(counts[:r]) for r in range(world)]
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_ragged_worker, args=(world, port, ret), nprocs=world, join=True)
        torch.manual_seed(13)
        est = torch.randn(sum(counts), 4, 6).requires_grad_(True)
        cand = torch.randn(sum(counts), 4, 6).requires_grad_(True)
        parts = [bm_oracle.clip_loss(est[starts[r]:starts[r] + counts[r]], cand, target_offset=starts[r]) for r in range(world)]
        (sum(parts) / world).backward()
        for r in range(world):
            loss_r, de_r, dc_r = ret[r]
            sl = slice(starts[r], starts[r] + counts[r])
            assert abs(loss_r - float(parts[r])) < 1e-5
            assert torch.allclose(de_r, world * est.grad[sl], atol=1e-6)
            assert torch.allclose(dc_r, world * cand.grad[sl], atol=1e-6)
