#!/usr/bin/env python
"""Benchmark of the contrastive training step (BASELINE.json metric: 3 s-segments/sec).

    python bench.py --gpus N --steps K --warmup W                 # B200 arm (this repo's CUDA path)
    python bench.py --impl reference --gpus N --steps K --warmup W  # reference arm: the CPU path on host cores

One "step" = SimpleConv forward + ClipLoss forward + backward of both + Adam(lr 3e-4) update on one synthetic
batch (bm/solver.py:297,373,384-387; bm/train.py:119); for N > 1 the candidates are all-gathered before the
contrastive matmul and the gradients all-reduced.  Workload at N=1 = BASELINE.json configs[1]:
synthetic gwilliams2022-like MEG (208 sensors, 3 s @ 120 Hz = 360 samples, 1024-d wav2vec-like features, 27
subjects), B = 256 per GPU.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOAD = dict(name="cfg2 synthetic gwilliams2022 (208 sensors, 3s@120Hz, F=1024, S=27)", C=208, T=360, F=1024, S=27)
CLIP_CONV = dict(hidden=dict(meg=320), batch_norm=True, depth=10, dilation_period=5, kernel_size=3, skip=True,
                 subject_layers=True, subject_dim=0, complex_out=True, glu=2, glu_context=1, merger=True,
                 initial_linear=270, gelu=True, merger_pos_dim=2048)       # conf/model/clip_conv.yaml:6-22


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"],
                    bf16_tflops_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


# ------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------------------
# synthetic workload: generated on CPU, pinned, shared by the e2e and resident runs
# ------------------------------------------------------------------------------------------------------
def make_host_batch(B, seed):
    g = torch.Generator().manual_seed(seed)
    C, T, F, S = WORKLOAD["C"], WORKLOAD["T"], WORKLOAD["F"], WORKLOAD["S"]
    meg = torch.randn(B, C, T, generator=g).clamp_(-20, 20)
    feats = torch.randn(B, F, T, generator=g)
    subj = torch.randint(0, S, (B,), generator=g)
    return meg, feats, subj


def run_b200(args):
    import torch.distributed as dist
    import brainmagick_b200 as bb
    from brainmagick_b200 import _lib, distrib, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback for the B200 arm)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    _lib.load()

    B = args.batch
    C, T, F, S = WORKLOAD["C"], WORKLOAD["T"], WORKLOAD["F"], WORKLOAD["S"]
    torch.manual_seed(2036)                                             # conf/config.yaml:33
    model = bb.SimpleConv(in_channels=dict(meg=C), out_channels=F, n_subjects=S,
                          **{k: (dict(v) if isinstance(v, dict) else v) for k, v in CLIP_CONV.items()}).to(dev)
    clip = bb.ClipLoss(global_negatives=world > 1).to(dev)
    model.train()
    clip.train()
    opt = torch.optim.Adam(model.parameters(), lr=3e-4, betas=(0.9, 0.999), fused=True)   # bm/train.py:119
    positions = synthetic.normalised_positions(S, C, seed=7)
    mask = torch.ones(B, 1, T, dtype=torch.bool, device=dev)

    n_host = 2                                                          # rotate two different host batches
    host = []
    for i in range(n_host):
        meg, feats, subj = make_host_batch(B, 2036 + 1000 * rank + i)
        host.append((meg.pin_memory(), feats.pin_memory(), subj.pin_memory(), subj.tolist()))
    resident = [tuple(t.to(dev) for t in hb[:3]) for hb in host]
    recs = [synthetic.SyntheticRecording(s, positions[s]) for s in range(S)]

    def make_batch(meg_d, subj_d, subj_h):
        return synthetic.SyntheticBatch(meg_d, subj_d, [recs[s] for s in subj_h])

    def step(meg_d, feats_d, subj_d, subj_h):
        opt.zero_grad(set_to_none=True)
        batch = make_batch(meg_d, subj_d, subj_h)
        clip.prefetch_candidates(feats_d)          # N > 1: the candidate all-gather overlaps the encoder forward
        est = model(dict(meg=meg_d), batch)
        loss = clip(est, feats_d, mask)
        loss.backward()
        if world > 1:
            distrib.sync_gradients(model.parameters())
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n_steps, fn):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n_steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- device-resident throughput (`value`) ---------------------------------------------------------
    def resident_step(i):
        meg_d, feats_d, subj_d = resident[i % n_host]
        step(meg_d, feats_d, subj_d, host[i % n_host][3])

    for i in range(args.warmup):
        resident_step(i)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    ms_total = timed(args.steps, resident_step)
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms_total / args.steps
    value = world * B / (ms_per_step / 1e3)

    # ---- end to end through the public API with HOST buffers (`e2e`) ------------------------------------
    # every step copies its inputs from pinned host memory (on a copy stream, one step ahead, like a DataLoader that
    # prefetches to the device) and reads the loss back to the host; all of it inside the timed region.
    last_loss = [0.0]
    copy_stream = torch.cuda.Stream(device=dev)
    slots = [None, None]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def issue_copy(i):
        meg_h, feats_h, subj_h, subj_l = host[i % n_host]
        k = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[k])                         # the step that used this slot has finished
            slots[k] = (meg_h.to(dev, non_blocking=True), feats_h.to(dev, non_blocking=True),
                        subj_h.to(dev, non_blocking=True), subj_l)
            ready[k].record(copy_stream)

    pending = [None]

    def e2e_run(n_steps):
        main = torch.cuda.current_stream()
        for k in range(2):
            consumed[k].record(main)
        issue_copy(0)
        for i in range(n_steps):
            if i + 1 < n_steps:
                issue_copy(i + 1)
            k = i % 2
            main.wait_event(ready[k])
            meg_d, feats_d, subj_d, subj_l = slots[k]
            loss = step(meg_d, feats_d, subj_d, subj_l)
            consumed[k].record(main)
            # device -> host read of a step's result EVERY step, pipelined by one step (the host reads step i-1's loss
            # while step i is already queued, like a training loop that logs the previous iteration's loss)
            if pending[0] is not None:
                last_loss[0] = pending[0].item()
            pending[0] = loss
        last_loss[0] = pending[0].item()
        pending[0] = None

    e2e_run(min(2, args.warmup))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_run(args.steps)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1) / args.steps
    if world > 1:
        tt = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_e2e = float(tt.item())
    h2d = sum(t.numel() * t.element_size() for t in host[0][:3])
    e2e = dict(value=world * B / (ms_e2e / 1e3), unit="segments/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=4,
               ms_per_step=ms_e2e, note="inputs copied from pinned host memory on a copy stream one step ahead; one loss.item() per step, read one step behind")

    # ---- roofline of the dominant kernel: the K3 dilated conv (320 -> 320, k=3) -------------------------
    roofline = None
    cpu_baseline = None
    also = None
    if rank == 0:
        roofline = conv_roofline(dev, B, T)
        if world == 1:
            also = also_measured(model, clip, make_batch, resident, host, n_host, mask, dev, B)
        if not args.no_cpu_baseline:
            cpu_baseline = run_cpu(steps=2, warmup=1, batch=16, threads=None)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    from brainmagick_b200 import functional as BF
    BF.check_tc_status()
    if rank != 0:
        return
    out = dict(
        metric="3s-segments/sec (training step: SimpleConv fwd + ClipLoss + bwd + Adam)", value=value,
        unit="segments/s", n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms_per_step,
        higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
        config=dict(workload=WORKLOAD["name"], batch_per_gpu=B, global_batch=B * world, sensors=C, T=T, F=F,
                    subjects=S, model="clip_conv (random init)", negatives="global (all-gather)" if world > 1 else "local",
                    l2="per-step working set (inputs 454 MB + ~5 GB saved activations) >> 126 MB L2; two input batches rotate",
                    last_loss=last_loss[0]),
        clocks=clocks, e2e=e2e, gpu_launches=int(launches), roofline=roofline, cpu_baseline=cpu_baseline, also=also)
    print(json.dumps(out))


def also_measured(model, clip, make_batch, resident, host, n_host, mask, dev, B, iters=5):
    """What SURVEY.md 8(d) asks to report beside the headline (N=1 only, after the headline has been timed; every part is
    optional and a failure is recorded instead of raised):
      * forward_only: the encoder in eval mode under no_grad (the evaluation-time cost), segments/s;
      * torch_eager_gpu: the SAME step (forward + ClipLoss + backward, no optimizer) written in plain PyTorch ops -- the
        oracle restatement of the reference modules -- run on this GPU with cuDNN / cuBLAS: the "library-kernel" bar,
        with TF32 off (fp32-faithful like this repo) and on."""
    out = {}

    def timed_ms(fn, n):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    try:
        model.eval()

        def fwd(i=0):
            meg_d, _, subj_d = resident[i % n_host]
            with torch.no_grad():
                model(dict(meg=meg_d), make_batch(meg_d, subj_d, host[i % n_host][3]))
        ms = timed_ms(fwd, iters)
        out["forward_only"] = dict(value=B / (ms / 1e3), unit="segments/s", ms_per_batch=ms, mode="eval, no_grad")
    except Exception as exc:
        out["forward_only"] = dict(error=f"{type(exc).__name__}: {exc}")
    finally:
        model.train()

    try:
        from oracle import bm_oracle
        from brainmagick_b200 import synthetic
        C, T, F, S = WORKLOAD["C"], WORKLOAD["T"], WORKLOAD["F"], WORKLOAD["S"]
        cfg = bm_oracle.Config(in_channels=C, out_channels=F, n_subjects=S)
        params = {k: v.detach().clone() for k, v in model.state_dict().items()}
        meg_d, feats_d, subj_d = resident[0]
        pos = synthetic.normalised_positions(S, C, seed=7).to(dev)
        ban = torch.tensor([0.5, 0.5], device=dev)
        saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        res = {}
        try:
            for label, tf32 in (("tf32_off", False), ("tf32_on", True)):
                torch.backends.cudnn.allow_tf32 = tf32
                torch.backends.cuda.matmul.allow_tf32 = tf32

                def eager(i=0):
                    bm_oracle.training_step(params, cfg, meg_d, pos, subj_d, subj_d, feats_d, ban_centre=ban, training=True)
                ms = timed_ms(eager, 3)
                res[label] = dict(value=B / (ms / 1e3), unit="segments/s", ms_per_step=ms)
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
        res["what"] = "oracle restatement of SimpleConv + ClipLoss in PyTorch ops on this GPU: forward + loss + backward, no optimizer"
        out["torch_eager_gpu"] = res
    except Exception as exc:
        out["torch_eager_gpu"] = dict(error=f"{type(exc).__name__}: {exc}")
    torch.cuda.empty_cache()
    return out


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu capture
    (profiles/*_ncu_conv3_summary.csv, `ncu --set full` of profiles/profile_kernels.py conv); None if absent."""
    import csv
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_conv3_summary.csv")))
    if not files:
        return None
    try:
        rows = list(csv.reader(open(files[-1])))
        hdr, units, vals = rows[0], rows[1], rows[-1]
        tot = 0.0
        for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(name)
            scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[units[i]]
            tot += float(vals[i]) * scale
        return tot
    except Exception:
        return None


def conv_roofline(dev, B, T, H=320, Kw=3, iters=20):
    """Times the dominant kernel alone -- bm_tc_conv1d_pair, the tcgen05 3xTF32 implicit-GEMM conv (K3: 320->320, k=3) --
    with CUDA events on the launching stream, L2 flushed between launches.  Algorithmic work: 2*H*H*Kw*T FLOP per
    segment (221.2 MFLOP/seg); the tensor pipe executes 3x that (hi*hi + lo*hi + hi*lo)."""
    from brainmagick_b200._lib import call, ptr, stream
    peaks = load_peaks()
    x = torch.randn(B, T, H, device=dev)
    w = torch.randn(H, H, Kw, device=dev) * 0.03
    f = torch.empty(Kw, H, H, device=dev)
    call("bm_tc_weight_split", ptr(w), H, H, Kw, ptr(f), None, None, None, stream())
    bias = torch.zeros(H, device=dev)
    y = torch.empty(B, T, H, device=dev)
    stats = torch.empty(2 * H, device=dev, dtype=torch.float64)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    times = []
    for i in range(iters + 3):
        flush.zero_()                                                   # L2 flush between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call("bm_tc_conv1d_persistent", ptr(x), ptr(f), ptr(bias), 0, B, T, H, H, Kw, 4, 1, 0, 0, 0, ptr(y), None, None,
             ptr(stats), ptr(status), stream())
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            times.append(e0.elapsed_time(e1))
    assert int(status.item()) == 0, "tcgen05 conv reported a pipeline timeout"
    ms = statistics.mean(times)
    flops = 2.0 * H * H * Kw * T * B
    achieved = flops / (ms / 1e3) / 1e12
    peak = peaks["bf16_tflops"]
    return dict(kernel="conv_pp_kernel via bm_tc_conv1d_persistent (K3: Conv1d 320->320 k3 d4 + BatchNorm statistics, persistent CTA pairs, tcgen05 cta_group::2 kind::tf32, 3xTF32)",
                bound="tensor", achieved=achieved, peak=peak, unit="TFLOP/s", frac=achieved / peak,
                traffic=ncu_traffic_bytes(), algorithmic_bytes=2.0 * B * T * H * 4,
                ms_per_launch=ms, peak_source=peaks["source"] + " cuBLAS bf16 burst (MEASURED_PEAKS.json)",
                executed_tflops=3 * achieved, tf32_pipe_peak=peak / 2, frac_of_tf32_pipe_executed=3 * achieved / (peak / 2),
                frac_of_3xtf32_ceiling=achieved / (peak / 6),
                note="achieved = ALGORITHMIC fp32 FLOPs (2*320*320*3*T*B per launch) / time. fp32-faithful parity (1e-4) "
                     "needs 3 tf32 MMAs per product, and kind::tf32 runs at half the bf16 rate, so the ceiling of "
                     "`frac` for this scheme is 1/6; frac_of_tf32_pipe_executed is the tensor-pipe view")


# ------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's CPU path on the host cores
# ------------------------------------------------------------------------------------------------------
def run_cpu(steps, warmup, batch, threads):
    from oracle import bm_oracle
    if threads:
        torch.set_num_threads(threads)
    cfg = bm_oracle.Config(in_channels=WORKLOAD["C"], out_channels=WORKLOAD["F"], n_subjects=WORKLOAD["S"])
    params = bm_oracle.init_state_dict(cfg, seed=0)
    tr = bm_oracle.CpuTrainer(cfg, params)
    d = bm_oracle.synthetic_batch(cfg, batch=batch, T=WORKLOAD["T"], seed=2036)
    a = (d["meg"], d["rec_positions"], d["rec_of_sample"], d["subject_index"], d["candidates"], d["ban_centre"])
    for _ in range(warmup):
        tr.step(*a)
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.step(*a)
    dt = (time.perf_counter() - t0) / steps
    return dict(value=batch / dt, unit="segments/s", cores=torch.get_num_threads(), kind="port",
                sample=f"{steps} steps of B={batch} at the cfg2 shapes (oracle/bm_oracle.py CpuTrainer: torch CPU ops "
                       f"fwd+loss+bwd+Adam), host has {os.cpu_count()} logical cores", s_per_step=dt)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 32
    res = run_cpu(steps=args.steps, warmup=min(args.warmup, 2), batch=batch, threads=None)
    out = dict(
        impl="reference",
        metric="3s-segments/sec (training step: SimpleConv fwd + ClipLoss + bwd + Adam)", value=res["value"],
        unit="segments/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=res["s_per_step"] * 1e3,
        higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
        config=dict(workload=WORKLOAD["name"], batch_per_step=batch, sensors=WORKLOAD["C"], T=WORKLOAD["T"],
                    F=WORKLOAD["F"], subjects=WORKLOAD["S"], model="clip_conv (random init)",
                    note="reference is pure Python/PyTorch and its deps (mne, flashy, dora) are absent, so its CPU path "
                         "is timed through the oracle port (same torch CPU ops), all host threads, bounded sample"),
        cpu_baseline=res,
        e2e=dict(value=res["value"], unit="segments/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="batch per GPU (256 = BASELINE.json configs[1])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--next-rows", action="store_true",
                    help="instead of the headline step: the SURVEY 8(f) rows (batch preparation, retrieval evaluation, "
                         "DeepMel), each beside a bounded CPU sample of its oracle; one GPU")
    args = ap.parse_args()
    if args.next_rows:
        from profiles import bench_next_rows      # its CPU legs are this file's cpu_baseline leg for those rows
        bench_next_rows.main()
        return
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
