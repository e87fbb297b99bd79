"""Times the ClipLoss forward (scores + norms + CE) and backward (dE) kernels at the N=1 (256 x 256) and N=8-per-rank
(256 x 2048) shapes of BASELINE cfg2 on ONE GPU, and prints the clip kernel's per-role cycle counters."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from brainmagick_b200 import _lib, functional as BF
from brainmagick_b200._lib import call, ptr, stream
dev = "cuda"; KT = 1024 * 360
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
def timeit(fn, n=8):
    ts = []
    for i in range(n + 2):
        flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(e0.elapsed_time(e1))
    return sum(ts) / len(ts)
for Bn, Bc in ((256, 256), (256, 2048), (64, 512), (128, 1024)):
    est = (torch.randn(Bn, KT, device=dev) * 0.01).requires_grad_(True)
    cand = torch.randn(Bc, KT, device=dev)
    loss = BF.clip_loss(est, cand, 0)
    ms_f = timeit(lambda: BF.clip_loss(est, cand, 0))
    def fb():
        l = BF.clip_loss(est, cand, 0); l.backward(); est.grad = None
    ms_fb = timeit(fb)
    fl = 2.0 * Bn * Bc * KT
    print(f"[{Bn} x {Bc}] forward {ms_f:.3f} ms = {fl / ms_f / 1e9:.0f} TF/s ({4.0 * KT * (Bn + Bc) / ms_f / 1e6:.0f} GB/s); "
          f"forward+backward {ms_fb:.3f} ms -> dE {ms_fb - ms_f:.3f} ms = {fl / (ms_fb - ms_f) / 1e9:.0f} TF/s")
    dbg = torch.zeros(8 * 148, device=dev, dtype=torch.int64)
    call("bm_set_debug_buffer", ptr(dbg))
    BF.clip_loss(est.detach(), cand, 0); torch.cuda.synchronize()
    call("bm_set_debug_buffer", None)
    d = dbg.reshape(148, 8).double(); lead = d[0::2]
    print("   MMA thread: wait converters %.0f, wait drain %.0f, of %.0f total; A converter: wait TMA %.0f, wait slot %.0f, work %.0f"
          % (lead[:, 0].mean(), lead[:, 6].mean(), lead[:, 1].mean(), lead[:, 2].mean(), lead[:, 3].mean(), lead[:, 4].mean()))
    del est, cand
BF.check_tc_status()
