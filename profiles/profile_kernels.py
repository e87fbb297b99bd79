"""Runs each hot kernel a few times at the BASELINE shape (B=256, T=360, H=320) so that one `ncu --set full`
capture per kernel is short (which = conv | wgrad | prep | topk | scores):   ncu --set full --clock-control none --import-source on -k regex:<name> -s 3 -c 2 \
                               -o gpurun_out/prof_<name> python profiles/profile_kernels.py <which>"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brainmagick_b200 import _lib  # noqa: E402
from brainmagick_b200._lib import call, ptr, stream  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "conv"
dev = "cuda"
B, T, H, Kw = 256, 360, 320, 3
status = torch.zeros(1, dtype=torch.int32, device=dev)
if which == "conv":
    x = torch.randn(B, T, H, device=dev)
    w = torch.randn(H, H, Kw, device=dev) * 0.03
    fh, fl = torch.empty(Kw, H, H, device=dev), torch.empty(Kw, H, H, device=dev)
    call("bm_tc_weight_split", ptr(w), H, H, Kw, ptr(fh), ptr(fl), None, None, stream())
    y = torch.empty(B, T, H, device=dev)
    for _ in range(6):
        call("bm_tc_conv1d_pair", ptr(x), ptr(fh), ptr(fl), None, None, B, T, H, H, Kw, 4, 1, 0, 0, 0, ptr(y), None, None,
             None, ptr(status), stream())
elif which == "wgrad":
    dy = torch.randn(B, T, H, device=dev)
    x = torch.randn(B, T, H, device=dev)
    ws = torch.empty(_lib.load().bm_tc_wgrad_workspace(B, H, H, Kw), device=dev)
    dw = torch.empty(H, H, Kw, device=dev)
    for _ in range(6):
        call("bm_tc_wgrad", ptr(dy), ptr(x), B, T, H, H, H, Kw, 4, ptr(ws), ptr(dw), None, ptr(status), stream())
elif which == "prep":          # -k regex:scale_clamp_crop_kernel   (batch preparation, HBM-bound)
    C, Tm, off, R = 273, 361, 18, 27
    x = torch.randn(B, C, Tm, device=dev)
    slot = torch.randint(0, R, (B,), device=dev, dtype=torch.int32)
    center, scale = torch.randn(R, C, device=dev) * 0.1, torch.rand(R, C, device=dev) + 0.5
    y = torch.empty(B, C, Tm - off, device=dev)
    for _ in range(6):
        call("bm_scale_clamp_crop", ptr(x), ptr(slot), ptr(center), ptr(scale), B, C, Tm, off, Tm - off, 20.0, 1, 0, ptr(y),
             None, stream())
elif which == "topk":          # -k regex:retrieval_topk_kernel     (fused softmax + top-k over an L2-resident score row)
    Bn, M = 1024, 4096
    scores = torch.randn(Bn, M, device=dev)
    labels = torch.arange(M, device=dev, dtype=torch.int64)
    targets = torch.randint(0, M, (Bn,), device=dev, dtype=torch.int64)
    top_idx = torch.empty(Bn, 10, device=dev, dtype=torch.int64)
    top_p = torch.empty(Bn, 10, device=dev)
    hit = torch.empty(Bn, device=dev, dtype=torch.int32)
    for _ in range(6):
        call("bm_retrieval_topk", ptr(scores), M, Bn, M, None, 0, 0, 10, ptr(labels), None, ptr(targets), ptr(top_idx),
             ptr(top_p), ptr(hit), None, None, None, stream())
elif which == "scores":        # -k regex:clip_scores_kernel           (retrieval / CLIP score GEMM, K = F*T = 368 640)
    Bn, M, KT = 1024, 2048, 1024 * 360
    est = torch.randn(Bn, KT, device=dev)
    cand = torch.randn(M, KT, device=dev)
    inv = torch.ones(M, device=dev)
    out = torch.empty(Bn, M, device=dev)
    need = int(_lib.load().bm_clip_workspace(Bn, M, KT))
    ws = torch.empty(max(need, 2), device=dev)
    for _ in range(3):
        call("bm_clip_scores", ptr(est), ptr(cand), Bn, M, KT, 1, ptr(inv), ptr(out), None, ptr(ws), ws.numel(), ptr(status),
             stream())
torch.cuda.synchronize()
assert int(status.item()) == 0
print("done", which)
