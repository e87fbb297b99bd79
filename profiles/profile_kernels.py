"""Runs each hot kernel a few times at the BASELINE shape (B=256, T=360, H=320) so that one `ncu --set full`
capture per kernel is short:   ncu --set full --clock-control none --import-source on -k regex:<name> -s 3 -c 2 \
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
        call("bm_tc_conv1d_pair", ptr(x), ptr(fh), ptr(fl), None, None, B, T, H, H, Kw, 4, 1, 0, 0, 0, ptr(y), None,
             None, ptr(status), stream())
elif which == "wgrad":
    dy = torch.randn(B, T, H, device=dev)
    x = torch.randn(B, T, H, device=dev)
    ws = torch.empty(_lib.load().bm_tc_wgrad_workspace(B, H, H, Kw), device=dev)
    dw = torch.empty(H, H, Kw, device=dev)
    for _ in range(6):
        call("bm_tc_wgrad", ptr(dy), ptr(x), B, T, H, H, H, Kw, 4, ptr(ws), ptr(dw), None, ptr(status), stream())
torch.cuda.synchronize()
assert int(status.item()) == 0
print("done", which)
