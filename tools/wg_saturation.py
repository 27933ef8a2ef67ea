"""How many of the dense kernel's 4 warpgroups does it take to saturate the MUFU pipe?  N = 128 / 256 / 384 / 512 keeps
1 / 2 / 3 / 4 warpgroups busy (warpgroup g owns j-tile [128 g, 128 g + 128) of every 512-wide j-block); B is chosen so
that every run has the same number of pairs.  Prints the fused-stage time and the SFU-roofline fraction per N.
    python tools/wg_saturation.py [--no-coors]"""
import ctypes as C
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO]
from egnn_pytorch_b200 import EGNN, _native as nat  # noqa: E402

torch.set_grad_enabled(False)
lib = nat.load()
dev = torch.device("cuda", 0)
kw = dict(update_coors=False) if "--no-coors" in sys.argv else {}
torch.manual_seed(0)
mod = EGNN(dim=512, **kw).bfloat16().to(dev).eval()
for n in (128, 256, 384, 512, 640, 768, 1024, 2048):
    b = max(1, (4 * 1024 * 1024) // (n * n))
    f = torch.randn(b, n, 512, device=dev).bfloat16()
    x = torch.randn(b, n, 3, device=dev)
    for _ in range(3):
        mod(f, x)
    torch.cuda.synchronize()
    lib.egnn_profile_read(None, None, None, 1)
    lib.egnn_profile_enable(1)
    for _ in range(10):
        mod(f, x)
    torch.cuda.synchronize()
    ms = (C.c_float * 4)(); spans = (C.c_int32 * 4)(); launches = C.c_int64()
    lib.egnn_profile_read(ms, spans, C.byref(launches), 1)
    lib.egnn_profile_enable(0)
    t = ms[2] / spans[2]
    pairs = b * n * n
    act = 2050 + 16 + (64 if not kw else 0)
    print(f"N={n:5d} B={b:4d} pairs={pairs:8d} edge {t:7.4f} ms  {pairs / t / 1e6:8.1f} Gpairs/s  sfu frac {pairs * act / (t * 1e-3) / 4653.12e9:.3f}", flush=True)
