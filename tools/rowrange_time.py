"""One rank's share of the row-sharded probe on a single GPU: EGNN(dim=512) bf16, N=8192, rows [0, N/P)."""
import os, sys, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [REPO]
from egnn_pytorch_b200 import EGNN
torch.set_grad_enabled(False)
dev = torch.device("cuda", 0); torch.manual_seed(0)
m = EGNN(dim=512).bfloat16().to(dev).eval()
n = 8192
f = torch.randn(1, n, 512, device=dev).bfloat16(); x = torch.randn(1, n, 3, device=dev)
def t(fn, it=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): out = fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / it, out
full_ms, full = t(lambda: m(f, x), 3)
print(f"full graph {full_ms:.3f} ms")
for P in (2, 4, 8, 16):
    r = n // P
    ms, out = t(lambda: m(f, x, _rows=(r, 2 * r)))
    ok = torch.equal(out[0][:, r:2 * r], full[0][:, r:2 * r]) and torch.equal(out[1][:, r:2 * r], full[1][:, r:2 * r])
    print(f"P={P:2d} rows {r:5d}: {ms:.3f} ms  ideal {full_ms / P:.3f}  efficiency {full_ms / P / ms:.3f}  bit-identical to the full forward: {ok}")
