"""Stage times of BASELINE config 4 (EGNN(256, edge_dim=4), k=32, N=4096, 8 graphs per GPU, bf16) and of c3 / c5 layers."""
import ctypes as C, os, sys, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [REPO]
from egnn_pytorch_b200 import EGNN, _native as nat
torch.set_grad_enabled(False); lib = nat.load(); dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(1)
m = EGNN(dim=256, edge_dim=4, num_nearest_neighbors=32).bfloat16().to(dev).eval()
f = torch.randn(8, 4096, 256, generator=g).to(dev, torch.bfloat16); x = torch.randn(8, 4096, 3, generator=g).to(dev)
e = torch.randn(8, 4096, 4096, 4, generator=g, dtype=torch.bfloat16).to(dev)
for _ in range(5): m(f, x, e)
torch.cuda.synchronize(); lib.egnn_profile_read(None, None, None, 1); lib.egnn_profile_enable(1)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); a.record()
for _ in range(20): m(f, x, e)
b.record(); torch.cuda.synchronize()
ms = (C.c_float * 4)(); sp = (C.c_int32 * 4)(); ln = C.c_int64(); lib.egnn_profile_read(ms, sp, C.byref(ln), 1)
print("c4 ms/fwd %.4f  select %.4f pre %.4f edge %.4f post %.4f" % (a.elapsed_time(b) / 20, ms[0] / 20, ms[1] / 20, ms[2] / 20, ms[3] / 20))
