"""Stage times (library event brackets) of the launch-bound BASELINE configs c1 / c3 / c5, warm loop."""
import ctypes as C, os, sys, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [REPO]
from egnn_pytorch_b200 import EGNN, EGNN_Network, _native as nat
torch.set_grad_enabled(False); lib = nat.load(); dev = torch.device("cuda", 0)
def stages(fn, iters=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); lib.egnn_profile_read(None, None, None, 1); lib.egnn_profile_enable(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    ms = (C.c_float * 4)(); sp = (C.c_int32 * 4)(); ln = C.c_int64(); lib.egnn_profile_read(ms, sp, C.byref(ln), 1); lib.egnn_profile_enable(0)
    return a.elapsed_time(b) / iters * 1e3, [ms[i] / iters * 1e3 for i in range(4)], ln.value / iters
g = torch.Generator().manual_seed(1)
m = EGNN(dim=512).to(dev).eval(); f, x = torch.randn(1, 16, 512, generator=g).to(dev), torch.randn(1, 16, 3, generator=g).to(dev)
print("c1 fp32 us/fwd %.1f  stages[select,pre,edge,post] %s launches %.0f" % stages(lambda: m(f, x)))
mb = EGNN(dim=512).bfloat16().to(dev).eval(); fb = f.bfloat16()
print("c1 bf16 us/fwd %.1f  stages %s launches %.0f" % stages(lambda: mb(fb, x)))
for dt in (torch.float32, torch.bfloat16):
    net = EGNN_Network(num_tokens=21, num_positions=1024, dim=32, depth=3, num_nearest_neighbors=8, coor_weights_clamp_value=2.0).to(dt).to(dev).eval()
    tok = torch.randint(0, 21, (1, 1024), generator=g).to(dev); xx = torch.randn(1, 1024, 3, generator=g).to(dev); mk = torch.ones(1, 1024, dtype=torch.bool, device=dev)
    print("c3 %s us/fwd %.1f  stages %s launches %.0f" % ((str(dt)[6:],) + stages(lambda: net(tok, xx, mask=mk))))
    n = 8192; i = torch.arange(n, device=dev); adj = (i[:, None] - i[None, :]).abs() <= 1
    net5 = EGNN_Network(num_tokens=21, dim=32, depth=3, num_adj_degrees=3, adj_dim=8, only_sparse_neighbors=True).to(dt).to(dev).eval()
    tok = torch.randint(0, 21, (1, n), generator=g).to(dev); xx = torch.randn(1, n, 3, generator=g).to(dev); mk = torch.ones(1, n, dtype=torch.bool, device=dev)
    print("c5 %s us/fwd %.1f  stages %s launches %.0f" % ((str(dt)[6:],) + stages(lambda: net5(tok, xx, adj_mat=adj, mask=mk), 10)))
