"""Per-stage device time (select / node tables / fused edge kernel / node update) of one layer config, from the
library's CUDA-event stage brackets.  python tools/stage_times.py c4|c2|c3layer [bf16|fp32]"""
import ctypes as C
import json
import os
import sys

import torch

torch.set_grad_enabled(False)      # forward benchmark: no autograd state is kept

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO]
from egnn_pytorch_b200 import EGNN, _native as nat  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c4"
dt = {"bf16": torch.bfloat16, "fp32": torch.float32}[sys.argv[2] if len(sys.argv) > 2 else "bf16"]
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(1)
cfgs = {
    "c4": (dict(dim=256, edge_dim=4, num_nearest_neighbors=32), 8, 4096),
    "c2": (dict(dim=512), 4, 1024),
    "c1": (dict(dim=512), 1, 16),
    "c3layer": (dict(dim=32, num_nearest_neighbors=8, norm_feats=True, coor_weights_clamp_value=2.0), 1, 1024),
}
kw, B, N = cfgs[name]
torch.manual_seed(0)
mod = EGNN(**kw).to(dt).to(dev).eval()
args = [torch.randn(B, N, kw["dim"], generator=g).to(dev, dt), torch.randn(B, N, 3, generator=g).to(dev)]
if kw.get("edge_dim", 0):
    args.append(torch.randn(B, N, N, kw["edge_dim"], generator=g).to(dev, dt))
lib = nat.load()
for _ in range(3):
    mod(*args)
torch.cuda.synchronize()
lib.egnn_profile_read(None, None, None, 1)
lib.egnn_profile_enable(1)
iters = 50
for _ in range(iters):
    mod(*args)
torch.cuda.synchronize()
ms = (C.c_float * 4)(); spans = (C.c_int32 * 4)(); launches = C.c_int64()
lib.egnn_profile_read(ms, spans, C.byref(launches), 1)
lib.egnn_profile_enable(0)
print(json.dumps(dict(config=name, dtype=str(dt), path=mod.last_path,
                      stage_ms={k: ms[i] / iters for i, k in enumerate(["select", "node_tables", "edge_kernel", "node_update"])},
                      launches_per_call=launches.value / iters)))
