"""EGNN(dim=512) bf16 forward at N=128 (only warpgroup 0 of the dense kernel has work): profiling target."""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO]
from egnn_pytorch_b200 import EGNN
torch.set_grad_enabled(False)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
b = (4 * 1024 * 1024) // (n * n)
mod = EGNN(dim=512).bfloat16().cuda().eval()
f = torch.randn(b, n, 512, device="cuda").bfloat16(); x = torch.randn(b, n, 3, device="cuda")
for _ in range(6):
    mod(f, x)
torch.cuda.synchronize()
