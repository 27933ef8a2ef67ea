"""Bring-up of the tcgen05 kernels on a real B200: runs the standalone GEMM and one fused-layer case
under every EGNN_TC_VARIANT (descriptor LBO/SBO order, bf16 pair order) in subprocesses with a timeout,
so a wrong guess cannot hang the box, and prints the errors.  python tools/tc_bringup.py"""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "tests")]

CHILD = r'''
import ctypes as C, json, os, sys
import numpy as np, torch
sys.path[:0] = [%(repo)r, %(repo)r + "/tests"]
from egnn_pytorch_b200 import _native as nat
import cases, util
lib = nat.load()
mode = sys.argv[1]
if mode == "gemm":
    torch.manual_seed(0)
    out = {}
    for (M, N, K) in [(128, 128, 64), (200, 136, 72), (512, 320, 512)]:
        A = torch.randn(M, K, device="cuda").bfloat16(); W = torch.randn(N, K, device="cuda").bfloat16()
        bias = torch.randn(N, device="cuda")
        o = torch.empty(M, N, device="cuda", dtype=torch.float32)
        rc = lib.egnn_gemm_bf16(M, N, K, A.data_ptr(), W.data_ptr(), bias.data_ptr(), 0.5, 0, o.data_ptr(), 1, None)
        torch.cuda.synchronize()
        ref = 0.5 * (A.float() @ W.float().t() + bias)
        out[f"{M}x{N}x{K}"] = dict(rc=rc, err=float((o - ref).abs().max()), ref=float(ref.abs().max()))
        ob = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        rc = lib.egnn_gemm_bf16(M, N, K, A.data_ptr(), W.data_ptr(), None, 1.0, 1, ob.data_ptr(), 0, None)
        torch.cuda.synchronize()
        ref = torch.nn.functional.silu(A.float() @ W.float().t())
        out[f"{M}x{N}x{K}_silu_bf16"] = dict(rc=rc, err=float((ob.float() - ref).abs().max()), ref=float(ref.abs().max()))
    print(json.dumps(out))
else:
    spec = dict(kind="layer", cfg=dict(dim=64), B=2, N=160, seed=99, init="xavier")
    case = cases.build_case(spec)
    r = lambda v: torch.from_numpy(np.asarray(v, np.float64)).bfloat16().double().numpy()
    case["params"] = {k: r(v) for k, v in case["params"].items()}
    case["inputs"]["feats"] = r(case["inputs"]["feats"])
    mod = util.make_module(case, torch.bfloat16)
    out = util.run_module(mod, case, torch.bfloat16)
    torch.cuda.synchronize()
    want = cases.run_oracle(case)
    print(json.dumps(dict(path=mod.last_path, feats_err=util.max_err(out[0], want[0]), feats_max=float(np.abs(want[0]).max()),
                          coors_err=util.max_err(out[1], want[1]),
                          dcoors_max=float(np.abs(want[1] - case["inputs"]["coors"]).max()))))
'''


def run(mode, variant):
    env = dict(os.environ, EGNN_TC_VARIANT=str(variant))
    try:
        res = subprocess.run([sys.executable, "-c", CHILD % dict(repo=REPO), mode], env=env, capture_output=True, text=True,
                             timeout=180)
        last = [l for l in res.stdout.strip().splitlines() if l.startswith("{")]
        return json.loads(last[-1]) if last else dict(error=(res.stderr or res.stdout)[-600:])
    except subprocess.TimeoutExpired:
        return dict(error="TIMEOUT (hang)")


if __name__ == "__main__":
    for v in (0, 1):
        print("gemm variant", v, json.dumps(run("gemm", v)), flush=True)
    for v in (0, 1, 2, 3):
        print("layer variant", v, json.dumps(run("layer", v)), flush=True)
