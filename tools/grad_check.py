"""Diagnostic table: per-case, per-tensor gradient error of the CUDA backward vs the numpy backward oracle.

    python tools/grad_check.py [case ...] [--fp32]
"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tests")):
    sys.path.insert(0, p)
import cases  # noqa: E402
from test_gpu_grad import module_grads  # noqa: E402

dtype = torch.float32 if "--fp32" in sys.argv else torch.float64
names = [a for a in sys.argv[1:] if not a.startswith("--")] or cases.GRAD_SPECS
for name in names:
    case = cases.build_case(cases.SPECS[name])
    try:
        got = module_grads(case, dtype)
    except Exception as e:  # noqa: BLE001
        print(f"{name:22s} FAILED: {type(e).__name__}: {e}")
        continue
    want = cases.flatten_grads(cases.run_oracle_grad(case))
    errs = {}
    for k in sorted(want):
        scale = max(1.0, float(np.abs(want[k]).max()))
        errs[k] = float(np.abs(got[k] - want[k]).max()) / scale if np.isfinite(got[k]).all() else float("nan")
    worst = max(errs.values(), key=lambda v: (np.isnan(v), v))
    bad = {k: v for k, v in errs.items() if not (v < (1e-6 if dtype == torch.float64 else 1e-3))}
    print(f"{name:22s} worst {worst:.2e}" + ("" if not bad else "  BAD: " + ", ".join(f"{k}={v:.2e}" for k, v in bad.items())))
