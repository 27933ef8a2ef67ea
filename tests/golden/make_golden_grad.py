"""Generate tests/golden/grad_*.npz: gradients of the UNMODIFIED reference (imported read-only from
/root/reference) by torch autograd in float64, for the cases of tests/cases.py:GRAD_SPECS and the fixed
cotangents of cases.upstream_grads.

Run in the build container only:   PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_grad.py

Each fixture stores the case checksum and the flat gradient dict of cases.flatten_grads ('in.feats',
'in.coors', 'in.edges', 'p.<state-dict key>').  Only numerical outputs of the reference are recorded.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.dont_write_bytecode = True
REF = os.environ.get("EGNN_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

import torch  # noqa: E402
import cases  # noqa: E402
from egnn_pytorch import EGNN, EGNN_Network  # noqa: E402  (the reference)
from make_golden import t  # noqa: E402


def reference_grads(case):
    spec, ins = case["spec"], case["inputs"]
    torch.set_default_dtype(torch.float64)
    mod = (EGNN_Network if case["kind"] == "network" else EGNN)(**spec["cfg"])
    missing, unexpected = mod.load_state_dict({k: t(v) for k, v in case["params"].items()}, strict=False)
    assert not missing and not unexpected
    mod.eval()
    gf, gx = (torch.from_numpy(g) for g in cases.upstream_grads(case))
    feats, coors, edges = t(ins["feats"]), t(ins["coors"]), t(ins.get("edges"))
    leaves = {"coors": coors.requires_grad_(True)}
    if feats.is_floating_point():
        leaves["feats"] = feats.requires_grad_(True)
    if edges is not None and edges.is_floating_point():
        leaves["edges"] = edges.requires_grad_(True)
    if case["kind"] == "network":
        fo, xo = mod(feats.clone() if not feats.is_floating_point() else feats, coors, adj_mat=t(ins.get("adj_mat")),
                     edges=edges, mask=t(ins.get("mask")))
    else:
        fo, xo = mod(feats, coors, edges, mask=t(ins.get("mask")), adj_mat=t(ins.get("adj_mat")))
    ((fo * gf).sum() + (xo * gx).sum()).backward()
    out = {f"in.{k}": v.grad.numpy() for k, v in leaves.items()}
    for k, p in mod.named_parameters():
        out[f"p.{k}"] = (torch.zeros_like(p) if p.grad is None else p.grad).numpy()
    return out


def main():
    names = sys.argv[1:] or cases.GRAD_SPECS
    worst = 0.0
    for name in names:
        case = cases.build_case(cases.SPECS[name])
        ref = reference_grads(case)
        mine = cases.flatten_grads(cases.run_oracle_grad(case))
        assert set(ref) == set(mine), (sorted(set(ref) ^ set(mine)))
        err = 0.0
        for k in ref:
            scale = max(1.0, float(np.abs(ref[k]).max()))
            err = max(err, float(np.abs(ref[k] - mine[k]).max()) / scale)
        tie = bool(cases.SPECS[name].get("tie_dependent", False))
        print(f"{name:22s} oracle-grad vs reference autograd: max rel err {err:.3e}{'  (tie-dependent)' if tie else ''}")
        if not tie:
            worst = max(worst, err)
        np.savez_compressed(os.path.join(HERE, f"grad_{name}.npz"), name=name, checksum=cases.case_checksum(case),
                            tie_dependent=tie, **ref)
    print(f"worst pinned error {worst:.3e}")


if __name__ == "__main__":
    main()
