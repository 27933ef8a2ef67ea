"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported read-only from
/root/reference) in float64 on the deterministic cases of tests/cases.py.

Run in the build container only (the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Each fixture stores: the case name, the sha256 of the regenerated parameters+inputs, and the
reference outputs (feats, coors[, coor_changes]) in float64.  Nothing from the reference's
source is copied; only its numerical outputs are recorded.
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


def t(x):
    if x is None:
        return None
    x = np.asarray(x)
    if x.dtype == bool:
        return torch.from_numpy(x.copy())
    if np.issubdtype(x.dtype, np.integer):
        return torch.from_numpy(x.astype(np.int64))
    return torch.from_numpy(x.astype(np.float64))


def run_reference(case):
    spec, ins = case["spec"], case["inputs"]
    torch.set_default_dtype(torch.float64)
    if case["kind"] == "network":
        mod = EGNN_Network(**spec["cfg"])
    else:
        mod = EGNN(**spec["cfg"])
    sd = {k: t(v) for k, v in case["params"].items()}
    missing, unexpected = mod.load_state_dict(sd, strict=False)
    # every generated parameter must land on a reference parameter and vice versa
    assert not unexpected, unexpected
    assert not missing, missing
    mod.eval()
    with torch.no_grad():
        if case["kind"] == "network":
            out = mod(t(ins["feats"]).clone(), t(ins["coors"]), adj_mat=t(ins.get("adj_mat")),
                      edges=t(ins.get("edges")), mask=t(ins.get("mask")), return_coor_changes=True)
            feats, coors, changes = out
            return feats.numpy(), coors.numpy(), np.stack([c.numpy() for c in changes])
        feats, coors = mod(t(ins["feats"]), t(ins["coors"]), t(ins.get("edges")),
                           mask=t(ins.get("mask")), adj_mat=t(ins.get("adj_mat")))
        return feats.numpy(), coors.numpy(), None


def main():
    names = sys.argv[1:] or list(cases.SPECS)
    worst = 0.0
    for name in names:
        spec = cases.SPECS[name]
        case = cases.build_case(spec)
        feats, coors, changes = run_reference(case)
        o = cases.run_oracle(case)
        err = max(np.abs(o[0] - feats).max(), np.abs(o[1] - coors).max())
        tie = bool(spec.get("tie_dependent", False))
        print(f"{name:22s} oracle-vs-reference max|err| = {err:.3e}{'  (tie-dependent, not pinned)' if tie else ''}")
        if not tie:
            worst = max(worst, err)
        payload = dict(name=name, checksum=cases.case_checksum(case), feats=feats, coors=coors,
                       tie_dependent=tie)
        if changes is not None:
            payload["coor_changes"] = changes
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **payload)
    print(f"worst pinned error {worst:.3e}")


if __name__ == "__main__":
    main()
