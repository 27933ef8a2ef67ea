"""The oracle (oracle/egnn_oracle.py) against the committed outputs of the reference itself
(tests/golden/*.npz, produced by tests/golden/make_golden.py).  CPU only."""
import glob
import os

import numpy as np
import pytest

import cases

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = sorted(n for n in (os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))
               if not n.startswith("grad_"))


def test_every_spec_has_a_fixture():
    assert set(NAMES) == set(cases.SPECS), set(NAMES) ^ set(cases.SPECS)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference_fp64(name):
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    case = cases.build_case(cases.SPECS[name])
    assert cases.case_checksum(case) == str(g["checksum"]), "regenerated inputs drifted from the fixture"
    if bool(g["tie_dependent"]):
        pytest.skip("reference output depends on torch.topk's tie order (see tests/cases.py)")
    if case["kind"] == "network":
        ins = case["inputs"]
        feats, coors, changes = cases.O.egnn_network_forward(
            case["params"], case["ncfg"], ins["feats"], ins["coors"], adj_mat=ins.get("adj_mat"),
            edges=ins.get("edges"), mask=ins.get("mask"), return_coor_changes=True)
        np.testing.assert_allclose(np.stack(changes), g["coor_changes"], rtol=1e-10, atol=1e-10)
    else:
        feats, coors = cases.run_oracle(case)
    # float64 vs float64: only summation order differs (numpy vs ATen)
    np.testing.assert_allclose(feats, g["feats"], rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(coors, g["coors"], rtol=1e-10, atol=1e-10)


def test_oracle_row_restriction_matches_full():
    case = cases.build_case(cases.SPECS["knn_edges_mask"])
    full = cases.run_oracle(case)
    part = cases.run_oracle(case, rows=(5, 11))
    np.testing.assert_array_equal(full[0][:, 5:11], part[0])
    np.testing.assert_array_equal(full[1][:, 5:11], part[1])


def test_oracle_equivariance_fp64():
    """The reference's own property test (tests/test_equivariance.py:8-34) on the oracle."""
    case = cases.build_case(cases.SPECS["dense_edges"])
    rs = np.random.RandomState(0)
    q, _ = np.linalg.qr(rs.standard_normal((3, 3)))
    t = rs.standard_normal((1, 1, 3))
    ins = dict(case["inputs"])
    f2, c2 = cases.run_oracle(case)
    ins_r = dict(ins, coors=ins["coors"] @ q + t)
    f1, c1 = cases.run_oracle(dict(case, inputs=ins_r))
    assert np.abs(f1 - f2).max() < 1e-12
    assert np.abs(c1 - (c2 @ q + t)).max() < 1e-12


# ---------------------------------------------------------------- backward oracle (oracle/egnn_oracle_grad.py)


def test_every_grad_spec_has_a_fixture():
    have = {os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(GOLDEN, "grad_*.npz"))}
    assert have == set(cases.GRAD_SPECS), have ^ set(cases.GRAD_SPECS)


@pytest.mark.parametrize("name", cases.GRAD_SPECS)
def test_grad_oracle_matches_reference_autograd_fp64(name):
    g = np.load(os.path.join(GOLDEN, f"grad_{name}.npz"))
    case = cases.build_case(cases.SPECS[name])
    assert cases.case_checksum(case) == str(g["checksum"])
    if bool(g["tie_dependent"]):
        pytest.skip("reference gradients depend on torch.topk's tie order")
    mine = cases.flatten_grads(cases.run_oracle_grad(case))
    keys = {k for k in g.files if k.startswith(("in.", "p."))}
    assert keys == set(mine)
    # CoorsNorm divides the self pair (|x_i - x_i| = 0) by eps = 1e-8: +-7e7-sized terms cancel in both
    # implementations and leave ~1e-9 of rounding noise; everything else agrees to summation order.
    tol = 1e-7 if "norm_coors" in str(case["spec"]["cfg"]) else 1e-11
    for k in sorted(keys):
        scale = max(1.0, float(np.abs(g[k]).max()))
        assert np.abs(mine[k] - g[k]).max() <= tol * scale, k


def test_grad_oracle_against_finite_differences():
    """Independent of the fixtures: directional derivative of the FORWARD oracle by central differences."""
    case = cases.build_case(cases.SPECS["knn_edges_mask"])
    gf, gx = cases.upstream_grads(case)
    grads = cases.flatten_grads(cases.run_oracle_grad(case))
    rs = np.random.RandomState(7)

    def loss(c):
        f, x = cases.run_oracle(c)
        return float((f * gf).sum() + (x * gx).sum())

    for group, prefix in (("inputs", "in."), ("params", "p.")):
        for key in [k[len(prefix):] for k in grads if k.startswith(prefix)]:
            v = rs.standard_normal(np.shape(case[group][key]))
            eps = 1e-6
            hi = dict(case, **{group: dict(case[group], **{key: case[group][key] + eps * v})})
            lo = dict(case, **{group: dict(case[group], **{key: case[group][key] - eps * v})})
            fd = (loss(hi) - loss(lo)) / (2 * eps)
            an = float((grads[prefix + key] * v).sum())
            assert abs(fd - an) <= 1e-6 * max(1.0, abs(an)), (key, fd, an)


def test_grad_oracle_is_equivariant():
    """Size-independent property of the backward: for coors -> coors Q + t and the coordinate cotangent G_x -> G_x Q,
    the loss is unchanged, so dL/dfeats and every parameter gradient are invariant and dL/dcoors rotates with Q."""
    from oracle import egnn_oracle_grad as G
    case = cases.build_case(cases.SPECS["knn_norm_coors"])
    ins = case["inputs"]
    gf, gx = cases.upstream_grads(case)
    rs = np.random.RandomState(1)
    q, _ = np.linalg.qr(rs.standard_normal((3, 3)))
    t = rs.standard_normal((1, 1, 3))
    a = G.egnn_layer_backward(case["params"], case["cfg"], ins["feats"], ins["coors"], ins.get("edges"), ins.get("mask"),
                              None, gf, gx)
    b = G.egnn_layer_backward(case["params"], case["cfg"], ins["feats"], ins["coors"] @ q + t, ins.get("edges"),
                              ins.get("mask"), None, gf, gx @ q)
    assert np.abs(a["feats"] - b["feats"]).max() < 1e-6        # (CoorsNorm's 1/eps self pair leaves ~1e-9 of noise)
    assert np.abs(a["coors"] @ q - b["coors"]).max() < 1e-6
    for k in a["params"]:
        assert np.abs(a["params"][k] - b["params"][k]).max() < 1e-6 * max(1.0, np.abs(a["params"][k]).max()), k
