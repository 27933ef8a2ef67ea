"""The oracle (oracle/egnn_oracle.py) against the committed outputs of the reference itself
(tests/golden/*.npz, produced by tests/golden/make_golden.py).  CPU only."""
import glob
import os

import numpy as np
import pytest

import cases

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


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
