"""The reference's own property tests (reference tests/test_equivariance.py:8-102, the four dense-API tests) run
against this package's `EGNN`, with the same shapes, the same float64 CPU tensors and the same atol=1e-6.  CPU
tensors are staged to the GPU by the module (transport only); the fp64 SIMT kernels do the arithmetic."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def rot(alpha, beta, gamma):
    """Rz(alpha) Ry(beta) Rz(gamma), as reference egnn_pytorch/utils.py:4-19 builds it."""
    def rz(t):
        return torch.tensor([[math.cos(t), -math.sin(t), 0], [math.sin(t), math.cos(t), 0], [0, 0, 1]], dtype=torch.float64)

    def ry(t):
        return torch.tensor([[math.cos(t), 0, math.sin(t)], [0, 1, 0], [-math.sin(t), 0, math.cos(t)]], dtype=torch.float64)
    return rz(alpha) @ ry(beta) @ rz(gamma)


@pytest.fixture(autouse=True)
def _fp64_default():
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)       # reference tests/test_equivariance.py:6
    yield
    torch.set_default_dtype(old)


def _check(layer, n, edge_dim, seed):
    torch.manual_seed(seed)
    R = rot(*torch.rand(3).tolist())
    T = torch.randn(1, 1, 3)
    feats = torch.randn(1, n, 512)
    coors = torch.randn(1, n, 3)
    edges = torch.randn(1, n, n, edge_dim)
    mask = torch.ones(1, n).bool()
    swapped = feats.clone()
    swapped[:, 0, :], swapped[:, 1, :] = feats[:, 1, :], feats[:, 0, :]
    feats1, coors1 = layer(feats, coors @ R + T, edges, mask=mask)
    feats2, coors2 = layer(feats, coors, edges, mask=mask)
    feats3, coors3 = layer(swapped, coors, edges, mask=mask)
    assert feats1.dtype == torch.float64 and feats1.device.type == "cpu"
    assert torch.allclose(feats1, feats2, atol=1e-6), "type 0 features are invariant"
    assert torch.allclose(coors1, (coors2 @ R + T), atol=1e-6), "type 1 features are equivariant"
    assert not torch.allclose(feats1, feats3, atol=1e-6), "layer must be equivariant to permutations of node order"
    assert layer.last_path == "fp64-simt"


def test_egnn_equivariance():                                   # reference :8-34
    from egnn_pytorch_b200 import EGNN
    _check(EGNN(dim=512, edge_dim=4), 16, 4, 0)


def test_higher_dimension():                                    # reference :36-45 (coordinate dimension 5)
    from egnn_pytorch_b200 import EGNN
    layer = EGNN(dim=512, edge_dim=4)
    feats, coors = layer(torch.randn(1, 16, 512), torch.randn(1, 16, 5), torch.randn(1, 16, 16, 4), mask=torch.ones(1, 16).bool())
    assert feats.shape == (1, 16, 512) and coors.shape == (1, 16, 5) and torch.isfinite(coors).all()


def test_egnn_equivariance_with_nearest_neighbors():            # reference :47-73
    from egnn_pytorch_b200 import EGNN
    _check(EGNN(dim=512, edge_dim=1, num_nearest_neighbors=8), 256, 1, 1)


def test_egnn_equivariance_with_coord_norm():                   # reference :76-102
    from egnn_pytorch_b200 import EGNN
    _check(EGNN(dim=512, edge_dim=1, num_nearest_neighbors=8, norm_coors=True), 256, 1, 2)
