"""EGNN_Network with only_sparse_neighbors + num_adj_degrees + a node mask: the neighbour lists depend on the adjacency
only (reference egnn_pytorch.py:249-260, :296), so the network builds them once (egnn_adj_neighbors) and every layer runs
in edge-list mode.  That must give exactly the per-layer adjacency-scan results."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _chain(n):
    i = torch.arange(n)
    return (i[:, None] - i[None, :]).abs() == 1


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_cached_lists_equal_per_layer_scan(dtype, monkeypatch):
    from egnn_pytorch_b200 import EGNN_Network
    torch.manual_seed(3)
    dev = torch.device("cuda", 0)
    n = 96
    net = EGNN_Network(num_tokens=11, dim=32, depth=3, num_adj_degrees=2, adj_dim=4, only_sparse_neighbors=True,
                       norm_coors=True, coor_weights_clamp_value=2.0).to(dev).to(dtype).eval()
    tok = torch.randint(0, 11, (2, n), device=dev)
    x = torch.randn(2, n, 3, device=dev)
    mask = torch.rand(2, n, device=dev) > 0.2
    adj = _chain(n).to(dev)
    with torch.no_grad():
        f1, x1 = net(tok, x, adj_mat=adj, mask=mask)
        assert net.__dict__["_adj_cache"][5] is not None            # the lists were built
        monkeypatch.setenv("EGNN_B200_NO_LIST_CACHE", "1")
        net.__dict__.pop("_adj_cache")
        f2, x2 = net(tok, x, adj_mat=adj, mask=mask)
        assert net.__dict__["_adj_cache"][5] is None
    assert torch.equal(f1, f2) and torch.equal(x1, x2)


def test_adj_neighbors_lists():
    from egnn_pytorch_b200 import _native as nat
    lib = nat.load()
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(5)
    b, n, k = 2, 70, 9
    adj = (torch.rand(b, n, n, generator=g) < 0.08)
    adj = adj | adj.transpose(1, 2)
    a = adj.to(torch.uint8).to(dev).contiguous()
    out = torch.empty(b, n, k, dtype=torch.int32, device=dev)
    nat.check("egnn_adj_neighbors", lib.egnn_adj_neighbors(b, n, k, C.c_void_p(a.data_ptr()), 1, C.c_void_p(out.data_ptr()), None,
                                                           C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    an = adj.numpy()
    for bi in range(b):
        for i in range(n):
            want = [i] + [j for j in range(n) if an[bi, i, j] and j != i]
            want = (want + [-1] * k)[:k]
            assert got[bi, i].tolist() == want
