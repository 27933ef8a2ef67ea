"""CPU oracle for the BACKWARD of the EGNN hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

The reference has no backward code of its own: gradients come from PyTorch autograd applied to
`EGNN.forward` (reference egnn_pytorch/egnn_pytorch.py:224-341) and `EGNN_Network.forward`
(:390-454).  This file restates that derivative by hand in numpy float64, walking the reference's
own formulation (concat edge input, Linear over the concatenation) in reverse, line by line.  The
CUDA product differentiates an algebraically split form with recomputation, so agreement is an
independent check.

Pinned against the reference itself: `tests/golden/make_golden.py` runs torch autograd on the
unmodified reference in float64 and commits the gradients (`tests/golden/grad_*.npz`);
`tests/test_oracle_golden.py` replays them through this file.

Same import rule as egnn_oracle.py: tests, smoke() and bench's CPU legs only.
"""
from __future__ import annotations

import numpy as np

from .egnn_oracle import (adjacency_degrees, fourier_features_of, layer_norm, linear, neighbour_selection,
                          sigmoid, silu)


def dsilu(x):
    """d/dx [x * sigmoid(x)]."""
    s = sigmoid(x)
    return s * (1.0 + x * (1.0 - s))


def _linear_bwd(gy, x, w):
    """y = x @ w.T + b  ->  (gx, gw, gb); leading axes of x are flattened for the parameter sums."""
    gx = gy @ w
    gw = gy.reshape(-1, gy.shape[-1]).T @ x.reshape(-1, x.shape[-1])
    gb = gy.reshape(-1, gy.shape[-1]).sum(0)
    return gx, gw, gb


def egnn_layer_backward(params, cfg, feats, coors, edges, mask, adj_mat, g_feats_out, g_coors_out, neighbors=None):
    """Gradient of sum(feats_out * g_feats_out) + sum(coors_out * g_coors_out) of one layer.

    `neighbors` int [B,N,k] (edge-list mode, see `egnn_oracle.egnn_layer_forward_edge_list`): these lists stand
    in for the top-k result; entries < 0 are empty slots (their pair is masked out of every sum).

    Returns dict(feats [B,N,dim], coors [B,N,C], edges [B,N,N,e] | None, params {state-dict key: grad}).
    Neighbour selection (egnn_pytorch.py:237-260) is piecewise constant and contributes no gradient, exactly
    as `topk` indices carry none in autograd."""
    f8 = np.float64
    P = {k: np.asarray(v, dtype=f8) for k, v in params.items()}
    feats = np.asarray(feats, dtype=f8)
    coors = np.asarray(coors, dtype=f8)
    edges = None if edges is None else np.asarray(edges, dtype=f8)
    mask = None if mask is None else np.asarray(mask).astype(bool)
    go = np.asarray(g_feats_out, dtype=f8)
    gxo = np.asarray(g_coors_out, dtype=f8)
    b, n, d = feats.shape
    F = cfg["fourier_features"]
    use_nearest = cfg["num_nearest_neighbors"] > 0 or cfg["only_sparse_neighbors"]
    bidx = np.arange(b)[:, None, None]
    iidx = np.arange(n)[None, :, None]

    # ------------------------------------------------------------ forward, keeping intermediates
    slot_ok = None
    if neighbors is not None:
        use_nearest = True
        nb = np.asarray(neighbors).astype(np.int64)
        slot_ok = nb >= 0
        jidx = np.where(slot_ok, nb, np.broadcast_to(iidx, nb.shape))
        nbhd_mask = np.ones(nb.shape, dtype=bool)
        xj, hj = coors[bidx, jidx], feats[bidx, jidx]
        eij = None if edges is None else edges[bidx, iidx, jidx]
    elif use_nearest:
        jidx, nbhd_mask, _ = neighbour_selection(cfg, coors, mask, adj_mat)
        xj, hj = coors[bidx, jidx], feats[bidx, jidx]
        eij = None if edges is None else edges[bidx, iidx, jidx]
    else:
        jidx = np.broadcast_to(np.arange(n)[None, None, :], (b, n, n))
        xj = np.broadcast_to(coors[:, None], (b, n, n, coors.shape[-1]))
        hj = np.broadcast_to(feats[:, None], (b, n, n, d))
        eij = edges
    J = jidx.shape[-1]
    rel = coors[:, :, None, :] - xj                                   # :232
    dist = (rel ** 2).sum(-1)                                         # :233
    dfeat = fourier_features_of(dist, F) if F > 0 else dist[..., None]
    hi_b = np.broadcast_to(feats[:, :, None, :], (b, n, J, d))
    edge_in = np.concatenate([hi_b, hj, dfeat] + ([eij] if eij is not None else []), axis=-1)
    W1, b1, W2, b2 = (P["edge_mlp.0.weight"], P["edge_mlp.0.bias"], P["edge_mlp.3.weight"], P["edge_mlp.3.bias"])
    pre1 = linear(edge_in, W1, b1)
    hid = silu(pre1)
    pre2 = linear(hid, W2, b2)
    s2 = silu(pre2)
    if cfg["soft_edges"]:
        gz = linear(s2, P["edge_gate.0.weight"], P["edge_gate.0.bias"])
        gate = sigmoid(gz)
        m_ij = s2 * gate
    else:
        m_ij = s2
    pmask = None
    if mask is not None:
        pmask = (mask[:, :, None] & mask[bidx, jidx] & nbhd_mask) if use_nearest else (mask[:, :, None] & mask[:, None, :])
    has_node_mask = pmask is not None
    if slot_ok is not None:
        pmask = slot_ok if pmask is None else (pmask & slot_ok)

    grads = {}
    g_m = np.zeros_like(m_ij)                                         # dL/dm_ij (after the gate)
    g_rel = np.zeros_like(rel)
    g_feats = np.zeros_like(feats)

    # ------------------------------------------------------------ node update :319-337, reversed
    if cfg["update_feats"]:
        mm = m_ij if pmask is None else np.where(pmask[..., None], m_ij, 0.0)
        if cfg["m_pool_method"] == "mean":
            if has_node_mask:
                cnt = pmask.sum(-1, keepdims=True).astype(f8)
                inv = np.where(cnt == 0, 0.0, 1.0 / np.maximum(cnt, 1e-8))
            else:
                inv = np.full((b, n, 1), 1.0 / J)
        else:
            inv = np.ones((b, n, 1))
        m_i = mm.sum(2) * inv
        normed = layer_norm(feats, P["node_norm.weight"], P["node_norm.bias"]) if cfg["norm_feats"] else feats
        node_in = np.concatenate([normed, m_i], axis=-1)
        h1pre = linear(node_in, P["node_mlp.0.weight"], P["node_mlp.0.bias"])
        h1 = silu(h1pre)
        g_h1, grads["node_mlp.3.weight"], grads["node_mlp.3.bias"] = _linear_bwd(go, h1, P["node_mlp.3.weight"])
        g_h1pre = g_h1 * dsilu(h1pre)
        g_node_in, grads["node_mlp.0.weight"], grads["node_mlp.0.bias"] = _linear_bwd(g_h1pre, node_in, P["node_mlp.0.weight"])
        g_normed, g_mi = g_node_in[..., :d], g_node_in[..., d:]
        g_feats += go                                                 # residual :337
        if cfg["norm_feats"]:
            mu = feats.mean(-1, keepdims=True)
            var = ((feats - mu) ** 2).mean(-1, keepdims=True)
            rstd = 1.0 / np.sqrt(var + 1e-5)
            xhat = (feats - mu) * rstd
            grads["node_norm.weight"] = (g_normed * xhat).reshape(-1, d).sum(0)
            grads["node_norm.bias"] = g_normed.reshape(-1, d).sum(0)
            gy = g_normed * P["node_norm.weight"]
            g_feats += rstd * (gy - gy.mean(-1, keepdims=True) - xhat * (gy * xhat).mean(-1, keepdims=True))
        else:
            g_feats += g_normed
        g_mm = np.broadcast_to((g_mi * inv)[:, :, None, :], m_ij.shape)
        g_m += g_mm if pmask is None else np.where(pmask[..., None], g_mm, 0.0)
    else:
        g_feats += go                                                 # :339

    # ------------------------------------------------------------ coordinate update :302-317, reversed
    g_coors = gxo.copy()                                              # x_i' = x_i + ...
    if cfg["update_coors"]:
        W3, b3, W4, b4 = (P["coors_mlp.0.weight"], P["coors_mlp.0.bias"], P["coors_mlp.3.weight"], P["coors_mlp.3.bias"])
        t = linear(m_ij, W3, b3)
        c = silu(t)
        w0 = linear(c, W4, b4)[..., 0]
        rel_n = rel
        if cfg["norm_coors"]:
            nrm = np.sqrt((rel ** 2).sum(-1, keepdims=True))
            den = np.maximum(nrm, 1e-8)
            scale = P["coors_norm.scale"]
            rel_n = rel / den * scale
        w1 = w0 if pmask is None else np.where(pmask, w0, 0.0)
        cv = cfg["coor_weights_clamp_value"]
        w2 = w1 if cv is None else np.clip(w1, -cv, cv)
        g_y = np.broadcast_to(gxo[:, :, None, :], rel.shape)          # y_ij = w2 * rel_n summed over j
        g_w2 = (g_y * rel_n).sum(-1)
        g_reln = g_y * w2[..., None]
        if cfg["norm_coors"]:
            grads["coors_norm.scale"] = np.array([(g_reln * rel / den).sum()])
            g_rel += g_reln * scale / den
            g_den = -(g_reln * rel).sum(-1, keepdims=True) * scale / den ** 2
            g_nrm = np.where(nrm >= 1e-8, g_den, 0.0)                 # clamp(min=eps)
            g_rel += np.where(nrm > 0, g_nrm * rel / np.where(nrm > 0, nrm, 1.0), 0.0)
        else:
            g_rel += g_reln
        g_w1 = g_w2 if cv is None else np.where((w1 >= -cv) & (w1 <= cv), g_w2, 0.0)
        g_w0 = g_w1 if pmask is None else np.where(pmask, g_w1, 0.0)
        g_c, grads["coors_mlp.3.weight"], grads["coors_mlp.3.bias"] = _linear_bwd(g_w0[..., None], c, W4)
        g_t = g_c * dsilu(t)
        g_m3, grads["coors_mlp.0.weight"], grads["coors_mlp.0.bias"] = _linear_bwd(g_t, m_ij, W3)
        g_m = g_m + g_m3

    # ------------------------------------------------------------ gate + edge MLP :287-290, reversed
    if cfg["soft_edges"]:
        g_gate = (g_m * s2).sum(-1, keepdims=True)
        g_gz = g_gate * gate * (1.0 - gate)
        g_s2g, grads["edge_gate.0.weight"], grads["edge_gate.0.bias"] = _linear_bwd(g_gz, s2, P["edge_gate.0.weight"])
        g_s2 = g_m * gate + g_s2g
    else:
        g_s2 = g_m
    g_pre2 = g_s2 * dsilu(pre2)
    g_hid, grads["edge_mlp.3.weight"], grads["edge_mlp.3.bias"] = _linear_bwd(g_pre2, hid, W2)
    g_pre1 = g_hid * dsilu(pre1)
    g_ein, grads["edge_mlp.0.weight"], grads["edge_mlp.0.bias"] = _linear_bwd(g_pre1, edge_in, W1)

    # ------------------------------------------------------------ edge-input assembly :270-285, reversed
    g_hi, g_hj = g_ein[..., :d], g_ein[..., d:2 * d]
    g_df = g_ein[..., 2 * d:2 * d + 2 * F + 1]
    g_e = g_ein[..., 2 * d + 2 * F + 1:]
    g_feats += g_hi.sum(2)
    np.add.at(g_feats, (np.broadcast_to(bidx, jidx.shape), jidx), g_hj)
    g_dist = g_df[..., 2 * F].copy()
    for q in range(F):
        sc = 2.0 ** q
        g_dist += g_df[..., q] * np.cos(dist / sc) / sc - g_df[..., F + q] * np.sin(dist / sc) / sc
    g_rel += 2.0 * rel * g_dist[..., None]
    g_coors += g_rel.sum(2)
    np.add.at(g_coors, (np.broadcast_to(bidx, jidx.shape), jidx), -g_rel)
    g_edges = None
    if edges is not None:
        g_edges = np.zeros_like(edges)
        np.add.at(g_edges, (np.broadcast_to(bidx, jidx.shape), np.broadcast_to(iidx, jidx.shape), jidx), g_e)
    return dict(feats=g_feats, coors=g_coors, edges=g_edges, params=grads)


def egnn_network_backward(params, cfg, feats, coors, adj_mat, edges, mask, g_feats_out, g_coors_out):
    """Gradient of `EGNN_Network.forward` (without global attention): layers in reverse, then the embedding
    look-ups of egnn_pytorch.py:401-411 and :430-432.  Returns dict(coors, feats (None for token input), edges
    (None for token / absent input), params {full state-dict key: grad})."""
    from .egnn_oracle import egnn_layer_forward
    f8 = np.float64
    P = {k: np.asarray(v, dtype=f8) for k, v in params.items()}
    coors = np.asarray(coors, dtype=f8)
    b = np.asarray(feats).shape[0]
    tokens = None
    if cfg["num_tokens"] is not None:
        tokens = np.asarray(feats).astype(np.int64)
        h = P["token_emb.weight"][tokens]
    else:
        h = np.asarray(feats, dtype=f8)
    n = h.shape[1]
    if cfg["num_positions"] is not None:
        h = h + P["pos_emb.weight"][:n][None]
    edge_tokens = None
    if edges is not None and cfg["num_edge_tokens"] is not None:
        edge_tokens = np.asarray(edges).astype(np.int64)
        edges = P["edge_emb.weight"][edge_tokens]
    elif edges is not None:
        edges = np.asarray(edges, dtype=f8)
    labels = None
    n_cont = 0 if edges is None else edges.shape[-1]
    if cfg["num_adj_degrees"] is not None:
        adj_mat, labels = adjacency_degrees(adj_mat, cfg["num_adj_degrees"], b)
        if cfg["adj_dim"] > 0:
            adj_e = P["adj_emb.weight"][labels]
            edges = adj_e if edges is None else np.concatenate([edges, adj_e], axis=-1)
    layer_params, states = [], [(h, coors)]
    for l in range(cfg["depth"]):
        prefix = f"layers.{l}.1."
        lp = {k[len(prefix):]: v for k, v in P.items() if k.startswith(prefix)}
        layer_params.append(lp)
        h, coors = egnn_layer_forward(lp, cfg["layer"], h, coors, edges=edges, mask=mask, adj_mat=adj_mat)
        states.append((h, coors))
    grads = {}
    gh, gx = np.asarray(g_feats_out, dtype=f8), np.asarray(g_coors_out, dtype=f8)
    g_edges = None if edges is None else np.zeros_like(edges)
    for l in reversed(range(cfg["depth"])):
        hin, xin = states[l]
        r = egnn_layer_backward(layer_params[l], cfg["layer"], hin, xin, edges, mask, adj_mat, gh, gx)
        gh, gx = r["feats"], r["coors"]
        if g_edges is not None:
            g_edges += r["edges"]
        for k, v in r["params"].items():
            grads[f"layers.{l}.1.{k}"] = v
    if labels is not None and cfg["adj_dim"] > 0:
        ga = np.zeros_like(P["adj_emb.weight"])
        np.add.at(ga, labels, g_edges[..., n_cont:])
        grads["adj_emb.weight"] = ga
    g_edges_in = None
    if n_cont > 0:
        g_cont = g_edges[..., :n_cont]
        if edge_tokens is not None:
            ge = np.zeros_like(P["edge_emb.weight"])
            np.add.at(ge, edge_tokens, g_cont)
            grads["edge_emb.weight"] = ge
        else:
            g_edges_in = g_cont
    if cfg["num_positions"] is not None:
        gp = np.zeros_like(P["pos_emb.weight"])
        gp[:n] = gh.sum(0)
        grads["pos_emb.weight"] = gp
    g_feats_in = None
    if tokens is not None:
        gt = np.zeros_like(P["token_emb.weight"])
        np.add.at(gt, tokens, gh)
        grads["token_emb.weight"] = gt
    else:
        g_feats_in = gh
    return dict(feats=g_feats_in, coors=gx, edges=g_edges_in, params=grads)
