"""CPU oracle for the EGNN forward hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

A numpy restatement of lucidrains/egnn-pytorch's dense-tensor message-passing layer
(`EGNN.forward`, reference egnn_pytorch/egnn_pytorch.py:224-341) and of its network wrapper
(`EGNN_Network.forward`, egnn_pytorch.py:390-454).  It keeps the reference's own
formulation (broadcast, concat, Linear over the concatenated edge input) on purpose: the
CUDA product uses an algebraically split form, so agreeing with this file is an
independent check, not a tautology.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
leg may import this module.  The product package `egnn_pytorch_b200` never does and fails
loudly without its CUDA library.

Pinning: the reference publishes no golden vectors, so this oracle is
pinned against outputs of the reference itself, generated in the build container by
`tests/golden/make_golden.py` (imports /root/reference read-only) and committed under
`tests/golden/*.npz`; `tests/test_oracle_golden.py` replays every fixture.

Conventions
* parameters are passed as a flat dict keyed by the reference's `state_dict()` names
  (`edge_mlp.0.weight`, `node_mlp.3.bias`, `coors_norm.scale`, ...), weights row-major
  `[out, in]` exactly as `nn.Linear` stores them;
* everything is evaluated in `dtype` (float64 by default, float32 for the timed baseline);
* rows of the (i, j) pair grid are processed in chunks of `row_chunk` so the
  `[rows, J, 2E]` hidden tensor stays bounded (the reference materialises all of it);
* top-k ties are broken towards the lowest neighbour index (`torch.topk` leaves the order
  unspecified; the CUDA select kernel uses the same rule).
"""
from __future__ import annotations

import numpy as np

# ----------------------------------------------------------------------------- helpers


def silu(x):
    """x * sigmoid(x) (reference `SiLU`, egnn_pytorch.py:56-60)."""
    return x / (1.0 + np.exp(-x))


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def linear(x, w, b=None):
    """nn.Linear: x @ w.T + b, w stored [out, in]."""
    y = x @ w.T
    if b is not None:
        y = y + b
    return y


def layer_norm(x, gamma, beta, eps=1e-5):
    """nn.LayerNorm over the last axis (biased variance), egnn_pytorch.py:191."""
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * gamma + beta


def fourier_features_of(d, num_encodings):
    """`fourier_encode_dist` (egnn_pytorch.py:34-41): [sin(d/2^k) k<F | cos(d/2^k) k<F | d]."""
    scales = (2.0 ** np.arange(num_encodings)).astype(d.dtype)
    scaled = d[..., None] / scales
    return np.concatenate([np.sin(scaled), np.cos(scaled), d[..., None]], axis=-1)


def smallest_k(ranking, k):
    """Indices and values of the k smallest entries along the last axis, ascending;
    ties go to the lowest index (stable sort).  Stands in for
    `ranking.topk(k, largest=False)` at egnn_pytorch.py:258."""
    order = np.argsort(ranking, axis=-1, kind="stable")[..., :k]
    return np.take_along_axis(ranking, order, axis=-1), order


DEFAULT_CFG = dict(
    dim=None, edge_dim=0, m_dim=16, fourier_features=0, num_nearest_neighbors=0,
    dropout=0.0, init_eps=1e-3, norm_feats=False, norm_coors=False,
    norm_coors_scale_init=1e-2, update_feats=True, update_coors=True,
    only_sparse_neighbors=False, valid_radius=float("inf"), m_pool_method="sum",
    soft_edges=False, coor_weights_clamp_value=None,
)


def layer_cfg(**kw):
    """Constructor keyword set of `EGNN.__init__` (egnn_pytorch.py:149-168) with its defaults."""
    cfg = dict(DEFAULT_CFG)
    unknown = set(kw) - set(cfg)
    assert not unknown, f"unknown EGNN kwargs {unknown}"
    cfg.update(kw)
    assert cfg["dim"] is not None
    assert cfg["m_pool_method"] in {"sum", "mean"}          # egnn_pytorch.py:170
    assert cfg["update_feats"] or cfg["update_coors"]        # egnn_pytorch.py:171
    return cfg


def edge_input_dim(cfg):
    """egnn_pytorch.py:175."""
    return cfg["fourier_features"] * 2 + cfg["dim"] * 2 + cfg["edge_dim"] + 1


# ----------------------------------------------------------------------------- the layer


def neighbour_selection(cfg, coors, mask, adj_mat):
    """Ranking + top-k of egnn_pytorch.py:237-260.

    Returns (nbr_idx [B,N,k] int64, nbhd_mask [B,N,k] bool, k).  Called only when
    `num_nearest_neighbors > 0 or only_sparse_neighbors`."""
    b, n, _ = coors.shape
    num_nearest = cfg["num_nearest_neighbors"]
    valid_radius = cfg["valid_radius"]
    rel = coors[:, :, None, :] - coors[:, None, :, :]
    ranking = (rel ** 2).sum(-1)                                     # :233, :238
    if mask is not None:
        rank_mask = mask[:, :, None] & mask[:, None, :]
        ranking = np.where(rank_mask, ranking, np.asarray(1e5, ranking.dtype))   # :240-242
    if adj_mat is not None:
        adj = np.asarray(adj_mat).astype(bool)
        if adj.ndim == 2:
            adj = np.broadcast_to(adj, (b, n, n))                    # :245-246
        if cfg["only_sparse_neighbors"]:
            num_nearest = int(adj.sum(-1).max())                     # :249 (diagonal still counted)
            valid_radius = 0
        eye = np.eye(n, dtype=bool)[None]
        adj = adj & ~eye                                             # :254
        ranking = np.where(eye, np.asarray(-1.0, ranking.dtype), ranking)   # :255
        ranking = np.where(adj, np.asarray(0.0, ranking.dtype), ranking)    # :256
    assert 0 < num_nearest <= n, "topk needs 0 < k <= N"
    vals, idx = smallest_k(ranking, num_nearest)                     # :258
    return idx, vals <= valid_radius, num_nearest                    # :260


def egnn_layer_forward(params, cfg, feats, coors, edges=None, mask=None, adj_mat=None,
                       dtype=np.float64, row_chunk=64, rows=None):
    """One EGNN layer, reference egnn_pytorch.py:224-341.

    feats [B,N,dim], coors [B,N,C], edges [B,N,N,edge_dim] | None, mask [B,N] bool | None,
    adj_mat [N,N] | [B,N,N] bool | None.  Returns (feats_out [B,N,dim], coors_out [B,N,C]).

    `rows=(r0, r1)` restricts the evaluated i-rows (used by the row-sharded multi-GPU test and
    by the bounded cpu_baseline sample); the returned arrays then cover only rows r0:r1.
    """
    P = {k: np.asarray(v, dtype=dtype) for k, v in params.items()}
    feats = np.asarray(feats, dtype=dtype)
    coors = np.asarray(coors, dtype=dtype)
    if edges is not None:
        edges = np.asarray(edges, dtype=dtype)
    if mask is not None:
        mask = np.asarray(mask).astype(bool)
    b, n, d = feats.shape
    F = cfg["fourier_features"]
    use_nearest = cfg["num_nearest_neighbors"] > 0 or cfg["only_sparse_neighbors"]   # :230

    nbr_idx = nbhd_mask = None
    if use_nearest:
        nbr_idx, nbhd_mask, _ = neighbour_selection(cfg, coors, mask, adj_mat)

    r0, r1 = (0, n) if rows is None else rows
    feats_out = np.empty((b, r1 - r0, d), dtype=dtype)
    coors_out = np.empty((b, r1 - r0, coors.shape[-1]), dtype=dtype)
    bidx = np.arange(b)[:, None, None]

    for s in range(r0, r1, row_chunk):
        e = min(s + row_chunk, r1)
        xi = coors[:, s:e]                                           # [B,R,C]
        hi = feats[:, s:e]
        if use_nearest:
            jidx = nbr_idx[:, s:e]                                   # [B,R,k]
            xj = coors[bidx, jidx]                                   # gather, :262 via :18-32
            hj = feats[bidx, jidx]                                   # :275
            eij = None if edges is None else edges[bidx, np.arange(s, e)[None, :, None], jidx]  # :266
        else:
            xj = coors[:, None, :, :]
            hj = np.broadcast_to(feats[:, None, :, :], (b, e - s, n, d))             # :277
            eij = None if edges is None else edges[:, s:e]
        rel = xi[:, :, None, :] - xj                                 # :232  x_i - x_j
        dist = (rel ** 2).sum(-1)                                    # :233  squared distance
        dfeat = fourier_features_of(dist, F) if F > 0 else dist[..., None]          # :270-272
        J = rel.shape[2]
        hi_b = np.broadcast_to(hi[:, :, None, :], (b, e - s, J, d))  # :279-280
        parts = [hi_b, hj, dfeat] + ([eij] if eij is not None else [])
        edge_in = np.concatenate(parts, axis=-1)                     # :282-285  [h_i | h_j | d | e]

        hid = silu(linear(edge_in, P["edge_mlp.0.weight"], P["edge_mlp.0.bias"]))   # :287 (:179-181)
        m_ij = silu(linear(hid, P["edge_mlp.3.weight"], P["edge_mlp.3.bias"]))      # :182-183
        del hid, edge_in
        if cfg["soft_edges"]:
            m_ij = m_ij * sigmoid(linear(m_ij, P["edge_gate.0.weight"], P["edge_gate.0.bias"]))  # :289-290

        pmask = None
        if mask is not None:                                         # :292-300
            mi = mask[:, s:e, None]
            if use_nearest:
                pmask = (mi & mask[bidx, jidx]) & nbhd_mask[:, s:e]
            else:
                pmask = mi & mask[:, None, :]

        if cfg["update_coors"]:                                      # :302-315
            cw = linear(silu(linear(m_ij, P["coors_mlp.0.weight"], P["coors_mlp.0.bias"])),
                        P["coors_mlp.3.weight"], P["coors_mlp.3.bias"])[..., 0]
            rel_n = rel
            if cfg["norm_coors"]:                                    # CoorsNorm :67-77
                nrm = np.sqrt((rel ** 2).sum(-1, keepdims=True))
                rel_n = rel / np.maximum(nrm, 1e-8) * P["coors_norm.scale"]
            if pmask is not None:
                cw = np.where(pmask, cw, 0.0)                        # :309
            cv = cfg["coor_weights_clamp_value"]
            if cv is not None:
                cw = np.clip(cw, -cv, cv)                            # :313
            coors_out[:, s - r0:e - r0] = (cw[..., None] * rel_n).sum(2) + xi       # :315
        else:
            coors_out[:, s - r0:e - r0] = xi                         # :317

        if cfg["update_feats"]:                                      # :319-337
            if pmask is not None:
                m_ij = np.where(pmask[..., None], m_ij, 0.0)         # :322
            if cfg["m_pool_method"] == "mean":
                if pmask is not None:                                # masked mean, safe_div :13-16
                    cnt = pmask.sum(-1, keepdims=True).astype(dtype)
                    m_i = m_ij.sum(2) / np.maximum(cnt, 1e-8)
                    m_i = np.where(cnt == 0, 0.0, m_i)
                else:
                    m_i = m_ij.mean(2)                               # :330
            else:
                m_i = m_ij.sum(2)                                    # :333
            normed = hi
            if cfg["norm_feats"]:
                normed = layer_norm(hi, P["node_norm.weight"], P["node_norm.bias"])  # :335
            node_in = np.concatenate([normed, m_i], axis=-1)
            h1 = silu(linear(node_in, P["node_mlp.0.weight"], P["node_mlp.0.bias"]))
            feats_out[:, s - r0:e - r0] = linear(h1, P["node_mlp.3.weight"], P["node_mlp.3.bias"]) + hi  # :337
        else:
            feats_out[:, s - r0:e - r0] = hi                         # :339

    return feats_out, coors_out


# ----------------------------------------------------------------------------- edge-list mode


def egnn_layer_forward_edge_list(params, cfg, feats, coors, neighbors, edges=None, mask=None, dtype=np.float64):
    """The layer on a caller-supplied edge list -- message passing in the flat
    per-edge form of the reference's PyG layer (`EGNN_Sparse.forward / message / propagate`,
    egnn_pytorch_geometric.py:182-267: per-edge `edge_mlp(cat[x_i, x_j, edge_attr])`, per-edge coordinate
    weight, `aggregate` = scatter-add onto the receiving node, `node_mlp(cat[norm(x), m_i]) + x`), written in the
    DENSE layer's conventions so that it is the same function as `EGNN.forward` restricted to these edges:
    edge-input column order [h_i | h_j | d | e_ij] (egnn_pytorch.py:282-285), rel = x_i - x_j (:232),
    masks / clamp / CoorsNorm / pooling exactly as :292-333.

    `neighbors` int [B,N,k]: neighbors[b,i,s] = j lists the edges j -> i; an entry < 0 is an empty slot and
    contributes to nothing.  `valid_radius` does not apply (it is a property of the top-k ranking, :260).
    Mean pooling: with a node mask the masked mean over existing, unmasked edges (:325-328, safe_div :13-16);
    without one the reference's plain `.mean` over the k slots (:330).

    Deliberately NOT the gather formulation of `egnn_layer_forward`: a flat list of E edges, one row per edge,
    `np.add.at` for the aggregation -- an independent restatement to check the CUDA neighbour-list kernels."""
    P = {k: np.asarray(v, dtype=dtype) for k, v in params.items()}
    feats = np.asarray(feats, dtype=dtype)
    coors = np.asarray(coors, dtype=dtype)
    nb = np.asarray(neighbors).astype(np.int64)
    b, n, d = feats.shape
    k = nb.shape[-1]
    assert nb.shape[:2] == (b, n) and (nb < n).all()
    F = cfg["fourier_features"]
    eb, ei, _ = np.nonzero(nb >= 0)                                  # one row per existing edge
    ej = nb[nb >= 0]
    x_i, x_j, h_i, h_j = coors[eb, ei], coors[eb, ej], feats[eb, ei], feats[eb, ej]
    rel = x_i - x_j                                                  # geometric.py:196 in the dense sign convention (:232)
    dist = (rel ** 2).sum(-1)                                        # geometric.py:197
    dfeat = fourier_features_of(dist, F) if F > 0 else dist[:, None]  # geometric.py:199-201
    parts = [h_i, h_j, dfeat]
    if edges is not None:
        parts.append(np.asarray(edges, dtype=dtype)[eb, ei, ej])
    edge_in = np.concatenate(parts, axis=-1)
    m = silu(linear(silu(linear(edge_in, P["edge_mlp.0.weight"], P["edge_mlp.0.bias"])),
                    P["edge_mlp.3.weight"], P["edge_mlp.3.bias"]))   # message(), geometric.py:214-216
    if cfg["soft_edges"]:
        m = m * sigmoid(linear(m, P["edge_gate.0.weight"], P["edge_gate.0.bias"]))   # geometric.py:257-258
    live = None
    if mask is not None:
        mk = np.asarray(mask).astype(bool)
        live = mk[eb, ei] & mk[eb, ej]                               # egnn_pytorch.py:292-297
    coors_out = coors.copy()
    if cfg["update_coors"]:
        w = linear(silu(linear(m, P["coors_mlp.0.weight"], P["coors_mlp.0.bias"])),
                   P["coors_mlp.3.weight"], P["coors_mlp.3.bias"])[:, 0]            # geometric.py:238
        if live is not None:
            w = np.where(live, w, 0.0)                               # egnn_pytorch.py:309
        cv = cfg["coor_weights_clamp_value"]
        if cv is not None:
            w = np.clip(w, -cv, cv)                                  # :313
        rel_n = rel
        if cfg["norm_coors"]:                                        # geometric.py:246, CoorsNorm
            rel_n = rel / np.maximum(np.sqrt((rel ** 2).sum(-1, keepdims=True)), 1e-8) * P["coors_norm.scale"]
        np.add.at(coors_out, (eb, ei), w[:, None] * rel_n)           # aggregate('add') + coors, geometric.py:248-249
    feats_out = feats.copy()
    if cfg["update_feats"]:
        mm = m if live is None else np.where(live[:, None], m, 0.0)  # :322
        m_i = np.zeros((b, n, m.shape[-1]), dtype=dtype)
        np.add.at(m_i, (eb, ei), mm)                                 # aggregate, geometric.py:259
        if cfg["m_pool_method"] == "mean":
            if live is not None:
                cnt = np.zeros((b, n, 1), dtype=dtype)
                np.add.at(cnt, (eb, ei), live[:, None].astype(dtype))
                m_i = np.where(cnt == 0, 0.0, m_i / np.maximum(cnt, 1e-8))
            else:
                m_i = m_i / k
        normed = layer_norm(feats, P["node_norm.weight"], P["node_norm.bias"]) if cfg["norm_feats"] else feats
        h1 = silu(linear(np.concatenate([normed, m_i], axis=-1), P["node_mlp.0.weight"], P["node_mlp.0.bias"]))
        feats_out = linear(h1, P["node_mlp.3.weight"], P["node_mlp.3.bias"]) + feats   # geometric.py:261-263
    return feats_out, coors_out


# ----------------------------------------------------------------------------- the network


def network_cfg(*, depth, dim, num_tokens=None, num_edge_tokens=None, num_positions=None,
                edge_dim=0, num_adj_degrees=None, adj_dim=0, global_linear_attn_every=0,
                global_linear_attn_heads=8, global_linear_attn_dim_head=64, num_global_tokens=4, **egnn_kwargs):
    """Keyword set of `EGNN_Network.__init__` (egnn_pytorch.py:344-388)."""
    assert not (num_adj_degrees is not None and num_adj_degrees < 1)               # :362
    has_edges = edge_dim > 0
    layer_edge = (edge_dim if has_edges else 0) + (adj_dim if num_adj_degrees is not None else 0)   # :372-373, :387
    every = global_linear_attn_every
    return dict(depth=depth, dim=dim, num_tokens=num_tokens, num_edge_tokens=num_edge_tokens,
                num_positions=num_positions, edge_dim=edge_dim, num_adj_degrees=num_adj_degrees,
                adj_dim=adj_dim,
                global_every=every, global_heads=global_linear_attn_heads, global_dim_head=global_linear_attn_dim_head,
                num_global_tokens=num_global_tokens,
                global_layers=[l for l in range(depth) if every > 0 and l % every == 0],            # :381-382
                layer=layer_cfg(dim=dim, edge_dim=layer_edge, norm_feats=True, **egnn_kwargs))


def gelu(x):
    """nn.GELU() (exact, erf form), egnn_pytorch.py:127."""
    from math import sqrt
    try:
        from scipy.special import erf
    except Exception:       # pragma: no cover
        erf = np.vectorize(__import__("math").erf)
    return 0.5 * x * (1.0 + erf(x / sqrt(2.0)))


def attention(P, prefix, x, context, heads, mask=None):
    """`Attention.forward` (egnn_pytorch.py:92-110): softmax(q k^T * dim_head^-0.5) v over `context`, heads split from
    the channel axis, masked keys filled with -finfo.max BEFORE the softmax (so a fully masked row is uniform)."""
    q = linear(x, P[prefix + "to_q.weight"])                                        # :95
    kv = linear(context, P[prefix + "to_kv.weight"])                                # :96
    inner = q.shape[-1]
    k, v = kv[..., :inner], kv[..., inner:]
    b, n = q.shape[:2]
    j = k.shape[1]
    dh = inner // heads
    split = lambda t_, m: t_.reshape(b, m, heads, dh).transpose(0, 2, 1, 3)        # b n (h d) -> b h n d, :98
    q, k, v = split(q, n), split(k, j), split(v, j)
    dots = np.einsum("bhid,bhjd->bhij", q, k) * dh ** -0.5                          # :99
    if mask is not None:
        dots = np.where(np.asarray(mask).astype(bool)[:, None, None, :], dots, -np.finfo(dots.dtype).max)   # :101-104
    dots = dots - dots.max(-1, keepdims=True)
    attn = np.exp(dots)
    attn = attn / attn.sum(-1, keepdims=True)                                       # :106
    out = np.einsum("bhij,bhjd->bhid", attn, v).transpose(0, 2, 1, 3).reshape(b, n, inner)   # :107-109
    return linear(out, P[prefix + "to_out.weight"], P[prefix + "to_out.bias"])     # :110


def global_linear_attention(P, prefix, x, queries, heads, mask=None):
    """`GlobalLinearAttention.forward` (egnn_pytorch.py:132-144): the global tokens attend over the (masked) nodes, the
    nodes attend over the induced tokens, residuals, then a pre-norm GELU feed-forward on the nodes."""
    res_x, res_q = x, queries
    xn = layer_norm(x, P[prefix + "norm_seq.weight"], P[prefix + "norm_seq.bias"])
    qn = layer_norm(queries, P[prefix + "norm_queries.weight"], P[prefix + "norm_queries.bias"])
    induced = attention(P, prefix + "attn1.", qn, xn, heads, mask)                  # :136
    out = attention(P, prefix + "attn2.", xn, induced, heads)                       # :137
    x = out + res_x                                                                 # :139
    queries = induced + res_q                                                       # :140
    y = layer_norm(x, P[prefix + "ff.0.weight"], P[prefix + "ff.0.bias"])
    y = linear(gelu(linear(y, P[prefix + "ff.1.weight"], P[prefix + "ff.1.bias"])), P[prefix + "ff.3.weight"], P[prefix + "ff.3.bias"])
    return y + x, queries                                                           # :142-143


def adjacency_degrees(adj_mat, num_adj_degrees, b):
    """N-th degree adjacency by repeated squaring, egnn_pytorch.py:414-428.

    Returns (expanded adjacency bool [B,N,N], degree labels int64 [B,N,N]).  Each round squares
    the *expanded* matrix, so reach grows 1 -> 2 -> 4 hops."""
    adj = np.asarray(adj_mat).astype(bool)
    if adj.ndim == 2:
        adj = np.broadcast_to(adj, (b,) + adj.shape).copy()
    labels = adj.astype(np.int64)                                                   # :420
    for ind in range(num_adj_degrees - 1):
        degree = ind + 2
        a = adj.astype(np.float64)
        nxt = (a @ a) > 0                                                           # :425
        # (next.float() - adj.float()).bool(): nonzero difference, i.e. next XOR adj
        newly = (nxt.astype(np.float64) - a) != 0                                   # :426
        labels = np.where(newly, degree, labels)                                    # :427
        adj = nxt                                                                   # :428
    return adj, labels


def egnn_network_forward(params, cfg, feats, coors, adj_mat=None, edges=None, mask=None,
                         return_coor_changes=False, dtype=np.float64, row_chunk=64):
    """`EGNN_Network.forward`, egnn_pytorch.py:390-454."""
    P = params
    coors = np.asarray(coors, dtype=dtype)
    b = np.asarray(feats).shape[0]
    if cfg["num_tokens"] is not None:
        feats = np.asarray(P["token_emb.weight"], dtype=dtype)[np.asarray(feats).astype(np.int64)]   # :401-402
    else:
        feats = np.asarray(feats, dtype=dtype)
    if cfg["num_positions"] is not None:
        n = feats.shape[1]
        assert n <= cfg["num_positions"]                                            # :406
        feats = feats + np.asarray(P["pos_emb.weight"], dtype=dtype)[:n][None]      # :407-408
    if edges is not None and cfg["num_edge_tokens"] is not None:
        edges = np.asarray(P["edge_emb.weight"], dtype=dtype)[np.asarray(edges).astype(np.int64)]    # :410-411
    if cfg["num_adj_degrees"] is not None:
        assert adj_mat is not None                                                  # :415
        adj_mat, labels = adjacency_degrees(adj_mat, cfg["num_adj_degrees"], b)
        if cfg["adj_dim"] > 0:
            adj_emb = np.asarray(P["adj_emb.weight"], dtype=dtype)[labels]          # :430-431
            edges = adj_emb if edges is None else np.concatenate(
                [np.asarray(edges, dtype=dtype), adj_emb], axis=-1)                 # :432
    global_tokens = None
    if cfg.get("global_layers"):
        global_tokens = np.broadcast_to(np.asarray(P["global_tokens"], dtype=dtype), (b,) + np.shape(P["global_tokens"]))   # :439-440
    coor_changes = [coors]
    for l in range(cfg["depth"]):
        if l in cfg.get("global_layers", []):                                       # :445-446
            PD = {k: np.asarray(v, dtype=dtype) for k, v in P.items() if k.startswith(f"layers.{l}.0.")}
            feats, global_tokens = global_linear_attention(PD, f"layers.{l}.0.", feats, global_tokens, cfg["global_heads"], mask)
        prefix = f"layers.{l}.1."
        lp = {k[len(prefix):]: v for k, v in P.items() if k.startswith(prefix)}
        feats, coors = egnn_layer_forward(lp, cfg["layer"], feats, coors, edges=edges, mask=mask,
                                          adj_mat=adj_mat, dtype=dtype, row_chunk=row_chunk)   # :448
        coor_changes.append(coors)
    if return_coor_changes:
        return feats, coors, coor_changes
    return feats, coors


# ----------------------------------------------------------------------------- work accounting


def flops_per_pair_reference(cfg):
    """F_ref of SURVEY.md section 8(d): 2*E*H + 2*H*m + 2*m*4m + 2*4m."""
    E = edge_input_dim(cfg)
    H, m = 2 * E, cfg["m_dim"]
    return 2 * E * H + 2 * H * m + 2 * m * 4 * m + 2 * 4 * m
