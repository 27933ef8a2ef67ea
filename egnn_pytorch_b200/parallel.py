"""Multi-GPU use of the layer.  One process per GPU, torch.distributed for
the plumbing.

* Batches of independent graphs shard over ranks with NO data-path collective
  (`batch_shard`, `batch_sharded_call`); outputs are gathered only if the caller asks.
* One huge graph: contiguous blocks of i-rows per rank.  Every rank needs all coordinates and all
  node features (for the per-node B_j table), so each layer starts with ONE all-gather of
  [coors | feats] and then evaluates only its own rows (`row_sharded_layer_call`); top-k, masks and
  both j-reductions are row-local, so nothing is reduced across ranks.

* Training on batch shards: every rank differentiates the loss of its own graphs; the parameter gradients are
  the only thing exchanged -- `allreduce_gradients` sums them in a few flat buckets (one collective per
  bucket: the whole EGNN(512) layer is 12.8 MB in fp32, i.e. one launch-latency-bound all-reduce over NVLink).

The compute callable is injected, so the partition/exchange logic is testable on CPU with gloo
(tests/test_multi_rank_cpu.py drives it with the oracle); on the GPU box it is the CUDA module.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int):
    """Contiguous, balanced [begin, end) of `total` units for `rank` (first ranks get the remainder)."""
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def batch_shard(tensors: dict, rank: int, world: int, batch: int):
    """Slice every tensor whose leading dimension is the batch; 2-D adjacency etc. pass through."""
    b0, b1 = shard_range(batch, rank, world)
    out = {}
    for k, v in tensors.items():
        if torch.is_tensor(v) and v.dim() >= 1 and v.shape[0] == batch and not (k == "adj_mat" and v.dim() == 2):
            out[k] = v[b0:b1]
        else:
            out[k] = v
    return out, (b0, b1)


def _all_gather_var(x: torch.Tensor, sizes, group=None):
    """all_gather of per-rank blocks with different leading sizes (pads to the maximum)."""
    world = dist.get_world_size(group)
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    pad[: x.shape[0]] = x
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)


def batch_sharded_call(fn, tensors: dict, batch: int, gather: bool = True, group=None):
    """Run `fn(**shard)` on this rank's graphs; optionally all-gather the (feats, coors) outputs."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    shard, (b0, b1) = batch_shard(tensors, rank, world, batch)
    outs = fn(**shard) if b1 > b0 else None
    if not gather:
        return outs
    sizes = [shard_range(batch, r, world)[1] - shard_range(batch, r, world)[0] for r in range(world)]
    return tuple(_all_gather_var(o, sizes, group) for o in outs)


def row_sharded_layer_call(layer_fn, feats_local, coors_local, n_total: int, group=None, **kw):
    """One layer of a row-sharded single graph.

    feats_local [B, R_rank, dim], coors_local [B, R_rank, C] hold this rank's node block.  The single
    exchange step all-gathers both along the node axis; `layer_fn(feats_all, coors_all, rows=(r0, r1), **kw)`
    must return full-size outputs of which only rows r0:r1 are meaningful.  Returns the local blocks."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    sizes = [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
    r0, r1 = shard_range(n_total, rank, world)
    # one payload: [coors | feats] along the channel axis, node axis first for the gather
    wide = torch.promote_types(coors_local.dtype, feats_local.dtype)      # bf16 feats ride in fp32: exact
    payload = torch.cat([coors_local.to(wide), feats_local.to(wide)], dim=-1).transpose(0, 1).contiguous()
    full = _all_gather_var(payload, sizes, group).transpose(0, 1)
    c = coors_local.shape[-1]
    coors_all = full[..., :c].to(coors_local.dtype).contiguous()
    feats_all = full[..., c:].to(feats_local.dtype).contiguous()
    f_out, x_out = layer_fn(feats_all, coors_all, rows=(r0, r1), **kw)
    return f_out[:, r0:r1], x_out[:, r0:r1]


def allreduce_gradients(params, group=None, average: bool = False, bucket_bytes: int = 64 << 20):
    """Sum (or average) `.grad` of the given parameters over the ranks, in place.

    Gradients are packed into flat buckets of at most `bucket_bytes` per dtype so that a layer costs one collective
    instead of one per tensor.  Parameters whose `.grad` is None on this rank (e.g. a rank with an empty shard)
    contribute zeros -- every rank must pass the same parameter list in the same order."""
    world = dist.get_world_size(group)
    params = [p for p in params if p.requires_grad]
    by_dtype = {}
    for p in params:
        by_dtype.setdefault((p.dtype, p.device), []).append(p)
    for (dtype, device), plist in by_dtype.items():
        bucket, size = [], 0
        buckets = []
        for p in plist:
            nbytes = p.numel() * p.element_size()
            if bucket and size + nbytes > bucket_bytes:
                buckets.append(bucket)
                bucket, size = [], 0
            bucket.append(p)
            size += nbytes
        if bucket:
            buckets.append(bucket)
        for bucket in buckets:
            flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in bucket])
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            if average:
                flat /= world
            off = 0
            for p in bucket:
                g = flat[off:off + p.numel()].view_as(p)
                if p.grad is None:
                    p.grad = g.clone()
                else:
                    p.grad.copy_(g)
                off += p.numel()
