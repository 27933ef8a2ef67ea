"""World-size-2 gloo tests of the multi-GPU host logic (egnn_pytorch_b200/parallel.py) on CPU.
The compute callable is the oracle here (no GPU in this container); on the GPU box the same
functions drive the CUDA modules (tests/test_gpu_multi.py, bench.py --gpus N)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases


def _oracle_layer(case):
    def fn(feats, coors, edges=None, mask=None, adj_mat=None, rows=None):
        np_ = lambda t: None if t is None else t.numpy()
        f, x = cases.O.egnn_layer_forward(case["params"], case["cfg"], np_(feats), np_(coors), edges=np_(edges),
                                          mask=np_(mask), adj_mat=np_(adj_mat), rows=rows)
        if rows is not None:          # contract of row_sharded_layer_call: full-size outputs
            F = np.zeros(feats.shape); X = np.zeros(coors.shape)
            F[:, rows[0]:rows[1]] = f; X[:, rows[0]:rows[1]] = x
            f, x = F, X
        return torch.from_numpy(f), torch.from_numpy(x)
    return fn


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from egnn_pytorch_b200 import parallel
        # --- batch sharding: independent graphs, no data-path collective
        case = cases.build_case(cases.SPECS["knn_edges_mask"])          # B = 3 -> shards of 2 and 1
        ins = {k: torch.from_numpy(np.asarray(v)) for k, v in case["inputs"].items()}
        fn = _oracle_layer(case)
        f, x = parallel.batch_sharded_call(lambda **kw: fn(**kw), ins, batch=3, gather=True)
        want = cases.run_oracle(case)
        assert np.abs(f.numpy() - want[0]).max() < 1e-12 and np.abs(x.numpy() - want[1]).max() < 1e-12
        # fewer graphs than ranks: every rank raises the same error BEFORE any collective (no hang on the empty shard)
        one = {k: (v[:1] if torch.is_tensor(v) and v.shape[0] == 3 else v) for k, v in ins.items()}
        try:
            parallel.batch_sharded_call(lambda **kw: fn(**kw), one, batch=1, gather=True)
            raise AssertionError("expected ValueError for batch < world")
        except ValueError:
            pass
        # ... and without the gather the rank with the empty shard simply has nothing to do
        outs = parallel.batch_sharded_call(lambda **kw: fn(**kw), one, batch=1, gather=False)
        assert (outs is None) == (rank == 1)
        # --- row sharding of one graph: a single all-gather of [coors | feats], rows stay local
        case = cases.build_case(cases.SPECS["knn_norm_coors"])          # B = 1, N = 40
        ins = {k: torch.from_numpy(np.asarray(v)) for k, v in case["inputs"].items()}
        n = ins["feats"].shape[1]
        r0, r1 = parallel.shard_range(n, rank, world)
        fn = _oracle_layer(case)
        f_loc, x_loc = parallel.row_sharded_layer_call(
            lambda fa, xa, rows, **kw: fn(fa, xa, rows=rows, **kw), ins["feats"][:, r0:r1], ins["coors"][:, r0:r1], n,
            edges=ins.get("edges"), mask=ins.get("mask"))
        want = cases.run_oracle(case)
        assert np.abs(f_loc.numpy() - want[0][:, r0:r1]).max() < 1e-12
        assert np.abs(x_loc.numpy() - want[1][:, r0:r1]).max() < 1e-12
        # --- batch-sharded training: each rank differentiates its own graphs (compute = backward oracle here),
        #     parameter gradients are summed with one bucketed all-reduce, input gradients stay local
        case = cases.build_case(cases.SPECS["knn_edges_mask"])          # B = 3 -> shards of 2 and 1
        gf, gx = cases.upstream_grads(case)
        full = cases.flatten_grads(cases.run_oracle_grad(case))
        b0, b1 = parallel.shard_range(3, rank, world)
        from oracle import egnn_oracle_grad as G
        ins = case["inputs"]
        part = G.egnn_layer_backward(case["params"], case["cfg"], ins["feats"][b0:b1], ins["coors"][b0:b1],
                                     ins["edges"][b0:b1], ins["mask"][b0:b1], None, gf[b0:b1], gx[b0:b1])
        params = {k: torch.nn.Parameter(torch.from_numpy(np.asarray(v, np.float64))) for k, v in case["params"].items()}
        for k, p_ in params.items():
            p_.grad = torch.from_numpy(part["params"][k]).reshape(p_.shape).clone()
        parallel.allreduce_gradients([params[k] for k in sorted(params)], bucket_bytes=4096)   # several buckets
        for k, p_ in params.items():
            assert np.abs(p_.grad.numpy() - full[f"p.{k}"].reshape(p_.shape)).max() < 1e-11, k
        assert np.abs(part["feats"] - full["in.feats"][b0:b1]).max() < 1e-12      # inputs: no exchange needed
        assert np.abs(part["coors"] - full["in.coors"][b0:b1]).max() < 1e-12
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_batch_and_row_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
