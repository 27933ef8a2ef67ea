"""Two-GPU checks of egnn_pytorch_b200/parallel.py with the CUDA modules over NCCL (skipped on a 1-GPU box;
the same logic is covered on CPU with gloo in tests/test_multi_rank_cpu.py)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
import util

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        from egnn_pytorch_b200 import parallel
        # ---- batch sharding: independent graphs, no data-path collective; outputs gathered for the check
        case = cases.build_case(cases.SPECS["knn_edges_mask"])          # B = 3
        mod = util.make_module(case, torch.float32, device=dev)
        ins = {k: util.to_torch(v, torch.float32, dev) for k, v in case["inputs"].items()}
        f, x = parallel.batch_sharded_call(lambda feats, coors, edges, mask: mod(feats, coors, edges, mask=mask), ins, batch=3)
        want = cases.run_oracle(case)
        util.assert_close(f, want[0], atol=2e-5, rtol=1e-4, what="batch-sharded feats")
        util.assert_close(x, want[1], atol=2e-5, rtol=1e-4, what="batch-sharded coors")
        # ---- one graph, i-rows sharded: a single all-gather of [coors | feats] per layer, rows stay local
        for name in ("knn_norm_coors", "dense_mask_padded"):
            case = cases.build_case(cases.SPECS[name])
            mod = util.make_module(case, torch.float32, device=dev)
            ins = {k: util.to_torch(v, torch.float32, dev) for k, v in case["inputs"].items()}
            n = ins["feats"].shape[1]
            r0, r1 = parallel.shard_range(n, rank, world)
            f_loc, x_loc = parallel.row_sharded_layer_call(
                lambda fa, xa, rows, **kw: mod(fa, xa, kw.get("edges"), mask=kw.get("mask"), _rows=rows),
                ins["feats"][:, r0:r1].contiguous(), ins["coors"][:, r0:r1].contiguous(), n, edges=ins.get("edges"),
                mask=ins.get("mask"))
            want = cases.run_oracle(case)
            util.assert_close(f_loc, want[0][:, r0:r1], atol=2e-5, rtol=1e-4, what=f"row-sharded {name} feats")
            util.assert_close(x_loc, want[1][:, r0:r1], atol=2e-5, rtol=1e-4, what=f"row-sharded {name} coors")
        dist.barrier()
        q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_batch_and_row_sharding_two_gpus():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
