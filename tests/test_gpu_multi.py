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
        torch.set_grad_enabled(False)                # inference: row ranges are a forward-only feature
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
        # ---- the same row sharding with the library's peer-memory all-gather over NVLink (egnn_comm_*): two layer calls
        #      back to back (exercises the double-buffered epochs), fp32 SIMT and bf16 tensor-core kernels, dense and kNN
        for name, dtype in (("dense_mask_padded", torch.float32), ("knn_norm_coors", torch.float32),
                            ("dense_xavier", torch.bfloat16), ("knn_basic", torch.bfloat16)):
            case = cases.build_case(cases.SPECS[name])
            if dtype == torch.bfloat16 and case["cfg"]["dim"] % 8:
                continue
            mod = util.make_module(case, dtype, device=dev)
            ins = {k: util.to_torch(v, dtype, dev) for k, v in case["inputs"].items()}
            ins["coors"] = ins["coors"].float()
            b, n, d = ins["feats"].shape
            r0, r1 = parallel.shard_range(n, rank, world)
            _, payload = parallel.row_payload_layout(n, ins["coors"].shape[-1], d, ins["feats"].element_size(), b)
            comm = parallel.PeerComm(payload)
            kw = dict(mask=ins.get("mask"))
            if ins.get("edges") is not None:
                kw["edges"] = ins["edges"]
            with torch.no_grad():
                full_f, full_x = mod(ins["feats"], ins["coors"], ins.get("edges"), mask=ins.get("mask"))
                f1, x1 = parallel.row_sharded_layer_peer(comm, mod, ins["feats"][:, r0:r1], ins["coors"][:, r0:r1], n, **kw)
                f2, x2 = parallel.row_sharded_layer_peer(comm, mod, f1, x1, n, **kw)          # second layer call on the outputs
                g_f, g_x = mod(full_f, full_x.float(), ins.get("edges"), mask=ins.get("mask"))
            torch.cuda.synchronize(dev)
            assert comm.status() == 0
            assert torch.equal(f1, full_f[:, r0:r1]) and torch.equal(x1.float(), full_x[:, r0:r1].float()), name
            assert torch.equal(f2, g_f[:, r0:r1]) and torch.equal(x2.float(), g_x[:, r0:r1].float()), name
            comm.close()
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
