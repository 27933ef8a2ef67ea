#!/usr/bin/env python
"""Benchmark of the EGNN forward hot path (BASELINE.json metric: node-pairs/sec, dim=512 N=1024).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|eager] [--dtype bf16|fp32]

One "step" = one EGNN(dim=512) layer forward over one batch (B=4 graphs x N=1024 nodes, dense
all-pairs = 4,194,304 node pairs) of synthetic N(0,1) inputs with the reference's default init
(BASELINE.json configs[1], SURVEY.md section 8(d) "c2").  Multi-GPU: one process per GPU under torchrun,
every rank runs its own batch (graphs are independent units: weak scaling, no data-path collective).

Prints ONE JSON line (rank 0).  `value` = whole-job pairs/s with inputs resident in HBM, timed
per step with CUDA events on the launch stream (L2 flushed between steps, flush not timed),
max over ranks.  `e2e` = the same through the public module API with pinned HOST tensors
(H2D + D2H inside the timed region).  `roofline` describes the fused edge kernel, timed live by
the library's own CUDA-event stage brackets (egnn_profile_*).  `cpu_baseline` = the oracle on the
host cores on a bounded sample.  `--impl reference` times the reference's algorithm on the CPU
(the oracle port; the Python reference itself cannot travel to the GPU box).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
for p in (REPO, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

torch.set_grad_enabled(False)      # forward benchmark: no autograd state is kept

WORKLOADS = {
    # BASELINE.json configs[1]
    "c2": dict(kind="layer", cfg=dict(dim=512), B=4, N=1024, C=3, label="EGNN(dim=512) dense all-pairs B=4 N=1024"),
    # BASELINE.json configs[0] (latency-bound; for reference only)
    "c1": dict(kind="layer", cfg=dict(dim=512), B=1, N=16, C=3, label="EGNN(dim=512) dense all-pairs B=1 N=16"),
}


def peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"], sm_max_mhz=p.get("sm_max_mhz", 1965.0),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, sm_max_mhz=1965.0, source="fallback (B200_PROFILING.md)")


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-i", str(index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except Exception:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    power_w_max=max(pw) if pw else None, samples=len(sm), reasons=sorted(reasons))


# ----------------------------------------------------------------------------- workload
def build_workload(name, dtype, device, seed=0):
    """Module with the reference's default init (weights N(0,1e-3), PyTorch-default biases) and
    synthetic N(0,1) inputs (SURVEY.md section 8(d) 'Synthetic inputs')."""
    from egnn_pytorch_b200 import EGNN
    w = WORKLOADS[name]
    torch.manual_seed(seed)
    mod = EGNN(**w["cfg"]).to(dtype).to(device).eval()
    g = torch.Generator().manual_seed(seed + 1)
    feats = torch.randn(w["B"], w["N"], w["cfg"]["dim"], generator=g)
    coors = torch.randn(w["B"], w["N"], w["C"], generator=g)
    return mod, feats, coors


def work_per_pair(cfg):
    """Per-pair work of SURVEY.md section 8(d): reference-formulation FLOPs, and the split formulation's
    activation count and implemented tensor FLOPs."""
    d, e, F, m = cfg["dim"], cfg.get("edge_dim", 0), cfg.get("fourier_features", 0), cfg.get("m_dim", 16)
    E = 2 * d + 2 * F + 1 + e
    H = 2 * E
    return dict(E=E, H=H, m=m, f_ref=2 * E * H + 2 * H * m + 2 * m * 4 * m + 2 * 4 * m, act=H + m + 4 * m,
                tensor_flops=2 * ((H + 63) // 64 * 64) * 16)


def compulsory_bytes(w, es):
    """2*B*N*d*s (feats in+out) + 2*B*N*C*4 (coors) + weights*s."""
    d, B, N, C = w["cfg"]["dim"], w["B"], w["N"], w["C"]
    E = 2 * d + 1
    H = 2 * E
    weights = H * E + H + 16 * H + 16 + (2 * d) * (d + 16) + 2 * d + d * 2 * d + d + 64 * 16 + 64 + 64 + 1
    return 2 * B * N * d * es + 2 * B * N * C * 4 + weights * es


def equivariance_error(mod, feats, coors, dtype, device):
    """tests/test_equivariance.py:8-34 as a number: max|feats(Rx+t) - feats(x)|, max|coors(Rx+t) - (coors(x)R+t)|."""
    g = torch.Generator().manual_seed(7)
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g, dtype=torch.float64))
    t = torch.randn(1, 1, 3, generator=g, dtype=torch.float64)
    f = feats[:1].to(device, dtype)
    x = coors[:1].double()
    xr = (x @ q + t)
    f2, c2 = mod(f, x.float().to(device))
    f1, c1 = mod(f, xr.float().to(device))
    ef = float((f1.double() - f2.double()).abs().max())
    ec = float((c1.double().cpu() - (c2.double().cpu() @ q + t)).abs().max())
    return dict(feats=ef, coors=ec)


# ----------------------------------------------------------------------------- CPU baseline (oracle port)
_CPU_JOB = {}


def _cpu_job(span):
    from oracle import egnn_oracle as O
    j = _CPU_JOB
    O.egnn_layer_forward(j["params"], j["cfg"], j["feats"], j["coors"], dtype=np.float32, row_chunk=j["block"], rows=span)
    return span[1] - span[0]


def cpu_baseline_sample(name, target_seconds=12.0, seed=0):
    """Time the oracle (numpy float32) on a bounded sample of the workload: `rows` i-rows of ONE
    graph against all N neighbours, the rows split over one forked worker PROCESS per host core
    (single-threaded BLAS in each), so both the Linear-1 GEMM and the elementwise SiLU use every
    core.  Must run in a process that has not initialised CUDA (bench.py's GPU arm calls it through
    `python bench.py --impl cpu-sample`)."""
    import multiprocessing as mp
    import cases
    try:
        from threadpoolctl import threadpool_limits
    except Exception:                       # pragma: no cover
        import contextlib
        threadpool_limits = lambda **kw: contextlib.nullcontext()
    w = WORKLOADS[name]
    cores = os.cpu_count() or 1
    spec = dict(kind="layer", cfg=w["cfg"], B=1, N=w["N"], C=w["C"], seed=seed)
    case = cases.build_case(spec)
    ins = case["inputs"]
    block = 8
    _CPU_JOB.update(params={k: np.asarray(v, np.float32) for k, v in case["params"].items()}, cfg=case["cfg"],
                    feats=ins["feats"].astype(np.float32), coors=ins["coors"].astype(np.float32), block=block)
    with threadpool_limits(limits=1):
        with mp.get_context("fork").Pool(cores) as pool:
            def run(rows):
                spans = [(a, min(a + block, rows)) for a in range(0, rows, block)]
                t0 = time.perf_counter()
                pool.map(_cpu_job, spans, chunksize=1)
                return time.perf_counter() - t0
            probe = min(block * cores, w["N"])
            run(probe)                               # warm-up (page faults, BLAS init in the workers)
            dt = run(probe)
            reps = int(max(1, min(64, target_seconds / max(dt, 1e-3))))
            rows_list = [min(w["N"], probe)] * reps
            if probe < w["N"]:
                rows = int(min(w["N"], max(probe, probe * target_seconds / max(dt, 1e-3))))
                rows_list = [max(block, rows // block * block)]
            t = sum(run(r) for r in rows_list)
    pairs = sum(rows_list) * w["N"]
    return dict(value=pairs / t, unit="pairs/s", cores=cores, kind="port",
                sample=f"{len(rows_list)} x {rows_list[0]} of {w['N']} i-rows of one graph x all {w['N']} neighbours "
                       f"({pairs} pairs, {t:.1f} s), numpy float32 oracle, {cores} worker processes"), pairs, t


def cpu_baseline_subprocess(name):
    """Run the sample in a fresh interpreter (no CUDA context, fork-safe)."""
    res = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "cpu-sample", "--workload", name],
                         capture_output=True, text=True, timeout=900)
    for line in reversed(res.stdout.strip().splitlines()):
        if line.startswith("{"):
            return json.loads(line)
    return dict(value=None, unit="pairs/s", cores=os.cpu_count(), kind="port", sample="failed: " + res.stderr[-300:])


# ----------------------------------------------------------------------------- arms
def dist_setup(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    return world, rank, local


def barrier_max(world, value, device):
    if world == 1:
        return value
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def arm_ours(args):
    from egnn_pytorch_b200 import _native as nat
    world, rank, local = dist_setup(args)
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    lib = nat.load()
    w = WORKLOADS[args.workload]
    dtype = {"bf16": torch.bfloat16, "fp32": torch.float32}[args.dtype]
    mod, feats, coors = build_workload(args.workload, dtype, dev, seed=rank)
    f_dev, x_dev = feats.to(dev, dtype), coors.to(dev)
    pairs_rank = w["B"] * w["N"] * w["N"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)        # > 126 MB L2

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- warm-up (also builds the packed-parameter cache)
    for _ in range(max(args.warmup, 3)):
        mod(f_dev, x_dev)
    path = mod.last_path
    sync_all()

    # ---- device-resident timing: K steps, CUDA events per step on the launch stream
    lib.egnn_profile_read(None, None, None, 1)
    lib.egnn_profile_enable(1)
    sampler = ClockSampler(local) if rank == 0 else None
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sync_all()
    for a, b in evs:
        flush.zero_()                      # L2 flush between timed iterations (not timed)
        a.record()
        mod(f_dev, x_dev)
        b.record()
    sync_all()
    clocks = sampler.stop() if sampler else None
    ms_total = sum(a.elapsed_time(b) for a, b in evs)
    ms = (C.c_float * 4)(); spans = (C.c_int32 * 4)(); launches = C.c_int64()
    lib.egnn_profile_read(ms, spans, C.byref(launches), 1)
    lib.egnn_profile_enable(0)
    ms_total = barrier_max(world, ms_total, dev)
    ms_per_step = ms_total / args.steps
    value = pairs_rank * world / (ms_per_step * 1e-3)

    # ---- end-to-end: pinned host tensors through the public module API, H2D + D2H timed
    hf = feats.to(dtype).pin_memory()
    hx = coors.pin_memory()
    for _ in range(2):
        mod(hf, hx)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        of, ox = mod(hf, hx)               # returns CPU tensors (synchronous D2H)
    torch.cuda.synchronize(dev)
    e2e_s = barrier_max(world, time.perf_counter() - t0, dev)
    e2e_value = pairs_rank * world / (e2e_s / args.steps)
    h2d = hf.numel() * hf.element_size() + hx.numel() * hx.element_size()
    d2h = of.numel() * of.element_size() + ox.numel() * ox.element_size()

    if rank != 0:
        return
    # ---- roofline of the fused edge kernel (stage 2), timed live by the library's event brackets
    pk = peaks()
    wp = work_per_pair(w["cfg"])
    pair_ms = ms[2] / max(1, spans[2])
    pair_s = pair_ms * 1e-3
    es = 2 if dtype == torch.bfloat16 else 4
    sfu_peak = 148 * 16 * pk["sm_max_mhz"] * 1e6 / 1e9                 # G activations / s  (16 MUFU/clk/SM)
    act_rate = pairs_rank * wp["act"] / pair_s / 1e9
    traffic = None
    tpath = os.path.join(REPO, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(f"{args.workload}:{path}")
    tensor_tf = pairs_rank * wp["tensor_flops"] / pair_s / 1e12 if path == "bf16-tcgen05" else 0.0
    roofline = dict(
        bound="sfu", kernel="fused edge kernel (stage 2 of egnn_layer_forward)", path=path,
        achieved=act_rate, peak=sfu_peak, unit="Gsilu/s", frac=act_rate / sfu_peak, traffic=traffic,
        note="split formulation: the binding unit is the MUFU/SFU pipe (SURVEY.md section 8(d)); "
             f"peak = 148 SM x 16 MUFU/clk x {pk['sm_max_mhz']:.0f} MHz",
        launch_ms=pair_ms, stage_ms_per_step={k: ms[i] / args.steps for i, k in
                                              enumerate(["select", "node_pre", "edge", "node_post"])},
        tensor=dict(achieved=tensor_tf, peak=pk["bf16_tflops"], unit="TFLOP/s", frac=tensor_tf / pk["bf16_tflops"],
                    note="implemented bf16 MMA FLOPs only (2*Hpad*16 per pair)"),
        hbm=dict(achieved=compulsory_bytes(w, es) / pair_s / 1e9, peak=pk["hbm_gbs"], unit="GB/s",
                 frac=compulsory_bytes(w, es) / pair_s / 1e9 / pk["hbm_gbs"], algorithmic_bytes=compulsory_bytes(w, es)),
        effective_reference_tflops=pairs_rank * wp["f_ref"] / pair_s / 1e12,
        peaks=pk["source"])

    out = dict(
        metric="EGNN fwd node-pairs/sec (dim=512 N=1024)", value=value, unit="pairs/s", n_gpus=world,
        steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=ms_per_step, higher_is_better=True, scaling="weak",
        vs_baseline=None, dtype="bf16" if path == "bf16-tcgen05" else "f32", data="synthetic",
        config=dict(workload=w["label"], per_gpu_batch=w["B"], nodes=w["N"], pairs_per_step=pairs_rank * world,
                    kernel_path=path, init="reference default init", l2="flushed between timed steps (256 MiB memset)",
                    parallelism=f"batch-sharded x{world} (independent graphs, no collective)"),
        e2e=dict(value=e2e_value, unit="pairs/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                 ms_per_step=e2e_s / args.steps * 1e3),
        gpu_launches=int(launches.value), clocks=clocks, roofline=roofline,
        equivariance_err=equivariance_error(mod, feats, coors, dtype, dev))
    if world == 1:
        out["train_step"] = train_step_probe(args.workload, dev, pairs_rank)
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline_subprocess(args.workload)
    print(json.dumps(out))


def train_step_probe(workload, dev, pairs):
    """Informational (not part of the metric): one forward + backward of the same workload through the autograd
    bridge (fp32 recompute-in-backward kernels, egnn_layer_backward), median of 3 after 2 warm-ups."""
    try:
        mod, feats, coors = build_workload(workload, torch.float32, dev, seed=0)
        mod.requires_grad_(True)
        f = feats.to(dev, torch.float32).requires_grad_(True)
        x = coors.to(dev).requires_grad_(True)
        ts = []
        with torch.enable_grad():
            for it in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fo, xo = mod(f, x)
                (fo.sum() + xo.sum()).backward()
                b.record()
                torch.cuda.synchronize(dev)
                if it >= 2:
                    ts.append(a.elapsed_time(b))
                mod.zero_grad(set_to_none=True)
                f.grad = x.grad = None
        ms = sorted(ts)[1]
        return dict(ms_per_step=ms, pairs_per_s=pairs / ms * 1e3, dtype="f32", what="forward + backward, device-resident")
    except Exception as e:  # noqa: BLE001  (a probe must never break the benchmark line)
        return dict(error=f"{type(e).__name__}: {e}"[:200])


def arm_reference(args):
    """The reference's algorithm on the host cores (oracle port), rank 0 only."""
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    if rank != 0:          # under torchrun the other ranks exit 0 without work (no process group needed)
        return
    w = WORKLOADS[args.workload]
    res, times = None, []
    per = max(2.0, min(12.0, 60.0 / max(1, args.steps + args.warmup)))
    for i in range(args.warmup + args.steps):
        res, pairs, dt = cpu_baseline_sample(args.workload, target_seconds=per)
        if i >= args.warmup:
            times.append((pairs, dt))
    pairs = sum(p for p, _ in times); dt = sum(t for _, t in times)
    value = pairs / dt
    res["value"] = value
    print(json.dumps(dict(
        impl="reference", metric="EGNN fwd node-pairs/sec (dim=512 N=1024)", value=value, unit="pairs/s",
        n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=dt / len(times) * 1e3, higher_is_better=True,
        scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
        config=dict(workload=w["label"], note="bounded sample per step, CPU"), cpu_baseline=res,
        e2e=dict(value=value, unit="pairs/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))))


def arm_eager(args):
    """Optional: the UNMODIFIED reference on the GPU (PyTorch eager), if it was installed into
    baseline/_ref (git-ignored).  This is the '>= 10x the reference's own GPU eager path' comparison
    of BASELINE.json's north_star; not part of the driver's contract."""
    ref = os.path.join(REPO, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "egnn_pytorch")):
        print(json.dumps(dict(impl="reference-gpu-eager", unavailable="baseline/_ref not installed")))
        return
    sys.path.insert(0, ref)
    from egnn_pytorch import EGNN as RefEGNN
    w = WORKLOADS[args.workload]
    dev = torch.device("cuda", 0)
    dtype = {"bf16": torch.bfloat16, "fp32": torch.float32}[args.dtype]
    torch.manual_seed(0)
    mod = RefEGNN(**w["cfg"]).to(dtype).to(dev).eval()
    g = torch.Generator().manual_seed(1)
    feats = torch.randn(w["B"], w["N"], w["cfg"]["dim"], generator=g).to(dev, dtype)
    coors = torch.randn(w["B"], w["N"], w["C"], generator=g).to(dev, dtype)
    with torch.no_grad():
        for _ in range(max(2, args.warmup)):
            mod(feats, coors)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            mod(feats, coors)
        b.record()
        torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.steps
    pairs = w["B"] * w["N"] * w["N"]
    print(json.dumps(dict(impl="reference-gpu-eager", metric="EGNN fwd node-pairs/sec (dim=512 N=1024)",
                          value=pairs / (ms * 1e-3), unit="pairs/s", ms_per_step=ms, dtype=args.dtype, steps=args.steps,
                          peak_mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30,
                          config=dict(workload=w["label"]))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "eager", "cpu-sample"])
    ap.add_argument("--workload", default="c2", choices=list(WORKLOADS))
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "cpu-sample":
        print(json.dumps(cpu_baseline_sample(args.workload)[0]))
        return
    {"ours": arm_ours, "reference": arm_reference, "eager": arm_eager}[args.impl](args)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
