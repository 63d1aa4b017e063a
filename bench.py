#!/usr/bin/env python
"""Headline benchmark: settles/sec on the serving batch (BASELINE.json configs[2]:
B independent lattices N=1200 D=384 k=8 settled concurrently on one B200).

One STEP = the reference's per-request sequence for every lattice of the batch
(cloud/app/main.py:916-939,1043,1061; scripts/benchmark.py:50-70):
    build (ctor) + set_query + settle(max_iters=12, tol=1e-3) + light receipt (U* solve + deltaH)

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

`value`  : lattices/s with the anchors already resident in HBM (CUDA events, max over ranks).
`e2e`    : the same through the public API with HOST buffers: pinned Y/psi -> H2D, build, settle,
           receipt, D2H of the per-lattice results, all inside the timed region.
`roofline`: the dominant kernel of the step, timed live with CUDA events on the launching stream.
`cpu_baseline` / --impl reference: the oracle's dense literal restatement of the reference
           (oracle/dense.py, kind "port": the reference is Python and cannot travel to the GPU box)
           on the host cores.
Multi-GPU: lattices are independent -> replicas only (the batch is sharded, no collective on the
data path); scaling is weak (B lattices per GPU).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_LAT, D_LAT, K_LAT = 1200, 384, 8
METRIC = "settles/sec (N=1200,D=384 batched)"
UNIT = "lattices/s"


# ----------------------------------------------------------------------------- CPU baseline
def cpu_reference_rate(n_lattices: int, workers: int) -> dict:
    """Time the dense oracle (reference restatement) on `n_lattices` lattices using `workers`
    processes (BLAS single-threaded inside each, which is the throughput-optimal CPU layout)."""
    import multiprocessing as mp

    from oracle import cpu_bench

    ctx = mp.get_context("spawn")
    old = {k: os.environ.get(k) for k in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS")}
    for k in old:
        os.environ[k] = "1"
    try:
        with ctx.Pool(workers) as pool:
            pool.map(cpu_bench.settle_one, range(workers))  # warm-up: imports + first BLAS call
            t0 = time.perf_counter()
            out = pool.map(cpu_bench.settle_one, range(n_lattices), chunksize=1)
            dt = time.perf_counter() - t0
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return {"value": n_lattices / dt, "seconds": dt, "deltaH0": out[0]}


def host_workers() -> int:
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return max(1, min(n, 64))


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workers = host_workers()
    per_step = max(workers, 16)
    for _ in range(args.warmup if args.warmup < 1 else 1):
        cpu_reference_rate(workers, workers)
    t_total, n_total = 0.0, 0
    for _ in range(args.steps):
        r = cpu_reference_rate(per_step, workers)
        t_total += r["seconds"]
        n_total += per_step
    value = n_total / t_total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * t_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"serving batch: independent lattices N={N_LAT} D={D_LAT} k={K_LAT}, "
                               "build + settle(12,1e-3) + light receipt per lattice",
                   "lattices_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port",
                         "sample": f"{per_step} lattices/step x {args.steps} steps, dense oracle "
                                   f"(oracle/dense.py) in {workers} processes, 1 BLAS thread each"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, bit in names.items():
                    if mask & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------------------- our arm
def run_ours(args) -> None:
    import numpy as np
    import torch
    import torch.distributed as dist

    from oscillink_b200 import BatchedLattices, settle_host_batch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    # synthetic Gaussian anchors; psi_b = normalise(mean(Y_b[:32]))  (scripts/benchmark.py:45-48)
    Y = torch.randn((B, N_LAT, D_LAT), generator=gen, device=dev, dtype=torch.float32)
    psi = Y[:, :32, :].mean(dim=1)
    psi = psi / (psi.norm(dim=1, keepdim=True) + 1e-12)
    Y_host = torch.empty((B, N_LAT, D_LAT), dtype=torch.float32, pin_memory=True)
    Y_host.copy_(Y)
    psi_host = torch.empty((B, D_LAT), dtype=torch.float32, pin_memory=True)
    psi_host.copy_(psi)
    res_host = torch.empty((B, 5), dtype=torch.float64, pin_memory=True)
    torch.cuda.synchronize()

    def step_resident():
        bl = BatchedLattices(Y, kneighbors=K_LAT)
        bl.set_query(psi)
        out = bl.settle(max_iters=12, tol=1e-3, receipt=True)
        return bl, out

    def step_e2e():
        # the public host-buffer call: chunked H2D on a copy stream overlapped with build + settle,
        # results D2H into pinned memory; synchronises before returning
        settle_host_batch(Y_host, psi_host, kneighbors=K_LAT, chunk=args.chunk, max_iters=12, tol=1e-3,
                          receipt=True, out_host=res_host, device=dev)
        return None, None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        phases = []
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            bl, _ = fn()
            if bl is not None:
                phases.append(bl.events)
            del bl
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, phases

    for _ in range(max(args.warmup, 3)):
        bl, out = step_resident()
        del bl
    torch.cuda.synchronize()
    check = {"iters_mean": float(out["iters"].mean().item()),
             "ustar_iters_mean": float(out["ustar_iters"].mean().item()),
             "deltaH_mean": float(out["deltaH"].mean().item())}
    engine = None
    with ClockSampler(local) as clk:
        ms, phases = timed(step_resident, args.steps)
    clocks = clk.summary()
    # per-kernel device time, averaged over the timed steps
    torch.cuda.synchronize()
    kern = {}
    for ev in phases:
        for name, (a, b) in ev.items():
            kern.setdefault(name, []).append(a.elapsed_time(b))
    kern_ms = {k: sum(v) / len(v) for k, v in kern.items()}
    bl, _ = step_resident()
    engine = getattr(bl, "engine_used", "simt")
    nnz_mean = float(bl.nnz.double().mean().item())
    check["knn_candidate_width"] = int(getattr(bl, "kc", 0))
    check["rows_recomputed_exhaustively_per_step"] = int(bl.n_exhaustive.item()) if hasattr(bl, "n_exhaustive") else 0
    del bl
    step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)

    value = world * B * args.steps / (ms / 1000.0)
    e2e_value = world * B * args.steps / (ms_e2e / 1000.0)

    # ---- roofline of the dominant kernel
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
        peak_src = "measured"
    except Exception:
        peaks = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
        peak_src = "fallback"
    dom = max(kern_ms, key=kern_ms.get)
    if dom == "knn_candidates":
        flops = 2.0 * N_LAT * N_LAT * D_LAT * B  # 2 N^2 D per lattice
        achieved = flops / (kern_ms[dom] / 1000.0) / 1e12
        # fp16 engine: the measured bf16 rate; TF32 engines: half of it
        half = engine != "tch"
        peak = (0.5 if half else 1.0) * float(peaks["bf16_tflops_sustained"])
        roof = {"kernel": "knn_candidates(" + engine + ")", "bound": "tensor", "achieved": achieved,
                "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                "peak_source": f"{peak_src}: " + ("0.5 x bf16_tflops_sustained (TF32 = half bf16 rate)" if half
                                                  else "bf16_tflops_sustained (fp16 operands)"),
                "algorithmic": "2*N^2*D flops per lattice (3xTF32 issues 3x that on the tensor pipe)"}
    else:
        V = N_LAT * D_LAT * 4.0
        if dom == "batched_settle":
            byts = (3.0 * V + 12.0 * nnz_mean + 2.0 * V) * B  # SURVEY 8(d) + U* solve re-reads Y, writes nothing else
        elif dom == "knn_rescore":
            byts = (K_LAT + 4 + 1) * V * B
        elif dom == "normalize":
            byts = 2.0 * V * B
        else:
            byts = 3.0 * N_LAT * K_LAT * 8.0 * B
        achieved = byts / (kern_ms[dom] / 1000.0) / 1e9
        peak = float(peaks["hbm_gbs"])
        roof = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src}
    # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture (per lattice,
    # scaled to this launch); profiles/ncu_traffic.json names the capture it came from
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tr = json.load(f).get(dom)
        if tr:
            roof["traffic"] = tr["dram_bytes_per_lattice"] * B
            roof["traffic_source"] = tr["source"]
    except Exception:
        pass
    if dom == "batched_settle":
        # What binds this kernel is the shared-memory data pipe, not HBM: p (the gathered vector) and the
        # graph image live in shared memory.  Algorithmic shared-memory bytes per lattice:
        # every gather pass moves Np*kp*(16 B of p + 6 B of graph) per 4-column slab, every CG iteration
        # re-reads and re-writes the thread's own rows of p (Np * 32 B).
        it_s, it_u = check["iters_mean"], check["ustar_iters_mean"]
        Np, kp, G = 1280, (K_LAT + 3) // 4 * 4, D_LAT // 4
        passes = it_s + it_u + 1.0          # one gather pass over Y serves both initial residuals
        smem_bytes = G * (passes * Np * kp * 22.0 + (it_s + it_u) * Np * 32.0) * B
        sm_clk = float(clocks.get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0) * 1e6
        smem_peak = 148 * 128.0 * sm_clk / 1e9   # 128 B per clock per SM
        onchip = {"bound": "smem", "achieved": smem_bytes / (kern_ms[dom] / 1000.0) / 1e9, "peak": smem_peak,
                  "unit": "GB/s", "algorithmic": "G*((it_s+it_u+1)*Np*kp*22 + (it_s+it_u)*Np*32) bytes per lattice, "
                  "conflict-free; peak = 148 SMs x 128 B/clk x sampled SM clock"}
        onchip["frac"] = onchip["achieved"] / onchip["peak"]
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                tr = json.load(f).get(dom) or {}
            if "l1tex_data_pipe_pct" in tr:
                onchip["ncu_l1tex_data_pipe_pct"] = tr["l1tex_data_pipe_pct"]
                onchip["ncu_source"] = tr.get("l1tex_source")
        except Exception:
            pass
        roof["onchip"] = onchip
        roof["note"] = ("everything but Y in / U out stays on chip, so HBM is not the binding resource; the "
                        "binding unit is the L1TEX/shared-memory data pipe (see `onchip`; the ncu capture counts "
                        "bank-conflict and reduction wavefronts on top of the algorithmic bytes)")
    roof["kernel_ms_per_step"] = kern_ms
    roof["share_of_step"] = {k: v / (ms / args.steps) for k, v in kern_ms.items()}

    line = None
    if rank == 0:
        workers = host_workers()
        sample = max(2 * workers, 32)
        cpu = cpu_reference_rate(sample, workers)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"serving batch: {B} independent lattices N={N_LAT} D={D_LAT} "
                                   f"k={K_LAT} per GPU; step = build + settle(12,1e-3) + light receipt",
                       "lattices_per_step_per_gpu": B, "parallelism": f"replicas x{world}",
                       "l2": "inputs (7.5 GB/step at B=4096) larger than L2", "knn_engine": engine,
                       "e2e_call": f"settle_host_batch(chunk={args.chunk}): pinned host Y/psi -> results in pinned host"},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(Y_host.numel() * 4 + psi_host.numel() * 4),
                    "d2h_bytes_per_step": int(res_host.numel() * 8)},
            # normalize, knn_tc, rescore, assemble(3 kernels), pack, settle, resolve, settle(fix), finalize
            # normalize, knn_tc2, rescore, exact_rows, assemble x3, pack, settle, resolve, settle(fix), finalize
            "gpu_launches": 12 * args.steps,
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": {"value": cpu["value"], "unit": UNIT, "cores": workers, "kind": "port",
                             "sample": f"{sample} lattices, dense oracle (oracle/dense.py), "
                                       f"{workers} processes x 1 BLAS thread"},
            "check": check,
        }
    if not args.no_large and world == 1:
        # the second half of BASELINE.json's metric (ms/settle of ONE big lattice) at the size that
        # builds in seconds; N=10M and the multi-GPU runs (--workload large) are under profiles/
        del Y, Y_host
        torch.cuda.empty_cache()
        big = measure_large(args.N, args.D, args.k, steps=3, warmup=3, partition="rows", world=1, rank=0,
                            local=local)
        line["large_lattice"] = {k: big[k] for k in ("metric", "value", "unit", "n_gpus", "config", "roofline",
                                                     "build_ms", "receipt_light_ms", "check")}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ----------------------------------------------------------------------------- large single lattice
def measure_large(N, D, k, *, steps, warmup, chain_len=0, partition="rows", world=1, rank=0, local=0, p2p=None):
    """BASELINE.json configs[3]/[4]: ONE big lattice (N up to 10M), kNN build + PCG settle, the rows
    (or column slabs) partitioned over the GPUs.  A step = one settle(12, 1e-3) from U = Y on the built
    lattice (the metric's `ms/settle`); build, U* + deltaH and the per-kernel rooflines ride along.
    Returns the JSON-able result dict on rank 0 (None elsewhere)."""
    import types

    import torch
    import torch.distributed as dist

    from oscillink_b200 import _cabi
    from oscillink_b200.sharded_api import ShardedLattice, _NativeKernels, gather_rows, shard_bounds

    dev = torch.device("cuda", local)
    args = types.SimpleNamespace(steps=steps, warmup=warmup, chain_len=chain_len, partition=partition)
    row0, n_local, _ = shard_bounds(N, world, rank)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    Y_local = torch.randn((n_local, D), generator=gen, device=dev, dtype=torch.float32)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lat = ShardedLattice(Y_local, N, kneighbors=k, mode="rows" if args.partition == "both" else args.partition,
                         p2p=p2p)
    e1.record()
    barrier()
    build_ms = max_over_ranks(e0.elapsed_time(e1))
    head = Y_local[:32].mean(dim=0)
    if world > 1:
        dist.broadcast(head, src=0)
    psi = (head / (head.norm() + 1e-12)).cpu().numpy()
    psi_host = torch.from_numpy(psi.copy()).pin_memory()
    lat.set_query(psi)
    if args.chain_len >= 2:
        lat.add_chain(list(range(args.chain_len)), lamP=0.2)
    def measure_mode(part):
        Y0 = lat._Y

        def one_settle():
            lat._U = Y0.clone()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            lat.set_query(psi_host.numpy())          # host psi -> device (the request's only host input)
            st = lat.settle(max_iters=12, tol=1e-3)  # returns host {iters,res}: D2H inside
            b.record()
            barrier()
            return a.elapsed_time(b), st

        for _ in range(max(args.warmup, 1)):
            one_settle()
        with ClockSampler(local) as clk:
            tot, st = 0.0, None
            for _ in range(args.steps):
                ms, st = one_settle()
                tot += ms
        clocks = clk.summary()
        ms_settle = max_over_ranks(tot / args.steps)

        barrier()
        e0.record()
        rec = lat.receipt()
        e1.record()
        barrier()
        receipt_ms = max_over_ranks(e0.elapsed_time(e1))

        # ---- per-kernel timings on the live state (one rank-local launch each, CUDA events)
        Dl = D if part == "rows" else lat.Dl
        n_loc = n_local if part == "rows" else N
        lat._Ustar = None
        fused = part == "rows" and lat._want_p2p and lat._peers is not None
        if fused:
            X = lat._peers.X[:n_loc]
            X.copy_(lat._U)
        else:
            X = lat._U  # clobbered below: the timed settles and the receipt are done
        torch.cuda.empty_cache()
        kf = _NativeKernels(lat, _cabi.MODE_SETTLE, 1.0, True, X, torch.zeros_like(X))
        ones = torch.ones(Dl, dtype=torch.float32, device=dev)

        def t_of(fn, reps=3):
            fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps

        gather = lambda v: gather_rows(v, N, lat.group).contiguous()  # noqa: E731
        if fused:
            full = lambda v: (kf.peer_sync(), None)[1]  # noqa: E731
        else:
            full = gather if part == "rows" else (lambda v: v)
        x_all = full(X)
        kf.residual0(x_all)
        del x_all
        p_all = full(kf.P)
        nnz = float(lat.nnz.item())
        nnz_loc = nnz * n_loc / max(N, 1) if part == "rows" else nnz
        V = n_loc * Dl * 4.0
        kms = {
            "pcg_spmm": t_of(lambda: kf.spmm(p_all)),
            "pcg_update": t_of(lambda: kf.update(ones, ones)),
            "pcg_pupdate": t_of(lambda: kf.pupdate(ones, ones)),
        }
        if world > 1 and part == "rows":
            # what the un-fused schedule pays in front of every SpMM (for reference when fused)
            kms["halo_allgather_p"] = t_of(lambda: gather(kf.P))
            kms["halo"] = "fused: peers' rows read over NVLink inside pcg_spmm" if fused else "NCCL all-gather"
        alg = {  # algorithmic bytes per launch
            "pcg_spmm": (nnz_loc / max(n_loc, 1) + 2.0) * V + 8.0 * nnz_loc,
            "pcg_update": 6.0 * V,
            "pcg_pupdate": 3.0 * V,
        }
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
            peak_src = "measured"
        except Exception:
            peaks = {"hbm_gbs": 6650.0}
            peak_src = "fallback"
        peak = float(peaks["hbm_gbs"])
        gbs = {n: alg[n] / (kms[n] / 1000.0) / 1e9 for n in alg}
        if fused:  # the SpMM is now bounded by NVLink for the remote share of its gathers
            roof_note = ("rows partition with the halo fused: (world-1)/world of the gathered rows arrive over "
                         "NVLink (~775 GB/s per GPU measured for LDG.128 peer reads), not HBM")
        else:
            roof_note = None
        iters = int(st["iters"])
        iter_bytes = alg["pcg_spmm"] + alg["pcg_update"] + alg["pcg_pupdate"]
        solve_bytes = (iters + 1) * alg["pcg_spmm"] + iters * alg["pcg_update"] + (iters - 1) * alg["pcg_pupdate"] + 4.0 * V
        roof = {"kernel": "pcg_spmm_kernel", "bound": "hbm", "achieved": gbs["pcg_spmm"], "peak": peak,
                "unit": "GB/s", "frac": gbs["pcg_spmm"] / peak, "traffic": None, "peak_source": peak_src,
                "kernel_ms": kms, "kernel_gbs": gbs,
                "whole_settle": {"algorithmic_bytes": solve_bytes, "achieved_gbs": solve_bytes / (ms_settle / 1e3) / 1e9,
                                 "frac": solve_bytes / (ms_settle / 1e3) / 1e9 / peak},
                "bytes_per_iteration": iter_bytes}
        return dict(ms_settle=ms_settle, st=st, rec=rec, receipt_ms=receipt_ms, roof=roof, clocks=clocks, V=V,
                    iters=iters, nnz=nnz)

    first = "rows" if args.partition == "both" else args.partition
    m = measure_mode(first)
    ms_settle, st, rec, receipt_ms, roof, clocks, V, iters, nnz = (m[k] for k in (
        "ms_settle", "st", "rec", "receipt_ms", "roof", "clocks", "V", "iters", "nnz"))
    other = None
    halo_first = "fused P2P" if lat._want_p2p else "NCCL all-gather"
    other_halo = None
    if args.partition == "both":
        def brief(mm, par):
            return {"parallelism": par, "value": mm["ms_settle"], "unit": "ms", "roofline": mm["roof"],
                    "receipt_light_ms": mm["receipt_ms"],
                    "check": {"iters": int(mm["st"]["iters"]), "res": float(mm["st"]["res"]),
                              "deltaH": mm["rec"]["deltaH_total"]}}

        if world > 1:  # the other halo strategy of the rows partition, same graph
            lat.set_halo(not lat._want_p2p)
            mh = measure_mode("rows")
            other_halo = brief(mh, f"rows x{world} ({'fused P2P' if lat._want_p2p else 'NCCL all-gather'} halo)")
        lat.repartition("columns")   # same graph, state transposed by one all-to-all
        other = brief(measure_mode("columns"), f"columns x{world}")
    line = None
    if rank == 0:
        line = {
            "metric": f"ms/settle at N={N},D={D}", "value": ms_settle, "unit": "ms", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": ms_settle,
            "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"one lattice N={N} D={D} k={k} chain_len={args.chain_len}: "
                                   "step = settle(12,1e-3) from U=Y on the built mutual-kNN graph",
                       "parallelism": f"{first} x{world}" + (f" ({halo_first} halo)"
                                                              if (first == "rows" and world > 1) else ""),
                       "l2": f"vectors ({V / 1e9:.2f} GB each per GPU) larger than L2"},
            "e2e": {"value": ms_settle, "unit": "ms", "h2d_bytes_per_step": int(D * 4),
                    "d2h_bytes_per_step": 8,
                    "note": "set_query(host psi) + settle() -> host {iters,res}; the lattice state is device-resident by API"},
            "gpu_launches": int(2 + 6 * iters) * args.steps,
            "clocks": clocks, "roofline": roof,
            "build_ms": build_ms, "receipt_light_ms": receipt_ms,
            "check": {"iters": iters, "res": float(st["res"]), "ustar_iters": rec["meta"]["ustar_iters"],
                      "ustar_res": rec["meta"]["ustar_res"], "deltaH": rec["deltaH_total"],
                      "avg_degree": rec["meta"]["avg_degree"], "nnz": nnz,
                      "rows_recomputed_exhaustively": int(getattr(lat, "n_exhaustive", torch.zeros(1)).item())},
        }
    if line is not None and other is not None:
        line["columns_partition"] = other
    if line is not None and other_halo is not None:
        line["rows_other_halo"] = other_halo
    lat.close()
    del lat
    torch.cuda.empty_cache()
    return line


def run_large(args) -> None:
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    line = measure_large(args.N, args.D, args.k, steps=args.steps, warmup=args.warmup, chain_len=args.chain_len,
                         partition=args.partition, world=world, rank=rank, local=local,
                         p2p=False if args.no_p2p else None)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=4096, help="lattices per step per GPU")
    ap.add_argument("--chunk", type=int, default=128, help="lattices per H2D/compute pipeline stage (e2e)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="serving", choices=["serving", "large"],
                    help="serving: the headline batch of N=1200 lattices; large: one big lattice (ms/settle)")
    ap.add_argument("--N", type=int, default=1_000_000)
    ap.add_argument("--D", type=int, default=768)
    ap.add_argument("--k", type=int, default=16)
    ap.add_argument("--chain-len", type=int, default=0)
    ap.add_argument("--partition", default="rows", choices=["rows", "columns", "both"])
    ap.add_argument("--no-p2p", action="store_true",
                    help="large workload, rows partition: NCCL all-gather halo instead of the fused P2P halo")
    ap.add_argument("--no-large", action="store_true",
                    help="serving workload only: skip the one-big-lattice block of the JSON line")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "large":
        run_large(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
