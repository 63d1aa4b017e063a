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
Multi-GPU: the serving lattices are independent -> replicas only (the batch is sharded, no collective on
the data path; weak scaling, B lattices per GPU).  The second half of BASELINE.json's metric -- ms/settle of
ONE lattice of N = 10M, D = 384, k = 10 with a chain prior -- is measured in the same run on the same ranks
(`large_lattice`): the lattice is sharded over the N GPUs (rows partition with an NVLink halo exchange +
NCCL all-reduces of the column dots, and the halo-free column-slab partition), strong scaling, with a
sampled-row parity check against an exhaustive fp64 scan (`parity_sample`).  `large_lattice_1M` is the
N = 1M, D = 768, k = 16 lattice of configs[3]; `single_lattice` times configs[0]/[1] through the
`OscillinkLattice` class with NumPy in and out (the protocol of scripts/benchmark.py:45-70).
"""
from __future__ import annotations

import argparse
import ctypes
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


# ----------------------------------------------------------------------------- helpers
def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def ncu_traffic(kernel: str):
    """DRAM bytes per lattice of `kernel` from the committed `ncu --set full` capture (profiles/)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


def dist_env():
    return (int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")),
            int(os.environ.get("LOCAL_RANK", "0")))


# ----------------------------------------------------------------------------- serving batch (headline)
def measure_serving(args, world, rank, local) -> dict | None:
    import torch
    import torch.distributed as dist

    from oscillink_b200 import BatchedLattices, settle_host_batch

    dev = torch.device("cuda", local)
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

    def step_e2e(U_host=None):
        # the public host-buffer call: chunked H2D on a copy stream overlapped with build + settle,
        # results D2H into pinned memory; synchronises before returning
        settle_host_batch(Y_host, psi_host, kneighbors=K_LAT, chunk=args.chunk, max_iters=12, tol=1e-3,
                          receipt=True, out_host=res_host, U_host=U_host, device=dev)
        return None, None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

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
        return max_over_ranks(e0.elapsed_time(e1)), phases

    for _ in range(max(args.warmup, 3)):
        bl, out = step_resident()
        del bl
    torch.cuda.synchronize()
    check = {"iters_mean": float(out["iters"].mean().item()),
             "ustar_iters_mean": float(out["ustar_iters"].mean().item()),
             "deltaH_mean": float(out["deltaH"].mean().item())}
    with ClockSampler(local) as clk:
        ms, phases = timed(step_resident, args.steps)
    clocks = clk.summary()
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
    # the same call that also returns the settled state U of every lattice (D2H on a third stream)
    e2e_steps = max(1, min(args.steps, 3))
    U_host = torch.empty((B, N_LAT, D_LAT), dtype=torch.float32, pin_memory=True)
    step_e2e(U_host)
    ms_e2e_u, _ = timed(lambda: step_e2e(U_host), e2e_steps)
    del U_host
    # ceiling of the end-to-end number: the H2D copy of one step's anchors alone (same pinned buffer, same
    # chunking, every rank at once -> at N > 1 this is the host's pinned-memory / PCIe-switch bandwidth)
    dst = [torch.empty((args.chunk, N_LAT, D_LAT), dtype=torch.float32, device=dev) for _ in range(2)]

    def h2d_only():
        for i in range(0, B, args.chunk):
            hi = min(B, i + args.chunk)
            dst[(i // args.chunk) & 1][: hi - i].copy_(Y_host[i:hi], non_blocking=True)
        return None, None

    h2d_only()
    ms_h2d, _ = timed(h2d_only, e2e_steps)
    del dst
    h2d_bytes = int(Y_host.numel() * 4 + psi_host.numel() * 4)

    value = world * B * args.steps / (ms / 1000.0)
    e2e_value = world * B * args.steps / (ms_e2e / 1000.0)

    # ---- roofline of the dominant kernel
    peaks, peak_src = load_peaks()
    dom = max(kern_ms, key=kern_ms.get)
    V = N_LAT * D_LAT * 4.0
    if dom == "knn_candidates":
        flops = 2.0 * N_LAT * N_LAT * D_LAT * B  # 2 N^2 D per lattice
        achieved = flops / (kern_ms[dom] / 1000.0) / 1e12
        half = engine != "tch"  # fp16 engine: the measured bf16 rate; TF32 engines: half of it
        peak = (0.5 if half else 1.0) * float(peaks["bf16_tflops_sustained"])
        roof = {"kernel": "knn_candidates(" + engine + ")", "bound": "tensor", "achieved": achieved,
                "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                "peak_source": f"{peak_src}: " + ("0.5 x bf16_tflops_sustained (TF32 = half bf16 rate)" if half
                                                  else "bf16_tflops_sustained (fp16 operands)"),
                "algorithmic": "2*N^2*D flops per lattice (3xTF32 issues 3x that on the tensor pipe)"}
    else:
        if dom == "batched_settle":
            byts = (3.0 * V + 12.0 * nnz_mean) * B  # read Y, read U, write U, graph
            alg = "(3V + 12 nnz) per lattice, SURVEY 8(d)"
        elif dom == "knn_rescore":
            byts = (K_LAT + 2.0) * V * B
            alg = "(k+2) V per lattice after pruning (DESIGN.md 4)"
        elif dom == "normalize":
            byts = 2.0 * V * B
            alg = "2V per lattice"
        else:
            byts = 3.0 * N_LAT * K_LAT * 8.0 * B
            alg = "3 N k 8 bytes per lattice"
        achieved = byts / (kern_ms[dom] / 1000.0) / 1e9
        peak = float(peaks["hbm_gbs"])
        roof = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src, "algorithmic": alg}
    tr = ncu_traffic(dom)
    if tr:
        roof["traffic"] = tr["dram_bytes_per_lattice"] * B
        roof["traffic_source"] = tr["source"]
    if dom == "batched_settle":
        # What binds this kernel is the shared-memory data pipe, not HBM: the gathered vector lives in shared
        # memory, everything else in registers.  Algorithmic shared-memory bytes per lattice,
        # multi-shift kernel: 1 + it_u gather passes of Np*kp*16 B (+ 6 B of graph image per entry when the
        # graph is staged in shared memory, 0 when it sits in registers) and one 16-B store per row and pass.
        it_u = check["ustar_iters_mean"]
        Np, kp, G = 1280, (K_LAT + 3) // 4 * 4, D_LAT // 4
        passes = it_u + 1.0
        smem_bytes = G * passes * (Np * kp * 16.0 + Np * 16.0) * B
        sm_clk = float(clocks.get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0) * 1e6
        smem_peak = 148 * 128.0 * sm_clk / 1e9   # 128 B per clock per SM
        onchip = {"bound": "smem", "achieved": smem_bytes / (kern_ms[dom] / 1000.0) / 1e9, "peak": smem_peak,
                  "unit": "GB/s", "algorithmic": "G*(1+it_u)*(Np*kp*16 + Np*16) bytes per lattice, conflict-free; "
                  "peak = 148 SMs x 128 B/clk x sampled SM clock"}
        onchip["frac"] = onchip["achieved"] / onchip["peak"]
        if tr and "l1tex_data_pipe_pct" in tr:
            onchip["ncu_l1tex_data_pipe_pct"] = tr["l1tex_data_pipe_pct"]
            onchip["ncu_source"] = tr.get("l1tex_source")
        roof["onchip"] = onchip
        roof["note"] = ("everything but Y in / U out stays on chip, so HBM is not the binding resource; the kernel "
                        "alternates between shared-memory gather passes and latency-bound reduction phases "
                        "(see `onchip` and profiles/)")
    roof["kernel_ms_per_step"] = kern_ms
    roof["share_of_step"] = {k: v / (ms / args.steps) for k, v in kern_ms.items()}

    del Y, Y_host, psi_host
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"serving batch: {B} independent lattices N={N_LAT} D={D_LAT} "
                               f"k={K_LAT} per GPU; step = build + settle(12,1e-3) + light receipt",
                   "lattices_per_step_per_gpu": B, "parallelism": f"replicas x{world}",
                   "l2": "inputs (7.5 GB/step at B=4096) larger than L2", "knn_engine": engine,
                   "e2e_call": f"settle_host_batch(chunk={args.chunk}): pinned host Y/psi -> results in pinned host"},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": int(res_host.numel() * 8),
                "returns": "per-lattice {iters, res, ustar_iters, ustar_res, deltaH} (a receipt-only request); "
                           "see e2e_with_U for the variant that also copies the settled state back",
                "h2d_ceiling": {"gbs_per_gpu": h2d_bytes / (ms_h2d / e2e_steps / 1000.0) / 1e9,
                                "ms_per_step": ms_h2d / e2e_steps,
                                "lattices_per_s_if_copy_bound": world * B / (ms_h2d / e2e_steps / 1000.0),
                                "how": f"the step's pinned->device copies alone, {world} rank(s) at once"}},
        "e2e_with_U": {"value": world * B * e2e_steps / (ms_e2e_u / 1000.0), "unit": UNIT,
                       "ms_per_step": ms_e2e_u / e2e_steps, "h2d_bytes_per_step": h2d_bytes,
                       "d2h_bytes_per_step": int(B * N_LAT * D_LAT * 4 + res_host.numel() * 8)},
        # normalize, knn_tc2, rescore, exact_rows, assemble x3, pack, settle, resolve, settle(fix), finalize
        "gpu_launches": 12 * args.steps,
        "clocks": clocks,
        "roofline": roof,
        "check": check,
    }


# ----------------------------------------------------------------------------- one lattice through the class
def measure_single_lattices() -> dict:
    """BASELINE.json configs[0]/[1] through `OscillinkLattice` with NumPy in and out, the protocol of the
    reference's scripts/benchmark.py:45-70 (build = ctor, settle(12, 1e-3), receipt; 1 warm-up + median of
    5), next to the reference restatement on this box's CPU and the README's published laptop numbers."""
    import numpy as np

    from oscillink_b200 import OscillinkLattice

    def inputs(N, D, head):
        rs = np.random.RandomState(0)
        Y = rs.randn(N, D).astype(np.float32)
        psi = Y[:head].mean(axis=0)
        return Y, (psi / (np.linalg.norm(psi) + 1e-12)).astype(np.float32)

    def med(v):
        v = sorted(v)
        return v[len(v) // 2]

    out = {}
    for name, N, D, k, head, chain in (("quickstart_N80_D128_k6", 80, 128, 6, 20, [2, 5, 7, 9]),
                                       ("readme_N1200_D384_k8", 1200, 384, 8, 32, None)):
        Y, psi = inputs(N, D, head)
        t = {"build_ms": [], "settle_ms": [], "receipt_light_ms": [], "receipt_full_ms": [], "e2e_light_ms": []}
        st = rec = None
        for rep in range(6):
            t0 = time.perf_counter()
            lat = OscillinkLattice(Y, kneighbors=k, deterministic_k=True)
            lat.set_query(psi)
            if chain is not None:
                lat.add_chain(chain, lamP=0.2)
            t1 = time.perf_counter()
            st = lat.settle(max_iters=12, tol=1e-3)
            t2 = time.perf_counter()
            lat.set_receipt_detail("light")
            rec = lat.receipt()
            t3 = time.perf_counter()
            lat.set_receipt_detail("full")
            full = lat.receipt()
            t4 = time.perf_counter()
            U = lat.U  # the settled state as a NumPy array (D2H)
            if rep == 0:
                continue  # warm-up
            t["build_ms"].append(1e3 * (t1 - t0))
            t["settle_ms"].append(1e3 * (t2 - t1))
            t["receipt_light_ms"].append(1e3 * (t3 - t2))
            t["receipt_full_ms"].append(1e3 * (t4 - t3))
            t["e2e_light_ms"].append(1e3 * (t3 - t0))
        out[name] = {k2: med(v) for k2, v in t.items()}
        out[name].update({"iters": int(st["iters"]), "res": float(st["res"]),
                          "deltaH": float(rec["deltaH_total"]), "null_points": len(full["null_points"]),
                          "U_shape": list(U.shape)})
    # the same protocol on the host cores with the dense restatement of the reference (oracle, checker leg)
    try:
        from oracle.dense import DenseLattice

        Y, psi = inputs(1200, 384, 32)
        t = {"build_ms": [], "settle_ms": [], "receipt_light_ms": []}
        for rep in range(4):
            t0 = time.perf_counter()
            o = DenseLattice(Y, k=8, deterministic=True)
            o.set_query(psi)
            t1 = time.perf_counter()
            o.settle(max_iters=12, tol=1e-3)
            t2 = time.perf_counter()
            us, _, _ = o.stationary()
            o.delta_h(us)
            t3 = time.perf_counter()
            if rep:
                t["build_ms"].append(1e3 * (t1 - t0))
                t["settle_ms"].append(1e3 * (t2 - t1))
                t["receipt_light_ms"].append(1e3 * (t3 - t2))
        out["cpu_port_readme_N1200_D384_k8"] = dict({k2: med(v) for k2, v in t.items()},
                                                    kind="port (oracle/dense.py, deterministic_k=True as "
                                                         "scripts/benchmark.py uses; the library default "
                                                         "argpartition path builds ~4x faster, BASELINE.md 2)",
                                                    blas_threads="default")
    except Exception as e:  # the checker is not part of the product
        out["cpu_port_readme_N1200_D384_k8"] = {"unavailable": repr(e)}
    out["reference_published_laptop_ms"] = {"build": 18, "settle": 10, "receipt_light": 3,
                                            "source": "README.md:176-182 (N~1200, laptop CPU)"}
    out["protocol"] = "scripts/benchmark.py:45-70: time.perf_counter, NumPy in/out, 1 warm-up + median of 5"
    return out


# ----------------------------------------------------------------------------- large single lattice
def measure_large(N, D, k, *, steps, warmup, chain_len=0, partition="both", world=1, rank=0, local=0,
                  halos=None, parity_rows=64, residual_rows=256):
    """BASELINE.json configs[3]/[4]: ONE big lattice (N up to 10M), kNN build + PCG settle, sharded over
    the GPUs.  A step = one settle(12, 1e-3) from U = Y on the built lattice (the metric's `ms/settle`);
    build, U* + deltaH, per-kernel rooflines and the sampled parity checks ride along.
    Returns the JSON-able result dict on rank 0 (None elsewhere)."""
    import torch
    import torch.distributed as dist

    from oscillink_b200 import _cabi
    from oscillink_b200.sharded_api import ShardedLattice, _NativeKernels, gather_rows, shard_bounds
    from tools import sampled_check as sc

    dev = torch.device("cuda", local)
    row0, n_local, _ = shard_bounds(N, world, rank)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    Y_local = torch.randn((n_local, D), generator=gen, device=dev, dtype=torch.float32)
    peaks, peak_src = load_peaks()
    peak = float(peaks["hbm_gbs"])

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
    lat = ShardedLattice(Y_local, N, kneighbors=k, mode="rows")
    e1.record()
    barrier()
    build_ms = max_over_ranks(e0.elapsed_time(e1))

    # ---- sampled-row parity of the graph: exact mutual-kNN sets from an exhaustive fp64 scan
    parity = None
    if parity_rows > 0:
        t0 = time.time()
        g = torch.Generator()
        g.manual_seed(20260)
        sample = torch.randperm(N, generator=g)[:parity_rows].sort().values.to(dev)
        want, min_gap, hop = sc.mutual_sets(Y_local, row0, N, sample, lat.k, group=None if world == 1 else lat.group)
        bad = sc.compare_neighbour_sets(lat._nbr[sample], want)
        barrier()
        parity = {"rows": int(parity_rows), "mismatches": int(bad), "rows_scanned_exhaustively": int(parity_rows + hop),
                  "min_k_gap": min_gap, "seconds": time.time() - t0,
                  "oracle": "tools/sampled_check.py: torch fp64 dot of the fp32-normalised rows against all N rows, "
                            "rounded once to fp32, (similarity desc, index asc) -- graph.py:35-52,64-65"}
        torch.cuda.empty_cache()

    head = Y_local[:32].mean(dim=0) if n_local >= 32 else torch.zeros(D, device=dev)
    if world > 1:
        dist.broadcast(head, src=0)
    psi = (head / (head.norm() + 1e-12)).cpu().numpy()
    psi_host = torch.from_numpy(psi.copy()).pin_memory()
    psi_dev = torch.from_numpy(psi.copy()).to(dev)
    lat.set_query(psi)
    if chain_len >= 2:
        lat.add_chain(list(range(chain_len)), lamP=0.2)
    lamP_eff = 0.2 if chain_len >= 2 else 0.0
    stride = max(1, (N - 2048) // max(residual_rows, 1))
    res_rows = (1024 + stride * torch.arange(residual_rows, device=dev)).clamp_(max=N - 1).unique()

    def residual_sample(vec, settle):
        """fp64 operator residual on a strided row sample -> estimate of max_c ||r_c||_2 over all N rows."""
        if residual_rows <= 0:
            return None
        Y0, U0 = lat._Y, lat._Y  # the timed settles start from U = Y
        r = sc.operator_residual_rows(
            res_rows, lambda ids: lat.rows_of(vec, ids), lambda ids: lat.rows_of(Y0, ids),
            lambda ids: lat.rows_of(U0, ids), lat._nbr, lat._W, psi_dev, (lat.lamG, lat.lamC, lat.lamQ),
            settle=settle, dt=1.0, lamP_eff=lamP_eff)
        est = float(torch.sqrt((r * r).sum(dim=0) * (N / r.shape[0])).max().item())
        # the TRUE residual of an fp32-stored iterate cannot go below ~eps32 * ||A|| * ||u_c||: the solver's
        # own (recurrence) residual does, which is why the two are compared through this floor
        un = float(lat.rows_of(vec, res_rows).double().pow(2).mean().sqrt().item()) * (N ** 0.5)
        floor = 2.0 * 1.19e-7 * (1.0 + lat.lamG + 2.0 * lat.lamC + lat.lamQ) * un
        return {"rows": int(r.shape[0]), "max_abs": float(r.abs().max().item()), "column_norm_estimate": est,
                "fp32_storage_floor": floor}

    def measure_mode(part, label):
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

        for _ in range(max(warmup, 1)):
            one_settle()
        with ClockSampler(local) as clk:
            tot, st = 0.0, None
            for _ in range(steps):
                ms, st = one_settle()
                tot += ms
        clocks = clk.summary()
        ms_settle = max_over_ranks(tot / steps)
        res_u = residual_sample(lat._U, settle=True)

        barrier()
        e0.record()
        rec = lat.receipt()
        e1.record()
        barrier()
        receipt_ms = max_over_ranks(e0.elapsed_time(e1))
        res_us = residual_sample(lat._Ustar, settle=False)

        # ---- per-kernel timings on the live state (one rank-local launch each, CUDA events)
        Dl = D if part == "rows" else lat.Dl
        n_loc = n_local if part == "rows" else N
        lat._Ustar = None
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
            "pcg_update": t_of(lambda: kf.update(ones, ones, with_x=False)),
            "pcg_pupdate_x": t_of(lambda: kf.pupdate_x(ones, ones, ones)),
        }
        halo = None
        if world > 1 and part == "rows":
            barrier()
            halo = {"strategy": lat.halo, "allgather_ms": max_over_ranks(t_of(lambda: gather(kf.P)))}
            if lat._pull is not None and lat.halo == "pull":
                ds = lat._dist_struct()
                flag = torch.zeros(1, dtype=torch.float32, device=dev)
                st_ptr = torch.cuda.current_stream().cuda_stream
                pull_ms = max_over_ranks(t_of(lambda: _cabi.check(
                    lat._lib.osc_dist_halo_exchange(ctypes.byref(ds), D, flag.data_ptr(), st_ptr))))
                halo["pull_ms"] = pull_ms
                halo["pull_gbs_per_gpu"] = float(lat._pull.n_halo) * D * 4.0 / (pull_ms / 1e3) / 1e9
                halo["nvlink_peak_gbs"] = 770.0  # measured peer copy per direction (B200_PROFILING.md)
            if lat._pull is not None:
                halo["pull_rows"] = int(lat._pull.n_halo)
                halo["pull_share_of_remote_rows"] = float(lat._pull.halo_fraction)
                halo["pull_bytes"] = float(lat._pull.n_halo) * D * 4.0
            barrier()
        del p_all
        alg = {  # algorithmic bytes per launch
            "pcg_spmm": (nnz_loc / max(n_loc, 1) + 2.0) * V + 8.0 * nnz_loc,
            "pcg_update": 3.0 * V,     # r, Ap -> r (+ the column dots)
            "pcg_pupdate_x": 5.0 * V,  # p, x, r -> p, x: the x update rides with the p update
        }
        gbs = {n: alg[n] / (kms[n] / 1000.0) / 1e9 for n in alg}
        iters = int(st["iters"])
        # the last iteration updates x alone (p, x -> x: 3 V); the first residual forms the right-hand side in
        # place (no setup pass): next to a SpMM's bytes it reads the Y row and writes z0 (2 V)
        solve_bytes = ((iters + 1) * alg["pcg_spmm"] + iters * alg["pcg_update"]
                       + (iters - 1) * alg["pcg_pupdate_x"] + 3.0 * V + 2.0 * V)
        roof = {"kernel": "pcg_spmm_kernel", "bound": "hbm", "achieved": gbs["pcg_spmm"], "peak": peak,
                "unit": "GB/s", "frac": gbs["pcg_spmm"] / peak, "traffic": None, "peak_source": peak_src,
                "kernel_ms": kms, "kernel_gbs": gbs,
                "whole_settle": {"algorithmic_bytes_per_gpu": solve_bytes,
                                 "achieved_gbs_per_gpu": solve_bytes / (ms_settle / 1e3) / 1e9,
                                 "frac": solve_bytes / (ms_settle / 1e3) / 1e9 / peak},
                "row_bytes": Dl * 4}
        del kf
        torch.cuda.empty_cache()
        return {"parallelism": label, "value": ms_settle, "unit": "ms", "receipt_light_ms": receipt_ms,
                "roofline": roof, "halo": halo, "clocks": clocks,
                "check": {"iters": iters, "res": float(st["res"]), "ustar_iters": rec["meta"]["ustar_iters"],
                          "ustar_res": rec["meta"]["ustar_res"], "deltaH": rec["deltaH_total"],
                          "settle_residual_sample": res_u, "ustar_residual_sample": res_us}}

    modes = []
    if partition in ("rows", "both"):
        hs = halos or (["pull", "allgather"] if world > 1 else [None])
        for h in hs:
            if h is not None:
                lat.set_halo(h)
            label = f"rows x{world}" + (f" ({lat.halo} halo, {'C ABI + NCCL' if lat._use_c_path() else 'python phases'})"
                                        if world > 1 else "")
            modes.append(measure_mode("rows", label))
    if partition in ("columns", "both") and (world > 1 or partition == "columns") and D % (4 * world) == 0:
        lat.repartition("columns")   # same graph, state transposed by one all-to-all
        modes.append(measure_mode("columns", f"columns x{world} (no halo)"))
    nnz = float(lat.nnz.item())
    n_exh = int(getattr(lat, "n_exhaustive", torch.zeros(1)).item())
    engine = getattr(lat, "engine_used", "?")
    lat.close()
    del lat
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    best = min(modes, key=lambda m: m["value"])
    return {
        "metric": f"ms/settle at N={N},D={D}", "value": best["value"], "unit": "ms", "n_gpus": world,
        "steps": steps, "warmup": max(warmup, 1), "ms_per_step": best["value"],
        "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"one lattice N={N} D={D} k={k} chain_len={chain_len}: "
                               "step = settle(12,1e-3) from U=Y on the built mutual-kNN graph",
                   "parallelism": best["parallelism"], "knn_engine": engine,
                   "l2": f"vectors ({N * D * 4.0 / world / 1e9:.2f} GB each per GPU) larger than L2"},
        "e2e": {"value": best["value"], "unit": "ms", "h2d_bytes_per_step": int(D * 4), "d2h_bytes_per_step": 8,
                "note": "set_query(host psi) + settle() -> host {iters,res}; the lattice state is device-resident by API"},
        "build_ms": build_ms, "partitions": modes, "parity_sample": parity,
        "graph": {"nnz": nnz, "avg_degree": nnz / max(N, 1), "rows_recomputed_exhaustively": n_exh},
    }


# ----------------------------------------------------------------------------- our arm
def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    world, rank, local = dist_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    line = measure_serving(args, world, rank, local)
    if rank == 0:
        workers = host_workers()
        sample = max(2 * workers, 32)
        cpu = cpu_reference_rate(sample, workers)
        line["cpu_baseline"] = {"value": cpu["value"], "unit": UNIT, "cores": workers, "kind": "port",
                                "sample": f"{sample} lattices, dense oracle (oracle/dense.py), "
                                          f"{workers} processes x 1 BLAS thread"}
        if not args.no_single:
            line["single_lattice"] = measure_single_lattices()
    if world > 1:
        dist.barrier()
    if not args.no_large:
        # the second half of BASELINE.json's metric: ms/settle of ONE lattice sharded over these GPUs
        big = measure_large(args.N, args.D, args.k, steps=3, warmup=2, chain_len=args.chain_len,
                            partition="both", world=world, rank=rank, local=local)
        if rank == 0:
            line["large_lattice"] = big
        if not args.no_large_1m:
            mid = measure_large(1_000_000, 768, 16, steps=3, warmup=2, chain_len=0, partition="both",
                                world=world, rank=rank, local=local)
            if rank == 0:
                line["large_lattice_1M"] = mid
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_large(args) -> None:
    import torch
    import torch.distributed as dist

    world, rank, local = dist_env()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    line = measure_large(args.N, args.D, args.k, steps=args.steps, warmup=args.warmup, chain_len=args.chain_len,
                         partition=args.partition, world=world, rank=rank, local=local,
                         halos=[args.halo] if args.halo else None, parity_rows=args.parity_rows)
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
                    help="serving: the headline batch of N=1200 lattices (+ the large-lattice blocks); "
                         "large: one big lattice only (ms/settle)")
    ap.add_argument("--N", type=int, default=10_000_000)
    ap.add_argument("--D", type=int, default=384)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--chain-len", type=int, default=8)
    ap.add_argument("--partition", default="both", choices=["rows", "columns", "both"])
    ap.add_argument("--halo", default=None, choices=["pull", "allgather", "fused"],
                    help="large workload, rows partition: measure this halo strategy only")
    ap.add_argument("--parity-rows", type=int, default=64)
    ap.add_argument("--no-large", action="store_true",
                    help="serving workload only: skip the one-big-lattice blocks of the JSON line")
    ap.add_argument("--no-large-1m", action="store_true", help="skip the N=1M D=768 k=16 block")
    ap.add_argument("--no-single", action="store_true", help="skip the single-lattice latency block")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "large":
        run_large(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
