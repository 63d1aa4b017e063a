"""Dev experiment: does the exact re-scoring of one half of a serving batch overlap with the candidate GEMM (K1)
of the other half when they run on two streams?  (K1: one persistent 226 KB CTA per SM, ALU-pipe bound in its
epilogue warps; re-scoring: small blocks, L2-latency bound.)
    B=4096 python tools/dev_overlap_build.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, ".")
from oscillink_b200 import _cabi  # noqa: E402

B = int(os.environ.get("B", "4096")); N, D, k = 1200, 384, 8
PARTS = int(os.environ.get("PARTS", "2"))
lib = _cabi.load()
dev = torch.device("cuda:0")
g = torch.Generator(device="cuda"); g.manual_seed(0)
Y = torch.randn((B, N, D), generator=g, device=dev)
eng, kc, eps = _cabi.knn_plan(N, N, D, k, _cabi.KNN_AUTO)
print("engine", _cabi.ENGINE_NAMES[eng], "kc", kc)
Yn = torch.empty_like(Y); hi = torch.empty(Y.shape, dtype=torch.float16, device=dev)
cand_idx = torch.empty((B, N, kc), dtype=torch.int32, device=dev); cand_sim = torch.empty((B, N, kc), device=dev)
top_idx = torch.empty((B, N, k), dtype=torch.int32, device=dev); top_sim = torch.empty((B, N, k), device=dev)
gap = torch.empty((B, N), device=dev)
nex = torch.zeros(PARTS + 1, dtype=torch.int32, device=dev)
need = C.c_size_t(0)
_cabi.check(lib.osc_knn_rescore_workspace(B, N, C.byref(need)))
ws_all = torch.empty(need.value, dtype=torch.uint8, device=dev)
Bh = B // PARTS
_cabi.check(lib.osc_knn_rescore_workspace(Bh, N, C.byref(need)))
ws_h = [torch.empty(need.value, dtype=torch.uint8, device=dev) for _ in range(PARTS)]
limit = int(lib.osc_knn_exhaustive_limit(B * N))
st0 = torch.cuda.current_stream()
_cabi.check(lib.osc_normalize_rows_f16(Y.data_ptr(), B * N, D, Yn.data_ptr(), hi.data_ptr(), st0.cuda_stream))


def k1(b0, nb, st):
    o = b0 * N
    _cabi.check(lib.osc_knn_candidates(Yn.data_ptr() + o * D * 4, Yn.data_ptr() + o * D * 4, hi.data_ptr() + o * D * 2, None,
                                       hi.data_ptr() + o * D * 2, None, nb, N, 0, N, D, kc, eng,
                                       cand_idx.data_ptr() + o * kc * 4, cand_sim.data_ptr() + o * kc * 4, None, 0, st))


def rescore(b0, nb, ws, cnt_slot, st):
    o = b0 * N
    _cabi.check(lib.osc_knn_rescore_guarded(Yn.data_ptr() + o * D * 4, Yn.data_ptr() + o * D * 4, nb, N, 0, N, D,
                                            cand_idx.data_ptr() + o * kc * 4, cand_sim.data_ptr() + o * kc * 4, kc, k, eps,
                                            limit, top_idx.data_ptr() + o * k * 4, top_sim.data_ptr() + o * k * 4,
                                            gap.data_ptr() + o * 4, nex.data_ptr() + 4 * cnt_slot, ws.data_ptr(), ws.numel(), st))


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def sequential():
    k1(0, B, st0.cuda_stream)
    rescore(0, B, ws_all, PARTS, st0.cuda_stream)


side = [torch.cuda.Stream() for _ in range(PARTS)]


def overlapped():
    # K1 of the parts runs back to back on the main stream; each part's re-scoring goes to its own side stream
    done = []
    for p in range(PARTS):
        k1(p * Bh, Bh, st0.cuda_stream)
        ev = torch.cuda.Event(); ev.record(st0)
        side[p].wait_event(ev)
        rescore(p * Bh, Bh, ws_h[p], p, side[p].cuda_stream)
        e2 = torch.cuda.Event(); e2.record(side[p]); done.append(e2)
    for e in done:
        st0.wait_event(e)


t_seq = timed(sequential)
ref = (top_idx.clone(), top_sim.clone())
t_ovl = timed(overlapped)
same = bool(torch.equal(ref[0], top_idx) and torch.equal(ref[1], top_sim))
print({"B": B, "parts": PARTS, "sequential_ms": t_seq, "overlapped_ms": t_ovl, "identical": same})
