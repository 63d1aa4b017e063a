"""ctypes binding of include/oscillink_b200.h (the C-ABI drop-in boundary).

The library is built in-tree by `oscillink_b200.build` (nvcc, sm_100a).  There is no CPU
fallback: if the shared object is missing and cannot be built, importing the numeric API
fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import build as _build

_lock = threading.Lock()
_lib = None

OK, ERR_INVALID, ERR_CUDA, ERR_WORKSPACE, ERR_UNSUPPORTED = 0, 1, 2, 3, 4
KNN_AUTO, KNN_SIMT, KNN_TC, KNN_TC1, KNN_TCH = 0, 1, 2, 3, 4
ENGINE_NAMES = {KNN_SIMT: "simt", KNN_TC: "tc", KNN_TC1: "tc1", KNN_TCH: "tch"}
MODE_SETTLE, MODE_STATIONARY = 0, 1
KNN_EPS = 1e-5  # OSC_KNN_EPS
KNN_EPS_TC1 = 1.25e-3  # OSC_KNN_EPS_TC1

c_i32, c_i64, c_f32, c_f64, c_void_p, c_size_t = C.c_int32, C.c_int64, C.c_float, C.c_double, C.c_void_p, C.c_size_t


class Graph(C.Structure):
    _fields_ = [("batch", c_i64), ("N", c_i64), ("k", c_i32), ("_pad", c_i32), ("nbr", c_void_p),
                ("A", c_void_p), ("W", c_void_p), ("deg", c_void_p), ("sqrt_deg", c_void_p)]


class Chain(C.Structure):
    _fields_ = [("n_rows", c_i32), ("nnz", c_i32), ("rows", c_void_p), ("rowptr", c_void_p),
                ("col", c_void_p), ("Wp", c_void_p), ("Ap", c_void_p), ("slot", c_void_p)]


class Params(C.Structure):
    _fields_ = [("lamG", c_f32), ("lamC", c_f32), ("lamQ", c_f32), ("lamP", c_f32),
                ("chain_present", c_i32), ("_pad", c_i32)]


class PcgDims(C.Structure):
    _fields_ = [("N", c_i64), ("row0", c_i64), ("n_local", c_i64), ("D", c_i32), ("n_blocks", c_i32)]


class BatchedArgs(C.Structure):
    _fields_ = [("Y", c_void_p), ("U_in", c_void_p), ("psi", c_void_p), ("gates", c_void_p),
                ("U_out", c_void_p), ("Ustar_out", c_void_p), ("stats", c_void_p), ("deltaH", c_void_p),
                ("D", c_i32), ("do_settle", c_i32), ("do_ustar", c_i32), ("do_deltaH", c_i32),
                ("dt", c_f32), ("tol_settle", c_f64), ("tol_ustar", c_f64),
                ("max_iters_settle", c_i32), ("max_iters_ustar", c_i32), ("unresolved", c_void_p)]


PART_ROWS, PART_COLUMNS = 0, 1
HALO_ALLGATHER, HALO_PULL = 0, 1


class Dist(C.Structure):
    """osc_dist_t: one rank's view of a sharded lattice (include/oscillink_b200.h)."""
    _fields_ = [("nccl_comm", c_void_p), ("world", c_i32), ("rank", c_i32), ("partition", c_i32),
                ("halo", c_i32), ("N", c_i64), ("shard", c_i64), ("d_peer_P", c_void_p),
                ("P_block", c_void_p), ("halo_rows", c_void_p), ("halo_nbr", c_void_p), ("n_halo", c_i64),
                ("halo_below", c_i64)]


P = C.POINTER
# name -> (restype, argtypes); every symbol declared in include/oscillink_b200.h
PROTOTYPES = {
    "osc_abi_version": (C.c_int, []),
    "osc_last_error": (C.c_char_p, []),
    "osc_device_info": (C.c_int, [C.c_int, P(C.c_int), P(C.c_int), P(C.c_int)]),
    "osc_normalize_rows": (C.c_int, [c_void_p, c_i64, c_i32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "osc_normalize_rows_f16": (C.c_int, [c_void_p, c_i64, c_i32, c_void_p, c_void_p, c_void_p]),
    "osc_chain_build_size": (C.c_int, [c_void_p, c_i32, c_i64, P(c_i32), P(c_i32)]),
    "osc_chain_build": (C.c_int, [c_void_p, c_i32, c_void_p, c_i64] + [c_void_p] * 6),
    "osc_knn_candidates": (C.c_int, [c_void_p] * 6 + [c_i64, c_i64, c_i64, c_i64, c_i32, c_i32, c_i32,
                                                     c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "osc_knn_tc_supported": (C.c_int, [c_i64, c_i32, c_i32]),
    "osc_knn_plan": (C.c_int, [c_i64, c_i64, c_i32, c_i32, c_i32, P(c_i32), P(c_i32), P(c_f32)]),
    "osc_knn_candidates_workspace": (C.c_int, [c_i64, c_i64, c_i64, c_i32, c_i32, c_i32, P(c_size_t)]),
    "osc_knn_rescore": (C.c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_i32, c_void_p, c_i32, c_i32,
                                  c_void_p, c_void_p, c_void_p, c_void_p]),
    "osc_knn_rescore_workspace": (C.c_int, [c_i64, c_i64, P(c_size_t)]),
    "osc_knn_rescore_checked": (C.c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_i64, c_i32, c_void_p,
                                          c_void_p, c_i32, c_i32, c_f32, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_size_t, c_void_p]),
    "osc_knn_rescore_guarded": (C.c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_i64, c_i32, c_void_p,
                                          c_void_p, c_i32, c_i32, c_f32, c_i64, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_size_t, c_void_p]),
    "osc_knn_exhaustive_limit": (c_i64, [c_i64]),
    "osc_graph_assemble": (C.c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i32, c_f32, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "osc_knn_build_workspace": (C.c_int, [c_i64, c_i64, c_i32, c_i32, c_i32, P(c_size_t)]),
    "osc_knn_build": (C.c_int, [c_void_p, c_i64, c_i64, c_i32, c_i32, c_f32, c_i32, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                c_void_p]),
    "osc_pcg_plan": (C.c_int, [P(PcgDims), P(c_size_t)]),
    "osc_pcg_max_ell_width": (C.c_int, [c_i32]),
    "osc_pcg_setup": (C.c_int, [P(PcgDims), P(Params), c_i32, c_f32, c_i32, c_f32, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "osc_pcg_residual0": (C.c_int, [P(PcgDims), P(Graph), P(Chain), P(Params), c_i32, c_f32, c_i32,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "osc_pcg_spmm_dot": (C.c_int, [P(PcgDims), P(Graph), P(Chain), P(Params), c_i32, c_f32, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p]),
    "osc_enable_peer_access": (C.c_int, [c_i32]),
    "osc_peer_alloc": (C.c_int, [c_size_t, P(c_void_p), c_void_p]),
    "osc_peer_open": (C.c_int, [c_void_p, P(c_void_p)]),
    "osc_peer_close": (C.c_int, [c_void_p]),
    "osc_peer_free": (C.c_int, [c_void_p]),
    "osc_pcg_residual0_p2p": (C.c_int, [P(PcgDims), P(Graph), P(Chain), P(Params), c_i32, c_f32, c_i32,
                                        c_void_p, c_void_p, c_i32, c_i64, c_void_p, c_void_p, c_void_p,
                                        c_void_p]),
    "osc_pcg_spmm_dot_p2p": (C.c_int, [P(PcgDims), P(Graph), P(Chain), P(Params), c_i32, c_f32, c_void_p,
                                       c_void_p, c_i32, c_i64, c_void_p, c_void_p, c_void_p]),
    "osc_pcg_reduce": (C.c_int, [c_void_p, c_i32, c_i32, c_void_p, c_void_p, c_void_p]),
    "osc_pcg_update": (C.c_int, [P(PcgDims), P(Params), c_i32, c_f32, c_i32, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "osc_pcg_pupdate": (C.c_int, [P(PcgDims), P(Params), c_i32, c_f32, c_i32, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p]),
    "osc_pcg_pupdate_x": (C.c_int, [P(PcgDims), P(Params), c_i32, c_f32, c_i32, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_void_p]),
    "osc_pcg_solve": (C.c_int, [P(Graph), P(Chain), P(Params), c_i32, c_f32, c_i32, c_f32, c_i32, c_f64,
                                c_i32, c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_void_p, P(c_i32),
                                P(c_f32), c_void_p, c_size_t, c_void_p]),
    "osc_pcg_solve_system": (C.c_int, [P(Graph), P(Chain), P(Params), c_i32, c_f32, c_i32, c_f64, c_i32,
                                       c_void_p, c_i32, c_void_p, c_void_p, P(c_i32), P(c_f32), c_void_p,
                                       c_size_t, c_void_p]),
    "osc_dist_nccl_version": (C.c_int, [P(c_i32)]),
    "osc_dist_unique_id": (C.c_int, [c_void_p]),
    "osc_dist_comm_init": (C.c_int, [c_void_p, c_i32, c_i32, P(c_void_p)]),
    "osc_dist_comm_destroy": (C.c_int, [c_void_p]),
    "osc_dist_halo_plan_workspace": (C.c_int, [c_i64, P(c_size_t)]),
    "osc_dist_halo_plan": (C.c_int, [c_void_p, c_i64, c_i32, c_void_p, c_i64, c_i64, c_i64, c_i64, c_void_p,
                                     c_i64, c_void_p, c_void_p, P(c_i64), c_void_p, c_size_t, c_void_p]),
    "osc_dist_halo_exchange": (C.c_int, [P(Dist), c_i32, c_void_p, c_void_p]),
    "osc_dist_pcg_workspace": (C.c_int, [P(Dist), c_i64, c_i32, P(c_size_t)]),
    "osc_dist_pcg_solve": (C.c_int, [P(Dist), P(Graph), P(Chain), P(Params), c_i32, c_f32, c_i32, c_f32, c_i32,
                                     c_f64, c_i32, c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_void_p,
                                     P(c_i32), P(c_f32), c_void_p, c_size_t, c_void_p]),
    "osc_dist_delta_h": (C.c_int, [P(Dist), P(Graph), P(Chain), P(Params), c_void_p, c_void_p, c_void_p, c_i32,
                                   P(c_f64), c_void_p, c_size_t, c_void_p]),
    "osc_delta_h": (C.c_int, [P(Graph), P(Chain), P(Params), c_void_p, c_void_p, c_void_p, c_i32,
                              P(c_f64), c_void_p, c_size_t, c_void_p]),
    "osc_receipt_full": (C.c_int, [P(Graph), P(Params), c_void_p, c_void_p, c_void_p, c_void_p, c_i32,
                                   c_f32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p]),
    "osc_row_align": (C.c_int, [c_void_p, c_void_p, c_i64, c_i32, c_void_p, c_void_p]),
    "osc_pair_d2": (C.c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_i32, c_void_p, c_void_p]),
    "osc_mmr_workspace": (C.c_int, [c_i64, P(c_size_t)]),
    "osc_mmr_select": (C.c_int, [c_void_p, c_void_p, c_i64, c_i32, c_i32, c_void_p, c_void_p, c_size_t,
                                 c_void_p]),
    "osc_batched_supported": (C.c_int, [c_i64, c_i32, c_i32]),
    "osc_batched_workspace": (C.c_int, [c_i64, c_i64, c_i32, P(c_size_t)]),
    "osc_batched_settle": (C.c_int, [P(Graph), P(Params), P(BatchedArgs), c_void_p, c_size_t, c_void_p]),
}


class OscillinkNativeError(RuntimeError):
    """CUDA / workspace / unsupported-shape failure reported by the native library."""


def lib_path() -> str:
    return _build.LIB


def load():
    """Load (building in-tree if needed) the native library.  Raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = os.environ.get("OSC_B200_LIB")  # dev-only: A/B a previously built variant
        if path:
            lib = C.CDLL(path)
            for name, (res, args) in PROTOTYPES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
            return _lib
        path = _build.LIB
        if not os.path.exists(path) or (_build.is_stale() and os.path.exists(_build.NVCC)):
            _build.build()
        lib = C.CDLL(path)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)  # AttributeError here == missing symbol: fail loudly
            fn.restype = res
            fn.argtypes = args
        if lib.osc_abi_version() != 1:
            raise OscillinkNativeError("libosc_b200 ABI version mismatch")
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc == OK:
        return
    msg = load().osc_last_error().decode("utf-8", "replace")
    if rc == ERR_INVALID:
        raise ValueError(msg)
    raise OscillinkNativeError(f"{what or 'native call'} failed (code {rc}): {msg}")


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def knn_plan(n_rows: int, N: int, D: int, k: int, flags: int) -> tuple[int, int, float]:
    """(engine, candidate width kc, error bound eps) that `flags` resolves to -- osc_knn_plan."""
    eng, kc, eps = c_i32(0), c_i32(0), c_f32(0.0)
    check(load().osc_knn_plan(n_rows, N, D, k, flags, C.byref(eng), C.byref(kc), C.byref(eps)), "osc_knn_plan")
    return eng.value, kc.value, eps.value
