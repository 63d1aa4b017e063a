"""Host-side mirror of the reference `OscillinkLattice` (oscillink/core/lattice.py:23-1014).

Same constructor, methods, attributes, return shapes, ValueErrors and logging events as the
reference class; every numeric step is a call into the C-ABI library (include/oscillink_b200.h)
running sm_100a kernels.  State lives in HBM as torch tensors (torch is used for allocation,
streams and host<->device copies only).  Hashing, HMAC signing, JSON, callbacks and caches stay
on the host exactly like the reference.

There is no CPU fallback: constructing a lattice without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import hmac
import json
import os
import time
from collections import deque
from typing import Any

import numpy as np
import torch

from . import _cabi
from ._cabi import Chain, Graph, Params

__all__ = ["OscillinkLattice", "Oscillink", "json_line_logger", "REFERENCE_VERSION"]

# version string reported in receipts / exported state: the reference release this mirrors
REFERENCE_VERSION = "0.1.13"

_F32 = np.float32


def _require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError(
            "oscillink_b200 needs a CUDA device (sm_100a); there is no CPU fallback by design"
        )
    return torch.device("cuda", torch.cuda.current_device())


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


class _Workspace:
    """Grow-only device scratch buffer owned by one lattice (one lattice == one host thread)."""

    def __init__(self, device):
        self.device = device
        self.buf = None

    def get(self, nbytes: int):
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=self.device)
        return self.buf


class OscillinkLattice:
    """B200-native short-term coherence lattice with the reference's public surface."""

    # ------------------------------------------------------------------ construction
    def __init__(
        self,
        Y: np.ndarray,
        kneighbors: int = 6,
        row_cap_val: float = 1.0,
        lamG: float = 1.0,
        lamC: float = 0.5,
        lamQ: float = 4.0,
        deterministic_k: bool = False,
        neighbor_seed: int | None = None,
    ):
        # same validation, same messages (lattice.py:45-53)
        if not isinstance(Y, np.ndarray) or Y.ndim != 2:
            raise ValueError("Y must be a 2D numpy array")
        if kneighbors < 1:
            raise ValueError("kneighbors must be >= 1")
        if lamG <= 0:
            raise ValueError("lamG must be > 0 for SPD")
        if lamC < 0:
            raise ValueError("lamC must be >= 0")
        if lamQ < 0:
            raise ValueError("lamQ must be >= 0")
        self._dev = _require_cuda()
        self._lib = _cabi.load()
        self._ws = _Workspace(self._dev)

        self.N, self.D = int(Y.shape[0]), int(Y.shape[1])
        host_Y = np.ascontiguousarray(Y, dtype=_F32)
        self._dY = torch.from_numpy(host_Y).to(self._dev)
        self._dU = self._dY.clone()
        self._hY: np.ndarray | None = host_Y.copy()
        self._hU: np.ndarray | None = None

        self._kneighbors = min(int(kneighbors), max(1, self.N - 1))  # lattice.py:60
        self._deterministic_k = bool(deterministic_k)
        self._neighbor_seed = neighbor_seed  # recorded, inert (seeded jitter)
        self._row_cap_val = float(row_cap_val)
        self._knn_engine = _cabi.KNN_AUTO
        self._build_graph()

        self._hB = np.ones(self.N, dtype=_F32)
        self._hpsi = np.zeros(self.D, dtype=_F32)
        self._dB = torch.ones(self.N, dtype=torch.float32, device=self._dev)
        self._dpsi = torch.zeros(self.D, dtype=torch.float32, device=self._dev)

        self.lamG, self.lamC, self.lamQ = lamG, lamC, lamQ
        self.lamP = 0.0
        self._chain: dict[str, Any] | None = None
        self._chain_nodes: list[int] | None = None
        self.last: dict[str, Any] = {"iters": 0, "res": None, "t_ms": None}

        self._Ustar_cache = None  # device tensor
        self._Ustar_host: np.ndarray | None = None
        self._Ustar_sig: str | None = None
        self.stats: dict[str, int] = {"ustar_solves": 0, "ustar_cache_hits": 0}
        self._settle_callbacks: list = []
        self._logger = None
        self._receipt_secret: bytes | None = None
        self._signature_mode = "minimal"
        self._receipt_detail = "full"
        self._last_dynamics: dict[str, Any] | None = None
        self._log(
            "init",
            {
                "N": self.N,
                "D": self.D,
                "kneighbors_requested": kneighbors,
                "kneighbors_effective": self._kneighbors,
                "deterministic_k": self._deterministic_k,
                "neighbor_seed": self._neighbor_seed,
            },
        )

    # ------------------------------------------------------------------ graph (K1 + K1b)
    def _build_graph(self) -> None:
        """graph.py:29-93 on device: fused similarity/top-k, canonical rescoring, assembly."""
        t0 = time.time()
        N, D, k = self.N, self.D, self._kneighbors
        dev = self._dev
        self._nbr = torch.full((max(N, 1), k), -1, dtype=torch.int32, device=dev)[:N]
        self._A = torch.zeros((max(N, 1), k), dtype=torch.float32, device=dev)[:N]
        self._W = torch.zeros((max(N, 1), k), dtype=torch.float32, device=dev)[:N]
        self._deg = torch.zeros(max(N, 1), dtype=torch.int32, device=dev)[:N]
        self._sd = torch.zeros(max(N, 1), dtype=torch.float32, device=dev)[:N]
        self._gap = torch.full((max(N, 1),), float("inf"), dtype=torch.float32, device=dev)[:N]
        nnz = torch.zeros(1, dtype=torch.int64, device=dev)
        if N > 0:
            need = C.c_size_t(0)
            _cabi.check(self._lib.osc_knn_build_workspace(1, N, D, k, self._knn_engine, C.byref(need)))
            ws = self._ws.get(need.value)
            _cabi.check(
                self._lib.osc_knn_build(
                    self._dY.data_ptr(), 1, N, D, k, self._row_cap_val, self._knn_engine,
                    self._nbr.data_ptr(), self._A.data_ptr(), self._W.data_ptr(), self._deg.data_ptr(),
                    self._sd.data_ptr(), nnz.data_ptr(), self._gap.data_ptr(), ws.data_ptr(),
                    ws.numel(), _stream_ptr(),
                ),
                "osc_knn_build",
            )
        torch.cuda.current_stream().synchronize()
        self._graph_build_ms = 1000.0 * (time.time() - t0)
        self._invalidate_graph_views()

    def _invalidate_graph_views(self) -> None:
        self._h_ell = None
        self._h_A = None
        self._h_L = None
        self._h_sd = None
        self._adj_sig = None
        self._nnz_pos = None

    def _graph_struct(self) -> Graph:
        return Graph(1, self.N, int(self._nbr.shape[1]) if self._nbr.ndim == 2 else 1, 0,
                     self._nbr.data_ptr(), self._A.data_ptr(), self._W.data_ptr(),
                     self._deg.data_ptr(), self._sd.data_ptr())

    def _params_struct(self) -> Params:
        return Params(float(self.lamG), float(self.lamC), float(self.lamQ), float(self.lamP),
                      1 if self._chain is not None else 0, 0)

    def _chain_struct(self):
        if self._chain is None:
            return None
        c = self._chain
        return Chain(c["n_rows"], c["nnz"], c["rows"].data_ptr(), c["rowptr"].data_ptr(),
                     c["col"].data_ptr(), c["Wp"].data_ptr(), c["Ap"].data_ptr(), c["slot"].data_ptr())

    def _ell_host(self):
        if self._h_ell is None:
            self._h_ell = (self._nbr.cpu().numpy(), self._A.cpu().numpy())
        return self._h_ell

    # ---- dense / host views the reference exposes as plain attributes
    @property
    def Y(self) -> np.ndarray:
        if self._hY is None:
            self._hY = self._dY.cpu().numpy()
        return self._hY

    @property
    def U(self) -> np.ndarray:
        if self._hU is None:
            self._hU = self._dU.cpu().numpy()
        return self._hU

    @U.setter
    def U(self, value: np.ndarray) -> None:
        arr = np.ascontiguousarray(value, dtype=_F32)
        if arr.shape != (self.N, self.D):
            raise ValueError("U shape mismatch")
        self._dU = torch.from_numpy(arr).to(self._dev)
        self._hU = arr.copy()

    @property
    def B_diag(self) -> np.ndarray:
        return self._hB

    @property
    def psi(self) -> np.ndarray:
        return self._hpsi

    @property
    def sqrt_deg(self) -> np.ndarray:
        if self._h_sd is None:
            self._h_sd = self._sd.cpu().numpy()
        return self._h_sd

    @property
    def A(self) -> np.ndarray:
        """Dense (N,N) adjacency, materialised lazily from the ELL graph (small N only)."""
        if self._h_A is None:
            nbr, a = self._ell_host()
            dense = np.zeros((self.N, self.N), dtype=_F32)
            if self.N:
                r, t = np.nonzero(nbr >= 0)
                dense[r, nbr[r, t]] = a[r, t]
            self._h_A = dense
        return self._h_A

    @A.setter
    def A(self, dense: np.ndarray) -> None:
        self._load_dense_adjacency(np.asarray(dense, dtype=_F32))

    @property
    def L_sym(self) -> np.ndarray:
        if self._h_L is None:
            nbr = self._ell_host()[0]
            w = self._W.cpu().numpy()
            L = np.eye(self.N, dtype=_F32)
            if self.N:
                r, t = np.nonzero(nbr >= 0)
                L[r, nbr[r, t]] -= w[r, t]
            self._h_L = L
        return self._h_L

    @property
    def A_path(self) -> np.ndarray | None:
        if self._chain is None:
            return None
        dense = np.zeros((self.N, self.N), dtype=_F32)
        for (u, v), w in self._chain["ap_host"].items():
            dense[u, v] = w
        return dense

    @property
    def L_path(self) -> np.ndarray | None:
        if self._chain is None:
            return None
        dense = np.eye(self.N, dtype=_F32)
        for (u, v), w in self._chain["wp_host"].items():
            dense[u, v] -= w
        return dense

    def _load_ell_adjacency(self, nbr: np.ndarray, a: np.ndarray) -> None:
        """from_state with the sparse graph format: adopt ELL rows (column ids ascending, -1 padded)
        and recompute sqrt_deg / normalised weights as graph.py:87-90 does (row sums in fp32)."""
        if nbr.ndim != 2 or nbr.shape != a.shape or nbr.shape[0] != self.N:
            raise ValueError("A_ell shape mismatch")
        if nbr.shape[1] == 0:
            nbr = np.full((self.N, 1), -1, dtype=np.int32)
            a = np.zeros((self.N, 1), dtype=_F32)
        nbr = np.asarray(nbr).astype(np.int64)
        if np.any(nbr >= self.N) or np.any(nbr < -1):
            raise ValueError("A_ell column index out of bounds")
        # the kernels read the first deg slots of a row: move the valid entries to the front, ascending
        # (padding anywhere in the row and unsorted rows are accepted; duplicates are not)
        key = np.where(nbr < 0, self.N, nbr)
        order = np.argsort(key, axis=1, kind="stable")
        nbr = np.take_along_axis(nbr, order, axis=1)
        a = np.take_along_axis(np.asarray(a, dtype=_F32), order, axis=1)
        if nbr.shape[1] > 1 and np.any((nbr[:, 1:] == nbr[:, :-1]) & (nbr[:, 1:] >= 0)):
            raise ValueError("A_ell has duplicate column indices in a row")
        self._check_ell_width(nbr.shape[1])
        a = np.where(nbr < 0, _F32(0), a).astype(_F32)
        d = a.sum(axis=1, dtype=_F32)
        sd = np.sqrt(np.maximum(d, _F32(1e-12))).astype(_F32)
        inv = (_F32(1.0) / sd).astype(_F32)
        safe = np.where(nbr < 0, 0, nbr)
        w = ((a * inv[:, None]).astype(_F32) * inv[safe]).astype(_F32)
        w[nbr < 0] = 0
        dev = self._dev
        self._nbr = torch.from_numpy(np.ascontiguousarray(nbr, dtype=np.int32)).to(dev)
        self._A = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        self._W = torch.from_numpy(np.ascontiguousarray(w)).to(dev)
        self._deg = torch.from_numpy((nbr >= 0).sum(axis=1).astype(np.int32)).to(dev)
        self._sd = torch.from_numpy(sd).to(dev)
        self._invalidate_graph_views()

    def _check_ell_width(self, width: int) -> None:
        """A graph adopted from a caller (from_state / the A setter) may have rows far wider than a kNN
        build produces; reject what the SpMM kernels cannot stage, before any device work."""
        lim = int(self._lib.osc_pcg_max_ell_width(int(self.D)))
        if width > lim:
            raise ValueError(f"adjacency row degree {width} exceeds the supported ELL width {lim} for D={self.D}")

    def _load_dense_adjacency(self, dense: np.ndarray) -> None:
        """from_state support (lattice.py:709-713): adopt a user-supplied dense adjacency and
        recompute sqrt_deg / normalised weights the way graph.py:87-90 does."""
        if dense.shape != (self.N, self.N):
            raise ValueError("A shape mismatch")
        mask = dense != 0
        width = max(1, int(mask.sum(axis=1).max()) if self.N else 1)
        self._check_ell_width(width)
        nbr = np.full((self.N, width), -1, dtype=np.int32)
        a = np.zeros((self.N, width), dtype=_F32)
        for i in range(self.N):
            cols = np.nonzero(mask[i])[0]
            nbr[i, : len(cols)] = cols
            a[i, : len(cols)] = dense[i, cols]
        d = dense.sum(axis=1)
        sd = np.sqrt(np.maximum(d, 1e-12)).astype(_F32)
        inv = (1.0 / sd).astype(_F32)
        safe = np.where(nbr < 0, 0, nbr)
        w = ((a * inv[:, None]) * inv[safe]).astype(_F32)
        w[nbr < 0] = 0
        dev = self._dev
        self._nbr = torch.from_numpy(nbr).to(dev)
        self._A = torch.from_numpy(a).to(dev)
        self._W = torch.from_numpy(w).to(dev)
        self._deg = torch.from_numpy((nbr >= 0).sum(axis=1).astype(np.int32)).to(dev)
        self._sd = torch.from_numpy(sd).to(dev)
        self._invalidate_graph_views()

    # ------------------------------------------------------------------ public API
    def set_query(self, psi: np.ndarray, gates: np.ndarray | None = None) -> None:
        # the reference fails with a NumPy broadcasting ValueError at settle time when psi does not have
        # D entries (lattice.py:184); here every kernel reads exactly D floats, so reject it up front
        hpsi = np.asarray(psi).astype(_F32).reshape(-1).copy()
        if hpsi.shape[0] != self.D:
            raise ValueError(f"psi must have D={self.D} entries, got {hpsi.shape[0]}")
        if gates is not None:
            gates = self._check_gates(gates)
        self._hpsi = hpsi
        self._dpsi = torch.from_numpy(self._hpsi).to(self._dev)
        if gates is not None:
            self._hB = gates
            self._dB = torch.from_numpy(self._hB).to(self._dev)
        self._invalidate_cache()

    def _check_gates(self, gates) -> np.ndarray:
        g = np.asarray(gates)
        if g.ndim < 1 or g.shape[0] != self.N:
            raise ValueError("gates length mismatch N")
        if g.ndim != 1:
            raise ValueError("gates must be a 1-D array of length N")
        return g.astype(_F32).copy()

    def set_gates(self, gates: np.ndarray) -> None:
        self._hB = self._check_gates(gates)
        self._dB = torch.from_numpy(self._hB).to(self._dev)
        self._invalidate_cache()

    def add_chain(self, chain: list[int], lamP: float = 0.2, weights: list[float] | None = None) -> None:
        if lamP < 0:
            raise ValueError("lamP must be >= 0")
        if any((c < 0 or c >= self.N) for c in chain):
            raise ValueError("chain indices out of bounds")
        if len(chain) < 2:
            raise ValueError("chain must contain at least two indices")
        if weights is not None and len(weights) != len(chain) - 1:
            raise ValueError("weights length must equal len(chain)-1")
        self._chain = self._make_chain(chain, weights)
        self.lamP = float(lamP)
        self._chain_nodes = list(map(int, chain))
        self._invalidate_cache()
        self._log("add_chain", {"length": len(chain), "lamP": lamP})

    def clear_chain(self) -> None:
        self._chain = None
        self.lamP = 0.0
        self._chain_nodes = None
        self._invalidate_cache()
        self._log("clear_chain", {})

    def _make_chain(self, chain, weights) -> dict[str, Any]:
        """graph.py:101-111 as a CSR over the distinct chain nodes (osc_chain_build, a host function of
        the C ABI): max-merged symmetric path weights, then Wp_uv = (Ap_uv / sdp_u) / sdp_v with
        sdp = sqrt(max(rowsum, 1e-12))."""
        lib = _cabi.load()
        ch = np.ascontiguousarray(np.asarray(chain, dtype=np.int64).astype(np.int32))
        wts = None if weights is None else np.ascontiguousarray(np.asarray(weights, dtype=_F32))
        n_rows, nnz = C.c_int32(0), C.c_int32(0)
        _cabi.check(lib.osc_chain_build_size(ch.ctypes.data, len(ch), self.N, C.byref(n_rows), C.byref(nnz)),
                    "osc_chain_build_size")
        rows = np.empty(n_rows.value, dtype=np.int32)
        rowptr = np.empty(n_rows.value + 1, dtype=np.int32)
        col = np.empty(nnz.value, dtype=np.int32)
        wp = np.empty(nnz.value, dtype=_F32)
        apv = np.empty(nnz.value, dtype=_F32)
        slot = np.empty(self.N, dtype=np.int32)
        _cabi.check(lib.osc_chain_build(ch.ctypes.data, len(ch), None if wts is None else wts.ctypes.data, self.N,
                                        rows.ctypes.data, rowptr.ctypes.data, col.ctypes.data, wp.ctypes.data,
                                        apv.ctypes.data, slot.ctypes.data), "osc_chain_build")
        # host copies for chain_receipt / export_state: {(u, v): value}
        ap, wp_host = {}, {}
        for r, u in enumerate(rows.tolist()):
            for e in range(int(rowptr[r]), int(rowptr[r + 1])):
                ap[(u, int(col[e]))] = _F32(apv[e])
                wp_host[(u, int(col[e]))] = _F32(wp[e])
        dev = self._dev
        as_dev = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
        return {
            "n_rows": int(n_rows.value), "nnz": int(nnz.value), "rows": as_dev(rows),
            "rowptr": as_dev(rowptr), "col": as_dev(col), "Wp": as_dev(wp), "Ap": as_dev(apv),
            "slot": as_dev(slot), "ap_host": ap, "wp_host": wp_host,
        }

    # ------------------------------------------------------------------ solves (K2 / K3)
    def _solve(self, mode: int, dt: float, warm: bool, inertia: float, jacobi: bool, tol: float,
               max_iters: int):
        """One PCG solve on device.  Returns (X tensor, iters, res)."""
        X = torch.empty_like(self._dY)
        if self.N == 0:
            return X, 0, float("nan")
        g, prm = self._graph_struct(), self._params_struct()
        batched_ok = (
            jacobi and self._chain is None and 1 <= max_iters <= 65535
            and (mode == _cabi.MODE_STATIONARY or (warm and float(inertia) <= 0.0))
            and self._lib.osc_batched_supported(self.N, self.D, g.k)
        )
        if batched_ok:
            stats = torch.zeros(5, dtype=torch.float32, device=self._dev)
            unres = stats[4:].view(torch.int32)
            need = C.c_size_t(0)
            _cabi.check(self._lib.osc_batched_workspace(1, self.N, self.D, C.byref(need)))
            ws = self._ws.get(need.value)
            settle = mode == _cabi.MODE_SETTLE
            args = _cabi.BatchedArgs(
                self._dY.data_ptr(), self._dU.data_ptr(), self._dpsi.data_ptr(), self._dB.data_ptr(),
                X.data_ptr() if settle else None, None if settle else X.data_ptr(), stats.data_ptr(),
                None, self.D, 1 if settle else 0, 0 if settle else 1, 0, float(dt), float(tol),
                float(tol), int(max_iters), int(max_iters), unres.data_ptr(),
            )
            _cabi.check(
                self._lib.osc_batched_settle(C.byref(g), C.byref(prm), C.byref(args), ws.data_ptr(),
                                             ws.numel(), _stream_ptr()),
                "osc_batched_settle",
            )
            s = stats.cpu()
            if (int(s[4:].view(torch.int32)[0]) & 1) == 0:
                o = 0 if settle else 2
                return X, int(s[o]), float(s[o + 1])
            # per-slab residuals not monotone around the stop: use the HBM-resident PCG below,
            # which evaluates the lattice-wide test every iteration
        dims = _cabi.PcgDims(self.N, 0, self.N, self.D, 0)
        need = C.c_size_t(0)
        _cabi.check(self._lib.osc_pcg_plan(C.byref(dims), C.byref(need)))
        ws = self._ws.get(need.value)
        ch = self._chain_struct()
        it, res = C.c_int32(0), C.c_float(0.0)
        _cabi.check(
            self._lib.osc_pcg_solve(
                C.byref(g), C.byref(ch) if ch is not None else None, C.byref(prm), mode, float(dt),
                1 if warm else 0, float(inertia), 1 if jacobi else 0, float(tol), int(max_iters),
                self._dY.data_ptr(), self._dU.data_ptr(), self._dpsi.data_ptr(), self._dB.data_ptr(),
                self.D, X.data_ptr(), C.byref(it), C.byref(res), ws.data_ptr(), ws.numel(),
                _stream_ptr(),
            ),
            "osc_pcg_solve",
        )
        return X, int(it.value), float(res.value)

    def settle(
        self,
        dt: float = 1.0,
        max_iters: int = 12,
        tol: float = 1e-3,
        precond: str = "jacobi",
        *,
        warm_start: bool = True,
        inertia: float = 0.0,
    ) -> dict[str, Any]:
        """Implicit Euler step (I + dt M) U+ = U + dt (lamG Y + lamQ B psi^T)  (lattice.py:159-230)."""
        dyn = os.getenv("OSCILLINK_RECEIPT_DYNAMICS", "0").strip().lower() in {"1", "true", "yes"}
        U_prev = self._dU.clone() if dyn else None
        t0 = time.time()
        X, iters, res = self._solve(_cabi.MODE_SETTLE, dt, warm_start, inertia, precond == "jacobi",
                                    tol, max_iters)
        self._dU = X
        self._hU = None
        self.last = {"iters": int(iters), "res": float(res), "t_ms": 1000.0 * (time.time() - t0)}
        self._log("settle", self.last)
        if res > tol * 10:
            self._log("settle_convergence_warn", {"res": float(res), "tol": tol, "iters": int(iters)})
        if dyn:
            try:
                self._last_dynamics = self._compute_dynamics(U_prev, int(iters))
            except Exception:
                self._last_dynamics = None
        for cb in list(self._settle_callbacks):
            try:
                cb(self, self.last)
            except Exception:
                pass
        return self.last

    def _solve_Ustar_device(self, tol: float, max_iters: int, use_cache: bool):
        sig = self._signature()
        if use_cache and self._Ustar_cache is not None and self._Ustar_sig == sig:
            self.stats["ustar_cache_hits"] += 1
            self._log("ustar_cache_hit", {"signature": sig})
            return self._Ustar_cache
        t0 = time.time()
        X, iters, res = self._solve(_cabi.MODE_STATIONARY, 0.0, False, 0.0, True, tol, max_iters)
        solve_ms = 1000.0 * (time.time() - t0)
        converged = bool(res <= tol)
        self.last_ustar = {"iters": int(iters), "res": float(res), "converged": converged,
                           "solve_ms": solve_ms}
        if use_cache:
            self._Ustar_cache = X
            self._Ustar_host = None
            self._Ustar_sig = sig
        self.stats["ustar_solves"] += 1
        self._log("ustar_solve", {"signature": sig, "tol": tol, "max_iters": max_iters,
                                  "iters": int(iters), "res": float(res), "converged": converged,
                                  "solve_ms": solve_ms})
        if not converged:
            self._log("ustar_convergence_warn", {"res": float(res), "tol": tol, "iters": int(iters)})
        return X

    def solve_Ustar(self, tol: float = 1e-4, max_iters: int = 64, use_cache: bool = True) -> np.ndarray:
        """Stationary U* (lattice.py:232-290); returns a host array like the reference."""
        X = self._solve_Ustar_device(tol, max_iters, use_cache)
        if use_cache and X is self._Ustar_cache:
            if self._Ustar_host is None:
                self._Ustar_host = X.cpu().numpy()
            return self._Ustar_host
        return X.cpu().numpy()

    def refresh_Ustar(self, tol: float = 1e-4, max_iters: int = 64) -> np.ndarray:
        self._invalidate_cache()
        self._log("refresh_ustar", {})
        return self.solve_Ustar(tol=tol, max_iters=max_iters, use_cache=True)

    # ------------------------------------------------------------------ receipts (K4)
    def _delta_h(self, Ua, Ub) -> float:
        """deltaH_trace (receipts.py:10-25) for device tensors Ua (state) and Ub (reference)."""
        if self.N == 0:
            return 0.0
        g, prm, ch = self._graph_struct(), self._params_struct(), self._chain_struct()
        dims = _cabi.PcgDims(self.N, 0, self.N, self.D, 0)
        need = C.c_size_t(0)
        _cabi.check(self._lib.osc_pcg_plan(C.byref(dims), C.byref(need)))
        ws = self._ws.get(need.value)
        out = C.c_double(0.0)
        _cabi.check(
            self._lib.osc_delta_h(C.byref(g), C.byref(ch) if ch is not None else None, C.byref(prm),
                                  Ua.data_ptr(), Ub.data_ptr(), self._dB.data_ptr(), self.D,
                                  C.byref(out), ws.data_ptr(), ws.numel(), _stream_ptr()),
            "osc_delta_h",
        )
        return float(np.float32(out.value))

    def _node_terms_device(self, Ustar, z_th: float = 3.0, row_stats: bool = False):
        dev, N = self._dev, self.N
        mu = torch.zeros(N, dtype=torch.float32, device=dev) if row_stats else None
        sg = torch.zeros(N, dtype=torch.float32, device=dev) if row_stats else None
        coh = torch.zeros(N, dtype=torch.float32, device=dev)
        anc = torch.zeros(N, dtype=torch.float32, device=dev)
        qry = torch.zeros(N, dtype=torch.float32, device=dev)
        nj = torch.full((N,), -1, dtype=torch.int32, device=dev)
        nz = torch.zeros(N, dtype=torch.float32, device=dev)
        nr = torch.zeros(N, dtype=torch.float32, device=dev)
        if N:
            g, prm = self._graph_struct(), self._params_struct()
            _cabi.check(
                self._lib.osc_receipt_full(C.byref(g), C.byref(prm), self._dY.data_ptr(),
                                           Ustar.data_ptr(), self._dpsi.data_ptr(), self._dB.data_ptr(),
                                           self.D, float(z_th), coh.data_ptr(), anc.data_ptr(),
                                           qry.data_ptr(), nj.data_ptr(), nz.data_ptr(), nr.data_ptr(),
                                           _cabi.ptr(mu), _cabi.ptr(sg), _stream_ptr()),
                "osc_receipt_full",
            )
        if row_stats:
            return coh, anc, qry, nj, nz, nr, mu, sg
        return coh, anc, qry, nj, nz, nr

    def receipt(self) -> dict[str, Any]:
        """Receipt dict with the reference's keys (lattice.py:298-455)."""
        Ustar = self._solve_Ustar_device(1e-4, 64, True)
        dH = self._delta_h(self._dU, Ustar)
        if self._receipt_detail == "light":
            coh_sum = anc_sum = qry_sum = 0.0
            nulls_full: list[dict[str, Any]] = []
        else:
            coh, anc, qry, nj, nz, nr = self._node_terms_device(Ustar, 3.0)
            coh_sum = float(np.sum(coh.cpu().numpy()))
            anc_sum = float(np.sum(anc.cpu().numpy()))
            qry_sum = float(np.sum(qry.cpu().numpy()))
            hj, hz, hr = nj.cpu().numpy(), nz.cpu().numpy(), nr.cpu().numpy()
            nulls_full = [
                {"edge": [int(i), int(hj[i])], "z": float(hz[i]), "residual": float(hr[i])}
                for i in np.nonzero(hj >= 0)[0]
            ]
        cap_raw = os.getenv("OSCILLINK_RECEIPT_NULL_CAP", "0").strip()
        try:
            cap_val = int(cap_raw)
        except ValueError:
            cap_val = 0
        capped = cap_val > 0 and len(nulls_full) > cap_val
        if capped:
            nulls = sorted(nulls_full, key=lambda e: e.get("z", 0.0), reverse=True)[:cap_val]
        else:
            nulls = nulls_full
        null_meta = {
            "total_null_points": len(nulls_full),
            "returned_null_points": cap_val if capped else len(nulls_full),
            "null_cap_applied": bool(capped),
        }
        lu = getattr(self, "last_ustar", {})
        nnz = self._count_positive_edges()
        sig = self._signature()
        meta: dict[str, Any] = {
            "ustar_cached": bool(self._Ustar_cache is not None and self._Ustar_sig == sig),
            "ustar_solves": int(self.stats["ustar_solves"]),
            "ustar_cache_hits": int(self.stats["ustar_cache_hits"]),
            "ustar_converged": bool(lu.get("converged", True)),
            "ustar_res": float(lu.get("res", 0.0)),
            "ustar_iters": int(lu.get("iters", 0)),
            "ustar_solve_ms": float(lu.get("solve_ms", 0.0)),
            "graph_build_ms": float(getattr(self, "_graph_build_ms", 0.0)),
            "last_settle_ms": float(self.last.get("t_ms") or 0.0),
            "avg_degree": float(nnz / max(self.N, 1)),
            "edge_density": float(nnz / max(self.N * (self.N - 1), 1)),
            "gates_min": float(np.min(self._hB)) if self.N else 0.0,
            "gates_max": float(np.max(self._hB)) if self.N else 0.0,
            "gates_mean": float(np.mean(self._hB)) if self.N else 0.0,
            "gates_uniform": bool(np.allclose(self._hB, self._hB[0])) if self.N else True,
            "state_sig": sig,
            "receipt_detail": self._receipt_detail,
            "null_points_summary": null_meta,
        }
        if self._receipt_secret is not None:
            payload: dict[str, Any] = {"sig_v": 1, "mode": self._signature_mode, "state_sig": sig,
                                       "deltaH_total": float(dH)}
            if self._signature_mode == "extended":
                payload.update({
                    "ustar_iters": int(lu.get("iters", 0)),
                    "ustar_res": float(lu.get("res", 0.0)),
                    "ustar_converged": bool(lu.get("converged", True)),
                    "params": {"lamG": self.lamG, "lamC": self.lamC, "lamQ": self.lamQ,
                               "lamP": self.lamP},
                    "graph": {"k": self._kneighbors, "deterministic_k": self._deterministic_k,
                              "neighbor_seed": self._neighbor_seed},
                })
            raw = json.dumps(payload, sort_keys=True).encode("utf-8")
            meta["signature"] = {
                "algorithm": "HMAC-SHA256",
                "payload": payload,
                "signature": hmac.new(self._receipt_secret, raw, hashlib.sha256).hexdigest(),
            }
        out = {
            "version": REFERENCE_VERSION,
            "deltaH_total": float(dH),
            "coh_drop_sum": coh_sum,
            "anchor_pen_sum": anc_sum,
            "query_term_sum": qry_sum,
            "cg_iters": int(self.last.get("iters") or 0),
            "residual": float(self.last.get("res") or 0.0),
            "t_ms": float(self.last.get("t_ms") or 0.0),
            "null_points": nulls,
            "meta": meta,
        }
        dyn = os.getenv("OSCILLINK_RECEIPT_DYNAMICS", "0").strip().lower() in {"1", "true", "yes"}
        if dyn and self._last_dynamics is not None:
            meta["dynamics"] = self._last_dynamics
        self._log("receipt", {"deltaH_total": out["deltaH_total"], "ustar_cached": meta["ustar_cached"]})
        return out

    # ------------------------------------------------------------------ chain receipt / bundle (f2, f1)
    def _pair_d2(self, V, pairs: np.ndarray) -> np.ndarray:
        """|| V_i/(sd_i+1e-12) - V_j/(sd_j+1e-12) ||^2 for (i,j) rows of `pairs` (device kernel)."""
        M = int(pairs.shape[0])
        if M == 0:
            return np.zeros(0, dtype=_F32)
        dp = torch.from_numpy(np.ascontiguousarray(pairs, dtype=np.int32)).to(self._dev)
        out = torch.empty(M, dtype=torch.float32, device=self._dev)
        _cabi.check(self._lib.osc_pair_d2(V.data_ptr(), self._sd.data_ptr(), dp.data_ptr(), M, self.D,
                                          out.data_ptr(), _stream_ptr()), "osc_pair_d2")
        return out.cpu().numpy()

    def _adjacency_lookup(self, i: int, j: int) -> float:
        nbr = self._nbr[i].cpu().numpy()
        a = self._A[i].cpu().numpy()
        hit = np.nonzero(nbr == j)[0]
        return float(a[hit[0]]) if len(hit) else 0.0

    def chain_receipt(self, chain: list[int], z_th: float = 2.5) -> dict[str, Any]:
        """Per-chain-edge z-scores, weakest link and coherence gain (lattice.py:466-528).

        The D-dimensional work (pair distances, per-row residual statistics) runs on device; the
        O(len(chain)) assembly of the verdict is host bookkeeping."""
        Ustar = self._solve_Ustar_device(1e-4, 64, True)
        *_, mu_s_d, sig_s_d = self._node_terms_device(Ustar, 3.0, row_stats=True)
        mu_s, sig_s = mu_s_d.cpu().numpy(), sig_s_d.cpu().numpy()
        path = self._chain if self._chain is not None else self._make_chain(chain, None)
        ap = path["ap_host"]
        edges_p = np.array(sorted(ap.keys()), dtype=np.int32).reshape(-1, 2)
        d2_p = dict(zip(map(tuple, edges_p.tolist()), self._pair_d2(Ustar, edges_p)))
        cp = _F32(max(self.lamC, 1e-6))
        N = self.N

        def path_stats(i: int):
            vals = np.array([_F32(_F32(cp * ap[(u, v)]) * d2_p[(u, v)]) for (u, v) in ap if u == i],
                            dtype=_F32)
            mu = _F32(vals.sum(dtype=_F32) / _F32(N)) if len(vals) else _F32(0)
            dev = float(((vals.astype(np.float64) - float(mu)) ** 2).sum()) + (N - len(vals)) * float(mu) ** 2
            return mu, _F32(np.sqrt(dev / N)) + _F32(1e-12)

        pairs = np.array([[int(chain[t]), int(chain[t + 1])] for t in range(len(chain) - 1)],
                         dtype=np.int32).reshape(-1, 2)
        d2u = self._pair_d2(Ustar, pairs)
        d2y = self._pair_d2(self._dY, pairs)
        edges: list[dict[str, Any]] = []
        worst = (-1, -1.0, (-1, -1))
        gain = 0.0
        for t, (i, j) in enumerate(pairs.tolist()):
            w_ij = self._adjacency_lookup(i, j)
            rs = _F32(_F32(_F32(self.lamC) * _F32(w_ij)) * d2u[t])
            rp = _F32(_F32(cp * ap.get((i, j), _F32(0))) * d2u[t])
            mu_p, sig_p = path_stats(i)
            z_struct = float((rs - mu_s[i]) / sig_s[i])
            z_path = float((rp - mu_p) / sig_p)
            edges.append({"k": int(t), "edge": [int(i), int(j)], "z_struct": z_struct, "z_path": z_path,
                          "r_struct": float(rs), "r_path": float(rp)})
            if max(z_struct, z_path) > worst[1]:
                worst = (t, max(z_struct, z_path), (i, j))
            gain += 0.5 * float(self.lamC) * max(w_ij, 0.0) * (float(d2y[t]) - float(d2u[t]))
        verdict = all(max(e["z_struct"], e["z_path"]) <= float(z_th) for e in edges)
        return {
            "verdict": bool(verdict),
            "weakest_link": {"k": int(worst[0]), "edge": [int(worst[2][0]), int(worst[2][1])],
                             "zscore": float(worst[1])},
            "coherence_gain": float(gain),
            "edges": edges,
        }

    def bundle(self, k: int = 8, alpha: float = 0.5) -> list[dict]:
        """Top-k diversified bundle: score = alpha*z(coherence drop) + (1-alpha)*cos(U*_i, psi), then
        greedy MMR (lambda 0.5) over the anchors (lattice.py:530-568, graph.py:114-133)."""
        Ustar = self._solve_Ustar_device(1e-4, 64, True)
        N = self.N
        align_d = torch.empty(N, dtype=torch.float32, device=self._dev)
        _cabi.check(self._lib.osc_row_align(Ustar.data_ptr(), self._dpsi.data_ptr(), N, self.D,
                                            align_d.data_ptr(), _stream_ptr()), "osc_row_align")
        coh_d = self._node_terms_device(Ustar, 3.0)[0]
        align, coh = align_d.cpu().numpy(), coh_d.cpu().numpy()
        mu, sigma = float(np.mean(coh)), float(np.std(coh) + 1e-12)
        z = (coh - mu) / sigma if sigma > 0 else np.zeros_like(coh)
        score = (alpha * z + (1 - alpha) * align).astype(_F32)
        if k <= 0 or N == 0:
            return []
        Yn = torch.empty_like(self._dY)
        _cabi.check(self._lib.osc_normalize_rows(self._dY.data_ptr(), N, self.D, Yn.data_ptr(), None, None,
                                                 _stream_ptr()), "osc_normalize_rows")
        score_d = torch.from_numpy(score).to(self._dev)
        steps = min(int(k), N)
        chosen = torch.full((steps,), -1, dtype=torch.int32, device=self._dev)
        need = C.c_size_t(0)
        _cabi.check(self._lib.osc_mmr_workspace(N, C.byref(need)))
        ws = self._ws.get(need.value)
        _cabi.check(self._lib.osc_mmr_select(Yn.data_ptr(), score_d.data_ptr(), N, self.D, steps,
                                             chosen.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr()),
                    "osc_mmr_select")
        order = chosen.cpu().numpy().tolist()
        return [{"id": int(i), "score": float(score[i]), "align": float(align[i])} for i in order]

    def verify_current_receipt(self, secret: bytes | str) -> bool:
        from .receipts import verify_receipt

        return verify_receipt(self.receipt(), secret)

    # ------------------------------------------------------------------ callbacks / logging / config
    def add_settle_callback(self, fn) -> None:
        self._settle_callbacks.append(fn)

    def remove_settle_callback(self, fn) -> None:
        try:
            self._settle_callbacks.remove(fn)
        except ValueError:
            pass

    def set_logger(self, logger_callable) -> None:
        self._logger = logger_callable

    def _log(self, event: str, payload: dict) -> None:
        if self._logger is not None:
            try:
                self._logger(event, payload)
            except Exception:
                pass

    def set_receipt_secret(self, secret: bytes | str | None) -> None:
        if isinstance(secret, str):
            secret = secret.encode("utf-8")
        self._receipt_secret = secret

    def set_signature_mode(self, mode: str) -> None:
        m = mode.lower().strip()
        if m not in {"minimal", "extended"}:
            raise ValueError("mode must be 'minimal' or 'extended'")
        self._signature_mode = m

    def set_receipt_detail(self, mode: str) -> None:
        m = mode.lower().strip()
        if m not in {"full", "light"}:
            raise ValueError("mode must be 'full' or 'light'")
        self._receipt_detail = m

    # ------------------------------------------------------------------ signature / cache
    def _count_positive_edges(self) -> int:
        if self._nnz_pos is None:
            self._nnz_pos = int((self._A > 0).sum().item()) if self.N else 0
        return self._nnz_pos

    def _first_edges(self, limit: int = 2048) -> np.ndarray:
        """argwhere(A > 0)[:limit] (lattice.py:731) without a dense A: ELL rows are already in
        row-major order with ascending columns."""
        if self.N == 0:
            return np.zeros((0, 2), dtype=np.int64)
        rows = min(self.N, limit)  # every row holds >= 0 edges; extend if that is not enough
        while True:
            nbr = self._nbr[:rows].cpu().numpy()
            a = self._A[:rows].cpu().numpy()
            r, t = np.nonzero((nbr >= 0) & (a > 0))
            if len(r) >= limit or rows >= self.N:
                break
            rows = min(self.N, rows * 4)
        pairs = np.stack([r.astype(np.int64), nbr[r, t].astype(np.int64)], axis=1)
        return np.ascontiguousarray(pairs[:limit])

    def _signature(self) -> str:
        if self._adj_sig is None:
            self._adj_sig = hashlib.sha256(self._first_edges().tobytes()).hexdigest()
        data = {
            "psi": np.round(self._hpsi, 6).tolist(),
            "B": np.round(self._hB, 6).tolist(),
            "lam": [self.lamG, self.lamC, self.lamQ, self.lamP],
            "chain_present": self._chain is not None,
            "chain_len": len(self._chain_nodes) if self._chain_nodes else 0,
            "k": self._kneighbors,
            "detk": self._deterministic_k,
            "adj": self._adj_sig,
        }
        return hashlib.sha256(json.dumps(data, sort_keys=True).encode("utf-8")).hexdigest()

    def _invalidate_cache(self) -> None:
        self._Ustar_cache = None
        self._Ustar_host = None
        self._Ustar_sig = None
        self._log("invalidate_cache", {})

    def rebuild_graph(self, *, row_cap_val: float | None = None, kneighbors: int | None = None,
                      deterministic_k: bool | None = None, neighbor_seed: int | None = None) -> None:
        if row_cap_val is not None:
            self._row_cap_val = float(row_cap_val)
        if kneighbors is not None:
            self._kneighbors = min(int(kneighbors), max(1, self.N - 1))
        if deterministic_k is not None:
            self._deterministic_k = bool(deterministic_k)
        if neighbor_seed is not None:
            self._neighbor_seed = neighbor_seed
        self._build_graph()
        self._invalidate_cache()
        self._log("rebuild_graph", {"k": int(self._kneighbors), "row_cap_val": float(self._row_cap_val),
                                    "deterministic_k": self._deterministic_k,
                                    "neighbor_seed": self._neighbor_seed})

    # ------------------------------------------------------------------ export / import
    def export_state(self, include_graph: bool = True, include_chain: bool = True, *,
                     graph_format: str = "dense") -> dict[str, Any]:
        """lattice.py:582-624.  graph_format="dense" (default) writes the reference's N x N `A`;
        graph_format="ell" writes the adjacency the device holds -- `A_ell = {"nbr": [N][k] column ids
        (-1 padded, ascending), "val": [N][k] weights}` -- O(N k) instead of O(N^2); `from_state`
        accepts either."""
        if graph_format not in {"dense", "ell"}:
            raise ValueError("graph_format must be 'dense' or 'ell'")
        h = hashlib.sha256()
        h.update(self.Y.tobytes())
        h.update(self._hpsi.tobytes())
        h.update(self._hB.tobytes())
        h.update(np.array([self.lamG, self.lamC, self.lamQ, self.lamP], dtype=np.float64).tobytes())
        h.update(self._first_edges().tobytes())
        state: dict[str, Any] = {
            "version": REFERENCE_VERSION,
            "shape": [int(self.N), int(self.D)],
            "params": {"lamG": self.lamG, "lamC": self.lamC, "lamQ": self.lamQ, "lamP": self.lamP},
            "Y": self.Y.tolist(),
            "psi": self._hpsi.tolist(),
            "B_diag": self._hB.tolist(),
            "kneighbors": int(self._kneighbors),
            "deterministic_k": bool(self._deterministic_k),
            "neighbor_seed": self._neighbor_seed,
            "provenance": h.hexdigest(),
        }
        if include_graph and graph_format == "dense":
            state["A"] = self.A.tolist()
        elif include_graph:
            nbr, a = self._ell_host()
            state["A_ell"] = {"nbr": nbr.tolist(), "val": a.tolist()}
        if include_chain and self._chain is not None:
            state["chain_edges"] = sorted([int(u), int(v)] for (u, v), w in self._chain["ap_host"].items()
                                          if u < v and w > 0)  # lattice.py:606-611: only A_path > 0
            if self._chain_nodes is not None:
                state["chain_nodes"] = list(self._chain_nodes)
        return state

    def save_state(self, path: str, format: str = "json", include_graph: bool = True,
                   include_chain: bool = True, *, graph_format: str = "dense") -> None:
        fmt = format.lower()
        state = self.export_state(include_graph=include_graph, include_chain=include_chain,
                                  graph_format=graph_format)
        if fmt == "json":
            with open(path, "w", encoding="utf-8") as f:
                json.dump(state, f, sort_keys=True)
        elif fmt == "npz":
            arrays: dict[str, np.ndarray] = {"Y": self.Y, "psi": self._hpsi, "B_diag": self._hB}
            if include_graph and graph_format == "dense":
                arrays["A"] = self.A
            elif include_graph:
                arrays["A_ell_nbr"], arrays["A_ell_val"] = self._ell_host()
            if include_chain and self._chain_nodes is not None:
                arrays["chain_nodes"] = np.array(self._chain_nodes, dtype=np.int32)
            meta = {k: v for k, v in state.items()
                    if k not in {"Y", "psi", "B_diag", "A", "A_ell", "chain_nodes"}}
            np.savez_compressed(path, __meta__=np.array(json.dumps(meta, sort_keys=True)), **arrays)
        else:
            raise ValueError("format must be 'json' or 'npz'")

    @classmethod
    def from_npz(cls, path: str) -> "OscillinkLattice":
        with np.load(path, allow_pickle=False) as data:
            state = dict(json.loads(str(data["__meta__"])))
            state["Y"] = data["Y"].astype(_F32)
            state["psi"] = data["psi"].astype(_F32)
            state["B_diag"] = data["B_diag"].astype(_F32)
            if "A" in data.files:
                state["A"] = data["A"].astype(_F32)
            if "A_ell_nbr" in data.files:
                state["A_ell"] = {"nbr": data["A_ell_nbr"], "val": data["A_ell_val"]}
            if "chain_nodes" in data.files:
                state["chain_nodes"] = data["chain_nodes"].astype(int).tolist()
        return cls.from_state(state)

    @classmethod
    def from_state(cls, state: dict[str, Any]) -> "OscillinkLattice":
        Y = np.array(state["Y"], dtype=_F32)
        params = state.get("params", {})
        lat = cls(Y, kneighbors=state.get("kneighbors", 6), lamG=params.get("lamG", 1.0),
                  lamC=params.get("lamC", 0.5), lamQ=params.get("lamQ", 4.0),
                  deterministic_k=state.get("deterministic_k", False),
                  neighbor_seed=state.get("neighbor_seed"))
        psi = np.array(state.get("psi", np.zeros(Y.shape[1], dtype=_F32)), dtype=_F32)
        B = np.array(state.get("B_diag", np.ones(Y.shape[0], dtype=_F32)), dtype=_F32)
        lat.set_query(psi, gates=B)
        if "A" in state:
            A = np.array(state["A"], dtype=_F32)
            if A.shape == (lat.N, lat.N):
                lat._load_dense_adjacency(A)
        elif "A_ell" in state:
            lat._load_ell_adjacency(np.asarray(state["A_ell"]["nbr"], dtype=np.int32),
                                    np.asarray(state["A_ell"]["val"], dtype=_F32))
        lamP = params.get("lamP", 0.0)
        if lamP > 0:
            if "chain_nodes" in state:
                lat.add_chain(list(map(int, state["chain_nodes"])), lamP=lamP)
            elif state.get("chain_edges"):
                flat = sorted({i for e in state["chain_edges"] for i in e})
                lat.add_chain(flat, lamP=lamP)
        if "provenance" in state:
            lat._imported_provenance = state["provenance"]
        return lat

    # ------------------------------------------------------------------ optional dynamics
    def _compute_dynamics(self, U_prev, iters: int) -> dict[str, Any]:
        """lattice.py:825-903 (env-gated, default off).  Energy step via the deltaH kernel;
        movement statistics and the BFS radius are host bookkeeping on downloaded rows."""
        up, un = U_prev.cpu().numpy(), self.U
        move2 = np.sum((un - up) ** 2, axis=1)
        dH_step = self._delta_h(U_prev, self._dU)
        nbr, a = self._ell_host()
        di = self.sqrt_deg + 1e-12
        Up, Un = up / di[:, None], un / di[:, None]
        flows, total = [], 0.0
        for i, t in zip(*np.nonzero((nbr >= 0) & (a > 0))):
            j = int(nbr[i, t])
            dp, dn = Up[i] - Up[j], Un[i] - Un[j]
            f = max(0.0, 0.5 * self.lamC * float(a[i, t]) * (float(dp @ dp) - float(dn @ dn)))
            if f > 0.0:
                total += f
                flows.append({"edge": [int(i), j], "flow": float(f)})
        flows.sort(key=lambda e: e["flow"], reverse=True)
        inf = np.sqrt(move2 + 1e-12)
        radius = 0
        if inf.size and float(np.max(inf)) > 1e-9:
            seeds = np.where(inf >= 0.1 * float(np.max(inf)))[0].tolist()
            dist = np.full(self.N, -1, dtype=int)
            q: deque[int] = deque()
            for s in seeds:
                dist[s] = 0
                q.append(int(s))
            while q:
                u = q.popleft()
                for v in nbr[u][(nbr[u] >= 0) & (a[u] > 0)]:
                    if dist[v] < 0:
                        dist[v] = dist[u] + 1
                        q.append(int(v))
            radius = int(np.max(dist)) if np.any(dist >= 0) else 0
        return {
            "temperature": float(np.mean(move2)), "step_deltaH": float(dH_step),
            "viscosity_step": float(iters) / (abs(dH_step) + 1e-12), "flow_total": float(total),
            "top_flows": flows[:16], "radius": int(radius),
            "move2_mean": float(np.mean(move2) if move2.size else 0.0),
            "move2_max": float(np.max(move2) if move2.size else 0.0),
        }

    def __repr__(self) -> str:  # pragma: no cover
        parts = [f"N={self.N}", f"D={self.D}", f"k={self._kneighbors}", f"lamG={self.lamG}",
                 f"lamC={self.lamC}", f"lamQ={self.lamQ}"]
        if self.lamP > 0 and self._chain_nodes is not None:
            parts += [f"chain_len={len(self._chain_nodes)}", f"lamP={self.lamP}"]
        if self._Ustar_cache is not None:
            parts.append("U*cached")
        return "OscillinkLattice(" + ", ".join(parts) + ")"


Oscillink = OscillinkLattice


def json_line_logger(stream=None):
    """Logger callable writing one compact JSON object per event (lattice.py:995-1014)."""
    import sys

    out = stream if stream is not None else sys.stderr

    def emit(event: str, payload: dict):  # pragma: no cover
        try:
            out.write(json.dumps({"event": event, **payload}, separators=(",", ":")) + "\n")
        except Exception:
            pass

    return emit
