"""Serving-batch API: many independent lattices of equal shape on one B200.

BASELINE.json config "serving batch: 4096 independent lattices N=1200 D=384 settled
concurrently".  Semantically each lattice b is exactly

    lat = OscillinkLattice(Y[b], kneighbors=k, deterministic_k=True, ...)
    lat.set_query(psi[b], gates[b]); lat.settle(...); lat.set_receipt_detail("light"); lat.receipt()

(the cloud handler's per-request sequence, cloud/app/main.py:916-939,1043,1061) but the whole
batch is built by one fused kNN pass and settled by ONE persistent kernel
(osc_batched_settle, csrc/batched.cu).
"""
from __future__ import annotations

import ctypes as C
from typing import Any

import numpy as np
import torch

from . import _cabi
from ._cabi import BatchedArgs, Graph, Params

__all__ = ["BatchedLattices", "settle_host_batch"]


class BatchedLattices:
    def __init__(self, Y, kneighbors: int = 6, row_cap_val: float = 1.0, lamG: float = 1.0,
                 lamC: float = 0.5, lamQ: float = 4.0, *, knn_engine: int = _cabi.KNN_AUTO,
                 device: torch.device | None = None):
        if not torch.cuda.is_available():
            raise RuntimeError("oscillink_b200 needs a CUDA device (sm_100a); no CPU fallback")
        if kneighbors < 1:
            raise ValueError("kneighbors must be >= 1")
        if lamG <= 0:
            raise ValueError("lamG must be > 0 for SPD")
        if lamC < 0 or lamQ < 0:
            raise ValueError("lamC and lamQ must be >= 0")
        self._dev = device or torch.device("cuda", torch.cuda.current_device())
        self._lib = _cabi.load()
        if isinstance(Y, np.ndarray):
            if Y.ndim != 3:
                raise ValueError("Y must be a (batch, N, D) array")
            Y = torch.from_numpy(np.ascontiguousarray(Y, dtype=np.float32)).to(self._dev, non_blocking=True)
        if Y.ndim != 3 or Y.dtype != torch.float32 or not Y.is_cuda:
            raise ValueError("Y must be a (batch, N, D) float32 array / CUDA tensor")
        self.Y = Y.contiguous()
        self.B, self.N, self.D = (int(s) for s in self.Y.shape)
        self.k = min(int(kneighbors), max(1, self.N - 1))
        self.row_cap_val = float(row_cap_val)
        self.lamG, self.lamC, self.lamQ = float(lamG), float(lamC), float(lamQ)
        self._engine = knn_engine
        self._ws = None
        self.psi = torch.zeros((self.B, self.D), dtype=torch.float32, device=self._dev)
        self.gates = None
        self.U = None
        self._U_prev = None
        self.Ustar = None
        self._build()

    def _workspace(self, nbytes: int):
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=self._dev)
        return self._ws

    def _stream(self) -> int:
        return torch.cuda.current_stream().cuda_stream

    def _build(self) -> None:
        """graph.py:29-93 for the whole batch, composed from the exported K1/K1b phases so each
        phase can be timed with CUDA events on the launching stream (bench.py roofline)."""
        B, N, D, k, dev, lib = self.B, self.N, self.D, self.k, self._dev, self._lib
        self.nbr = torch.empty((B, N, k), dtype=torch.int32, device=dev)
        self.A = torch.empty((B, N, k), dtype=torch.float32, device=dev)
        self.W = torch.empty((B, N, k), dtype=torch.float32, device=dev)
        self.deg = torch.empty((B, N), dtype=torch.int32, device=dev)
        self.sqrt_deg = torch.empty((B, N), dtype=torch.float32, device=dev)
        self.nnz = torch.zeros(B, dtype=torch.int64, device=dev)
        self.gap = torch.empty((B, N), dtype=torch.float32, device=dev)
        self.events = {}
        if N < 2:
            need = C.c_size_t(0)
            _cabi.check(lib.osc_knn_build_workspace(B, N, D, k, self._engine, C.byref(need)))
            ws = self._workspace(need.value)
            _cabi.check(lib.osc_knn_build(self.Y.data_ptr(), B, N, D, k, self.row_cap_val, self._engine,
                                          self.nbr.data_ptr(), self.A.data_ptr(), self.W.data_ptr(),
                                          self.deg.data_ptr(), self.sqrt_deg.data_ptr(),
                                          self.nnz.data_ptr(), self.gap.data_ptr(), ws.data_ptr(),
                                          ws.numel(), self._stream()), "osc_knn_build")
            return
        rows = B * N
        eng, kc, eps = _cabi.knn_plan(N, N, D, k, self._engine)
        use_tc = eng in (_cabi.KNN_TC, _cabi.KNN_TC1)
        self.engine_used = _cabi.ENGINE_NAMES[eng]
        self.kc = kc
        Yn = torch.empty_like(self.Y)
        hi = torch.empty_like(self.Y) if use_tc else None
        lo = torch.empty_like(self.Y) if eng == _cabi.KNN_TC else None
        cand_idx = torch.empty((B, N, kc), dtype=torch.int32, device=dev)
        cand_sim = torch.empty((B, N, kc), dtype=torch.float32, device=dev)
        top_idx = torch.empty((B, N, k), dtype=torch.int32, device=dev)
        top_sim = torch.empty((B, N, k), dtype=torch.float32, device=dev)
        scratch = torch.empty(rows, dtype=torch.float32, device=dev)
        st = self._stream()

        def phase(name, rc_fn):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _cabi.check(rc_fn(), name)
            e1.record()
            self.events[name] = (e0, e1)

        P = _cabi.ptr
        if eng == _cabi.KNN_TCH:  # fp16 rows ride in the q_hi / all_hi arguments
            hi = torch.empty(self.Y.shape, dtype=torch.float16, device=dev)
            phase("normalize", lambda: lib.osc_normalize_rows_f16(self.Y.data_ptr(), rows, D, Yn.data_ptr(),
                                                                  hi.data_ptr(), st))
        else:
            phase("normalize", lambda: lib.osc_normalize_rows(self.Y.data_ptr(), rows, D, Yn.data_ptr(),
                                                              P(hi), P(lo), st))
        phase("knn_candidates", lambda: lib.osc_knn_candidates(
            Yn.data_ptr(), Yn.data_ptr(), P(hi), P(lo), P(hi), P(lo), B, N, 0, N, D, kc,
            eng, cand_idx.data_ptr(), cand_sim.data_ptr(), None, 0, st))
        # canonical re-scoring + completeness check of every candidate list (rows that cannot be proven
        # complete are recomputed exhaustively on device; n_exhaustive counts them)
        self.n_exhaustive = torch.zeros(1, dtype=torch.int32, device=dev)
        need = C.c_size_t(0)
        _cabi.check(lib.osc_knn_rescore_workspace(B, N, C.byref(need)))
        rws = self._workspace(need.value)
        # single-product engines: the exhaustive path is bounded on the device (no host sync here); if the
        # bound was exceeded (clustered / near-duplicate anchors) `_verify_build` rebuilds with 3xTF32
        single = eng in (_cabi.KNN_TC1, _cabi.KNN_TCH)
        self._exh_limit = int(lib.osc_knn_exhaustive_limit(rows)) if single else -1
        self._build_verified = not single
        phase("knn_rescore", lambda: lib.osc_knn_rescore_guarded(
            Yn.data_ptr(), Yn.data_ptr(), B, N, 0, N, D, cand_idx.data_ptr(), cand_sim.data_ptr(), kc, k,
            eps, self._exh_limit, top_idx.data_ptr(), top_sim.data_ptr(), self.gap.data_ptr(),
            self.n_exhaustive.data_ptr(), rws.data_ptr(), rws.numel(), st))
        phase("graph_assemble", lambda: lib.osc_graph_assemble(
            top_idx.data_ptr(), top_sim.data_ptr(), B, N, k, self.row_cap_val, self.nbr.data_ptr(),
            self.A.data_ptr(), self.W.data_ptr(), self.deg.data_ptr(), self.sqrt_deg.data_ptr(),
            self.nnz.data_ptr(), scratch.data_ptr(), st))

    def build_exceeded(self) -> bool:
        """True if the single-product kNN engine flagged more rows than the exhaustive path is allowed to
        take (synchronises on first call): the graph then came from unproven candidate lists and must be
        rebuilt with the 3xTF32 engine -- `_verify_build` does it."""
        if getattr(self, "_build_verified", True):
            return False
        self._build_verified = True
        return int(self.n_exhaustive.item()) > self._exh_limit >= 0

    def _verify_build(self) -> bool:
        if not self.build_exceeded():
            return False
        was = getattr(self, "engine_used", "?")
        self._engine = _cabi.KNN_TC
        self._ws = None
        self._build()
        self.engine_used = f"{self.engine_used} (fallback from {was})"
        return True

    def phase_ms(self) -> dict:
        """Device time of every recorded phase (synchronises)."""
        torch.cuda.current_stream().synchronize()
        return {k: e0.elapsed_time(e1) for k, (e0, e1) in self.events.items()}

    def set_query(self, psi, gates=None) -> None:
        psi_t = torch.as_tensor(psi, dtype=torch.float32)
        if psi_t.ndim == 1:
            psi_t = psi_t.expand(self.B, self.D)
        if tuple(psi_t.shape) != (self.B, self.D):
            raise ValueError("psi must be (D,) or (batch, D)")
        self.psi = psi_t.to(self._dev, non_blocking=True).contiguous()
        if gates is not None:
            g = torch.as_tensor(gates, dtype=torch.float32)
            if tuple(g.shape) != (self.B, self.N):
                raise ValueError("gates length mismatch N")
            self.gates = g.to(self._dev, non_blocking=True).contiguous()

    def supported(self) -> bool:
        return bool(self._lib.osc_batched_supported(self.N, self.D, self.k))

    def settle(self, dt: float = 1.0, max_iters: int = 12, tol: float = 1e-3, *, receipt: bool = True,
               ustar_tol: float = 1e-4, ustar_max_iters: int = 64, keep_ustar: bool = False,
               strict: bool = True) -> dict[str, Any]:
        """settle() [+ light receipt()] for every lattice, one persistent kernel launch.

        Returns device tensors: iters[B], res[B] and, with receipt=True, ustar_iters[B],
        ustar_res[B], deltaH[B] (float64), plus unresolved[B] (int32 flags: bit 0 = needs the
        fallback below, bit 1 = some slab was re-run by the kernel's own fix pass).

        The slab kernel resolves the lattice-wide stop test after the fact; a lattice whose
        per-slab residuals are not monotone around the stop is flagged in `unresolved`.  With
        strict=True (default) the call synchronises, and flagged lattices are settled again
        with the HBM-resident PCG (osc_pcg_solve) so every returned row is final; strict=False
        returns without synchronising and leaves that to `resolve_flagged(out)`."""
        if not self.supported():
            raise _cabi.OscillinkNativeError(
                "batched kernel does not cover this shape; use OscillinkLattice per lattice")
        if strict:
            self._verify_build()  # (strict=False callers get the same check from resolve_flagged)
        B, dev = self.B, self._dev
        U_in = self.U
        U_out = torch.empty_like(self.Y)
        stats = torch.zeros((B, 4), dtype=torch.float32, device=dev)
        dH = torch.zeros(B, dtype=torch.float64, device=dev)
        unres = torch.zeros(B, dtype=torch.int32, device=dev)
        Us = torch.empty_like(self.Y) if (receipt and keep_ustar) else None
        g = Graph(B, self.N, self.k, 0, self.nbr.data_ptr(), self.A.data_ptr(), self.W.data_ptr(),
                  self.deg.data_ptr(), self.sqrt_deg.data_ptr())
        prm = Params(self.lamG, self.lamC, self.lamQ, 0.0, 0, 0)
        args = BatchedArgs(self.Y.data_ptr(), _cabi.ptr(U_in), self.psi.data_ptr(), _cabi.ptr(self.gates),
                           U_out.data_ptr(), _cabi.ptr(Us), stats.data_ptr(), dH.data_ptr(), self.D,
                           1, 1 if receipt else 0, 1 if receipt else 0, float(dt), float(tol),
                           float(ustar_tol), int(max_iters), int(ustar_max_iters), unres.data_ptr())
        need = C.c_size_t(0)
        _cabi.check(self._lib.osc_batched_workspace(B, self.N, self.D, C.byref(need)))
        ws = self._workspace(need.value)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _cabi.check(
            self._lib.osc_batched_settle(C.byref(g), C.byref(prm), C.byref(args), ws.data_ptr(),
                                         ws.numel(), self._stream()),
            "osc_batched_settle",
        )
        e1.record()
        self.events["batched_settle"] = (e0, e1)
        self._U_prev = U_in
        self.U = U_out
        self.Ustar = Us
        out = {"iters": stats[:, 0], "res": stats[:, 1], "unresolved": unres, "_stats": stats}
        if receipt:
            out.update({"ustar_iters": stats[:, 2], "ustar_res": stats[:, 3], "deltaH": dH})
        self._last_call = dict(dt=float(dt), max_iters=int(max_iters), tol=float(tol), receipt=bool(receipt),
                               ustar_tol=float(ustar_tol), ustar_max_iters=int(ustar_max_iters))
        if strict:
            self.resolve_flagged(out)
        return out

    def resolve_flagged(self, out: dict[str, Any]) -> int:
        """Synchronise and settle every lattice flagged `unresolved` with the HBM-resident PCG
        (csrc/pcg.cu: global stop test every iteration).  Returns how many were redone."""
        if self._verify_build():  # the graph itself had to be rebuilt: settle the whole batch again
            a = self._last_call
            self.U = self._U_prev
            new = self.settle(dt=a["dt"], max_iters=a["max_iters"], tol=a["tol"], receipt=a["receipt"],
                              ustar_tol=a["ustar_tol"], ustar_max_iters=a["ustar_max_iters"],
                              keep_ustar=self.Ustar is not None, strict=True)
            out.update(new)
            return self.B
        flagged = torch.nonzero(out["unresolved"] & 1).flatten().tolist()
        if not flagged:
            return 0
        a = self._last_call
        lib, D, N = self._lib, self.D, self.N
        prm = Params(self.lamG, self.lamC, self.lamQ, 0.0, 0, 0)
        dims = _cabi.PcgDims(N, 0, N, D, 0)
        need = C.c_size_t(0)
        _cabi.check(lib.osc_pcg_plan(C.byref(dims), C.byref(need)))
        ws = torch.empty(max(need.value, 256), dtype=torch.uint8, device=self._dev)
        ones = torch.ones(N, dtype=torch.float32, device=self._dev)
        stats = out["_stats"]
        for b in flagged:
            g = Graph(1, N, self.k, 0, self.nbr[b].data_ptr(), self.A[b].data_ptr(), self.W[b].data_ptr(),
                      self.deg[b].data_ptr(), self.sqrt_deg[b].data_ptr())
            gates = self.gates[b] if self.gates is not None else ones
            U0 = self._U_prev[b] if self._U_prev is not None else self.Y[b]
            it, res = C.c_int32(0), C.c_float(0.0)
            X = torch.empty_like(self.Y[b])
            _cabi.check(lib.osc_pcg_solve(C.byref(g), None, C.byref(prm), _cabi.MODE_SETTLE, a["dt"], 1, 0.0,
                                          1, a["tol"], a["max_iters"], self.Y[b].data_ptr(), U0.data_ptr(),
                                          self.psi[b].data_ptr(), gates.data_ptr(), D, X.data_ptr(),
                                          C.byref(it), C.byref(res), ws.data_ptr(), ws.numel(),
                                          self._stream()), "osc_pcg_solve")
            self.U[b].copy_(X)
            stats[b, 0], stats[b, 1] = float(it.value), float(res.value)
            if a["receipt"]:
                Xs = torch.empty_like(self.Y[b])
                _cabi.check(lib.osc_pcg_solve(C.byref(g), None, C.byref(prm), _cabi.MODE_STATIONARY, 0.0, 0,
                                              0.0, 1, a["ustar_tol"], a["ustar_max_iters"],
                                              self.Y[b].data_ptr(), self.Y[b].data_ptr(),
                                              self.psi[b].data_ptr(), gates.data_ptr(), D, Xs.data_ptr(),
                                              C.byref(it), C.byref(res), ws.data_ptr(), ws.numel(),
                                              self._stream()), "osc_pcg_solve")
                stats[b, 2], stats[b, 3] = float(it.value), float(res.value)
                if self.Ustar is not None:
                    self.Ustar[b].copy_(Xs)
                dh = C.c_double(0.0)
                _cabi.check(lib.osc_delta_h(C.byref(g), None, C.byref(prm), X.data_ptr(), Xs.data_ptr(),
                                            gates.data_ptr(), D, C.byref(dh), ws.data_ptr(), ws.numel(),
                                            self._stream()), "osc_delta_h")
                out["deltaH"][b] = dh.value
            out["unresolved"][b] &= 2
        return len(flagged)


def settle_host_batch(Y_host: torch.Tensor, psi_host: torch.Tensor, kneighbors: int = 6, *,
                      chunk: int = 128, max_iters: int = 12, tol: float = 1e-3, receipt: bool = True,
                      out_host: torch.Tensor | None = None, U_host: torch.Tensor | None = None,
                      device: torch.device | None = None, **lattice_kw) -> torch.Tensor:
    """End-to-end serving call on HOST buffers: for every lattice b of Y_host[B,N,D] (pinned fp32)
    run ctor + set_query(psi_host[b]) + settle + light receipt (cloud/app/main.py:916-939,1043,1061)
    and return a pinned [B,5] float64 host tensor {iters, res, ustar_iters, ustar_res, deltaH}.
    U_host (optional, pinned [B,N,D] fp32) also receives the settled state of every lattice: its D2H copy
    runs on a third stream, concurrently with the next chunk's H2D copy (PCIe is full duplex) and compute.

    The batch is cut into chunks; the H2D copy of chunk i+1 runs on a copy stream while chunk i is
    built and settled, so PCIe and the SMs work concurrently (two device staging buffers)."""
    dev = device or torch.device("cuda", torch.cuda.current_device())
    B, N, D = (int(v) for v in Y_host.shape)
    chunk = max(1, min(int(chunk), B))
    if out_host is None:
        out_host = torch.empty((B, 5), dtype=torch.float64, pin_memory=True)
    res_dev = torch.empty((B, 5), dtype=torch.float64, device=dev)
    comp = torch.cuda.current_stream(dev)
    copy = torch.cuda.Stream(dev)
    back = torch.cuda.Stream(dev) if U_host is not None else None
    bufY = [torch.empty((chunk, N, D), dtype=torch.float32, device=dev) for _ in range(2)]
    bufP = [torch.empty((chunk, D), dtype=torch.float32, device=dev) for _ in range(2)]
    copy.wait_stream(comp)
    free_ev = [None, None]
    pending = []
    n_chunks = (B + chunk - 1) // chunk
    for i in range(n_chunks):
        lo, hi = i * chunk, min(B, (i + 1) * chunk)
        s = i & 1
        with torch.cuda.stream(copy):
            if free_ev[s] is not None:
                copy.wait_event(free_ev[s])
            bufY[s][: hi - lo].copy_(Y_host[lo:hi], non_blocking=True)
            bufP[s][: hi - lo].copy_(psi_host[lo:hi], non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(copy)
        comp.wait_event(ready)
        bl = BatchedLattices(bufY[s][: hi - lo], kneighbors=kneighbors, device=dev, **lattice_kw)
        bl.set_query(bufP[s][: hi - lo])
        out = bl.settle(max_iters=max_iters, tol=tol, receipt=receipt, strict=False)
        r = res_dev[lo:hi]
        r[:, 0], r[:, 1] = out["iters"], out["res"]
        if receipt:
            r[:, 2], r[:, 3], r[:, 4] = out["ustar_iters"], out["ustar_res"], out["deltaH"]
        free_ev[s] = torch.cuda.Event()
        free_ev[s].record(comp)
        if back is not None:
            back.wait_event(free_ev[s])
            with torch.cuda.stream(back):
                U_host[lo:hi].copy_(bl.U, non_blocking=True)
        pending.append((bl, out, lo, hi))
    out_host.copy_(res_dev, non_blocking=True)
    comp.synchronize()
    if back is not None:
        back.synchronize()
    # flagged lattices (pathological, see BatchedLattices.settle): redo with the global-test PCG.
    # bufY has been reused by then, so the chunk is fetched again from the host copy.
    for bl, out, lo, hi in pending:
        if bl.build_exceeded() or bool((out["unresolved"] & 1).any()):
            bl._build_verified = False  # (let resolve_flagged see the exceeded build again)
            bl.Y = Y_host[lo:hi].to(dev)
            bl.psi = psi_host[lo:hi].to(dev)
            bl.resolve_flagged(out)
            r = res_dev[lo:hi]
            r[:, 0], r[:, 1] = out["iters"], out["res"]
            if receipt:
                r[:, 2], r[:, 3], r[:, 4] = out["ustar_iters"], out["ustar_res"], out["deltaH"]
            out_host[lo:hi].copy_(r)
            if U_host is not None:
                U_host[lo:hi].copy_(bl.U)
    return out_host
