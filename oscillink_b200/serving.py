"""Request coalescing in front of the batched kernel.

The reference service settles one lattice per HTTP request on a threadpool thread
(cloud/app/main.py:887-947 `_build_lattice`, :1030-1150 `POST /v1/settle`): ctor, set_query,
settle, receipt.  On a B200 one lattice of N ~ 1200 fills a fraction of the machine, so this shim
lets those request threads hand their inputs to ONE worker that groups requests of equal shape and
parameters and runs each group through `BatchedLattices` (one fused kNN pass + one persistent settle
kernel per group).  Per request the result carries what the handler puts in its `ReceiptResponse`
(cloud/app/models.py:36-41): `state_sig`, a light `receipt`, `timings_ms`, and the settle statistics.

    co = SettleCoalescer(max_batch=256, max_wait_ms=2.0)
    fut = co.submit(Y, psi, kneighbors=8)          # from any request thread
    out = fut.result()                             # {"settle": {...}, "receipt": {...}, "state_sig": ...}

Requests the batched kernel does not cover (a chain prior, N > 2560, D % 4 != 0, k > 16) are routed
to a plain `OscillinkLattice`, one by one, on the same worker.  Only host logic lives here; the
backend is injectable so the grouping policy is testable without a GPU.
"""
from __future__ import annotations

import hashlib
import json
import queue
import threading
import time
from concurrent.futures import Future
from dataclasses import dataclass, field
from typing import Any, Callable

import numpy as np

__all__ = ["SettleCoalescer", "SettleRequest", "state_signature"]


@dataclass
class SettleRequest:
    Y: np.ndarray
    psi: np.ndarray
    gates: np.ndarray | None = None
    chain: list[int] | None = None
    kneighbors: int = 6
    row_cap_val: float = 1.0
    lamG: float = 1.0
    lamC: float = 0.5
    lamQ: float = 4.0
    lamP: float = 0.2
    dt: float = 1.0
    max_iters: int = 12
    tol: float = 1e-3
    include_receipt: bool = True
    future: Future = field(default_factory=Future, repr=False)

    def group_key(self):
        """Requests with equal keys can share one BatchedLattices call."""
        n, d = self.Y.shape
        return (n, d, min(int(self.kneighbors), max(1, n - 1)), float(self.row_cap_val), float(self.lamG),
                float(self.lamC), float(self.lamQ), float(self.dt), int(self.max_iters), float(self.tol),
                bool(self.include_receipt), self.chain is None)


def state_signature(psi, gates, lam, k: int, detk: bool, first_edges: np.ndarray, chain_len: int = 0) -> str:
    """lattice.py:729-744: sha256 over the canonical JSON of the rounded query state and the sha256 of
    the first 2048 (i, j) index pairs of {A_ij > 0} in row-major order (int64)."""
    adj = hashlib.sha256(np.ascontiguousarray(first_edges, dtype=np.int64).tobytes()).hexdigest()
    data = {
        "psi": np.round(np.asarray(psi, dtype=np.float32), 6).tolist(),
        "B": np.round(np.asarray(gates, dtype=np.float32), 6).tolist(),
        "lam": [lam[0], lam[1], lam[2], lam[3]],
        "chain_present": chain_len > 0,
        "chain_len": int(chain_len),
        "k": int(k),
        "detk": bool(detk),
        "adj": adj,
    }
    return hashlib.sha256(json.dumps(data, sort_keys=True).encode("utf-8")).hexdigest()


def _validate(r: SettleRequest) -> None:
    """The ValueErrors `_build_lattice` / the lattice constructor raise (lattice.py:45-53,117-125),
    raised on the submitting thread before anything is queued."""
    if not isinstance(r.Y, np.ndarray) or r.Y.ndim != 2:
        raise ValueError("Y must be a 2D numpy array")
    if r.kneighbors < 1:
        raise ValueError("kneighbors must be >= 1")
    if r.lamG <= 0:
        raise ValueError("lamG must be > 0 for SPD")
    if r.lamC < 0:
        raise ValueError("lamC must be >= 0")
    if r.lamQ < 0:
        raise ValueError("lamQ must be >= 0")
    if r.psi.shape[0] != r.Y.shape[1]:
        raise ValueError("psi dimension mismatch")
    if r.gates is not None and r.gates.shape[0] != r.Y.shape[0]:
        raise ValueError("gates length mismatch N")
    if r.kneighbors > 128:
        raise ValueError("kneighbors must be <= 128")
    if r.chain is not None:  # lattice.py:135-142, checked here so that one bad request cannot fail its group
        if r.lamP < 0:
            raise ValueError("lamP must be >= 0")
        if any((c < 0 or c >= r.Y.shape[0]) for c in r.chain):
            raise ValueError("chain indices out of bounds")
        if len(r.chain) < 2:
            raise ValueError("chain must contain at least two indices")


# ----------------------------------------------------------------------------- CUDA backend
def _first_edges_batched(nbr: np.ndarray, a: np.ndarray, limit: int = 2048) -> np.ndarray:
    r, t = np.nonzero((nbr >= 0) & (a > 0))
    pairs = np.stack([r.astype(np.int64), nbr[r, t].astype(np.int64)], axis=1)
    return np.ascontiguousarray(pairs[:limit])


def cuda_backend(group: list[SettleRequest]) -> list[dict[str, Any]]:
    """Run one group (equal group_key) on the device; one result dict per request, in order."""
    import torch

    from .batched_api import BatchedLattices
    from .lattice_api import REFERENCE_VERSION, OscillinkLattice

    r0 = group[0]
    n, d = r0.Y.shape
    t0 = time.time()
    lib_ok = False
    if r0.chain is None and n >= 2:
        from . import _cabi

        lib_ok = bool(_cabi.load().osc_batched_supported(n, d, min(r0.kneighbors, n - 1)))
    if not lib_ok:
        out: list[Any] = []
        for r in group:  # shapes the slab kernel does not cover: the general single-lattice path
            t1 = time.time()
            try:  # requests are independent here: a failing one becomes ITS result, not the group's
                lat = OscillinkLattice(r.Y, kneighbors=r.kneighbors, row_cap_val=r.row_cap_val, lamG=r.lamG,
                                       lamC=r.lamC, lamQ=r.lamQ, deterministic_k=True)
                lat.set_query(r.psi, gates=r.gates)
                if r.chain is not None:
                    lat.add_chain(r.chain, lamP=r.lamP)
                st = lat.settle(dt=r.dt, max_iters=r.max_iters, tol=r.tol)
                rec = None
                if r.include_receipt:
                    lat.set_receipt_detail("light")
                    rec = lat.receipt()
                out.append({"settle": dict(st), "receipt": rec, "state_sig": lat._signature(),
                            "timings_ms": {"total_settle_ms": 1000.0 * (time.time() - t1)},
                            "meta": {"N": n, "D": d, "batch_size": 1, "path": "single"}})
            except Exception as e:  # noqa: BLE001
                out.append(e)
        return out
    B = len(group)
    Y = torch.from_numpy(np.stack([np.ascontiguousarray(r.Y, dtype=np.float32) for r in group])).pin_memory()
    psi = np.stack([np.asarray(r.psi, dtype=np.float32) for r in group])
    any_gates = any(r.gates is not None for r in group)
    gates = (np.stack([np.ones(n, np.float32) if r.gates is None else np.asarray(r.gates, np.float32)
                       for r in group]) if any_gates else None)
    bl = BatchedLattices(Y.cuda(non_blocking=True), kneighbors=r0.kneighbors, row_cap_val=r0.row_cap_val,
                         lamG=r0.lamG, lamC=r0.lamC, lamQ=r0.lamQ)
    bl.set_query(psi, gates)
    res = bl.settle(dt=r0.dt, max_iters=r0.max_iters, tol=r0.tol, receipt=r0.include_receipt)
    iters = res["iters"].cpu().numpy()
    resid = res["res"].cpu().numpy()
    rows = min(n, 2048)
    nbr_h = bl.nbr[:, :rows].cpu().numpy()
    a_h = bl.A[:, :rows].cpu().numpy()
    nnz = bl.nnz.cpu().numpy()
    if r0.include_receipt:
        uit, ures, dh = (res[k].cpu().numpy() for k in ("ustar_iters", "ustar_res", "deltaH"))
    ms = 1000.0 * (time.time() - t0)
    out = []
    for b, r in enumerate(group):
        edges = _first_edges_batched(nbr_h[b], a_h[b])
        if len(edges) < 2048 and rows < n:  # sparse head of the table: take every row
            edges = _first_edges_batched(bl.nbr[b].cpu().numpy(), bl.A[b].cpu().numpy())
        g = np.ones(n, np.float32) if r.gates is None else np.asarray(r.gates, np.float32)
        sig = state_signature(psi[b], g, [r.lamG, r.lamC, r.lamQ, 0.0], bl.k, True, edges)
        st = {"iters": int(iters[b]), "res": float(resid[b]), "t_ms": ms / B}
        rec = None
        if r.include_receipt:
            rec = {
                "version": REFERENCE_VERSION, "deltaH_total": float(np.float32(dh[b])), "coh_drop_sum": 0.0,
                "anchor_pen_sum": 0.0, "query_term_sum": 0.0, "cg_iters": st["iters"], "residual": st["res"],
                "t_ms": st["t_ms"], "null_points": [],
                "meta": {"ustar_iters": int(uit[b]), "ustar_res": float(ures[b]),
                         "ustar_converged": bool(ures[b] <= 1e-4), "avg_degree": float(nnz[b] / max(n, 1)),
                         "edge_density": float(nnz[b] / max(n * (n - 1), 1)), "gates_min": float(g.min()),
                         "gates_max": float(g.max()), "gates_mean": float(g.mean()),
                         "gates_uniform": bool(np.allclose(g, g[0])), "state_sig": sig,
                         "receipt_detail": "light"},
            }
        out.append({"settle": st, "receipt": rec, "state_sig": sig,
                    "timings_ms": {"total_settle_ms": ms / B, "batch_ms": ms},
                    "meta": {"N": n, "D": d, "batch_size": B, "path": "batched"}})
    return out


# ----------------------------------------------------------------------------- the coalescer
class SettleCoalescer:
    """One worker thread; request threads `submit()` and wait on the returned Future.

    Policy: the worker takes the oldest request, then keeps draining the queue until either
    `max_batch` requests with the SAME group key are collected or `max_wait_ms` has passed since the
    first one arrived; requests with other keys seen meanwhile are kept and served next (FIFO per key).
    """

    def __init__(self, max_batch: int = 256, max_wait_ms: float = 2.0,
                 backend: Callable[[list[SettleRequest]], list[dict[str, Any]]] | None = None):
        if max_batch < 1:
            raise ValueError("max_batch must be >= 1")
        self.max_batch, self.max_wait = int(max_batch), float(max_wait_ms) / 1000.0
        self._backend = backend or cuda_backend
        self._q: "queue.Queue[SettleRequest | None]" = queue.Queue()
        self._held: list[SettleRequest] = []
        self.stats = {"requests": 0, "batches": 0, "max_batch_seen": 0}
        self._closed = False
        self._worker = threading.Thread(target=self._run, name="osc-coalescer", daemon=True)
        self._worker.start()

    # ---- request side
    def submit(self, Y, psi=None, gates=None, chain=None, **kw) -> Future:
        if self._closed:
            raise RuntimeError("coalescer is closed")
        Y = np.asarray(Y, dtype=np.float32) if not isinstance(Y, np.ndarray) else Y
        d = Y.shape[1] if Y.ndim == 2 else 0
        psi = np.zeros(d, np.float32) if psi is None else np.asarray(psi, dtype=np.float32)
        req = SettleRequest(Y=Y, psi=psi, gates=None if gates is None else np.asarray(gates, np.float32),
                            chain=None if chain is None else [int(c) for c in chain], **kw)
        _validate(req)
        self._q.put(req)
        return req.future

    def settle(self, Y, psi=None, **kw) -> dict[str, Any]:
        return self.submit(Y, psi, **kw).result()

    def close(self) -> None:
        self._closed = True
        self._q.put(None)
        self._worker.join(timeout=30)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- worker side
    def _next_group(self) -> list[SettleRequest] | None:
        first = self._held.pop(0) if self._held else self._q.get()
        if first is None:
            return None
        key, group = first.group_key(), [first]
        rest = []
        for r in self._held:  # earlier arrivals with the same key ride along
            (group if (r.group_key() == key and len(group) < self.max_batch) else rest).append(r)
        self._held = rest
        deadline = time.monotonic() + self.max_wait
        while len(group) < self.max_batch:
            left = deadline - time.monotonic()
            try:
                r = self._q.get(timeout=max(left, 0.0)) if left > 0 else self._q.get_nowait()
            except queue.Empty:
                break
            if r is None:
                self._q.put(None)  # leave the sentinel for the main loop
                break
            (group if r.group_key() == key else self._held).append(r)
        return group

    def _run(self) -> None:
        while True:
            group = self._next_group()
            if group is None:
                for r in self._held:
                    r.future.set_exception(RuntimeError("coalescer closed"))
                return
            self.stats["requests"] += len(group)
            self.stats["batches"] += 1
            self.stats["max_batch_seen"] = max(self.stats["max_batch_seen"], len(group))
            try:
                results = self._backend(group)
                for r, out in zip(group, results):
                    # a backend may return an Exception in a request's slot (per-request failure)
                    if isinstance(out, BaseException):
                        r.future.set_exception(out)
                    else:
                        r.future.set_result(out)
            except BaseException as e:  # noqa: BLE001 -- the request threads must always be released
                for r in group:
                    if not r.future.done():
                        r.future.set_exception(e)
