"""Screened-diffusion gates on device.

Mirror of `oscillink.preprocess.diffusion.compute_diffusion_gates`
(oscillink/preprocess/diffusion.py:35-129): same signature, same ValueErrors, same output.

    (L_sym + gamma I) h = beta * max(0, cos(Y_i, psi))          diffusion.py:113-123
    h <- (h - min) / (max - min), clipped to [0, 1]              diffusion.py:125-129

The structural graph is the lattice's own mutual-kNN build (K1/K1b kernels).  The SPD system is the
stationary lattice operator with (lamG, lamC, lamQ) = (gamma, 1, 0): M = (gamma + 1) I - W =
L_sym + gamma I, so the solve is `osc_pcg_solve_system` (K2) with one right-hand side.

* method="cg"     (diffusion.py:138-150): the reference runs its Jacobi-PCG with M_diag = diag(L_sym)
                  + gamma = 1 + gamma, a scalar multiple of I, from x0 = 0 with an absolute tolerance.
                  A scalar preconditioner leaves every CG iterate unchanged, so the device solve uses
                  the same recurrences, the same x0, tol and max_iters.
* method="direct" (diffusion.py:152-163): `np.linalg.solve`.  There is no dense factorisation here
                  (N x N never exists); the same system is iterated to the fp32 floor
                  (||r|| <= 2e-7 ||s||, cap 512 iterations), which agrees with LAPACK's fp32
                  solution to ~1e-6 (condition number <= (2 + gamma) / gamma).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

__all__ = ["compute_diffusion_gates"]


def compute_diffusion_gates(
    Y: np.ndarray,
    psi: np.ndarray,
    *,
    kneighbors: int = 6,
    row_cap_val: float = 1.0,
    beta: float = 1.0,
    gamma: float = 0.1,
    similarity: str = "cosine",
    deterministic_k: bool = False,
    neighbor_seed: Optional[int] = None,
    clamp: bool = True,
    method: str = "direct",
    tol: float = 1e-4,
    max_iters: int = 256,
) -> np.ndarray:
    # validation in the reference's order and wording (diffusion.py:86-94)
    if Y.ndim != 2:
        raise ValueError("Y must be 2D")
    N, D = Y.shape
    if psi.shape[0] != D:
        raise ValueError("psi dimension mismatch")
    if gamma <= 0:
        raise ValueError("gamma must be > 0 for SPD")
    if kneighbors < 1:
        raise ValueError("kneighbors must be >=1")
    if similarity != "cosine":
        raise ValueError("unsupported similarity metric")

    import torch

    from . import _cabi
    from .lattice_api import OscillinkLattice, _stream_ptr

    # 1. structural adjacency with the lattice's own build (diffusion.py:100-108)
    lat = OscillinkLattice(np.asarray(Y, dtype=np.float32), kneighbors=kneighbors, row_cap_val=row_cap_val,
                           lamG=float(gamma), lamC=1.0, lamQ=0.0, deterministic_k=deterministic_k,
                           neighbor_seed=neighbor_seed)
    if N == 0:
        return np.zeros(0, dtype=np.float32)
    lib, dev = lat._lib, lat._dev
    st = _stream_ptr()
    # 2. source strengths s = beta * max(0, <Y_i/(|Y_i|+1e-12), psi/(|psi|+1e-12)>)  (diffusion.py:111-118)
    dpsi = torch.from_numpy(np.ascontiguousarray(psi, dtype=np.float32)).to(dev)
    s = torch.empty(N, dtype=torch.float32, device=dev)
    _cabi.check(lib.osc_row_align(lat._dY.data_ptr(), dpsi.data_ptr(), N, D, s.data_ptr(), st), "osc_row_align")
    s = (float(beta) * torch.clamp_min(s, 0.0)).to(torch.float32).contiguous()
    # 3. (L_sym + gamma I) h = s
    if method == "cg":
        solve_tol, solve_max = float(tol), int(max_iters)
    else:
        solve_tol, solve_max = 2e-7 * float(s.norm().item()), 512
    h = torch.zeros(N, dtype=torch.float32, device=dev)  # x0 = 0 (solver.py:17-18)
    rhs = s.clone()
    g, prm = lat._graph_struct(), lat._params_struct()
    dims = _cabi.PcgDims(N, 0, N, 1, 0)
    need = C.c_size_t(0)
    _cabi.check(lib.osc_pcg_plan(C.byref(dims), C.byref(need)))
    ws = lat._ws.get(need.value)
    it, res = C.c_int32(0), C.c_float(0.0)
    if solve_max >= 1:
        _cabi.check(
            lib.osc_pcg_solve_system(C.byref(g), None, C.byref(prm), _cabi.MODE_STATIONARY, 0.0, 1, solve_tol,
                                     solve_max, None, 1, h.data_ptr(), rhs.data_ptr(), C.byref(it),
                                     C.byref(res), ws.data_ptr(), ws.numel(), st),
            "osc_pcg_solve_system",
        )
    hh = h.cpu().numpy().astype(np.float32)
    if not np.all(np.isfinite(hh)):  # the reference falls back to uniform gates on failure (:149-150,162-163)
        hh = np.ones(N, dtype=np.float32)
    # 4. clamp / normalise (diffusion.py:125-129), on the host like the reference (N floats)
    if clamp:
        h_min = float(np.min(hh))
        h_max = float(np.max(hh))
        hh = np.ones(N, dtype=np.float32) if h_max - h_min < 1e-12 else (hh - h_min) / (h_max - h_min)
    return np.clip(hh, 0.0, 1.0).astype(np.float32)
