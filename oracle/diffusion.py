"""ORACLE (test infrastructure, NOT product code) -- screened-diffusion gates.

Restates oscillink/preprocess/diffusion.py:35-163 on top of oracle.dense's graph functions (which
restate oscillink/core/graph.py and are pinned to the real reference by tests/test_oracle_golden.py).
Parity: PINNED -- tests/golden/diffusion_*.npz are produced by oracle/make_golden_diffusion.py, which
imports the unmodified reference, and tests/test_oracle_golden.py::test_diffusion_oracle_* checks this
file against them.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

import numpy as np

from .dense import cap_rows, knn_adjacency, pcg, sym_laplacian

F32 = np.float32

# the seeded cases shared by the golden generator and the tests
CASES = {
    # scripts/benchmark_gating_compare.py:27-52 recipe behind gating_result.json (N=400 D=64 k=6)
    "diffusion_gating_400": dict(N=400, D=64, k=6, seed=123, beta=1.0, gamma=0.15, det=False, method="direct"),
    "diffusion_cg_300": dict(N=300, D=48, k=5, seed=5, beta=1.2, gamma=0.1, det=True, method="cg"),
    "diffusion_direct_60": dict(N=60, D=32, k=5, seed=42, beta=1.2, gamma=0.15, det=True, method="direct"),
    "diffusion_noclamp_120": dict(N=120, D=24, k=4, seed=9, beta=0.7, gamma=0.3, det=True, method="cg",
                                  clamp=False),
}


def case_inputs(name):
    c = CASES[name]
    rng = np.random.default_rng(c["seed"])
    Y = rng.normal(size=(c["N"], c["D"])).astype(F32)
    psi = rng.normal(size=(c["D"],)).astype(F32)
    if name == "diffusion_gating_400":  # benchmark_gating_compare.py:31-33: psi from the first rows
        psi = (Y[:20].mean(axis=0) / (np.linalg.norm(Y[:20].mean(axis=0)) + 1e-12)).astype(F32)
    return Y, psi, c


def diffusion_gates(Y, psi, *, k=6, cap=1.0, beta=1.0, gamma=0.1, deterministic=False, clamp=True,
                    method="direct", tol=1e-4, max_iters=256):
    """diffusion.py:96-129; returns (h, iters) -- iters is None for the direct solve."""
    Yf = np.asarray(Y, dtype=F32)
    psif = np.asarray(psi, dtype=F32)
    n = Yf.shape[0]
    A = cap_rows(knn_adjacency(Yf, k, deterministic=deterministic), cap)     # :100-107
    L, _ = sym_laplacian(A)                                                   # :108
    Yn = Yf / (np.linalg.norm(Yf, axis=1, keepdims=True) + 1e-12)              # :112
    psin = psif / (np.linalg.norm(psif) + 1e-12)                               # :113
    s = (Yn @ psin).astype(F32)                                                # :114
    s = beta * np.maximum(0.0, s)                                              # :118
    iters = None
    if method == "cg":                                                         # :138-150
        md = np.diag(L).astype(F32) + float(gamma)
        h, iters, _ = pcg(lambda x: (L @ x) + gamma * x, s.astype(F32)[:, None],
                          np.zeros((n, 1), dtype=F32), md, tol, max_iters)
        h = h[:, 0].astype(F32)
    else:                                                                      # :152-157
        h = np.linalg.solve(L + gamma * np.eye(n, dtype=F32), s).astype(F32)
    if clamp:                                                                  # :125-128
        lo, hi = float(np.min(h)), float(np.max(h))
        h = np.ones(n, dtype=F32) if hi - lo < 1e-12 else (h - lo) / (hi - lo)
    return np.clip(h, 0.0, 1.0).astype(F32), iters                             # :129
