"""CPU-baseline worker (test/bench infrastructure): one end-to-end lattice of the serving batch
through the dense oracle, i.e. the reference's per-request sequence
(scripts/benchmark.py:45-70): build (deterministic kNN) + set_query + settle(12,1e-3) +
light receipt (stationary solve + deltaH).  Imports NumPy only (safe to spawn)."""
from __future__ import annotations

import numpy as np

N, D, K = 1200, 384, 8


def settle_one(b: int, n: int = N, d: int = D, k: int = K) -> float:
    from oracle.dense import DenseLattice

    rs = np.random.RandomState(b)
    Y = rs.randn(n, d).astype(np.float32)
    psi = Y[: min(32, n)].mean(axis=0)
    psi = (psi / (np.linalg.norm(psi) + 1e-12)).astype(np.float32)
    lat = DenseLattice(Y, k=k, deterministic=True)
    lat.set_query(psi)
    lat.settle(max_iters=12, tol=1e-3)
    us, _, _ = lat.stationary()
    return lat.delta_h(us)
