"""Seeded input recipes shared by oracle/make_golden.py (which runs the REAL reference on them)
and by the parity tests (which run the oracle / the CUDA path on them).  Test infrastructure.

Recipes follow the reference's own scripts so the committed artefacts apply:
  scripts/benchmark.py:45-48,59-63   (perf_400, config2_1200)
  scripts/scale_benchmark.py:23-35   (scale_*)
  examples/quickstart.py:7-18        (quickstart_120)
"""
from __future__ import annotations

import numpy as np


def unit(v):
    v = v.astype(np.float32)
    return v / (np.linalg.norm(v) + 1e-12)


def _case(Y, k, psi, **kw):
    d = dict(Y=Y, k=k, psi=psi, det=True, chain=None, lamP=0.2, weights=None, gates=None,
             settle_kw={}, full=True, bundle_k=0, chain_receipt=False, lam=(1.0, 0.5, 4.0), cap=1.0,
             save_rows=True, second_settle=None)
    d.update(kw)
    return d


def build(name):
    if name == "quickstart_120":
        Y = np.random.RandomState(0).randn(120, 128).astype(np.float32)
        return _case(Y, 6, unit(Y[:20].mean(axis=0)), det=False, chain=[2, 5, 7, 9],
                     settle_kw=dict(dt=1.0, max_iters=12, tol=1e-3), bundle_k=6, chain_receipt=True)
    if name == "readme_80":
        Y = np.random.RandomState(0).randn(80, 128).astype(np.float32)
        return _case(Y, 6, unit(Y[:20].mean(axis=0)), settle_kw=dict(max_iters=12, tol=1e-3), bundle_k=5)
    if name == "config2_1200":
        Y = np.random.RandomState(0).randn(1200, 384).astype(np.float32)
        return _case(Y, 8, unit(Y[:32].mean(axis=0)), settle_kw=dict(max_iters=12, tol=1e-3), bundle_k=8)
    if name == "perf_400":
        Y = np.random.RandomState(0).randn(400, 64).astype(np.float32)
        return _case(Y, 6, unit(Y[:32].mean(axis=0)), chain=list(range(8)),
                     settle_kw=dict(max_iters=12, tol=1e-3), chain_receipt=True)
    if name.startswith("scale_"):
        _, n, d = name.split("_")
        n, d = int(n), int(d)
        rs = np.random.RandomState(0)
        Y = rs.randn(n, d).astype(np.float32)
        psi = rs.randn(d).astype(np.float32)
        return _case(Y, 6, psi / (np.linalg.norm(psi) + 1e-12), chain=list(range(min(4, n))),
                     settle_kw=dict(max_iters=6, tol=1e-3), full=False, save_rows=(n <= 500))
    if name == "gates_300":
        rs = np.random.RandomState(7)
        Y = rs.randn(300, 96).astype(np.float32)
        gates = rs.uniform(0.1, 1.0, size=300).astype(np.float32)
        return _case(Y, 7, unit(Y[:10].mean(axis=0)), chain=[5, 17, 5, 40, 41, 42], lamP=0.35,
                     weights=[1.0, 0.5, 2.0, 1.0, 0.25], gates=gates, lam=(0.8, 0.9, 2.5), cap=0.7,
                     settle_kw=dict(dt=0.5, max_iters=12, tol=1e-4, warm_start=False),
                     second_settle=dict(dt=1.0, max_iters=12, tol=1e-3, inertia=0.4),
                     chain_receipt=True, bundle_k=4)
    if name == "ties_10":
        return _case(np.ones((10, 4), dtype=np.float32), 4, unit(np.ones(4)), settle_kw=dict(max_iters=4))
    rs = np.random.RandomState(3)
    Y6 = rs.randn(6, 3).astype(np.float32)
    Y2 = rs.randn(2, 5).astype(np.float32)
    Y1 = rs.randn(1, 4).astype(np.float32)
    Y40 = rs.randn(40, 16).astype(np.float32)
    Y40[3] = 0.0
    if name == "tiny_clamp_6":
        return _case(Y6, 10, unit(Y6[0]), settle_kw=dict(max_iters=8))
    if name == "tiny_2":
        return _case(Y2, 3, unit(Y2[0]), settle_kw=dict(max_iters=8))
    if name == "tiny_1":
        return _case(Y1, 6, unit(Y1[0]), settle_kw=dict(max_iters=8), full=False)
    if name == "zero_row_40":
        return _case(Y40, 5, unit(Y40[1]), settle_kw=dict(max_iters=12))
    if name == "noprecond_150":
        Y = np.random.RandomState(11).randn(150, 32).astype(np.float32)
        return _case(Y, 5, unit(Y[:8].mean(axis=0)),
                     settle_kw=dict(max_iters=20, tol=1e-4, precond="none"))
    raise KeyError(name)


NAMES = [
    "quickstart_120", "readme_80", "config2_1200", "perf_400",
    "scale_100_128", "scale_500_128", "scale_1000_128", "scale_2000_128",
    "scale_400_64", "scale_800_64", "scale_1200_64",
    "gates_300", "ties_10", "tiny_clamp_6", "tiny_2", "tiny_1", "zero_row_40", "noprecond_150",
]
