#!/usr/bin/env python
"""Generate tests/golden/*.json(.npz) by running the REAL reference.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

The reference is imported unmodified from /root/reference; nothing from it is
copied.  The fixtures are small (index tables, scalars, column norms, a strided
sample of rows) and are committed together with this script.

Cases:
  quickstart_120   examples/quickstart.py recipe, RandomState(0)
  readme_80        README snippet N=80 D=128 k=6
  config2_1200     N=1200 D=384 k=8 (BASELINE.json configs[1]), light + full receipt
  perf_400         scripts/benchmark.py N=400 D=64 k=6 chain range(8)  (perf_snapshot.json)
  scale_*          scripts/scale_benchmark.py N in {100,500,1000,2000} D=128, {400,800,1200} D=64
  gates_300        non-uniform gates + chain + weights + inertia start
  ties_10          all-equal rows (tests/test_new_invariants.py:28-40)
  tiny_*           N=1, N=2, k>=N clamp, zero row
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

REF = os.environ.get("OSC_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.dont_write_bytecode = True
from oscillink import OscillinkLattice  # noqa: E402  (the real reference)

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def nbr_table(A, k):
    """Row-major neighbour index table (-1 padded) of the dense adjacency."""
    n = A.shape[0]
    t = np.full((n, max(k, 1)), -1, dtype=np.int32)
    for i in range(n):
        nz = np.nonzero(A[i] > 0)[0]
        t[i, : len(nz)] = nz
    return t


def summarise(X):
    X = np.asarray(X, dtype=np.float64)
    return {
        "fro": float(np.sqrt((X * X).sum())),
        "colnorm_head": np.sqrt((X * X).sum(axis=0))[:8].tolist(),
        "sum": float(X.sum()),
    }


def run_case(name, Y, k, psi, *, det=True, chain=None, lamP=0.2, weights=None, gates=None,
             settle_kw=None, full=True, bundle_k=0, chain_receipt=False, lam=(1.0, 0.5, 4.0),
             cap=1.0, save_rows=True, second_settle=None):
    settle_kw = settle_kw or {}
    lat = OscillinkLattice(Y, kneighbors=k, row_cap_val=cap, lamG=lam[0], lamC=lam[1], lamQ=lam[2],
                           deterministic_k=det)
    if psi is not None:
        lat.set_query(psi, gates=gates)
    if chain is not None:
        lat.add_chain(chain, lamP=lamP, weights=weights)
    g = {
        "name": name, "N": int(lat.N), "D": int(lat.D), "k_requested": int(k),
        "k_effective": int(lat._kneighbors), "deterministic": bool(det), "lam": list(lam),
        "cap": cap, "chain": chain, "lamP": lamP if chain is not None else 0.0,
        "weights": weights, "settle_kw": settle_kw,
        "nnz": int((lat.A > 0).sum()),
        "A_sum": float(lat.A.astype(np.float64).sum()),
        "A_rowsum_max": float(lat.A.sum(axis=1).max()) if lat.N else 0.0,
        "sqrt_deg_sum": float(lat.sqrt_deg.astype(np.float64).sum()),
        "state_sig_init": lat._signature(),
    }
    arrays = {"nbr": nbr_table(lat.A, lat._kneighbors), "sqrt_deg": lat.sqrt_deg.astype(np.float32)}
    st = lat.settle(**settle_kw)
    g["settle"] = {"iters": st["iters"], "res": st["res"]}
    g["U"] = summarise(lat.U)
    if second_settle is not None:
        st2 = lat.settle(**second_settle)
        g["settle2"] = {"iters": st2["iters"], "res": st2["res"], "kw": second_settle}
        g["U2"] = summarise(lat.U)
    lat.set_receipt_detail("light")
    rec = lat.receipt()
    g["deltaH"] = rec["deltaH_total"]
    g["ustar"] = {"iters": rec["meta"]["ustar_iters"], "res": rec["meta"]["ustar_res"]}
    g["avg_degree"] = rec["meta"]["avg_degree"]
    g["edge_density"] = rec["meta"]["edge_density"]
    g["state_sig"] = rec["meta"]["state_sig"]
    Ustar = lat.solve_Ustar()
    g["Ustar"] = summarise(Ustar)
    if save_rows:
        step = max(1, lat.N // 16)
        arrays["U_rows"] = lat.U[::step].astype(np.float32)
        arrays["Ustar_rows"] = Ustar[::step].astype(np.float32)
        arrays["row_step"] = np.array(step)
    if full:
        lat.set_receipt_detail("full")
        rf = lat.receipt()
        g["full"] = {
            "coh_drop_sum": rf["coh_drop_sum"], "anchor_pen_sum": rf["anchor_pen_sum"],
            "query_term_sum": rf["query_term_sum"], "n_null": len(rf["null_points"]),
            "null_head": rf["null_points"][:5],
        }
        arrays["null_edges"] = np.array([e["edge"] for e in rf["null_points"]], dtype=np.int32).reshape(-1, 2)
        arrays["null_z"] = np.array([e["z"] for e in rf["null_points"]], dtype=np.float64)
        arrays["null_R"] = np.array([e["residual"] for e in rf["null_points"]], dtype=np.float64)
    if bundle_k:
        g["bundle"] = lat.bundle(k=bundle_k)
    if chain_receipt and chain is not None:
        cr = lat.chain_receipt(chain)
        g["chain_receipt"] = cr
    with open(os.path.join(OUT, f"{name}.json"), "w") as f:
        json.dump(g, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **arrays)
    print(f"{name}: nnz={g['nnz']} settle={g['settle']} ustar={g['ustar']} dH={g['deltaH']}")


def main():
    os.makedirs(OUT, exist_ok=True)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from oracle import cases

    for name in cases.NAMES:
        c = cases.build(name)
        Y, k, psi = c.pop("Y"), c.pop("k"), c.pop("psi")
        run_case(name, Y, k, psi, **c)


if __name__ == "__main__":
    main()
