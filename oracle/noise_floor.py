#!/usr/bin/env python
"""Measure the REFERENCE's own floating-point noise floor (test infrastructure; build container only).

The parity tests compare a few receipt quantities with tolerances looser than the 1e-5 of north_star
(coh_drop_sum, null-point z / residual, bundle score, the residual scalars).  Those quantities are
differences of nearly equal numbers or z-scores, so fp32 rounding is amplified; how much is measured here
on the reference itself: the lattice is a permutation-equivariant function of its rows, so running the
UNMODIFIED reference (imported from /root/reference) on a row-permuted copy of a case changes nothing but
the order in which NumPy/OpenBLAS sums -- the spread between the runs is the reference's own noise.

    PYTHONDONTWRITEBYTECODE=1 python oracle/noise_floor.py     -> tests/golden/noise_floor.json

tests/test_gpu_parity.py asserts each of those quantities within max(1e-5, 10 x measured floor).
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

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
from oracle import cases  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden", "noise_floor.json")


def run(c, perm):
    """Reference results on rows permuted by `perm` (new row t = old row perm[t]), mapped back to old ids."""
    inv = np.argsort(perm)
    Y = c["Y"][perm]
    gates = None if c["gates"] is None else c["gates"][perm]
    lam = c["lam"]
    lat = OscillinkLattice(Y, kneighbors=c["k"], row_cap_val=c["cap"], lamG=lam[0], lamC=lam[1], lamQ=lam[2],
                           deterministic_k=c["det"])
    lat.set_query(c["psi"], gates=gates)
    if c["chain"] is not None:
        lat.add_chain([int(inv[i]) for i in c["chain"]], lamP=c["lamP"], weights=c["weights"])
    st = lat.settle(**c["settle_kw"])
    lat.set_receipt_detail("full")
    rec = lat.receipt()
    out = {"settle_res": st["res"], "settle_iters": st["iters"], "ustar_res": rec["meta"]["ustar_res"],
           "deltaH": rec["deltaH_total"], "coh_drop_sum": rec["coh_drop_sum"],
           "anchor_pen_sum": rec["anchor_pen_sum"], "query_term_sum": rec["query_term_sum"]}
    out["null"] = {(int(perm[e["edge"][0]]), int(perm[e["edge"][1]])): (e["z"], e["residual"])
                   for e in rec["null_points"]}
    if c["bundle_k"]:
        out["bundle"] = {int(perm[b["id"]]): (b["score"], b["align"]) for b in lat.bundle(k=c["bundle_k"])}
    return out


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-30)


def main():
    res = {"_how": "max over 3 row permutations of |permuted - identity| / |identity|, unmodified reference, "
                   "NumPy " + np.__version__}
    for name in ("config2_1200", "perf_400", "gates_300", "quickstart_120", "readme_80"):
        c = cases.build(name)
        n = c["Y"].shape[0]
        base = run(c, np.arange(n))
        floor = {k: 0.0 for k in ("settle_res", "ustar_res", "deltaH", "coh_drop_sum", "anchor_pen_sum",
                                  "query_term_sum", "null_z", "null_residual", "bundle_score", "bundle_align")}
        notes = {"null_edge_sets_equal": True, "bundle_ids_equal": True, "iters_equal": True}
        for seed in (1, 2, 3):
            p = run(c, np.random.RandomState(seed).permutation(n))
            for k in ("settle_res", "ustar_res", "deltaH", "coh_drop_sum", "anchor_pen_sum", "query_term_sum"):
                floor[k] = max(floor[k], rel(p[k], base[k]))
            notes["iters_equal"] &= p["settle_iters"] == base["settle_iters"]
            notes["null_edge_sets_equal"] &= set(p["null"]) == set(base["null"])
            for e in set(p["null"]) & set(base["null"]):
                floor["null_z"] = max(floor["null_z"], rel(p["null"][e][0], base["null"][e][0]))
                floor["null_residual"] = max(floor["null_residual"], rel(p["null"][e][1], base["null"][e][1]))
            if "bundle" in base:
                notes["bundle_ids_equal"] &= set(p["bundle"]) == set(base["bundle"])
                for i in set(p["bundle"]) & set(base["bundle"]):
                    floor["bundle_score"] = max(floor["bundle_score"], rel(p["bundle"][i][0], base["bundle"][i][0]))
                    floor["bundle_align"] = max(floor["bundle_align"], rel(p["bundle"][i][1], base["bundle"][i][1]))
        res[name] = dict(floor, **notes)
        print(name, {k: f"{v:.2e}" for k, v in floor.items()}, notes)
    with open(OUT, "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
