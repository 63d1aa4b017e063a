#!/usr/bin/env python
"""Generate tests/golden/diffusion_*.npz by running the REAL reference's compute_diffusion_gates
(oscillink/preprocess/diffusion.py), build container only:

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_diffusion.py
"""
from __future__ import annotations

import os
import sys

import numpy as np

REF = os.environ.get("OSC_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.dont_write_bytecode = True
from oscillink import OscillinkLattice, compute_diffusion_gates  # noqa: E402  (the real reference)

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
from oracle.diffusion import CASES, case_inputs  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")


def main():
    for name in CASES:
        Y, psi, c = case_inputs(name)
        h = compute_diffusion_gates(Y, psi, kneighbors=c["k"], beta=c["beta"], gamma=c["gamma"],
                                    deterministic_k=c["det"], method=c["method"], clamp=c.get("clamp", True))
        out = {"h": h.astype(np.float32)}
        # the gated lattice the gates are meant for (benchmark_gating_compare.py:54-70): deltaH pins the
        # end-to-end effect of the gates on the settle path
        lat = OscillinkLattice(Y, kneighbors=c["k"], deterministic_k=c["det"])
        lat.set_query(psi, gates=np.clip(h, 0.0, 1.0))
        lat.settle(max_iters=12, tol=1e-3)
        lat.set_receipt_detail("light")
        out["deltaH_gated"] = np.float64(lat.receipt()["deltaH_total"])
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **out)
        print(name, "h mean", float(h.mean()), "min", float(h.min()), "max", float(h.max()),
              "dH gated", float(out["deltaH_gated"]))


if __name__ == "__main__":
    main()
