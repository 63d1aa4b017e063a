#!/usr/bin/env python
"""Golden values of the env-gated receipt dynamics (oscillink/core/lattice.py:825-927,
OSCILLINK_RECEIPT_DYNAMICS=1) from the REAL reference.  Build container only:

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_dynamics.py   -> tests/golden/dynamics.json
"""
import json
import os
import sys

os.environ["OSCILLINK_RECEIPT_DYNAMICS"] = "1"
REF = os.environ.get("OSC_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.dont_write_bytecode = True
from oscillink import OscillinkLattice  # noqa: E402  (the real reference)

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
from oracle import cases  # noqa: E402

out = {}
for name in ("perf_400", "quickstart_120", "gates_300"):
    c = cases.build(name)
    lam = c["lam"]
    lat = OscillinkLattice(c["Y"], kneighbors=c["k"], row_cap_val=c["cap"], lamG=lam[0], lamC=lam[1], lamQ=lam[2],
                           deterministic_k=c["det"])
    lat.set_query(c["psi"], gates=c["gates"])
    if c["chain"] is not None:
        lat.add_chain(c["chain"], lamP=c["lamP"], weights=c["weights"])
    lat.settle(**c["settle_kw"])
    rec = lat.receipt()
    dyn = rec["meta"]["dynamics"]
    out[name] = dyn
    print(name, {k: v for k, v in dyn.items() if k != "top_flows"}, len(dyn["top_flows"]))
with open(os.path.join(HERE, "..", "tests", "golden", "dynamics.json"), "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
