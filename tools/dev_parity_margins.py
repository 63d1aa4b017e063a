"""Dev (GPU box): observed deviation of the CUDA path from the reference goldens for the quantities whose
test tolerances are looser than 1e-5, next to the reference's own permutation noise floor
(tests/golden/noise_floor.json).  Prints one JSON line; kept under profiles/ as the measurement that
justifies the tolerances in tests/test_gpu_parity.py."""
import json, os, sys
import numpy as np
sys.path.insert(0, ".")
from oracle import cases
from tests.helpers import load_golden, rel
import oscillink_b200 as api

floor = json.load(open("tests/golden/noise_floor.json"))
out = {}
for name in cases.NAMES:
    g, z = load_golden(name)
    c = cases.build(name)
    lat = api.OscillinkLattice(c["Y"], kneighbors=c["k"], row_cap_val=c["cap"], lamG=c["lam"][0], lamC=c["lam"][1],
                               lamQ=c["lam"][2], deterministic_k=c["det"])
    lat.set_query(c["psi"], gates=c["gates"])
    if c["chain"] is not None:
        lat.add_chain(c["chain"], lamP=c["lamP"], weights=c["weights"])
    st = lat.settle(**c["settle_kw"])
    m = {}
    if st["iters"] == g["settle"]["iters"] and g["settle"]["res"] > 1e-6:
        m["settle_res"] = rel(st["res"], g["settle"]["res"])
    if not c["second_settle"] and "U_rows" in z:
        step = int(z["row_step"])
        m["U_rows"] = float(np.linalg.norm(lat.U[::step].astype(np.float64) - z["U_rows"]) / max(np.linalg.norm(z["U_rows"]), 1e-30))
    if c["second_settle"]:
        lat.settle(**c["second_settle"])
    lat.set_receipt_detail("light")
    rec = lat.receipt()
    m["deltaH"] = rel(rec["deltaH_total"], g["deltaH"])
    if rec["meta"]["ustar_iters"] == g["ustar"]["iters"] and g["ustar"]["res"] > 1e-7:
        m["ustar_res"] = rel(rec["meta"]["ustar_res"], g["ustar"]["res"])
    if c["full"]:
        lat.set_receipt_detail("full")
        rf = lat.receipt(); gf = g["full"]
        for k in ("coh_drop_sum", "anchor_pen_sum", "query_term_sum"):
            if abs(gf[k]) > 1e-6:
                m[k] = rel(rf[k], gf[k])
        if gf["n_null"] and [e["edge"] for e in rf["null_points"]] == z["null_edges"].tolist():
            m["null_z"] = float(np.max(np.abs(np.array([e["z"] for e in rf["null_points"]]) / z["null_z"] - 1)))
            m["null_residual"] = float(np.max(np.abs(np.array([e["residual"] for e in rf["null_points"]]) / z["null_R"] - 1)))
    if c["bundle_k"]:
        b = lat.bundle(k=c["bundle_k"])
        if [e["id"] for e in b] == [e["id"] for e in g["bundle"]]:
            m["bundle_score"] = max(rel(x["score"], y["score"]) for x, y in zip(b, g["bundle"]))
            m["bundle_align"] = max(rel(x["align"], y["align"]) for x, y in zip(b, g["bundle"]))
    out[name] = m
worst = {}
for name, m in out.items():
    for k, v in m.items():
        if v >= worst.get(k, (0, ""))[0]:
            worst[k] = (v, name)
ref_floor = {}
for name, f in floor.items():
    if name.startswith("_"): continue
    for k, v in f.items():
        if isinstance(v, float): ref_floor[k] = max(ref_floor.get(k, 0.0), v)
print(json.dumps({"worst_observed": {k: {"rel": v[0], "case": v[1]} for k, v in worst.items()},
                  "reference_permutation_floor": ref_floor, "per_case": out}))
