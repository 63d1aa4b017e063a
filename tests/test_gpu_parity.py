"""GPU parity tests (run on the B200 box: `pytest -m gpu`).

Every test drives the CUDA path through the reference-facing class / the C ABI and compares
with (a) the committed golden fixtures produced by the real reference, (b) the reference's own
artefacts, (c) the sparse oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): neighbour index sets bit-exact; U, U*, deltaH within
1e-5 relative (norm-wise for arrays); CG iteration counts within +-1; residual
scalars compared at equal iteration count.

Every other receipt quantity (coh_drop / anchor / query sums, null-point z and residual, bundle score and
alignment, the residual scalars) is held to max(1e-5, 10 x the reference's OWN noise floor), measured by
oracle/noise_floor.py (tests/golden/noise_floor.json): the unmodified reference run on row-permuted copies
of the same case, i.e. only the order in which NumPy/OpenBLAS sums changes.  The floors are 1e-7 .. 8e-7 for
the receipt terms and 4e-6 for the residual scalars (a difference of numbers five orders of magnitude
larger); the deviations of the CUDA path measured on B200 are 1.2e-7 .. 1.2e-6 and 2.5e-6
(profiles/r02_parity_margins.json, tools/dev_parity_margins.py).
"""
import numpy as np
import pytest

from oracle import cases
from oracle.sparse import SparseLattice
from tests.helpers import artefacts, load_golden, rel

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _floor_tol(key):
    """max(1e-5, 10 x the reference's own permutation noise floor for `key`)."""
    import json
    import os

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "noise_floor.json")) as f:
        fl = json.load(f)
    worst = max(v[key] for name, v in fl.items() if not name.startswith("_"))
    return max(TOL, 10.0 * worst)


TOL_RES = _floor_tol("settle_res")          # residual scalars: 4.0e-5
TOL_NULL = max(_floor_tol("null_z"), _floor_tol("null_residual"))   # 1e-5
TOL_COH = _floor_tol("coh_drop_sum")        # 1e-5
TOL_BUNDLE = max(_floor_tol("bundle_score"), _floor_tol("bundle_align"))  # 1e-5


@pytest.fixture(scope="module")
def api():
    import torch

    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import oscillink_b200

    return oscillink_b200


def _make(api, c, **extra):
    lat = api.OscillinkLattice(c["Y"], kneighbors=c["k"], row_cap_val=c["cap"], lamG=c["lam"][0],
                               lamC=c["lam"][1], lamQ=c["lam"][2], deterministic_k=c["det"], **extra)
    lat.set_query(c["psi"], gates=c["gates"])
    if c["chain"] is not None:
        lat.add_chain(c["chain"], lamP=c["lamP"], weights=c["weights"])
    return lat


def _nbr_table(lat):
    nbr = lat._nbr.cpu().numpy()
    return nbr


@pytest.mark.parametrize("name", cases.NAMES)
def test_lattice_matches_reference_golden(api, name):
    g, z = load_golden(name)
    c = cases.build(name)
    lat = _make(api, c)
    # ---- graph: bit-exact neighbour sets, degree statistics
    assert lat._kneighbors == g["k_effective"]
    nbr = _nbr_table(lat)
    kk = nbr.shape[1]
    assert np.array_equal(nbr, z["nbr"][:, :kk]), "neighbour index table differs from the reference"
    assert int((lat._A > 0).sum().item()) == g["nnz"]
    np.testing.assert_allclose(lat.sqrt_deg, z["sqrt_deg"], rtol=2e-6)
    assert rel(float(lat._A.double().sum().item()), g["A_sum"]) < 1e-6 or g["A_sum"] == 0
    assert lat._signature() == g["state_sig_init"]
    # ---- settle
    st = lat.settle(**c["settle_kw"])
    assert abs(st["iters"] - g["settle"]["iters"]) <= 1
    if st["iters"] == g["settle"]["iters"] and g["settle"]["res"] > 1e-6:
        assert rel(st["res"], g["settle"]["res"]) < TOL_RES
    U1 = lat.U.copy()
    if c["second_settle"]:
        st2 = lat.settle(**c["second_settle"])
        assert abs(st2["iters"] - g["settle2"]["iters"]) <= 1
        assert rel(np.sqrt((lat.U.astype(np.float64) ** 2).sum()), g["U2"]["fro"]) < TOL
    else:
        assert rel(np.sqrt((U1.astype(np.float64) ** 2).sum()), g["U"]["fro"]) < TOL
    # ---- light receipt
    lat.set_receipt_detail("light")
    rec = lat.receipt()
    assert abs(rec["meta"]["ustar_iters"] - g["ustar"]["iters"]) <= 1
    assert rel(rec["deltaH_total"], g["deltaH"]) < TOL
    assert rec["meta"]["avg_degree"] == pytest.approx(g["avg_degree"], rel=1e-12)
    assert rec["meta"]["edge_density"] == pytest.approx(g["edge_density"], rel=1e-12)
    assert rec["meta"]["state_sig"] == g["state_sig"]
    assert rec["coh_drop_sum"] == 0.0 and rec["null_points"] == []
    Us = lat.solve_Ustar()
    assert rel(np.sqrt((Us.astype(np.float64) ** 2).sum()), g["Ustar"]["fro"]) < TOL
    np.testing.assert_allclose(np.sqrt((Us.astype(np.float64) ** 2).sum(axis=0))[:8],
                               g["Ustar"]["colnorm_head"], rtol=TOL)
    if "Ustar_rows" in z:
        step = int(z["row_step"])
        num = np.linalg.norm(Us[::step].astype(np.float64) - z["Ustar_rows"])
        assert num / max(np.linalg.norm(z["Ustar_rows"]), 1e-30) < TOL
        if not c["second_settle"]:
            num = np.linalg.norm(U1[::step].astype(np.float64) - z["U_rows"])
            assert num / max(np.linalg.norm(z["U_rows"]), 1e-30) < TOL
    # ---- full receipt
    if c["full"]:
        lat.set_receipt_detail("full")
        rf = lat.receipt()
        gf = g["full"]
        assert rel(rf["coh_drop_sum"], gf["coh_drop_sum"]) < TOL_COH or abs(
            rf["coh_drop_sum"] - gf["coh_drop_sum"]) < 1e-6
        assert rel(rf["anchor_pen_sum"], gf["anchor_pen_sum"]) < TOL
        assert rel(rf["query_term_sum"], gf["query_term_sum"]) < TOL
        assert len(rf["null_points"]) == gf["n_null"]
        assert [e["edge"] for e in rf["null_points"]] == z["null_edges"].tolist()
        if gf["n_null"]:
            np.testing.assert_allclose([e["z"] for e in rf["null_points"]], z["null_z"], rtol=TOL_NULL)
            np.testing.assert_allclose([e["residual"] for e in rf["null_points"]], z["null_R"], rtol=TOL_NULL)


@pytest.mark.parametrize("name", sorted(artefacts()))
def test_lattice_matches_reference_artefacts(api, name):
    """Known answers committed in the reference repository itself."""
    ref = artefacts()[name]
    c = cases.build(name)
    lat = _make(api, c)
    lat.settle(**c["settle_kw"])
    lat.set_receipt_detail("full" if "null_points" in ref else "light")
    rec = lat.receipt()
    assert rec["meta"]["ustar_iters"] == ref["ustar_iters"]
    assert rel(rec["meta"]["ustar_res"], ref["ustar_res"]) < TOL_RES
    assert rel(rec["deltaH_total"], ref["deltaH"]) < TOL
    if "null_points" in ref:
        assert len(rec["null_points"]) == ref["null_points"]
        first = rec["null_points"][0]
        assert first["edge"] == ref["sample_null"]["edge"]
        assert rel(first["z"], ref["sample_null"]["z"]) < TOL_NULL
        assert rel(first["residual"], ref["sample_null"]["residual"]) < TOL_NULL


@pytest.mark.parametrize("name", ["config2_1200", "gates_300", "perf_400", "zero_row_40"])
def test_graph_arrays_match_sparse_oracle(api, name):
    """Array-by-array comparison with the oracle that shares the ELL layout."""
    c = cases.build(name)
    lat = _make(api, c)
    o = SparseLattice(c["Y"], k=c["k"], cap=c["cap"], lamG=c["lam"][0], lamC=c["lam"][1], lamQ=c["lam"][2])
    nbr = lat._nbr.cpu().numpy()
    assert np.array_equal(nbr, o.nbr.astype(np.int32))
    assert np.array_equal(lat._deg.cpu().numpy(), o.deg)
    np.testing.assert_allclose(lat._A.cpu().numpy(), o.A, rtol=2e-6, atol=1e-9)
    np.testing.assert_allclose(lat._W.cpu().numpy(), o.W, rtol=4e-6, atol=1e-9)
    np.testing.assert_allclose(lat.sqrt_deg, o.sd, rtol=2e-6)
    np.testing.assert_allclose(lat._gap.cpu().numpy(), o.gap, rtol=1e-3, atol=2e-7)


def test_simt_engine_and_default_engine_agree(api):
    """Both kNN engines feed the same canonical rescoring; the tables must be identical."""
    c = cases.build("config2_1200")
    a = _make(api, c)
    from oscillink_b200 import _cabi

    b = api.OscillinkLattice.__new__(api.OscillinkLattice)
    b.__init__(c["Y"], kneighbors=c["k"], deterministic_k=True)
    b._knn_engine = _cabi.KNN_SIMT
    b._build_graph()
    assert np.array_equal(a._nbr.cpu().numpy(), b._nbr.cpu().numpy())
    np.testing.assert_array_equal(a._A.cpu().numpy(), b._A.cpu().numpy())


def test_general_pcg_path_matches_batched_kernel(api):
    """config #2 goes through the slab kernel; force the HBM-resident K2 path and compare."""
    c = cases.build("config2_1200")
    g, _ = load_golden("config2_1200")
    a = _make(api, c)
    sa = a.settle(max_iters=12, tol=1e-3)
    b = _make(api, c)
    sb = b.settle(max_iters=12, tol=1e-3, inertia=1e-30)  # inertia>0 routes to osc_pcg_solve; w~0 => x0=U
    assert sa["iters"] == sb["iters"] == g["settle"]["iters"]
    assert rel(sa["res"], sb["res"]) < 1e-3
    num = np.linalg.norm(a.U.astype(np.float64) - b.U.astype(np.float64))
    assert num / np.linalg.norm(a.U.astype(np.float64)) < 1e-6


def test_reference_invariants(api):
    """Property tests mirrored from the reference suite (tests/test_new_invariants.py:21-47,
    tests/test_spd_and_deltaH.py:6-17, tests/test_graph_helpers.py:6-11)."""
    rs = np.random.RandomState(5)
    Y = rs.randn(24, 12).astype(np.float32)
    lat = api.OscillinkLattice(Y, kneighbors=5, deterministic_k=True)
    assert np.allclose(lat.A, lat.A.T, atol=1e-6)
    # tie-break determinism on all-equal rows
    ones = np.ones((10, 4), dtype=np.float32)
    l1 = api.OscillinkLattice(ones, kneighbors=4, deterministic_k=True)
    l2 = api.OscillinkLattice(ones, kneighbors=4, deterministic_k=True)
    assert np.array_equal(l1._nbr.cpu().numpy(), l2._nbr.cpu().numpy())
    # k clamp
    l3 = api.OscillinkLattice(rs.randn(6, 3).astype(np.float32), kneighbors=10)
    assert l3._kneighbors == 5
    # N = 1 -> empty graph
    l4 = api.OscillinkLattice(np.zeros((1, 4), dtype=np.float32), kneighbors=6)
    assert l4.A.shape == (1, 1) and float(l4.A.sum()) == 0.0
    # deltaH >= 0 with a chain
    Y = rs.randn(80, 64).astype(np.float32)
    psi = Y[:20].mean(axis=0)
    psi = (psi / (np.linalg.norm(psi) + 1e-12)).astype(np.float32)
    lat = api.OscillinkLattice(Y, kneighbors=6, lamG=1.0, lamC=0.5, lamQ=4.0)
    lat.set_query(psi=psi)
    lat.add_chain([1, 3, 5, 7], lamP=0.2)
    lat.settle(dt=1.0, max_iters=8, tol=1e-3)
    assert lat.receipt()["deltaH_total"] >= -1e-5


def test_cache_counters_signing_and_state_roundtrip(api, tmp_path):
    """tests/test_export_import_and_cache.py:6-57, tests/test_signature_roundtrip.py:6-15,
    tests/test_receipts_verify.py."""
    rs = np.random.RandomState(9)
    Y = rs.randn(60, 16).astype(np.float32)
    lat = api.OscillinkLattice(Y, kneighbors=4, deterministic_k=True)
    lat.set_query(rs.randn(16).astype(np.float32))
    events = []
    lat.set_logger(lambda ev, p: events.append(ev))
    lat.solve_Ustar()
    lat.solve_Ustar()
    assert lat.stats == {"ustar_solves": 1, "ustar_cache_hits": 1}
    assert "ustar_solve" in events and "ustar_cache_hit" in events
    lat.rebuild_graph(kneighbors=3)
    assert lat._Ustar_cache is None
    lat.set_receipt_secret("s3cret")
    lat.settle()
    rec = lat.receipt()
    assert api.verify_receipt(rec, "s3cret") and not api.verify_receipt(rec, "wrong")
    lat.set_signature_mode("extended")
    ok, payload = api.verify_receipt_mode(lat.receipt(), "s3cret", require_mode="extended")
    assert ok and payload["graph"]["k"] == 3
    # JSON + NPZ round trip keeps the signature and deltaH
    sig = lat._signature()
    for fmt in ("json", "npz"):
        p = str(tmp_path / f"state.{fmt}")
        lat.save_state(p, format=fmt)
        if fmt == "json":
            import json

            twin = api.OscillinkLattice.from_state(json.load(open(p)))
        else:
            twin = api.OscillinkLattice.from_npz(p)
        assert twin._signature() == sig
        twin.settle()
        assert abs(twin.receipt()["deltaH_total"] - rec["deltaH_total"]) <= 1e-2 * max(1.0, abs(rec["deltaH_total"]))
    with pytest.raises(ValueError):
        lat.set_gates(np.ones(3, dtype=np.float32))
    with pytest.raises(ValueError):
        lat.add_chain([0], lamP=0.1)
    with pytest.raises(ValueError):
        lat.set_receipt_detail("verbose")


def test_null_cap_env(api, monkeypatch):
    c = cases.build("readme_80")
    lat = _make(api, c)
    lat.settle()
    monkeypatch.setenv("OSCILLINK_RECEIPT_NULL_CAP", "5")
    rec = lat.receipt()
    assert len(rec["null_points"]) == 5
    s = rec["meta"]["null_points_summary"]
    assert s["null_cap_applied"] and s["total_null_points"] == 80 and s["returned_null_points"] == 5
    zs = [e["z"] for e in rec["null_points"]]
    assert zs == sorted(zs, reverse=True)


@pytest.mark.parametrize("shape", [(1, 300, 64, 6), (2, 1200, 384, 8), (1, 97, 20, 5), (3, 640, 128, 8),
                                   (1, 3000, 768, 16), (1, 130, 4, 28)])
def test_tc_and_simt_engines_build_identical_graphs(shape):
    """tcgen05 3xTF32 engine vs CUDA-core engine: candidate order may differ inside near-ties,
    the canonical rescoring must make the final top-k tables and weights bit-identical."""
    import ctypes as C

    import torch

    from oscillink_b200 import _cabi

    B, N, D, k = shape
    lib = _cabi.load()
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev)
    gen.manual_seed(N + D)
    Y = torch.randn((B, N, D), generator=gen, device=dev)
    res = {}
    engines = [("simt", _cabi.KNN_SIMT), ("tc", _cabi.KNN_TC)]
    if N > 128:  # the single-product engine is a 2-CTA kernel: more than one 128-row panel
        engines.append(("tc1", _cabi.KNN_TC1))
        assert _cabi.knn_plan(N, N, D, k, _cabi.KNN_TC1)[2] == pytest.approx(1.25e-3)
        if D % 8 == 0:  # fp16 single-product engine
            engines.append(("tch", _cabi.KNN_TCH))
            assert _cabi.knn_plan(N, N, D, k, _cabi.KNN_TCH)[2] == pytest.approx(1.25e-3)
    for name, eng in engines:
        nbr = torch.empty((B, N, k), dtype=torch.int32, device=dev)
        A = torch.empty((B, N, k), dtype=torch.float32, device=dev)
        W = torch.empty_like(A)
        deg = torch.empty((B, N), dtype=torch.int32, device=dev)
        sd = torch.empty((B, N), dtype=torch.float32, device=dev)
        nnz = torch.zeros(B, dtype=torch.int64, device=dev)
        gap = torch.empty((B, N), dtype=torch.float32, device=dev)
        need = C.c_size_t(0)
        _cabi.check(lib.osc_knn_build_workspace(B, N, D, k, eng, C.byref(need)))
        ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
        _cabi.check(lib.osc_knn_build(Y.data_ptr(), B, N, D, k, 1.0, eng, nbr.data_ptr(), A.data_ptr(),
                                      W.data_ptr(), deg.data_ptr(), sd.data_ptr(), nnz.data_ptr(),
                                      gap.data_ptr(), ws.data_ptr(), ws.numel(),
                                      torch.cuda.current_stream().cuda_stream), name)
        torch.cuda.synchronize()
        res[name] = (nbr.cpu().numpy(), A.cpu().numpy(), W.cpu().numpy(), nnz.cpu().numpy())
    for name, _ in engines[1:]:
        for a, b in zip(res["simt"], res[name]):
            assert np.array_equal(a, b), name


@pytest.mark.parametrize("shape,eng_name", [((2, 1200, 384, 8), "tc"), ((2, 1200, 384, 8), "tc1"),
                                            ((2, 1200, 384, 8), "tch"), ((1, 700, 72, 10), "tch"),
                                            ((1, 700, 64, 10), "tc1"), ((1, 300, 32, 6), "simt")])
def test_pruned_rescoring_equals_full_rescoring(shape, eng_name):
    """osc_knn_rescore_checked skips candidates that the engine's error bound proves to lie below the
    exact (k+1)-th score; tables, weights and the k/(k+1) gap must equal the unpruned pass."""
    import ctypes as C

    import torch

    from oscillink_b200 import _cabi

    B, N, D, k = shape
    lib = _cabi.load()
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev)
    gen.manual_seed(7 * N + D)
    Y = torch.randn((B, N, D), generator=gen, device=dev)
    flags = {"simt": _cabi.KNN_SIMT, "tc": _cabi.KNN_TC, "tc1": _cabi.KNN_TC1, "tch": _cabi.KNN_TCH}[eng_name]
    eng, kc, eps = _cabi.knn_plan(N, N, D, k, flags)
    assert eng == flags and kc >= k + 4
    st = torch.cuda.current_stream().cuda_stream
    Yn, hi, lo = torch.empty_like(Y), torch.empty_like(Y), torch.empty_like(Y)
    if eng == _cabi.KNN_TCH:
        hi = torch.empty(Y.shape, dtype=torch.float16, device=dev)
        _cabi.check(lib.osc_normalize_rows_f16(Y.data_ptr(), B * N, D, Yn.data_ptr(), hi.data_ptr(), st))
        assert torch.equal(hi, Yn.to(torch.float16)), "fp16 rows = round-to-nearest of Yn"
    else:
        _cabi.check(lib.osc_normalize_rows(Y.data_ptr(), B * N, D, Yn.data_ptr(), hi.data_ptr(),
                                           lo.data_ptr() if eng == _cabi.KNN_TC else None, st))
    ci = torch.empty((B, N, kc), dtype=torch.int32, device=dev)
    cs = torch.empty((B, N, kc), dtype=torch.float32, device=dev)
    _cabi.check(lib.osc_knn_candidates(Yn.data_ptr(), Yn.data_ptr(), hi.data_ptr(), lo.data_ptr(), hi.data_ptr(),
                                       lo.data_ptr(), B, N, 0, N, D, kc, eng, ci.data_ptr(), cs.data_ptr(),
                                       None, 0, st), "candidates")
    assert bool((cs[..., :-1] >= cs[..., 1:]).all()), "candidate lists must be sorted descending"
    # the engine's scores stay within the error bound the completeness check relies on
    exact = torch.gather(Yn.double() @ Yn.double().transpose(1, 2), 2, ci.clamp(min=0).long())
    err = float(((cs.double() - exact).abs() * (ci >= 0)).max())
    assert err <= eps, (eng_name, err, eps)
    out = {}
    for mode in ("full", "pruned"):
        ti = torch.empty((B, N, k), dtype=torch.int32, device=dev)
        ts = torch.empty((B, N, k), dtype=torch.float32, device=dev)
        gap = torch.empty((B, N), dtype=torch.float32, device=dev)
        if mode == "full":
            _cabi.check(lib.osc_knn_rescore(Yn.data_ptr(), Yn.data_ptr(), B, N, N, D, ci.data_ptr(), kc, k,
                                            ti.data_ptr(), ts.data_ptr(), gap.data_ptr(), st), "rescore")
        else:
            nflag = torch.zeros(1, dtype=torch.int32, device=dev)
            need = C.c_size_t(0)
            _cabi.check(lib.osc_knn_rescore_workspace(B, N, C.byref(need)))
            ws = torch.empty(max(need.value, 256), dtype=torch.uint8, device=dev)
            _cabi.check(lib.osc_knn_rescore_checked(Yn.data_ptr(), Yn.data_ptr(), B, N, 0, N, D, ci.data_ptr(),
                                                    cs.data_ptr(), kc, k, eps, ti.data_ptr(), ts.data_ptr(),
                                                    gap.data_ptr(), nflag.data_ptr(), ws.data_ptr(), ws.numel(),
                                                    st), "rescore_checked")
            out["flagged"] = int(nflag.item())
        torch.cuda.synchronize()
        out[mode] = (ti.cpu().numpy(), ts.cpu().numpy(), gap.cpu().numpy())
    assert out["flagged"] == 0  # (an exhaustive row sees columns outside the list: not comparable)
    for a, b in zip(out["full"], out["pruned"]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("N", [264, 2048, 2056])
def test_packed_key_lists_handle_negative_scores_and_the_column_limit(N):
    """The fp16 engine keeps (score image | column) keys in one register for N <= 2048 and score/column
    lists above.  Near-simplex anchors make every similarity slightly NEGATIVE (-1/(N-1) + noise), which
    exercises the sign handling of the key order; 2048 / 2056 straddle the 11-bit column limit.  The
    scores are closer together than the engine's error bound, so the completeness check sends the rows
    through the exhaustive kernel as well.  The canonical top-k after re-scoring must equal the
    CUDA-core engine's, and the engine's scores must stay inside its error bound."""
    import ctypes as C

    import torch

    from oscillink_b200 import _cabi

    lib = _cabi.load()
    dev = torch.device("cuda")
    D, k = N, 6
    gen = torch.Generator(device=dev)
    gen.manual_seed(N)
    Y = (torch.eye(N, device=dev) - 1.0 / N + (0.1 / N) * torch.randn((N, D), generator=gen, device=dev))[None]
    st = torch.cuda.current_stream().cuda_stream
    tops = {}
    for name, flags in (("simt", _cabi.KNN_SIMT), ("tch", _cabi.KNN_TCH)):
        eng, kc, eps = _cabi.knn_plan(N, N, D, k, flags)
        assert eng == flags
        Yn = torch.empty_like(Y)
        hi = torch.empty(Y.shape, dtype=torch.float16, device=dev)
        _cabi.check(lib.osc_normalize_rows_f16(Y.data_ptr(), N, D, Yn.data_ptr(), hi.data_ptr(), st))
        ci = torch.empty((1, N, kc), dtype=torch.int32, device=dev)
        cs = torch.empty((1, N, kc), dtype=torch.float32, device=dev)
        _cabi.check(lib.osc_knn_candidates(Yn.data_ptr(), Yn.data_ptr(), hi.data_ptr(), None, hi.data_ptr(), None,
                                           1, N, 0, N, D, kc, eng, ci.data_ptr(), cs.data_ptr(), None, 0, st), name)
        assert bool((ci >= 0).all()) and bool((cs[..., :-1] >= cs[..., 1:]).all())
        exact = torch.gather(Yn.double() @ Yn.double().transpose(1, 2), 2, ci.long())
        assert float((cs.double() - exact).abs().max()) <= eps, name
        assert float(exact.max()) < 0.0  # the case really is all-negative
        ti = torch.empty((1, N, k), dtype=torch.int32, device=dev)
        ts = torch.empty((1, N, k), dtype=torch.float32, device=dev)
        gap = torch.empty((1, N), dtype=torch.float32, device=dev)
        nflag = torch.zeros(1, dtype=torch.int32, device=dev)
        need = C.c_size_t(0)
        _cabi.check(lib.osc_knn_rescore_workspace(1, N, C.byref(need)))
        ws = torch.empty(max(need.value, 256), dtype=torch.uint8, device=dev)
        _cabi.check(lib.osc_knn_rescore_checked(Yn.data_ptr(), Yn.data_ptr(), 1, N, 0, N, D, ci.data_ptr(),
                                                cs.data_ptr(), kc, k, eps, ti.data_ptr(), ts.data_ptr(),
                                                gap.data_ptr(), nflag.data_ptr(), ws.data_ptr(), ws.numel(), st),
                    "rescore_checked")
        torch.cuda.synchronize()
        tops[name] = (ti.cpu().numpy(), ts.cpu().numpy())
    assert np.array_equal(tops["simt"][0], tops["tch"][0])
    assert np.array_equal(tops["simt"][1], tops["tch"][1])


@pytest.mark.parametrize("name", ["quickstart_120", "readme_80", "config2_1200", "gates_300"])
def test_bundle_matches_reference(api, name):
    """f1: bundle() ids identical, scores / alignments within max(1e-5, 10 x the reference's own floor)."""
    g, _ = load_golden(name)
    c = cases.build(name)
    lat = _make(api, c)
    lat.settle(**c["settle_kw"])
    if c["second_settle"]:
        lat.settle(**c["second_settle"])
    out = lat.bundle(k=c["bundle_k"])
    assert [e["id"] for e in out] == [e["id"] for e in g["bundle"]]
    np.testing.assert_allclose([e["score"] for e in out], [e["score"] for e in g["bundle"]], rtol=TOL_BUNDLE, atol=1e-7)
    np.testing.assert_allclose([e["align"] for e in out], [e["align"] for e in g["bundle"]], rtol=TOL_BUNDLE, atol=1e-7)
    assert lat.bundle(k=0) == []


@pytest.mark.parametrize("name", ["quickstart_120", "perf_400", "gates_300"])
def test_chain_receipt_matches_reference(api, name):
    """f2: chain_receipt() verdict, weakest link and per-edge z-scores."""
    g, _ = load_golden(name)
    c = cases.build(name)
    lat = _make(api, c)
    lat.settle(**c["settle_kw"])
    if c["second_settle"]:
        lat.settle(**c["second_settle"])
    cr = lat.chain_receipt(c["chain"])
    ref = g["chain_receipt"]
    assert cr["verdict"] == ref["verdict"]
    assert cr["weakest_link"]["k"] == ref["weakest_link"]["k"]
    assert cr["weakest_link"]["edge"] == ref["weakest_link"]["edge"]
    assert rel(cr["weakest_link"]["zscore"], ref["weakest_link"]["zscore"]) < TOL_NULL
    assert abs(cr["coherence_gain"] - ref["coherence_gain"]) <= TOL_NULL * max(1.0, abs(ref["coherence_gain"]))
    for a, b in zip(cr["edges"], ref["edges"]):
        assert a["edge"] == b["edge"]
        for key in ("z_struct", "z_path", "r_struct", "r_path"):
            assert abs(a[key] - b[key]) <= TOL_NULL * max(1.0, abs(b[key])), (key, a, b)


def test_perf_snapshot_weakest_link(api):
    """The reference's own perf_snapshot.json pins the chain verdict of the N=400 benchmark."""
    ref = artefacts()["perf_400"]
    c = cases.build("perf_400")
    lat = _make(api, c)
    lat.settle(**c["settle_kw"])
    cr = lat.chain_receipt(c["chain"])
    assert cr["verdict"] == ref["chain_verdict"]
    assert cr["weakest_link"]["edge"] == ref["weakest_link"]["edge"]
    assert rel(cr["weakest_link"]["zscore"], ref["weakest_link"]["zscore"]) < TOL_NULL


# --------------------------------------------------------------------------- candidate-list completeness
def test_exact_ties_beyond_the_candidate_margin_take_the_exhaustive_path(api):
    """200 identical rows: every one of them has 199 columns tied at similarity 1.0, far more than the
    k+4 candidates the approximate pass keeps.  graph.py:46-49 wants the k lowest indices among the
    ties; the completeness check must notice the list cannot be proven complete and redo those rows."""
    import torch

    rs = np.random.RandomState(3)
    base = rs.randn(400, 48).astype(np.float32)
    Y = np.concatenate([base, np.repeat(base[:1], 200, axis=0)], axis=0)
    Y = Y[rs.permutation(len(Y))]
    o = SparseLattice(Y, k=6)
    lat = api.OscillinkLattice(Y, kneighbors=6, deterministic_k=True)
    assert np.array_equal(lat._nbr.cpu().numpy(), o.nbr.astype(np.int32))
    np.testing.assert_allclose(lat._A.cpu().numpy(), o.A, rtol=2e-6, atol=1e-9)
    bl = api.BatchedLattices(torch.from_numpy(Y[None]).cuda(), kneighbors=6)
    assert int(bl.n_exhaustive.item()) >= 200
    assert np.array_equal(bl.nbr[0].cpu().numpy(), o.nbr.astype(np.int32))


def test_exhaustive_rows_equal_candidate_rows(api):
    """eps = 10 forces EVERY row through knn_exact_rows_kernel; the tables must equal the normal path."""
    import ctypes as C

    import torch

    from oscillink_b200 import _cabi

    c = cases.build("perf_400")
    lib = _cabi.load()
    dev = torch.device("cuda")
    Y = torch.from_numpy(c["Y"]).to(dev)
    N, D = Y.shape
    k, kc = 6, 10
    Yn = torch.empty_like(Y)
    st = torch.cuda.current_stream().cuda_stream
    _cabi.check(lib.osc_normalize_rows(Y.data_ptr(), N, D, Yn.data_ptr(), None, None, st))
    ci = torch.empty((N, kc), dtype=torch.int32, device=dev)
    cs = torch.empty((N, kc), dtype=torch.float32, device=dev)
    _cabi.check(lib.osc_knn_candidates(Yn.data_ptr(), Yn.data_ptr(), None, None, None, None, 1, N, 0, N, D, kc,
                                       _cabi.KNN_SIMT, ci.data_ptr(), cs.data_ptr(), None, 0, st))
    need = C.c_size_t(0)
    _cabi.check(lib.osc_knn_rescore_workspace(1, N, C.byref(need)))
    ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
    outs = []
    for eps in (_cabi.KNN_EPS, 10.0):
        ti = torch.empty((N, k), dtype=torch.int32, device=dev)
        ts = torch.empty((N, k), dtype=torch.float32, device=dev)
        gp = torch.empty(N, dtype=torch.float32, device=dev)
        nf = torch.zeros(1, dtype=torch.int32, device=dev)
        _cabi.check(lib.osc_knn_rescore_checked(Yn.data_ptr(), Yn.data_ptr(), 1, N, 0, N, D, ci.data_ptr(),
                                                cs.data_ptr(), kc, k, eps, ti.data_ptr(), ts.data_ptr(),
                                                gp.data_ptr(), nf.data_ptr(), ws.data_ptr(), ws.numel(), st))
        outs.append((ti.cpu().numpy(), ts.cpu().numpy(), gp.cpu().numpy(), int(nf.item())))
    assert outs[0][3] == 0 and outs[1][3] == N
    assert np.array_equal(outs[0][0], outs[1][0])
    assert np.array_equal(outs[0][1], outs[1][1])
    assert np.array_equal(outs[0][2], outs[1][2])


def test_sparse_state_format_roundtrip(api, tmp_path):
    """Row f4: the O(N k) ELL state format restores the same lattice as the reference's dense format
    (lattice.py:582-726): identical signature, identical graph, deltaH within the reference test's bound
    (tests/test_export_import_and_cache.py:6-25 allows 1e-2; here 1e-6)."""
    c = cases.build("perf_400")
    lat = _make(api, c)
    lat.settle()
    lat.set_receipt_detail("light")
    dh = lat.receipt()["deltaH_total"]
    st_d = lat.export_state()
    st_s = lat.export_state(graph_format="ell")
    assert "A" in st_d and "A_ell" in st_s and "A" not in st_s
    assert st_d["provenance"] == st_s["provenance"]
    assert len(st_s["A_ell"]["nbr"]) == lat.N and len(st_s["A_ell"]["nbr"][0]) == lat._nbr.shape[1]
    with pytest.raises(ValueError):
        lat.export_state(graph_format="csr")
    for fmt in ("json", "npz"):
        path = str(tmp_path / f"state_ell.{fmt}")
        lat.save_state(path, format=fmt, graph_format="ell")
        if fmt == "npz":
            back = api.OscillinkLattice.from_npz(path)
        else:
            import json

            with open(path) as f:
                back = api.OscillinkLattice.from_state(json.load(f))
        assert back._signature() == lat._signature()
        assert np.array_equal(back._nbr.cpu().numpy(), lat._nbr.cpu().numpy())
        np.testing.assert_allclose(back._W.cpu().numpy(), lat._W.cpu().numpy(), rtol=6e-7, atol=0)
        back.settle()
        back.set_receipt_detail("light")
        assert rel(back.receipt()["deltaH_total"], dh) < 1e-6
    # the dense and the sparse import agree with each other
    bd = api.OscillinkLattice.from_state(st_d)
    bs = api.OscillinkLattice.from_state(st_s)
    assert np.array_equal(bd._nbr.cpu().numpy(), bs._nbr.cpu().numpy())
    np.testing.assert_allclose(bd._W.cpu().numpy(), bs._W.cpu().numpy(), rtol=6e-7, atol=0)


@pytest.mark.parametrize("name", ["perf_400", "quickstart_120", "gates_300"])
def test_receipt_dynamics_match_reference(api, name, monkeypatch):
    """OSCILLINK_RECEIPT_DYNAMICS=1 (lattice.py:825-927): temperature, step deltaH, edge flows and the BFS
    radius of the first settle against the real reference (oracle/make_golden_dynamics.py)."""
    import json
    import os

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dynamics.json")) as f:
        ref = json.load(f)[name]
    monkeypatch.setenv("OSCILLINK_RECEIPT_DYNAMICS", "1")
    c = cases.build(name)
    lat = _make(api, c)
    lat.settle(**c["settle_kw"])
    dyn = lat.receipt()["meta"]["dynamics"]
    for key in ("temperature", "step_deltaH", "flow_total", "move2_mean", "move2_max"):
        assert rel(dyn[key], ref[key]) < TOL, key
    assert rel(dyn["viscosity_step"], ref["viscosity_step"]) < TOL
    assert dyn["radius"] == ref["radius"]
    assert [e["edge"] for e in dyn["top_flows"]] == [e["edge"] for e in ref["top_flows"]]
    np.testing.assert_allclose([e["flow"] for e in dyn["top_flows"]], [e["flow"] for e in ref["top_flows"]], rtol=TOL)


def test_deferred_x_update_is_bit_identical(monkeypatch):
    """osc_pcg_solve applies x += alpha p in the pass that rewrites p (pcg_pupdate_x_kernel) instead of the r
    update (solver.py:25 next to :26): the same operations on the same operands, so settle and U* must not
    change by a single bit -- converged (stops in the middle) and capped (max_iters reached) alike."""
    import oscillink_b200 as api

    rs = np.random.RandomState(21)
    Y = rs.randn(3000, 96).astype(np.float32)
    psi = Y[:32].mean(axis=0)
    psi = (psi / np.linalg.norm(psi)).astype(np.float32)
    out = {}
    for fuse in ("0", "1"):
        monkeypatch.setenv("OSC_PCG_FUSE_X", fuse)
        lat = api.OscillinkLattice(Y, kneighbors=8, deterministic_k=True)
        lat.set_query(psi)
        lat.add_chain([3, 9, 27, 81], lamP=0.2)
        capped = lat.settle(max_iters=2, tol=1e-9)
        U2 = lat.U.copy()
        st = lat.settle(max_iters=12, tol=1e-3)
        out[fuse] = (capped["iters"], capped["res"], U2, st["iters"], st["res"], lat.U.copy(), lat.solve_Ustar().copy())
    a, b = out["0"], out["1"]
    assert a[0] == b[0] == 2 and a[1] == b[1]
    assert np.array_equal(a[2], b[2])
    assert a[3] == b[3] and a[4] == b[4]
    assert np.array_equal(a[5], b[5])
    assert np.array_equal(a[6], b[6])


@pytest.mark.parametrize("warm", [False, True])
def test_fused_first_residual_is_bit_identical(monkeypatch, warm):
    """osc_pcg_solve / osc_dist_pcg_solve skip the setup pass when the start vector is Y or U itself: the first
    residual forms the right-hand side in place (pcg_setup_kernel's arithmetic) and X is first written by the x
    update of iteration 1.  Cold start, warm start from a previous settle, gates, a chain, and U*: no bit moves."""
    import oscillink_b200 as api
    from oscillink_b200.sharded_api import ShardedLattice

    rs = np.random.RandomState(33)
    Y = rs.randn(2500, 64).astype(np.float32)
    psi = Y[:32].mean(axis=0)
    psi = (psi / np.linalg.norm(psi)).astype(np.float32)
    gates = rs.uniform(0.2, 1.0, size=2500).astype(np.float32)
    out = {}
    for fuse in ("0", "1"):
        monkeypatch.setenv("OSC_PCG_FUSE_INIT", fuse)
        lat = api.OscillinkLattice(Y, kneighbors=6, deterministic_k=True)
        lat.set_query(psi, gates=gates)
        lat.add_chain([5, 50, 500, 1500], lamP=0.2)
        a = lat.settle(max_iters=3, tol=1e-9, warm_start=warm)
        U1 = lat.U.copy()
        b = lat.settle(max_iters=12, tol=1e-3, warm_start=warm)
        import torch

        sl = ShardedLattice(torch.as_tensor(Y).cuda(), 2500, kneighbors=6)
        sl.set_query(psi)
        c = sl.settle(max_iters=12, tol=1e-3)
        out[fuse] = (a["iters"], a["res"], U1, b["iters"], b["res"], lat.U.copy(), lat.solve_Ustar().copy(),
                     c["iters"], c["res"], sl.U_full())
    for x, y in zip(out["0"], out["1"]):
        assert np.array_equal(np.asarray(x), np.asarray(y))
