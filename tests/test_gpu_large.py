"""Parity beyond the dense reference's reach (sparse restatement + sampled rows).

* N = 20 000: the whole lattice against oracle/sparse.py (itself pinned to the reference through
  oracle/dense.py): neighbour tables bit-exact, PCG iteration counts, U / U* / deltaH within 1e-5.
* N = 200 000 (OSC_TEST_LARGE_N overrides): the dense reference cannot run and a full CPU kNN is
  minutes, so neighbour SETS are compared on sampled rows (oracle top-k of the sample and of every
  neighbour of the sample gives the exact mutual sets), plus size-independent properties: graph
  symmetry, an independent fp64 evaluation of the settle residual, deltaH >= 0.
"""
import os

import numpy as np
import pytest

from oracle.sparse import SparseLattice, normalise_rows, topk_canonical

pytestmark = pytest.mark.gpu

TOL = 1e-5  # north_star: U*, deltaH and residual-normed quantities within 1e-5 relative


@pytest.fixture(scope="module")
def api():
    import torch

    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import oscillink_b200

    return oscillink_b200


def _psi(Y):
    p = Y[:32].mean(axis=0)
    return (p / (np.linalg.norm(p) + 1e-12)).astype(np.float32)


def _frob_rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def test_n20000_full_lattice_matches_sparse_oracle(api):
    rs = np.random.RandomState(7)
    N, D, k = 20000, 96, 10
    Y = rs.randn(N, D).astype(np.float32)
    psi = _psi(Y)
    chain = list(range(0, 64, 2))
    lat = api.OscillinkLattice(Y, kneighbors=k, deterministic_k=True)
    lat.set_query(psi)
    lat.add_chain(chain, lamP=0.2)
    st = lat.settle(max_iters=12, tol=1e-3)
    lat.set_receipt_detail("light")
    rec = lat.receipt()

    o = SparseLattice(Y, k=k)
    o.set_query(psi)
    o.add_chain(chain, lamP=0.2)
    so = o.settle(max_iters=12, tol=1e-3)
    us, uit, ures = o.stationary()

    # smallest k/(k+1) margin of this fixture is one fp32 ulp (1.19e-7): the canonical rule (exact dot
    # rounded once, then index order) is what both sides implement, so the tables must still agree
    assert np.array_equal(lat._nbr.cpu().numpy(), o.nbr.astype(np.int32))
    np.testing.assert_allclose(lat._W.cpu().numpy(), o.W, rtol=4e-6, atol=1e-9)
    assert abs(st["iters"] - so["iters"]) <= 1
    assert _frob_rel(lat.U, o.U) < TOL
    assert abs(rec["meta"]["ustar_iters"] - uit) <= 1
    assert _frob_rel(lat.solve_Ustar(), us) < TOL
    dh = o.delta_h(us)
    assert abs(rec["deltaH_total"] - dh) <= TOL * abs(dh)
    if st["iters"] == so["iters"]:
        assert abs(st["res"] - so["res"]) <= 1e-3 * so["res"]


def test_large_lattice_sampled_rows_and_properties(api):
    import torch

    N = int(os.environ.get("OSC_TEST_LARGE_N", "200000"))
    D, k, n_sample = 64, 10, 48
    rs = np.random.RandomState(11)
    Y = rs.randn(N, D).astype(np.float32)
    psi = _psi(Y)
    lat = api.OscillinkLattice(Y, kneighbors=k, deterministic_k=True)
    lat.set_query(psi)
    nbr = lat._nbr.cpu().numpy()
    W = lat._W.cpu().numpy()
    A = lat._A.cpu().numpy()

    # ---- sampled neighbour sets, bit-exact (graph.py:46-52,64-65 on the sample's 1-hop closure)
    Yn = normalise_rows(Y)
    sample = rs.choice(N, size=n_sample, replace=False)
    idx_s, sim_s, gap_s = topk_canonical(Yn, k, rows=sample)
    hop = np.unique(idx_s.reshape(-1))
    idx_h, sim_h, gap_h = topk_canonical(Yn, k, rows=hop)
    table = {int(r): (idx_h[t], sim_h[t]) for t, r in enumerate(hop)}
    for t, i in enumerate(sample):
        want = []
        for j, s in zip(idx_s[t], sim_s[t]):
            jj, js = table[int(j)]
            back = (jj == i) & (js > 0)
            if s > 0 and back.any():
                want.append(int(j))
        got = [int(j) for j in nbr[i] if j >= 0]
        assert got == sorted(want), f"row {i}: {got} != {sorted(want)}"

    # ---- symmetry of the assembled graph (test_new_invariants.py:21-25), all rows
    rows = np.repeat(np.arange(N), nbr.shape[1])[nbr.reshape(-1) >= 0]
    cols = nbr.reshape(-1)[nbr.reshape(-1) >= 0]
    w = W.reshape(-1)[nbr.reshape(-1) >= 0]
    a = A.reshape(-1)[nbr.reshape(-1) >= 0]
    key = rows.astype(np.int64) * N + cols
    rkey = cols.astype(np.int64) * N + rows
    assert np.array_equal(np.sort(key), np.sort(rkey)), "edge set is not symmetric"
    order_f, order_r = np.argsort(key), np.argsort(rkey)
    # the capped adjacency is exactly symmetric (graph.py:80-83); W = (A*inv_i)*inv_j is symmetric only
    # up to fp32 rounding of the two multiplications, in the reference as well (graph.py:89-90)
    assert np.array_equal(a[order_f], a[order_r]), "capped adjacency is not symmetric"
    np.testing.assert_allclose(w[order_f], w[order_r], rtol=3e-7)
    assert len(key) > N  # average degree well above 1

    # ---- settle, then an independent fp64 evaluation of the residual (solver.py:29)
    st = lat.settle(max_iters=12, tol=1e-3)
    assert 1 <= st["iters"] <= 12 and st["res"] <= 1e-3
    dev = lat._dU.device
    Ud = lat._dU.double()
    Yd = lat._dY.double()
    safe = torch.as_tensor(np.where(nbr < 0, 0, nbr), device=dev, dtype=torch.long)
    Wd = torch.as_tensor(W, device=dev).double()
    gath = torch.zeros_like(Ud)
    for t in range(nbr.shape[1]):
        gath += Wd[:, t, None] * Ud[safe[:, t]]
    lamG, lamC, lamQ = 1.0, 0.5, 4.0
    psid = torch.as_tensor(psi, device=dev).double()
    AU = Ud + (lamG + lamC + lamQ) * Ud - lamC * gath
    b = Yd + (lamG * Yd + lamQ * psid[None, :])  # first settle: U_in == Y, dt = 1
    res = float((b - AU).norm(dim=0).max().item())
    # the recurrence residual the solver reports and the true residual differ by the fp32 storage
    # floor of U: ~eps32 * ||A|| * max_c ||U_c||  (||A|| <= 1 + lamG + 2 lamC + lamQ = 7)
    floor = 2.0 * 1.19e-7 * 7.0 * float(Ud.norm(dim=0).max().item())
    assert res <= 1e-3 + floor
    assert abs(res - st["res"]) <= floor

    lat.set_receipt_detail("light")
    rec = lat.receipt()
    assert rec["deltaH_total"] >= 0.0
    assert rec["meta"]["ustar_converged"]


def test_n1m_sampled_rows_against_exhaustive_fp64_scan(api):
    """SURVEY 8(c) at the size of BASELINE configs[3]: N = 1M anchors; 32 random rows' mutual-kNN sets from an
    exhaustive fp64 scan of the sample and of its neighbours (tools/sampled_check.py, the checker bench.py's
    `parity_sample` uses) against the lattice's graph, plus the fp64 operator residual of U on a strided sample."""
    import torch

    from oscillink_b200.sharded_api import ShardedLattice
    from tools import sampled_check as sc

    N, D, k = 1_000_000, 64, 10
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    Y = torch.randn((N, D), generator=g, device="cuda")
    lat = ShardedLattice(Y, N, kneighbors=k)
    gs = torch.Generator()
    gs.manual_seed(1)
    sample = torch.randperm(N, generator=gs)[:32].sort().values.cuda()
    want, min_gap, hop = sc.mutual_sets(Y, 0, N, sample, k)
    assert sc.compare_neighbour_sets(lat._nbr[sample], want) == 0
    assert hop > 200
    psi = Y[:32].mean(dim=0)
    psi = psi / psi.norm()
    lat.set_query(psi.cpu().numpy())
    st = lat.settle(max_iters=12, tol=1e-3)
    assert st["res"] <= 1e-3
    rows = (1000 + 3907 * torch.arange(256, device="cuda")).clamp_(max=N - 1)
    r = sc.operator_residual_rows(rows, lambda i: lat._U[i], lambda i: lat._Y[i], lambda i: lat._Y[i], lat._nbr,
                                  lat._W, psi, (lat.lamG, lat.lamC, lat.lamQ), settle=True)
    est = float(torch.sqrt((r * r).sum(dim=0) * (N / r.shape[0])).max().item())
    floor = 2.0 * 1.19e-7 * 7.0 * float(lat._U.double().pow(2).sum(dim=0).max().sqrt().item())
    assert est <= 2.0 * (1e-3 + floor)
