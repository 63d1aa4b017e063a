"""GPU tests of the input-hardening fixes: wide ELL rows, bad psi / gates shapes, zero-iteration and
dt = 0 solves, malformed sparse state.  Everything goes through the reference-facing class."""
import numpy as np
import pytest

from oracle.sparse import SparseLattice
from tests.helpers import rel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    import torch

    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import oscillink_b200

    return oscillink_b200


def _inputs(N, D, seed=0):
    rs = np.random.RandomState(seed)
    Y = rs.randn(N, D).astype(np.float32)
    psi = Y[: min(32, N)].mean(axis=0)
    return Y, (psi / (np.linalg.norm(psi) + 1e-12)).astype(np.float32)


def test_kneighbors_128_settles_and_matches_oracle(api):
    """osc_knn_build documents k <= 128: at D = 384 the staged SpMM chunk then exceeds the 48 KB default
    of dynamic shared memory (the launch used to fail with 'invalid argument')."""
    Y, psi = _inputs(300, 384, seed=3)
    lat = api.OscillinkLattice(Y, kneighbors=128, deterministic_k=True)
    lat.set_query(psi)
    st = lat.settle(max_iters=12, tol=1e-3)
    o = SparseLattice(Y, k=128)
    o.set_query(psi)
    so = o.settle(max_iters=12, tol=1e-3)
    assert np.array_equal(lat._nbr.cpu().numpy(), o.nbr.astype(np.int32))
    assert abs(st["iters"] - so["iters"]) <= 1
    assert np.linalg.norm(lat.U - o.U) / np.linalg.norm(o.U) < 1e-5
    lat.set_receipt_detail("light")
    us, _, _ = o.stationary()
    assert rel(lat.receipt()["deltaH_total"], o.delta_h(us)) < 1e-5


def test_from_state_dense_adjacency_with_a_wide_row(api):
    """A caller-supplied dense A (lattice.py:709-713) whose hub row touches every node: ELL width N-1."""
    N, D = 700, 1024
    Y, psi = _inputs(N, D, seed=5)
    lat = api.OscillinkLattice(Y, kneighbors=6, deterministic_k=True)
    lat.set_query(psi)
    A = lat.A.copy()
    A[0, 1:] = np.maximum(A[0, 1:], 0.01)
    A[1:, 0] = A[0, 1:]
    st = lat.export_state()
    st["A"] = A.tolist()
    back = api.OscillinkLattice.from_state(st)
    assert back._nbr.shape[1] == N - 1
    s = back.settle(max_iters=60, tol=1e-4)
    assert np.isfinite(s["res"]) and 1 <= s["iters"] < 60
    # independent check of the solve: fp64 residual of (I + M) U = Y + RHS with the dense operator
    L = back.L_sym.astype(np.float64)
    M = np.eye(N) * (back.lamG + back.lamQ) + back.lamC * L
    rhs = back.lamG * Y.astype(np.float64) + back.lamQ * np.outer(np.ones(N), psi.astype(np.float64))
    r = (Y.astype(np.float64) + rhs) - (back.U.astype(np.float64) + M @ back.U.astype(np.float64))
    assert np.sqrt((r * r).sum(axis=0)).max() < 1e-3
    too_wide = api._cabi.load().osc_pcg_max_ell_width(D)
    assert too_wide >= N - 1


def test_wrong_length_psi_and_gates_raise_before_device_work(api):
    Y, psi = _inputs(64, 32)
    lat = api.OscillinkLattice(Y, kneighbors=5)
    for bad in (psi[:31], np.concatenate([psi, psi])):
        with pytest.raises(ValueError):
            lat.set_query(bad)
    with pytest.raises(ValueError):
        lat.set_query(psi, gates=np.ones(63, np.float32))
    with pytest.raises(ValueError):
        lat.set_query(psi, gates=np.ones((64, 2), np.float32))
    with pytest.raises(ValueError):
        lat.set_gates(np.ones((64, 2), np.float32))
    lat.set_query(psi.reshape(1, -1))  # a (1, D) row is accepted like the reference's broadcasting does
    assert lat.settle()["iters"] >= 1


def test_zero_iterations_leave_the_start_vector(api):
    """settle(max_iters=0): the state must be x0, never uninitialised memory."""
    Y, psi = _inputs(200, 32)
    lat = api.OscillinkLattice(Y, kneighbors=5)
    lat.set_query(psi)
    lat.settle(max_iters=0)
    assert np.array_equal(lat.U, Y)
    lat.settle(max_iters=3)
    U1 = lat.U.copy()
    lat.settle(max_iters=0)
    assert np.array_equal(lat.U, U1)


def test_batched_dt_zero_gives_finite_deltaH(api):
    from oscillink_b200 import BatchedLattices

    Y = np.stack([_inputs(96, 16, seed=s)[0] for s in range(2)])
    psi = np.stack([_inputs(96, 16, seed=s)[1] for s in range(2)])
    bl = BatchedLattices(Y, kneighbors=4)
    bl.set_query(psi)
    out = bl.settle(dt=0.0, receipt=True)
    dh = out["deltaH"].cpu().numpy()
    assert np.all(np.isfinite(dh))
    # dt = 0: (I + 0 M) U = U  ->  U stays Y and deltaH = <Y - U*, M (Y - U*)>
    assert np.array_equal(bl.U.cpu().numpy(), Y)
    for b in range(2):
        o = SparseLattice(Y[b], k=4)
        o.set_query(psi[b])
        us, _, _ = o.stationary()
        assert rel(float(dh[b]), o.delta_h(us)) < 1e-5


def test_sparse_state_with_scattered_padding_is_compacted(api):
    Y, psi = _inputs(120, 16, seed=9)
    lat = api.OscillinkLattice(Y, kneighbors=6, deterministic_k=True)
    lat.set_query(psi)
    st = lat.export_state(graph_format="ell")
    nbr = np.array(st["A_ell"]["nbr"])
    val = np.array(st["A_ell"]["val"], dtype=np.float32)
    # reverse every row: padding first, columns descending
    st2 = dict(st)
    st2["A_ell"] = {"nbr": nbr[:, ::-1].tolist(), "val": val[:, ::-1].tolist()}
    back = api.OscillinkLattice.from_state(st2)
    assert np.array_equal(back._nbr.cpu().numpy(), lat._nbr.cpu().numpy())
    assert back._signature() == lat._signature()
    a, b = lat.settle(), back.settle()
    assert a["iters"] == b["iters"] and np.allclose(lat.U, back.U, rtol=0, atol=1e-6)
    bad = dict(st)
    dup = nbr.copy()
    row = int(np.argmax((nbr >= 0).sum(axis=1)))
    dup[row, 1] = dup[row, 0]
    bad["A_ell"] = {"nbr": dup.tolist(), "val": val.tolist()}
    with pytest.raises(ValueError):
        api.OscillinkLattice.from_state(bad)
    bad["A_ell"] = {"nbr": np.where(nbr < 0, -2, nbr).tolist(), "val": val.tolist()}
    with pytest.raises(ValueError):
        api.OscillinkLattice.from_state(bad)


def _clustered(N, D, n_clusters, sigma, seed):
    rs = np.random.RandomState(seed)
    centers = rs.randn(n_clusters, D).astype(np.float32)
    centers /= np.linalg.norm(centers, axis=1, keepdims=True)
    lab = rs.randint(0, n_clusters, size=N)
    return (centers[lab] + sigma * rs.randn(N, D)).astype(np.float32)


def test_clustered_anchors_hand_over_to_the_3xtf32_engine(api):
    """64 tight clusters: almost every row has more than kc neighbours within the single-product engines'
    error bound (1.25e-3), so their candidate lists cannot be proven complete.  The build must NOT send
    those rows through the exhaustive kernel (N*D fp64 FMAs per row): it re-runs the candidate pass with the
    3xTF32 engine, and the graph is still the canonical one."""
    import time

    import torch

    from oracle.sparse import assemble, normalise_rows, topk_canonical
    from oscillink_b200.sharded_api import ShardedLattice

    N, D, k = 20000, 64, 10
    Y = _clustered(N, D, 64, 0.01, seed=21)
    idx, sim, gap = topk_canonical(normalise_rows(Y), k)
    want = assemble(idx, sim, 1.0)[0].astype(np.int32)
    torch.cuda.synchronize()
    t0 = time.time()
    lat = api.OscillinkLattice(Y, kneighbors=k, deterministic_k=True)
    torch.cuda.synchronize()
    t_class = time.time() - t0
    assert np.array_equal(lat._nbr.cpu().numpy(), want)
    t0 = time.time()
    sl = ShardedLattice(Y, N, kneighbors=k)
    torch.cuda.synchronize()
    t_sharded = time.time() - t0
    assert np.array_equal(sl._nbr.cpu().numpy(), want)
    assert "fallback" in sl.engine_used, sl.engine_used
    # an exhaustive scan of every row would be N^2*D = 2.6e10 fp64 FMAs plus 20000 serial list merges
    assert t_class < 5.0 and t_sharded < 5.0, (t_class, t_sharded)


def test_clustered_batch_is_rebuilt_after_the_fact(api):
    """BatchedLattices bounds the exhaustive path on the device (no host sync in the build) and rebuilds
    with the 3xTF32 engine when the bound was exceeded: strict settle and the pipelined host call."""
    import torch

    from oracle.sparse import SparseLattice
    from oscillink_b200 import BatchedLattices, settle_host_batch

    B, N, D, k = 3, 1000, 32, 8
    Y = np.stack([_clustered(N, D, 16, 0.01, seed=40 + b) for b in range(B)])
    psi = np.stack([Y[b, :32].mean(axis=0) / (np.linalg.norm(Y[b, :32].mean(axis=0)) + 1e-12) for b in range(B)])
    psi = psi.astype(np.float32)
    bl = BatchedLattices(Y, kneighbors=k)
    bl.set_query(psi)
    out = bl.settle(receipt=True)
    assert "fallback" in bl.engine_used, bl.engine_used
    got = settle_host_batch(torch.from_numpy(Y).pin_memory(), torch.from_numpy(psi).pin_memory(), kneighbors=k,
                            chunk=2).numpy()
    for b in range(B):
        o = SparseLattice(Y[b], k=k)
        o.set_query(psi[b])
        st = o.settle()
        us, it, _ = o.stationary()
        assert np.array_equal(bl.nbr[b].cpu().numpy(), o.nbr.astype(np.int32))
        assert int(out["iters"][b].item()) == st["iters"] == int(got[b, 0])
        assert int(out["ustar_iters"][b].item()) == it == int(got[b, 2])
        assert rel(float(out["deltaH"][b].item()), o.delta_h(us)) < 1e-5
        assert rel(float(got[b, 4]), o.delta_h(us)) < 1e-5
