"""GPU parity of the serving-batch path (osc_batched_settle) against the per-lattice class,
the sparse oracle and the reference golden values."""
import numpy as np
import pytest

from oracle import cases
from oracle.sparse import SparseLattice
from tests.helpers import load_golden, rel

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _inputs(B, N, D, seed0=0):
    Ys, psis = [], []
    for b in range(B):
        rs = np.random.RandomState(seed0 + b)
        Y = rs.randn(N, D).astype(np.float32)
        psi = Y[: min(32, N)].mean(axis=0)
        psis.append((psi / (np.linalg.norm(psi) + 1e-12)).astype(np.float32))
        Ys.append(Y)
    return np.stack(Ys), np.stack(psis)


@pytest.mark.parametrize("B,N,D,k", [(5, 300, 64, 6), (3, 1200, 384, 8), (4, 97, 20, 5), (2, 640, 128, 8)])
def test_batched_matches_oracle(B, N, D, k):
    import torch

    from oscillink_b200 import BatchedLattices

    Y, psi = _inputs(B, N, D)
    bl = BatchedLattices(Y, kneighbors=k)
    bl.set_query(psi)
    assert bl.supported()
    out = bl.settle(max_iters=12, tol=1e-3, receipt=True, keep_ustar=True)
    torch.cuda.synchronize()
    nbr = bl.nbr.cpu().numpy()
    U = bl.U.cpu().numpy()
    Us = bl.Ustar.cpu().numpy()
    for b in range(B):
        o = SparseLattice(Y[b], k=k)
        o.set_query(psi[b])
        assert np.array_equal(nbr[b], o.nbr.astype(np.int32))
        st = o.settle(max_iters=12, tol=1e-3)
        assert abs(int(out["iters"][b].item()) - st["iters"]) <= 1
        ous, it, res = o.stationary()
        assert abs(int(out["ustar_iters"][b].item()) - it) <= 1
        assert np.linalg.norm(U[b] - o.U) / np.linalg.norm(o.U) < TOL
        assert np.linalg.norm(Us[b] - ous) / np.linalg.norm(ous) < TOL
        assert rel(float(out["deltaH"][b].item()), o.delta_h(ous)) < TOL
        if int(out["iters"][b].item()) == st["iters"]:
            assert rel(float(out["res"][b].item()), st["res"]) < 1e-4


def test_batched_lattice0_is_config2_golden():
    """Lattice b of the serving batch uses RandomState(b); b = 0 is BASELINE config #2."""
    from oscillink_b200 import BatchedLattices

    g, z = load_golden("config2_1200")
    Y, psi = _inputs(2, 1200, 384)
    bl = BatchedLattices(Y, kneighbors=8)
    bl.set_query(psi)
    out = bl.settle(max_iters=12, tol=1e-3, receipt=True)
    assert np.array_equal(bl.nbr[0].cpu().numpy(), z["nbr"])
    assert int(out["iters"][0].item()) == g["settle"]["iters"]
    assert int(out["ustar_iters"][0].item()) == g["ustar"]["iters"]
    assert rel(float(out["res"][0].item()), g["settle"]["res"]) < 1e-4
    assert rel(float(out["deltaH"][0].item()), g["deltaH"]) < TOL
    assert int(bl.nnz[0].item()) == g["nnz"]


def test_batched_with_gates_and_second_settle():
    """Non-uniform gates; a second settle warm-starts from the first one's U."""
    from oscillink_b200 import BatchedLattices

    B, N, D, k = 3, 256, 32, 6
    Y, psi = _inputs(B, N, D, seed0=40)
    gates = np.random.RandomState(1).uniform(0.2, 1.0, size=(B, N)).astype(np.float32)
    bl = BatchedLattices(Y, kneighbors=k, lamG=0.9, lamC=0.7, lamQ=3.0, row_cap_val=0.8)
    bl.set_query(psi, gates)
    bl.settle(dt=0.5, max_iters=3, tol=1e-9, receipt=False)
    out = bl.settle(dt=1.0, max_iters=12, tol=1e-3, receipt=True)
    U = bl.U.cpu().numpy()
    for b in range(B):
        o = SparseLattice(Y[b], k=k, cap=0.8, lamG=0.9, lamC=0.7, lamQ=3.0)
        o.set_query(psi[b], gates[b])
        o.settle(dt=0.5, max_iters=3, tol=1e-9)
        st = o.settle(dt=1.0, max_iters=12, tol=1e-3)
        assert abs(int(out["iters"][b].item()) - st["iters"]) <= 1
        assert np.linalg.norm(U[b] - o.U) / np.linalg.norm(o.U) < TOL
        ous, _, _ = o.stationary()
        assert rel(float(out["deltaH"][b].item()), o.delta_h(ous)) < 1e-4


def _uneven_inputs(B, N, D, seed0=70):
    """Columns 0-7 carry almost no signal, so their slabs satisfy the stop test iterations before
    the rest of the lattice does: exercises the resolve + re-run pass of the slab kernel."""
    Y, _ = _inputs(B, N, D, seed0)
    Y[:, :, :8] *= 1e-4
    psi = np.zeros((B, D), dtype=np.float32)
    psi[:, 8:] = Y[:, :16, 8:].mean(axis=1)
    psi /= np.linalg.norm(psi, axis=1, keepdims=True) + 1e-12
    return Y, psi.astype(np.float32)


def test_batched_uneven_slabs_are_rerun_to_lattice_count():
    import torch

    from oscillink_b200 import BatchedLattices

    B, N, D, k = 3, 320, 48, 6
    Y, psi = _uneven_inputs(B, N, D)
    bl = BatchedLattices(Y, kneighbors=k)
    bl.set_query(psi)
    out = bl.settle(max_iters=12, tol=1e-3, receipt=True, keep_ustar=True)
    torch.cuda.synchronize()
    flags = out["unresolved"].cpu().numpy()
    assert (flags & 2).all(), "expected the early slabs to be re-run"
    assert not (flags & 1).any()
    U, Us = bl.U.cpu().numpy(), bl.Ustar.cpu().numpy()
    for b in range(B):
        o = SparseLattice(Y[b], k=k)
        o.set_query(psi[b])
        st = o.settle(max_iters=12, tol=1e-3)
        ous, it, res = o.stationary()
        assert int(out["iters"][b].item()) == st["iters"]
        assert int(out["ustar_iters"][b].item()) == it
        assert rel(float(out["res"][b].item()), st["res"]) < 1e-4
        assert np.linalg.norm(U[b] - o.U) / np.linalg.norm(o.U) < TOL
        # the quiet columns must carry exactly the lattice's iteration count as well
        assert np.linalg.norm(U[b][:, :8] - o.U[:, :8]) / np.linalg.norm(o.U[:, :8]) < 1e-4
        assert np.linalg.norm(Us[b] - ous) / np.linalg.norm(ous) < TOL
        assert rel(float(out["deltaH"][b].item()), o.delta_h(ous)) < TOL


def test_batched_unresolved_falls_back_to_global_pcg(monkeypatch):
    """Host fallback (osc_pcg_solve per flagged lattice), forced through the test hook."""
    from oscillink_b200 import BatchedLattices

    B, N, D, k = 2, 200, 32, 5
    Y, psi = _inputs(B, N, D, seed0=90)
    ref = BatchedLattices(Y, kneighbors=k)
    ref.set_query(psi)
    want = ref.settle(receipt=True)
    monkeypatch.setenv("OSC_BATCHED_FORCE_UNRESOLVED", "1")
    bl = BatchedLattices(Y, kneighbors=k)
    bl.set_query(psi)
    lazy = bl.settle(receipt=True, strict=False)
    assert (lazy["unresolved"].cpu().numpy() & 1).all()
    assert bl.resolve_flagged(lazy) == B
    assert not (lazy["unresolved"].cpu().numpy() & 1).any()
    for key in ("iters", "ustar_iters"):
        assert np.array_equal(lazy[key].cpu().numpy(), want[key].cpu().numpy())
    assert np.allclose(lazy["deltaH"].cpu().numpy(), want["deltaH"].cpu().numpy(), rtol=1e-5)
    U0, U1 = ref.U.cpu().numpy(), bl.U.cpu().numpy()
    assert np.linalg.norm(U0 - U1) / np.linalg.norm(U0) < TOL


def test_settle_host_batch_pipelined_matches_resident():
    import torch

    from oscillink_b200 import BatchedLattices, settle_host_batch

    B, N, D, k = 7, 160, 32, 6
    Y, psi = _inputs(B, N, D, seed0=120)
    Yh = torch.from_numpy(Y).pin_memory()
    ph = torch.from_numpy(psi).pin_memory()
    got = settle_host_batch(Yh, ph, kneighbors=k, chunk=3).numpy()
    bl = BatchedLattices(Y, kneighbors=k)
    bl.set_query(psi)
    out = bl.settle(receipt=True)
    assert np.array_equal(got[:, 0], out["iters"].cpu().numpy().astype(np.float64))
    assert np.array_equal(got[:, 2], out["ustar_iters"].cpu().numpy().astype(np.float64))
    assert np.allclose(got[:, 4], out["deltaH"].cpu().numpy(), rtol=1e-12)


# ------------------------------------------------------------------ multi-shift fast path (batched_ms.cu)
@pytest.mark.parametrize("kw", [
    dict(dt=1.0, max_iters=12, tol=1e-3, ustar_tol=1e-4),          # the serving call
    dict(dt=0.5, max_iters=12, tol=1e-3, ustar_tol=1e-4),          # sigma = 2
    dict(dt=1.0, max_iters=12, tol=1e-6, ustar_tol=1e-2),          # the stationary system stops FIRST
    dict(dt=2.0, max_iters=2, tol=1e-9, ustar_tol=1e-4),           # settle capped by max_iters
    dict(dt=1.0, max_iters=12, tol=1e-3, ustar_tol=1e-9, ustar_max_iters=3),  # U* capped by max_iters
    dict(dt=1.0, max_iters=12, tol=1e-4, ustar_tol=1e-4),          # both stop at (about) the same iteration
])
def test_multishift_matches_two_solve_oracle(kw):
    """settle + U* + deltaH from ONE multi-shift CG (first settle, uniform gates) against the oracle's two
    separate PCG solves: same iteration counts, U / U* / deltaH within 1e-5."""
    import torch

    from oscillink_b200 import BatchedLattices

    B, N, D, k = 3, 500, 40, 7
    Y, psi = _inputs(B, N, D, seed0=200)
    bl = BatchedLattices(Y, kneighbors=k, lamG=1.1, lamC=0.6, lamQ=3.5)
    bl.set_query(psi)
    out = bl.settle(receipt=True, keep_ustar=True, **kw)
    torch.cuda.synchronize()
    U, Us = bl.U.cpu().numpy(), bl.Ustar.cpu().numpy()
    for b in range(B):
        o = SparseLattice(Y[b], k=k, lamG=1.1, lamC=0.6, lamQ=3.5)
        o.set_query(psi[b])
        st = o.settle(dt=kw["dt"], max_iters=kw["max_iters"], tol=kw["tol"])
        ous, it, res = o.stationary(tol=kw["ustar_tol"], max_iters=kw.get("ustar_max_iters", 64))
        assert int(out["iters"][b].item()) == st["iters"]
        assert int(out["ustar_iters"][b].item()) == it
        assert rel(float(out["res"][b].item()), st["res"]) < 1e-4
        assert rel(float(out["ustar_res"][b].item()), res) < 1e-4
        assert np.linalg.norm(U[b] - o.U) / np.linalg.norm(o.U) < TOL
        assert np.linalg.norm(Us[b] - ous) / np.linalg.norm(ous) < TOL
        assert rel(float(out["deltaH"][b].item()), o.delta_h(ous)) < TOL


@pytest.mark.parametrize("N,D,k", [(1200, 64, 8), (1210, 16, 12), (1280, 16, 16), (130, 8, 3), (1000, 24, 9)])
def test_multishift_equals_two_solve_kernel(N, D, k, monkeypatch):
    """The fast path and the two-solve kernel (OSC_BATCHED_MS=0) are the same recurrences up to fp32
    rounding: every block size / ELL width variant, pad rows included."""
    import torch

    from oscillink_b200 import BatchedLattices

    B = 2
    Y, psi = _inputs(B, N, D, seed0=300)
    outs = []
    for ms in ("1", "0"):
        monkeypatch.setenv("OSC_BATCHED_MS", ms)
        bl = BatchedLattices(Y, kneighbors=k)
        bl.set_query(psi)
        out = bl.settle(receipt=True, keep_ustar=True)
        torch.cuda.synchronize()
        outs.append((bl.U.cpu().numpy(), bl.Ustar.cpu().numpy(), out["iters"].cpu().numpy(),
                     out["ustar_iters"].cpu().numpy(), out["deltaH"].cpu().numpy(), out["res"].cpu().numpy()))
    a, b = outs
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
    assert np.linalg.norm(a[0] - b[0]) / np.linalg.norm(b[0]) < 2e-6
    assert np.linalg.norm(a[1] - b[1]) / np.linalg.norm(b[1]) < 2e-6
    assert np.allclose(a[4], b[4], rtol=1e-5)
    assert np.allclose(a[5], b[5], rtol=1e-4)


def test_multishift_uneven_slabs_forced_counts(monkeypatch):
    """The fix pass of the fast path: slabs that stop early are re-run with forced (T_s, T_u)."""
    import torch

    from oscillink_b200 import BatchedLattices

    B, N, D, k = 2, 320, 48, 6
    Y, psi = _uneven_inputs(B, N, D)
    res = {}
    for ms in ("1", "0"):
        monkeypatch.setenv("OSC_BATCHED_MS", ms)
        bl = BatchedLattices(Y, kneighbors=k)
        bl.set_query(psi)
        out = bl.settle(receipt=True, keep_ustar=True)
        torch.cuda.synchronize()
        assert (out["unresolved"].cpu().numpy() & 2).all()
        assert not (out["unresolved"].cpu().numpy() & 1).any()
        res[ms] = (bl.U.cpu().numpy(), bl.Ustar.cpu().numpy(), out["iters"].cpu().numpy(),
                   out["ustar_iters"].cpu().numpy(), out["deltaH"].cpu().numpy())
    assert np.array_equal(res["1"][2], res["0"][2]) and np.array_equal(res["1"][3], res["0"][3])
    for i in (0, 1):
        assert np.linalg.norm(res["1"][i] - res["0"][i]) / np.linalg.norm(res["0"][i]) < 2e-6
    assert np.allclose(res["1"][4], res["0"][4], rtol=1e-5)


@pytest.mark.parametrize("N,D,k", [(1200, 384, 8), (500, 64, 6), (300, 36, 12)])
def test_duplicate_free_rescoring_is_bit_identical(N, D, k, monkeypatch):
    """knn_rescore_owned_kernel + knn_rescore_rank_kernel (every mutual candidate pair scored once) against
    knn_rescore_kernel (every pair scored by both rows): identical neighbour tables, weights and gaps."""
    import torch

    from oscillink_b200 import BatchedLattices

    Y, psi = _inputs(3, N, D, seed0=500)
    Y[1, 7] = Y[1, 3]          # duplicate anchors: exact ties, handled by the canonical (sim desc, index asc) rule
    res = []
    for flag in ("1", "0"):
        monkeypatch.setenv("OSC_RESCORE_DEDUP", flag)
        bl = BatchedLattices(Y, kneighbors=k)
        torch.cuda.synchronize()
        res.append((bl.nbr.cpu().numpy(), bl.A.cpu().numpy(), bl.W.cpu().numpy(), bl.gap.cpu().numpy(),
                    int(bl.n_exhaustive.item())))
    for a, b in zip(res[0][:4], res[1][:4]):
        assert np.array_equal(a, b)
    assert res[0][4] == res[1][4]


@pytest.mark.parametrize("N,D,k,switch", [(1200, 384, 8, "OSC_ASSEMBLE_REG"), (500, 64, 4, "OSC_ASSEMBLE_REG"),
                                          (300, 36, 12, "OSC_ASSEMBLE_REG"), (1200, 384, 8, "OSC_RESCORE_RANK16"),
                                          (700, 128, 5, "OSC_RESCORE_RANK16")])
def test_register_resident_build_kernels_are_bit_identical(N, D, k, switch, monkeypatch):
    """assemble_mutual_reg_kernel (k in 4/8/12/16: mutual filter with the lists in registers) against the generic
    assemble_mutual_kernel, and knn_rescore_rank16_kernel (half a warp per row, shuffles) against the
    shared-memory ranking pass: identical graphs, weights, degrees and gaps."""
    import torch

    from oscillink_b200 import BatchedLattices

    Y, psi = _inputs(3, N, D, seed0=900)
    Y[2, 11] = Y[2, 5]
    res = []
    for flag in ("1", "0"):
        monkeypatch.setenv(switch, flag)
        bl = BatchedLattices(Y, kneighbors=k)
        torch.cuda.synchronize()
        res.append((bl.nbr.cpu().numpy(), bl.A.cpu().numpy(), bl.W.cpu().numpy(), bl.gap.cpu().numpy(),
                    bl.deg.cpu().numpy(), bl.sqrt_deg.cpu().numpy()))
    for a, b in zip(res[0], res[1]):
        assert np.array_equal(a, b)


def test_register_resident_packer_gives_the_same_graph_image(monkeypatch):
    """batched_pack8_kernel (k <= 8: bank-residue scheduling of the neighbour slots with the row entries and
    the residue counts in registers) against the generic batched_pack_kernel: the slot order fixes the order of
    the gather sums, so identical images <=> bit-identical settles."""
    import torch

    from oscillink_b200 import BatchedLattices

    Y, psi = _inputs(4, 1200, 384, seed0=1300)
    res = []
    for flag in ("1", "0"):
        monkeypatch.setenv("OSC_BATCHED_PACK8", flag)
        bl = BatchedLattices(Y, kneighbors=8)
        bl.set_query(psi)
        out = bl.settle(max_iters=12, tol=1e-3, receipt=True, keep_ustar=True)
        torch.cuda.synchronize()
        res.append((bl.U.cpu().numpy(), bl.Ustar.cpu().numpy(), out["deltaH"].cpu().numpy(),
                    out["iters"].cpu().numpy(), out["res"].cpu().numpy()))
    for a, b in zip(res[0], res[1]):
        assert np.array_equal(a, b)


def test_settle_host_batch_returns_the_settled_state():
    """settle_host_batch(U_host=...): the settled U of every lattice comes back through a third stream."""
    import torch

    from oscillink_b200 import BatchedLattices, settle_host_batch

    B, N, D, k = 5, 200, 32, 6
    Y, psi = _inputs(B, N, D, seed0=700)
    Yh, ph = torch.from_numpy(Y).pin_memory(), torch.from_numpy(psi).pin_memory()
    Uh = torch.empty((B, N, D), dtype=torch.float32).pin_memory()
    got = settle_host_batch(Yh, ph, kneighbors=k, chunk=2, U_host=Uh).numpy()
    bl = BatchedLattices(Y, kneighbors=k)
    bl.set_query(psi)
    out = bl.settle(receipt=True)
    assert np.array_equal(got[:, 0], out["iters"].cpu().numpy().astype(np.float64))
    assert np.array_equal(Uh.numpy(), bl.U.cpu().numpy())
