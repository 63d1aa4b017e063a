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
            assert rel(float(out["res"][b].item()), st["res"]) < 1e-3


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
    assert rel(float(out["res"][0].item()), g["settle"]["res"]) < 1e-3
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
