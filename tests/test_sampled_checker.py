"""The sampled-row checker used at N = 1M / 10M (tools/sampled_check.py: bench.py `parity_sample`,
tests/test_gpu_large.py) is itself checked here, on CPU, against the sparse oracle that the reference's
golden vectors pin (tests/test_oracle_golden.py): same canonical top-k, same mutual sets."""
import numpy as np
import torch

from oracle.sparse import SparseLattice, normalise_rows, topk_canonical
from tools import sampled_check as sc


def _data(N, D, seed):
    rs = np.random.RandomState(seed)
    Y = rs.randn(N, D).astype(np.float32)
    Y[17] = Y[5]  # duplicate anchors: exact ties, decided by (similarity desc, index asc)
    return Y


def test_exact_topk_is_the_oracles_canonical_topk():
    N, D, k = 1500, 24, 6
    Y = _data(N, D, 3)
    rows = np.arange(0, N, 7)
    idx_o, sim_o, _ = topk_canonical(normalise_rows(Y), k, rows=rows)
    idx, sim, gap = sc.exact_topk(torch.from_numpy(Y), 0, N, torch.from_numpy(rows), k, chunk=400)
    assert np.array_equal(idx.numpy(), idx_o)
    np.testing.assert_allclose(sim.numpy(), sim_o, rtol=3e-7, atol=1e-9)
    assert float(gap.min()) >= 0.0


def test_mutual_sets_are_the_oracles_graph_rows():
    N, D, k = 1500, 24, 6
    Y = _data(N, D, 4)
    o = SparseLattice(Y, k=k)
    sample = torch.arange(0, N, 11)
    want, min_gap, hop = sc.mutual_sets(torch.from_numpy(Y), 0, N, sample, k, chunk=512)
    assert hop > sample.numel()
    assert sc.compare_neighbour_sets(torch.from_numpy(o.nbr[sample.numpy()].astype(np.int64)), want) == 0
    # and the comparison does notice a wrong row
    bad = o.nbr[sample.numpy()].astype(np.int64).copy()
    bad[3, 0] = (bad[3, 0] + 1) % N
    assert sc.compare_neighbour_sets(torch.from_numpy(bad), want) == 1


def test_operator_residual_rows_vanish_at_the_oracles_solution():
    N, D, k = 800, 16, 5
    Y = _data(N, D, 5)
    psi = Y[:32].mean(axis=0)
    psi = (psi / np.linalg.norm(psi)).astype(np.float32)
    o = SparseLattice(Y, k=k)
    o.set_query(psi)
    us, it, res = o.stationary(tol=1e-7, max_iters=200)
    rows = torch.arange(0, N, 9)
    Yt, Ut = torch.from_numpy(Y), torch.from_numpy(np.asarray(us, dtype=np.float32))
    r = sc.operator_residual_rows(rows, lambda i: Ut[i], lambda i: Yt[i], lambda i: Yt[i],
                                  torch.from_numpy(o.nbr.astype(np.int64)), torch.from_numpy(o.W.astype(np.float32)),
                                  torch.from_numpy(psi), (o.lamG, o.lamC, o.lamQ), settle=False)
    assert float(r.abs().max()) < 5e-5
