"""osc_chain_build (host function of the C ABI, a6: oscillink/core/graph.py:96-111 via
lattice.py:129-149) against the oracle's sparse restatement -- bit-exact in fp32, no GPU needed."""
import ctypes as C

import numpy as np
import pytest

from oracle.sparse import chain_rows
from oscillink_b200 import _cabi


def _build(chain, weights, N):
    lib = _cabi.load()
    ch = np.ascontiguousarray(np.asarray(chain, dtype=np.int32))
    w = None if weights is None else np.ascontiguousarray(np.asarray(weights, dtype=np.float32))
    n_rows, nnz = C.c_int32(0), C.c_int32(0)
    _cabi.check(lib.osc_chain_build_size(ch.ctypes.data, len(ch), N, C.byref(n_rows), C.byref(nnz)))
    rows = np.empty(n_rows.value, np.int32)
    rowptr = np.empty(n_rows.value + 1, np.int32)
    col = np.empty(nnz.value, np.int32)
    wp = np.empty(nnz.value, np.float32)
    ap = np.empty(nnz.value, np.float32)
    slot = np.empty(N, np.int32)
    _cabi.check(lib.osc_chain_build(ch.ctypes.data, len(ch), None if w is None else w.ctypes.data, N,
                                    rows.ctypes.data, rowptr.ctypes.data, col.ctypes.data, wp.ctypes.data,
                                    ap.ctypes.data, slot.ctypes.data))
    return rows, rowptr, col, wp, ap, slot


CASES = [
    ([2, 5, 7, 9], None, 12),                                   # examples/quickstart.py
    (list(range(8)), None, 400),                                # scripts/benchmark.py:60
    ([0, 1, 2, 1, 3, 0], [0.5, 2.0, 0.25, 1.5, 3.0], 6),         # revisits: max-merge of (1,2)/(2,1)
    ([4, 4, 1], [0.7, 0.3], 5),                                 # a self edge
    ([7, 3], [1e-20], 8),                                       # degree below the 1e-12 clamp
    (list(np.random.RandomState(0).randint(0, 50, size=200)), list(np.random.RandomState(1).rand(199)), 50),
]


@pytest.mark.parametrize("chain,weights,N", CASES)
def test_chain_csr_equals_the_oracle(chain, weights, N):
    rows, rowptr, col, wp, ap, slot = _build(chain, weights, N)
    wp_o, ap_o = chain_rows(N, chain, weights)
    assert rows.tolist() == sorted(wp_o)
    assert np.array_equal(np.nonzero(slot >= 0)[0], rows) and np.array_equal(slot[rows], np.arange(len(rows)))
    assert rowptr[0] == 0 and rowptr[-1] == len(col)
    for r, u in enumerate(rows.tolist()):
        cols = col[rowptr[r]:rowptr[r + 1]].tolist()
        assert cols == sorted(wp_o[u])
        for e, v in zip(range(rowptr[r], rowptr[r + 1]), cols):
            assert ap[e] == ap_o[u][v]            # bit-exact fp32
            assert wp[e] == wp_o[u][v]
    # the path graph is symmetric up to the rounding order of (Ap * 1/sd_u) * 1/sd_v (graph.py:90-91)
    d = {(int(u), int(col[e])): float(wp[e]) for r, u in enumerate(rows) for e in range(rowptr[r], rowptr[r + 1])}
    assert all(abs(d[(v, u)] - w) <= 2e-7 * abs(w) for (u, v), w in d.items())


def test_chain_validation_errors():
    lib = _cabi.load()
    n, z = C.c_int32(0), C.c_int32(0)
    one = np.array([3], np.int32)
    with pytest.raises(ValueError):
        _cabi.check(lib.osc_chain_build_size(one.ctypes.data, 1, 10, C.byref(n), C.byref(z)))
    bad = np.array([1, 10], np.int32)
    with pytest.raises(ValueError):
        _cabi.check(lib.osc_chain_build_size(bad.ctypes.data, 2, 10, C.byref(n), C.byref(z)))
    neg = np.array([-1, 2], np.int32)
    with pytest.raises(ValueError):
        _cabi.check(lib.osc_chain_build_size(neg.ctypes.data, 2, 10, C.byref(n), C.byref(z)))
