"""CPU model of the packed-key top-k lists of csrc/knn_tc.cu (pk_key / pk_score / pk_col / pk_insert).

The kernel keeps (order-preserving integer image of the score | column) in ONE register per list
entry and inserts with a min/max network.  This file restates that arithmetic in NumPy and pins the
properties the completeness check of osc_knn_rescore_checked relies on (include/oscillink_b200.h,
OSC_KNN_EPS_TC1): keys order by (truncated score desc, column asc), the truncation costs at most
2^-12 relative, and the network keeps exactly the KC largest keys of a stream, sorted."""
import numpy as np

COL_BITS = 11
COL_MASK = (1 << COL_BITS) - 1


def pk_key(s, col):
    b = np.asarray(s, dtype=np.float32).view(np.int32).astype(np.int64)
    t = b ^ ((b >> 31) & 0x7FFFFFFF)          # arithmetic shift: all-ones for negative floats
    t = ((t + 2**31) % 2**32) - 2**31           # stay in int32 range
    return (t & ~COL_MASK) | (COL_MASK - np.asarray(col, dtype=np.int64))


def pk_score(key):
    t = np.asarray(key, dtype=np.int64) & ~COL_MASK
    b = t ^ ((t >> 31) & 0x7FFFFFFF)
    return (((b + 2**31) % 2**32) - 2**31).astype(np.int32).view(np.float32)


def pk_col(key):
    return COL_MASK - (np.asarray(key, dtype=np.int64) & COL_MASK)


def pk_insert(keys, k):
    old = keys.copy()
    for q in range(len(keys) - 1, 0, -1):
        keys[q] = max(old[q], min(old[q - 1], k))
    keys[0] = max(old[0], k)


def test_key_roundtrip_and_truncation_bound():
    rs = np.random.RandomState(0)
    s = np.concatenate([rs.uniform(-1, 1, 20000), rs.normal(0, 0.05, 20000), [0.0, 1.0, -1.0, 1e-6, -1e-6]])
    s = s.astype(np.float32)
    col = rs.randint(0, 2048, size=s.size)
    key = pk_key(s, col)
    assert np.array_equal(pk_col(key), col)
    back = pk_score(key)
    assert np.all(np.abs(back.astype(np.float64) - s.astype(np.float64)) <= 2.0**-12 * np.abs(s) + 1e-45)
    pos = s > 0
    assert np.all(back[pos] <= s[pos])


def test_key_order_is_score_then_column():
    rs = np.random.RandomState(1)
    s = rs.uniform(-1, 1, 4000).astype(np.float32)
    col = rs.randint(0, 2048, size=s.size)
    key = pk_key(s, col)
    i, j = rs.randint(0, s.size, size=(2, 200000))
    far = np.abs(s[i].astype(np.float64) - s[j]) > 2.0**-11 * np.maximum(np.abs(s[i]), np.abs(s[j]))
    lt = s[i] < s[j]
    assert np.all(key[i][far & lt] < key[j][far & lt])          # order follows the score beyond the truncation
    same = pk_key(np.float32(0.125), np.arange(2048))
    assert np.all(np.diff(same) < 0)                             # equal scores: the smaller column wins
    assert np.all(pk_key(np.float32(-0.3), 5) < pk_key(np.float32(-0.2), 2047))
    assert np.all(pk_key(np.float32(-1e-3), 0) < pk_key(np.float32(1e-3), 2047))
    assert np.all(pk_key(np.float32(-1.0), 0) > -2**31)          # INT_MIN stays free for "empty"


def test_minmax_network_keeps_the_largest_keys_sorted():
    rs = np.random.RandomState(2)
    for kc in (12, 16):
        for _ in range(20):
            n = 300
            s = rs.normal(0, 0.05, n).astype(np.float32)
            keys_in = pk_key(s, rs.permutation(2048)[:n])
            lst = np.full(kc, -2**31, dtype=np.int64)
            for k in keys_in:
                pk_insert(lst, int(k))
            want = np.sort(keys_in)[::-1][:kc]
            assert np.array_equal(lst, want)
