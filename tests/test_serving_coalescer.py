"""Row f4 host logic: the request coalescer's grouping policy with an injected backend (no GPU),
plus the canonical state signature against the reference's own value."""
import threading
import time

import numpy as np
import pytest

from oscillink_b200.serving import SettleCoalescer, state_signature
from tests.helpers import load_golden


def _fake_backend(log):
    def run(group):
        log.append([r.group_key() for r in group])
        time.sleep(0.01)
        return [{"settle": {"iters": 1, "res": 0.0}, "n": r.Y.shape[0], "batch": len(group)} for r in group]
    return run


def test_requests_with_equal_shape_are_coalesced_and_others_kept_apart():
    log = []
    with SettleCoalescer(max_batch=8, max_wait_ms=200.0, backend=_fake_backend(log)) as co:
        Ya, Yb = np.zeros((10, 4), np.float32), np.zeros((12, 4), np.float32)
        futs = [co.submit(Ya, np.zeros(4, np.float32), kneighbors=3) for _ in range(5)]
        futs += [co.submit(Yb, np.zeros(4, np.float32), kneighbors=3) for _ in range(3)]
        futs += [co.submit(Ya, np.zeros(4, np.float32), kneighbors=3, tol=1e-4)]
        outs = [f.result(timeout=10) for f in futs]
    assert [o["n"] for o in outs] == [10] * 5 + [12] * 3 + [10]
    assert sorted(len(g) for g in log) == [1, 3, 5]
    for g in log:
        assert len(set(g)) == 1  # one key per backend call
    assert co.stats["requests"] == 9 and co.stats["batches"] == 3 and co.stats["max_batch_seen"] == 5


def test_max_batch_splits_a_burst_and_order_is_fifo_per_key():
    log = []
    with SettleCoalescer(max_batch=4, max_wait_ms=100.0, backend=_fake_backend(log)) as co:
        Y = np.zeros((6, 4), np.float32)
        futs = [co.submit(Y, np.full(4, float(i), np.float32)) for i in range(10)]
        outs = [f.result(timeout=10) for f in futs]
    assert [o["batch"] for o in outs] == [4] * 4 + [4] * 4 + [2] * 2
    assert [len(g) for g in log] == [4, 4, 2]


def test_concurrent_request_threads_share_batches():
    log = []
    co = SettleCoalescer(max_batch=64, max_wait_ms=150.0, backend=_fake_backend(log))
    outs = [None] * 16
    start = threading.Barrier(16)

    def worker(i):
        start.wait()
        outs[i] = co.settle(np.zeros((9, 8), np.float32), np.zeros(8, np.float32), kneighbors=4)

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(16)]
    [t.start() for t in ts]
    [t.join(timeout=20) for t in ts]
    co.close()
    assert all(o is not None for o in outs)
    assert sum(len(g) for g in log) == 16 and len(log) <= 3  # a 150 ms window catches the burst


def test_validation_errors_are_raised_on_the_submitting_thread():
    with SettleCoalescer(backend=_fake_backend([])) as co:
        with pytest.raises(ValueError):
            co.submit(np.zeros(4, np.float32), np.zeros(4, np.float32))
        with pytest.raises(ValueError):
            co.submit(np.zeros((4, 4), np.float32), np.zeros(3, np.float32))
        with pytest.raises(ValueError):
            co.submit(np.zeros((4, 4), np.float32), np.zeros(4, np.float32), kneighbors=0)
        with pytest.raises(ValueError):
            co.submit(np.zeros((4, 4), np.float32), np.zeros(4, np.float32), lamG=0.0)
        with pytest.raises(ValueError):
            co.submit(np.zeros((4, 4), np.float32), np.zeros(4, np.float32), gates=np.ones(3, np.float32))


def test_backend_failure_releases_every_waiter():
    def boom(group):
        raise RuntimeError("device lost")

    with SettleCoalescer(max_wait_ms=50.0, backend=boom) as co:
        futs = [co.submit(np.zeros((5, 4), np.float32), np.zeros(4, np.float32)) for _ in range(3)]
        for f in futs:
            with pytest.raises(RuntimeError):
                f.result(timeout=10)


def test_state_signature_matches_reference_value():
    """lattice.py:729-744 through the golden fixture the real reference produced."""
    from oracle import cases

    g, z = load_golden("config2_1200")
    c = cases.build("config2_1200")
    nbr = z["nbr"]
    r, t = np.nonzero(nbr >= 0)
    pairs = np.stack([r.astype(np.int64), nbr[r, t].astype(np.int64)], axis=1)[:2048]
    sig = state_signature(c["psi"], np.ones(1200, np.float32), [1.0, 0.5, 4.0, 0.0], 8, True, pairs)
    assert sig == g["state_sig"]


def test_a_failing_request_only_fails_its_own_future():
    """A backend may put an Exception in one request's slot; the other requests of the group get results."""
    def backend(group):
        return [ValueError("bad request") if float(r.psi[0]) == 1.0 else {"ok": True} for r in group]

    with SettleCoalescer(max_batch=8, max_wait_ms=200.0, backend=backend) as co:
        Y = np.zeros((6, 4), np.float32)
        futs = [co.submit(Y, np.full(4, float(i == 2), np.float32)) for i in range(5)]
        for i, f in enumerate(futs):
            if i == 2:
                with pytest.raises(ValueError):
                    f.result(timeout=10)
            else:
                assert f.result(timeout=10) == {"ok": True}


def test_chain_and_kneighbors_are_validated_on_the_submitting_thread():
    with SettleCoalescer(max_batch=2, max_wait_ms=1.0, backend=lambda g: [{} for _ in g]) as co:
        Y = np.zeros((6, 4), np.float32)
        with pytest.raises(ValueError):
            co.submit(Y, np.zeros(4, np.float32), chain=[0, 9])
        with pytest.raises(ValueError):
            co.submit(Y, np.zeros(4, np.float32), chain=[1])
        with pytest.raises(ValueError):
            co.submit(Y, np.zeros(4, np.float32), chain=[0, 1], lamP=-1.0)
        with pytest.raises(ValueError):
            co.submit(Y, np.zeros(4, np.float32), kneighbors=129)
        assert co.submit(Y, np.zeros(4, np.float32), chain=[0, 1]).result(timeout=10) == {}
