"""Row f4 on device: coalesced requests return what the per-request reference sequence returns
(cloud/app/main.py:916-939,1043,1061): same state_sig, iteration counts, deltaH."""
import numpy as np
import pytest

from oracle.sparse import SparseLattice

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    import torch

    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import oscillink_b200

    return oscillink_b200


def _req(seed, n=160, d=64):
    rs = np.random.RandomState(seed)
    Y = rs.randn(n, d).astype(np.float32)
    p = Y[:16].mean(axis=0)
    return Y, (p / (np.linalg.norm(p) + 1e-12)).astype(np.float32)


def test_coalesced_requests_match_single_lattices(api):
    from oscillink_b200.serving import SettleCoalescer

    reqs = [_req(s) for s in range(6)]
    gates = np.linspace(0.2, 1.0, 160).astype(np.float32)
    with SettleCoalescer(max_batch=16, max_wait_ms=300.0) as co:
        futs = [co.submit(Y, psi, kneighbors=6) for Y, psi in reqs]
        futs.append(co.submit(reqs[0][0], reqs[0][1], gates=gates, kneighbors=6))
        futs.append(co.submit(reqs[1][0], reqs[1][1], chain=[3, 9, 27], kneighbors=6))  # single path
        outs = [f.result(timeout=120) for f in futs]
    assert co.stats["max_batch_seen"] >= 6
    for (Y, psi), out in zip(reqs, outs[:6]):
        lat = api.OscillinkLattice(Y, kneighbors=6, deterministic_k=True)
        lat.set_query(psi)
        st = lat.settle()
        lat.set_receipt_detail("light")
        rec = lat.receipt()
        assert out["meta"]["path"] == "batched"
        assert out["state_sig"] == rec["meta"]["state_sig"]
        assert out["settle"]["iters"] == st["iters"]
        assert abs(out["receipt"]["deltaH_total"] - rec["deltaH_total"]) <= 1e-5 * abs(rec["deltaH_total"])
        assert out["receipt"]["meta"]["ustar_iters"] == rec["meta"]["ustar_iters"]
        assert abs(out["receipt"]["meta"]["avg_degree"] - rec["meta"]["avg_degree"]) < 1e-9
    # gated request vs the oracle
    o = SparseLattice(reqs[0][0], k=6)
    o.set_query(reqs[0][1], gates)
    so = o.settle()
    us, _, _ = o.stationary()
    assert outs[6]["settle"]["iters"] == so["iters"]
    dh = o.delta_h(us)
    assert abs(outs[6]["receipt"]["deltaH_total"] - dh) <= 1e-5 * abs(dh)
    # chain request took the general path and still carries a receipt
    assert outs[7]["meta"]["path"] == "single" and outs[7]["receipt"]["deltaH_total"] >= 0.0
