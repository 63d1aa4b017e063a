"""GPU test of the sharded lattice: two ranks (gloo, both on cuda:0 -- the test box has one GPU;
NCCL runs of the same code are in tools/sharded_check.py) against the single-GPU class and the
reference golden values."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mode, case, q, p2p=None):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import cases
        from oscillink_b200 import OscillinkLattice
        from oscillink_b200.sharded_api import ShardedLattice, shard_bounds

        c = cases.build(case)
        Y = c["Y"]
        N = Y.shape[0]
        r0, nl, _ = shard_bounds(N, world, rank)
        sl = ShardedLattice(Y[r0:r0 + nl], N, kneighbors=c["k"], row_cap_val=c["cap"], lamG=c["lam"][0],
                            lamC=c["lam"][1], lamQ=c["lam"][2], mode=mode, p2p=p2p)
        sl.set_query(c["psi"], c["gates"])
        if c["chain"] is not None:
            sl.add_chain(c["chain"], lamP=c["lamP"], weights=c["weights"])
        kw = {k: v for k, v in c["settle_kw"].items()}
        st = sl.settle(**kw)
        rec = sl.receipt()
        U = sl.U_full()
        Us = sl.Ustar_full()
        out = None
        if rank == 0:
            ref = OscillinkLattice(Y, kneighbors=c["k"], row_cap_val=c["cap"], lamG=c["lam"][0],
                                   lamC=c["lam"][1], lamQ=c["lam"][2], deterministic_k=True)
            ref.set_query(c["psi"], gates=c["gates"])
            if c["chain"] is not None:
                ref.add_chain(c["chain"], lamP=c["lamP"], weights=c["weights"])
            rst = ref.settle(**kw)
            ref.set_receipt_detail("light")
            rrec = ref.receipt()
            out = {
                "nbr_equal": bool(np.array_equal(sl._nbr.cpu().numpy(), ref._nbr.cpu().numpy())),
                "iters": (st["iters"], rst["iters"]), "res": (st["res"], rst["res"]),
                "dH": (rec["deltaH_total"], rrec["deltaH_total"]),
                "ustar_iters": (rec["meta"]["ustar_iters"], rrec["meta"]["ustar_iters"]),
                "U_err": float(np.linalg.norm(U - ref.U) / np.linalg.norm(ref.U)),
                "Us_err": float(np.linalg.norm(Us - ref.solve_Ustar()) / np.linalg.norm(ref.solve_Ustar())),
            }
        if out is not None:
            out["p2p_used"] = sl._peers is not None
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode,case,p2p", [("rows", "config2_1200", None), ("columns", "config2_1200", None),
                                           ("rows", "perf_400", None), ("columns", "perf_400", None),
                                           ("rows", "gates_300", None),
                                           # fused halo: peers' blocks read in place through CUDA IPC
                                           ("rows", "config2_1200", True), ("rows", "perf_400", True),
                                           ("rows", "gates_300", True)])
def test_sharded_matches_single_gpu(mode, case, p2p):
    import torch.multiprocessing as mp

    from tests.helpers import load_golden, rel

    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, case, q, p2p)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    res = dict(q.get() for _ in range(world))[0]
    g, _ = load_golden(case)
    assert res["p2p_used"] == bool(p2p)
    assert res["nbr_equal"]
    assert res["iters"][0] == res["iters"][1]
    assert rel(res["res"][0], res["res"][1]) < 1e-3
    assert res["ustar_iters"][0] == res["ustar_iters"][1] == g["ustar"]["iters"]
    assert rel(res["dH"][0], res["dH"][1]) < 1e-5
    from oracle import cases

    if not cases.build(case)["second_settle"]:  # golden deltaH of gates_300 is after its 2nd settle
        assert rel(res["dH"][0], g["deltaH"]) < 1e-5
    assert res["U_err"] < 1e-6 and res["Us_err"] < 1e-6


def _repartition_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import cases
        from oscillink_b200 import OscillinkLattice
        from oscillink_b200.sharded_api import ShardedLattice, shard_bounds

        c = cases.build("config2_1200")
        Y = c["Y"]
        N = Y.shape[0]
        r0, nl, _ = shard_bounds(N, world, rank)
        sl = ShardedLattice(Y[r0:r0 + nl], N, kneighbors=c["k"], mode="rows")
        sl.set_query(c["psi"])
        s1 = sl.settle()
        U_rows = sl.U_full()
        sl.repartition("columns")
        same = bool(np.array_equal(U_rows, sl.U_full()))
        s2 = sl.settle()           # second settle runs on column slabs, no graph rebuild
        sl.repartition("rows")
        U2 = sl.U_full()
        out = None
        if rank == 0:
            ref = OscillinkLattice(Y, kneighbors=c["k"], deterministic_k=True)
            ref.set_query(c["psi"])
            r1 = ref.settle()
            r2 = ref.settle()
            out = {"same": same, "iters": (s1["iters"], r1["iters"], s2["iters"], r2["iters"]),
                   "U2_err": float(np.linalg.norm(U2 - ref.U) / np.linalg.norm(ref.U))}
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_repartition_rows_columns_keeps_state_and_graph():
    import torch.multiprocessing as mp

    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_repartition_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    res = dict(q.get() for _ in range(world))[0]
    assert res["same"]
    assert res["iters"][0] == res["iters"][1] and res["iters"][2] == res["iters"][3]
    assert res["U2_err"] < 1e-6
