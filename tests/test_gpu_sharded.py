"""GPU tests of the sharded lattice against the single-GPU class and the reference golden values:
two gloo ranks on cuda:0 (the Python-driven phase schedule, incl. the IPC-mapped fused halo), the C-ABI
distributed solve at world 1, the halo plan, and -- whenever two or more GPUs are visible -- real NCCL
ranks through osc_dist_pcg_solve (test_nccl_ranks_match_single_gpu; the round-1 tools/sharded_check.py)."""
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


# ------------------------------------------------------------------ C-ABI distributed solve (csrc/dist.cu)
@pytest.mark.parametrize("mode,case", [("rows", "config2_1200"), ("columns", "config2_1200"),
                                       ("rows", "gates_300"), ("rows", "perf_400"), ("columns", "perf_400")])
def test_dist_c_abi_world1_matches_class_and_golden(mode, case):
    """One rank: ShardedLattice routes settle / U* / deltaH through osc_dist_pcg_solve / osc_dist_delta_h
    (device-side stop test, lagged host poll) -- same results as the class and the reference goldens."""
    from oracle import cases
    from oscillink_b200 import OscillinkLattice
    from oscillink_b200.sharded_api import ShardedLattice
    from tests.helpers import load_golden, rel

    c = cases.build(case)
    Y = c["Y"]
    sl = ShardedLattice(Y, Y.shape[0], kneighbors=c["k"], row_cap_val=c["cap"], lamG=c["lam"][0],
                        lamC=c["lam"][1], lamQ=c["lam"][2], mode=mode)
    assert sl._use_c_path()
    sl.set_query(c["psi"], c["gates"])
    if c["chain"] is not None:
        sl.add_chain(c["chain"], lamP=c["lamP"], weights=c["weights"])
    st = sl.settle(**c["settle_kw"])
    rec = sl.receipt()
    ref = OscillinkLattice(Y, kneighbors=c["k"], row_cap_val=c["cap"], lamG=c["lam"][0], lamC=c["lam"][1],
                           lamQ=c["lam"][2], deterministic_k=True)
    ref.set_query(c["psi"], gates=c["gates"])
    if c["chain"] is not None:
        ref.add_chain(c["chain"], lamP=c["lamP"], weights=c["weights"])
    rst = ref.settle(**c["settle_kw"])
    ref.set_receipt_detail("light")
    rrec = ref.receipt()
    g, _ = load_golden(case)
    assert st["iters"] == rst["iters"] == g["settle"]["iters"]
    assert rel(st["res"], rst["res"]) < 1e-3
    assert rec["meta"]["ustar_iters"] == rrec["meta"]["ustar_iters"] == g["ustar"]["iters"]
    assert rel(rec["deltaH_total"], rrec["deltaH_total"]) < 1e-5
    assert np.linalg.norm(sl.U_full() - ref.U) / np.linalg.norm(ref.U) < 1e-6
    # max_iters caps and zero iterations through the same entry point
    sl2 = ShardedLattice(Y, Y.shape[0], kneighbors=c["k"], mode=mode)
    sl2.set_query(c["psi"])
    s2 = sl2.settle(max_iters=2, tol=1e-12)
    assert s2["iters"] == 2
    s0 = sl2.settle(max_iters=0)
    assert s0["iters"] == 0


def test_halo_plan_matches_numpy():
    """osc_dist_halo_plan: ascending unique remote rows + neighbour ids remapped into the block."""
    import ctypes as C

    import torch

    from oscillink_b200 import _cabi

    lib = _cabi.load()
    rs = np.random.RandomState(3)
    N, k, world = 5000, 7, 4
    shard = (N + world - 1) // world
    for rank in (0, 2, 3):
        r0 = rank * shard
        nl = min(N, r0 + shard) - r0
        nbr = rs.randint(0, N, size=(nl, k)).astype(np.int32)
        nbr[rs.rand(nl, k) < 0.2] = -1
        extra = rs.randint(0, N, size=11).astype(np.int32)
        d_nbr, d_extra = torch.from_numpy(nbr).cuda(), torch.from_numpy(extra).cuda()
        need = C.c_size_t(0)
        _cabi.check(lib.osc_dist_halo_plan_workspace(N, C.byref(need)))
        ws = torch.empty(need.value, dtype=torch.uint8, device="cuda")
        n_halo = C.c_int64(0)
        args = (d_nbr.data_ptr(), nl, k, d_extra.data_ptr(), extra.size, N, r0, shard)
        _cabi.check(lib.osc_dist_halo_plan(*args, None, 0, None, None, C.byref(n_halo), ws.data_ptr(), ws.numel(), None))
        ids = np.concatenate([nbr.reshape(-1), extra])
        remote = np.unique(ids[(ids >= 0) & ((ids < r0) | (ids >= r0 + nl))])
        assert n_halo.value == len(remote)
        rows = torch.empty(len(remote), dtype=torch.int32, device="cuda")
        out = torch.empty_like(d_nbr)
        eout = torch.empty_like(d_extra)
        _cabi.check(lib.osc_dist_halo_plan(*args, rows.data_ptr(), len(remote), out.data_ptr(), eout.data_ptr(),
                                           C.byref(n_halo), ws.data_ptr(), ws.numel(), None))
        assert np.array_equal(rows.cpu().numpy(), remote.astype(np.int32))
        pos = {int(j): t for t, j in enumerate(remote)}

        def remap(j):
            if j < 0:
                return j
            return j - r0 if r0 <= j < r0 + nl else shard + pos[int(j)]

        want = np.vectorize(remap)(nbr).astype(np.int32)
        assert np.array_equal(out.cpu().numpy(), want)
        assert np.array_equal(eout.cpu().numpy(), np.vectorize(remap)(extra).astype(np.int32))


def _nccl_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from oscillink_b200 import OscillinkLattice
        from oscillink_b200.sharded_api import ShardedLattice, shard_bounds

        rs = np.random.RandomState(5)
        N, D, k = 6001, 64, 9  # N not divisible by the world size: the last shard is short
        Y = rs.randn(N, D).astype(np.float32)
        psi = Y[:32].mean(axis=0)
        psi = (psi / (np.linalg.norm(psi) + 1e-12)).astype(np.float32)
        chain = [3, N - 2, 17, N // 2]  # crosses the shards
        r0, nl, _ = shard_bounds(N, world, rank)
        out = {}
        ref = None
        if rank == 0:
            ref = OscillinkLattice(Y, kneighbors=k, deterministic_k=True)
            ref.set_query(psi)
            ref.add_chain(chain, lamP=0.2)
            rst = ref.settle()
            ref.set_receipt_detail("light")
            rrec = ref.receipt()
        for mode, halo in (("rows", "pull"), ("rows", "allgather"), ("columns", None)):
            if mode == "columns" and D % (4 * world):
                continue
            sl = ShardedLattice(Y[r0:r0 + nl], N, kneighbors=k, mode=mode, halo=halo)
            sl.set_query(psi)
            sl.add_chain(chain, lamP=0.2)
            st = sl.settle()
            rec = sl.receipt()
            U = sl.U_full()
            c_path = sl._use_c_path()
            used = sl.halo
            if rank == 0:
                out[f"{mode}/{halo}"] = {
                    "c_path": bool(c_path), "halo": used,
                    "nbr_equal": bool(np.array_equal(sl._nbr.cpu().numpy(), ref._nbr.cpu().numpy())),
                    "iters": (st["iters"], rst["iters"]), "res": (st["res"], rst["res"]),
                    "dH": (rec["deltaH_total"], rrec["deltaH_total"]),
                    "ustar_iters": (rec["meta"]["ustar_iters"], rrec["meta"]["ustar_iters"]),
                    "U_err": float(np.linalg.norm(U - ref.U) / np.linalg.norm(ref.U))}
            sl.close()
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_nccl_ranks_match_single_gpu():
    """Real NCCL ranks (one per GPU): rows partition with the pull halo and with the all-gather halo, and
    column slabs, all through osc_dist_pcg_solve -- equal to the single-GPU class.  Needs >= 2 GPUs."""
    import torch
    import torch.multiprocessing as mp

    from tests.helpers import rel

    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    res = dict(q.get() for _ in range(world))[0]
    assert "rows/pull" in res and "rows/allgather" in res
    for name, r in res.items():
        assert r["c_path"], name
        assert r["nbr_equal"], name
        assert r["iters"][0] == r["iters"][1], name
        assert rel(r["res"][0], r["res"][1]) < 1e-3, name
        assert r["ustar_iters"][0] == r["ustar_iters"][1], name
        assert rel(r["dH"][0], r["dH"][1]) < 1e-5, name
        assert r["U_err"] < 1e-6, name
    assert res["rows/pull"]["halo"] == "pull"
