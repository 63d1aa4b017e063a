"""NCCL multi-GPU check + timing of the sharded lattice (run under torchrun on a multi-GPU box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/sharded_check.py --N 20000 --D 384 --k 10 --mode rows

Compares against the single-GPU class when N is small enough, prints device timings (max over
ranks) for build / settle / receipt."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oscillink_b200.sharded_api import ShardedLattice, shard_bounds  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=20000)
    ap.add_argument("--D", type=int, default=384)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--mode", default="rows")
    ap.add_argument("--chain", type=int, default=8)
    ap.add_argument("--check", type=int, default=1)
    ap.add_argument("--p2p", type=int, default=-1, help="-1 auto (on under NCCL), 0 all-gather halo, 1 fused P2P halo")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    r0, nl, _ = shard_bounds(a.N, world, rank)
    # counter-based generation: every rank can regenerate any row block (seed = block start)
    gen = torch.Generator(device=dev)
    Yl = torch.empty((nl, a.D), device=dev)
    CH = 65536
    for s in range(r0 - r0 % CH, r0 + nl, CH):
        gen.manual_seed(1000003 + s)
        blk = torch.randn((CH, a.D), generator=gen, device=dev)
        lo, hi = max(s, r0), min(s + CH, r0 + nl)
        Yl[lo - r0:hi - r0] = blk[lo - s:hi - s]
    head = torch.zeros((32, a.D), device=dev)
    if r0 < 32:
        head[r0:min(32, r0 + nl)] = Yl[:max(0, min(32, r0 + nl) - r0)]
    if world > 1:
        dist.all_reduce(head)
    psi = head.mean(0)
    psi = (psi / (psi.norm() + 1e-12)).cpu().numpy()

    def timed(fn):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return out, float(dt.item()) * 1e3

    sl, t_build = timed(lambda: ShardedLattice(Yl, a.N, kneighbors=a.k, mode=a.mode,
                                                    p2p=None if a.p2p < 0 else bool(a.p2p)))
    sl.set_query(psi)
    if a.chain >= 2:
        sl.add_chain(list(range(a.chain)), lamP=0.2)
    st, t_settle = timed(lambda: sl.settle(max_iters=12, tol=1e-3))
    rec, t_rec = timed(sl.receipt)
    res = {"N": a.N, "D": a.D, "k": a.k, "mode": a.mode, "world": world, "p2p_halo": sl._peers is not None,
           "build_ms": t_build,
           "settle_ms": t_settle, "receipt_ms": t_rec, "settle": {k: st[k] for k in ("iters", "res")},
           "deltaH": rec["deltaH_total"], "ustar_iters": rec["meta"]["ustar_iters"],
           "avg_degree": rec["meta"]["avg_degree"]}
    if a.check and a.N <= 60000:
        U = sl.U_full()
        if rank == 0:
            from oscillink_b200 import OscillinkLattice

            Yfull = np.empty((a.N, a.D), dtype=np.float32)
        Yall = [None]
        from oscillink_b200.sharded_api import gather_rows

        Yfull_t = gather_rows(Yl, a.N).cpu().numpy()
        if rank == 0:
            ref = OscillinkLattice(Yfull_t, kneighbors=a.k, deterministic_k=True)
            ref.set_query(psi)
            if a.chain >= 2:
                ref.add_chain(list(range(a.chain)), lamP=0.2)
            rst = ref.settle(max_iters=12, tol=1e-3)
            ref.set_receipt_detail("light")
            rr = ref.receipt()
            res["check"] = {
                "nbr_equal": bool(np.array_equal(sl._nbr.cpu().numpy(), ref._nbr.cpu().numpy())),
                "iters_ref": rst["iters"], "U_err": float(np.linalg.norm(U - ref.U) / np.linalg.norm(ref.U)),
                "dH_ref": rr["deltaH_total"],
            }
    if rank == 0:
        print(json.dumps(res))
    sl.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
