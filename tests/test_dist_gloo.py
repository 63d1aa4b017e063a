"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: block partition, padded row
gather, and the distributed PCG schedule (which collectives sit between which phases).

The schedule is exercised with an oracle-backed kernel facade on CPU tensors (tests may use the
oracle; the product only ever constructs the native facade) and must reproduce the
single-process oracle solve in both partitions (rows: halo all-gather + dot all-reduce;
columns: MAX all-reduce only)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oscillink_b200.sharded_api import gather_rows, pcg_schedule, shard_bounds


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_bounds_cover_rows_exactly():
    for N in [0, 1, 2, 7, 10, 1200, 1201]:
        for G in [1, 2, 3, 4, 8]:
            blocks = [shard_bounds(N, G, r) for r in range(G)]
            assert sum(b[1] for b in blocks) == N
            pos = 0
            for r0, n, shard in blocks:
                assert r0 == min(N, pos) and 0 <= n <= shard
                pos += n


class OracleKernels:
    """pcg_schedule facade backed by oracle.sparse (CPU, test-only)."""

    def __init__(self, lat, mode, rank, world, dt, X, Bv):
        self.lat, self.mode, self.dt = lat, mode, dt
        self.N = lat.N
        self.row0, self.n_loc, _ = shard_bounds(lat.N, world, rank)
        self.cols = None
        if mode == "columns":
            Dl = lat.D // world
            self.cols = slice(rank * Dl, (rank + 1) * Dl)
        self.X, self.R = X.clone(), Bv.clone()
        self.P = torch.zeros_like(self.X)
        self.AP = torch.zeros_like(self.X)
        self.sums = {}
        md = (1.0 + dt * lat._diag_base()) if dt is not None else lat._diag_base()
        self.md = torch.from_numpy(md.astype(np.float32))
        if mode == "rows":
            self.md = self.md[self.row0:self.row0 + self.n_loc]

    def _apply(self, v_all):
        full = np.zeros((self.lat.N, self.lat.D), dtype=np.float32)
        if self.mode == "rows":
            full[:] = v_all.numpy()
            out = self.lat.apply_M(full, self.dt)
            return torch.from_numpy(out[self.row0:self.row0 + self.n_loc])
        full[:, self.cols] = v_all.numpy()
        return torch.from_numpy(self.lat.apply_M(full, self.dt)[:, self.cols].copy())

    def x_local(self):
        return self.X

    def p_local(self):
        return self.P

    def residual0(self, x_all):
        self.R = self.R - self._apply(x_all)
        self.P = self.R / (self.md[:, None] + 1e-12)
        self.sums["rz"] = (self.R * self.P).sum(0)

    def spmm(self, p_all):
        self.AP = self._apply(p_all)
        self.sums["pap"] = (self.P * self.AP).sum(0)

    def reduce(self, which):
        return self.sums[which].clone()

    def update(self, rz, pap):
        alpha = rz / (pap + 1e-18)
        self.X = self.X + self.P * alpha
        self.R = self.R - self.AP * alpha
        z = self.R / (self.md[:, None] + 1e-12)
        self.sums["rr"] = (self.R * self.R).sum(0)
        self.sums["rz_new"] = (self.R * z).sum(0)

    def pupdate(self, rz_new, rz_old):
        z = self.R / (self.md[:, None] + 1e-12)
        self.P = z + self.P * (rz_new / (rz_old + 1e-18))


def _worker(rank, world, port, mode, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.sparse import SparseLattice

        N, D, k = 203, 16, 5
        rs = np.random.RandomState(0)
        Y = rs.randn(N, D).astype(np.float32)
        psi = rs.randn(D).astype(np.float32)
        lat = SparseLattice(Y, k=k)
        lat.set_query(psi / np.linalg.norm(psi))
        lat.add_chain([3, 9, 11], lamP=0.2)
        # gather_rows reassembles an uneven block partition
        r0, n_loc, _ = shard_bounds(N, world, rank)
        full = gather_rows(torch.from_numpy(Y[r0:r0 + n_loc]), N)
        assert torch.equal(full, torch.from_numpy(Y))
        # distributed schedule == single-process oracle
        dt = 1.0
        bvec = torch.from_numpy((lat.U + dt * lat.rhs()).astype(np.float32))
        x0 = torch.from_numpy(lat.U.copy())
        if mode == "rows":
            sl = slice(r0, r0 + n_loc)
            kern = OracleKernels(lat, mode, rank, world, dt, x0[sl], bvec[sl])
        else:
            Dl = D // world
            cs = slice(rank * Dl, (rank + 1) * Dl)
            kern = OracleKernels(lat, mode, rank, world, dt, x0[:, cs], bvec[:, cs])
        it, res = pcg_schedule(kern, mode=mode, tol=1e-3, max_iters=12)
        ref = SparseLattice(Y, k=k)
        ref.set_query(psi / np.linalg.norm(psi))
        ref.add_chain([3, 9, 11], lamP=0.2)
        st = ref.settle(dt=dt, max_iters=12, tol=1e-3)
        mine = kern.X.numpy()
        want = ref.U[r0:r0 + n_loc] if mode == "rows" else ref.U[:, cs]
        err = float(np.linalg.norm(mine - want) / np.linalg.norm(want))
        out.put((rank, it, res, st["iters"], st["res"], err))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["rows", "columns"])
def test_distributed_pcg_schedule_matches_single_process(mode):
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    got = [q.get() for _ in range(world)]
    for rank, it, res, it_ref, res_ref, err in got:
        assert it == it_ref
        assert abs(res - res_ref) <= 1e-3 * res_ref
        assert err < 1e-5
