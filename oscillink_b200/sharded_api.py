"""One lattice sharded over the GPUs of a box (BASELINE.json configs #4/#5).

One process per GPU (`torch.distributed`, NCCL over NVLink/NVSwitch; gloo also works and is what
the CPU tests use for the host-side schedule).  Rank g owns the contiguous row block
[g*S, min(N,(g+1)*S)) with S = ceil(N/G) of Y, U and of the graph.

Build (graph.py:29-93):
  all-gather the anchor rows -> every rank normalises all N rows with the same kernel (bit-identical
  Yn everywhere) -> each rank runs the fused similarity/top-k kernel for ITS row panel against all N
  columns -> canonical rescoring of its rows -> all-gather of the N x k top tables -> every rank
  assembles the (tiny, O(N k)) mutual/cap/degree arrays for all rows redundantly, so no further
  exchange of c_j / 1/sd_j is needed.

Solve (solver.py:15-37), two exact partitions of the same recurrences:
  mode="rows"    (north_star): per iteration the search direction p is all-gathered (the halo
                 exchange; on kNN graphs of random anchors ~2/3 of every remote shard is needed
                 anyway), the SpMM runs on local rows, and the per-column dot products
                 (p.Ap, r.r, r.z: D floats each) are all-reduced.
  mode="columns" every reduction of the solver is per column, so a rank that owns ALL rows of a
                 D/G column slab needs no halo and no dot all-reduce -- only a 1-float MAX
                 all-reduce for the stop test.  Results are identical.

  mode="rows", p2p=True (default under NCCL): the halo exchange is FUSED into the SpMM.  Every rank
                 keeps its block of x0 / p in a buffer that its peers map with CUDA IPC; the SpMM
                 kernel fetches remote neighbour rows with plain loads over NVLink while it computes
                 (osc_pcg_spmm_dot_p2p), so there is no all-gather and no N x D staging buffer --
                 only a stream-ordered barrier between "p written" and "p read by peers".

Under NCCL the whole solve is ONE call into the C ABI (osc_dist_pcg_solve, csrc/dist.cu): NCCL is called
from C on the current stream and the stop test runs on the device.  halo="pull" (rows partition): every
rank pulls the unique remote rows its graph references from the peers' blocks over NVLink (CUDA IPC
mappings) -- one fetch per ROW, ascending, instead of the all-gather's every row or the fused kernel's
one fetch per REFERENCE.  Without NCCL (gloo: the CPU / single-device test rigs) the same recurrences are
driven from Python through the exported phase entry points (osc_pcg_*), `pcg_schedule` below.
"""
from __future__ import annotations

import ctypes as C
import time
from typing import Any

import numpy as np

__all__ = ["ShardedLattice", "shard_bounds", "gather_rows", "pcg_schedule"]


# ----------------------------------------------------------------------------- partition helpers
def shard_bounds(N: int, world: int, rank: int) -> tuple[int, int, int]:
    """(row0, n_local, shard) for the contiguous block partition with shard = ceil(N/world)."""
    shard = (N + world - 1) // world if world > 0 else N
    row0 = min(N, rank * shard)
    return row0, max(0, min(N, row0 + shard) - row0), shard


def gather_rows(local, N: int, group=None):
    """All-gather row blocks of the block partition into the full [N, ...] tensor (every rank).
    Blocks are padded to the common shard size so a plain all_gather_into_tensor applies."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local[:N]
    rank = dist.get_rank(group)
    _, n_loc, shard = shard_bounds(N, world, rank)
    assert local.shape[0] == n_loc, (local.shape, n_loc)
    tail = tuple(local.shape[1:])
    padded = local
    if n_loc < shard:
        padded = torch.zeros((shard,) + tail, dtype=local.dtype, device=local.device)
        padded[:n_loc] = local
    if padded.is_cuda and dist.get_backend(group) == "gloo":
        # gloo has no CUDA all_gather: stage through the host (single-GPU test rigs only;
        # production runs use NCCL)
        host = torch.empty((world * shard,) + tail, dtype=local.dtype)
        dist.all_gather_into_tensor(host, padded.cpu().contiguous(), group=group)
        return host.to(local.device)[:N]
    out = torch.empty((world * shard,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    return out[:N]


def _allreduce(t, op, group=None):
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=op, group=group)
    return t


def pcg_schedule(k, *, mode: str, tol: float, max_iters: int, group=None):
    """The distributed PCG driver (solver.py:19-37) written against a kernel facade `k`:

        k.residual0(x_all) -> None          r = b - A x ; p = precond(r); fills k.part_rz
        k.reduce(which)    -> tensor[D]     column sums of the named partial ("rz","pap","rr","rz_new")
        k.spmm(p_all)      -> None          Ap = A p ; fills part_pap
        k.update(rz, pap)  -> None          x += alpha p ; r -= alpha Ap ; fills part_rr, part_rz_new
        k.pupdate(rz_new, rz) -> None       p = z + beta p
        k.p_local() / k.x_local()           local row blocks (mode rows) or column slabs

    Returns (iters, res).  `mode` selects which collectives sit between the phases."""
    import torch
    import torch.distributed as dist

    rows = mode == "rows"
    SUM, MAX = dist.ReduceOp.SUM, dist.ReduceOp.MAX

    def full(vec):
        if rows and getattr(k, "p2p", False):
            k.peer_sync()  # every rank's block is written before any peer's SpMM reads it
            return None    # the kernel reads the peers' blocks in place
        return gather_rows(vec, k.N, group) if rows else vec

    def colsum(which):
        s = k.reduce(which)
        return _allreduce(s, SUM, group) if rows else s

    k.residual0(full(k.x_local()))
    rz = colsum("rz").clone()
    it, res = 0, float("nan")
    for it in range(1, max_iters + 1):
        k.spmm(full(k.p_local()))
        pap = colsum("pap")
        k.update(rz, pap)
        if rows:
            both = torch.stack([k.reduce("rr"), k.reduce("rz_new")])
            _allreduce(both, SUM, group)
            rr, rz_new = both[0], both[1]
            res = float(torch.sqrt(torch.clamp(rr.max(), min=0)).item())
        else:
            rr, rz_new = k.reduce("rr"), k.reduce("rz_new")
            m = torch.sqrt(torch.clamp(rr.max(), min=0)).reshape(1)
            res = float(_allreduce(m, MAX, group).item())
        if res <= tol or it == max_iters:
            break
        k.pupdate(rz_new, rz)
        rz = rz_new.clone()
    return it, res


# ----------------------------------------------------------------------------- native kernel facade
class _NativeKernels:
    """osc_pcg_* phase calls for one (rows | columns) partition on the current device."""

    def __init__(self, lat: "ShardedLattice", mode_id: int, dt: float, jacobi: bool, X, Bv):
        import torch

        from . import _cabi

        self.cabi, self.lib, self.lat = _cabi, lat._lib, lat
        self.N = lat.N
        self.mode_id, self.dt, self.jacobi = mode_id, float(dt), 1 if jacobi else 0
        dev = lat._dev
        if lat.mode == "rows":
            self.dims = _cabi.PcgDims(lat.N, lat.row0, lat.n_local, lat.D, 0)
            self.graph = lat._graph_struct(local=True)
            self.gates = lat._dB_loc
            self.Dl = lat.D
        else:
            self.dims = _cabi.PcgDims(lat.N, 0, lat.N, lat.Dl, 0)
            self.graph = lat._graph_struct(local=False)
            self.gates = lat._dB_all
            self.Dl = lat.Dl
        need = C.c_size_t(0)
        _cabi.check(self.lib.osc_pcg_plan(C.byref(self.dims), C.byref(need)))
        nb, Dl = self.dims.n_blocks, self.Dl
        self.X, self.R = X, Bv
        self.p2p = lat.mode == "rows" and lat._peers is not None and X.data_ptr() == lat._peers.X.data_ptr()
        self.P = lat._peers.P[: X.shape[0]] if self.p2p else torch.empty_like(X)
        self.AP = torch.empty_like(X)
        self.parts = {n: torch.zeros((nb, Dl), dtype=torch.float64, device=dev)
                      for n in ("rz", "pap", "rr", "rz_new")}
        self.out = {n: torch.zeros(Dl, dtype=torch.float32, device=dev) for n in self.parts}
        self.prm = lat._params_struct()
        self.chain = lat._chain_struct()

    def _st(self):
        import torch

        return torch.cuda.current_stream().cuda_stream

    def _chain_arg(self):
        return C.byref(self.chain) if self.chain is not None else None

    def x_local(self):
        return self.X

    def p_local(self):
        return self.P

    def peer_sync(self):
        self.lat._peers.sync()

    def residual0(self, x_all):
        if x_all is None:  # fused halo: x0 blocks are read from the peers' buffers
            pe = self.lat._peers
            self.cabi.check(self.lib.osc_pcg_residual0_p2p(
                C.byref(self.dims), C.byref(self.graph), self._chain_arg(), C.byref(self.prm), self.mode_id,
                self.dt, self.jacobi, self.gates.data_ptr(), pe.tabX.data_ptr(), pe.world, pe.shard,
                self.R.data_ptr(), self.P.data_ptr(), self.parts["rz"].data_ptr(), self._st()),
                "osc_pcg_residual0_p2p")
            return
        self.cabi.check(self.lib.osc_pcg_residual0(
            C.byref(self.dims), C.byref(self.graph), self._chain_arg(), C.byref(self.prm), self.mode_id,
            self.dt, self.jacobi, self.gates.data_ptr(), x_all.data_ptr(), self.R.data_ptr(),
            self.P.data_ptr(), self.parts["rz"].data_ptr(), self._st()), "osc_pcg_residual0")

    def spmm(self, p_all):
        if p_all is None:
            pe = self.lat._peers
            self.cabi.check(self.lib.osc_pcg_spmm_dot_p2p(
                C.byref(self.dims), C.byref(self.graph), self._chain_arg(), C.byref(self.prm), self.mode_id,
                self.dt, self.gates.data_ptr(), pe.tabP.data_ptr(), pe.world, pe.shard, self.AP.data_ptr(),
                self.parts["pap"].data_ptr(), self._st()), "osc_pcg_spmm_dot_p2p")
            return
        self.cabi.check(self.lib.osc_pcg_spmm_dot(
            C.byref(self.dims), C.byref(self.graph), self._chain_arg(), C.byref(self.prm), self.mode_id,
            self.dt, self.gates.data_ptr(), p_all.data_ptr(), self.AP.data_ptr(),
            self.parts["pap"].data_ptr(), self._st()), "osc_pcg_spmm_dot")

    def reduce(self, which):
        self.cabi.check(self.lib.osc_pcg_reduce(self.parts[which].data_ptr(), self.dims.n_blocks, self.Dl,
                                                self.out[which].data_ptr(), None, self._st()),
                        "osc_pcg_reduce")
        return self.out[which]

    def update(self, rz, pap, with_x: bool = True):
        self.cabi.check(self.lib.osc_pcg_update(
            C.byref(self.dims), C.byref(self.prm), self.mode_id, self.dt, self.jacobi,
            self.gates.data_ptr(), rz.data_ptr(), pap.data_ptr(), self.P.data_ptr(), self.AP.data_ptr(),
            self.X.data_ptr() if with_x else None, self.R.data_ptr(), self.parts["rr"].data_ptr(),
            self.parts["rz_new"].data_ptr(), self._st()), "osc_pcg_update")

    def pupdate_x(self, rz_new, rz_old, pap, last: bool = False):
        """x += alpha p fused with p = z + beta p: the pair osc_pcg_solve / osc_dist_pcg_solve run (after
        update(with_x=False)); exposed for per-kernel timing."""
        self.cabi.check(self.lib.osc_pcg_pupdate_x(
            C.byref(self.dims), C.byref(self.prm), self.mode_id, self.dt, self.jacobi,
            self.gates.data_ptr(), rz_new.data_ptr(), rz_old.data_ptr(), pap.data_ptr(), self.R.data_ptr(),
            self.P.data_ptr(), self.X.data_ptr(), 1 if last else 0, self._st()), "osc_pcg_pupdate_x")

    def pupdate(self, rz_new, rz_old):
        self.cabi.check(self.lib.osc_pcg_pupdate(
            C.byref(self.dims), C.byref(self.prm), self.mode_id, self.dt, self.jacobi,
            self.gates.data_ptr(), rz_new.data_ptr(), rz_old.data_ptr(), self.R.data_ptr(),
            self.P.data_ptr(), self._st()), "osc_pcg_pupdate")


# ----------------------------------------------------------------------------- peer-mapped state
class _RawDeviceArray:
    """__cuda_array_interface__ holder so torch can view a cudaMalloc'd block as a tensor."""

    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 2}


class _PeerBuffers:
    """Per-lattice X (iterate / x0) and P (search direction) row blocks that every peer rank maps into
    its own address space with CUDA IPC (osc_peer_alloc / osc_peer_open).  tabX / tabP are the device
    tables of the `world` block base pointers the *_p2p kernels index."""

    def __init__(self, lat: "ShardedLattice"):
        import torch
        import torch.distributed as dist

        cabi, lib = _cabi_mod(), lat._lib
        self._lib, self._cabi = lib, cabi
        self.group, self.world, self.shard = lat.group, lat.world, lat.shard
        dev = lat._dev
        rows = max(lat.shard, 1)
        nbytes = rows * lat.D * 4
        self._own, self._opened = [], []
        handles = []
        for _ in range(2):
            ptr, h = C.c_void_p(), (C.c_ubyte * 64)()
            cabi.check(lib.osc_peer_alloc(nbytes, C.byref(ptr), h), "osc_peer_alloc")
            self._own.append(ptr.value)
            handles.append(bytes(h))
        self.X = torch.as_tensor(_RawDeviceArray(self._own[0], (rows, lat.D)), device=dev)
        self.P = torch.as_tensor(_RawDeviceArray(self._own[1], (rows, lat.D)), device=dev)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, (lat.rank, dev.index, handles[0], handles[1]), group=self.group)
        ptrX, ptrP = [0] * self.world, [0] * self.world
        for rank, dev_index, hx, hp in everyone:
            if rank == lat.rank:
                ptrX[rank], ptrP[rank] = self._own
                continue
            if dev_index != dev.index:
                cabi.check(lib.osc_enable_peer_access(dev_index), "osc_enable_peer_access")
            for tab, h in ((ptrX, hx), (ptrP, hp)):
                ptr = C.c_void_p()
                buf = (C.c_ubyte * 64).from_buffer_copy(h)
                cabi.check(lib.osc_peer_open(buf, C.byref(ptr)), "osc_peer_open")
                self._opened.append(ptr.value)
                tab[rank] = ptr.value
        self.tabX = torch.tensor(ptrX, dtype=torch.int64, device=dev)
        self.tabP = torch.tensor(ptrP, dtype=torch.int64, device=dev)
        self._flag = torch.zeros(1, dtype=torch.float32, device=dev)
        self._nccl = dist.get_backend(self.group) == "nccl"
        self.sync()

    def sync(self) -> None:
        """Cross-rank ordering point: work enqueued before it on every rank is complete before work
        enqueued after it starts on any rank.  NCCL: a 1-float all-reduce, stream-ordered, no host
        synchronisation.  gloo (single-GPU test rigs): device synchronise + host barrier."""
        import torch
        import torch.distributed as dist

        if self._nccl:
            dist.all_reduce(self._flag, op=dist.ReduceOp.MAX, group=self.group)
        else:
            torch.cuda.synchronize()
            dist.barrier(group=self.group)

    def close(self) -> None:
        """Unmap the peers' blocks and free this rank's (collective: every rank calls it)."""
        import torch

        if not self._own:
            return
        torch.cuda.synchronize()
        try:
            self.sync()  # nobody is still reading a block that is about to be freed
        except Exception:
            pass
        for p in self._opened:
            self._lib.osc_peer_close(C.c_void_p(p))
        self.X = self.P = None
        for p in self._own:
            self._lib.osc_peer_free(C.c_void_p(p))
        self._own, self._opened = [], []


class _PullHalo:
    """State of the OSC_HALO_PULL strategy for one rank: the halo plan (osc_dist_halo_plan) and the
    peer-mapped block [shard own rows | n_halo pulled rows][D] whose first `shard` rows every peer maps with
    CUDA IPC (osc_peer_alloc / osc_peer_open).  Collective: every rank builds it at the same point."""

    def __init__(self, lat: "ShardedLattice"):
        import torch
        import torch.distributed as dist

        cabi, lib, dev = _cabi_mod(), lat._lib, lat._dev
        self._lib, self._cabi, self.group, self.world = lib, cabi, lat.group, lat.world
        st = torch.cuda.current_stream().cuda_stream
        r0, nl, k = lat.row0, lat.n_local, lat.k
        nbr_loc = lat._nbr[r0:r0 + nl].contiguous()
        extra = lat._chain["col"] if lat._chain is not None else None
        n_extra = int(extra.numel()) if extra is not None else 0
        need = C.c_size_t(0)
        cabi.check(lib.osc_dist_halo_plan_workspace(lat.N, C.byref(need)))
        ws = torch.empty(max(need.value, 256), dtype=torch.uint8, device=dev)
        n_halo = C.c_int64(0)
        args = (nbr_loc.data_ptr() if nl else None, nl, k, extra.data_ptr() if n_extra else None, n_extra,
                lat.N, r0, lat.shard)
        cabi.check(lib.osc_dist_halo_plan(*args, None, 0, None, None, C.byref(n_halo), ws.data_ptr(),
                                          ws.numel(), st), "osc_dist_halo_plan")
        self.n_halo = int(n_halo.value)
        self.halo_rows = torch.empty(max(self.n_halo, 1), dtype=torch.int32, device=dev)
        self.nbr_local = torch.full((max(nl, 1), k), -1, dtype=torch.int32, device=dev)
        self.chain_col = torch.empty(max(n_extra, 1), dtype=torch.int32, device=dev)
        cabi.check(lib.osc_dist_halo_plan(*args, self.halo_rows.data_ptr(), self.n_halo,
                                          self.nbr_local.data_ptr(), self.chain_col.data_ptr() if n_extra else None,
                                          C.byref(n_halo), ws.data_ptr(), ws.numel(), st), "osc_dist_halo_plan")
        del ws
        rows = max(lat.shard + self.n_halo, 1)
        ptr, h = C.c_void_p(), (C.c_ubyte * 64)()
        cabi.check(lib.osc_peer_alloc(rows * lat.D * 4, C.byref(ptr), h), "osc_peer_alloc")
        self.block_ptr, self._opened = ptr.value, []
        everyone = [None] * self.world
        dist.all_gather_object(everyone, (lat.rank, dev.index, bytes(h)), group=self.group)
        tab = [0] * self.world
        for rank, dev_index, hb in everyone:
            if rank == lat.rank:
                tab[rank] = self.block_ptr
                continue
            if dev_index != dev.index:
                cabi.check(lib.osc_enable_peer_access(dev_index), "osc_enable_peer_access")
            p = C.c_void_p()
            cabi.check(lib.osc_peer_open((C.c_ubyte * 64).from_buffer_copy(hb), C.byref(p)), "osc_peer_open")
            self._opened.append(p.value)
            tab[rank] = p.value
        self.tab = torch.tensor(tab, dtype=torch.int64, device=dev)
        self.halo_fraction = self.n_halo / max(lat.N - nl, 1)  # share of the remote rows that is needed
        # halo rows owned by lower ranks: the pull starts behind them (osc_dist_t.halo_below)
        self.n_below = int((self.halo_rows[: self.n_halo] < r0).sum().item()) if self.n_halo else 0
        torch.cuda.synchronize()
        dist.barrier(group=self.group)

    def chain_struct(self, chain):
        c = self._cabi.Chain
        return c(chain["n_rows"], chain["nnz"], chain["rows"].data_ptr(), chain["rowptr"].data_ptr(),
                 self.chain_col.data_ptr(), chain["Wp"].data_ptr(), chain["Ap"].data_ptr(),
                 chain["slot"].data_ptr())

    def close(self) -> None:
        """Unmap the peers' blocks and free this rank's (collective)."""
        import torch
        import torch.distributed as dist

        if self.block_ptr is None:
            return
        torch.cuda.synchronize()
        try:
            dist.barrier(group=self.group)  # nobody is still reading a block that is about to be freed
        except Exception:
            pass
        for p in self._opened:
            self._lib.osc_peer_close(C.c_void_p(p))
        self._lib.osc_peer_free(C.c_void_p(self.block_ptr))
        self.block_ptr, self._opened = None, []



def _cabi_mod():
    from . import _cabi

    return _cabi


# ----------------------------------------------------------------------------- the lattice
class ShardedLattice:
    """Row-sharded lattice.  `Y_local` is this rank's row block (see shard_bounds)."""

    def __init__(self, Y_local, N: int, kneighbors: int = 6, row_cap_val: float = 1.0, lamG: float = 1.0,
                 lamC: float = 0.5, lamQ: float = 4.0, *, mode: str = "rows", group=None,
                 knn_engine: int = 0, p2p: bool | None = None, halo: str | None = None):
        import torch
        import torch.distributed as dist

        from . import _cabi

        if mode not in {"rows", "columns"}:
            raise ValueError("mode must be 'rows' or 'columns'")
        if kneighbors < 1:
            raise ValueError("kneighbors must be >= 1")
        if lamG <= 0:
            raise ValueError("lamG must be > 0 for SPD")
        if lamC < 0 or lamQ < 0:
            raise ValueError("lamC and lamQ must be >= 0")
        if not torch.cuda.is_available():
            raise RuntimeError("oscillink_b200 needs a CUDA device (sm_100a); no CPU fallback")
        self._cabi, self._lib = _cabi, _cabi.load()
        self._dev = torch.device("cuda", torch.cuda.current_device())
        self.group, self.mode = group, mode
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.N = int(N)
        self.row0, self.n_local, self.shard = shard_bounds(self.N, self.world, self.rank)
        if isinstance(Y_local, np.ndarray):
            Y_local = torch.from_numpy(np.ascontiguousarray(Y_local, dtype=np.float32))
        Y_local = Y_local.to(self._dev, dtype=torch.float32).contiguous()
        if Y_local.ndim != 2 or Y_local.shape[0] != self.n_local:
            raise ValueError(f"Y_local must hold this rank's {self.n_local} rows")
        self.D = int(Y_local.shape[1])
        if mode == "columns":
            if self.D % (4 * self.world) != 0:
                raise ValueError("mode='columns' needs D divisible by 4*world_size")
            self.Dl = self.D // self.world
            self.c0 = self.rank * self.Dl
        self.k = min(int(kneighbors), max(1, self.N - 1))
        self.row_cap_val = float(row_cap_val)
        self.lamG, self.lamC, self.lamQ, self.lamP = float(lamG), float(lamC), float(lamQ), 0.0
        self._chain = None
        self._chain_nodes = None
        self._engine = knn_engine
        if halo not in (None, "allgather", "pull", "fused"):
            raise ValueError("halo must be None, 'allgather', 'pull' or 'fused'")
        if halo is None and p2p is not None:  # legacy switch: p2p=True is the fused kernel
            halo = "fused" if p2p else "allgather"
        self._halo_arg = halo  # resolved after the build, see _resolve_halo
        self.halo = "allgather"
        self._want_p2p = False
        self._peers = None
        self._pull = None
        self._comm = None      # ncclComm_t of the C-ABI solve (None: not tried yet, 0: unavailable)
        self._comm_owned = False
        self.last: dict[str, Any] = {"iters": 0, "res": None, "t_ms": None}
        self.timings: dict[str, float] = {}
        self._build(Y_local)
        self._resolve_halo()
        self._dB_all = torch.ones(self.N, dtype=torch.float32, device=self._dev)
        self._dB_loc = self._dB_all[self.row0:self.row0 + self.n_local]
        self._dpsi = torch.zeros(self.D, dtype=torch.float32, device=self._dev)
        self._Ustar = None

    # ---- local views of the state: rows mode [n_local, D]; columns mode [N, D/G]
    def _slab(self, full):
        return full[:, self.c0:self.c0 + self.Dl].contiguous()

    def _build(self, Y_local) -> None:
        import torch

        cabi, lib, dev = self._cabi, self._lib, self._dev
        N, D, k = self.N, self.D, self.k
        t0 = time.time()
        Y_all = gather_rows(Y_local, N, self.group).contiguous()
        st = torch.cuda.current_stream().cuda_stream
        nbr = torch.full((N, k), -1, dtype=torch.int32, device=dev)
        A = torch.zeros((N, k), dtype=torch.float32, device=dev)
        W = torch.zeros((N, k), dtype=torch.float32, device=dev)
        deg = torch.zeros(N, dtype=torch.int32, device=dev)
        sd = torch.full((N,), 1e-6, dtype=torch.float32, device=dev)
        self.nnz = torch.zeros(1, dtype=torch.int64, device=dev)
        if N >= 2:
            eng, kc, eps = cabi.knn_plan(max(self.n_local, 1), N, D, k, self._engine)
            use_tc = eng in (cabi.KNN_TC, cabi.KNN_TC1)
            self.engine_used = cabi.ENGINE_NAMES[eng]
            Yn = torch.empty_like(Y_all)
            hi = torch.empty_like(Y_all) if use_tc else None
            lo = torch.empty_like(Y_all) if eng == cabi.KNN_TC else None
            P = cabi.ptr
            if eng == cabi.KNN_TCH:  # fp16 rows ride in the q_hi / all_hi arguments
                hi = torch.empty(Y_all.shape, dtype=torch.float16, device=dev)
                cabi.check(lib.osc_normalize_rows_f16(Y_all.data_ptr(), N, D, Yn.data_ptr(), hi.data_ptr(), st))
            else:
                cabi.check(lib.osc_normalize_rows(Y_all.data_ptr(), N, D, Yn.data_ptr(), P(hi), P(lo), st))
            nl, r0 = self.n_local, self.row0
            top_idx = torch.full((max(nl, 1), k), -1, dtype=torch.int32, device=dev)
            top_sim = torch.zeros((max(nl, 1), k), dtype=torch.float32, device=dev)
            gap = torch.empty(max(nl, 1), dtype=torch.float32, device=dev)
            if nl > 0:
                q = lambda t: None if t is None else t.data_ptr() + r0 * D * t.element_size()  # noqa: E731
                self.n_exhaustive = torch.zeros(1, dtype=torch.int32, device=dev)
                need = C.c_size_t(0)
                cabi.check(lib.osc_knn_rescore_workspace(1, nl, C.byref(need)))
                rws = torch.empty(max(need.value, 256), dtype=torch.uint8, device=dev)

                def run(eng, kc, eps, hi, lo, limit):
                    cand_idx = torch.empty((nl, kc), dtype=torch.int32, device=dev)
                    cand_sim = torch.empty((nl, kc), dtype=torch.float32, device=dev)
                    cabi.check(lib.osc_knn_candidates(
                        q(Yn), Yn.data_ptr(), q(hi), q(lo), P(hi), P(lo), 1, nl, r0, N, D, kc,
                        eng, cand_idx.data_ptr(), cand_sim.data_ptr(), None, 0, st), "osc_knn_candidates")
                    cabi.check(lib.osc_knn_rescore_guarded(
                        q(Yn), Yn.data_ptr(), 1, nl, r0, N, D, cand_idx.data_ptr(), cand_sim.data_ptr(), kc, k,
                        eps, limit, top_idx.data_ptr(), top_sim.data_ptr(), gap.data_ptr(),
                        self.n_exhaustive.data_ptr(), rws.data_ptr(), rws.numel(), st),
                        "osc_knn_rescore_guarded")

                # single-product engines: bounded exhaustive path, hand-over to 3xTF32 if too many rows of this
                # rank cannot be proven complete (clustered / near-duplicate anchors) -- rank-local decision
                single = eng in (cabi.KNN_TC1, cabi.KNN_TCH)
                limit = int(lib.osc_knn_exhaustive_limit(nl)) if single else -1
                run(eng, kc, eps, hi, lo, limit)
                if single and int(self.n_exhaustive.item()) > limit:
                    e3, kc3, eps3 = cabi.knn_plan(nl, N, D, k, cabi.KNN_TC)
                    del hi, lo
                    hi, lo = torch.empty_like(Y_all), torch.empty_like(Y_all)
                    cabi.check(lib.osc_normalize_rows(Y_all.data_ptr(), N, D, Yn.data_ptr(), hi.data_ptr(),
                                                      lo.data_ptr(), st))
                    run(e3, kc3, eps3, hi, lo, -1)
                    self.engine_used = cabi.ENGINE_NAMES[e3] + " (fallback from " + self.engine_used + ")"
            self.gap_local = gap[:nl]
            top_idx_all = gather_rows(top_idx[:nl], N, self.group).contiguous()
            top_sim_all = gather_rows(top_sim[:nl], N, self.group).contiguous()
            scratch = torch.empty(N, dtype=torch.float32, device=dev)
            cabi.check(lib.osc_graph_assemble(top_idx_all.data_ptr(), top_sim_all.data_ptr(), 1, N, k,
                                              self.row_cap_val, nbr.data_ptr(), A.data_ptr(),
                                              W.data_ptr(), deg.data_ptr(), sd.data_ptr(),
                                              self.nnz.data_ptr(), scratch.data_ptr(), st),
                       "osc_graph_assemble")
            del Yn, hi, lo
        self._nbr, self._A, self._W, self._deg, self._sd = nbr, A, W, deg, sd
        if self.mode == "rows":
            self._Y = Y_local
            del Y_all
        else:
            self._Y = self._slab(Y_all)
            del Y_all
        self._U = self._Y.clone()
        torch.cuda.current_stream().synchronize()
        self.timings["graph_build_ms"] = 1000.0 * (time.time() - t0)

    def _resolve_halo(self) -> None:
        """Halo strategy of the rows partition: "allgather" (NCCL all-gather of p in front of every SpMM),
        "pull" (unique remote rows fetched over NVLink peer memory into a local halo block, C path only) or
        "fused" (remote rows read inside the SpMM, once per reference: measured 7x slower than the
        all-gather on kNN graphs of random anchors, kept for graphs with locality).  Auto: "pull" under NCCL,
        "allgather" otherwise."""
        import torch.distributed as dist

        nccl = self.world > 1 and dist.is_initialized() and dist.get_backend(self.group) == "nccl"
        if self.world <= 1:
            self.halo = "allgather"
        elif self._halo_arg is None:
            self.halo = "pull" if nccl else "allgather"
        else:
            self.halo = self._halo_arg
        if self.halo == "pull" and not nccl:
            self.halo = "allgather"  # the pull halo lives in the C path, which needs NCCL
        self._want_p2p = self.halo == "fused"

    def set_halo(self, halo) -> None:
        """Switch the rows-partition halo strategy on a built lattice (collective).  Accepts the strategy
        name or the legacy bool (True = fused P2P, False = all-gather)."""
        if isinstance(halo, bool):
            halo = "fused" if halo else "allgather"
        if halo not in (None, "allgather", "pull", "fused"):
            raise ValueError("halo must be None, 'allgather', 'pull' or 'fused'")
        self._halo_arg = halo
        self._resolve_halo()

    def close(self) -> None:
        """Release the peer-mapped buffers of the fused halo (collective; safe to call twice)."""
        if self._peers is not None:
            self._peers.close()
            self._peers = None
        if self._pull is not None:
            self._pull.close()
            self._pull = None
        if self._comm and self._comm_owned:
            import torch

            torch.cuda.synchronize()
            self._lib.osc_dist_comm_destroy(C.c_void_p(self._comm))
        self._comm, self._comm_owned = None, False

    def repartition(self, mode: str) -> None:
        """Switch between the row-block and the column-slab partition WITHOUT rebuilding the graph
        (one all-to-all transposes the state from the layout the build wants to the layout
        the solve wants).  Y and U move; the cached U* is dropped."""
        import torch
        import torch.distributed as dist

        if mode not in {"rows", "columns"}:
            raise ValueError("mode must be 'rows' or 'columns'")
        if mode == self.mode:
            return
        G, N, D = self.world, self.N, self.D
        if mode == "columns" and D % (4 * G) != 0:
            raise ValueError("mode='columns' needs D divisible by 4*world_size")
        nccl = G > 1 and dist.get_backend(self.group) == "nccl" and N % G == 0

        def to_columns(t):  # [n_local, D] -> [N, D/G]
            Dl = D // G
            if G == 1:
                return t
            if nccl:
                send = t.view(self.n_local, G, Dl).permute(1, 0, 2).contiguous()  # [G][n_local][Dl]
                recv = torch.empty((G, self.shard, Dl), dtype=t.dtype, device=t.device)
                dist.all_to_all_single(recv, send, group=self.group)
                return recv.view(N, Dl)
            full = gather_rows(t, N, self.group)
            return full[:, self.rank * Dl:(self.rank + 1) * Dl].contiguous()

        def to_rows(t):  # [N, D/G] -> [n_local, D]
            Dl = D // G
            if G == 1:
                return t
            if nccl:
                send = t.view(G, self.shard, Dl).contiguous()
                recv = torch.empty((G, self.n_local, Dl), dtype=t.dtype, device=t.device)
                dist.all_to_all_single(recv, send, group=self.group)
                return recv.permute(1, 0, 2).contiguous().view(self.n_local, D)
            src = t.contiguous()
            if dist.get_backend(self.group) == "gloo":
                src = src.cpu()
            parts = [torch.empty_like(src) for _ in range(G)]
            dist.all_gather(parts, src, group=self.group)
            full = torch.cat(parts, dim=1).to(t.device)
            return full[self.row0:self.row0 + self.n_local].contiguous()

        move = to_columns if mode == "columns" else to_rows
        self._Y, self._U = move(self._Y), move(self._U)
        self.mode = mode
        if mode == "columns":
            self.Dl = D // G
            self.c0 = self.rank * self.Dl
        self._Ustar = None

    # ---- C-ABI structs
    def _graph_struct(self, local: bool):
        g = self._cabi.Graph
        if local:
            r0, r1 = self.row0, self.row0 + self.n_local
            return g(1, self.n_local, self.k, 0, self._nbr[r0:r1].data_ptr() if self.n_local else 0,
                     self._A[r0:r1].data_ptr() if self.n_local else 0,
                     self._W[r0:r1].data_ptr() if self.n_local else 0,
                     self._deg[r0:r1].data_ptr() if self.n_local else 0,
                     self._sd[r0:r1].data_ptr() if self.n_local else 0)
        return g(1, self.N, self.k, 0, self._nbr.data_ptr(), self._A.data_ptr(), self._W.data_ptr(),
                 self._deg.data_ptr(), self._sd.data_ptr())

    def _params_struct(self):
        return self._cabi.Params(self.lamG, self.lamC, self.lamQ, self.lamP,
                                 1 if self._chain is not None else 0, 0)

    def _chain_struct(self):
        if self._chain is None:
            return None
        c = self._chain
        return self._cabi.Chain(c["n_rows"], c["nnz"], c["rows"].data_ptr(), c["rowptr"].data_ptr(),
                                c["col"].data_ptr(), c["Wp"].data_ptr(), c["Ap"].data_ptr(),
                                c["slot"].data_ptr())

    # ---- public API (subset of OscillinkLattice that makes sense for a sharded lattice)
    def set_query(self, psi, gates=None) -> None:
        import torch

        hpsi = np.asarray(psi, dtype=np.float32).reshape(-1)
        if hpsi.shape[0] != self.D:
            raise ValueError(f"psi must have D={self.D} entries, got {hpsi.shape[0]}")
        if gates is not None:
            hg = np.asarray(gates, dtype=np.float32)
            if hg.ndim != 1 or hg.shape[0] != self.N:
                raise ValueError("gates length mismatch N")
        self._dpsi = torch.as_tensor(hpsi.copy()).to(self._dev)
        if gates is not None:
            self._dB_all = torch.as_tensor(hg.copy()).to(self._dev)
            self._dB_loc = self._dB_all[self.row0:self.row0 + self.n_local]
        self._Ustar = None

    def add_chain(self, chain, lamP: float = 0.2, weights=None) -> None:
        from .lattice_api import OscillinkLattice

        if lamP < 0:
            raise ValueError("lamP must be >= 0")
        if any((c < 0 or c >= self.N) for c in chain):
            raise ValueError("chain indices out of bounds")
        if len(chain) < 2:
            raise ValueError("chain must contain at least two indices")
        if weights is not None and len(weights) != len(chain) - 1:
            raise ValueError("weights length must equal len(chain)-1")
        self._chain = OscillinkLattice._make_chain(self, chain, weights)  # same CSR builder
        self.lamP = float(lamP)
        self._chain_nodes = list(map(int, chain))
        self._Ustar = None
        if self._pull is not None:  # the chain's columns join the halo plan
            self._pull.close()
            self._pull = None

    def _psi_view(self):
        return self._dpsi if self.mode == "rows" else self._dpsi[self.c0:self.c0 + self.Dl].contiguous()

    # ---- NCCL communicator of the C-ABI solve
    def _nccl_comm(self):
        """ncclComm_t for osc_dist_* (0 if the C path cannot be used: no NCCL backend / binding failed).
        The library creates its own communicator (ncclGetUniqueId on rank 0, the 128-byte id broadcast
        through torch.distributed, ncclCommInitRank on every rank); if that fails, torch's own communicator
        of this group is borrowed (same NCCL library instance)."""
        import torch
        import torch.distributed as dist

        if self._comm is not None:
            return self._comm
        self._comm = 0
        if self.world <= 1 or not dist.is_initialized() or dist.get_backend(self.group) != "nccl":
            return self._comm
        lib = self._lib
        ok = torch.zeros(1, dtype=torch.int32, device=self._dev)
        try:
            box = [None]
            if self.rank == 0:
                buf = (C.c_ubyte * 128)()
                box[0] = bytes(buf) if lib.osc_dist_unique_id(buf) == 0 else None
            src = dist.get_global_rank(self.group, 0) if self.group is not None else 0
            dist.broadcast_object_list(box, src=src, group=self.group)
            if box[0] is not None:
                comm = C.c_void_p()
                idb = (C.c_ubyte * 128).from_buffer_copy(box[0])
                if lib.osc_dist_comm_init(idb, self.world, self.rank, C.byref(comm)) == 0 and comm.value:
                    self._comm, self._comm_owned = comm.value, True
                    ok += 1
        except Exception:
            pass
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)  # all ranks or none
        if int(ok.item()) == 0:
            if self._comm and self._comm_owned:
                lib.osc_dist_comm_destroy(C.c_void_p(self._comm))
            self._comm, self._comm_owned = 0, False
            try:  # borrow torch's communicator
                pg = self.group if self.group is not None else dist.distributed_c10d._get_default_group()
                self._comm = int(pg._get_backend(self._dev)._comm_ptr())
            except Exception:
                self._comm = 0
        return self._comm

    def _dist_struct(self):
        cabi = self._cabi
        rows = self.mode == "rows"
        pull = rows and self.world > 1 and self.halo == "pull"
        if pull and self._pull is None:
            self._pull = _PullHalo(self)
        ds = cabi.Dist(self._nccl_comm() or None, self.world, self.rank,
                       cabi.PART_ROWS if rows else cabi.PART_COLUMNS,
                       cabi.HALO_PULL if pull else cabi.HALO_ALLGATHER, self.N, self.shard,
                       None, None, None, None, 0, 0)
        if pull:
            pl = self._pull
            ds.d_peer_P, ds.P_block = pl.tab.data_ptr(), pl.block_ptr
            ds.halo_rows = pl.halo_rows.data_ptr() if pl.n_halo else None
            ds.halo_nbr, ds.n_halo, ds.halo_below = pl.nbr_local.data_ptr(), pl.n_halo, pl.n_below
        return ds

    def _use_c_path(self) -> bool:
        """The C-ABI solve: always on one GPU, and under NCCL unless the fused-P2P halo was asked for."""
        if self.world == 1:
            return True
        if self.mode == "rows" and self.halo == "fused":
            return False
        return bool(self._nccl_comm())

    def _c_args(self):
        rows = self.mode == "rows"
        g = self._graph_struct(local=rows)
        ch = self._chain_struct()
        if ch is not None and rows and self.world > 1 and self.halo == "pull":
            ch = self._pull.chain_struct(self._chain)  # column ids as rows of the block
        gates = self._dB_loc if rows else self._dB_all
        Dl = self.D if rows else self.Dl
        n_loc = self.n_local if rows else self.N
        return g, ch, gates, Dl, n_loc

    def _solve(self, mode_id: int, dt: float, tol: float, max_iters: int, jacobi: bool = True,
               warm_start: bool = True, inertia: float = 0.0):
        import torch

        cabi, lib = self._cabi, self._lib
        rows = self.mode == "rows"
        if self._use_c_path():
            ds = self._dist_struct()  # (collective on first use: communicator, halo plan)
            g, ch, gates, Dl, n_loc = self._c_args()
            prm = self._params_struct()
            psi = self._psi_view()
            X = torch.empty_like(self._Y)
            need = C.c_size_t(0)
            cabi.check(lib.osc_dist_pcg_workspace(C.byref(ds), n_loc, Dl, C.byref(need)), "osc_dist_pcg_workspace")
            ws = torch.empty(max(need.value, 256), dtype=torch.uint8, device=self._dev)
            it, res = C.c_int32(0), C.c_float(0.0)
            cabi.check(lib.osc_dist_pcg_solve(
                C.byref(ds), C.byref(g), C.byref(ch) if ch is not None else None, C.byref(prm), mode_id,
                float(dt), 1 if warm_start else 0, float(inertia), 1 if jacobi else 0, float(tol),
                int(max_iters), self._Y.data_ptr(), self._U.data_ptr(), psi.data_ptr(), gates.data_ptr(), Dl,
                X.data_ptr(), C.byref(it), C.byref(res), ws.data_ptr(), ws.numel(),
                torch.cuda.current_stream().cuda_stream), "osc_dist_pcg_solve")
            return X, int(it.value), float(res.value)
        # ---- no NCCL (gloo test rigs) or the fused-P2P halo: the phase entry points driven from Python
        use_p2p = rows and self._want_p2p
        if use_p2p and self._peers is None:
            self._peers = _PeerBuffers(self)
        if not use_p2p and self._peers is not None and rows:
            self._peers.close()  # halo strategy switched back to the all-gather
            self._peers = None
        X = self._peers.X[: self.n_local] if use_p2p else torch.empty_like(self._Y)
        Bv = torch.empty_like(self._Y)
        Dl = self.D if rows else self.Dl
        n_loc = self.n_local if rows else self.N
        dims = cabi.PcgDims(self.N, self.row0 if rows else 0, n_loc, Dl, 0)
        prm = self._params_struct()
        gates = self._dB_loc if rows else self._dB_all
        psi = self._psi_view()
        if n_loc > 0:
            cabi.check(lib.osc_pcg_setup(C.byref(dims), C.byref(prm), mode_id, float(dt),
                                         1 if warm_start else 0, float(inertia), self._Y.data_ptr(),
                                         self._U.data_ptr(), psi.data_ptr(), gates.data_ptr(),
                                         X.data_ptr(), Bv.data_ptr(),
                                         torch.cuda.current_stream().cuda_stream), "osc_pcg_setup")
        k = _NativeKernels(self, mode_id, dt, jacobi, X, Bv)
        it, res = pcg_schedule(k, mode=self.mode, tol=tol, max_iters=max_iters, group=self.group)
        if use_p2p:
            X = X.clone()  # the peer-mapped buffer is reused by the next solve
        return X, it, res

    def settle(self, dt: float = 1.0, max_iters: int = 12, tol: float = 1e-3, precond: str = "jacobi", *,
               warm_start: bool = True, inertia: float = 0.0) -> dict[str, Any]:
        import torch

        t0 = time.time()
        X, it, res = self._solve(self._cabi.MODE_SETTLE, dt, tol, max_iters, precond == "jacobi",
                                 warm_start, inertia)
        self._U = X
        torch.cuda.current_stream().synchronize()
        self.last = {"iters": int(it), "res": float(res), "t_ms": 1000.0 * (time.time() - t0)}
        return self.last

    def solve_Ustar(self, tol: float = 1e-4, max_iters: int = 64):
        if self._Ustar is None:
            X, it, res = self._solve(self._cabi.MODE_STATIONARY, 0.0, tol, max_iters)
            self._Ustar = X
            self.last_ustar = {"iters": int(it), "res": float(res), "converged": bool(res <= tol)}
        return self._Ustar

    def deltaH(self) -> float:
        """receipts.py:21-25 over the sharded state: <U-U*, M(U-U*)> summed over ranks."""
        import torch
        import torch.distributed as dist

        Us = self.solve_Ustar()
        if self._use_c_path():
            cabi, lib = self._cabi, self._lib
            ds = self._dist_struct()
            g, ch, gates, Dl, n_loc = self._c_args()
            prm = self._params_struct()
            need = C.c_size_t(0)
            cabi.check(lib.osc_dist_pcg_workspace(C.byref(ds), n_loc, Dl, C.byref(need)), "osc_dist_pcg_workspace")
            ws = torch.empty(max(need.value, 256), dtype=torch.uint8, device=self._dev)
            out = C.c_double(0.0)
            cabi.check(lib.osc_dist_delta_h(
                C.byref(ds), C.byref(g), C.byref(ch) if ch is not None else None, C.byref(prm),
                self._U.data_ptr(), Us.data_ptr(), gates.data_ptr(), Dl, C.byref(out), ws.data_ptr(),
                ws.numel(), torch.cuda.current_stream().cuda_stream), "osc_dist_delta_h")
            return float(np.float32(out.value))
        diff = self._U - Us
        if self.mode == "rows" and self._peers is not None and self._want_p2p:
            # fused halo: U - U* goes into the peer-mapped P block and the SpMM reads the peers' blocks
            k = _NativeKernels(self, self._cabi.MODE_STATIONARY, 0.0, True, self._peers.X[: self.n_local],
                               torch.empty_like(diff))
            k.P.copy_(diff)
            k.peer_sync()
            k.spmm(None)
        else:
            k = _NativeKernels(self, self._cabi.MODE_STATIONARY, 0.0, True, diff, torch.empty_like(diff))
            full = gather_rows(diff, self.N, self.group).contiguous() if self.mode == "rows" else diff
            k.spmm(full)
        tot = k.parts["pap"].sum().reshape(1)
        _allreduce(tot, dist.ReduceOp.SUM, self.group)
        return float(np.float32(tot.item()))

    def receipt(self) -> dict[str, Any]:
        """Light receipt (lattice.py:298-318,427-439): deltaH + solve statistics."""
        dH = self.deltaH()
        lu = getattr(self, "last_ustar", {})
        return {
            "deltaH_total": dH, "cg_iters": int(self.last.get("iters") or 0),
            "residual": float(self.last.get("res") or 0.0), "t_ms": float(self.last.get("t_ms") or 0.0),
            "meta": {"ustar_iters": int(lu.get("iters", 0)), "ustar_res": float(lu.get("res", 0.0)),
                     "ustar_converged": bool(lu.get("converged", True)),
                     "graph_build_ms": float(self.timings.get("graph_build_ms", 0.0)),
                     "avg_degree": float(int((self._A > 0).sum().item()) / max(self.N, 1)),
                     "world_size": self.world, "partition": self.mode},
        }

    def rows_of(self, t, ids):
        """Full-width rows `ids` (global row ids, 1-D int64 tensor, same on every rank) of a state tensor
        held in this lattice's partition (Y / U / U*), on every rank: [len(ids), D] fp32."""
        import torch
        import torch.distributed as dist

        ids = ids.to(self._dev).long()
        if self.world == 1:
            return t[ids]
        if self.mode == "rows":
            out = torch.zeros((ids.numel(), self.D), dtype=t.dtype, device=self._dev)
            mine = (ids >= self.row0) & (ids < self.row0 + self.n_local)
            if bool(mine.any()):
                out[mine] = t[ids[mine] - self.row0]
            dist.all_reduce(out, group=self.group)  # the other ranks add exact zeros
            return out
        part = t[ids].contiguous()  # [R, Dl]
        if dist.get_backend(self.group) == "gloo":
            part = part.cpu()
        parts = [torch.empty_like(part) for _ in range(self.world)]
        dist.all_gather(parts, part, group=self.group)
        return torch.cat(parts, dim=1).to(self._dev)

    # ---- host views for tests / callers
    def U_full(self) -> np.ndarray:
        """The settled state gathered on every rank as a host array (N, D)."""
        return self._gather_state(self._U)

    def Ustar_full(self) -> np.ndarray:
        return self._gather_state(self.solve_Ustar())

    def _gather_state(self, t) -> np.ndarray:
        import torch
        import torch.distributed as dist

        if self.mode == "rows":
            return gather_rows(t, self.N, self.group).cpu().numpy()
        if self.world == 1:
            return t.cpu().numpy()
        src = t.contiguous()
        if dist.get_backend(self.group) == "gloo":
            src = src.cpu()
        parts = [torch.empty_like(src) for _ in range(self.world)]
        dist.all_gather(parts, src, group=self.group)
        return torch.cat(parts, dim=1).cpu().numpy()
