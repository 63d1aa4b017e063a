"""Sampled-row parity checks for lattices the dense reference cannot hold (N = 1M / 10M; SURVEY 8c:
"4096 random rows' neighbour sets + a strided sample of U* rows").

CHECKER code (bench.py `parity_sample`, tests/test_gpu_large.py) -- independent of the product kernels:
everything below is plain torch fp64 on the rows each rank holds.

* `exact_topk`     exhaustive canonical top-k of chosen rows against ALL N rows: rows normalised as
                   graph.py:35 (fp32 division by fp32(norm) + 1e-12), similarity = fp64 dot rounded once to
                   fp32 (an ideal sgemm, graph.py:36), ranked by (similarity desc, index asc) (graph.py:46-49),
                   diagonal excluded (graph.py:37).
* `mutual_sets`    the exact mutual-kNN neighbour set of every sampled row (graph.py:50-52,64-65): top-k of
                   the sample and of every neighbour of the sample.
* `operator_residual_rows`  fp64 evaluation of rows of  b - A u  for the settle / stationary systems
                   (lattice.py:171-184,245-256) from the lattice's ELL graph, for a strided row sample.

Multi-GPU: every rank passes its row block; queries are exchanged with all-reduce / all-gather over the
given process group, every rank returns the same result.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def _world(group):
    return dist.get_world_size(group) if dist.is_initialized() else 1


def _normalised(Yc: torch.Tensor) -> torch.Tensor:
    den = Yc.double().pow(2).sum(dim=1).sqrt().float() + 1e-12
    return Yc / den[:, None]


def _lex_topk(vals: torch.Tensor, ids: torch.Tensor, kk: int):
    """vals/ids: [C, R] candidates per query (columns).  Best kk by (value desc, id asc)."""
    o = torch.sort(ids, dim=0, stable=True)
    ids, vals = o.values, torch.gather(vals, 0, o.indices)
    o = torch.sort(vals, dim=0, descending=True, stable=True)
    vals, ids = o.values, torch.gather(ids, 0, o.indices)
    return vals[:kk], ids[:kk]


def exact_topk(Y_local: torch.Tensor, row0: int, N: int, rows: torch.Tensor, k: int, group=None,
               chunk: int = 1 << 20, margin: int = 8):
    """(idx[R,k] int64, sim[R,k] fp32, gap[R] fp32) for the global row ids `rows` (same tensor on every
    rank).  A block keeps k+margin candidates per query, so the result is canonical unless more than
    `margin` columns tie EXACTLY with a row's k-th similarity (duplicate anchors)."""
    dev = Y_local.device
    n_loc, D = Y_local.shape
    rows = rows.to(dev).long()
    R = rows.numel()
    k = max(1, min(k, N - 1))
    kk = min(k + margin, N - 1)
    # queries: owners contribute their normalised rows, the others zeros (x + 0 is exact)
    Q = torch.zeros((R, D), dtype=torch.float32, device=dev)
    mine = (rows >= row0) & (rows < row0 + n_loc)
    if bool(mine.any()):
        Q[mine] = _normalised(Y_local[rows[mine] - row0])
    if _world(group) > 1:
        dist.all_reduce(Q, group=group)
    Qd = Q.double().t().contiguous()  # [D, R]
    best_v = torch.full((0, R), 0.0, dtype=torch.float32, device=dev)
    best_i = torch.zeros((0, R), dtype=torch.int64, device=dev)
    qcol = torch.arange(R, device=dev)
    for c0 in range(0, n_loc, chunk):
        c1 = min(n_loc, c0 + chunk)
        S = (_normalised(Y_local[c0:c1]).double() @ Qd).float()  # [c, R], rounded once
        own = (rows >= row0 + c0) & (rows < row0 + c1)
        if bool(own.any()):
            S[rows[own] - row0 - c0, qcol[own]] = float("-inf")
        t = min(kk, c1 - c0)
        v, i = torch.topk(S, t, dim=0)
        best_v, best_i = _lex_topk(torch.cat([best_v, v]), torch.cat([best_i, i + (row0 + c0)]), kk)
        del S
    if best_v.shape[0] < kk:  # fewer local rows than candidates: pad so that all_gather shapes agree
        pad = kk - best_v.shape[0]
        best_v = torch.cat([best_v, torch.full((pad, R), float("-inf"), device=dev)])
        best_i = torch.cat([best_i, torch.full((pad, R), N, dtype=torch.int64, device=dev)])
    G = _world(group)
    if G > 1:
        allv = torch.empty((G * kk, R), dtype=torch.float32, device=dev)
        alli = torch.empty((G * kk, R), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allv, best_v.contiguous(), group=group)
        dist.all_gather_into_tensor(alli, best_i.contiguous(), group=group)
        best_v, best_i = _lex_topk(allv, alli, kk)
    idx = best_i[:k].t().contiguous()
    sim = best_v[:k].t().contiguous()
    gap = (best_v[k - 1] - best_v[k]) if kk > k else torch.full((R,), float("inf"), device=dev)
    return idx, sim, gap


def mutual_sets(Y_local: torch.Tensor, row0: int, N: int, sample: torch.Tensor, k: int, group=None,
                chunk: int = 1 << 20):
    """Exact mutual-kNN neighbour lists (ascending ids) of the sampled rows + the smallest k/(k+1)
    similarity gap met on the way (a gap below ~1e-7 is where the reference's own sgemm noise decides)."""
    idx_s, sim_s, gap_s = exact_topk(Y_local, row0, N, sample, k, group, chunk)
    hop = torch.unique(idx_s.reshape(-1))
    idx_h, sim_h, gap_h = exact_topk(Y_local, row0, N, hop, k, group, chunk)
    pos = {int(r): t for t, r in enumerate(hop.tolist())}
    idx_s_h, sim_s_h = idx_s.cpu(), sim_s.cpu()
    idx_h_h, sim_h_h = idx_h.cpu(), sim_h.cpu()
    out = []
    for t, i in enumerate(sample.tolist()):
        want = []
        for j, s in zip(idx_s_h[t].tolist(), sim_s_h[t].tolist()):
            u = pos[int(j)]
            back = (idx_h_h[u] == i) & (sim_h_h[u] > 0)
            if s > 0 and bool(back.any()):
                want.append(int(j))
        out.append(sorted(want))
    min_gap = float(torch.minimum(gap_s.min(), gap_h.min()).item())
    return out, min_gap, int(hop.numel())


def compare_neighbour_sets(nbr_rows: torch.Tensor, want: list[list[int]]) -> int:
    """nbr_rows: the lattice's ELL rows [R, k] (-1 padded) of the sampled rows.  Number of mismatching rows."""
    bad = 0
    for t, row in enumerate(nbr_rows.cpu().tolist()):
        got = [int(j) for j in row if j >= 0]
        if got != want[t]:
            bad += 1
    return bad


def operator_residual_rows(rows: torch.Tensor, u_rows, y_rows, u0_rows, nbr, W, psi, lam, *, settle: bool,
                           dt: float = 1.0, gates_rows=None, lamP_eff: float = 0.0):
    """fp64 rows of  b - A u  for `rows` (global ids; none of them may be a chain node: with a chain and
    lamP > 0 the reference's L_path = I - W_path spans all N rows, so every OFF-chain row just gains
    lamP * x_i -- pass that lamP as `lamP_eff`, graph.py:101-111 / SURVEY a6).

    u_rows(ids) / y_rows(ids) / u0_rows(ids): callables returning the [len(ids), D] fp32 rows of the
    solution, the anchors and the settle start state on THIS device (full D columns).
    nbr / W: the lattice's ELL graph [N, k] (global ids, -1 padded).  lam = (lamG, lamC, lamQ).
    Returns r [len(rows), D] fp64."""
    lamG, lamC, lamQ = (float(v) for v in lam)
    rows = rows.long()
    nb = nbr[rows].long()  # [R, k]
    w = W[rows].double()
    valid = nb >= 0
    flat = torch.where(valid, nb, torch.zeros_like(nb)).reshape(-1)
    u = u_rows(rows).double()
    un = u_rows(flat).double().reshape(nb.shape[0], nb.shape[1], -1)
    gath = (w[:, :, None] * un * valid[:, :, None]).sum(dim=1)
    b = torch.ones(rows.numel(), dtype=torch.float64, device=u.device) if gates_rows is None \
        else gates_rows(rows).double()
    y = y_rows(rows).double()
    rhs = lamG * y + lamQ * b[:, None] * psi.double()[None, :]
    Mu = (lamG + lamC + float(lamP_eff)) * u + lamQ * b[:, None] * u - lamC * gath
    if settle:
        return (u0_rows(rows).double() + dt * rhs) - (u + dt * Mu)
    return rhs - Mu
