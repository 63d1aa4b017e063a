"""ORACLE (test infrastructure, NOT product code) -- sparse restatement.

The reference materialises several dense N x N fp32 matrices and therefore stops
near N ~ 16-30k.  This module restates the same algorithm in the sparse form of
SURVEY.md Appendix A (ELL neighbour lists of width k instead of dense A / L_sym)
so the oracle can be run at sizes the reference cannot reach, and so the CUDA
path (which uses the same data layout) can be compared array by array.

Canonical neighbour rule (DESIGN.md "near-ties"): similarities are the exact
dot products of the fp32-normalised rows, rounded once to fp32 (an ideal
correctly-rounded sgemm), ranked by (similarity desc, index asc) exactly as
oscillink/core/graph.py:46-49 ranks the OpenBLAS result.  Whenever the k-th/(k+1)-th
gap of a row exceeds the reference's own sgemm noise (~1e-7) the two rules select
identical sets; tests/test_oracle_golden.py checks that on every fixture and
reports the smallest gap seen.

Parity: PINNED through oracle/dense.py -- every function here is compared with the
dense literal restatement (itself pinned to the real reference) in
tests/test_oracle_golden.py.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import
this module.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------- graph
def normalise_rows(Y):
    """graph.py:35 with a correctly rounded norm: fp32(sqrt(sum_d fp64(y)^2)) + 1e-12."""
    Y = np.ascontiguousarray(Y, dtype=F32)
    nrm = np.sqrt(np.einsum("ij,ij->i", Y.astype(np.float64), Y.astype(np.float64))).astype(F32)
    return (Y / (nrm[:, None] + F32(1e-12))).astype(F32)


def topk_canonical(Yn, k, rows=None, block=512):
    """graph.py:36-52 in canonical form.  Returns (idx[R,k] int64, sim[R,k] fp32,
    gap[R] fp32) for the requested rows (default: all): the k best columns by
    (fp32(exact dot) desc, index asc), diagonal excluded, and the margin between the
    k-th and (k+1)-th similarity (inf when N-1 == k)."""
    n = Yn.shape[0]
    rows = np.arange(n) if rows is None else np.asarray(rows)
    k = int(max(1, min(k, n - 1)))
    Y64 = Yn.astype(np.float64)
    idx = np.empty((len(rows), k), dtype=np.int64)
    sim = np.empty((len(rows), k), dtype=F32)
    gap = np.full(len(rows), np.inf, dtype=F32)
    col = np.arange(n)
    for s in range(0, len(rows), block):
        r = rows[s : s + block]
        S = (Y64[r] @ Y64.T).astype(F32)
        S[np.arange(len(r)), r] = -np.inf
        for t in range(len(r)):
            row = S[t]
            kk = min(k + 1, n - 1)
            part = np.argpartition(-row, kk - 1)[:kk]
            thr = row[part].min()
            pool = np.nonzero(row >= thr)[0]  # every column tied with the pool boundary
            order = pool[np.lexsort((col[pool], -row[pool]))]
            idx[s + t] = order[:k]
            sim[s + t] = row[order[:k]]
            if len(order) > k:
                gap[s + t] = row[order[k - 1]] - row[order[k]]
    return idx, sim, gap


def assemble(idx, sim, cap=1.0):
    """graph.py:50-52,64-65,77-83,87-92 on neighbour lists.

    idx/sim: directed top-k lists for ALL rows.  Returns ELL arrays with rows
    compacted left and columns ascending:
      nbr[N,k] int64 (-1 pad), A[N,k] fp32 capped weights, W[N,k] fp32 normalised
      weights, deg[N] int32, sd[N] fp32 (sqrt_deg)."""
    n, k = idx.shape
    keep = sim > 0  # graph.py:51 / :62
    # mutual test: i must appear in row j's kept list (graph.py:64)
    mutual = np.zeros_like(keep)
    for t in range(k):
        j = idx[:, t]
        back = (idx[j] == np.arange(n)[:, None]) & keep[j]
        mutual[:, t] = keep[:, t] & back.any(axis=1)
    nbr = np.where(mutual, idx, n)  # pad sorts last
    order = np.argsort(nbr, axis=1, kind="stable")
    nbr = np.take_along_axis(nbr, order, axis=1)
    a = np.take_along_axis(np.where(mutual, sim, F32(0)), order, axis=1).astype(F32)
    deg = mutual.sum(axis=1).astype(np.int32)
    nbr[nbr == n] = -1
    safe = np.where(nbr < 0, 0, nbr)
    # graph.py:77-83 (fp32, sequential left-to-right row sums)
    s = np.zeros(n, dtype=F32)
    for t in range(k):
        s = (s + a[:, t]).astype(F32)
    s = (s + F32(1e-12)).astype(F32)
    c = np.minimum(F32(1.0), (F32(cap) / s).astype(F32)).astype(F32)
    A = (a * np.sqrt((c[:, None] * c[safe]).astype(F32)).astype(F32)).astype(F32)
    A[nbr < 0] = 0
    # graph.py:87-90
    d = np.zeros(n, dtype=F32)
    for t in range(k):
        d = (d + A[:, t]).astype(F32)
    sd = np.sqrt(np.maximum(d, F32(1e-12))).astype(F32)
    inv = (F32(1.0) / sd).astype(F32)
    W = ((A * inv[:, None]).astype(F32) * inv[safe]).astype(F32)
    W[nbr < 0] = 0
    return nbr, A, W, deg, sd


def chain_rows(n, chain, weights=None):
    """graph.py:101-111 sparse: returns dict {u: {v: Wp_uv}} of the normalised path
    adjacency (max-merged weights, 1/sqrt(dp_u dp_v) scaling) plus raw Ap."""
    if weights is None:
        weights = [1.0] * max(0, len(chain) - 1)
    ap = {}
    for t in range(len(chain) - 1):
        u, v = int(chain[t]), int(chain[t + 1])
        w = F32(weights[t])
        if 0 <= u < n and 0 <= v < n:
            ap.setdefault(u, {})
            ap.setdefault(v, {})
            ap[u][v] = max(ap[u].get(v, F32(0)), w)
            ap[v][u] = max(ap[v].get(u, F32(0)), w)
    sd = {}
    for u, row in ap.items():
        d = F32(0)
        for v in sorted(row):
            d = F32(d + row[v])
        sd[u] = np.sqrt(max(d, F32(1e-12))).astype(F32)
    wp = {u: {v: F32(F32(row[v] * (F32(1) / sd[u])) * (F32(1) / sd[v])) for v in sorted(row)}
          for u, row in ap.items()}
    return wp, ap


class SparseLattice:
    """Sparse twin of oracle.dense.DenseLattice (same method names/returns)."""

    def __init__(self, Y, k=6, cap=1.0, lamG=1.0, lamC=0.5, lamQ=4.0):
        self.Y = np.ascontiguousarray(Y, dtype=F32).copy()
        self.U = self.Y.copy()
        self.N, self.D = self.Y.shape
        self.k = min(k, max(1, self.N - 1))
        if self.N <= 1:
            self.nbr = np.full((self.N, 1), -1, dtype=np.int64)
            self.A = np.zeros((self.N, 1), F32)
            self.W = np.zeros((self.N, 1), F32)
            self.deg = np.zeros(self.N, np.int32)
            self.sd = np.full(self.N, F32(1e-6))
            self.gap = np.full(self.N, np.inf, F32)
        else:
            idx, sim, self.gap = topk_canonical(normalise_rows(self.Y), self.k)
            self.nbr, self.A, self.W, self.deg, self.sd = assemble(idx, sim, cap)
        self.b = np.ones(self.N, dtype=F32)
        self.psi = np.zeros(self.D, dtype=F32)
        self.lamG, self.lamC, self.lamQ, self.lamP = F32(lamG), F32(lamC), F32(lamQ), F32(0)
        self.wp = None

    def set_query(self, psi, gates=None):
        self.psi = psi.astype(F32).copy()
        if gates is not None:
            self.b = gates.astype(F32).copy()

    def add_chain(self, chain, lamP=0.2, weights=None):
        self.wp, self.ap = chain_rows(self.N, chain, weights)
        self.lamP = F32(lamP)

    @property
    def nnz(self):
        return int(self.deg.sum())

    # ---- operators
    def _gather(self, X):
        acc = np.zeros_like(X)
        safe = np.where(self.nbr < 0, 0, self.nbr)
        for t in range(self.nbr.shape[1]):
            acc += self.W[:, t, None] * X[safe[:, t]]
        return acc

    def _chain_gather(self, X):
        acc = np.zeros_like(X)
        for u, row in self.wp.items():
            for v, w in row.items():
                acc[u] += w * X[v]
        return acc

    def _lamP_eff(self):
        return self.lamP if (self.wp is not None and self.lamP > 0) else F32(0)

    def apply_M(self, X, dt=None):
        """M x (dt None) or (I + dt M) x -- lattice.py:173-182 / :247-255 in sparse form."""
        lp = self._lamP_eff()
        diag = (self.lamG + self.lamC + lp) + self.lamQ * self.b
        offc, offp = self.lamC, lp
        if dt is not None:
            diag = F32(1) + F32(dt) * diag
            offc, offp = F32(dt) * offc, F32(dt) * offp
        out = diag.astype(F32)[:, None] * X - offc * self._gather(X)
        if lp > 0:
            out = out - offp * self._chain_gather(X)
        return out.astype(F32)

    def rhs(self):
        return (self.lamG * self.Y + self.lamQ * (self.b[:, None] * self.psi[None, :])).astype(F32)

    def _diag_base(self):
        return (self.lamG + self.lamQ * self.b + (self.lamP if self.wp is not None else F32(0))).astype(F32)

    def settle(self, dt=1.0, max_iters=12, tol=1e-3, jacobi=True, warm_start=True, inertia=0.0,
               trace=None):
        from .dense import pcg

        bvec = (self.U + F32(dt) * self.rhs()).astype(F32)
        md = (F32(1) + F32(dt) * self._diag_base()).astype(F32) if jacobi else None
        if not warm_start:
            x0 = self.Y
        else:
            w = float(max(0.0, min(1.0, inertia)))
            x0 = self.U if w <= 0.0 else ((1.0 - w) * self.Y + w * self.U).astype(F32)
        x, it, res = pcg(lambda X: self.apply_M(X, dt), bvec, x0, md, tol, max_iters)
        self.U = x.astype(F32)
        return {"iters": int(it), "res": float(res)}

    def stationary(self, tol=1e-4, max_iters=64):
        from .dense import pcg

        x, it, res = pcg(self.apply_M, self.rhs(), self.Y, self._diag_base(), tol, max_iters)
        return x.astype(F32), int(it), float(res)

    def delta_h(self, Ustar):
        diff = (self.U - Ustar).astype(F32)
        return float(np.sum(diff.astype(np.float64) * self.apply_M(diff).astype(np.float64)))

    # ---- receipts (Appendix A.5)
    def _edge_d2(self, V):
        """||V_i/(sd_i+1e-12) - V_j/(sd_j+1e-12)||^2 for every ELL slot."""
        Vn = (V / (self.sd[:, None] + F32(1e-12))).astype(F32)
        safe = np.where(self.nbr < 0, 0, self.nbr)
        out = np.zeros(self.nbr.shape, dtype=F32)
        for t in range(self.nbr.shape[1]):
            d = Vn - Vn[safe[:, t]]
            out[:, t] = np.einsum("ij,ij->i", d, d)
        out[self.nbr < 0] = 0
        return out

    def node_terms(self, Ustar):
        dy = self._edge_d2(self.Y)
        du = self._edge_d2(Ustar)
        coh = (F32(0.5) * self.lamC * self.A * (dy - du)).sum(axis=1).astype(F32)
        anchor = (self.lamG * np.sum((Ustar - self.Y) ** 2, axis=1)).astype(F32)
        q = Ustar - self.psi[None, :]
        query = (self.lamQ * self.b * np.sum(q * q, axis=1)).astype(F32)
        return coh, anchor, query

    def null_edges(self, Ustar, z_th=3.0):
        """receipts.py:70-82 sparse: mu/sigma over all N columns (zeros included),
        j* = neighbour with the largest residual, lowest column on ties."""
        R = (self.lamC * self.A * self._edge_d2(Ustar)).astype(F32)
        n = F32(self.N)
        mu = (R.sum(axis=1) / n).astype(F32)
        var = np.maximum((R * R).sum(axis=1) / n - mu * mu, 0)
        sg = (np.sqrt(var) + F32(1e-12)).astype(F32)
        out = []
        for i in range(self.N):
            if self.deg[i] == 0:
                continue
            t = int(np.argmax(R[i, : self.deg[i]]))
            r = R[i, t]
            z = (r - mu[i]) / sg[i]
            if r > 0 and z > z_th:
                out.append({"edge": [i, int(self.nbr[i, t])], "z": float(z), "residual": float(r)})
        return out

    def dense_A(self):
        A = np.zeros((self.N, self.N), dtype=F32)
        for t in range(self.nbr.shape[1]):
            m = self.nbr[:, t] >= 0
            A[np.nonzero(m)[0], self.nbr[m, t]] = self.A[m, t]
        return A
