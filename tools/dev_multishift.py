"""Dev check (CPU, oracle only): the multi-shift CG that the batched kernel's fast path runs
(settle and stationary systems are shifts of one another when U=Y and the gates are uniform)
against the two separate PCG solves of the sparse oracle."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.sparse import SparseLattice

F = np.float32

def multishift(o, dt=1.0, tol_s=1e-3, max_s=12, tol_u=1e-4, max_u=64):
    lamG, lamC, lamQ = o.lamG, o.lamC, o.lamQ
    Y = o.Y
    diag = F(lamG + lamC + lamQ)
    im = F(1.0) / (F(lamG + lamQ) + F(1e-12))
    sigma = F(1.0) / F(dt)
    rhs = o.rhs()
    g0 = o._gather(Y)
    R = (rhs - (diag * Y - lamC * g0)).astype(F)
    Xu, Xs = Y.copy(), Y.copy()
    P = (im * R).astype(F)
    PS = R.copy()
    G = np.zeros_like(Y)
    rr = np.einsum("ij,ij->j", R, R).astype(F)
    rz = (im * rr).astype(F)
    D = Y.shape[1]
    zeta = np.ones(D, F); zeta_p = np.ones(D, F); a_prev = np.ones(D, F); b_prev = np.zeros(D, F)
    beta = np.zeros(D, F)
    Ts = Tu = None
    k = 0
    out = {}
    while True:
        k += 1
        g = o._gather(R)
        G = (im * g + beta[None, :] * G).astype(F)
        AP = (diag * P - lamC * G).astype(F)
        pap = np.einsum("ij,ij->j", P, AP).astype(F)
        alpha = (rz / (pap + F(1e-18))).astype(F)
        a = (alpha * im).astype(F)
        num = zeta * zeta_p * a_prev
        den = a * b_prev * (zeta_p - zeta) + zeta_p * a_prev * (F(1) + sigma * a)
        zn = np.where(den != 0, num / np.where(den != 0, den, 1), zeta).astype(F)
        ratio = np.where(zeta != 0, zn / np.where(zeta != 0, zeta, 1), 0).astype(F)
        a_s = (a * ratio).astype(F)
        if Tu is None:
            Xu = (Xu + alpha[None, :] * P).astype(F)
        if Ts is None:
            Xs = (Xs + a_s[None, :] * PS).astype(F)
        R = (R - alpha[None, :] * AP).astype(F)
        rr_new = np.einsum("ij,ij->j", R, R).astype(F)
        rzn = (im * rr_new).astype(F)
        beta = (rzn / (rz + F(1e-18))).astype(F)
        b_s = (beta * ratio * ratio).astype(F)
        res_u = float(np.sqrt(rr_new.max()))
        res_s = float(np.sqrt(((F(dt) * zn) ** 2 * rr_new).max()))
        if Ts is None and (res_s <= tol_s or k >= max_s):
            Ts = k; out["res_s"] = res_s
            out["t1"] = ((Y - Xs) * sigma - zn[None, :] * R).astype(F)
        if Tu is None and (res_u <= tol_u or k >= max_u):
            Tu = k; out["res_u"] = res_u; out["Ru"] = R.copy()
        if Ts is not None and Tu is not None:
            break
        P = (im * R + beta[None, :] * P).astype(F)
        PS = (zn[None, :] * R + b_s[None, :] * PS).astype(F)
        zeta_p, zeta, a_prev, b_prev, rz = zeta, zn, a, beta, rzn
    d = (Xs - Xu).astype(np.float64)
    dh = float(np.sum(d * (out["t1"].astype(np.float64) + out["Ru"].astype(np.float64))))
    return Xs, Xu, Ts, Tu, out["res_s"], out["res_u"], dh

if __name__ == "__main__":
    N, D, k = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (1200, 384, 8)))
    for seed in range(int(sys.argv[4]) if len(sys.argv) > 4 else 2):
        rs = np.random.RandomState(seed)
        Y = rs.randn(N, D).astype(F)
        psi = Y[:32].mean(0); psi = (psi / (np.linalg.norm(psi) + 1e-12)).astype(F)
        o = SparseLattice(Y, k=k); o.set_query(psi)
        t = time.time()
        Xs, Xu, Ts, Tu, rs_, ru_, dh = multishift(o)
        st = o.settle()
        us, itu, resu = o.stationary()
        dh0 = o.delta_h(us)
        eU = np.linalg.norm(Xs - o.U) / np.linalg.norm(o.U)
        eS = np.linalg.norm(Xu - us) / np.linalg.norm(us)
        print(f"seed {seed}: iters ms ({Ts},{Tu}) ref ({st['iters']},{itu}) res_s {rs_:.4e}/{st['res']:.4e} "
              f"res_u {ru_:.4e}/{resu:.4e} relU {eU:.2e} relU* {eS:.2e} dH {dh:.4f}/{dh0:.4f} rel {abs(dh-dh0)/abs(dh0):.2e}")
