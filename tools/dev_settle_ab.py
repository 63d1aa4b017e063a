"""Dev: A/B of environment switches on the settle of one large lattice, in ONE process (same clocks, same graph):
    N=1000000 D=768 K=16 AB="OSC_PCG_FUSE_X=0,OSC_PCG_FUSE_X=1" python tools/dev_settle_ab.py
Every setting: 2 warm-up settles from U = Y, then the median of REPS timed ones (CUDA events)."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from oscillink_b200.sharded_api import ShardedLattice  # noqa: E402

N = int(os.environ.get("N", "1000000"))
D = int(os.environ.get("D", "768"))
K = int(os.environ.get("K", "16"))
REPS = int(os.environ.get("REPS", "7"))
settings = [s for s in os.environ.get("AB", "OSC_PCG_FUSE_X=0,OSC_PCG_FUSE_X=1").split(",") if s]
g = torch.Generator(device="cuda")
g.manual_seed(1)
Y = torch.randn((N, D), generator=g, device="cuda")
lat = ShardedLattice(Y, N, kneighbors=K)
psi = Y[:32].mean(0)
lat.set_query((psi / psi.norm()).cpu().numpy())


def one():
    lat._U.copy_(lat._Y)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    st = lat.settle()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b), st


out = {}
for rnd in range(2):  # ABAB: drift shows up as a difference between the rounds
    for s in settings:
        k, v = s.split("=")
        os.environ[k] = v
        for _ in range(2):
            one()
        ts = sorted(one()[0] for _ in range(REPS))
        st = one()[1]
        out.setdefault(s, []).append({"median_ms": ts[len(ts) // 2], "min_ms": ts[0], "iters": st["iters"],
                                      "res": st["res"]})
print(json.dumps({"N": N, "D": D, "k": K, "results": out}))

if os.environ.get("KERN"):  # per-kernel times of the phases (one launch each, live state)
    from oscillink_b200 import _cabi
    from oscillink_b200.sharded_api import _NativeKernels

    X = lat._U
    kf = _NativeKernels(lat, _cabi.MODE_SETTLE, 1.0, True, X, torch.zeros_like(X))
    kf.residual0(X)
    ones = torch.ones(D, dtype=torch.float32, device="cuda")

    def t_of(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    V = N * D * 4.0
    nnz = float(lat.nnz.item())
    rows = {"pcg_spmm": (t_of(lambda: kf.spmm(kf.P)), (nnz / N + 2.0) * V + 8.0 * nnz),
            "pcg_update(r only)": (t_of(lambda: kf.update(ones, ones, with_x=False)), 3.0 * V),
            "pcg_update(x, r)": (t_of(lambda: kf.update(ones, ones)), 6.0 * V),
            "pcg_pupdate": (t_of(lambda: kf.pupdate(ones, ones)), 3.0 * V),
            "pcg_pupdate_x": (t_of(lambda: kf.pupdate_x(ones, ones, ones)), 5.0 * V),
            "pcg_pupdate_x(last)": (t_of(lambda: kf.pupdate_x(ones, ones, ones, last=True)), 3.0 * V)}
    print(json.dumps({k: {"ms": round(v[0], 4), "GB/s": round(v[1] / v[0] / 1e6, 1)} for k, v in rows.items()}))
