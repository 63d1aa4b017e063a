"""Dev: the SpMM at the row width a column slab has on G GPUs (D/G floats), on ONE GPU:
    N=10000000 D=48 K=10 python tools/dev_spmm_rows.py      (D = 384/8)
Builds the lattice at that width and times settle + the phase kernels (CUDA events)."""
import os, sys, json
import torch
sys.path.insert(0, ".")
from oscillink_b200.sharded_api import ShardedLattice, _NativeKernels
from oscillink_b200 import _cabi

if os.environ.get("L2G"):  # experiment: cudaLimitMaxL2FetchGranularity (0x05) = 32 / 64 / 128 bytes
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    torch.cuda.init(); torch.zeros(1, device="cuda")
    val = ctypes.c_size_t(0)
    rt.cudaDeviceGetLimit(ctypes.byref(val), 5); before = val.value
    rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(int(os.environ["L2G"])))
    rt.cudaDeviceGetLimit(ctypes.byref(val), 5)
    print("L2 fetch granularity: before", before, "set rc", rc, "now", val.value)
N = int(os.environ.get("N", "10000000")); D = int(os.environ.get("D", "48")); K = int(os.environ.get("K", "10"))
g = torch.Generator(device="cuda"); g.manual_seed(1)
Y = torch.randn((N, D), generator=g, device="cuda")
lat = ShardedLattice(Y, N, kneighbors=K)
psi = Y[:32].mean(0); psi = (psi / psi.norm()).cpu().numpy()
lat.set_query(psi)
st = lat.settle()
torch.cuda.synchronize()
X = lat._U
kf = _NativeKernels(lat, _cabi.MODE_SETTLE, 1.0, True, X, torch.zeros_like(X))
kf.residual0(X)
def t_of(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
ms = t_of(lambda: kf.spmm(kf.P))
nnz = float(lat.nnz.item()); V = N * D * 4.0
alg = (nnz / N + 2.0) * V + 8.0 * nnz
print(json.dumps({"N": N, "D": D, "k": K, "settle": st, "spmm_ms": ms, "spmm_gbs": alg / ms / 1e6, "nnz_per_row": nnz / N,
                  "engine": lat.engine_used}))
