"""Developer check for the tcgen05 kNN engine (run on the GPU box under `timeout`):
compares the candidate lists of the TC engine with the SIMT engine through the C ABI."""
import ctypes as C
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from oscillink_b200 import _cabi  # noqa: E402


def run(B, N, D, k, seed=0):
    lib = _cabi.load()
    dev = torch.device("cuda")
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    Y = torch.randn((B, N, D), generator=g, device=dev)
    Yn, hi, lo = torch.empty_like(Y), torch.empty_like(Y), torch.empty_like(Y)
    st = torch.cuda.current_stream().cuda_stream
    _cabi.check(lib.osc_normalize_rows(Y.data_ptr(), B * N, D, Yn.data_ptr(), hi.data_ptr(), lo.data_ptr(), st))
    kc = min(k + 4, N - 1)
    outs = {}
    for name, eng in (("simt", _cabi.KNN_SIMT), ("tc", _cabi.KNN_TC)):
        ci = torch.full((B, N, kc), -7, dtype=torch.int32, device=dev)
        cs = torch.zeros((B, N, kc), dtype=torch.float32, device=dev)
        torch.cuda.synchronize()
        t0 = time.time()
        _cabi.check(lib.osc_knn_candidates(Yn.data_ptr(), Yn.data_ptr(), hi.data_ptr(), lo.data_ptr(),
                                           hi.data_ptr(), lo.data_ptr(), B, N, 0, N, D, kc, eng,
                                           ci.data_ptr(), cs.data_ptr(), None, 0, st), name)
        torch.cuda.synchronize()
        outs[name] = (ci.cpu().numpy(), cs.cpu().numpy(), time.time() - t0)
    (ia, sa, ta), (ib, sb, tb) = outs["simt"], outs["tc"]
    same = np.array_equal(ia, ib)
    print(f"B={B} N={N} D={D} k={k}: idx equal={same} max|ds|={np.abs(sa - sb).max():.3e} "
          f"simt {ta*1e3:.2f} ms tc {tb*1e3:.2f} ms", flush=True)
    if not same:
        bad = np.argwhere(ia != ib)
        print("  mismatches:", len(bad), "first:", bad[:5].tolist())
        b, r, c = bad[0]
        print("  simt", ia[b, r], sa[b, r])
        print("  tc  ", ib[b, r], sb[b, r])
    return same


if __name__ == "__main__":
    ok = True
    for shape in [(1, 300, 64, 6), (2, 1200, 384, 8), (1, 97, 20, 5), (3, 640, 128, 8), (1, 5000, 768, 16),
                  (64, 1200, 384, 8)]:
        ok &= run(*shape)
    print("ALL OK" if ok else "MISMATCH")
