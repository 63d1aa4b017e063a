"""In-tree nvcc build of the C-ABI library (sm_100a only).

    python -m oscillink_b200.build            # build if stale
    python -m oscillink_b200.build --force

Produces oscillink_b200/_lib/libosc_b200.so (git-ignored; travels to the GPU box with the
gpurun snapshot).  cudart is linked statically so the library also loads on a box with no
driver, which is what the CPU-side symbol test relies on.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_lib")
LIB = os.path.join(OUT_DIR, "libosc_b200.so")
SOURCES = ["cabi.cu", "knn.cu", "knn_tc.cu", "graph.cu", "pcg.cu", "dist.cu", "receipt.cu", "batched.cu", "batched_ms.cu", "bundle.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _deps():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    files.append(os.path.join(HERE, "..", "include", "oscillink_b200.h"))
    return files


def _digest() -> str:
    import hashlib

    h = hashlib.sha256()
    for f in sorted(_deps()):
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


STAMP = os.path.join(OUT_DIR, "build.sha256")


def is_stale() -> bool:
    """Content-hash staleness (mtimes do not survive the snapshot copy to the GPU box)."""
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    return open(STAMP).read().strip() != _digest()


def _compile(src: str) -> str:
    obj = os.path.join(OUT_DIR, src.replace(".cu", ".o"))
    cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(OUT_DIR, src.replace(".cu", ".ptxas.log"))
    with open(log, "w") as f:
        f.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(_compile, SOURCES))
    cmd = [NVCC, "-shared", "-o", LIB, *objs, "-cudart", "static", "-lpthread", "-ldl", "-lrt"]  # NCCL: dlopen
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(STAMP, "w") as f:
        f.write(_digest())
    if verbose:
        print(f"built {LIB}")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
