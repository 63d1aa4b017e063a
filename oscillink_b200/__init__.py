"""oscillink_b200 -- B200-native lattice-settle hot path behind the Oscillink API.

    from oscillink_b200 import OscillinkLattice      # drop-in for oscillink.OscillinkLattice

The numeric path (mutual-kNN build, Jacobi-PCG settle, deltaH / null-point receipts) runs as
hand-written sm_100a CUDA kernels behind the C ABI in include/oscillink_b200.h.  Importing
the package is cheap and GPU-free; the native library is loaded on first use.
"""
from __future__ import annotations

__version__ = "0.1.0"

from .receipts import verify_receipt, verify_receipt_mode  # noqa: F401


def __getattr__(name):  # lazy: torch / the native library load only when the API is touched
    if name in {"OscillinkLattice", "Oscillink", "json_line_logger"}:
        from . import lattice_api

        return getattr(lattice_api, name)
    if name in {"BatchedLattices", "settle_host_batch"}:
        from . import batched_api

        return getattr(batched_api, name)
    if name == "compute_diffusion_gates":
        from . import diffusion

        return diffusion.compute_diffusion_gates
    if name in {"ShardedLattice"}:
        from . import sharded_api

        return getattr(sharded_api, name)
    raise AttributeError(name)


__all__ = ["OscillinkLattice", "Oscillink", "BatchedLattices", "settle_host_batch", "ShardedLattice", "compute_diffusion_gates",
           "verify_receipt",
           "verify_receipt_mode", "json_line_logger", "__version__"]
