"""The product package must never import, link or execute anything under oracle/ and must not
carry a CPU fallback for the numeric path."""
import ast
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "oscillink_b200")


def _py_files():
    for d, _, files in os.walk(PKG):
        for f in files:
            if f.endswith(".py"):
                yield os.path.join(d, f)


def test_product_never_imports_oracle():
    for path in _py_files():
        tree = ast.parse(open(path).read())
        for node in ast.walk(tree):
            mods = []
            if isinstance(node, ast.Import):
                mods = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                mods = [node.module or ""]
            for m in mods:
                assert not m.split(".")[0] == "oracle", f"{path} imports {m}"
        assert "/root/reference" not in open(path).read()


def test_constructing_without_cuda_fails_loudly():
    import numpy as np
    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from oscillink_b200 import OscillinkLattice

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        OscillinkLattice(np.zeros((4, 3), dtype=np.float32))


def test_validation_errors_precede_device_use():
    """ValueErrors of lattice.py:45-53 are raised before any device work (so they also fire here)."""
    import numpy as np
    import pytest

    from oscillink_b200 import OscillinkLattice

    with pytest.raises(ValueError):
        OscillinkLattice([[1.0, 2.0]])
    with pytest.raises(ValueError):
        OscillinkLattice(np.zeros((4, 3), dtype=np.float32), kneighbors=0)
    with pytest.raises(ValueError):
        OscillinkLattice(np.zeros((4, 3), dtype=np.float32), lamG=0.0)
    with pytest.raises(ValueError):
        OscillinkLattice(np.zeros((4, 3), dtype=np.float32), lamC=-1.0)
