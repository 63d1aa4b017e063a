"""Shared loaders for the parity tests."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with open(os.path.join(GOLDEN, f"{name}.json")) as f:
        g = json.load(f)
    z = dict(np.load(os.path.join(GOLDEN, f"{name}.npz")))
    return g, z


def artefacts():
    with open(os.path.join(GOLDEN, "reference_artefacts.json")) as f:
        raw = json.load(f)
    out = {}
    for src, cases in raw.items():
        if src.startswith("_"):
            continue
        for name, vals in cases.items():
            out[name] = dict(vals, _source=src)
    return out


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-30)


def summarise(X):
    X = np.asarray(X, dtype=np.float64)
    return {"fro": float(np.sqrt((X * X).sum())), "colnorm_head": np.sqrt((X * X).sum(axis=0))[:8].tolist(),
            "sum": float(X.sum())}
