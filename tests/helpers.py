"""Shared test helpers: module loading and the parity metrics used across the suite."""
import importlib
import importlib.util
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_synthetic():
    """`mr-mt3_b200/synthetic.py` by path: no CUDA library needed (CPU tests use it too)."""
    spec = importlib.util.spec_from_file_location(
        "mrmt3_synthetic", os.path.join(ROOT, "mr-mt3_b200", "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def package():
    return importlib.import_module("mr-mt3_b200")


def golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name))


def logmel_rel_err(a, b):
    """The frontend parity metric: |a-b| / max(|b|, 1), max over elements."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0)))


def first_divergence(a, b):
    """Index of the first differing position of two 1-D int sequences (len if none)."""
    a = np.asarray(a)
    b = np.asarray(b)
    n = min(len(a), len(b))
    d = np.nonzero(a[:n] != b[:n])[0]
    return int(d[0]) if len(d) else n


def top2_margin(logits):
    v = torch.as_tensor(logits).topk(2, dim=-1).values
    return (v[..., 0] - v[..., 1])
