"""Shared helpers for the test-suite (unique module name: an unrelated ``tests`` package is installed in this image)."""

import glob
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden_module_files():
    return sorted(glob.glob(os.path.join(GOLDEN, "modules_*.npz")))


def golden_ids():
    return [os.path.basename(p)[len("modules_"):-4] for p in golden_module_files()]


def rel_err(a, b):
    """max|a-b| / max|b|: the gradient parity measure (SURVEY.md §7 hard part 5)."""
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    if b.size == 0:
        return 0.0
    scale = max(float(np.max(np.abs(b))), 1e-300)
    return float(np.max(np.abs(a - b))) / scale
