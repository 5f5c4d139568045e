"""Shared helpers for the parity tests (test infrastructure)."""
import ctypes as C

import numpy as np

from ctsm_b200 import abi


def relerr(a, b, floor_frac=1e-12):
    """max elementwise |a-b| / max(|b|, floor) with floor = floor_frac * max|b| (guards exact zeros)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    scale = float(np.max(np.abs(b))) if b.size else 0.0
    floor = max(scale * floor_frac, 1e-300)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


def to_device(arrays):
    import torch
    return {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in arrays.items()}


def to_host(tensors):
    return {k: v.cpu().numpy() for k, v in tensors.items()}


def copy_state(S):
    return {k: v.copy() for k, v in S.items()}


def group_arrays(S, group):
    return {fs.name: S[fs.name] for fs in abi.FIELDS[group]}
