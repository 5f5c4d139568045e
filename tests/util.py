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


def compare_step_fields(sg, S, got, ref, loose_p, groups, rtol=1e-10, floor_frac=1e-2):
    """Compare every output field of `groups` between a CUDA step and the oracle step, leaving out the columns (and their
    patches / gridcells) that own a patch in `loose_p` (iteration-count ties, capped or ill-conditioned canopy patches: see
    tests/test_gpu_canopy.py).  Integers must be identical; reals are judged relative to max(|value|, floor_frac x field
    range) because the routines after CanopyFluxes form differences of its outputs.  Returns the worst error per field."""
    loose_c = np.zeros(sg.ncol, dtype=bool)
    loose_c[S["column"][loose_p] - 1] = True
    loose_g = np.zeros(sg.ngrc, dtype=bool)
    loose_g[sg.col_gridcell[loose_c] - 1] = True
    skip_of = {"PATCH": loose_c[S["column"] - 1], "COL": loose_c, "GRC": loose_g}
    worst = {}
    for g in groups:
        for fs in abi.FIELDS[g]:
            if fs.intent == "IN" or fs.sub not in skip_of or fs.name.startswith("err"):
                continue
            a, b = got[fs.name], ref[fs.name]
            keep = ~skip_of[fs.sub]
            if fs.ctype == "int":
                assert np.array_equal(a[..., keep], b[..., keep]), fs.name
                continue
            fin = np.abs(b) < 1e30
            assert np.array_equal(fin, np.abs(a) < 1e30), fs.name
            bb, aa = np.where(fin, b, 0.0)[..., keep], np.where(fin, a, 0.0)[..., keep]
            scale = float(np.max(np.abs(bb))) if bb.size else 0.0
            e = float(np.max(np.abs(aa - bb) / np.maximum(np.abs(bb), floor_frac * scale + 1e-300))) if bb.size else 0.0
            worst[fs.name] = max(worst.get(fs.name, 0.0), e)
    bad = {k: v for k, v in worst.items() if not v <= rtol}
    assert not bad, bad
    return worst
