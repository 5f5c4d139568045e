"""Host-side mirror of CTSM's clump decomposition for the multi-GPU layout.

Reference: src/main/decompInitMod.F90:96-161 (clumps dealt round-robin to processes, gridcells dealt to
clumps in `nsegspc` segments per clump) and decompMod.F90:349-424 (get_clump_bounds: proc-local 1-based
bounds).  One process per GPU = one "pe"; clump_pproc clumps per process.  The path has no exchange
step: a rank only ever touches the gridcells of its own clumps, so the only
multi-rank operation is the optional reduction of BalanceCheck's clump maxima for a global report.
Pure integer arithmetic; bit-exact against the oracle's restatement (tests/test_decomp.py).
"""
from __future__ import annotations

from typing import List

import numpy as np


def gridcell_to_clump(numg: int, nclumps: int, nsegspc: int = 35) -> np.ndarray:
    """cid (1-based) of every land gridcell ng = 1..numg, decompInitMod.F90:117-145."""
    ng = np.arange(1, numg + 1, dtype=np.int64)
    if float(np.float32(numg) / np.float32(nclumps)) < float(np.float32(nsegspc)):     # float() = REAL(4) in the reference
        return ((ng - 1) % nclumps + 1).astype(np.int32)
    rcid = ((ng - 1).astype(np.float64) / float(numg)) * float(nsegspc) * float(nclumps)
    return (rcid.astype(np.int64) % nclumps + 1).astype(np.int32)


def clump_owner(nclumps: int, npes: int) -> np.ndarray:
    """owner pe (0-based) of clump n = 1..nclumps: round robin, decompInitMod.F90:96-114."""
    return ((np.arange(1, nclumps + 1) - 1) % npes).astype(np.int32)


def rank_gridcells(numg: int, npes: int, clump_pproc: int, rank: int, nsegspc: int = 35) -> List[np.ndarray]:
    """For one rank: the global gridcell numbers (1-based, ascending) of each of its clumps, in clump order."""
    nclumps = npes * clump_pproc
    cid = gridcell_to_clump(numg, nclumps, nsegspc)
    own = clump_owner(nclumps, npes)
    mine = [n for n in range(1, nclumps + 1) if own[n - 1] == rank]
    return [np.nonzero(cid == n)[0].astype(np.int64) + 1 for n in mine]


def balanced_slabs(cost_per_gridcell, nparts: int) -> np.ndarray:
    """Contiguous gridcell slabs of (nearly) equal cost for the GPU layout of SURVEY.md 8e: the reference balances clumps
    by gridcell COUNT (decompInitMod.F90:117-145), but the cost of the hot path is proportional to the number of
    exposed-vegetation patches (CanopyFluxes is >90 % of the step), which varies by an order of magnitude between
    gridcells of a real surface dataset.  Returns edges e (len nparts+1, e[0] = 0, e[-1] = numg): part k owns gridcells
    e[k]+1 .. e[k+1] (1-based).  Slabs are contiguous so that each part's landunits, columns and patches are contiguous
    index ranges (initGridCellsMod nests g > l > c > p) and every routine can be called with plain clump bounds.
    Greedy prefix-sum cut: part k ends at the first gridcell where the running cost reaches (k+1)/nparts of the total;
    the imbalance is bounded by the cost of one gridcell."""
    c = np.asarray(cost_per_gridcell, dtype=np.float64)
    numg = len(c)
    nparts = max(1, int(nparts))
    cum = np.concatenate([[0.0], np.cumsum(c)])
    total = cum[-1]
    edges = np.zeros(nparts + 1, dtype=np.int64)
    edges[-1] = numg
    for k in range(1, nparts):
        if total > 0:
            e = int(np.searchsorted(cum, total * k / nparts, side="left"))
        else:
            e = (numg * k) // nparts
        edges[k] = min(max(e, edges[k - 1]), numg)
    return edges


def slab_imbalance(cost_per_gridcell, edges) -> float:
    """max part cost / mean part cost (1.0 = perfect)."""
    c = np.asarray(cost_per_gridcell, dtype=np.float64)
    parts = np.array([c[edges[k]:edges[k + 1]].sum() for k in range(len(edges) - 1)])
    return float(parts.max() / parts.mean()) if parts.mean() > 0 else 1.0


def reduce_balance_report(max_abs, dist=None):
    """Global maxima of the per-rank BalanceCheck report (reporting diagnostic only; the reference has no
    such collective, BalanceCheckMod.F90 is clump-local).  `dist` is torch.distributed or None."""
    import torch
    t = torch.tensor(list(max_abs), dtype=torch.float64)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.cpu().tolist()
