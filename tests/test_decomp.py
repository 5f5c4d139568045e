"""Clump decomposition (decompInitMod.F90:96-161) and the world_size-2 gloo path of the multi-GPU layout:
every gridcell is owned by exactly one rank, and running the oracle step per rank on its own clumps gives
bit-identical results to the undecomposed run (the reference's ERP/PEM invariant)."""
import os
import sys

import numpy as np
import pytest

from ctsm_b200 import decomp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("numg,npes,cpp", [(5500, 8, 1), (21000, 8, 4), (100, 2, 3), (7, 2, 2), (336000, 8, 16)])
def test_partition_is_exact(numg, npes, cpp):
    seen = np.zeros(numg + 1, dtype=np.int32)
    for r in range(npes):
        for cells in decomp.rank_gridcells(numg, npes, cpp, r):
            assert np.all(np.diff(cells) > 0)
            seen[cells] += 1
    assert np.all(seen[1:] == 1)
    # gridcells are dealt in nsegspc segments per clump when there are enough of them (:117-123)
    cid = decomp.gridcell_to_clump(numg, npes * cpp)
    if numg / (npes * cpp) >= 35:
        nseg = 1 + int(np.sum(np.diff(cid) != 0))
        assert abs(nseg - 35 * npes * cpp) <= npes * cpp
    else:
        assert np.array_equal(cid, (np.arange(numg) % (npes * cpp)) + 1)


# argument order of oracle_fullstep_clumps: the whole step of clm_drv (CanopyFluxes -> SoilTemperature -> SoilFluxes ->
# clm_drv_patch2col -> root-water sink -> SoilWater -> BalanceCheck)
STEP_GROUPS = ("soiltemperature", "soilwater", "canopyfluxes", "plantsink", "balancecheck", "soilfluxes", "patch2col")
CHECKED = ("t_veg", "num_iter", "t_soisno", "h2osoi_liq", "qflx_tran_veg", "t_grnd", "eflx_soil_grnd", "errsoi_col",
           "qflx_evap_soi_col", "qflx_rootsoi", "errh2o")


def _global_case(synthetic_canopy):
    sg, S = synthetic_canopy.make_full_case(48, seed=5)
    synthetic_canopy.balance_state(sg, S, np.random.Generator(np.random.PCG64(6)), 1.0e-11)
    synthetic_canopy.soilfluxes_state(sg, S, np.random.Generator(np.random.PCG64(7)))
    return sg, S


def test_cost_balanced_slabs():
    """Contiguous slabs balanced by exposed-patch count: exact cover, order preserved, imbalance bounded by one
    gridcell, and much better than equal-count slabs on a skewed cost field."""
    rng = np.random.Generator(np.random.PCG64(3))
    numg = 5000
    cost = rng.integers(0, 15, numg).astype(float)
    cost[:1500] = rng.integers(0, 2, 1500)             # a sparsely vegetated band (e.g. high latitudes come first)
    for nparts in (1, 2, 4, 8, 64):
        e = decomp.balanced_slabs(cost, nparts)
        assert e[0] == 0 and e[-1] == numg and np.all(np.diff(e) >= 0) and len(e) == nparts + 1
        parts = np.array([cost[e[k]:e[k + 1]].sum() for k in range(nparts)])
        assert abs(parts.sum() - cost.sum()) < 1e-9
        assert parts.max() - parts.mean() <= cost.max() + 1e-9
    e8 = decomp.balanced_slabs(cost, 8)
    equal = np.linspace(0, numg, 9).astype(np.int64)
    assert decomp.slab_imbalance(cost, e8) < 1.02 < decomp.slab_imbalance(cost, equal)
    assert np.array_equal(decomp.balanced_slabs(np.zeros(10), 3), [0, 3, 6, 10])


def test_driver_slabs_partition_the_subgrid_and_cut_the_filters():
    """driver.make_slabs (what bench.py's e2e clump loop and the multi-GPU layout use): contiguous, exhaustive, whole
    gridcells, filters cut at the bounds in order, balanced by exposed-vegetation patches."""
    from ctsm_b200 import driver, synthetic_canopy
    sg, S = synthetic_canopy.make_full_case(300, seed=9)
    for k in (1, 2, 3, 7):
        slabs = driver.make_slabs(sg, k)
        assert len(slabs) == k
        assert slabs[0][0].begg == sg.bounds.begg and slabs[-1][0].endg == sg.bounds.endg
        for (b0, _), (b1, _) in zip(slabs[:-1], slabs[1:]):
            assert (b1.begg, b1.begc, b1.begp) == (b0.endg + 1, b0.endc + 1, b0.endp + 1)
        for name, f in sg.filters.items():
            assert np.array_equal(np.concatenate([fl[name] for _, fl in slabs]), f), name
        for b, fl in slabs:
            assert np.all(sg.col_gridcell[b.begc - 1:b.endc] >= b.begg) and np.all(sg.col_gridcell[b.begc - 1:b.endc] <= b.endg)
            assert np.array_equal(fl["allc"], np.arange(b.begc, b.endc + 1))
        cost = [len(fl["exposedvegp"]) for _, fl in slabs]
        assert max(cost) - min(cost) <= 30          # one gridcell holds at most 15 patches


def _worker(rank, world, port, q):
    import ctypes as C
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from ctsm_b200 import abi, synthetic_canopy
    from oracle import oracle
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    OL = oracle.lib()
    prm = abi.default_params()
    sg, S = _global_case(synthetic_canopy)                        # every rank builds the same global case
    cells = decomp.rank_gridcells(sg.ngrc, world, 3, rank, nsegspc=2)
    prm.balance_skip_steps = int(OL.oracle_balancecheck_skip_steps(prm.dtime))
    structs = [abi.make_struct(g, S, sg.bounds) for g in STEP_GROUPS]
    touched = np.zeros(sg.ngrc + 1, dtype=bool)
    for cl in cells:                                              # contiguous runs of gridcells = segments
        if len(cl) == 0:
            continue
        runs = np.split(cl, np.nonzero(np.diff(cl) != 1)[0] + 1)
        for run in runs:
            sub = oracle.clump_for_gridcells(sg, int(run[0]), int(run[-1]))
            arr, keep = sub
            assert OL.oracle_fullstep_clumps(C.byref(prm), 1, arr, *[C.byref(x) for x in structs], 1, 127) == 0
            touched[run] = True
    import torch
    mx = torch.tensor([float(np.nanmax(np.where(touched[S["gridcell"]], np.where(S["t_veg"] < 1e30, S["t_veg"], 0), 0)))], dtype=torch.float64)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    got = decomp.reduce_balance_report([float(rank), 2.0], dist)
    pmask = touched[S["gridcell"]]
    cmask = touched[sg.col_gridcell]
    q.put((rank, touched, {k: (S[k][..., pmask] if S[k].shape[-1] == sg.npatch else S[k][..., cmask])
                           for k in CHECKED}, float(mx), got))
    dist.destroy_process_group()


def test_two_rank_gloo_run_matches_single(oracle_lib):
    import ctypes as C
    import torch.multiprocessing as mp
    from ctsm_b200 import abi, synthetic_canopy
    from oracle import oracle
    world, port = 2, 29000 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference run
    prm = abi.default_params()
    prm.balance_skip_steps = int(oracle_lib.oracle_balancecheck_skip_steps(prm.dtime))
    sg, S = _global_case(synthetic_canopy)
    arr, keep = oracle.make_clumps(sg, 1)
    structs = [abi.make_struct(g, S, sg.bounds) for g in STEP_GROUPS]
    assert oracle_lib.oracle_fullstep_clumps(C.byref(prm), 1, arr, *[C.byref(x) for x in structs], 1, 127) == 0
    owned = np.zeros(sg.ngrc + 1, dtype=np.int32)
    for rank, touched, fields, mx, got in res:
        owned += touched
        pmask, cmask = touched[S["gridcell"]], touched[sg.col_gridcell]
        for k, v in fields.items():
            want = S[k][..., pmask] if S[k].shape[-1] == sg.npatch else S[k][..., cmask]
            assert np.array_equal(v, want), (rank, k)             # bit-identical for any partition
        assert got == [float(world - 1), 2.0]
    assert np.all(owned[1:] == 1)
    assert abs(res[0][3] - res[1][3]) == 0.0
