"""The N > 1 layout through the CUDA library (VERDICT r01 "multi-GPU invariant proven on the wrong implementation"):
two processes (torch.distributed, world_size 2; both on cuda:0, which is all a one-GPU test box has) each run the
seven-routine step on THEIR contiguous gridcell slab of one global grid - call bounds = the slab, arrays allocated with
the global bounds - once device-resident and once with host arrays, and the union must be bit-identical to the
single-process CUDA run over the whole grid (the reference's ERP/PEM invariant).  The ranks
also MAX-reduce BalanceCheck's maxima from the library's device slot (driver.HotPath.enable_global_balance) and must
arrive at the single-process figures."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECKED = ("t_veg", "num_iter", "t_soisno", "h2osoi_liq", "h2osoi_ice", "qflx_tran_veg", "t_grnd", "eflx_soil_grnd", "errsoi_col",
           "qflx_evap_soi_col", "qflx_rootsoi", "errh2o", "errseb", "smp_l", "num_substeps", "imelt", "vegwp", "btran")


def _global_case():
    from ctsm_b200 import synthetic_canopy
    sg, S = synthetic_canopy.make_full_case(400, seed=5)
    synthetic_canopy.balance_state(sg, S, np.random.Generator(np.random.PCG64(6)), 1.0e-11)
    synthetic_canopy.soilfluxes_state(sg, S, np.random.Generator(np.random.PCG64(7)))
    return sg, S


def _worker(rank, world, port, mode, q):
    try:
        _worker_body(rank, world, port, mode, q)
    except BaseException:                      # a dead worker must not leave the parent waiting on the queue
        import traceback
        q.put(("error", rank, traceback.format_exc()))


def _worker_body(rank, world, port, mode, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from ctsm_b200 import abi, driver
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    sg, S = _global_case()                                   # every rank holds the global arrays; it touches its slab only
    ctx = driver.Context(abi.default_params())
    names = sorted({fs.name for g in driver.ROUTINES for fs in abi.FIELDS[g]})
    if mode == "device":
        A = {k: torch.from_numpy(np.ascontiguousarray(S[k])).cuda() for k in names}
        hp = driver.HotPath(ctx, sg, A, abi.MEM_DEVICE, nslab=world)
        hp.enable_global_balance(dist)
    else:
        A = S
        hp = driver.HotPath(ctx, sg, A, abi.MEM_HOST, nslab=world, window=(mode == "window"))
    b = hp.slabs[rank][0]
    if hp.window:
        assert ctx.L.ctsm_b200_host_window_begin(ctx.h) == 0
    hp._select(rank)
    for g in hp.routines:
        hp.call(g)
    if hp.window:
        import ctypes as C
        st = abi.Status()
        assert ctx.L.ctsm_b200_host_window_end(ctx.h, C.byref(st)) == 0
    else:
        ctx.sync()
    gmax = hp.global_balance()
    own = list(hp.balance_report.max_abs)
    out = {}
    for k in CHECKED:
        a = A[k].cpu().numpy() if mode == "device" else A[k]
        lo, hi = (b.begp, b.endp) if a.shape[-1] == sg.npatch else (b.begc, b.endc)
        out[k] = a[..., lo - 1:hi].copy()
        if mode == "device":                  # nothing outside the slab was touched (host modes work on S itself)
            rest = np.ones(a.shape[-1], dtype=bool); rest[lo - 1:hi] = False
            assert np.array_equal(a[..., rest], S[k][..., rest], equal_nan=True), k
    q.put((rank, (b.begc, b.endc, b.begp, b.endp), out, gmax, own))
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["device", "host", "window"])
def test_two_rank_cuda_run_matches_single_rank_cuda_run(mode):
    import torch
    import torch.multiprocessing as mp
    from ctsm_b200 import abi, driver
    world, port = 2, 29500 + (os.getpid() % 1000) + {"device": 0, "host": 1, "window": 2}[mode]
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_worker, args=(r, world, port, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for r in res:
        assert r[0] != "error", r[2]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process CUDA run over the whole grid
    sg, S = _global_case()
    ctx = driver.Context(abi.default_params())
    try:
        names = sorted({fs.name for g in driver.ROUTINES for fs in abi.FIELDS[g]})
        D = {k: torch.from_numpy(np.ascontiguousarray(S[k])).cuda() for k in names}
        hp = driver.HotPath(ctx, sg, D, abi.MEM_DEVICE)
        hp.step()
        ctx.sync()
        single_max = list(hp.balance_report.max_abs)
        whole = {k: D[k].cpu().numpy() for k in CHECKED}
    finally:
        ctx.close()
    covered_c = np.zeros(sg.ncol, dtype=np.int32)
    for rank, (bc, ec, bp, ep), out, gmax, own in res:
        covered_c[bc - 1:ec] += 1
        for k, v in out.items():
            lo, hi = (bp, ep) if whole[k].shape[-1] == sg.npatch else (bc, ec)
            assert np.array_equal(v, whole[k][..., lo - 1:hi], equal_nan=True), (mode, rank, k)
        if mode == "device":
            assert gmax is not None and np.array_equal(np.asarray(gmax), np.asarray(single_max)), (gmax, single_max)
    assert np.all(covered_c == 1)
    # the global maximum is the maximum of the ranks' own maxima
    own_max = np.max(np.asarray([r[4] for r in res]), axis=0)
    assert np.array_equal(own_max, np.asarray(single_max))
