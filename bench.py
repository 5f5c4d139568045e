#!/usr/bin/env python
"""bench.py — column-timesteps/s of the CTSM biogeophysics hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--size f02] [--impl reference]

One "step" = one pass of the hot path in clm_drv call order (clm_driver.F90:766,900,921,950,1422),
CanopyFluxes (+PHS) -> SoilTemperature -> SoilFluxes -> clm_drv_patch2col -> root-water sink -> SoilWater -> BalanceCheck,
over the rank's
synthetic grid (BASELINE.json config 4 on one GPU; `--size f09 --routines soiltemperature,soilwater` is
config 2, `--size f09 --routines canopyfluxes` config 3).  `value` is whole-job throughput with all state resident in HBM; `e2e` is
the same step driven through the C ABI with pinned HOST buffers, host<->device copies inside the
timed region.  `roofline` describes the dominant routine's kernels, `cpu_baseline` the CPU oracle
(C restatement of the reference, OpenMP over clumps) on a bounded sample of the same workload.
Under torchrun ONE grid of the named size is dealt to the ranks in contiguous gridcell slabs (clump decomposition,
decompInitMod.F90:96-161; BASELINE.json config 4): strong scaling.  The path has no exchange step; the one collective in
the timed region is the NCCL MAX-reduction of BalanceCheck's residual maxima on a side stream (`--scaling weak` gives
every rank its own full-size grid instead).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "column_timesteps_per_sec"
UNIT = "column-steps/s"
PRE_ROUTINES = ("preflux", "surfacehumidity", "baregroundfluxes")     # clm_driver.F90:680, 702, 711
ALL_ROUTINES = ("canopyfluxes", "soiltemperature", "soilfluxes", "patch2col", "plantsink", "soilwater", "balancecheck")   # clm_drv order
# the rest of HydrologyNoDrainage around the sink and SoilWater (HydrologyNoDrainageMod.F90:279-757): `--routines hydro,...`
HYDRO_ROUTINES = ("snowwater", "infiltration", "watertable", "snowcapping", "snowlayers", "hydrodiag")
FULL_ORDER = PRE_ROUTINES + ("canopyfluxes", "soiltemperature", "soilfluxes", "patch2col", "snowwater", "infiltration", "plantsink", "soilwater",
                             "watertable", "snowcapping", "snowlayers", "hydrodiag", "balancecheck")
KERNEL_OF = {"plantsink": "plantsink_kernel", "soilfluxes": "soilfluxes_patch_kernel + soilfluxes_p2c_kernel", "patch2col": "patch2col_kernel<false/true>", "balancecheck": "balance_col/grc/patch/loc kernels",
             "canopyfluxes": "CanopyFluxes kernel chain of one call (init, then per ITERATION pass close/fric/leaf, "
                             "phs_ci x4, phs_newton x4, phs_end; final) - largest member: phs_newton_kernel",
             "soiltemperature": "soiltemp_kernel", "soilwater": "soilwater_kernel"}
KERNEL_OF.update({"preflux": "preflux_patch_a / preflux_col / preflux_patch_b kernels", "surfacehumidity": "surface_humidity_kernel",
                  "baregroundfluxes": "bareground_kernel + bareground_colcopy_kernel"})
KERNEL_OF.update({"snowwater": "aerosol_dep_kernel + snow_water_kernel", "infiltration": "floodc_kernel + infiltration_kernel",
                  "watertable": "water_table_kernel", "snowcapping": "snow_capping_init_kernel + snow_capping_kernel",
                  "snowlayers": "snow_layers_kernel", "hydrodiag": "hydrodiag_nolake / snowdp / snow / soil kernels"})
NAME_OF = {"snowwater": "BuildSnowFilter+SnowWater", "infiltration": "SetSoilWaterFractions..TotalSurfaceRunoff",
           "watertable": "PerchedWaterTable+ThetaBasedWaterTable+RenewCondensation", "snowcapping": "SnowCapping",
           "snowlayers": "SnowCompaction+CombineSnowLayers+DivideSnowLayers+ZeroEmptySnowLayers",
           "hydrodiag": "BuildSnowFilter+HydrologyNoDrainage diagnostics",
           "preflux": "BiogeophysPreFluxCalcs", "surfacehumidity": "CalculateSurfaceHumidity", "baregroundfluxes": "BareGroundFluxes",
           "canopyfluxes": "CanopyFluxes", "soiltemperature": "SoilTemperature", "soilwater": "SoilWater",
           "plantsink": "VertTranSink_HydStress", "balancecheck": "BalanceCheck", "soilfluxes": "SoilFluxes", "patch2col": "clm_drv_patch2col"}


def make_workload(size, seed, hydro=False):
    """Synthetic subgrid + state of every routine of the step (SURVEY.md 8d generators)."""
    from ctsm_b200 import synthetic_canopy
    sg, S = synthetic_canopy.make_full_case(size, seed=seed)
    synthetic_canopy.balance_state(sg, S, np.random.Generator(np.random.PCG64(seed + 1)), 1.0e-11)
    synthetic_canopy.soilfluxes_state(sg, S, np.random.Generator(np.random.PCG64(seed + 2)))
    synthetic_canopy.preflux_state(sg, S, np.random.Generator(np.random.PCG64(seed + 3)))     # own generator: adds fields only
    if hydro:                                               # inputs of the HydrologyNoDrainage routines (own generators as well)
        synthetic_canopy.hydrology_state(sg, S, np.random.Generator(np.random.PCG64(seed + 4)))
        synthetic_canopy.snow_state(sg, S, np.random.Generator(np.random.PCG64(seed + 5)))
        synthetic_canopy.watertable_state(sg, S, np.random.Generator(np.random.PCG64(seed + 6)), saturate=False)
        S["topo"] = np.random.Generator(np.random.PCG64(seed + 7)).uniform(0.0, 3000.0, sg.ncol)
        for k in ("qflx_snwcp_ice", "qflx_snwcp_liq", "qflx_snwcp_discarded_ice", "qflx_snwcp_discarded_liq"):
            S[k] = np.full(sg.ncol, 1.0e36)
    return sg, S


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks line of B200_PROFILING.md, sampled while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.path = device, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                t = [x.strip() for x in line.split(",")]
                if len(t) < 9:
                    continue
                try:
                    sm.append(float(t[1])); mx.append(float(t[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), t[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            hi = [s for s in sm if s >= 0.5 * max(sm)]
            out["sm_mhz"] = float(np.median(hi))
            out["sm_max_mhz"] = float(max(mx))
            out["reasons"] = sorted(reasons)
        return out


def read_profile(size):
    """The committed ncu launch-list summary of this command (tools/launch_summary.py -> profiles/r02_traffic.json): DRAM
    bytes and FP64-pipe instructions per call of each routine.  None when the workload differs from the profiled one
    (the counters are per launch of THAT grid)."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            d = json.load(open(os.path.join(ROOT, "profiles", name)))
            if d.get("size") == str(size):
                d["file"] = "profiles/" + name
                return d
        except Exception:
            continue
    return None


FP64_LANES_PER_SM, N_SM = 64, 148          # B200: 148 SMs x 4 SMSPs x 16 FP64 lanes; a warp instruction holds an SMSP's pipe 2 cycles


def which_mask(routines):
    return sum({"soiltemperature": 1, "soilwater": 2, "canopyfluxes": 4, "plantsink": 8, "balancecheck": 16, "soilfluxes": 32, "patch2col": 64}.get(g, 0) for g in routines)   # (the CPU arm times the seven-routine step)


def cpu_reference_run(size_label, routines, steps, warmup, sample_gridcells, seed):
    """The reference-arm / cpu_baseline measurement: oracle routines driven clump-parallel
    (OpenMP over clumps like clm_driver.F90:525) on a bounded sample of the workload."""
    from ctsm_b200 import abi, synthetic_canopy
    from oracle import oracle
    OL = oracle.lib()
    # all the host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which is not the reference's setup)
    OL.oracle_set_num_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    nthreads = int(OL.oracle_num_threads())
    sg, S = make_workload(sample_gridcells, seed)
    prm = oracle.default_params()          # from liboracle.so: this arm never loads the CUDA library
    prm.balance_skip_steps = int(OL.oracle_balancecheck_skip_steps(prm.dtime))
    clumps, keep = oracle.make_clumps(sg, nthreads * 4)
    inout = {fs.name for g in ALL_ROUTINES for fs in abi.FIELDS[g] if fs.intent != "IN"}
    pristine = {k: S[k].copy() for k in inout}
    ft = abi.make_struct("soiltemperature", S, sg.bounds)
    fw = abi.make_struct("soilwater", S, sg.bounds)
    fc = abi.make_struct("canopyfluxes", S, sg.bounds)
    fs_ = abi.make_struct("plantsink", S, sg.bounds)
    fb = abi.make_struct("balancecheck", S, sg.bounds)
    fx = abi.make_struct("soilfluxes", S, sg.bounds)
    f2c = abi.make_struct("patch2col", S, sg.bounds)
    times = []
    for it in range(warmup + steps):
        for k, v in pristine.items():
            S[k][...] = v
        t0 = time.perf_counter()
        rc = OL.oracle_fullstep_clumps(C.byref(prm), len(clumps), clumps, C.byref(ft), C.byref(fw), C.byref(fc),
                                       C.byref(fs_), C.byref(fb), C.byref(fx), C.byref(f2c), 1, which_mask(routines))
        t1 = time.perf_counter()
        assert rc == 0, "oracle step failed rc=%d" % rc
        if it >= warmup:
            times.append(t1 - t0)
    tot = float(np.sum(times))
    return {"value": sg.ncol * steps / tot, "ms_per_step": 1e3 * tot / steps, "cores": nthreads, "columns": sg.ncol,
            "sample": "%d-gridcell (%d columns, %d patches, %d exposed-veg patches) sample of the %s workload, %d steps of %s, "
            "C restatement of the reference (gcc -O2 -ffp-contract=off -fopenmp, one clump per task, %d threads); "
            "not gfortran: no Fortran compiler in the image" % (
                sg.ngrc, sg.ncol, sg.npatch, len(sg.filters["exposedvegp"]), size_label, steps,
                "+".join(NAME_OF[g] for g in routines), nthreads)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--under-profiler", action="store_true", help="run under ncu: allows --warmup < 3; the printed line is not a measurement")
    ap.add_argument("--size", default="f02", help="tiny|f19|f09|f02 or a gridcell count (per GPU)")
    ap.add_argument("--routines", default=",".join(ALL_ROUTINES))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=20000, help="gridcells in the CPU baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-slabs", type=int, default=6,
                    help="gridcell slabs the e2e step is issued over (clump loop).  Measured on B200: 6 slabs beat 3 at one GPU "
                         "(f02) and at two (1.82 M against 1.45 M column-steps/s): the overlap of uploads, kernels and downloads "
                         "is worth more than the canopy call's latency floor that every slab pays")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--members", type=int, default=1,
                    help="perturbed-parameter ensemble (BASELINE config 5): this many parameter sets x the --size grid, batched "
                         "as independent columns; dealt to the ranks under torchrun")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="under torchrun: strong = ONE grid of --size dealt to the ranks (default), weak = one grid of --size per rank")
    a = ap.parse_args()
    # timing rule: at least 3 warm-up steps; --under-profiler (ncu launch lists, never a bench value) lifts it
    a.warmup = max(a.warmup, 3) if (a.impl == "b200" and not a.under_profiler) else a.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    size = a.size if not a.size.isdigit() else int(a.size)
    # --routines names a subset of the step; "pre" (or the three names) adds the routines clm_drv runs before CanopyFluxes
    want = a.routines.replace("pre,", ",".join(PRE_ROUTINES) + ",").replace("hydro,", ",".join(HYDRO_ROUTINES) + ",").split(",")
    routines = tuple(g for g in FULL_ORDER if g in want)
    strong = a.scaling == "strong" or world == 1
    from ctsm_b200 import synthetic
    members_local = a.members
    if a.members > 1:
        if a.members % world != 0:
            raise SystemExit("bench.py: --members must be a multiple of the number of ranks")
        per_member_g = synthetic.GRID_SIZES[size] if isinstance(size, str) else int(size)
        members_local = a.members // world if strong else a.members
        size = per_member_g * a.members                     # the ensemble as ONE grid of independent columns
    if strong and world > 1:
        total_g = synthetic.GRID_SIZES[size] if isinstance(size, str) else int(size)
        g0, g1 = rank * total_g // world, (rank + 1) * total_g // world      # contiguous slab of the one grid (clump range)
        local_size = g1 - g0
    else:
        local_size = size
    wl_name = "%s one 1800 s step, ONE %s synthetic grid (15 patches per soil column)%s" % (
        "->".join(NAME_OF[g] for g in routines),
        ("%s-sized" % a.size) if a.members == 1 else ("%d-member perturbed-parameter ensemble x %s-sized" % (a.members, a.size)),
        "" if world == 1 else (" dealt to %d GPUs in contiguous gridcell slabs" % world if strong else " PER GPU (weak scaling)"))
    config = {"workload": wl_name, "grid": str(a.size), "routines": [NAME_OF[g] for g in routines],
              "state": "restored from a pristine device snapshot before every step (untimed D2D copies)",
              "l2": "per-step working set (GBs at f02) exceeds the 126 MB L2; the untimed state restore between steps "
                    "streams >L2 bytes through the cache",
              "parallelism": "gridcells/clumps sharded by rank (contiguous slabs), no data-path collective; NCCL all-reduce(MAX) of the "
                             "7 BalanceCheck maxima per step on a side stream"}

    if a.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_run(a.size, routines, a.steps, a.warmup, a.cpu_sample, 20260101)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": a.gpus,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    from ctsm_b200 import abi, synthetic_canopy, driver
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; ctsm_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    prm = abi.default_params(device=local_rank)
    sg, S = make_workload(local_size, 20260101 + 1000 * rank, hydro=any(g in HYDRO_ROUTINES for g in routines))
    if a.members > 1:
        # config 5: members = contiguous gridcell ranges; PFT tables per member (medlynslope, kmax, psi50, ck, krmax perturbed)
        # and per-member scalars e_ice, csoilc, cv, a_coef, z_dl
        rngm = np.random.Generator(np.random.PCG64(20260105 + rank))
        synthetic_canopy.make_ensemble(sg, S, members_local, rngm, spread=0.2)
        prm.npft_table = members_local * (abi.MXPFT + 1)
    ctx = driver.Context(prm)
    if a.members > 1:
        member_g = np.minimum((np.arange(sg.ngrc) * members_local) // sg.ngrc, members_local - 1)
        scal = {k: getattr(prm, k) * rngm.uniform(0.8, 1.2, members_local) for k in ("e_ice", "csoilc", "cv", "a_coef", "z_dl")}
        ctx.set_member_params(members_local, member_g[sg.col_gridcell - 1], sg.bounds.begc, sg.bounds.endc, **scal)
    names = sorted({fs.name for g in routines for fs in abi.FIELDS[g]})
    D = {k: torch.from_numpy(S[k]).cuda() for k in names}
    restore = sorted({fs.name for g in routines for fs in abi.FIELDS[g] if fs.intent != "IN"})
    pristine = {k: D[k].clone() for k in restore}
    hp = driver.HotPath(ctx, sg, D, abi.MEM_DEVICE, routines)
    if dist is not None and "balancecheck" in routines:
        hp.enable_global_balance(dist)        # BalanceCheckMod's global water / energy figures: NCCL MAX on a side stream
    stream = torch.cuda.ExternalStream(ctx.stream_ptr, device=torch.device("cuda", local_rank))

    def reset_state():
        with torch.cuda.stream(stream):
            for k, v in pristine.items():
                D[k].copy_(v)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    ncol = sg.ncol
    for _ in range(a.warmup):
        reset_state(); hp.step()
    ctx.sync()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launches
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(len(routines) + 1)] for _ in range(a.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for it in range(a.steps):
        reset_state()
        ev[it][0].record(stream)
        for i, g in enumerate(routines):
            hp.call(g)
            ev[it][i + 1].record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    st = ctx.sync()
    launches = ctx.launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [ev[it][0].elapsed_time(ev[it][-1]) for it in range(a.steps)]
    rt_ms = {g: float(np.mean([ev[it][i].elapsed_time(ev[it][i + 1]) for it in range(a.steps)])) for i, g in enumerate(routines)}
    total_s = float(np.sum(step_ms)) / 1e3
    t = torch.tensor([total_s], dtype=torch.float64, device="cuda")
    cols = torch.tensor([float(ncol)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cols, op=dist.ReduceOp.SUM)
    total_s, total_cols = float(t.item()), float(cols.item())
    value = total_cols * a.steps / total_s

    # Rooflines (SURVEY.md 8d asks for both and for which one binds), per routine and for the dominant one.
    #   hbm : algorithmic bytes of the field table / event-timed duration, against the measured copy bandwidth
    #   fp64: FP64-pipe warp instructions of the call (committed ncu pass of this command, profiles/r02_traffic.json) x 32
    #         thread slots / event-timed duration, against 148 SMs x 64 FP64 lanes x the SM clock sampled DURING this run.
    #         The code is compiled -fmad=false (the reference's -ffp-contract=off), so one slot is one flop.
    peak, peak_src = read_peaks()
    prof = read_profile(a.size) if (world == 1 and a.members == 1 and tuple(routines) == ALL_ROUTINES) else None
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    fp64_peak = N_SM * FP64_LANES_PER_SM * sm_mhz * 1e6 / 1e9          # G thread-slots / s
    per_routine = {}
    for g in routines:
        ab = driver.algorithmic_bytes(sg, S, g)
        e = {"ms": rt_ms[g], "algorithmic_bytes": ab["bytes"], "GBps": ab["bytes"] / (rt_ms[g] * 1e-3) / 1e9,
             "frac_of_hbm_peak": ab["bytes"] / (rt_ms[g] * 1e-3) / 1e9 / peak, "columns": ab["columns"], "patches": ab["patches"]}
        if prof is not None and prof.get("fp64_warp_inst_per_call", {}).get(NAME_OF[g]):
            w = prof["fp64_warp_inst_per_call"][NAME_OF[g]]
            t = prof.get("fp64_thread_inst_per_call", {}).get(NAME_OF[g], 0.0)
            e["fp64_Gslots_per_s"] = 32.0 * w / (rt_ms[g] * 1e-3) / 1e9
            e["frac_of_fp64_peak"] = e["fp64_Gslots_per_s"] / fp64_peak
            e["fp64_lanes_per_warp_inst"] = t / w if w else None
            e["dram_traffic_bytes"] = prof["dram_bytes_per_call"].get(NAME_OF[g])
        e["bound"] = "fp64" if e.get("frac_of_fp64_peak", 0.0) > e["frac_of_hbm_peak"] else "hbm"
        per_routine[NAME_OF[g]] = e
    dom = max(rt_ms, key=rt_ms.get)
    ab = driver.algorithmic_bytes(sg, S, dom)
    d = per_routine[NAME_OF[dom]]
    step_bytes = sum(v["algorithmic_bytes"] for v in per_routine.values())
    hbm = {"achieved": d["GBps"], "peak": peak, "unit": "GB/s", "frac": d["frac_of_hbm_peak"], "peak_source": peak_src,
           "algorithmic_bytes_per_launch": ab["bytes"]}
    fp64 = None
    if "frac_of_fp64_peak" in d:
        fp64 = {"achieved": d["fp64_Gslots_per_s"], "peak": fp64_peak, "unit": "G FP64 thread-slots/s", "frac": d["frac_of_fp64_peak"],
                "peak_source": "148 SMs x 64 FP64 lanes x %.0f MHz (SM clock sampled under load in this run)" % sm_mhz,
                "warp_instructions_per_launch": prof["fp64_warp_inst_per_call"][NAME_OF[dom]],
                "lanes_per_warp_instruction": d["fp64_lanes_per_warp_inst"], "source": prof["file"]}
    binds = fp64 if d["bound"] == "fp64" else hbm
    roofline = {"bound": d["bound"], "kernel": KERNEL_OF[dom], "achieved": binds["achieved"], "peak": binds["peak"],
                "peak_source": binds["peak_source"], "unit": binds["unit"], "frac": binds["frac"],
                "traffic": prof["dram_bytes_per_call"].get(NAME_OF[dom]) if prof else None,
                "algorithmic_bytes_per_launch": ab["bytes"], "ms_per_launch": rt_ms[dom], "hbm": hbm, "fp64": fp64,
                "note": "a 'launch' is one call of the dominant routine (its chain of kernels, timed with CUDA events on the "
                        "library's stream).  `bound` is the roofline the routine sits closer to; for CanopyFluxes + PHS that is "
                        "the FP64 pipe (hundreds of exp/log/pow per patch-pass), but the ITERATION loop's 41 data-dependent "
                        "rounds make it latency-bound below either roof (DESIGN.md section 4.1: the worst patch's Newton "
                        "chain sets a floor per round).  fp64 is null when the workload is not the profiled one.",
                "whole_step": {"algorithmic_bytes": step_bytes, "GBps": step_bytes / (float(np.mean(step_ms)) * 1e-3) / 1e9,
                               "frac": step_bytes / (float(np.mean(step_ms)) * 1e-3) / 1e9 / peak},
                "routines": per_routine}

    # e2e: same step through the C ABI with pinned HOST buffers (H2D + kernels + D2H per step), issued clump after clump
    # inside a resident window (include/ctsm_b200.h): every IN/INOUT field is uploaded once per step, every OUT/INOUT
    # field downloaded once per routine that writes it; bytes are counted by the library from the copies it issues
    e2e = None
    if not a.no_e2e:
        H = {k: torch.from_numpy(S[k]).pin_memory() for k in names}
        Hn = {k: v.numpy() for k, v in H.items()}
        Hp = {k: S[k].copy() for k in restore}
        a.e2e_slabs = max(1, a.e2e_slabs)
        hph = driver.HotPath(ctx, sg, Hn, abi.MEM_HOST, routines, nslab=a.e2e_slabs, window=True)
        e2e_steps = max(2, min(a.steps, 3))
        ts = []
        h2d = d2h = 0
        for it in range(1 + e2e_steps):                  # the first step also creates the device mirrors (one-time)
            for k, v in Hp.items():
                Hn[k][...] = v
            barrier()
            t0 = time.perf_counter()
            hph.step()
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            if it >= 1:
                ts.append(t1 - t0)
                h2d, d2h = hph.window_bytes()
        te = torch.tensor([float(np.sum(ts))], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": total_cols * e2e_steps / float(te.item()), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "ms_per_step": 1e3 * float(te.item()) / e2e_steps,
               "slabs": a.e2e_slabs,
               "mode": "CTSM_MEM_HOST inside ctsm_b200_host_window_begin/_end, the step issued over %d contiguous gridcell slabs "
                       "(clump loop): per step every IN/INOUT field crosses PCIe once, every OUT/INOUT field once per routine that "
                       "writes it; uploads, kernels and downloads of successive slabs overlap on three streams; wall clock from "
                       "window_begin to window_end" % a.e2e_slabs}
        del hph, H, Hn
        ctx.L.ctsm_b200_host_invalidate(ctx.h, None)

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        r = cpu_reference_run(a.size, routines, 3, 1, a.cpu_sample, 20260101)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}

    if rank == 0:
        nit = S["num_iter"] if False else None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": 1e3 * total_s / a.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak",
                "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": dict(config, columns_per_gpu=ncol, patches_per_gpu=sg.npatch,
                               exposedveg_patches_per_gpu=int(len(sg.filters["exposedvegp"])), members=a.members,
                               members_per_gpu=members_local),
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
                "warnings_in_timed_region": int(st.n_warnings), "wall_s_timed_region": t_wall,
                "balance_global_max": hp.global_balance(), "total_columns": int(total_cols)}
        if a.under_profiler:
            line["under_profiler"] = True          # not a measurement
        print(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
