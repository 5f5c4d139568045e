"""Host-side mirror of the hot-path call sites in clm_drv.

Reference call order inside the clump loop (src/main/clm_driver.F90):
    CanopyFluxes :766  ->  SoilTemperature :900  ->  HydrologyNoDrainage/SoilWater :950
(+ BalanceCheck :1422 in the second clump loop).  This module owns a library
context and the device-resident state and issues those calls through the C ABI
(include/ctsm_b200.h).  It contains no arithmetic.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, Optional

import numpy as np

from . import abi

# clm_drv call order: CanopyFluxes clm_driver.F90:766, SoilTemperature :900, SoilFluxes :921, clm_drv_patch2col :936,
# HydrologyNoDrainage :950
# (root-water sink HydrologyNoDrainageMod.F90:339, SoilWater :346), BalanceCheck :1422
ROUTINES = ("canopyfluxes", "soiltemperature", "soilfluxes", "patch2col", "plantsink", "soilwater", "balancecheck")
# the routines clm_drv runs just before CanopyFluxes (clm_driver.F90:680, :702, :711); PRE_ROUTINES + ROUTINES
# is the step from BiogeophysPreFluxCalcs to BalanceCheck
PRE_ROUTINES = ("preflux", "surfacehumidity", "baregroundfluxes")
# HydrologyNoDrainage's routines in front of the root-water sink (HydrologyNoDrainageMod.F90:297-337): with
# them SoilWater's icefrac / eff_porosity / qflx_infl are produced on the device instead of being inputs
HYDRO_ROUTINES = ("snowwater", "infiltration", "watertable", "snowcapping", "snowlayers", "hydrodiag")
# HydrologyNoDrainage as far as it is built (HydrologyNoDrainageMod.F90:279-757): BuildSnowFilter, SnowWater, the infiltration chain,
# the root-water sink, SoilWater, PerchedWaterTable / ThetaBasedWaterTable / RenewCondensation, SnowCompaction / CombineSnowLayers /
# SnowCapping, SnowCompaction / CombineSnowLayers / DivideSnowLayers / ZeroEmptySnowLayers, BuildSnowFilter, the closing diagnostics
ROUTINES_HYDRO = ("canopyfluxes", "soiltemperature", "soilfluxes", "patch2col", "snowwater", "infiltration", "plantsink", "soilwater",
                  "watertable", "snowcapping", "snowlayers", "hydrodiag", "balancecheck")
FILTER_OF = {"preflux": ("nolakec", "nolakep"), "surfacehumidity": ("nolakec",), "baregroundfluxes": ("noexposedvegp",),
             "canopyfluxes": ("exposedvegp",), "soiltemperature": ("nolakep", "nolakec"), "soilfluxes": ("nolakep", "nolakec"),
             "patch2col": ("allc", "nolakec"), "infiltration": ("nolakec", "hydrologyc"), "plantsink": ("hydrologyc",),
             "snowwater": ("nolakec",), "snowlayers": ("nolakec",), "watertable": ("hydrologyc",),
             "hydrodiag": ("nolakec", "hydrologyc"), "snowcapping": ("nolakec",), "soilwater": ("hydrologyc",), "balancecheck": ("allc",)}


class CtsmError(RuntimeError):
    """Raised where the reference would call endrun()."""

    def __init__(self, st: abi.Status, rc: int):
        self.code, self.subgrid_level, self.subgrid_index, self.info = rc, st.subgrid_level, st.subgrid_index, st.info
        super().__init__("ctsm_b200 rc=%d level=%d index=%d info=%d: %s" % (
            rc, st.subgrid_level, st.subgrid_index, st.info, st.msg.decode()))


class Context:
    def __init__(self, prm: Optional[abi.Params] = None):
        self.L = abi.lib()
        self.prm = prm if prm is not None else abi.default_params()
        self.h = C.c_void_p()
        self._pending = []      # objects the library writes at the next synchronisation (deferred BalanceCheck reports)
        rc = self.L.ctsm_b200_init(C.byref(self.prm), C.byref(self.h))
        if rc != 0:
            why = {1: "no usable CUDA device; there is no CPU fallback", 2: "bad argument / unsupported configuration",
                   3: "device memory allocation failed", 4: "CUDA runtime error"}.get(rc, "error")
            raise RuntimeError("ctsm_b200_init failed (rc=%d): %s %s" % (rc, why, self.L.ctsm_b200_last_cuda_error().decode()))

    def set_tuning(self, tail_max: int = -1, nt_budget: int = -1, tail_lanes: int = -1, nt_split: int = -1):
        """Scheduling knobs of CanopyFluxes' ITERATION loop (include/ctsm_b200.h); results do not depend on them."""
        self.L.ctsm_b200_set_tuning(self.h, tail_max, nt_budget, tail_lanes, nt_split)

    def set_member_params(self, nmember, col_member, begc, endc, **tables):
        """Per-member values of e_ice / csoilc / cv / a_coef / z_dl (ctsm_b200_set_member_params)."""
        arr = {k: (np.ascontiguousarray(v, dtype=np.float64) if v is not None else None)
               for k, v in ((k, tables.get(k)) for k in ("e_ice", "csoilc", "cv", "a_coef", "z_dl"))}
        cm = np.ascontiguousarray(col_member, dtype=np.int32) if col_member is not None else None
        rc = self.L.ctsm_b200_set_member_params(self.h, int(nmember), *[abi.f64p(arr[k]) for k in ("e_ice", "csoilc", "cv", "a_coef", "z_dl")],
                                                abi.i32p(cm), int(begc), int(endc))
        if rc != 0:
            raise RuntimeError("ctsm_b200_set_member_params rc=%d" % rc)

    def close(self):
        if self.h:
            self.L.ctsm_b200_finalize(self.h)
            self.h = C.c_void_p()

    def sync(self) -> abi.Status:
        st = abi.Status()
        rc = self.L.ctsm_b200_sync(self.h, C.byref(st))
        self._pending.clear()
        if rc != 0:
            raise CtsmError(st, rc)
        return st

    @property
    def stream_ptr(self) -> int:
        return int(self.L.ctsm_b200_stream(self.h) or 0)

    @property
    def launches(self) -> int:
        return int(self.L.ctsm_b200_launch_count(self.h))


def make_slabs(sg, nslab: int, cost: str = "exposedvegp"):
    """Contiguous gridcell slabs (clump bounds + the rank-level filters cut at them), balanced by the number of
    exposed-vegetation patches (decomp.balanced_slabs; decompInitMod.F90:137-161 gives clumps whole gridcells).
    Returns [(Bounds, {filter name: int32 array})]."""
    from . import decomp
    ng = sg.ngrc
    nslab = max(1, min(int(nslab), ng))
    cost_g = np.zeros(ng, dtype=np.float64)
    f = sg.filters.get(cost)
    if f is not None and len(f):
        gcell = sg.col_gridcell[sg.patch_column[f - 1] - 1]
        np.add.at(cost_g, gcell - sg.bounds.begg, 1.0)
    cost_g += 1.0e-3                                              # so that empty stretches are still dealt out
    edges = decomp.balanced_slabs(cost_g, nslab)
    out = []
    for k in range(nslab):
        g0, g1 = int(edges[k]) + sg.bounds.begg, int(edges[k + 1]) + sg.bounds.begg - 1
        if g1 < g0:
            continue
        cols = np.nonzero((sg.col_gridcell >= g0) & (sg.col_gridcell <= g1))[0]
        b = sg.bounds.copy()
        b.begg, b.endg = g0, g1
        b.begc, b.endc = int(cols[0]) + sg.bounds.begc, int(cols[-1]) + sg.bounds.begc
        b.begl, b.endl = b.begc, b.endc                           # one landunit per column in the synthetic subgrid
        b.begp, b.endp = int(sg.col_patchi[cols[0]]), int(sg.col_patchf[cols[-1]])
        b.level, b.clump_index = 2, k + 1
        fl = {}
        for name, arr in sg.filters.items():
            lo, hi = (b.begp, b.endp) if name.endswith("p") else (b.begc, b.endc)
            fl[name] = np.ascontiguousarray(arr[(arr >= lo) & (arr <= hi)])
        fl["allc"] = np.arange(b.begc, b.endc + 1, dtype=np.int32)
        out.append((b, fl))
    return out


class HotPath:
    """One rank's share of the grid: subgrid topology, filters and state arrays
    (numpy = host-owned as in the Fortran model, torch = device-resident).

    nslab > 1 issues the step clump after clump (clm_drv's clump loop, clm_driver.F90:525) over nslab contiguous
    slabs; with host-owned arrays and window=True the step runs inside a resident window
    (ctsm_b200_host_window_begin/_end), where uploads, kernels and downloads of successive slabs overlap."""

    def __init__(self, ctx: Context, sg, arrays: Dict[str, object], mem: int, routines: Iterable[str] = ROUTINES,
                 nslab: int = 1, window: bool = False):
        self.ctx, self.sg, self.arrays, self.mem = ctx, sg, arrays, mem
        self.routines = tuple(routines)
        self.window = bool(window) and mem != abi.MEM_DEVICE
        # the sink SoilWater consumes follows use_hydrstress (Compute_EffecRootFrac_And_VertTranSink, SoilWaterPlantSinkMod.F90:77-131)
        self.group_of = {g: ("plantsinkdefault" if g == "plantsink" and not ctx.prm.use_hydrstress else g) for g in self.routines}
        self.structs = {g: abi.make_struct(self.group_of[g], arrays, sg.bounds) for g in self.routines}
        # DAnstep: 1 = inside BalanceCheck's skip steps (BalanceCheckMod.F90:91,754): all residuals, maxima and warnings
        # are computed, the abort is suppressed - the synthetic water/energy terms are not a closed budget after the step
        self.danstep = 1
        if "balancecheck" in self.routines:
            self.ctx.L.ctsm_b200_balancecheck_init(self.ctx.h)
        if nslab > 1:
            slabs = make_slabs(sg, nslab)
        else:
            fl = dict(sg.filters)
            fl.setdefault("allc", np.arange(sg.bounds.begc, sg.bounds.endc + 1, dtype=np.int32))
            slabs = [(sg.bounds, fl)]
        self.slabs = []
        for b, fl in slabs:
            nf = {k: len(v) for k, v in fl.items()}
            if mem == abi.MEM_DEVICE:
                import torch
                fl = {k: torch.from_numpy(v if len(v) else np.zeros(1, np.int32)).cuda() for k, v in fl.items()}
            else:
                fl = {k: (v if len(v) else np.zeros(1, np.int32)) for k, v in fl.items()}
            self.slabs.append((b, fl, nf, abi.BalanceReport()))
        self._select(0)

    def _select(self, k: int):
        self.bounds, self.filters, self.nfilter, self.balance_report = self.slabs[k]

    # -- individual routines (names and argument meaning follow the Fortran) ---------------
    def SoilTemperature(self):
        st = abi.Status()
        rc = self.ctx.L.ctsm_b200_soiltemperature(
            self.ctx.h, C.byref(self.bounds), self.nfilter["nolakep"], abi.i32p(self.filters["nolakep"]),
            self.nfilter["nolakec"], abi.i32p(self.filters["nolakec"]), C.byref(self.structs["soiltemperature"]),
            self.mem, C.byref(st))
        if rc != 0:
            raise CtsmError(st, rc)

    def SoilWater(self):
        st = abi.Status()
        rc = self.ctx.L.ctsm_b200_soilwater(
            self.ctx.h, C.byref(self.bounds), self.nfilter["hydrologyc"], abi.i32p(self.filters["hydrologyc"]),
            C.byref(self.structs["soilwater"]), self.mem, C.byref(st))
        if rc != 0:
            raise CtsmError(st, rc)

    def CanopyFluxes(self):
        st = abi.Status()
        rc = self.ctx.L.ctsm_b200_canopyfluxes(
            self.ctx.h, C.byref(self.bounds), self.nfilter["exposedvegp"], abi.i32p(self.filters["exposedvegp"]),
            C.byref(self.structs["canopyfluxes"]), self.mem, C.byref(st))
        if rc != 0:
            raise CtsmError(st, rc)

    def SoilFluxes(self):
        """SoilFluxes (SoilFluxesMod.F90:37; the urban filters of the reference's dummy list are empty on this path)"""
        st = abi.Status()
        rc = self.ctx.L.ctsm_b200_soilfluxes(
            self.ctx.h, C.byref(self.bounds), self.nfilter["nolakec"], abi.i32p(self.filters["nolakec"]),
            self.nfilter["nolakep"], abi.i32p(self.filters["nolakep"]), C.byref(self.structs["soilfluxes"]), self.mem,
            C.byref(st))
        if rc != 0:
            raise CtsmError(st, rc)

    def Patch2Col(self):
        """clm_drv_patch2col (clm_driver.F90:1655)"""
        st = abi.Status()
        rc = self.ctx.L.ctsm_b200_patch2col(
            self.ctx.h, C.byref(self.bounds), self.nfilter["allc"], abi.i32p(self.filters["allc"]),
            self.nfilter["nolakec"], abi.i32p(self.filters["nolakec"]), C.byref(self.structs["patch2col"]), self.mem,
            C.byref(st))
        if rc != 0:
            raise CtsmError(st, rc)

    def VertTranSink(self):
        """Compute_EffecRootFrac_And_VertTranSink (SoilWaterPlantSinkMod.F90:18-142): _HydStress (:236-328) with plant
        hydraulic stress, _Default (:332-424) without"""
        st = abi.Status()
        fn = self.ctx.L.ctsm_b200_vert_tran_sink_hydstress if self.ctx.prm.use_hydrstress else self.ctx.L.ctsm_b200_vert_tran_sink_default
        rc = fn(
            self.ctx.h, C.byref(self.bounds), self.nfilter["hydrologyc"], abi.i32p(self.filters["hydrologyc"]),
            C.byref(self.structs["plantsink"]), self.mem, C.byref(st))
        if rc != 0:
            raise CtsmError(st, rc)

    def BalanceCheck(self):
        """BalanceCheck + EnergyBalanceCheck over all columns in bounds (BalanceCheckMod.F90:445,859)"""
        st = abi.Status()
        self.ctx._pending.append(self.balance_report)        # filled at the next sync when the call is asynchronous
        rc = self.ctx.L.ctsm_b200_balancecheck(
            self.ctx.h, C.byref(self.bounds), self.nfilter["allc"], abi.i32p(self.filters["allc"]),
            C.byref(self.structs["balancecheck"]), self.danstep, self.mem, C.byref(self.balance_report), C.byref(st))
        if rc != 0:
            raise CtsmError(st, rc)
        if getattr(self, "_dist", None) is not None and self.mem == abi.MEM_DEVICE:
            self._reduce_balance()

    # -- the one collective of the path: global BalanceCheck figures (an addition, the reference is clump-local)
    def enable_global_balance(self, dist):
        """After every BalanceCheck call, MAX-reduce its 7 residual maxima over the ranks of `dist` (NCCL) on a side
        stream, straight from the device slot the kernels wrote (ctsm_b200_balance_device_maxima): no host round trip,
        nothing on the compute stream waits for it."""
        import torch
        self._dist = dist
        self._side = torch.cuda.Stream()
        self._gmax = torch.zeros(7, dtype=torch.float64, device="cuda")
        self._ext = torch.cuda.ExternalStream(self.ctx.stream_ptr)

    def _reduce_balance(self):
        import torch

        class _Dev:                      # the library's device slot as a CUDA array (7 doubles)
            def __init__(self, ptr):
                self.__cuda_array_interface__ = {"shape": (7,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}

        ptr = self.ctx.L.ctsm_b200_balance_device_maxima(self.ctx.h)
        if not ptr:
            return
        src = torch.as_tensor(_Dev(ptr), device="cuda")
        self._side.wait_stream(self._ext)
        with torch.cuda.stream(self._side):
            local = src.clone()
            self._dist.all_reduce(local, op=self._dist.ReduceOp.MAX)
            torch.maximum(self._gmax, local, out=self._gmax)

    def global_balance(self):
        """max |residual| over all ranks and all BalanceCheck calls since enable_global_balance (order CTSM_BAL_*)."""
        if getattr(self, "_dist", None) is None:
            return None
        import torch
        self._side.synchronize()
        return [float(x) for x in self._gmax.cpu()]

    def BiogeophysPreFluxCalcs(self, time_flags: int = 0):
        """BiogeophysPreFluxCalcs (BiogeophysPreFluxCalcsMod.F90:58-118); no urban columns"""
        st = abi.Status()
        rc = self.ctx.L.ctsm_b200_biogeophys_pre_flux_calcs(
            self.ctx.h, C.byref(self.bounds), self.nfilter["nolakec"], abi.i32p(self.filters["nolakec"]),
            self.nfilter["nolakep"], abi.i32p(self.filters["nolakep"]), 0, None, int(time_flags),
            C.byref(self.structs["preflux"]), self.mem, C.byref(st))
        if rc != 0:
            raise CtsmError(st, rc)

    def CalculateSurfaceHumidity(self):
        """CalculateSurfaceHumidity (SurfaceHumidityMod.F90:41-239)"""
        st = abi.Status()
        rc = self.ctx.L.ctsm_b200_calculate_surface_humidity(
            self.ctx.h, C.byref(self.bounds), self.nfilter["nolakec"], abi.i32p(self.filters["nolakec"]),
            C.byref(self.structs["surfacehumidity"]), self.mem, C.byref(st))
        if rc != 0:
            raise CtsmError(st, rc)

    def BareGroundFluxes(self):
        """BareGroundFluxes (BareGroundFluxesMod.F90:63-529) over filter_noexposedvegp"""
        st = abi.Status()
        rc = self.ctx.L.ctsm_b200_bare_ground_fluxes(
            self.ctx.h, C.byref(self.bounds), self.nfilter["noexposedvegp"], abi.i32p(self.filters["noexposedvegp"]),
            C.byref(self.structs["baregroundfluxes"]), self.mem, C.byref(st))
        if rc != 0:
            raise CtsmError(st, rc)

    def HydrologyInfiltration(self):
        """SetSoilWaterFractions ... TotalSurfaceRunoff, the call sequence HydrologyNoDrainageMod.F90:297-337; no urban columns"""
        st = abi.Status()
        rc = self.ctx.L.ctsm_b200_hydrology_infiltration(
            self.ctx.h, C.byref(self.bounds), self.nfilter["nolakec"], abi.i32p(self.filters["nolakec"]),
            self.nfilter["hydrologyc"], abi.i32p(self.filters["hydrologyc"]), 0, None,
            C.byref(self.structs["infiltration"]), self.mem, C.byref(st))
        if rc != 0:
            raise CtsmError(st, rc)

    def BuildSnowFilter(self):
        """BuildSnowFilter (SnowHydrologyMod.F90:3975): filter_snowc / filter_nosnowc of the current clump from col%snl, kept on the
        side the arrays live on; returns the two lists as numpy arrays"""
        n = self.nfilter["nolakec"]
        na, nb = C.c_int32(0), C.c_int32(0)
        snl = self.arrays["snl"]
        bufs = self.filters.get("_snowbufs")                  # the two output lists live as long as the clump (no per-step allocation)
        if bufs is None:
            if self.mem == abi.MEM_DEVICE:
                import torch
                # torch fills on ITS stream; the library's stream is non-blocking, so an asynchronous fill could land after the
                # library has written the lists: allocate without a fill and drain torch's stream once
                bufs = (torch.empty(max(n, 1), dtype=torch.int32, device="cuda"), torch.empty(max(n, 1), dtype=torch.int32, device="cuda"))
                torch.cuda.current_stream().synchronize()
            else:
                bufs = (np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.int32))
            self.filters["_snowbufs"] = bufs
        a, b = bufs
        rc = self.ctx.L.ctsm_b200_build_snow_filter(
            self.ctx.h, C.byref(self.bounds), n, abi.i32p(self.filters["nolakec"]), abi.i32p(snl), self.sg.bounds.begc,
            self.sg.bounds.endc, abi.i32p(a), C.byref(na), abi.i32p(b), C.byref(nb), abi.MEM_DEVICE if self.mem == abi.MEM_DEVICE else abi.MEM_HOST)
        if rc != 0:
            raise RuntimeError("ctsm_b200_build_snow_filter rc=%d" % rc)
        self.filters["snowc"], self.filters["nosnowc"] = a, b
        self.nfilter["snowc"], self.nfilter["nosnowc"] = int(na.value), int(nb.value)
        to_np = (lambda t, k: t[:k].cpu().numpy()) if self.mem == abi.MEM_DEVICE else (lambda t, k: t[:k].copy())
        return to_np(a, na.value), to_np(b, nb.value)

    def SnowWater(self):
        """BuildSnowFilter + SnowWater, HydrologyNoDrainageMod.F90:279-285 (SnowHydrologyMod.F90:1015)"""
        self.BuildSnowFilter()
        st = abi.Status()
        rc = self.ctx.L.ctsm_b200_snow_water(
            self.ctx.h, C.byref(self.bounds), self.nfilter["snowc"], abi.i32p(self.filters["snowc"]), self.nfilter["nosnowc"],
            abi.i32p(self.filters["nosnowc"]), C.byref(self.structs["snowwater"]), self.mem, C.byref(st))
        if rc != 0:
            raise CtsmError(st, rc)

    def SnowCapping(self, nstep: int = 1000):
        """SnowCapping over (filter_nolakec, filter_snowc), HydrologyNoDrainageMod.F90:377 (SnowHydrologyMod.F90:3121)"""
        if "snowc" not in self.filters:
            self.BuildSnowFilter()
        st = abi.Status()
        rc = self.ctx.L.ctsm_b200_snow_capping(
            self.ctx.h, C.byref(self.bounds), self.nfilter["nolakec"], abi.i32p(self.filters["nolakec"]), self.nfilter["snowc"],
            abi.i32p(self.filters["snowc"]), C.byref(self.structs["snowcapping"]), int(nstep), self.mem, C.byref(st))
        if rc != 0:
            raise CtsmError(st, rc)

    def SnowLayers(self):
        """SnowCompaction, CombineSnowLayers, DivideSnowLayers, ZeroEmptySnowLayers over the snow filter built at the start of
        HydrologyNoDrainage (HydrologyNoDrainageMod.F90:381-399)"""
        if "snowc" not in self.filters:
            self.BuildSnowFilter()
        st = abi.Status()
        rc = self.ctx.L.ctsm_b200_snow_layers(
            self.ctx.h, C.byref(self.bounds), self.nfilter["snowc"], abi.i32p(self.filters["snowc"]),
            C.byref(self.structs["snowlayers"]), self.mem, C.byref(st))
        if rc != 0:
            raise CtsmError(st, rc)
        self.filters.pop("snowc", None)            # snl has changed: the next user rebuilds the filter (:402)

    def WaterTable(self):
        """PerchedWaterTable, ThetaBasedWaterTable, RenewCondensation (HydrologyNoDrainageMod.F90:359-373)"""
        st = abi.Status()
        rc = self.ctx.L.ctsm_b200_water_table(
            self.ctx.h, C.byref(self.bounds), self.nfilter["hydrologyc"], abi.i32p(self.filters["hydrologyc"]), 0, None,
            C.byref(self.structs["watertable"]), self.mem, C.byref(st))
        if rc != 0:
            raise CtsmError(st, rc)

    def HydrologyDiagnostics(self):
        """BuildSnowFilter (:402) + the diagnostics that close HydrologyNoDrainage (:420-757)"""
        if "snowc" not in self.filters:
            self.BuildSnowFilter()
        st = abi.Status()
        rc = self.ctx.L.ctsm_b200_hydrology_diagnostics(
            self.ctx.h, C.byref(self.bounds), self.nfilter["nolakec"], abi.i32p(self.filters["nolakec"]),
            self.nfilter["snowc"], abi.i32p(self.filters["snowc"]), self.nfilter["nosnowc"], abi.i32p(self.filters["nosnowc"]),
            self.nfilter["hydrologyc"], abi.i32p(self.filters["hydrologyc"]), 0, None,
            C.byref(self.structs["hydrodiag"]), self.mem, C.byref(st))
        if rc != 0:
            raise CtsmError(st, rc)

    def call(self, g):
        {"snowcapping": self.SnowCapping, "watertable": self.WaterTable, "hydrodiag": self.HydrologyDiagnostics, "snowwater": self.SnowWater, "snowlayers": self.SnowLayers, "infiltration": self.HydrologyInfiltration, "preflux": self.BiogeophysPreFluxCalcs, "surfacehumidity": self.CalculateSurfaceHumidity, "baregroundfluxes": self.BareGroundFluxes,
         "canopyfluxes": self.CanopyFluxes, "soiltemperature": self.SoilTemperature, "soilwater": self.SoilWater,
         "plantsink": self.VertTranSink, "balancecheck": self.BalanceCheck, "soilfluxes": self.SoilFluxes, "patch2col": self.Patch2Col}[g]()

    def step(self):
        if self.window:
            rc = self.ctx.L.ctsm_b200_host_window_begin(self.ctx.h)
            if rc != 0:
                raise RuntimeError("ctsm_b200_host_window_begin rc=%d" % rc)
        try:
            for k in range(len(self.slabs)):
                self._select(k)
                for g in self.routines:
                    self.call(g)
        finally:
            self._select(0)
            if self.window:
                st = abi.Status()
                rc = self.ctx.L.ctsm_b200_host_window_end(self.ctx.h, C.byref(st))
                self.ctx._pending.clear()
                if rc != 0:
                    raise CtsmError(st, rc)

    def window_bytes(self):
        """(h2d, d2h) bytes the last resident window moved, counted by the library from the copies it issued."""
        a, b = C.c_uint64(), C.c_uint64()
        self.ctx.L.ctsm_b200_host_window_bytes(self.ctx.h, C.byref(a), C.byref(b))
        return int(a.value), int(b.value)


def staged_bytes(sg, routines: Iterable[str], preserve_out: bool = True):
    """Bytes a CTSM_MEM_HOST step moves over PCIe (fields + filters), from the field table."""
    h2d = d2h = 0
    for g in routines:
        for fs in abi.FIELDS[g]:
            n = sg.bounds.extent(fs.sub) * fs.nlev * (8 if fs.ctype == "double" else 4)
            if fs.intent in ("IN", "INOUT") or preserve_out:
                h2d += n
            if fs.intent in ("OUT", "INOUT"):
                d2h += n
    for g in routines:
        h2d += 4 * sum(len(sg.filters[k]) if k in sg.filters else sg.ncol for k in FILTER_OF[g])
    return h2d, d2h


def algorithmic_bytes(sg, S, group: str) -> Dict[str, float]:
    """Algorithmic bytes per launch of one routine, from the field table:
    for every field, (levels >= 1 touched + active snow levels if touched) x element size x units in the
    filter, counted once for IN or OUT and twice for INOUT."""
    if group == "soilwater":
        ncol = len(sg.filters["hydrologyc"]); cols = sg.filters["hydrologyc"] - 1
        npat = 0; pats = np.zeros(0, dtype=np.int64)
    elif group == "plantsink":
        ncol = len(sg.filters["hydrologyc"]); cols = sg.filters["hydrologyc"] - 1
        pats = np.nonzero(np.isin(sg.patch_column, sg.filters["hydrologyc"]))[0]; npat = len(pats)
    elif group == "patch2col":
        ncol = sg.ncol; cols = np.arange(sg.ncol)
        npat = len(sg.filters["nolakep"]); pats = sg.filters["nolakep"] - 1
    elif group == "soilfluxes":
        ncol = len(sg.filters["nolakec"]); cols = sg.filters["nolakec"] - 1
        npat = len(sg.filters["nolakep"]); pats = sg.filters["nolakep"] - 1
    elif group == "balancecheck":
        ncol = sg.ncol; cols = np.arange(sg.ncol)
        npat = sg.npatch; pats = np.arange(sg.npatch)
    elif group == "baregroundfluxes":
        pats = sg.filters["noexposedvegp"] - 1; npat = len(pats)
        cols = np.unique(sg.patch_column[pats]) - 1; ncol = len(cols)
    elif group == "canopyfluxes":
        pats = sg.filters["exposedvegp"] - 1; npat = len(pats)
        cols = np.unique(sg.patch_column[pats]) - 1; ncol = len(cols)      # each column counted once per step
    else:
        ncol = len(sg.filters["nolakec"]); cols = sg.filters["nolakec"] - 1
        npat = len(sg.filters["nolakep"]); pats = sg.filters["nolakep"] - 1
    snow_c = float(np.mean(-S["snl"][cols])) if ncol else 0.0
    snow_p = float(np.mean(-S["snl"][sg.patch_column[pats] - 1])) if npat else 0.0
    total = 0.0
    for fs in abi.FIELDS[group]:
        es = 8 if fs.ctype == "double" else 4
        if fs.sub in ("PFT", "GRC"):
            continue          # parameter tables / per-gridcell scalars: negligible, shared by many units
        units, snow = (npat, snow_p) if fs.sub == "PATCH" else (ncol, snow_c)
        lev = fs.used_soil + fs.used_snow * snow
        mult = 2 if fs.intent == "INOUT" else 1
        total += units * lev * es * mult
    return {"bytes": total, "columns": ncol, "patches": npat, "bytes_per_column": total / max(ncol, 1),
            "bytes_per_patch": total / max(npat, 1)}
