"""ctypes loader for the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs import this module; the
product package ctsm_b200 never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

from ctsm_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
_lib = None


class Clump(C.Structure):
    _fields_ = [("bounds", abi.Bounds),
                ("num_nolakep", C.c_int32), ("filter_nolakep", C.POINTER(C.c_int32)),
                ("num_nolakec", C.c_int32), ("filter_nolakec", C.POINTER(C.c_int32)),
                ("num_hydrologyc", C.c_int32), ("filter_hydrologyc", C.POINTER(C.c_int32)),
                ("num_exposedvegp", C.c_int32), ("filter_exposedvegp", C.POINTER(C.c_int32))]


def make_clumps(sg, nclumps):
    """Deal contiguous gridcell ranges to clumps (decompInitMod.F90:96-161 gives each clump whole
    gridcells) and cut the proc-level filters at the clump bounds.  Returns (ctypes array, keepalive)."""
    import numpy as np
    ng = sg.ngrc
    nclumps = max(1, min(nclumps, ng))
    edges = np.linspace(0, ng, nclumps + 1).astype(np.int64)
    arr = (Clump * nclumps)()
    keep = []
    for k in range(nclumps):
        g0, g1 = int(edges[k]) + 1, int(edges[k + 1])
        cols = np.nonzero((sg.col_gridcell >= g0) & (sg.col_gridcell <= g1))[0]
        b = sg.bounds.copy()
        b.begg, b.endg = g0, g1
        b.begc, b.endc = int(cols[0]) + 1, int(cols[-1]) + 1
        b.begl, b.endl = b.begc, b.endc
        b.begp, b.endp = int(sg.col_patchi[cols[0]]), int(sg.col_patchf[cols[-1]])
        b.level, b.clump_index = 2, k + 1
        arr[k].bounds = b
        for name, lo, hi in (("nolakep", b.begp, b.endp), ("nolakec", b.begc, b.endc),
                             ("hydrologyc", b.begc, b.endc), ("exposedvegp", b.begp, b.endp)):
            f = sg.filters.get(name)
            if f is None:
                f = np.zeros(0, dtype=np.int32)
            sub = np.ascontiguousarray(f[(f >= lo) & (f <= hi)])
            keep.append(sub)
            setattr(arr[k], "num_" + name, len(sub))
            setattr(arr[k], "filter_" + name, abi.i32p(sub))
    return arr, keep


def clump_for_gridcells(sg, g0, g1):
    """One clump covering the contiguous gridcell range [g0, g1] (1-based, inclusive)."""
    import numpy as np
    arr = (Clump * 1)()
    keep = []
    cols = np.nonzero((sg.col_gridcell >= g0) & (sg.col_gridcell <= g1))[0]
    b = sg.bounds.copy()
    b.begg, b.endg = g0, g1
    b.begc, b.endc = int(cols[0]) + 1, int(cols[-1]) + 1
    b.begl, b.endl = b.begc, b.endc
    b.begp, b.endp = int(sg.col_patchi[cols[0]]), int(sg.col_patchf[cols[-1]])
    b.level, b.clump_index = 2, 1
    arr[0].bounds = b
    for name, lo, hi in (("nolakep", b.begp, b.endp), ("nolakec", b.begc, b.endc),
                         ("hydrologyc", b.begc, b.endc), ("exposedvegp", b.begp, b.endp)):
        f = sg.filters.get(name)
        sub = np.ascontiguousarray(f[(f >= lo) & (f <= hi)]) if f is not None else np.zeros(0, dtype=np.int32)
        if len(sub) == 0:
            sub = np.zeros(1, dtype=np.int32); n = 0
        else:
            n = len(sub)
        keep.append(sub)
        setattr(arr[0], "num_" + name, n)
        setattr(arr[0], "filter_" + name, abi.i32p(sub))
    return arr, keep


def build(force: bool = False) -> str:
    srcs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(".c") or f.endswith(".h")]
    srcs += [os.path.join(HERE, "..", "include", f) for f in ("ctsm_b200.h", "ctsm_b200_fields.def", "ctsm_b200_defaults.h")]
    stale = (not os.path.exists(LIB)) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", HERE, "liboracle.so"])
    return LIB


_pert = None


def lib_perturbed():
    """The conditioning probe (oracle_pert.h): same exports as lib(), CanopyFluxes / PHS libm results moved by one ulp."""
    global _pert
    if _pert is None:
        subprocess.check_call(["make", "-s", "-C", HERE, "liboracle_pert.so"])
        _pert = _bind(C.CDLL(os.path.join(HERE, "liboracle_pert.so")))
    return _pert


def default_params():
    """ctsm_params_t with the clm6_0 defaults, from the oracle library alone (does not load libctsm_b200.so)."""
    p = abi.Params()
    lib().oracle_default_params(C.byref(p))
    return p


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    _lib = _bind(C.CDLL(LIB))
    return _lib


def _bind(L):
    i32p, f64p = C.POINTER(C.c_int32), C.POINTER(C.c_double)
    B, S, P = C.POINTER(abi.Bounds), C.POINTER(abi.Status), C.POINTER(abi.Params)
    L.oracle_dgbsv.argtypes = [C.c_int] * 4 + [f64p, C.c_int, i32p, f64p, C.c_int, i32p]
    L.oracle_dgbsv.restype = None
    L.oracle_dgtsv.argtypes = [C.c_int, C.c_int, f64p, f64p, f64p, f64p, C.c_int, i32p]
    L.oracle_dgtsv.restype = None
    L.oracle_tridiagonal.argtypes = [B, C.c_int, C.c_int, i32p, C.c_int, i32p, f64p, f64p, f64p, f64p, f64p]
    L.oracle_tridiagonal.restype = None
    L.oracle_banddiagonal.argtypes = [B, C.c_int, C.c_int, i32p, i32p, C.c_int, i32p, C.c_int, f64p, f64p, f64p, S]
    L.oracle_dgtsv_batch.argtypes = [B, C.c_int, i32p, C.c_int, i32p, f64p, f64p, f64p, f64p, f64p, S]
    L.oracle_soilwater.argtypes = [P, B, C.c_int, i32p, C.POINTER(abi.STRUCTS["soilwater"]), S]
    L.oracle_soiltemperature.argtypes = [P, B, C.c_int, i32p, C.c_int, i32p,
                                         C.POINTER(abi.STRUCTS["soiltemperature"]), S]
    L.oracle_canopyfluxes.argtypes = [P, B, C.c_int, i32p, C.POINTER(abi.STRUCTS["canopyfluxes"]), S]
    L.oracle_set_exposedvegp_filter.argtypes = [B, C.c_int, i32p, i32p, i32p, i32p, i32p, i32p]
    L.oracle_set_exposedvegp_filter.restype = None
    L.oracle_qsat.argtypes = [C.c_double, C.c_double, f64p, f64p, f64p]
    L.oracle_qsat.restype = None
    L.oracle_moninobukini.argtypes = [C.c_double] * 6 + [f64p, f64p]
    L.oracle_moninobukini.restype = None
    L.oracle_quadratic.argtypes = [C.c_double] * 3 + [f64p, f64p]
    L.oracle_plc.argtypes = [C.c_double] * 3
    L.oracle_plc.restype = C.c_double
    L.oracle_d1plc.argtypes = [C.c_double] * 3
    L.oracle_d1plc.restype = C.c_double
    L.oracle_balancecheck.argtypes = [P, B, C.c_int, i32p, C.POINTER(abi.STRUCTS["balancecheck"]), C.c_int,
                                      C.POINTER(abi.BalanceReport), S]
    L.oracle_balancecheck_skip_steps.argtypes = [C.c_double]
    L.oracle_vert_tran_sink_hydstress.argtypes = [B, C.c_int, i32p, C.POINTER(abi.STRUCTS["plantsink"])]
    L.oracle_biogeophys_pre_flux_calcs.argtypes = [P, B, C.c_int, i32p, C.c_int, i32p, C.c_int, C.c_int,
                                                   C.POINTER(abi.STRUCTS["preflux"]), C.POINTER(abi.Status)]
    L.oracle_calculate_surface_humidity.argtypes = [B, C.c_int, i32p, C.POINTER(abi.STRUCTS["surfacehumidity"]), C.POINTER(abi.Status)]
    L.oracle_bare_ground_fluxes.argtypes = [P, B, C.c_int, i32p, C.POINTER(abi.STRUCTS["baregroundfluxes"]), C.POINTER(abi.Status)]
    L.oracle_hydrology_infiltration.argtypes = [P, B, C.c_int, i32p, C.c_int, i32p, C.c_int, C.POINTER(abi.STRUCTS["infiltration"]),
                                                C.POINTER(abi.Status)]
    dp = C.POINTER(C.c_double)
    L.oracle_snow_dz_limits.argtypes = [P, dp, dp, dp]
    L.oracle_snow_dz_limits.restype = None
    L.oracle_build_snow_filter.argtypes = [C.c_int, i32p, i32p, C.c_int, i32p, i32p, i32p, i32p]
    L.oracle_build_snow_filter.restype = None
    L.oracle_snow_water.argtypes = [P, B, C.c_int, i32p, C.c_int, i32p, C.POINTER(abi.STRUCTS["snowwater"]), C.POINTER(abi.Status)]
    L.oracle_snow_capping.argtypes = [P, B, C.c_int, i32p, C.c_int, i32p, C.POINTER(abi.STRUCTS["snowcapping"]), C.c_int, C.POINTER(abi.Status)]
    L.oracle_snow_layers.argtypes = [P, B, C.c_int, i32p, C.POINTER(abi.STRUCTS["snowlayers"]), C.POINTER(abi.Status)]
    L.oracle_calc_ozone_uptake.argtypes = [P, B, C.c_int, i32p, C.POINTER(abi.STRUCTS["ozone"]), C.POINTER(abi.Status)]
    L.oracle_calc_ozone_stress.argtypes = [B, C.c_int, i32p, C.c_int, i32p, C.c_int, C.c_int, C.POINTER(abi.STRUCTS["ozone"]), C.POINTER(abi.Status)]
    L.oracle_water_table.argtypes = [P, B, C.c_int, i32p, C.c_int, C.POINTER(abi.STRUCTS["watertable"]), C.POINTER(abi.Status)]
    L.oracle_hydrology_diagnostics.argtypes = [P, B, C.c_int, i32p, C.c_int, i32p, C.c_int, i32p, C.c_int, i32p, C.c_int,
                                               C.POINTER(abi.STRUCTS["hydrodiag"]), C.POINTER(abi.Status)]
    L.oracle_vert_tran_sink_default.argtypes = [B, C.c_int, i32p, C.POINTER(abi.STRUCTS["plantsinkdefault"])]
    L.oracle_set_plantsink_default.argtypes = [C.POINTER(abi.STRUCTS["plantsinkdefault"])]
    L.oracle_set_plantsink_default.restype = None
    L.oracle_num_threads.restype = C.c_int
    L.oracle_set_pert_mode.argtypes = [C.c_int]
    L.oracle_set_pert_mode.restype = None
    L.oracle_default_params.argtypes = [P]
    L.oracle_default_params.restype = None
    L.oracle_step_clumps.argtypes = [P, C.c_int, C.POINTER(Clump), C.POINTER(abi.STRUCTS["soiltemperature"]),
                                     C.POINTER(abi.STRUCTS["soilwater"]), C.POINTER(abi.STRUCTS["canopyfluxes"]), C.c_int]
    L.oracle_fullstep_clumps.argtypes = [P, C.c_int, C.POINTER(Clump), C.POINTER(abi.STRUCTS["soiltemperature"]),
                                         C.POINTER(abi.STRUCTS["soilwater"]), C.POINTER(abi.STRUCTS["canopyfluxes"]),
                                         C.POINTER(abi.STRUCTS["plantsink"]), C.POINTER(abi.STRUCTS["balancecheck"]),
                                         C.POINTER(abi.STRUCTS["soilfluxes"]), C.POINTER(abi.STRUCTS["patch2col"]),
                                         C.c_int, C.c_int]
    L.oracle_begin_water_column_balance.argtypes = [B, C.c_int, i32p, C.c_int, i32p, C.POINTER(abi.STRUCTS["waterbalance"]),
                                                    C.c_double, S]
    L.oracle_water_gridcell_balance.argtypes = [B, C.c_int, i32p, C.c_int, i32p, C.POINTER(abi.STRUCTS["watergridbalance"]),
                                                C.c_double, C.c_int, S]
    L.oracle_set_filters.argtypes = [B, C.POINTER(abi.FilterInputs), C.POINTER(abi.Filters)]
    L.oracle_patch2col.argtypes = [B, C.c_int, i32p, C.c_int, i32p, C.POINTER(abi.STRUCTS["patch2col"])]
    L.oracle_soilfluxes.argtypes = [P, B, C.c_int, i32p, C.c_int, i32p, C.POINTER(abi.STRUCTS["soilfluxes"]), S]
    return L
