"""ctypes binding of the C ABI in include/ctsm_b200.h.

The field structs are generated from include/ctsm_b200_fields.def (the same
X-macro table the C header and the CUDA library are generated from), so the
three views cannot drift apart.

Nothing in this module computes: it loads ``libctsm_b200.so`` (the CUDA
library; there is no CPU fallback) and marshals pointers.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEF_PATH = os.path.join(ROOT, "include", "ctsm_b200_fields.def")
LIB_PATH = os.environ.get("CTSM_B200_LIB", os.path.join(ROOT, "ctsm_b200", "lib", "libctsm_b200.so"))

NLEVSNO, NLEVGRND, NLEVSOI, NVEGWCS, NLEVCAN, MXPFT = 12, 25, 20, 4, 1, 78

# LEV token -> (lower bound, number of levels); see ctsm_b200_fields.def header
LEV = {
    "L1": (1, 1),
    "SNOSOI": (-NLEVSNO + 1, NLEVSNO + NLEVGRND),
    "SNOSOI0": (-NLEVSNO, NLEVSNO + NLEVGRND + 1),
    "GRND": (1, NLEVGRND),
    "SOI": (1, NLEVSOI),
    "SNO": (-NLEVSNO + 1, NLEVSNO),
    "SNO1": (-NLEVSNO + 1, NLEVSNO + 1),
    "VEGWCS": (1, NVEGWCS),
    "CAN": (1, NLEVCAN),
    "LAK": (1, 10),
    "PHS2": (1, 2 * NLEVCAN),
    "NUMRAD": (1, 2),
    "AER": (1, 14),
    "PFT": (0, MXPFT + 1),
    "PFTVEGWCS": (0, (MXPFT + 1) * NVEGWCS),
}

MEM_DEVICE, MEM_HOST, MEM_HOST_NOPRESERVE = 0, 1, 3

ISTSOIL, ISTCROP, ISTICE, ISTDLAK, ISTWET, ISTURB_MIN, ISTURB_MAX = 1, 2, 4, 5, 6, 7, 9


class Bounds(C.Structure):
    """decompMod.F90:60-68 bounds_type."""
    _fields_ = [(n, C.c_int32) for n in (
        "begg", "endg", "begl", "endl", "begc", "endc", "begp", "endp",
        "begCohort", "endCohort", "level", "clump_index")]

    def extent(self, sub: str) -> int:
        if sub == "PFT":
            return MXPFT + 1
        b, e = SUB_BOUNDS[sub]
        return getattr(self, e) - getattr(self, b) + 1

    def beg(self, sub: str) -> int:
        if sub == "PFT":
            return 0
        return getattr(self, SUB_BOUNDS[sub][0])

    def copy(self) -> "Bounds":
        o = Bounds()
        C.memmove(C.byref(o), C.byref(self), C.sizeof(Bounds))
        return o


SUB_BOUNDS = {"GRC": ("begg", "endg"), "LUN": ("begl", "endl"), "COL": ("begc", "endc"),
              "PATCH": ("begp", "endp"), "PFT": (None, None)}


class Status(C.Structure):
    _fields_ = [("code", C.c_int32), ("subgrid_level", C.c_int32), ("subgrid_index", C.c_int32),
                ("info", C.c_int32), ("value", C.c_double), ("n_warnings", C.c_int32),
                ("reserved", C.c_int32), ("msg", C.c_char * 160)]


class Params(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("device", C.c_int32),
                ("nlevsno", C.c_int32), ("nlevgrnd", C.c_int32), ("nlevsoi", C.c_int32),
                ("dtime", C.c_double),
                ("upper_boundary_condition", C.c_int32), ("lower_boundary_condition", C.c_int32),
                ("flux_calculation", C.c_int32),
                ("dtmin", C.c_double), ("verySmall", C.c_double), ("xTolerUpper", C.c_double),
                ("xTolerLower", C.c_double), ("e_ice", C.c_double),
                ("snow_thermal_cond_method", C.c_int32), ("snow_thermal_cond_glc_method", C.c_int32),
                ("itmax_canopy_fluxes", C.c_int32), ("use_undercanopy_stability", C.c_int32),
                ("use_biomass_heat_storage", C.c_int32), ("z0param_method", C.c_int32),
                ("soil_resis_method", C.c_int32), ("use_hydrstress", C.c_int32), ("use_luna", C.c_int32),
                ("stomatalcond_mtd", C.c_int32), ("light_inhibit", C.c_int32),
                ("modifyphoto_and_lmr_forcrop", C.c_int32)] + [
                (n, C.c_double) for n in (
                    "lai_dl", "z_dl", "a_coef", "a_exp", "csoilc", "cv", "wind_min", "zetamaxstable", "leaf_mr_vcm",
                    "act25", "fnr", "cp25_yr2000", "kc25_coef", "ko25_coef", "fnps", "theta_psii", "theta_ip",
                    "vcmaxha", "jmaxha", "tpuha", "lmrha", "kcha", "koha", "cpha",
                    "vcmaxhd", "jmaxhd", "tpuhd", "lmrhd", "lmrse",
                    "tpu25ratio", "kp25ratio", "vcmaxse_sf", "jmaxse_sf", "tpuse_sf", "jmax25top_sf")] + [
                ("balance_skip_steps", C.c_int32),
                ("npft_table", C.c_int32), ("calc_human_stress_indices", C.c_int32), ("use_z0m_snowmelt", C.c_int32),
                ("h2osfcflag", C.c_int32), ("crop_fsat_equals_zero", C.c_int32), ("reserved_i", C.c_int32 * 2)] + [
                (n, C.c_double) for n in ("zlnd", "zsno", "zglc", "d_max", "frac_sat_soil_dsl_init", "fff", "pc", "mu")] + [
                (n, C.c_int32) for n in ("snow_overburden_compaction_method", "wind_dependent_snow_density", "use_subgrid_fluxes",
                                         "snicar_use_aerosol")] + [
                (n, C.c_double) for n in (
                    "snow_dzmin_1", "snow_dzmin_2", "snow_dzmax_l_1", "snow_dzmax_l_2", "snow_dzmax_u_1", "snow_dzmax_u_2",
                    "overburden_compress_Tfactor", "int_snow_max", "wimp", "ssi", "drift_gs", "eta0_anderson", "eta0_vionnet",
                    "rho_max", "tau_ref", "ceta", "snw_rds_min", "upplim_destruct_metamorph", "scvng_fct_mlt_sf",
                    "scvng_fct_mlt_bcphi", "scvng_fct_mlt_bcpho", "scvng_fct_mlt_dst1", "scvng_fct_mlt_dst2", "scvng_fct_mlt_dst3",
                    "scvng_fct_mlt_dst4", "h2osno_max", "reset_snow_glc_ela")] + [
                ("reset_snow", C.c_int32), ("reset_snow_glc", C.c_int32)]


def default_params(dtime: float = 1800.0, device: int = 0) -> Params:
    """ctsm_b200_default_params (clm6_0 namelist defaults, namelist_defaults_ctsm.xml) with dtime/device set.
    Pure host code inside the library: works without a GPU."""
    p = Params()
    lib().ctsm_b200_default_params(C.byref(p))
    p.device = device
    p.dtime = dtime
    return p


@dataclass(frozen=True)
class FieldSpec:
    name: str
    ctype: str      # "double" | "int"
    sub: str        # GRC | LUN | COL | PATCH | PFT
    lev: str        # key of LEV
    intent: str     # IN | OUT | INOUT
    used_soil: int
    used_snow: int
    ref: str

    @property
    def dtype(self):
        return np.float64 if self.ctype == "double" else np.int32

    @property
    def lo(self) -> int:
        return LEV[self.lev][0]

    @property
    def nlev(self) -> int:
        return LEV[self.lev][1]


_F_RE = re.compile(r'CTSM_F\(\s*(\w+)\s*,\s*(\w+)\s*,\s*(\w+)\s*,\s*(\w+)\s*,\s*(\w+)\s*,'
                   r'\s*(\d+)\s*,\s*(\d+)\s*,\s*"([^"]*)"\s*\)')


def parse_fields(path: str = DEF_PATH) -> Dict[str, List[FieldSpec]]:
    groups: Dict[str, List[FieldSpec]] = {}
    cur: Optional[str] = None
    with open(path) as fh:
        for line in fh:
            m = re.match(r"\s*#ifdef\s+CTSM_FIELDS_(\w+)", line)
            if m:
                cur = m.group(1).lower()
                groups[cur] = []
                continue
            if re.match(r"\s*#endif", line):
                cur = None
                continue
            m = _F_RE.search(line)
            if m and cur is not None:
                n, t, s, lv, it, us, usn, ref = m.groups()
                groups[cur].append(FieldSpec(n, t, s, lv, it, int(us), int(usn), ref))
    return groups


FIELDS = parse_fields()


def _make_struct(group: str):
    flds = [("alloc", Bounds)]
    for fs in FIELDS[group]:
        flds.append((fs.name, C.POINTER(C.c_double if fs.ctype == "double" else C.c_int32)))
    return type("ctsm_%s_fields_t" % group, (C.Structure,), {"_fields_": flds})


STRUCTS = {g: _make_struct(g) for g in FIELDS}


def _ptr(a, ctype):
    """Raw pointer of a numpy array or a torch tensor (host or device)."""
    if a is None:
        return C.cast(None, C.POINTER(ctype))
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"], "field arrays must be contiguous (level-major, subgrid fastest)"
        return a.ctypes.data_as(C.POINTER(ctype))
    # torch tensor
    assert a.is_contiguous()
    return C.cast(a.data_ptr(), C.POINTER(ctype))


def make_struct(group: str, arrays: dict, alloc: Bounds):
    """Fill ctsm_<group>_fields_t from name -> array (numpy or torch).

    Arrays are stored level-major: shape (nlev, n_subgrid) C-contiguous, which
    is the Fortran (n_subgrid, nlev) column-major layout; 1-D fields have
    shape (n_subgrid,).
    """
    st = STRUCTS[group]()
    st.alloc = alloc
    for fs in FIELDS[group]:
        a = arrays[fs.name]
        n = alloc.extent(fs.sub)
        if fs.sub == "PFT":         # parameter tables: (mxpft+1) x members (ctsm_params_t::npft_table)
            n = int(a.shape[-1])
            assert n % (MXPFT + 1) == 0, "%s.%s: table length %d is not a multiple of mxpft+1" % (group, fs.name, n)
        want = (n,) if fs.lev == "L1" else (fs.nlev, n)
        assert tuple(a.shape) == want, "%s.%s: shape %s != %s" % (group, fs.name, tuple(a.shape), want)
        kind = "float64" if fs.ctype == "double" else "int32"
        assert str(a.dtype).endswith(kind), "%s.%s: dtype %s" % (group, fs.name, a.dtype)
        setattr(st, fs.name, _ptr(a, C.c_double if fs.ctype == "double" else C.c_int32))
    return st


class BalanceReport(C.Structure):
    """ctsm_balance_report_t"""
    KINDS = ("errh2o_col", "errh2o_grc", "errh2osno", "errsol", "errlon", "errseb", "errsoi_col")
    _fields_ = [("max_abs", C.c_double * 7), ("index", C.c_int32 * 7), ("warn", C.c_int32 * 7),
                ("abort_kind", C.c_int32), ("skip_steps", C.c_int32)]


FILTER_NAMES = ("allc", "lakec", "nolakec", "bgc_soilc", "soilc", "hydrologyc", "urbanc", "nourbanc", "icec", "do_smb_c",
                "lakep", "nolakep", "nolakeurbanp", "bgc_vegp", "soilp", "pcropp", "soilnopcropp", "urbanp", "nourbanp",
                "urbanl", "nourbanl")            # CTSM_FLT_* order (filterMod.F90:29-115)
FILTER_LEVEL = ("COL",) * 10 + ("PATCH",) * 9 + ("LUN",) * 2


class FilterInputs(C.Structure):
    """ctsm_filter_inputs_t"""
    _ARRAYS = ("col_active", "col_landunit", "col_gridcell", "col_hydrologically_active", "lun_active", "lun_lakpoi",
               "lun_urbpoi", "lun_itype", "patch_active", "patch_landunit", "patch_itype", "melt_replaced_by_ice_grc")
    _fields_ = [("alloc", Bounds)] + [(n, C.POINTER(C.c_int32)) for n in _ARRAYS] + [
        (n, C.c_int32) for n in ("include_inactive", "use_cn", "use_fates", "use_fates_bgc", "npcropmin", "npcropmax")]


class Filters(C.Structure):
    """ctsm_filters_t"""
    _fields_ = [("list", C.POINTER(C.c_int32) * len(FILTER_NAMES)), ("num", C.c_int32 * len(FILTER_NAMES))]


class LibraryMissing(RuntimeError):
    pass


_lib = None


def lib():
    """Load libctsm_b200.so (built in-tree by `make` / __graft_entry__.build()).

    Fails loudly when the CUDA extension is missing: there is no CPU path.
    """
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            "%s not found: build the CUDA library with `make` (or __graft_entry__.build()). "
            "ctsm_b200 has no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32p, f64p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_double)
    L.ctsm_b200_default_params.argtypes = [C.POINTER(Params)]
    L.ctsm_b200_default_params.restype = None
    L.ctsm_b200_init.argtypes = [C.POINTER(Params), C.POINTER(vp)]
    L.ctsm_b200_finalize.argtypes = [vp]
    L.ctsm_b200_sync.argtypes = [vp, C.POINTER(Status)]
    L.ctsm_b200_stream.argtypes = [vp]
    L.ctsm_b200_stream.restype = vp
    L.ctsm_b200_stream_of.argtypes = [vp, C.c_int]
    L.ctsm_b200_stream_of.restype = vp
    L.ctsm_b200_launch_count.argtypes = [vp]
    L.ctsm_b200_launch_count.restype = C.c_int64
    L.ctsm_b200_host_register.argtypes = [vp, C.c_uint64]
    L.ctsm_b200_host_unregister.argtypes = [vp]
    L.ctsm_b200_version.restype = C.c_char_p
    L.ctsm_b200_last_cuda_error.restype = C.c_char_p
    L.ctsm_b200_set_member_params.argtypes = [vp, C.c_int, f64p, f64p, f64p, f64p, f64p, i32p, C.c_int, C.c_int]
    L.ctsm_b200_set_member_params.restype = C.c_int
    L.ctsm_b200_set_tuning.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int]
    L.ctsm_b200_set_tuning.restype = C.c_int
    L.ctsm_b200_host_window_begin.argtypes = [vp]
    L.ctsm_b200_host_window_begin.restype = C.c_int
    L.ctsm_b200_host_window_end.argtypes = [vp, C.POINTER(Status)]
    L.ctsm_b200_host_window_end.restype = C.c_int
    L.ctsm_b200_host_window_bytes.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.ctsm_b200_host_window_bytes.restype = C.c_int
    L.ctsm_b200_host_invalidate.argtypes = [vp, vp]
    L.ctsm_b200_host_invalidate.restype = C.c_int
    L.ctsm_b200_balance_device_maxima.argtypes = [vp]
    L.ctsm_b200_balance_device_maxima.restype = vp
    L.ctsm_b200_canopy_round_stats.argtypes = [vp, i32p, i32p, C.c_int]
    L.ctsm_b200_canopy_round_stats.restype = C.c_int
    L.ctsm_b200_tridiagonal.argtypes = [vp, C.POINTER(Bounds), C.c_int, C.c_int, i32p, C.c_int, i32p,
                                        f64p, f64p, f64p, f64p, f64p, C.c_int]
    L.ctsm_b200_banddiagonal.argtypes = [vp, C.POINTER(Bounds), C.c_int, C.c_int, i32p, i32p, C.c_int, i32p,
                                         C.c_int, f64p, f64p, f64p, C.c_int, C.POINTER(Status)]
    L.ctsm_b200_dgtsv_batch.argtypes = [vp, C.POINTER(Bounds), C.c_int, i32p, C.c_int, i32p,
                                        f64p, f64p, f64p, f64p, f64p, C.c_int, C.POINTER(Status)]
    L.ctsm_b200_soilwater.argtypes = [vp, C.POINTER(Bounds), C.c_int, i32p,
                                      C.POINTER(STRUCTS["soilwater"]), C.c_int, C.POINTER(Status)]
    L.ctsm_b200_soiltemperature.argtypes = [vp, C.POINTER(Bounds), C.c_int, i32p, C.c_int, i32p,
                                            C.POINTER(STRUCTS["soiltemperature"]), C.c_int, C.POINTER(Status)]
    L.ctsm_b200_canopyfluxes.argtypes = [vp, C.POINTER(Bounds), C.c_int, i32p,
                                         C.POINTER(STRUCTS["canopyfluxes"]), C.c_int, C.POINTER(Status)]
    L.ctsm_b200_set_exposedvegp_filter.argtypes = [vp, C.POINTER(Bounds), C.c_int, i32p, i32p, i32p, i32p, i32p, i32p,
                                                   C.c_int]
    L.ctsm_b200_vert_tran_sink_hydstress.argtypes = [vp, C.POINTER(Bounds), C.c_int, i32p,
                                                     C.POINTER(STRUCTS["plantsink"]), C.c_int, C.POINTER(Status)]
    L.ctsm_b200_biogeophys_pre_flux_calcs.argtypes = [vp, C.POINTER(Bounds), C.c_int, i32p, C.c_int, i32p, C.c_int, i32p, C.c_int,
                                                      C.POINTER(STRUCTS["preflux"]), C.c_int, C.POINTER(Status)]
    L.ctsm_b200_calculate_surface_humidity.argtypes = [vp, C.POINTER(Bounds), C.c_int, i32p,
                                                       C.POINTER(STRUCTS["surfacehumidity"]), C.c_int, C.POINTER(Status)]
    L.ctsm_b200_bare_ground_fluxes.argtypes = [vp, C.POINTER(Bounds), C.c_int, i32p,
                                               C.POINTER(STRUCTS["baregroundfluxes"]), C.c_int, C.POINTER(Status)]
    L.ctsm_b200_hydrology_infiltration.argtypes = [vp, C.POINTER(Bounds), C.c_int, i32p, C.c_int, i32p, C.c_int, i32p,
                                                   C.POINTER(STRUCTS["infiltration"]), C.c_int, C.POINTER(Status)]
    L.ctsm_b200_calc_ozone_uptake.argtypes = [vp, C.POINTER(Bounds), C.c_int, i32p, C.POINTER(STRUCTS["ozone"]), C.c_int, C.POINTER(Status)]
    L.ctsm_b200_calc_ozone_stress.argtypes = [vp, C.POINTER(Bounds), C.c_int, i32p, C.c_int, i32p, C.c_int, C.c_int, C.POINTER(STRUCTS["ozone"]),
                                              C.c_int, C.POINTER(Status)]
    L.ctsm_b200_build_snow_filter.argtypes = [vp, C.POINTER(Bounds), C.c_int, i32p, i32p, C.c_int, C.c_int, i32p, i32p, i32p, i32p, C.c_int]
    L.ctsm_b200_snow_water.argtypes = [vp, C.POINTER(Bounds), C.c_int, i32p, C.c_int, i32p, C.POINTER(STRUCTS["snowwater"]), C.c_int,
                                       C.POINTER(Status)]
    L.ctsm_b200_snow_capping.argtypes = [vp, C.POINTER(Bounds), C.c_int, i32p, C.c_int, i32p, C.POINTER(STRUCTS["snowcapping"]), C.c_int, C.c_int,
                                         C.POINTER(Status)]
    L.ctsm_b200_snow_layers.argtypes = [vp, C.POINTER(Bounds), C.c_int, i32p, C.POINTER(STRUCTS["snowlayers"]), C.c_int, C.POINTER(Status)]
    L.ctsm_b200_water_table.argtypes = [vp, C.POINTER(Bounds), C.c_int, i32p, C.c_int, i32p, C.POINTER(STRUCTS["watertable"]), C.c_int,
                                        C.POINTER(Status)]
    L.ctsm_b200_hydrology_diagnostics.argtypes = [vp, C.POINTER(Bounds), C.c_int, i32p, C.c_int, i32p, C.c_int, i32p, C.c_int, i32p, C.c_int,
                                                  i32p, C.POINTER(STRUCTS["hydrodiag"]), C.c_int, C.POINTER(Status)]
    L.ctsm_b200_vert_tran_sink_default.argtypes = [vp, C.POINTER(Bounds), C.c_int, i32p,
                                                     C.POINTER(STRUCTS["plantsinkdefault"]), C.c_int, C.POINTER(Status)]
    L.ctsm_b200_soilfluxes.argtypes = [vp, C.POINTER(Bounds), C.c_int, i32p, C.c_int, i32p,
                                       C.POINTER(STRUCTS["soilfluxes"]), C.c_int, C.POINTER(Status)]
    L.ctsm_b200_begin_water_column_balance.argtypes = [vp, C.POINTER(Bounds), C.c_int, i32p, C.c_int, i32p,
                                                       C.POINTER(STRUCTS["waterbalance"]), C.c_double, C.c_int, C.POINTER(Status)]
    L.ctsm_b200_water_gridcell_balance.argtypes = [vp, C.POINTER(Bounds), C.c_int, i32p, C.c_int, i32p,
                                                   C.POINTER(STRUCTS["watergridbalance"]), C.c_double, C.c_int, C.c_int,
                                                   C.POINTER(Status)]
    L.ctsm_b200_water_gridcell_balance.restype = C.c_int
    L.ctsm_b200_begin_water_column_balance.restype = C.c_int
    L.ctsm_b200_set_filters.argtypes = [vp, C.POINTER(Bounds), C.POINTER(FilterInputs), C.POINTER(Filters), C.c_int]
    L.ctsm_b200_set_filters.restype = C.c_int
    L.ctsm_b200_patch2col.argtypes = [vp, C.POINTER(Bounds), C.c_int, i32p, C.c_int, i32p,
                                      C.POINTER(STRUCTS["patch2col"]), C.c_int, C.POINTER(Status)]
    L.ctsm_b200_balancecheck_init.argtypes = [vp]
    L.ctsm_b200_balancecheck.argtypes = [vp, C.POINTER(Bounds), C.c_int, i32p, C.POINTER(STRUCTS["balancecheck"]),
                                         C.c_int, C.c_int, C.POINTER(BalanceReport), C.POINTER(Status)]
    L.ctsm_b200_set_soil_tuning.argtypes = [vp, C.c_int]
    L.ctsm_b200_set_soilwater_tuning.argtypes = [vp, C.c_int]
    L.ctsm_b200_set_sink_tuning.argtypes = [vp, C.c_int]
    for fn in ("vert_tran_sink_hydstress", "vert_tran_sink_default", "biogeophys_pre_flux_calcs", "calculate_surface_humidity",
               "bare_ground_fluxes", "hydrology_infiltration", "calc_ozone_uptake", "calc_ozone_stress", "build_snow_filter", "snow_water", "snow_capping", "snow_layers", "water_table", "hydrology_diagnostics", "balancecheck_init", "balancecheck", "soilfluxes", "patch2col"):
        getattr(L, "ctsm_b200_" + fn).restype = C.c_int
    for fn in ("init", "finalize", "sync", "host_register", "host_unregister", "tridiagonal", "banddiagonal",
               "dgtsv_batch", "soilwater", "soiltemperature", "canopyfluxes", "set_exposedvegp_filter"):
        getattr(L, "ctsm_b200_" + fn).restype = C.c_int
    _lib = L
    return L


def i32p(a):
    return _ptr(a, C.c_int32)


def f64p(a):
    return _ptr(a, C.c_double)
