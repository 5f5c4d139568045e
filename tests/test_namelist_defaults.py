"""ctsm_b200_default_params pinned against the reference's clm6_0 namelist defaults
(tests/golden/namelist_defaults_clm6_0.json, generated from bld/namelist_files/namelist_defaults_ctsm.xml by
tests/golden/make_namelist_golden.py).  The oracle and the CUDA path share default_params, so a wrong default
is invisible to every parity test; this is the test that sees it."""
import ctypes as C
import json
import os

from ctsm_b200 import abi

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "namelist_defaults_clm6_0.json")))["values"]


def _logical(s):
    return {".true.": 1, ".false.": 0}[s]


def _real(s):
    return float(s.lower().replace("d", "e"))


def test_default_params_follow_clm6_0_namelist_defaults():
    L = abi.lib()
    p = abi.Params()
    L.ctsm_b200_default_params(C.byref(p))
    snow = {"Jordan1991": 1, "Sturm1997": 2}
    assert p.snow_thermal_cond_method == snow[GOLD["snow_thermal_cond_method"]]
    assert p.snow_thermal_cond_glc_method == snow[GOLD["snow_thermal_cond_glc_method"]]
    assert p.upper_boundary_condition == int(GOLD["upper_boundary_condition"])
    assert p.lower_boundary_condition == int(GOLD["lower_boundary_condition"])
    assert p.flux_calculation == int(GOLD["flux_calculation"])
    assert p.dtmin == _real(GOLD["dtmin"]) and p.verySmall == _real(GOLD["verySmall"])
    assert p.xTolerUpper == _real(GOLD["xTolerUpper"]) and p.xTolerLower == _real(GOLD["xTolerLower"])
    assert p.itmax_canopy_fluxes == int(GOLD["itmax_canopy_fluxes"])
    assert p.use_undercanopy_stability == _logical(GOLD["use_undercanopy_stability"])
    assert p.use_biomass_heat_storage == _logical(GOLD["use_biomass_heat_storage"])
    assert p.z0param_method == {"ZengWang2007": 1, "Meier2022": 2}[GOLD["z0param_method"]]
    assert p.soil_resis_method == int(GOLD["soil_resis_method"])
    assert p.use_hydrstress == _logical(GOLD["use_hydrstress"])
    assert p.use_luna == _logical(GOLD["use_luna"])
    assert p.stomatalcond_mtd == {"Ball-Berry1987": 1, "Medlyn2011": 2}[GOLD["stomatalcond_method"]]
    assert p.light_inhibit == _logical(GOLD["light_inhibit"])
    assert p.modifyphoto_and_lmr_forcrop == _logical(GOLD["modifyphoto_and_lmr_forcrop"])
    assert p.zetamaxstable == _real(GOLD["zetamaxstable"])
    assert p.leaf_mr_vcm == _real(GOLD["leaf_mr_vcm"])
    assert p.calc_human_stress_indices == {"NONE": 0, "FAST": 1, "ALL": 2}[GOLD["calc_human_stress_indices"]]
    assert p.nlevsno == int(GOLD["nlevsno"])
    assert GOLD["soil_layerstruct_predefined"] == "20SL_8.5m" and (p.nlevsoi, p.nlevgrnd) == (20, 25)
    assert int(GOLD["soilwater_movement_method"]) == 1      # moisture form + adaptive time stepping: the only one built


def test_snow_defaults_follow_clm6_0_namelist_defaults():
    L = abi.lib()
    p = abi.Params()
    L.ctsm_b200_default_params(C.byref(p))
    for nm in ("snow_dzmin_1", "snow_dzmin_2", "snow_dzmax_l_1", "snow_dzmax_l_2", "snow_dzmax_u_1", "snow_dzmax_u_2", "int_snow_max"):
        assert getattr(p, nm) == _real(GOLD[nm]), nm
    assert p.overburden_compress_Tfactor == _real(GOLD["overburden_compress_tfactor"])
    assert p.wind_dependent_snow_density == _logical(GOLD["wind_dependent_snow_density"])
    assert p.use_subgrid_fluxes == _logical(GOLD["use_subgrid_fluxes"])
    assert p.snicar_use_aerosol == _logical(GOLD["snicar_use_aerosol"])
    assert p.snow_overburden_compaction_method == {"Anderson1976": 1, "Vionnet2012": 2}[GOLD["snow_overburden_compaction_method"].strip("'")]
    assert GOLD["snow_cover_fraction_method"] == "SwensonLawrence2012"      # FracSnowDuringMelt: the only method built
