#!/usr/bin/env python
"""Writes tests/golden/namelist_defaults_clm6_0.json: the clm6_0 namelist defaults of every switch that
ctsm_params_t mirrors, resolved from the reference's own defaults file

    bld/namelist_files/namelist_defaults_ctsm.xml

the way bld/CLMBuildNamelist.pm resolves them: among the entries of a variable whose attributes all
match the configuration, the one with the most matching attributes wins.  Configuration = the default
CTSM case of the hot path: phys=clm6_0, structure=standard, use_fates=.false., configuration=clm,
use_bedrock=.true., soilwater_movement_method=1, vichydro=0.

Run in the build container (needs /root/reference); the JSON is committed so the test that pins
ctsm_b200_default_params against it (tests/test_namelist_defaults.py) runs anywhere.
"""
import json
import os
import re
import sys

REF = os.environ.get("CTSM_REFERENCE", "/root/reference")
XML = os.path.join(REF, "bld", "namelist_files", "namelist_defaults_ctsm.xml")

CONFIG = {"phys": "clm6_0", "structure": "standard", "use_fates": ".false.", "configuration": "clm",
          "use_bedrock": ".true.", "soilwater_movement_method": "1", "vichydro": "0", "use_cn": ".false.",
          "use_biomass_heat_storage": ".true.", "z0param_method": "Meier2022"}
VARS = ["upper_boundary_condition", "lower_boundary_condition", "flux_calculation", "dtmin", "verySmall", "xTolerUpper",
        "xTolerLower", "snow_thermal_cond_method", "snow_thermal_cond_glc_method", "itmax_canopy_fluxes",
        "use_undercanopy_stability", "use_biomass_heat_storage", "z0param_method", "soil_resis_method", "use_hydrstress",
        "use_luna", "stomatalcond_method", "light_inhibit", "modifyphoto_and_lmr_forcrop", "zetamaxstable", "leaf_mr_vcm",
        "soilwater_movement_method", "nlevsno", "soil_layerstruct_predefined", "calc_human_stress_indices",
        "snow_dzmin_1", "snow_dzmin_2", "snow_dzmax_l_1", "snow_dzmax_l_2", "snow_dzmax_u_1", "snow_dzmax_u_2", "int_snow_max",
        "wind_dependent_snow_density", "snow_overburden_compaction_method", "overburden_compress_tfactor", "use_subgrid_fluxes",
        "snow_cover_fraction_method", "snicar_use_aerosol"]


def resolve(text, var):
    best, best_n, best_line = None, -1, None
    for m in re.finditer(r"<%s(\s[^>]*)?>([^<]*)</%s>" % (var, var), text):
        attrs = dict(re.findall(r'(\w+)\s*=\s*"([^"]*)"', m.group(1) or ""))
        if all(CONFIG.get(k) == v for k, v in attrs.items()):
            if len(attrs) > best_n:
                best, best_n = m.group(2).strip(), len(attrs)
                best_line = text.count("\n", 0, m.start()) + 1
    return best, best_line


def main():
    text = open(XML).read()
    out = {"_source": "bld/namelist_files/namelist_defaults_ctsm.xml", "_config": CONFIG, "values": {}, "lines": {}}
    for v in VARS:
        val, line = resolve(text, v)
        if val is None:
            sys.exit("no default for %s" % v)
        out["values"][v] = val
        out["lines"][v] = line
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "namelist_defaults_clm6_0.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out["values"], indent=1))


if __name__ == "__main__":
    main()
