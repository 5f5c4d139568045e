"""Pins the C oracle (oracle/*.c, the checker of every GPU parity test) against a SECOND restatement of the same
Fortran that shares nothing with it: NumPy whole-array code + LAPACK's own dgtsv / dgbsv through scipy
(oracle/independent/).  Both were written from the reference source; agreement at <= 1e-13 over thousands of random
columns, including every edge regime, is what stands in for a run of the reference (no Fortran compiler exists in the
build container or on the GPU box: profiles/r02_fortran_probe.txt)."""
import ctypes as C

import numpy as np
import pytest

from ctsm_b200 import abi, synthetic
from tests.util import relerr, copy_state

RTOL = 1e-13


@pytest.mark.parametrize("size,seed,lbc", [(3000, 20260103, 2), (800, 7, 1), ("tiny", 3, 2)])
def test_soilwater_c_oracle_matches_numpy_restatement(oracle_lib, size, seed, lbc):
    from oracle.independent import soilwater_np
    prm = abi.default_params()
    prm.lower_boundary_condition = lbc
    sg, S = synthetic.make_case(size, seed=seed)
    a, b = copy_state(S), copy_state(S)
    fh = sg.filters["hydrologyc"]
    st = abi.Status()
    f = abi.make_struct("soilwater", a, sg.bounds)
    assert oracle_lib.oracle_soilwater(C.byref(prm), C.byref(sg.bounds), len(fh), abi.i32p(fh), C.byref(f), C.byref(st)) == 0
    soilwater_np.soilwater(b, fh, dtime=prm.dtime, dtmin=prm.dtmin, very_small=prm.verySmall, x_toler_upper=prm.xTolerUpper,
                           x_toler_lower=prm.xTolerLower, e_ice=prm.e_ice, lower_boundary_condition=lbc)
    assert np.array_equal(a["num_substeps"], b["num_substeps"])
    assert (a["num_substeps"][fh - 1] > 1).any() and a["num_substeps"][fh - 1].max() >= 4
    for name in ("h2osoi_liq", "smp_l", "hk_l", "qin", "qout", "qcharge"):
        e = relerr(a[name], b[name])
        assert e <= RTOL, (name, e)
    # columns outside the filter untouched by both
    out = np.setdiff1d(np.arange(sg.ncol), fh - 1)
    assert np.array_equal(b["h2osoi_liq"][:, out], S["h2osoi_liq"][:, out])


@pytest.mark.parametrize("size,seed,snow,glc", [(3000, 20260102, 2, 2), (1500, 5, 1, 1), ("tiny", 3, 2, 1)])
def test_soiltemperature_c_oracle_matches_numpy_restatement(oracle_lib, size, seed, snow, glc):
    from oracle.independent import soiltemp_np
    prm = abi.default_params()
    prm.snow_thermal_cond_method, prm.snow_thermal_cond_glc_method = snow, glc
    sg, S = synthetic.make_case(size, seed=seed)
    a, b = copy_state(S), copy_state(S)
    fc, fp = sg.filters["nolakec"], sg.filters["nolakep"]
    st = abi.Status()
    f = abi.make_struct("soiltemperature", a, sg.bounds)
    assert oracle_lib.oracle_soiltemperature(C.byref(prm), C.byref(sg.bounds), len(fp), abi.i32p(fp), len(fc), abi.i32p(fc),
                                             C.byref(f), C.byref(st)) == 0
    soiltemp_np.soiltemperature(b, fp, fc, dtime=prm.dtime, snow_method=snow, snow_glc_method=glc)
    worst = {}
    for fs in abi.FIELDS["soiltemperature"]:
        if fs.intent == "IN":
            assert np.array_equal(b[fs.name], S[fs.name]), fs.name
            continue
        if fs.ctype == "int":
            assert np.array_equal(a[fs.name], b[fs.name]), fs.name          # imelt: identical melt / freeze decisions
            continue
        x, y = a[fs.name], b[fs.name]
        fin = np.abs(x) < 1e30
        assert np.array_equal(fin, np.abs(y) < 1e30), fs.name
        if not fin.any():
            continue
        # The two codes share no arithmetic below the formula level (the band solve is LAPACK's FMA-using dgbsv here,
        # a restated non-FMA LU in the C oracle), so temperatures differ in the last bits (~1e-15 relative) and the
        # phase-change terms, which are differences against the freezing point, inherit that as an ABSOLUTE error:
        # the bar is 1e-13 of the field's scale plus 1e-11 relative, three orders below the 1e-10 parity tolerance.
        scale = float(np.max(np.abs(x[fin])))
        err = np.abs(x[fin] - y[fin]) / (1e-13 * scale + 1e-11 * np.abs(x[fin]) + 1e-300)
        worst[fs.name] = float(err.max())
    bad = {k: v for k, v in worst.items() if not v <= 1.0}
    assert not bad, bad
    if size != "tiny":      # every regime present: melting, freezing, all snow depths, standing water, unlayered snow
        assert set(np.unique(a["imelt"])) >= {0, 1, 2}
        assert (S["snl"] == 0).any() and (S["snl"] == -12).any() and (S["frac_h2osfc"] > 0).any()
        assert (a["xmf_h2osfc"] != 0).any() and (S["h2osno_no_layers"] > 0).any()
