#!/usr/bin/env python
"""tests/golden/make_oracle_golden.py: writes tests/golden/oracle_step_tiny.npz, a regression fixture of the ORACLE (not of
the reference: CTSM cannot be built here).  It freezes the outputs of the oracle's seven-routine step
(CanopyFluxes -> SoilTemperature -> SoilFluxes -> clm_drv_patch2col -> root-water sink -> SoilWater -> BalanceCheck) on
the 64-gridcell synthetic case, so that an accidental edit of oracle/*.c shows up as a diff in the CPU suite
(tests/test_oracle_regression.py).  Regenerate only when the oracle is changed on purpose:
    python tests/golden/make_oracle_golden.py"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ctsm_b200 import abi, synthetic_canopy  # noqa: E402
from oracle import oracle  # noqa: E402

FIELDS = ("t_veg", "num_iter", "taf", "qaf", "eflx_sh_veg", "qflx_evap_veg", "qflx_tran_veg", "fpsn", "btran", "bsun", "bsha",
          "vegwp", "rssun", "rssha", "t_soisno", "h2osoi_liq", "h2osoi_ice", "t_grnd", "imelt", "eflx_fgr", "smp_l", "hk_l",
          "num_substeps", "qflx_rootsoi", "eflx_soil_grnd", "eflx_lwrad_out", "errsoi_col", "qflx_evap_soi_col", "errh2o", "errseb")
GROUPS = ("soiltemperature", "soilwater", "canopyfluxes", "plantsink", "balancecheck", "soilfluxes", "patch2col")


def run():
    OL = oracle.lib()
    prm = abi.default_params()
    prm.balance_skip_steps = int(OL.oracle_balancecheck_skip_steps(prm.dtime))
    sg, S = synthetic_canopy.make_full_case("tiny", seed=20260101)
    synthetic_canopy.balance_state(sg, S, np.random.Generator(np.random.PCG64(20260102)), 1.0e-11)
    synthetic_canopy.soilfluxes_state(sg, S, np.random.Generator(np.random.PCG64(20260103)))
    clumps, keep = oracle.make_clumps(sg, 3)
    structs = [abi.make_struct(g, S, sg.bounds) for g in GROUPS]
    rc = OL.oracle_fullstep_clumps(C.byref(prm), len(clumps), clumps, *[C.byref(x) for x in structs], 1, 127)
    assert rc == 0, rc
    return {k: S[k].copy() for k in FIELDS}


if __name__ == "__main__":
    out = run()
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_step_tiny.npz"), **out)
    print("wrote oracle_step_tiny.npz:", {k: v.shape for k, v in out.items()})
