"""Host-side logic of the drop-in layer that needs no GPU: the byte audits bench.py reports (algorithmic bytes per routine,
PCIe bytes of a CTSM_MEM_HOST step) follow from the field table, the step order is clm_drv's, and every routine of the
step has a field group, a filter list and a C-ABI symbol."""
import ctypes as C

import numpy as np

from ctsm_b200 import abi, driver, synthetic_canopy


def _case():
    sg, S = synthetic_canopy.make_full_case(64, seed=3)
    synthetic_canopy.balance_state(sg, S, np.random.Generator(np.random.PCG64(4)), 1.0e-11)
    synthetic_canopy.soilfluxes_state(sg, S, np.random.Generator(np.random.PCG64(5)))
    return sg, S


def test_step_order_is_clm_drv_order_and_every_routine_is_bound():
    assert driver.ROUTINES == ("canopyfluxes", "soiltemperature", "soilfluxes", "patch2col", "plantsink", "soilwater",
                               "balancecheck")                     # clm_driver.F90:766, 900, 921, 936, 950 (339, 346), 1422
    L = abi.lib()
    symbol = {"plantsink": "ctsm_b200_vert_tran_sink_hydstress"}
    for g in driver.ROUTINES:
        assert g in abi.FIELDS and g in driver.FILTER_OF
        assert hasattr(L, symbol.get(g, "ctsm_b200_" + g))
    sg, S = _case()
    for g in driver.ROUTINES:
        abi.make_struct(g, S, sg.bounds)                           # every field of every group exists with the right shape


def test_pre_flux_routines_are_bound_in_clm_drv_order():
    assert driver.PRE_ROUTINES == ("preflux", "surfacehumidity", "baregroundfluxes")        # clm_driver.F90:680, 702, 711
    L = abi.lib()
    symbol = {"preflux": "ctsm_b200_biogeophys_pre_flux_calcs", "surfacehumidity": "ctsm_b200_calculate_surface_humidity",
              "baregroundfluxes": "ctsm_b200_bare_ground_fluxes"}
    sg, S = _case()
    synthetic_canopy.preflux_state(sg, S, np.random.Generator(np.random.PCG64(6)))
    for g in driver.PRE_ROUTINES:
        assert g in abi.FIELDS and g in driver.FILTER_OF and hasattr(L, symbol[g])
        abi.make_struct(g, S, sg.bounds)
        for k in driver.FILTER_OF[g]:
            assert k in sg.filters


def test_algorithmic_bytes_follow_the_field_table():
    sg, S = _case()
    for g in driver.ROUTINES:
        ab = driver.algorithmic_bytes(sg, S, g)
        assert ab["bytes"] > 0 and ab["columns"] > 0
        # recompute from the table for the simplest group: 1-D fields only count once per unit (twice for INOUT)
    g = "patch2col"
    ab = driver.algorithmic_bytes(sg, S, g)
    want = 0.0
    for fs in abi.FIELDS[g]:
        es = 8 if fs.ctype == "double" else 4
        units = ab["patches"] if fs.sub == "PATCH" else ab["columns"]
        want += units * fs.used_soil * es * (2 if fs.intent == "INOUT" else 1)
    assert abs(ab["bytes"] - want) < 1e-6
    # CanopyFluxes counts each column once although ~7 exposed patches share it
    ac = driver.algorithmic_bytes(sg, S, "canopyfluxes")
    assert ac["patches"] == len(sg.filters["exposedvegp"]) and ac["columns"] <= ac["patches"]
    assert 2000 < ac["bytes_per_patch"] < 6000


def test_staged_bytes_count_every_field_once_per_direction():
    sg, S = _case()
    h2d, d2h = driver.staged_bytes(sg, ("soilwater",), preserve_out=True)
    up = down = 0
    for fs in abi.FIELDS["soilwater"]:
        n = sg.bounds.extent(fs.sub) * fs.nlev * (8 if fs.ctype == "double" else 4)
        up += n
        down += n if fs.intent in ("OUT", "INOUT") else 0
    assert h2d == up + 4 * len(sg.filters["hydrologyc"]) and d2h == down
    h2d_min, _ = driver.staged_bytes(sg, ("soilwater",), preserve_out=False)
    assert h2d_min < h2d
    tot = driver.staged_bytes(sg, driver.ROUTINES)
    assert tot[0] == sum(driver.staged_bytes(sg, (g,))[0] for g in driver.ROUTINES)


def test_params_roundtrip_and_ensemble_table_length():
    prm = abi.default_params()
    assert prm.npft_table == abi.MXPFT + 1 and prm.use_hydrstress == 1 and prm.itmax_canopy_fluxes == 40
    assert C.sizeof(abi.Params) % 8 == 0
