"""The C-ABI library must load (no GPU needed for that) and export every symbol
include/ctsm_b200.h declares.  No compute call is made here."""
import ctypes as C
import os
import re

from ctsm_b200 import abi


def _declared_symbols():
    hdr = open(os.path.join(abi.ROOT, "include", "ctsm_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(ctsm_b200_\w+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    L = abi.lib()
    syms = _declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(L, s), "libctsm_b200.so does not export %s" % s


def test_version_and_defaults_match_python_twin():
    L = abi.lib()
    assert b"sm_100a" in L.ctsm_b200_version()
    p = abi.Params()
    L.ctsm_b200_default_params(C.byref(p))
    q = abi.default_params()
    for name, _ in abi.Params._fields_:
        if name.startswith("reserved"):
            continue
        assert getattr(p, name) == getattr(q, name), name
    # layout check of the ctypes mirror: first, middle and last members as ctsm_b200_default_params wrote them
    assert (p.abi_version, p.nlevsno, p.nlevgrnd, p.nlevsoi) == (5, 12, 25, 20)
    assert (p.dtmin, p.xTolerUpper, p.snow_thermal_cond_method) == (60.0, 0.1, 2)
    assert (p.itmax_canopy_fluxes, p.z0param_method, p.stomatalcond_mtd) == (40, 2, 2)
    assert (p.csoilc, p.zetamaxstable, p.lmrhd, p.jmax25top_sf, p.balance_skip_steps) == (0.004, 2.0, 150650.0, 1.0, -1)
    assert all(v == 0 for v in p.reserved_i) and (p.fff, p.pc, p.mu, p.h2osfcflag) == (0.5, 0.4, 0.13889, 1)


def test_struct_layout_matches_def_table():
    # one pointer per CTSM_F line + the alloc bounds
    for g, specs in abi.FIELDS.items():
        st = abi.STRUCTS[g]
        assert C.sizeof(st) == C.sizeof(abi.Bounds) + 8 * len(specs)
        names = [n for n, _ in st._fields_][1:]
        assert names == [fs.name for fs in specs]


def test_no_device_is_loud():
    """Without a GPU the init call must fail with CTSM_ERR_NO_DEVICE, never fall back."""
    import torch
    if torch.cuda.is_available():
        return
    L = abi.lib()
    p = abi.default_params()
    ctx = C.c_void_p()
    assert L.ctsm_b200_init(C.byref(p), C.byref(ctx)) == 1
    assert not ctx.value


def test_ctypes_mirrors_have_the_c_layout(tmp_path):
    """sizeof / offsetof of every hand-written ctypes mirror against the C compiler's view of include/ctsm_b200.h."""
    import subprocess
    probes = {"ctsm_bounds_t": (abi.Bounds, ["begg", "endp", "clump_index"]),
              "ctsm_status_t": (abi.Status, ["code", "value", "n_warnings", "msg"]),
              "ctsm_params_t": (abi.Params, ["abi_version", "dtime", "e_ice", "itmax_canopy_fluxes", "lai_dl", "jmax25top_sf",
                                             "balance_skip_steps", "npft_table", "calc_human_stress_indices", "reserved_i", "fff", "mu", "snicar_use_aerosol", "snow_dzmin_1", "scvng_fct_mlt_dst4"]),
              "ctsm_balance_report_t": (abi.BalanceReport, ["max_abs", "index", "warn", "abort_kind", "skip_steps"]),
              "ctsm_filter_inputs_t": (abi.FilterInputs, ["alloc", "col_active", "melt_replaced_by_ice_grc", "include_inactive",
                                                          "npcropmax"]),
              "ctsm_filters_t": (abi.Filters, ["list", "num"])}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "ctsm_b200.h"', "int main(void) {"]
    for cname, (_, members) in probes.items():
        lines.append('  printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for m in members:
            lines.append('  printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, m, cname, m))
    lines += ["  return 0;", "}"]
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = str(tmp_path / "probe")
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(abi.ROOT, "include"), str(src), "-o", exe])
    out = dict(l.split() for l in subprocess.run([exe], capture_output=True, text=True).stdout.splitlines())
    for cname, (ct, members) in probes.items():
        assert C.sizeof(ct) == int(out[cname]), cname
        for m in members:
            assert getattr(ct, m).offset == int(out["%s.%s" % (cname, m)]), (cname, m)
    assert len(abi.FILTER_NAMES) == C.sizeof(abi.Filters().num) // 4
