"""integration/ctsm_b200_mod.F90 (the ISO_C_BINDING module a CTSM maintainer compiles with the host model) is generated
from the header and the field table; the committed copy must be current, declare one derived type per field group with
one c_ptr member per table entry, and bind every entry point the header declares."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fortran_shim_is_current_and_complete():
    from ctsm_b200 import abi
    gen = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_fortran_shim.py")], capture_output=True, text=True)
    assert gen.returncode == 0, gen.stderr
    committed = open(os.path.join(ROOT, "integration", "ctsm_b200_mod.F90")).read()
    assert gen.stdout == committed, "integration/ctsm_b200_mod.F90 is stale: rerun tools/gen_fortran_shim.py"
    for g, fields in abi.FIELDS.items():
        block = re.search(r"type, bind\(C\) :: ctsm_%s_fields_t(.*?)end type" % g, committed, re.S).group(1)
        assert block.count("type(c_ptr) ::") == len(fields), g
    hdr = open(os.path.join(ROOT, "include", "ctsm_b200.h")).read()
    for fn in set(re.findall(r"\b(ctsm_b200_\w+)\s*\(", hdr)):
        assert 'bind(C, name="%s")' % fn in committed, fn
    assert not [l for l in committed.splitlines() if len(l) > 132 and not l.lstrip().startswith("!")]
    # the per-group c_loc blocks: one assignment (or one GATHER note) per table entry
    fill = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_fortran_shim.py"), "--fill"], capture_output=True, text=True)
    assert fill.returncode == 0, fill.stderr
    inc = open(os.path.join(ROOT, "integration", "ctsm_b200_fill.inc")).read()
    assert fill.stdout == inc, "integration/ctsm_b200_fill.inc is stale: rerun tools/gen_fortran_shim.py --fill"
    for g, fields in abi.FIELDS.items():
        block = re.search(r"#ifdef CTSM_FILL_%s\n(.*?)#undef" % g.upper(), inc, re.S).group(1)
        n = len(re.findall(r"^  f%\w+ = c_loc", block, re.M)) + len(re.findall(r"^  ! GATHER", block, re.M))
        assert n == len(fields), g
