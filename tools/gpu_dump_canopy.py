#!/usr/bin/env python
"""tools/gpu_dump_canopy.py SIZE SEED OUT.npz: CanopyFluxes on cuda:0 (device-resident) with the library selected by
CTSM_B200_LIB; dumps every OUT/INOUT field.  Experiment aid: bitwise A/B comparison of library variants."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ctsm_b200 import abi, synthetic_canopy
from tests.util import to_device, group_arrays
size, seed, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
L = abi.lib(); prm = abi.default_params(); ctx = C.c_void_p()
assert L.ctsm_b200_init(C.byref(prm), C.byref(ctx)) == 0
sg, S = synthetic_canopy.make_full_case(size, seed=seed)
D = to_device(group_arrays(S, "canopyfluxes"))
fe = sg.filters["exposedvegp"]; dfe = to_device({"f": fe})["f"]
f = abi.make_struct("canopyfluxes", D, sg.bounds); st = abi.Status()
assert L.ctsm_b200_canopyfluxes(ctx, C.byref(sg.bounds), len(fe), abi.i32p(dfe), C.byref(f), abi.MEM_DEVICE, C.byref(st)) == 0
assert L.ctsm_b200_sync(ctx, C.byref(st)) == 0
np.savez(out, **{fs.name: D[fs.name].cpu().numpy() for fs in abi.FIELDS["canopyfluxes"] if fs.intent != "IN"})
print("dumped", out, "warnings", st.n_warnings)
