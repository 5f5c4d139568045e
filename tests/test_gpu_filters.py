"""GPU parity: ctsm_b200_set_filters (setFiltersOneGroup, filterMod.F90:303-592) vs the CPU oracle.  Integer index
lists: bit-exact, counts included; entries beyond a list's length must stay untouched."""
import ctypes as C

import numpy as np
import pytest

from ctsm_b200 import abi
from tests.filters_util import random_topology, make_inputs, make_outputs
from tests.util import to_device

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ng,mem,inc,cn,fates,fbgc", [(40, abi.MEM_HOST, 0, 0, 0, 0), (40, abi.MEM_DEVICE, 1, 1, 0, 1),
                                                    (3000, abi.MEM_DEVICE, 0, 1, 0, 0), (3000, abi.MEM_HOST, 0, 0, 1, 1),
                                                    (1, abi.MEM_DEVICE, 0, 1, 0, 0)])
def test_set_filters_bit_exact(gpu_ctx, oracle_lib, ng, mem, inc, cn, fates, fbgc):
    L, ctx, prm = gpu_ctx
    b, T = random_topology(ng, 300 + ng + inc)
    ref_in = make_inputs(b, T, inc, cn, fates, fbgc)
    ref, rbuf = make_outputs(b)
    assert oracle_lib.oracle_set_filters(C.byref(b), C.byref(ref_in), C.byref(ref)) == 0
    if mem == abi.MEM_DEVICE:
        D = to_device(T)
        fin = make_inputs(b, T, inc, cn, fates, fbgc, arrays=D)
        out, bufs = make_outputs(b, device=True)
    else:
        fin = make_inputs(b, T, inc, cn, fates, fbgc)
        out, bufs = make_outputs(b)
    assert L.ctsm_b200_set_filters(ctx, C.byref(b), C.byref(fin), C.byref(out), mem) == 0
    for k, name in enumerate(abi.FILTER_NAMES):
        got = bufs[k].cpu().numpy() if mem == abi.MEM_DEVICE else bufs[k]
        assert out.num[k] == ref.num[k], name
        assert np.array_equal(got[:out.num[k]], rbuf[k][:ref.num[k]]), name
        assert np.all(got[out.num[k]:] == -7), "%s: wrote past its length" % name
    assert sum(out.num) > 0


def test_set_filters_clump_bounds(gpu_ctx, oracle_lib):
    L, ctx, prm = gpu_ctx
    b, T = random_topology(200, 9)
    lg = T["col_gridcell"]
    cols = np.nonzero((lg >= 71) & (lg <= 150))[0] + 1
    luns = np.unique(T["col_landunit"][cols - 1])
    pats = np.nonzero(np.isin(T["patch_landunit"], luns))[0] + 1
    k = b.copy()
    k.begg, k.endg, k.begl, k.endl = 71, 150, int(luns[0]), int(luns[-1])
    k.begc, k.endc, k.begp, k.endp = int(cols[0]), int(cols[-1]), int(pats[0]), int(pats[-1])
    fin = make_inputs(b, T, 0, 1, 0, 0)
    ref, rbuf = make_outputs(k)
    assert oracle_lib.oracle_set_filters(C.byref(k), C.byref(fin), C.byref(ref)) == 0
    out, bufs = make_outputs(k)
    assert L.ctsm_b200_set_filters(ctx, C.byref(k), C.byref(fin), C.byref(out), abi.MEM_HOST) == 0
    for i, name in enumerate(abi.FILTER_NAMES):
        assert out.num[i] == ref.num[i], name
        assert np.array_equal(bufs[i][:out.num[i]], rbuf[i][:ref.num[i]]), name
