"""CPU checks of the setFiltersOneGroup oracle (filterMod.F90:303-592) against an independent numpy statement of the
lists, for every combination of the switches that change them, on whole and on clump-sized bounds."""
import ctypes as C
import itertools

import numpy as np
import pytest

from ctsm_b200 import abi
from oracle import oracle
from tests.filters_util import random_topology, expected_lists, make_inputs, make_outputs


@pytest.mark.parametrize("inc,cn,fates,fbgc", list(itertools.product((0, 1), (0, 1), (0, 1), (0, 1))))
def test_oracle_filters_match_numpy(inc, cn, fates, fbgc):
    OL = oracle.lib()
    b, T = random_topology(60, 100 + inc + 2 * cn + 4 * fates + 8 * fbgc)
    fin = make_inputs(b, T, inc, cn, fates, fbgc)
    out, bufs = make_outputs(b)
    assert OL.oracle_set_filters(C.byref(b), C.byref(fin), C.byref(out)) == 0
    E = expected_lists(b, T, inc, cn, fates, fbgc)
    for k, name in enumerate(abi.FILTER_NAMES):
        assert out.num[k] == len(E[name]), name
        assert np.array_equal(bufs[k][:out.num[k]], E[name]), name


def test_oracle_filters_on_clump_bounds():
    OL = oracle.lib()
    b, T = random_topology(80, 7)
    # a clump = gridcells 31..60: contiguous landunit / column / patch ranges (initGridCellsMod nests g > l > c > p)
    lg = T["col_gridcell"]
    cols = np.nonzero((lg >= 31) & (lg <= 60))[0] + 1
    luns = np.unique(T["col_landunit"][cols - 1])
    pats = np.nonzero(np.isin(T["patch_landunit"], luns))[0] + 1
    k = b.copy()
    k.begg, k.endg, k.begl, k.endl = 31, 60, int(luns[0]), int(luns[-1])
    k.begc, k.endc, k.begp, k.endp = int(cols[0]), int(cols[-1]), int(pats[0]), int(pats[-1])
    fin = make_inputs(b, T, 0, 1, 0, 0)
    out, bufs = make_outputs(k)
    assert OL.oracle_set_filters(C.byref(k), C.byref(fin), C.byref(out)) == 0
    E = expected_lists(k, T, 0, 1, 0, 0)
    for i, name in enumerate(abi.FILTER_NAMES):
        assert np.array_equal(bufs[i][:out.num[i]], E[name]), name
    # nolakec / nolakep / hydrologyc of the synthetic generator are these lists
    assert out.num[abi.FILTER_NAMES.index("allc")] == int((T["col_active"][cols - 1] != 0).sum())
