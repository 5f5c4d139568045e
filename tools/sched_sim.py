#!/usr/bin/env python
"""tools/sched_sim.py [NGRIDCELLS]: lane-occupancy model of the calcstress task kernel (phs_newton_kernel, canopy.cu).

Takes the distribution of Newton iterations per calcstress call from the oracle's counters on a synthetic case
(oracle_phs_newton_hist, night and day solves separately) and simulates one warp of the lane-refill scheduler for a long
queue of tasks: a task costs `begin` + iterations x 1 + `finish` warp-steps, every scheduler round costs the same whatever
the number of active lanes, and the policy is the kernel's: refill when >= REFILL_MIN lanes are idle (or nothing runs), run
the epilogue when >= FIN_MIN lanes wait (or nothing runs), else step the running lanes.  Prints the fraction of
lane-rounds doing useful work for several thresholds - the numbers behind REFILL_MIN = 8 / FIN_MIN = 16.
Experiment aid; test infrastructure only (it loads the oracle)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ctsm_b200 import abi, synthetic_canopy  # noqa: E402
from oracle import oracle  # noqa: E402

CB, CFD, CFN = 0.3, 0.35, 0.9        # cost of the task prologue, the day epilogue, the night epilogue (in Newton iterations)


def histogram(ngrc):
    OL = oracle.lib()
    sg, S = synthetic_canopy.make_full_case(ngrc, seed=13)
    prm = abi.default_params()
    hist = ((C.c_longlong * 64) * 2).in_dll(OL, "oracle_phs_newton_hist")
    f = abi.make_struct("canopyfluxes", S, sg.bounds)
    fe = sg.filters["exposedvegp"]
    st = abi.Status()
    assert OL.oracle_canopyfluxes(C.byref(prm), C.byref(sg.bounds), len(fe), abi.i32p(fe), C.byref(f), C.byref(st)) == 0
    return np.array([list(hist[0]), list(hist[1])], dtype=float)


def simulate(h, ntasks, refill_min, fin_min, rng):
    night = rng.random(ntasks) < 0.5
    it = np.where(night, rng.choice(64, size=ntasks, p=h[1] / h[1].sum()), rng.choice(64, size=ntasks, p=h[0] / h[0].sum()))
    useful = float((it + CB + np.where(night, CFN, CFD)).sum())
    pos, t = 0, 0.0
    rem = np.zeros(32, int); st = np.zeros(32, int); isn = np.zeros(32, bool)      # 0 idle, 1 run, 2 waiting for the epilogue
    while True:
        idle, run, fin = st == 0, st == 1, st == 2
        if pos < ntasks and (idle.sum() >= refill_min or run.sum() + fin.sum() == 0):
            for lane in np.nonzero(idle)[0]:
                if pos < ntasks:
                    rem[lane], isn[lane] = it[pos], night[pos]
                    st[lane] = 1 if rem[lane] > 0 else 2
                    pos += 1
            t += CB
            continue
        if run.sum() + fin.sum() == 0:
            break
        if fin.sum() > 0 and (fin.sum() + idle.sum() >= fin_min or run.sum() == 0):
            t += CFN if isn[fin].any() else CFD
            st[fin] = 0
            continue
        rem[run] -= 1
        st[run & (rem == 0)] = 2
        t += 1.0
    return useful / (32.0 * t)


if __name__ == "__main__":
    h = histogram(int(sys.argv[1]) if len(sys.argv) > 1 else 2000)
    for k, name in ((0, "day"), (1, "night")):
        tot = h[k].sum()
        print("%-5s solves: mean %.2f iterations; <=4: %.1f %%, >20: %.2f %%" % (
            name, (h[k] * np.arange(64)).sum() / tot, 100 * h[k][:5].sum() / tot, 100 * h[k][21:].sum() / tot))
    rng = np.random.default_rng(1)
    for r, f in ((2, 2), (4, 4), (8, 8), (8, 16), (16, 16), (12, 24)):
        print("REFILL_MIN %2d FIN_MIN %2d: useful lane-rounds %.1f %%" % (r, f, 100 * simulate(h, 32 * 400, r, f, rng)))
