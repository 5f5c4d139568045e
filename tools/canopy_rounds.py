#!/usr/bin/env python
"""Per-round statistics and timing of one CanopyFluxes call on the GPU, for tuning the bulk/tail schedule
(ctsm_b200_set_tuning).  usage: canopy_rounds.py SIZE [tail_max nt_budget tail_lanes]..."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctsm_b200 import abi, driver, synthetic_canopy

size = sys.argv[1] if len(sys.argv) > 1 else "f09"
size = int(size) if size.isdigit() else size
confs = [tuple(int(x) for x in a.split(",")) for a in sys.argv[2:]] or [(0, 0, 1, 1)]
sg, S = synthetic_canopy.make_full_case(size, seed=20260101)
ctx = driver.Context(abi.default_params())
names = [fs.name for fs in abi.FIELDS["canopyfluxes"]]
D = {k: torch.from_numpy(S[k]).cuda() for k in names}
pristine = {fs.name: D[fs.name].clone() for fs in abi.FIELDS["canopyfluxes"] if fs.intent != "IN"}
hp = driver.HotPath(ctx, sg, D, abi.MEM_DEVICE, ("canopyfluxes",))
stream = torch.cuda.ExternalStream(ctx.stream_ptr)
for conf in confs:
    ctx.set_tuning(*conf)
    ts = []
    for it in range(4):
        with torch.cuda.stream(stream):
            for k, v in pristine.items():
                D[k].copy_(v)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); l0 = ctx.launches
        hp.CanopyFluxes()
        e1.record(stream); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1)); nl = ctx.launches - l0
    ll, te = np.zeros(64, np.int32), np.zeros(64, np.int32)
    n = ctx.L.ctsm_b200_canopy_round_stats(ctx.h, abi.i32p(ll), abi.i32p(te), 64)
    print("conf %s (tail_max, nt_budget, lanes, nt_split): ms %s launches %d" % (conf, ["%.2f" % t for t in ts], nl))
    print("  list_len", ll[:n].tolist())
    print("  tail_new", np.diff(np.concatenate([[0], te[:n]])).tolist())
ctx.close()
