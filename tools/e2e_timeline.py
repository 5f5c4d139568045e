#!/usr/bin/env python
"""Timeline of the end-to-end (host arrays, resident window) step: when does the host return from each call, how long
does the whole window take, for several slab counts.  usage: e2e_timeline.py SIZE K [K ...]"""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctsm_b200 import abi, driver
import bench

size = sys.argv[1] if len(sys.argv) > 1 else "f02"
size = int(size) if size.isdigit() else size
ks = [int(x) for x in sys.argv[2:]] or [1, 4]
sg, S = bench.make_workload(size, 20260101)
ctx = driver.Context(abi.default_params())
names = sorted({fs.name for g in driver.ROUTINES for fs in abi.FIELDS[g]})
restore = sorted({fs.name for g in driver.ROUTINES for fs in abi.FIELDS[g] if fs.intent != "IN"})
H = {k: torch.from_numpy(S[k]).pin_memory() for k in names}
Hn = {k: v.numpy() for k, v in H.items()}
for K in ks:
    hp = driver.HotPath(ctx, sg, Hn, abi.MEM_HOST, driver.ROUTINES, nslab=K, window=True)
    for it in range(3):
        for k in restore:
            Hn[k][...] = S[k]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        assert ctx.L.ctsm_b200_host_window_begin(ctx.h) == 0
        marks = []
        streams = [torch.cuda.ExternalStream(int(ctx.L.ctsm_b200_stream_of(ctx.h, kind))) for kind in (2, 0, 3)]
        ev0 = torch.cuda.Event(enable_timing=True); ev0.record(streams[1])
        devs = []
        for sl in range(len(hp.slabs)):
            hp._select(sl)
            for g in hp.routines:
                hp.call(g)
                marks.append((sl, g, time.perf_counter() - t0))
                if g in ("canopyfluxes", "balancecheck"):
                    evs = []
                    for stq in streams:
                        e = torch.cuda.Event(enable_timing=True); e.record(stq); evs.append(e)
                    devs.append((sl, g, evs))
        t_issue = time.perf_counter() - t0
        st = abi.Status()
        assert ctx.L.ctsm_b200_host_window_end(ctx.h, C.byref(st)) == 0
        t_all = time.perf_counter() - t0
        hp._select(0)
    print("K=%d: host issued everything after %.1f ms, window done after %.1f ms, bytes %s" % (K, 1e3 * t_issue, 1e3 * t_all, hp.window_bytes()))
    print("   host:", ["%d:%s %.1f" % (sl, g[:6], 1e3 * t) for sl, g, t in marks if g in ("canopyfluxes", "balancecheck")])
    print("   device (h2d, compute, d2h done, ms after window start):", ["%d:%s %.0f/%.0f/%.0f" % ((sl, g[:6]) + tuple(ev0.elapsed_time(e) for e in evs)) for sl, g, evs in devs])
    ctx.L.ctsm_b200_host_invalidate(ctx.h, None)
ctx.close()
