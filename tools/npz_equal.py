#!/usr/bin/env python
import sys, numpy as np
a, b = np.load(sys.argv[1]), np.load(sys.argv[2])
bad = [k for k in a.files if not np.array_equal(a[k], b[k], equal_nan=True)]
for k in bad:
    d = np.abs(a[k].astype(float) - b[k].astype(float)); m = np.isfinite(d)
    print("DIFF", k, "n=", int((a[k] != b[k]).sum()), "max abs", float(d[m].max()) if m.any() else None)
print("identical" if not bad else "%d fields differ" % len(bad))
