"""Compares the activation dumps of tools/dbg_merge.py runs: python tools/dbg_merge_cmp.py C2 0 5"""
import sys
import numpy as np
name, a, b = sys.argv[1:4]
A = np.load("/tmp/dbg_acts_%s_%s.npz" % (name, a)); B = np.load("/tmp/dbg_acts_%s_%s.npz" % (name, b))
for k in A.files:
    x, y = A[k].astype(np.float64), B[k].astype(np.float64)
    d = np.abs(x - y)
    print("%-12s rel-l2 %.2e  max-abs %.2e  at %d  (|x| rms %.2e)" % (k, np.linalg.norm(x - y) / max(np.linalg.norm(x), 1e-300), d.max(), int(d.argmax()), (x ** 2).mean() ** 0.5))
