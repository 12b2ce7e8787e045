"""one launch per pattern for ncu: python profiles/stub_ncu.py  (run under ncu --metrics ...)"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
m = importlib.import_module("md-bench_b200")
na = 2097152
s = m.Simulation(m.default_params(precision=m.DP, layout=m.SOA, nx=1, ny=1, nz=1, cutforce=1.0e6, skin=0.0))
x = np.repeat((np.arange(na) * 1e-5)[:, None], 3, axis=1)
s.setAtoms(x, None)
for pat, w in (("seq", 0), ("local", 64), ("localbank", 64), ("local", 256), ("localbank", 256), ("local", 512), ("localbank", 512), ("local", 1024), ("localbank", 1024), ("local", 4096)):
    s.stubNeighbors(pat, 76, 1, w or 12345)
    s.computeForceLJFullNeigh()
    s.computeForceLJFullNeigh()
s.close()
