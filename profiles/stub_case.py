"""force kernel on the micro-benchmark's synthetic lists, per kernel variant. usage: stub_case.py [--na N] [--opt k=v ...]"""
import argparse, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--na", type=int, default=8388608)
ap.add_argument("--precision", default="dp")
ap.add_argument("--opt", action="append", default=[])
a = ap.parse_args()
m = importlib.import_module("md-bench_b200")
s = m.Simulation(m.default_params(precision=m.DP if a.precision == "dp" else m.SP, layout=m.SOA, nx=1, ny=1, nz=1, cutforce=1.0e6, skin=0.0))
for kv in a.opt:
    k, v = kv.split("=")
    s.setOption(k, float(v))
x = np.repeat((np.arange(a.na) * 1e-5)[:, None], 3, axis=1)
s.setAtoms(x, None)
out = []
for pat in ("seq", "fix") + tuple("%s:%d" % (p, w) for w in (16, 32, 64, 128, 256, 1024, 8192) for p in ("local", "localbank")):
    pat, _, w = pat.partition(":")
    s.stubNeighbors(pat, 76, 1, int(w or 12345))
    for _ in range(3):
        s.computeForceLJFullNeigh()
    t = min(s.computeForceLJFullNeigh() for _ in range(10))
    out.append("%s%s %.3f" % (pat, w, t * 1e3))
print(a.precision, a.opt, " ".join(out))
s.close()
