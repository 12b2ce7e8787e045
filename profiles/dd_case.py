"""force-kernel time per launch over the course of a run, single domain vs brick mode (one GPU).
usage: python profiles/dd_case.py [--nx 64] [--chunks 10] [--bricks 1,1,1]"""
import argparse
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=64)
ap.add_argument("--chunks", type=int, default=10)
ap.add_argument("--bricks", default="1,1,1")
ap.add_argument("--opt", action="append", default=[])
a = ap.parse_args()
m = importlib.import_module("md-bench_b200")
grid = tuple(int(v) for v in a.bricks.split(","))
for mode in ("plain", "brick"):
    if mode == "plain":
        s = m.Simulation(m.default_params(nx=a.nx * grid[0], ny=a.nx * grid[1], nz=a.nx * grid[2]))
    else:
        s = m.Decomposition(m.default_params(nx=a.nx * grid[0], ny=a.nx * grid[1], nz=a.nx * grid[2]), grid)
    for kv in a.opt:
        k, v = kv.split("=")
        s.setOption(k, float(v))
    s.createAtom()
    s.setup(adjust=True)
    s.setTiming(True)
    out = []
    for c in range(a.chunks):
        s.resetKernelStats()
        rec, tm = s.run(20)
        k = s.kernelStats()
        out.append("%.3f/%.2f" % (k["force_ms"] / max(1, k["force_launches"]), k["neigh_ms"] / max(1, k["neigh_launches"])))
    print(mode, "force ms per launch / neigh ms per rebuild, per 20-step chunk:", " ".join(out), "| counts", s.counts())
    s.close()
