"""Small driver for ncu captures: Cu FCC nx^3 LJ, setup + nsteps through the C ABI.
usage: python profiles/profile_case.py [--nx 64] [--steps 25] [--precision dp] [--half 0]"""
import argparse
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=64)
ap.add_argument("--steps", type=int, default=25)
ap.add_argument("--precision", default="dp")
ap.add_argument("--half", type=int, default=0)
ap.add_argument("--opt", action="append", default=[])
a = ap.parse_args()
m = importlib.import_module("md-bench_b200")
s = m.Simulation(m.default_params(precision=m.DP if a.precision == "dp" else m.SP, nx=a.nx, ny=a.nx, nz=a.nx,
                                  half_neigh=a.half))
for o in a.opt:
    k, v = o.split("=")
    s.setOption(k, float(v))
s.createAtom()
s.setup(adjust=True)
rec, tm = s.run(a.steps)
print("nx", a.nx, "steps", a.steps, "T", rec[-1][1], "TOTAL %.4fs" % tm["TOTAL"], s.kernelStats())
s.close()
