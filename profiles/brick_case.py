"""Small driver for ncu captures of the BRICK variant of the fused force kernel (k_force_lj_full_fi<.., XY, ZG>): a decomposed box on
ONE GPU (bricks in one process, halo by device copies), Cu FCC nx^3 unit cells per brick.
usage: python profiles/brick_case.py [--nx 128] [--bricks 2,1,1] [--steps 45] [--opt k=v ...]"""
import argparse
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=128)
ap.add_argument("--bricks", default="2,1,1")
ap.add_argument("--steps", type=int, default=45)
ap.add_argument("--opt", action="append", default=[])
a = ap.parse_args()
m = importlib.import_module("md-bench_b200")
g = tuple(int(v) for v in a.bricks.split(","))
d = m.Decomposition(m.default_params(nx=a.nx * g[0], ny=a.nx * g[1], nz=a.nx * g[2]), g)
for kv in a.opt:
    k, v = kv.split("=")
    d.setOption(k, float(v))
d.createAtom()
d.setup(adjust=True)
rec, tm = d.run(a.steps)
print("bricks", g, "nx", a.nx, "steps", a.steps, "T", rec[-1][1], "TOTAL %.4fs" % tm["TOTAL"], d.kernelStats(), d.counts())
d.close()
