"""Small driver for timing / ncu captures of the CLUSTERPAIR scheme: Cu FCC nx^3 LJ, setup + nsteps through the C ABI.
usage: python profiles/cp_case.py [--nx 64] [--steps 100] [--precision sp] [--n 4] [--half 0] [--opt k=v ...]"""
import argparse
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=64)
ap.add_argument("--steps", type=int, default=100)
ap.add_argument("--precision", default="sp")
ap.add_argument("--n", type=int, default=4)
ap.add_argument("--half", type=int, default=0)
ap.add_argument("--timing", type=int, default=1)
ap.add_argument("--opt", action="append", default=[])
a = ap.parse_args()
m = importlib.import_module("md-bench_b200")
s = m.ClusterSimulation(m.default_params(precision=m.DP if a.precision == "dp" else m.SP, nx=a.nx, ny=a.nx, nz=a.nx,
                                         half_neigh=a.half), cluster_n=a.n)
for kv in a.opt:
    k, v = kv.split("=")
    s.setOption(k, float(v))
n = s.createAtom()
t0 = time.time()
s.setup(adjust=True)
s.sync()
t1 = time.time()
s.setTiming(bool(a.timing))
s.resetKernelStats()
rec, tm = s.run(a.steps)
k = s.kernelStats()
cp, inside = s.countPairs()
c = s.counts()
print("CP nx %d %s 4x%d half %d: setup %.3fs, %d steps TOTAL %.4fs -> %.3f G atom-steps/s; force %.3f ms/launch (%d), neigh %.2f ms/rebuild (%d); "
      "T %.9f; cluster pairs/cluster %.2f, in-cutoff pairs/atom %.2f, ghosts %d, maxneighs %d, launches %d"
      % (a.nx, a.precision, a.n, a.half, t1 - t0, a.steps, tm["TOTAL"], n * a.steps / tm["TOTAL"] / 1e9,
         k["force_ms"] / max(1, k["force_launches"]), k["force_launches"], k["neigh_ms"] / max(1, k["neigh_launches"]),
         k["neigh_launches"], rec[-1][1], cp / c["Nclusters_local"], inside / n, c["Nclusters_ghost"], c["maxneighs"],
         k["launches"]))
s.close()
