"""A/B of kernel variants on one GPU: per-launch force time, per-rebuild list time, loop total.
usage: python profiles/ab.py [--nx 128] [--steps 40] [--precision dp] name=value,name=value ...
each positional arg is one configuration (comma-separated mdb_setOption pairs; 'default' = none)"""
import argparse
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=128)
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--precision", default="dp")
ap.add_argument("--half", type=int, default=0)
ap.add_argument("configs", nargs="*", default=["default"])
a = ap.parse_args()
m = importlib.import_module("md-bench_b200")
for cfg in a.configs:
    s = m.Simulation(m.default_params(precision=m.DP if a.precision == "dp" else m.SP, nx=a.nx, ny=a.nx, nz=a.nx,
                                      half_neigh=a.half))
    if cfg != "default":
        for kv in cfg.split(","):
            k, v = kv.split("=")
            s.setOption(k, float(v))
    n = s.createAtom()
    s.setup(adjust=True)
    s.run(20)                      # warm-up (includes one rebuild)
    s.setTiming(True)
    s.resetKernelStats()
    rec, tm = s.run(a.steps)
    ks = s.kernelStats()
    s.setTiming(False)
    rec2, tm2 = s.run(a.steps)     # untimed-per-kernel loop: pure stream time
    print("%-40s atoms %d  force %.4f ms/launch  neigh %.3f ms/rebuild  loop(untimed) %.2f ms/%d steps = %.3f G atom-steps/s  T=%.12f"
          % (cfg, n, ks["force_ms"] / max(1, ks["force_launches"]), ks["neigh_ms"] / max(1, ks["neigh_launches"]),
             tm2["TOTAL"] * 1e3, a.steps, n * a.steps / tm2["TOTAL"] * 1e-9, rec[-1][1]), flush=True)
    s.close()
