"""Small cases for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): every kernel family of the hot path once
at a size that finishes in seconds under the tool -- verletlist LJ full lists (fused step, two rebuilds), half lists (RED
atomics), EAM-free; clusterpair 4x4 SP (two-lane kernel) and DP, half; a 2x1x1 decomposition on one GPU (bricks in one process:
in-place fused kernel, gather copies, migration, halo by device copies).
usage: compute-sanitizer --tool memcheck python profiles/sanitizer_case.py"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
m = importlib.import_module("md-bench_b200")


def vl(dp, half, nx=8, steps=45, **opt):
    s = m.Simulation(m.default_params(precision=m.DP if dp else m.SP, nx=nx, ny=nx, nz=nx, half_neigh=half))
    for k, v in opt.items():
        s.setOption(k, v)
    s.createAtom()
    s.setup(adjust=True)
    rec, _ = s.run(steps)
    print("verletlist dp=%d half=%d %s: T(%d) = %.12f" % (dp, half, opt, steps, rec[-1][1]))
    s.close()


def cp(dp, half, n, nx=8, steps=45):
    s = m.ClusterSimulation(m.default_params(precision=m.DP if dp else m.SP, nx=nx, ny=nx, nz=nx, half_neigh=half), cluster_n=n)
    s.createAtom()
    s.setup(adjust=True)
    rec, _ = s.run(steps)
    print("clusterpair dp=%d half=%d 4x%d: T(%d) = %.9f" % (dp, half, n, steps, rec[-1][1]))
    s.close()


def dd(grid=(2, 1, 1), nx=6, steps=45):
    s = m.Decomposition(m.default_params(nx=nx * grid[0], ny=nx * grid[1], nz=nx * grid[2]), grid)
    s.createAtom()
    s.setup(adjust=True)
    rec, _ = s.run(steps)
    print("decomposition %s: T(%d) = %.12f" % (grid, steps, rec[-1][1]))
    s.close()


vl(True, 0)
vl(False, 0)
vl(True, 1)
vl(True, 0, sort_atoms=1, sort_block=2)
cp(False, 0, 4)
cp(False, 0, 8)
cp(True, 0, 4)
cp(True, 1, 4)
dd()
print("sanitizer_case done")
