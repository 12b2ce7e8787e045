"""ncu target: one decomposed (or plain) run of 120 steps at nx^3 per brick. usage: dd_prof.py plain|brick [nx] [bricks]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
m = importlib.import_module("md-bench_b200")
mode, nx = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 64
grid = tuple(int(v) for v in (sys.argv[3] if len(sys.argv) > 3 else "2,1,1").split(","))
if mode == "plain":
    s = m.Simulation(m.default_params(nx=nx, ny=nx, nz=nx))
else:
    s = m.Decomposition(m.default_params(nx=nx * grid[0], ny=nx * grid[1], nz=nx * grid[2]), grid)
s.createAtom(); s.setup(adjust=True)
rec, tm = s.run(120)
print(mode, rec[-1], tm)
s.close()
