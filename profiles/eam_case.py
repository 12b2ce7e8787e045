"""EAM (BASELINE config 4 physics: Cu_u3 funcfl tables, generated Cu lattice) timing through the C ABI.
usage: python profiles/eam_case.py [--nx 64] [--steps 100] [--precision dp] [--opt eam_variant=2]
The funcfl tables come from the committed fixture tests/golden/eam_cu_nx5.npz (made from the reference's data/Cu_u3.eam)."""
import argparse, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=64)
ap.add_argument("--steps", type=int, default=100)
ap.add_argument("--precision", default="dp")
ap.add_argument("--opt", action="append", default=[], help="name=value for mdb_setOption (A/B)")
a = ap.parse_args()
m = importlib.import_module("md-bench_b200")
g = np.load(os.path.join(ROOT, "tests", "golden", "eam_cu_nx5.npz"))
s = m.Simulation(m.default_params(precision=m.DP if a.precision == "dp" else m.SP, force_field=m.FF_EAM, nx=a.nx, ny=a.nx, nz=a.nx))
for kv in a.opt:
    s.setOption(kv.split("=")[0], float(kv.split("=")[1]))
s.setEam(int(g["funcfl_nrho"]), float(g["funcfl_drho"]), int(g["funcfl_nr"]), float(g["funcfl_dr"]), float(g["funcfl_cut"]),
         float(g["funcfl_mass"]), g["funcfl_frho"], g["funcfl_zr"], g["funcfl_rhor"])
n = s.createAtom()
s.setup(adjust=True)
listed, inside = s.countPairs()
s.run(20)                      # warm-up
s.setTiming(True); s.resetKernelStats()
rec, tm = s.run(a.steps)
k = s.kernelStats()
print("EAM nx %d %s: %d atoms, %d steps TOTAL %.4fs -> %.3f G atom-steps/s; force (3 passes) %.3f ms/call, neigh %.2f ms/rebuild; "
      "listed %.1f / in-cutoff %.1f pairs per atom at t=0; T %.6e; ghosts %d"
      % (a.nx, a.precision, n, a.steps, tm["TOTAL"], n * a.steps / tm["TOTAL"] / 1e9, k["force_ms"] / max(1, k["force_launches"]),
         k["neigh_ms"] / max(1, k["neigh_launches"]), listed / n, inside / n, rec[-1][1], s.counts()["Nghost"]))
s.close()
