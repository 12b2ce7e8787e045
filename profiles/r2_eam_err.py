"""EAM parity margins per kernel generation against the reference fixtures (tests/golden/eam_cu_*.npz): prints the actual
relative errors so that the tolerance written in tests/test_gpu_parity.py is the measured one."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from cases import eam_cuda, eam_oracle
from parity import rel_err
G = os.path.join(ROOT, "tests", "golden")
g = np.load(os.path.join(G, "eam_cu_nx5.npz"))
o = eam_oracle(g)
o.set_atoms(g["x0"], g["v0"]); o.setup(create=False)
o.run(int(g["nsteps"]))
print("oracle vs reference fixture after %d steps: x %.2e v %.2e" % (int(g["nsteps"]), rel_err(o.get("x"), g["xN"]), rel_err(o.get("v"), g["vN"])))
for variant in (0, 1, 2):
    s = eam_cuda(g)
    s.setOption("eam_variant", variant)
    s.setAtoms(g["x0"], g["v0"]); s.setup(adjust=False)
    s.computeForceEam()
    e0 = np.abs(s.get("f") - g["f0"]).max()
    rec, _ = s.run(int(g["nsteps"]))
    print("lattice  eam_variant %d: |f0 - ref| %.2e (abs), after %d steps: T %.2e x %.2e v %.2e f %.2e" % (
        variant, e0, int(g["nsteps"]), abs(rec[-1][1] - g["thermoN"][0]) / g["thermoN"][0], rel_err(s.get("x"), g["xN"]),
        rel_err(s.get("v"), g["vN"]), np.abs(s.get("f") - g["fN"]).max() / np.abs(g["fN"]).max()))
    s.close()
g = np.load(os.path.join(G, "eam_cu_melting.npz"))
for variant in (0, 1, 2):
    s = eam_cuda(g, from_dump=True)
    s.setOption("eam_variant", variant)
    s.computeForceEam()
    rec, _ = s.run(200)
    err = max(max(abs(T - gT) / gT, abs(P - gP) / abs(gP)) for (st, T, P), (gs, gT, gP) in zip(rec, g["records"]))
    print("melting  eam_variant %d: printed thermo records (7 digits) %.2e, final T vs reference %.2e" % (variant, err, abs(rec[-1][1] - g["thermoN"][0]) / g["thermoN"][0]))
    s.close()
