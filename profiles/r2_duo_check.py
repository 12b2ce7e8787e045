"""duo rows vs the one-atom-per-lane kernels: forces after a melted start and 200-step trajectories (DP rel 1e-10 asked)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
m = importlib.import_module("md-bench_b200")

def run(nx, prec, opts, steps=200):
    s = m.Simulation(m.default_params(precision=prec, nx=nx, ny=nx, nz=nx))
    for k, v in opts.items():
        s.setOption(k, float(v))
    s.createAtom()
    s.setup(adjust=True)
    rec, tm = s.run(steps)
    s.computeForce()
    out = (np.array(rec), s.get("x"), s.get("v"), s.get("f"))
    nn, nb = s.neighbors()
    s.close()
    return out + (nn, nb)

for prec, tol in ((m.DP, 1e-10), (m.SP, 1e-4)):
    for nx in (6, 17):
        ref = run(nx, prec, {})
        for opts in [dict(kv.split("=") for kv in a.split(",")) for a in sys.argv[1:]] or ({"duo": 1}, {"duo": 4}, {"duo": 4, "duo_bf": 1}, {"duo": 4, "duo_u": 4, "duo_minb": 4}, {"duo": 4, "fuse_force": 0}):
            got = run(nx, prec, opts)
            errs = [np.abs(g - r).max() / np.abs(r).max() for g, r in zip(got[:4], ref[:4])]
            same_sets = all(np.array_equal(np.sort(got[5][i, :got[4][i]]), np.sort(ref[5][i, :ref[4][i]])) for i in range(len(ref[4])))
            ok = max(errs) <= tol and np.array_equal(got[4], ref[4]) and same_sets
            print("prec %d nx %2d %-40s thermo %.2e x %.2e v %.2e f %.2e lists %s -> %s" % (prec, nx, opts, *errs, same_sets, "ok" if ok else "FAIL"))
