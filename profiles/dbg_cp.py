import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/tests"); sys.path.insert(0, ROOT + "/oracle")
import numpy as np
from test_gpu_cp import make_cp, make_oracle, assert_same_structure, check_forces, initial_atoms
dp, N, nx = True, 8, int(sys.argv[1]) if len(sys.argv) > 1 else 32
x, v = initial_atoms(dp, nx)
s, o = make_cp(dp, N, nx), make_oracle(dp, N, nx)
s.setAtoms(x, v); o.set_atoms(x, v)
s.setup(adjust=False); o.setup()
assert_same_structure(o, s); print("structure t0 ok")
s.computeForce(); o.computeForce(); check_forces(s, o, dp); print("forces t0 ok")
for n in range(41):
    s.step(n); o.step(n)
    fs, fo = np.nan_to_num(s.cl("f")), np.nan_to_num(o.cl("f"))
    vs, vo = np.nan_to_num(s.cl("v")), np.nan_to_num(o.cl("v"))
    real = np.isfinite(o.cl("x")[:len(fo)])
    print(n, "df", np.abs(fs - fo)[real].max(), "fsum", fs[real.nonzero()].sum(), "dv", np.abs(vs - vo)[real].max(), "vsum gpu", (vs * real).sum(axis=(0, 2)), "oracle", (vo * real).sum(axis=(0, 2)))
    if (n + 1) % 20 == 0:
        try:
            assert_same_structure(o, s, tol=1e-10); print("structure ok at", n)
        except AssertionError as e:
            print("STRUCTURE MISMATCH", e)
