"""Generate the golden fixtures in tests/golden/ from the UNMODIFIED reference (oracle/_ref, built
from /root/reference by `make -C oracle ref`).  Run in the build container:

    OMP_NUM_THREADS=1 python tests/golden/make_golden.py

The reference ships no tests or golden vectors, so these files ARE the pinned
reference outputs: every value in them was produced by reference code, none by our own.
"""
import os
import sys

import numpy as np

os.environ["OMP_NUM_THREADS"] = "1"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from refbind import RefVL  # noqa: E402


def rows_to_csr(nn, nb):
    """numneigh + padded rows -> (offsets, concatenated sorted rows)"""
    off = np.zeros(len(nn) + 1, np.int64)
    off[1:] = np.cumsum(nn)
    flat = np.concatenate([np.sort(nb[i, :nn[i]]) for i in range(len(nn))]).astype(np.int32)
    return off, flat


def snapshot(r, tag, d):
    d[tag + "_x"] = r.get("x", ghosts=True)
    d[tag + "_v"] = r.get("v")
    d[tag + "_f"] = r.get("f")
    d[tag + "_border_map"] = r.get("border_map")
    for k in ("PBCx", "PBCy", "PBCz"):
        d[tag + "_" + k] = r.get(k)
    nn, nb = r.get("numneigh"), r.get("neighbors")
    d[tag + "_numneigh"] = nn
    d[tag + "_nbr_off"], d[tag + "_nbr_flat"] = rows_to_csr(nn, nb)
    # unsorted first rows, to pin the row ORDER too (stencil order x ascending index)
    d[tag + "_row0_raw"] = nb[0, :nn[0]].copy()
    d[tag + "_maxneighs"] = np.int32(r.neighbor.maxneighs)
    d[tag + "_thermo"] = np.array(r.thermo())


def lj_case(variant, nx, half, nsteps, name):
    r = RefVL(variant)
    r.param.nx = r.param.ny = r.param.nz = nx
    r.param.half_neigh = half
    r.setup()
    d = {"nx": np.int32(nx), "half": np.int32(half), "nsteps": np.int32(nsteps)}
    g = r.neigh_globals()
    for k, v in g.items():
        d["ng_" + k] = np.asarray(v)
    r.computeForce()
    snapshot(r, "t0", d)
    for n in range(nsteps):
        r.step(n)
    snapshot(r, "tN", d)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    print(name, "Nlocal", r.atom.Nlocal, "Nghost", r.atom.Nghost, "T", d["tN_thermo"])


def thermo_case(variant, nx, nsteps, half=0):
    r = RefVL(variant)
    r.param.nx = r.param.ny = r.param.nz = nx
    r.param.ntimes = nsteps
    r.param.half_neigh = half
    r.setup()
    rec = [r.computeThermo(0)]
    r.computeForce()
    for n in range(nsteps):
        r.step(n)
        if (n + 1) % r.param.nstat == 0 and (n + 1) < nsteps:
            rec.append(r.computeThermo(n + 1))
    rec.append(r.computeThermo(-1))
    return dict(variant=variant, nx=nx, nsteps=nsteps, half=half, records=rec, nghost=int(r.atom.Nghost),
                T_full=r.thermo()[0], P_full=r.thermo()[1])


if __name__ == "__main__":
    import json
    lj_case("vl_dp_aos", 6, 0, 45, "lj_dp_full_nx6")
    lj_case("vl_dp_aos", 6, 1, 45, "lj_dp_half_nx6")
    lj_case("vl_sp_soa", 6, 0, 45, "lj_sp_full_nx6")
    th = [thermo_case("vl_dp_aos", 32, 200), thermo_case("vl_sp_soa", 32, 200),
          thermo_case("vl_dp_aos", 8, 200), thermo_case("vl_dp_aos", 8, 200, half=1),
          thermo_case("vl_sp_soa", 8, 200)]
    with open(os.path.join(HERE, "thermo_lj.json"), "w") as f:
        json.dump(th, f, indent=1)
    for t in th:
        print(t)
