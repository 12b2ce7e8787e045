"""Generate the CLUSTERPAIR golden fixtures in tests/golden/ from the reference's clusterpair builds
(oracle/_ref, `make -C oracle ref-cp ref-cpref`).  Run in the build container:

    OMP_NUM_THREADS=1 python tests/golden/make_golden_cp.py

cp44_{sp,dp}_nx6.npz : the reference's scalar kernel (computeForceLJRef, M x N = 4 x 4; oracle/Makefile
                       ref-cpref): cluster structures and lists at t = 0, forces, and the state after 45 steps.
cp48_{sp,dp}_nx6.npz : the UNMODIFIED AVX-512 builds (4 x 8): structures and lists at t = 0.
thermo_cp.json       : thermo records of BASELINE config 2 (Cu FCC 32^3, 200 steps) from the scalar 4x4 builds.
Every value was produced by reference code, none by our own.
"""
import json
import os
import sys

import numpy as np

os.environ["OMP_NUM_THREADS"] = "1"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from cpbind import RefCP  # noqa: E402


def structures(r, tag, d):
    d[tag + "_counts"] = np.array([r.geti(k) for k in ("Nlocal", "Nghost", "Nclusters_local", "Nclusters_ghost", "dummy_cj")],
                                  np.int32)
    d[tag + "_inat"], d[tag + "_ibb"] = r.iclusters()
    d[tag + "_jnat"], d[tag + "_jbb"] = r.jclusters()
    d[tag + "_ibin"] = r.icluster_bin()
    for k, v in r.ghost_map().items():
        d[tag + "_" + k] = v
    d[tag + "_clx"] = r.cl("x")
    d[tag + "_clv"] = r.cl("v")
    nn, nm, rows = r.cluster_lists()
    d[tag + "_numneigh"], d[tag + "_numneigh_masked"] = nn, nm
    d[tag + "_nnz"] = np.array([len(q) for q in rows], np.int32)     # without dummy padding
    d[tag + "_nbr_flat"] = np.concatenate(rows).astype(np.int32)


def case(variant, nx, nsteps, name, half=0):
    r = RefCP(variant)
    r.configure(nx=nx, half_neigh=half)
    r.setup()
    d = {"nx": np.int32(nx), "N": np.int32(r.N), "half": np.int32(half), "nsteps": np.int32(nsteps)}
    d["x0"], d["v0"] = r.atoms("x"), r.atoms("v")
    structures(r, "t0", d)
    if variant.startswith("cpref"):
        r.computeForce()
        d["t0_clf"] = r.cl("f")
        for n in range(nsteps):
            r.step(n)
        r.updateSingleAtoms()
        d["tN_x"], d["tN_v"] = r.atoms("x"), r.atoms("v")
        d["tN_thermo"] = np.array(r.thermo())
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    print(name, d["t0_counts"], d.get("tN_thermo"))


REFDATA = "/root/reference/data"


def argon_case(variant, name, nsteps=105):
    """clusterpair with -p data/argon/mdbench_params.conf -i data/argon/input.gro (box from the reader, 1000 atoms,
    cutneigh 1.9 > L/2, one rebuild at step 100): structures at t = 0, forces, atoms after nsteps"""
    r = RefCP(variant)
    r.setup_from_files(REFDATA + "/argon/mdbench_params.conf", REFDATA + "/argon/input.gro")
    p = r.param
    d = {"nx": np.int32(1), "N": np.int32(r.N), "half": np.int32(p.half_neigh), "nsteps": np.int32(nsteps),
         "box": np.array([p.xprd, p.yprd, p.zprd], np.float64),
         "params": np.array([p.epsilon, p.sigma, p.cutforce, p.skin, p.dt, p.temp, p.rho, p.mass], np.float64),
         "ints": np.array([p.reneigh_every, p.nstat], np.int32)}
    d["x0"], d["v0"] = r.atoms("x"), r.atoms("v")
    structures(r, "t0", d)
    r.computeForce()
    d["t0_clf"] = r.cl("f")
    for n in range(nsteps):
        r.step(n)
    r.updateSingleAtoms()
    d["tN_x"], d["tN_v"] = r.atoms("x"), r.atoms("v")
    d["tN_thermo"] = np.array(r.thermo())
    d["tN_counts"] = np.array([r.geti(k) for k in ("Nclusters_local", "Nclusters_ghost")], np.int32)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    print(name, d["t0_counts"], d["tN_thermo"], d["tN_counts"])


def thermo_run(variant, nx=32, nsteps=200):
    r = RefCP(variant)
    r.configure(nx=nx, ntimes=nsteps)
    r.setup()
    rec = [(0,) + r.thermo()]
    r.computeForce()
    for n in range(nsteps):
        r.step(n)
        if (n + 1) % r.param.nstat == 0 and n + 1 < nsteps:
            rec.append((n + 1,) + r.thermo())     # atom arrays as they are (half-step v after a rebuild)
    r.updateSingleAtoms()
    rec.append((nsteps,) + r.thermo())
    print(variant, rec)
    return {"variant": variant, "nx": nx, "nsteps": nsteps, "N": r.N, "records": [list(map(float, q)) for q in rec],
            "nghost_atoms": int(r.geti("Nghost")), "nclusters_ghost": int(r.geti("Nclusters_ghost"))}


if __name__ == "__main__":
    if "--argon-only" in sys.argv:
        argon_case("cpref44_dp", "cp44_dp_argon")
        sys.exit(0)
    case("cpref44_sp", 6, 45, "cp44_sp_nx6")
    case("cpref44_dp", 6, 45, "cp44_dp_nx6")
    case("cpref44_dp", 6, 45, "cp44_dp_half_nx6", half=1)
    case("cpref48_dp", 6, 45, "cp48ref_dp_nx6")
    case("cp_dp_aos", 6, 0, "cp48_dp_nx6")
    case("cp_sp_aos", 6, 0, "cp48_sp_nx6")
    argon_case("cpref44_dp", "cp44_dp_argon")
    out = [thermo_run("cpref44_sp"), thermo_run("cpref44_dp")]
    json.dump(out, open(os.path.join(HERE, "thermo_cp.json"), "w"), indent=1)
