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


REFDATA = "/root/reference/data"


def row_checksums(nn, nb):
    """order-independent per-row fingerprints of a neighbor list (small fixture instead of the rows)"""
    s1 = np.array([int(nb[i, :nn[i]].astype(np.int64).sum()) for i in range(len(nn))], np.int64)
    s2 = np.array([int((nb[i, :nn[i]].astype(np.int64) ** 2).sum() % 2147483647) for i in range(len(nn))], np.int64)
    return s1, s2


def parse_funcfl(path):
    """funcfl tables exactly as the reference's readEamFile parses them (before its 1-shift)"""
    with open(path) as f:
        f.readline()
        l2 = f.readline().split()
        l3 = f.readline().split()
        vals = np.array(f.read().split(), dtype=np.float64)
    mass = float(l2[1])
    nrho, drho, nr, dr, cut = int(l3[0]), float(l3[1]), int(l3[2]), float(l3[3]), float(l3[4])
    return dict(mass=mass, nrho=nrho, drho=drho, nr=nr, dr=dr, cut=cut, frho=vals[:nrho], zr=vals[nrho:nrho + nr],
                rhor=vals[nrho + nr:nrho + 2 * nr])


def eam_tables(r):
    """spline tables of the reference's global `eam` (common/eam.h:20-29) after initEam"""
    import ctypes as C

    class Funcfl(C.Structure):
        _fields_ = [("nrho", C.c_int), ("nr", C.c_int), ("drho", r.real), ("dr", r.real), ("cut", r.real),
                    ("mass", r.real), ("frho", C.POINTER(r.real)), ("rhor", C.POINTER(r.real)), ("zr", C.POINTER(r.real))]

    class Eam(C.Structure):
        _fields_ = [("fp", C.POINTER(r.real)), ("nmax", C.c_int), ("nrho", C.c_int), ("nr", C.c_int),
                    ("nrho_tot", C.c_int), ("nr_tot", C.c_int), ("dr", r.real), ("rdr", r.real), ("drho", r.real),
                    ("rdrho", r.real), ("frho", C.POINTER(r.real)), ("rhor", C.POINTER(r.real)),
                    ("z2r", C.POINTER(r.real)), ("rhor_spline", C.POINTER(r.real)), ("frho_spline", C.POINTER(r.real)),
                    ("z2r_spline", C.POINTER(r.real)), ("file", Funcfl)]
    e = Eam.in_dll(r.lib, "eam")
    A = np.ctypeslib.as_array
    d = dict(nr=e.nr, nrho=e.nrho, nr_tot=e.nr_tot, nrho_tot=e.nrho_tot, rdr=float(e.rdr), rdrho=float(e.rdrho),
             rhor_spline=A(e.rhor_spline, shape=(e.nr_tot,)).copy(), frho_spline=A(e.frho_spline, shape=(e.nrho_tot,)).copy(),
             z2r_spline=A(e.z2r_spline, shape=(e.nr_tot,)).copy())
    # knot 0 of each table is never written by the reference (uninitialised memory): zero it in the fixture
    # and so is the padding behind knot n (array2spline pads the length to a multiple of 64)
    for k, n in (("rhor_spline", e.nr), ("frho_spline", e.nrho), ("z2r_spline", e.nr)):
        d[k][:7] = 0.0
        d[k][(n + 1) * 7:] = 0.0
    d["_eam"] = e
    return d


def argon_case():
    """BASELINE config 3: data/argon, LJ, half neighbor lists, fixed-interval rebuilds"""
    r = RefVL("vl_dp_aos")
    r.lib.readParameter(__import__("ctypes").byref(r.param), (REFDATA + "/argon/mdbench_params.conf").encode())
    r.set_str("input_file", REFDATA + "/argon/input.gro")
    r.param.half_neigh = 1
    r.param.ntimes = 200
    r.setup()
    p = r.param
    d = dict(box=np.array([p.xlo, p.xhi, p.ylo, p.yhi, p.zlo, p.zhi]),
             params=np.array([p.epsilon, p.sigma, p.cutforce, p.skin, p.dt, p.temp, p.rho, p.mass]),
             ints=np.array([p.reneigh_every, p.nstat, p.half_neigh, p.ntimes]),
             x0=r.get("x"), v0=r.get("v"), nghost0=np.int32(r.atom.Nghost), maxneighs0=np.int32(r.neighbor.maxneighs))
    nn, nb = r.get("numneigh"), r.get("neighbors")
    d["numneigh0"] = nn
    d["rowsum0"], d["rowsq0"] = row_checksums(nn, nb)
    d["thermo0"] = np.array(r.thermo())
    r.computeForce()
    d["f0"] = r.get("f")
    for n in range(200):
        r.step(n)
    d["xN"], d["vN"], d["fN"] = r.get("x"), r.get("v"), r.get("f")
    d["thermoN"] = np.array(r.thermo())
    d["nghostN"] = np.int32(r.atom.Nghost)
    nn, nb = r.get("numneigh"), r.get("neighbors")
    d["numneighN"] = nn
    d["rowsumN"], d["rowsqN"] = row_checksums(nn, nb)
    np.savez_compressed(os.path.join(HERE, "argon_half.npz"), **d)
    print("argon", r.atom.Nlocal, d["nghost0"], d["thermo0"], d["thermoN"], "maxneighs", r.neighbor.maxneighs)


def eam_case(nx, nsteps, name, dmp=None):
    """EAM (verletlist only): generated Cu lattice, or BASELINE config 4 (copper_melting .dmp)"""
    r = RefVL("vl_dp_aos")
    r.param.force_field = 1
    r.set_str("eam_file", REFDATA + "/Cu_u3.eam")
    r.param.nx = r.param.ny = r.param.nz = nx
    r.param.ntimes = nsteps
    if dmp:
        r.set_str("input_file", dmp)
    r.setup()
    p = r.param
    t = eam_tables(r)
    e = t.pop("_eam")
    d = {("eam_" + k): np.asarray(v) for k, v in t.items()}
    ff = parse_funcfl(REFDATA + "/Cu_u3.eam")
    d.update({("funcfl_" + k): np.asarray(v) for k, v in ff.items()})
    d.update(nx=np.int32(nx), nsteps=np.int32(nsteps), box=np.array([p.xlo, p.xhi, p.ylo, p.yhi, p.zlo, p.zhi]),
             params=np.array([p.cutforce, p.cutneigh, p.dt, p.dtforce, p.temp, p.rho, p.mass]),
             x0=r.get("x"), v0=r.get("v"), nghost0=np.int32(r.atom.Nghost), thermo0=np.array(r.thermo()))
    nn, nb = r.get("numneigh"), r.get("neighbors")
    d["numneigh0"] = nn
    d["rowsum0"], d["rowsq0"] = row_checksums(nn, nb)
    r.computeForce()
    d["f0"] = r.get("f")
    d["fp0"] = np.ctypeslib.as_array(e.fp, shape=(r.atom.Nlocal + r.atom.Nghost,)).copy()
    rec = [r.computeThermo(0)]
    for n in range(nsteps):
        r.step(n)
        if (n + 1) % r.param.nstat == 0 and (n + 1) < nsteps:
            rec.append(r.computeThermo(n + 1))
    rec.append(r.computeThermo(-1))
    d["records"] = np.array(rec)
    d["thermoN"] = np.array(r.thermo())
    d["nghostN"] = np.int32(r.atom.Nghost)
    if r.atom.Nlocal <= 4000:
        d["xN"], d["vN"], d["fN"] = r.get("x"), r.get("v"), r.get("f")
    else:   # keep the big fixture small: positions/velocities have few digits in the input file
        d["x0"] = d["x0"].astype(np.float64)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    print(name, r.atom.Nlocal, d["nghost0"], rec)


if __name__ == "__main__":
    import json
    if "--extra" in sys.argv:
        if "--eam-only" not in sys.argv:
            argon_case()
        eam_case(5, 45, "eam_cu_nx5")
        eam_case(20, 200, "eam_cu_melting", dmp=REFDATA + "/copper_melting/input_eam_cu_one_atomtype_20x20x20.dmp")
        sys.exit(0)
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
