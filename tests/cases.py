"""Builders for the argon (BASELINE config 3) and EAM (config 4) fixtures: the same inputs fed to
the oracle restatement and to the CUDA Simulation."""
import numpy as np

from conftest import load_pkg
from portbind import OracleVL


def row_checksums(nn, nb):
    s1 = np.array([int(nb[i, :nn[i]].astype(np.int64).sum()) for i in range(len(nn))], np.int64)
    s2 = np.array([int((nb[i, :nn[i]].astype(np.int64) ** 2).sum() % 2147483647) for i in range(len(nn))], np.int64)
    return s1, s2


# ---- argon ---------------------------------------------------------------------------------------
def argon_oracle(g):
    eps, sig, cutf, skin, dt, temp, rho, mass = [float(v) for v in g["params"]]
    reneigh, nstat, half, ntimes = [int(v) for v in g["ints"]]
    o = OracleVL(True)
    o.configure(nx=1, ntimes=ntimes, nstat=nstat, reneigh_every=reneigh, half_neigh=half, epsilon=eps, sigma=sig,
                cutforce=cutf, skin=skin, dt=dt, temp=temp, rho=rho, mass=mass)
    o.set_box(*[float(v) for v in g["box"]])
    o.set_atoms(g["x0"], g["v0"])
    o.setup(create=False)
    return o


def argon_cuda(g, sort=False):
    m = load_pkg()
    eps, sig, cutf, skin, dt, temp, rho, mass = [float(v) for v in g["params"]]
    reneigh, nstat, half, ntimes = [int(v) for v in g["ints"]]
    b = [float(v) for v in g["box"]]
    p = m.default_params(epsilon=eps, sigma=sig, cutforce=cutf, skin=skin, dt=dt, temp=temp, rho=rho, mass=mass,
                         reneigh_every=reneigh, nstat=nstat, half_neigh=half, ntimes=ntimes, from_input=1,
                         xlo=b[0], xhi=b[1], ylo=b[2], yhi=b[3], zlo=b[4], zhi=b[5])
    s = m.Simulation(p)
    s.setOption("sort_atoms", 1 if sort else 0)
    s.setAtoms(g["x0"], g["v0"])
    s.setup(adjust=False)
    return s


# ---- EAM -----------------------------------------------------------------------------------------
def funcfl_args(g):
    return (int(g["funcfl_nrho"]), float(g["funcfl_drho"]), int(g["funcfl_nr"]), float(g["funcfl_dr"]),
            float(g["funcfl_cut"]), float(g["funcfl_mass"]), g["funcfl_frho"], g["funcfl_zr"], g["funcfl_rhor"])


def eam_oracle(g, from_dump=False, tables="funcfl"):
    nx = int(g["nx"])
    o = OracleVL(True)
    o.configure(nx=nx, force_field=1, ntimes=int(g["nsteps"]))
    o.eam_from_funcfl(*funcfl_args(g))
    if tables == "fixture":   # spline tables verbatim from the reference instead of our table builder
        o.eam_set(int(g["eam_nr"]), int(g["eam_nrho"]), int(g["eam_nr_tot"]), int(g["eam_nrho_tot"]),
                  float(g["eam_rdr"]), float(g["eam_rdrho"]), g["eam_rhor_spline"], g["eam_frho_spline"],
                  g["eam_z2r_spline"])
    if from_dump:
        o.set_box(*[float(v) for v in g["box"]])
        o.set_atoms(g["x0"], g["v0"])
        o.setup(create=False)
    else:
        o.setup(create=True)
    return o


def eam_cuda(g, from_dump=False, sort=False, dp=True):
    m = load_pkg()
    nx = int(g["nx"])
    kw = dict(precision=m.DP if dp else m.SP, force_field=m.FF_EAM, nx=nx, ny=nx, nz=nx, ntimes=int(g["nsteps"]))
    if from_dump:
        b = [float(v) for v in g["box"]]
        kw.update(from_input=1, xlo=b[0], xhi=b[1], ylo=b[2], yhi=b[3], zlo=b[4], zhi=b[5])
    s = m.Simulation(m.default_params(**kw))
    s.setOption("sort_atoms", 1 if sort else 0)
    s.setEam(*funcfl_args(g))
    if from_dump:
        s.setAtoms(g["x0"], g["v0"])
        s.setup(adjust=False)
    else:
        s.createAtom()
        s.setup(adjust=True)
    return s
