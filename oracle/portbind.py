"""ctypes binding to the oracle restatement (oracle/vl_oracle.c -> libmdoracle_{dp,sp}.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg; never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def build_port():
    subprocess.check_call(["make", "-s", "-C", HERE, "port"])


def _load(dp):
    path = os.path.join(HERE, "libmdoracle_%s.so" % ("dp" if dp else "sp"))
    if not os.path.exists(path):
        build_port()
    lib = C.CDLL(path)
    lib.ovl_new.restype = C.c_void_p
    lib.ovl_ptr.restype = C.c_void_p
    lib.ovl_get_real.restype = C.c_double
    lib.ovl_get_counter.restype = C.c_longlong
    return lib


class OracleVL:
    """Operator-by-operator access to the restated verletlist path (same method names as
    oracle/refbind.RefVL so tests can drive either)."""

    def __init__(self, dp=True):
        self.dp = dp
        self.lib = _load(dp)
        self.np_real = np.float64 if dp else np.float32
        self.h = C.c_void_p(self.lib.ovl_new())
        self._ff = 0

    def __del__(self):
        try:
            self.lib.ovl_free(self.h)
        except Exception:
            pass

    def _call(self, name, *a):
        return getattr(self.lib, name)(self.h, *a)

    def configure(self, nx=32, ny=None, nz=None, ntimes=200, nstat=100, reneigh_every=20,
                  half_neigh=0, pbc=(1, 1, 1), epsilon=1.0, sigma=1.0, cutforce=2.5, skin=0.3,
                  dt=0.005, temp=1.44, rho=0.8442, mass=1.0, force_field=0):
        ny = nx if ny is None else ny
        nz = nx if nz is None else nz
        d = C.c_double
        self._call("ovl_set_lj", d(epsilon), d(sigma), d(cutforce), d(skin), d(dt), d(temp), d(rho),
                   d(mass))
        self._call("ovl_set_run", nx, ny, nz, ntimes, nstat, reneigh_every, half_neigh, *pbc)
        self._call("ovl_set_force_field", force_field)
        self._ff = force_field

    def set_box(self, xlo, xhi, ylo, yhi, zlo, zhi):
        d = C.c_double
        self._call("ovl_set_box", d(xlo), d(xhi), d(ylo), d(yhi), d(zlo), d(zhi))

    def set_atoms(self, x, v=None):
        r = self.np_real
        x = np.ascontiguousarray(x, dtype=r)
        cols = [np.ascontiguousarray(x[:, k]) for k in range(3)]
        if v is not None:
            v = np.ascontiguousarray(v, dtype=r)
            cols += [np.ascontiguousarray(v[:, k]) for k in range(3)]
            ptrs = [c.ctypes.data_as(C.c_void_p) for c in cols]
        else:
            ptrs = [c.ctypes.data_as(C.c_void_p) for c in cols] + [None] * 3
        self._call("ovl_set_atoms", x.shape[0], *ptrs)

    def eam_set(self, nr, nrho, nr_tot, nrho_tot, rdr, rdrho, rhor, frho, z2r):
        r = self.np_real
        a = [np.ascontiguousarray(t, dtype=r) for t in (rhor, frho, z2r)]
        self._call("ovl_eam_set", nr, nrho, nr_tot, nrho_tot, C.c_double(rdr), C.c_double(rdrho),
                   *[t.ctypes.data_as(C.c_void_p) for t in a])

    def eam_from_funcfl(self, nrho, drho, nr, dr, cut, mass, frho, zr, rhor):
        r = self.np_real
        a = [np.ascontiguousarray(t, dtype=r) for t in (frho, zr, rhor)]
        d = C.c_double
        self._call("ovl_eam_from_funcfl", nrho, d(drho), nr, d(dr), d(cut), d(mass),
                   *[t.ctypes.data_as(C.c_void_p) for t in a])

    # operators ---------------------------------------------------------------------------
    def setup(self, create=True):
        self._call("ovl_setup", 1 if create else 0)

    def derive(self): self._call("ovl_derive")
    def create_atoms(self): self._call("ovl_create_atoms")
    def setup_neighbor(self): self._call("ovl_setup_neighbor")
    def setup_thermo(self): self._call("ovl_setup_thermo")
    def adjust_thermo(self): self._call("ovl_adjust_thermo")
    def setupPbc(self): self._call("ovl_setup_pbc")
    def updatePbc(self, reneigh=False): self._call("ovl_update_pbc")
    def updateAtomsPbc(self): self._call("ovl_update_atoms_pbc")
    def buildNeighbor(self): self._call("ovl_build_neighbor")
    def computeForce(self): self._call("ovl_compute_force")
    def initialIntegrate(self, reneigh=False): self._call("ovl_initial_integrate")
    def finalIntegrate(self, reneigh=False): self._call("ovl_final_integrate")
    def reneighbour(self): self._call("ovl_reneighbour")

    def thermo(self):
        T, P = C.c_double(), C.c_double()
        self._call("ovl_thermo", C.byref(T), C.byref(P))
        return T.value, P.value

    def run(self, nsteps):
        out = np.zeros(3 * (nsteps // max(1, self.geti("nstat")) + 3))
        n = self._call("ovl_run", nsteps, out.ctypes.data_as(C.c_void_p), len(out) // 3)
        return out[:3 * n].reshape(n, 3)

    def step(self, n):
        reneigh = (n + 1) % self.geti("reneigh_every") == 0
        self.initialIntegrate(reneigh)
        if reneigh:
            self.reneighbour()
        else:
            self.updatePbc(False)
        self.computeForce()
        self.finalIntegrate(reneigh)
        return reneigh

    # accessors ---------------------------------------------------------------------------
    def geti(self, k): return self._call("ovl_get_int", k.encode())
    def getr(self, k): return self._call("ovl_get_real", k.encode())
    def counter(self, which): return self._call("ovl_get_counter", which)

    def _arr(self, k, n, dtype):
        p = self._call("ovl_ptr", k.encode())
        if not p:
            return np.zeros(0, dtype=dtype)
        ct = {np.float64: C.c_double, np.float32: C.c_float, np.int32: C.c_int}[dtype]
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), shape=(n,))

    def get(self, what, ghosts=False):
        nl, ng = self.geti("Nlocal"), self.geti("Nghost")
        n = nl + (ng if ghosts else 0)
        if what in ("x", "v", "f"):
            names = {"x": "xyz", "v": ("vx", "vy", "vz"), "f": ("fx", "fy", "fz")}[what]
            return np.stack([self._arr(c, n, self.np_real) for c in names], axis=1).copy()
        if what in ("border_map", "PBCx", "PBCy", "PBCz"):
            return self._arr(what, ng, np.int32).copy()
        if what == "numneigh":
            return self._arr("numneigh", nl, np.int32).copy()
        if what == "neighbors":
            m = self.geti("maxneighs")
            return self._arr("neighbors", nl * m, np.int32).reshape(nl, m).copy()
        if what == "fp":
            return self._arr("fp", n, self.np_real).copy()
        if what == "stencil":
            return self._arr("stencil", self.geti("nstencil"), np.int32).copy()
        if what == "bincount":
            return self._arr("bincount", self.geti("mbins"), np.int32).copy()
        raise KeyError(what)

    def sorted_neighbor_sets(self):
        nn = self.get("numneigh")
        nb = self.get("neighbors")
        return nn, [np.sort(nb[i, :nn[i]]) for i in range(len(nn))]
