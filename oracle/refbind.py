"""ctypes binding to the UNMODIFIED MD-Bench reference built by oracle/Makefile into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (md-bench_b200/) may import this module;
it is used by tests/, by tests/golden/make_golden.py (fixture generation) and by bench.py's
cpu_baseline / --impl reference legs.

The ctypes structures below mirror the reference's plugin-boundary structs so that the reference's
own functions can be called on them:
  Parameter  <- reference src/common/parameter.h:27-61
  Atom       <- reference src/verletlist/atom.h:12-39
  Neighbor   <- reference src/verletlist/neighbor.h:18-28
  Stats      <- reference src/verletlist/stats.h:13-18
The driver flow (setup / reneighbour / time loop) restated in RefVL.setup()/run() follows
reference src/verletlist/main.c:36-95 and 244-288 call for call.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(HERE, "_ref")

FF_LJ, FF_EAM = 0, 1


def _structs(real):
    P = C.POINTER

    class Parameter(C.Structure):
        _fields_ = [
            ("force_field", C.c_int),
            ("param_file", C.c_char_p), ("input_file", C.c_char_p), ("vtk_file", C.c_char_p),
            ("xtc_file", C.c_char_p), ("write_atom_file", C.c_char_p),
            ("epsilon", real), ("sigma", real), ("sigma6", real), ("temp", real), ("rho", real),
            ("mass", real),
            ("ntypes", C.c_int), ("ntimes", C.c_int), ("nstat", C.c_int), ("reneigh_every", C.c_int),
            ("resort_every", C.c_int), ("prune_every", C.c_int), ("x_out_every", C.c_int),
            ("v_out_every", C.c_int), ("half_neigh", C.c_int),
            ("dt", real), ("dtforce", real), ("skin", real), ("cutforce", real), ("cutneigh", real),
            ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
            ("pbc_x", C.c_int), ("pbc_y", C.c_int), ("pbc_z", C.c_int),
            ("lattice", real),
            ("xlo", real), ("xhi", real), ("ylo", real), ("yhi", real), ("zlo", real), ("zhi", real),
            ("xprd", real), ("yprd", real), ("zprd", real),
            ("proc_freq", C.c_double),
            ("eam_file", C.c_char_p),
        ]

    class DeviceAtom(C.Structure):
        _fields_ = [(n, C.c_void_p) for n in
                    ("x", "y", "z", "vx", "vy", "vz", "fx", "fy", "fz", "border_map", "type",
                     "epsilon", "sigma6", "cutforcesq", "cutneighsq")]

    class Atom(C.Structure):
        _fields_ = [
            ("Natoms", C.c_int), ("Nlocal", C.c_int), ("Nghost", C.c_int), ("Nmax", C.c_int),
            ("x", P(real)), ("y", P(real)), ("z", P(real)),
            ("vx", P(real)), ("vy", P(real)), ("vz", P(real)),
            ("fx", P(real)), ("fy", P(real)), ("fz", P(real)),
            ("border_map", P(C.c_int)), ("type", P(C.c_int)), ("ntypes", C.c_int),
            ("epsilon", P(real)), ("sigma6", P(real)), ("cutforcesq", P(real)), ("cutneighsq", P(real)),
            ("d_atom", DeviceAtom),
        ]

    class Neighbor(C.Structure):
        _fields_ = [
            ("every", C.c_int), ("ncalls", C.c_int), ("maxneighs", C.c_int), ("half_neigh", C.c_int),
            ("neighbors", P(C.c_int)), ("numneigh", P(C.c_int)),
            ("d_neighbors", C.c_void_p), ("d_numneigh", C.c_void_p),
        ]

    class Stats(C.Structure):
        _fields_ = [(n, C.c_longlong) for n in
                    ("total_force_neighs", "total_force_iters", "atoms_within_cutoff",
                     "atoms_outside_cutoff")]

    return Parameter, Atom, Neighbor, Stats


def ref_available(variant="vl_dp_aos"):
    return os.path.exists(os.path.join(REFDIR, "libmdref_%s.so" % variant))


def ref_binary(variant="vl_dp_aos"):
    p = os.path.join(REFDIR, "MDBench-%s" % variant)
    return p if os.path.exists(p) else None


class RefVL:
    """The reference verletlist build, driven operator by operator.

    variant: 'vl_{dp|sp}_{aos|soa}'.  Set OMP_NUM_THREADS=1 in the environment BEFORE the first
    load for deterministic (parity) runs (the stats counters and half-list updates
    are racy under OpenMP).
    """

    def __init__(self, variant="vl_dp_aos"):
        self.variant = variant
        self.dp = "_dp_" in variant
        self.aos = variant.endswith("aos")
        self.real = C.c_double if self.dp else C.c_float
        self.np_real = np.float64 if self.dp else np.float32
        path = os.path.join(REFDIR, "libmdref_%s.so" % variant)
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle ref`)")
        self.lib = C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        self.Parameter, self.Atom, self.Neighbor, self.Stats = _structs(self.real)
        self.param = self.Parameter()
        self.atom = self.Atom()
        self.neighbor = self.Neighbor()
        self.stats = self.Stats()
        L = self.lib
        L.initParameter(C.byref(self.param))
        self._keep = []

    # -- helpers ---------------------------------------------------------------------------
    def _fp(self, name, restype, *argtypes):
        """Fetch a global function POINTER variable (computeForce, buildNeighbor, ...)."""
        addr = C.c_void_p.in_dll(self.lib, name).value
        return C.CFUNCTYPE(restype, *argtypes)(addr)

    def _glob(self, name, ctype):
        return ctype.in_dll(self.lib, name)

    def set_str(self, field, s):
        b = C.create_string_buffer(s.encode())
        self._keep.append(b)
        setattr(self.param, field, C.cast(b, C.c_char_p))

    # -- driver flow: reference src/verletlist/main.c:36-74 -----------------------------------
    def setup(self):
        L, p, a, n, s = self.lib, self.param, self.atom, self.neighbor, self.stats
        p.cutneigh = p.cutforce + p.skin                      # main.c:233
        if p.force_field == FF_EAM:
            L.initEam(C.byref(p))
        p.lattice = self.np_real((4.0 / float(p.rho)) ** (1.0 / 3.0))
        # the reference evaluates pow() in double and narrows on assignment (main.c:42-45)
        p.xprd = self.np_real(p.nx * self.np_real(p.lattice))
        p.yprd = self.np_real(p.ny * self.np_real(p.lattice))
        p.zprd = self.np_real(p.nz * self.np_real(p.lattice))
        L.initAtom(C.byref(a))
        L.initPbc(C.byref(a))
        L.initStats(C.byref(s))
        L.initNeighbor(C.byref(n), C.byref(p))
        if not p.input_file:
            L.createAtom(C.byref(a), C.byref(p))
        else:
            L.readAtom(C.byref(a), C.byref(p))
        L.setupNeighbor(C.byref(p))
        L.setupThermo(C.byref(p), a.Natoms)
        if not p.input_file:
            L.adjustThermo(C.byref(p), C.byref(a))
        L.setupPbc(C.byref(a), C.byref(p))
        self.updatePbc(True)
        self.buildNeighbor()
        L.initForce(C.byref(p))

    def updatePbc(self, reneigh=False):
        f = self._fp("updatePbc", None, C.c_void_p, C.c_void_p, C.c_bool)
        f(C.addressof(self.atom), C.addressof(self.param), reneigh)

    def updateAtomsPbc(self):
        f = self._fp("updateAtomsPbc", None, C.c_void_p, C.c_void_p, C.c_bool)
        f(C.addressof(self.atom), C.addressof(self.param), True)

    def setupPbc(self):
        self.lib.setupPbc(C.byref(self.atom), C.byref(self.param))

    def buildNeighbor(self):
        f = self._fp("buildNeighbor", None, C.c_void_p, C.c_void_p)
        f(C.addressof(self.atom), C.addressof(self.neighbor))

    def computeForce(self):
        f = self._fp("computeForce", C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p)
        return f(C.addressof(self.param), C.addressof(self.atom), C.addressof(self.neighbor),
                 C.addressof(self.stats))

    def initialIntegrate(self, reneigh=False):
        f = self._fp("initialIntegrate", None, C.c_bool, C.c_void_p, C.c_void_p)
        f(reneigh, C.addressof(self.param), C.addressof(self.atom))

    def finalIntegrate(self, reneigh=False):
        f = self._fp("finalIntegrate", None, C.c_bool, C.c_void_p, C.c_void_p)
        f(reneigh, C.addressof(self.param), C.addressof(self.atom))

    def reneighbour(self):                                  # main.c:76-95 (SORT_ATOMS off)
        self.updateAtomsPbc()
        self.setupPbc()
        self.updatePbc(True)
        self.buildNeighbor()

    def thermo(self):
        """(T, P) exactly as computeThermo prints them (common/thermo.c:55-66) without the print:
        recomputed here from v in the build's precision, serial order."""
        v = self.get("v")
        p = self.param
        r = self.np_real
        t = r(0.0)
        # the reference sums with -Ofast (vectorised); compare with tolerance, not bitwise
        t = r(np.sum((v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1] + v[:, 2] * v[:, 2]) * r(p.mass), dtype=r))
        if p.force_field == FF_LJ:
            mvv2e, dof = 1.0, (self.atom.Natoms * 3 - 3)
            pscale = 1.0 / 3 / p.xprd / p.yprd / p.zprd
        else:
            mvv2e, dof = 1.036427e-04, (self.atom.Natoms * 3 - 3) * 8.617343e-05
            pscale = 1.602176e+06 / 3 / p.xprd / p.yprd / p.zprd
        T = t * (mvv2e / dof)
        return float(T), float(T * dof * pscale)

    def computeThermo(self, iflag):
        """Call the reference's computeThermo (common/thermo.c:55-80) and parse the line it prints
        ("%i\t%e\t%e") -> (step, T, P) as the 7-significant-digit values of the report."""
        import tempfile
        libc = C.CDLL(None)
        libc.fflush(None)
        with tempfile.TemporaryFile(mode="w+b") as tf:
            saved = os.dup(1)
            os.dup2(tf.fileno(), 1)
            try:
                self.lib.computeThermo(C.c_int(iflag), C.byref(self.param), C.byref(self.atom))
                libc.fflush(None)
            finally:
                os.dup2(saved, 1)
                os.close(saved)
            tf.seek(0)
            line = tf.read().decode().strip().splitlines()[-1]
        a, b, c = line.split()
        return int(a), float(b), float(c)

    def step(self, n):
        """One iteration of the reference time loop, main.c:258-273. Returns True if it rebuilt."""
        reneigh = (n + 1) % self.param.reneigh_every == 0
        self.initialIntegrate(reneigh)
        if reneigh:
            self.reneighbour()
        else:
            self.updatePbc(False)
        self.computeForce()
        self.finalIntegrate(reneigh)
        return reneigh

    # -- array access ----------------------------------------------------------------------
    def get(self, what, ghosts=False):
        a = self.atom
        n = a.Nlocal + (a.Nghost if ghosts else 0)
        if what in ("x", "v", "f"):
            px = {"x": (a.x, a.y, a.z), "v": (a.vx, a.vy, a.vz), "f": (a.fx, a.fy, a.fz)}[what]
            if self.aos:
                return np.ctypeslib.as_array(px[0], shape=(n * 3,)).reshape(n, 3).copy()
            return np.stack([np.ctypeslib.as_array(q, shape=(n,)) for q in px], axis=1).copy()
        if what == "type":
            return np.ctypeslib.as_array(a.type, shape=(n,)).copy()
        if what == "border_map":
            return np.ctypeslib.as_array(a.border_map, shape=(a.Nghost,)).copy()
        if what in ("PBCx", "PBCy", "PBCz"):
            ptr = C.POINTER(C.c_int).in_dll(self.lib, what)
            return np.ctypeslib.as_array(ptr, shape=(a.Nghost,)).copy()
        if what == "numneigh":
            return np.ctypeslib.as_array(self.neighbor.numneigh, shape=(a.Nlocal,)).copy()
        if what == "neighbors":
            m = self.neighbor.maxneighs
            return np.ctypeslib.as_array(self.neighbor.neighbors, shape=(a.Nlocal * m,)).reshape(
                a.Nlocal, m).copy()
        raise KeyError(what)

    def set(self, what, arr, n=None):
        """Overwrite x or v of the first len(arr) atoms (used to feed identical inputs)."""
        a = self.atom
        arr = np.ascontiguousarray(arr, dtype=self.np_real)
        n = arr.shape[0]
        px = {"x": (a.x, a.y, a.z), "v": (a.vx, a.vy, a.vz), "f": (a.fx, a.fy, a.fz)}[what]
        if self.aos:
            np.ctypeslib.as_array(px[0], shape=(n * 3,))[:] = arr.reshape(-1)
        else:
            for k in range(3):
                np.ctypeslib.as_array(px[k], shape=(n,))[:] = arr[:, k]

    def neigh_globals(self):
        g = {}
        for nme in ("nbinx", "nbiny", "nbinz", "mbinx", "mbiny", "mbinz", "mbinxlo", "mbinylo",
                    "mbinzlo", "mbins", "atoms_per_bin", "nstencil"):
            g[nme] = self._glob(nme, C.c_int).value
        for nme in ("bininvx", "bininvy", "bininvz", "binsizex", "binsizey", "binsizez",
                    "cutneighsq", "cutneigh", "xprd", "yprd", "zprd"):
            g[nme] = self._glob(nme, self.real).value
        sp = C.POINTER(C.c_int).in_dll(self.lib, "stencil")
        g["stencil"] = np.ctypeslib.as_array(sp, shape=(g["nstencil"],)).copy()
        return g

    def sorted_neighbor_sets(self):
        """Rows of the neighbor list as sorted index arrays (the parity object of north_star)."""
        nn = self.get("numneigh")
        nb = self.get("neighbors")
        return nn, [np.sort(nb[i, :nn[i]]) for i in range(len(nn))]
