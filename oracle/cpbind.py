"""ctypes bindings for the CLUSTERPAIR scheme: OracleCP (oracle/cp_oracle.c, our restatement) and RefCP
(the reference's clusterpair libraries built by oracle/Makefile into oracle/_ref/).

TEST INFRASTRUCTURE ONLY: imported by tests/, tests/golden/make_golden.py, __graft_entry__.smoke() and
bench.py's cpu_baseline leg; never by the product package.

Both classes expose the reference driver's operator names (clusterpair/main.c:40-93, neighbor.h:42-50):
buildClusters, defineJClusters, setupPbc, binClusters, buildNeighbor, updateSingleAtoms, updateAtomsPbc,
updatePbc, computeForce, initialIntegrate, finalIntegrate, reneighbour -- and the same accessors
(clusters(), cluster_lists(), ...), so a test can drive either.
"""
import ctypes as C
import os

import numpy as np

from portbind import OracleVL, build_port
from refbind import REFDIR, _structs as _vl_structs

HERE = os.path.dirname(os.path.abspath(__file__))
M = 4


def initial_atoms(dp, nx, ny=None, nz=None):
    """createAtom + adjustThermo (identical in both schemes, clusterpair/atom.c:49-180 == verletlist/atom.c):
    x, v of the generated lattice from the verletlist oracle"""
    o = OracleVL(dp)
    o.configure(nx=nx, ny=ny, nz=nz)
    o.derive(); o.create_atoms(); o.setup_neighbor(); o.setup_thermo(); o.adjust_thermo()
    return o.get("x"), o.get("v")


def cluster_view(N, ncl_tiles, flat, comps=3):
    """cl_x-like flat array -> (tiles, comps, N): tile t holds j-cluster t (and i-clusters 2t, 2t+1 when N = 8)"""
    return np.asarray(flat[:ncl_tiles * comps * N]).reshape(ncl_tiles, comps, N)


class _Common:
    """accessors shared by OracleCP and RefCP (subclasses provide _ints(), _arr())"""

    def tiles(self):
        ncl, ngh = self.geti("Nclusters_local"), self.geti("Nclusters_ghost")
        jfac = max(1, self.N // M)
        return ncl // jfac, ngh

    def iclusters(self):
        """(natoms[ncl], bbox[ncl, 6]) of the local i-clusters"""
        return self._clusters("iclusters", self.geti("Nclusters_local"))

    def jclusters(self):
        ncj, ngh = self.tiles()
        return self._clusters("jclusters", ncj + ngh)

    def cl(self, what):
        """cluster data 'x' | 'v' | 'f' as (tiles, 3, N), local tiles then ghost tiles (x only)"""
        ncj, ngh = self.tiles()
        nt = ncj + (ngh if what == "x" else 0)
        return cluster_view(self.N, nt, self._arr("cl_" + what, nt * 3 * self.N, self.np_real)).copy()

    def cluster_lists(self, strip_dummy=True):
        """numneigh, numneigh_masked, [sorted j-cluster ids per i-cluster]"""
        ncl = self.geti("Nclusters_local")
        mx = self.geti("maxneighs")
        nn = self._arr("numneigh", ncl, np.int32).copy()
        nm = self._arr("numneigh_masked", ncl, np.int32).copy()
        nb = self._arr("neighbors", ncl * mx, np.int32).reshape(ncl, mx)
        dummy = self.geti("dummy_cj")
        rows = []
        for ci in range(ncl):
            r = nb[ci, :nn[ci]]
            if strip_dummy:
                r = r[r != dummy]
            rows.append(np.sort(r))
        return nn, nm, rows

    def raw_lists(self):
        """numneigh, numneigh_masked, rows (ncl, maxneighs) in list order"""
        ncl, mx = self.geti("Nclusters_local"), self.geti("maxneighs")
        return (self._arr("numneigh", ncl, np.int32).copy(), self._arr("numneigh_masked", ncl, np.int32).copy(),
                self._arr("neighbors", ncl * mx, np.int32).reshape(ncl, mx).copy())

    def ghost_map(self):
        ngh = self.geti("Nclusters_ghost")
        return {k: self._arr(k, ngh, np.int32).copy() for k in ("border_map", "PBCx", "PBCy", "PBCz")}

    def atoms(self, what):
        n = self.geti("Nlocal")
        return self._atoms(what, n)

    def step(self, n):
        """one iteration of the reference loop, clusterpair/main.c:246-266 (prune_every never fires)"""
        reneigh = (n + 1) % self.geti("reneigh_every") == 0
        self.initialIntegrate()
        if reneigh:
            self.reneighbour()
        else:
            self.updatePbc(False)
        self.computeForce()
        self.finalIntegrate()
        return reneigh


class OracleCP(_Common):
    def __init__(self, dp=True, N=4, vector_width=0):
        path = os.path.join(HERE, "libmdoracle_%s.so" % ("dp" if dp else "sp"))
        if not os.path.exists(path):
            build_port()
        self.lib = C.CDLL(path)
        if not hasattr(self.lib, "ocp_new"):
            build_port()
            self.lib = C.CDLL(path)
        self.lib.ocp_new.restype = C.c_void_p
        self.lib.ocp_ptr.restype = C.c_void_p
        self.lib.ocp_get_real.restype = C.c_double
        self.dp, self.N = dp, N
        self.np_real = np.float64 if dp else np.float32
        self.h = C.c_void_p(self.lib.ocp_new(N, vector_width))

    def __del__(self):
        try:
            self.lib.ocp_free(self.h)
        except Exception:
            pass

    def _call(self, name, *a):
        return getattr(self.lib, name)(self.h, *a)

    def configure(self, nx=32, ny=None, nz=None, ntimes=200, nstat=100, reneigh_every=20, half_neigh=0, epsilon=1.0,
                  sigma=1.0, cutforce=2.5, skin=0.3, dt=0.005, temp=1.44, rho=0.8442, mass=1.0):
        ny = nx if ny is None else ny
        nz = nx if nz is None else nz
        d = C.c_double
        self._call("ocp_set_lj", d(epsilon), d(sigma), d(cutforce), d(skin), d(dt), d(temp), d(rho), d(mass))
        self._call("ocp_set_run", nx, ny, nz, ntimes, nstat, reneigh_every, half_neigh)

    def set_box(self, xprd, yprd, zprd):
        d = C.c_double
        self._call("ocp_set_box", d(xprd), d(yprd), d(zprd))

    def set_atoms(self, x, v):
        r = self.np_real
        cols = [np.ascontiguousarray(np.asarray(a, dtype=r)[:, k]) for a in (x, v) for k in range(3)]
        self._call("ocp_set_atoms", len(x), *[c.ctypes.data_as(C.c_void_p) for c in cols])

    def setup(self): self._call("ocp_setup")
    def setupNeighbor(self): self._call("ocp_setup_neighbor")
    def buildClusters(self): self._call("ocp_build_clusters")
    def defineJClusters(self): self._call("ocp_define_jclusters")
    def setupPbc(self): self._call("ocp_setup_pbc")
    def binClusters(self): self._call("ocp_bin_clusters")
    def buildNeighbor(self): self._call("ocp_build_neighbor")
    def pruneNeighbor(self): self._call("ocp_prune_neighbor")
    def updateSingleAtoms(self): self._call("ocp_update_single_atoms")
    def updateAtomsPbc(self): self._call("ocp_update_atoms_pbc")
    def updatePbc(self, first=False): self._call("ocp_update_pbc", int(first))
    def computeForce(self): self._call("ocp_compute_force")
    def initialIntegrate(self): self._call("ocp_initial_integrate")
    def finalIntegrate(self): self._call("ocp_final_integrate")
    def reneighbour(self): self._call("ocp_reneighbour")

    def thermo(self):
        T, P = C.c_double(), C.c_double()
        self._call("ocp_thermo", C.byref(T), C.byref(P))
        return T.value, P.value

    def run(self, nsteps):
        out = np.zeros(3 * (nsteps // max(1, self.geti("nstat")) + 3))
        n = self._call("ocp_run", nsteps, out.ctypes.data_as(C.c_void_p), len(out) // 3)
        return out[:3 * n].reshape(n, 3)

    def geti(self, k): return self._call("ocp_get_int", k.encode())
    def getr(self, k): return self._call("ocp_get_real", k.encode())

    def _arr(self, k, n, dtype):
        p = self._call("ocp_ptr", k.encode())
        if not p or n == 0:
            return np.zeros(0, dtype=dtype)
        ct = {np.float64: C.c_double, np.float32: C.c_float, np.int32: C.c_int}[dtype]
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), shape=(n,))

    def _clusters(self, name, n):
        p = self._call("ocp_ptr", name.encode())
        sz = self.lib.ocp_sizeof_cluster()
        raw = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_ubyte)), shape=(n * sz,)).reshape(n, sz)
        nat = raw[:, :4].copy().view(np.int32).reshape(n)
        off = 8 if self.dp else 4
        bb = raw[:, off:off + 6 * (8 if self.dp else 4)].copy().view(self.np_real).reshape(n, 6)
        return nat, bb

    def _atoms(self, what, n):
        names = {"x": "xyz", "v": ("vx", "vy", "vz")}[what]
        return np.stack([self._arr(c, n, self.np_real) for c in names], axis=1).copy()

    def tags(self):
        """original index of the atom now stored at each atom slot / each cluster slot (-1 = padding)"""
        n = self.geti("Nlocal")
        ncj, ngh = self.tiles()
        return self._arr("tag", n, np.int32).copy(), self._arr("cl_tag", (ncj + ngh) * self.N, np.int32).copy()

    def neigh_params(self):
        d = {k: self.geti(k) for k in ("nbinx", "nbiny", "mbinx", "mbiny", "mbins", "mbinxlo", "mbinylo", "nstencil")}
        d.update({k: self.getr(k) for k in ("binsizex", "binsizey", "bininvx", "bininvy", "cutneighsq", "xprd")})
        d["stencil"] = self._arr("stencil", d["nstencil"], np.int32).copy()
        return d

    def bin_nclusters(self):
        return self._arr("bin_nclusters", self.geti("mbins"), np.int32).copy()

    def icluster_bin(self):
        return self._arr("icluster_bin", self.geti("Nclusters_local"), np.int32).copy()


# ---------------------------------------------------------------------------------------------------
def _cp_structs(real):
    P = C.POINTER
    Parameter = _vl_structs(real)[0]

    class Cluster(C.Structure):
        _fields_ = [("natoms", C.c_int)] + [(n, real) for n in ("bbminx", "bbmaxx", "bbminy", "bbmaxy", "bbminz", "bbmaxz")]

    class Atom(C.Structure):  # clusterpair/atom.h:26-60
        _fields_ = [
            ("Natoms", C.c_int), ("Nlocal", C.c_int), ("Nghost", C.c_int), ("Nmax", C.c_int),
            ("Nclusters", C.c_int), ("Nclusters_local", C.c_int), ("Nclusters_ghost", C.c_int), ("Nclusters_max", C.c_int),
            ("x", P(real)), ("y", P(real)), ("z", P(real)), ("vx", P(real)), ("vy", P(real)), ("vz", P(real)),
            ("border_map", P(C.c_int)), ("type", P(C.c_int)), ("ntypes", C.c_int),
            ("epsilon", P(real)), ("sigma6", P(real)), ("cutforcesq", P(real)), ("cutneighsq", P(real)),
            ("PBCx", P(C.c_int)), ("PBCy", P(C.c_int)), ("PBCz", P(C.c_int)),
            ("cl_x", P(real)), ("cl_v", P(real)), ("cl_f", P(real)), ("cl_type", P(C.c_int)),
            ("iclusters", P(Cluster)), ("jclusters", P(Cluster)), ("icluster_bin", P(C.c_int)), ("dummy_cj", C.c_int),
            ("exclusion_filter", C.c_void_p), ("diagonal_4xn_j_minus_i", C.c_void_p), ("diagonal_2xnn_j_minus_i", C.c_void_p),
            ("masks_2xnn_hn", C.c_uint * 8), ("masks_2xnn_fn", C.c_uint * 8),
            ("masks_4xn_hn", C.c_uint * 16), ("masks_4xn_fn", C.c_uint * 16),
        ]

    class Neighbor(C.Structure):  # clusterpair/neighbor.h:31-40
        _fields_ = [("every", C.c_int), ("ncalls", C.c_int), ("maxneighs", C.c_int), ("numneigh", P(C.c_int)),
                    ("numneigh_masked", P(C.c_int)), ("half_neigh", C.c_int), ("neighbors", P(C.c_int)),
                    ("neighbors_imask", P(C.c_uint))]

    class Stats(C.Structure):  # clusterpair/stats.h:12-20
        _fields_ = [(n, C.c_longlong) for n in ("calculated_forces", "num_neighs", "force_iters", "atoms_within_cutoff",
                                                "atoms_outside_cutoff", "clusters_within_cutoff", "clusters_outside_cutoff")]

    return Parameter, Atom, Neighbor, Stats, Cluster


def refcp_available(variant):
    return os.path.exists(os.path.join(REFDIR, "libmdref_%s.so" % variant))


class RefCP(_Common):
    """The reference clusterpair build, driven operator by operator.
    variant: 'cp_dp_aos' / 'cp_sp_aos' (unmodified AVX-512 builds: 4x8, SIMD kernels with rcp14) or
    'cpref44_sp' / 'cpref44_dp' / 'cpref48_sp' / 'cpref48_dp' (scalar kernel computeForceLJRef, see Makefile)."""

    def __init__(self, variant):
        self.variant = variant
        self.dp = "dp" in variant
        self.N = 4 if "44" in variant else 8
        self.real = C.c_double if self.dp else C.c_float
        self.np_real = np.float64 if self.dp else np.float32
        path = os.path.join(REFDIR, "libmdref_%s.so" % variant)
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle ref-cp ref-cpref`)")
        self.lib = C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        self.Parameter, self.Atom, self.Neighbor, self.Stats, self.Cluster = _cp_structs(self.real)
        self.param, self.atom, self.neighbor, self.stats = self.Parameter(), self.Atom(), self.Neighbor(), self.Stats()
        self.lib.initParameter(C.byref(self.param))

    def _fp(self, name, restype, *argtypes):
        addr = C.c_void_p.in_dll(self.lib, name).value
        return C.CFUNCTYPE(restype, *argtypes)(addr)

    def configure(self, nx=32, ny=None, nz=None, half_neigh=0, reneigh_every=20, nstat=100, ntimes=200):
        p = self.param
        p.nx, p.ny, p.nz = nx, nx if ny is None else ny, nx if nz is None else nz
        p.half_neigh, p.reneigh_every, p.nstat, p.ntimes = half_neigh, reneigh_every, nstat, ntimes

    def setup(self, upto="all"):   # clusterpair/main.c:40-76
        L, p, a, n, s = self.lib, self.param, self.atom, self.neighbor, self.stats
        p.cutneigh = p.cutforce + p.skin    # clusterpair/main.c:216
        p.lattice = self.np_real((4.0 / float(p.rho)) ** (1.0 / 3.0))
        p.xprd = self.np_real(p.nx * self.np_real(p.lattice))
        p.yprd = self.np_real(p.ny * self.np_real(p.lattice))
        p.zprd = self.np_real(p.nz * self.np_real(p.lattice))
        L.initAtom(C.byref(a)); L.initForce(C.byref(p)); L.initPbc(C.byref(a)); L.initStats(C.byref(s))
        L.initNeighbor(C.byref(n), C.byref(p))
        L.createAtom(C.byref(a), C.byref(p))
        L.setupNeighbor(C.byref(p), C.byref(a))
        L.setupThermo(C.byref(p), a.Natoms)
        L.adjustThermo(C.byref(p), C.byref(a))
        if upto == "atoms":
            return
        self.buildClusters(); self.defineJClusters(); self.setupPbc(); self.binClusters(); self.buildNeighbor()

    def setup_from_files(self, param_file, input_file, upto="all", **over):
        """clusterpair/main.c with -p <param_file> -i <input_file>: readParameter, readAtom (sets the box), no adjustThermo"""
        L, p, a, n, s = self.lib, self.param, self.atom, self.neighbor, self.stats
        L.readParameter(C.byref(p), param_file.encode())
        for k, v in over.items():
            setattr(p, k, v)
        self._input = input_file.encode()
        p.input_file = self._input
        p.cutneigh = p.cutforce + p.skin
        p.lattice = self.np_real((4.0 / float(p.rho)) ** (1.0 / 3.0))
        L.initAtom(C.byref(a)); L.initForce(C.byref(p)); L.initPbc(C.byref(a)); L.initStats(C.byref(s))
        L.initNeighbor(C.byref(n), C.byref(p))
        L.readAtom(C.byref(a), C.byref(p))
        L.setupNeighbor(C.byref(p), C.byref(a))
        L.setupThermo(C.byref(p), a.Natoms)
        if upto == "atoms":
            return
        self.buildClusters(); self.defineJClusters(); self.setupPbc(); self.binClusters(); self.buildNeighbor()

    def buildClusters(self): self.lib.buildClusters(C.byref(self.atom))
    def defineJClusters(self): self.lib.defineJClusters(C.byref(self.atom))
    def setupPbc(self): self.lib.setupPbc(C.byref(self.atom), C.byref(self.param))
    def binClusters(self): self.lib.binClusters(C.byref(self.atom))
    def buildNeighbor(self): self.lib.buildNeighborCPU(C.byref(self.atom), C.byref(self.neighbor))
    def pruneNeighbor(self): self.lib.pruneNeighbor(C.byref(self.param), C.byref(self.atom), C.byref(self.neighbor))
    def updateSingleAtoms(self): self.lib.updateSingleAtoms(C.byref(self.atom))
    def updateAtomsPbc(self): self.lib.updateAtomsPbcCPU(C.byref(self.atom), C.byref(self.param), C.c_bool(False))
    def updatePbc(self, first=False): self.lib.updatePbcCPU(C.byref(self.atom), C.byref(self.param), C.c_bool(first))
    def initialIntegrate(self): self.lib.initialIntegrateCPU(C.byref(self.param), C.byref(self.atom))
    def finalIntegrate(self): self.lib.finalIntegrateCPU(C.byref(self.param), C.byref(self.atom))

    def computeForce(self):
        f = self._fp("computeForce", C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p)
        return f(C.addressof(self.param), C.addressof(self.atom), C.addressof(self.neighbor), C.addressof(self.stats))

    def reneighbour(self):   # clusterpair/main.c:78-93
        self.updateSingleAtoms(); self.updateAtomsPbc(); self.buildClusters(); self.defineJClusters()
        self.setupPbc(); self.binClusters(); self.buildNeighbor()

    def thermo(self):
        """T, P of common/thermo.c:55-66 recomputed from the atom arrays in the build's precision"""
        v = self._atoms("v", self.atom.Nlocal)
        p, r = self.param, self.np_real
        t = r(np.sum((v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1] + v[:, 2] * v[:, 2]) * r(p.mass), dtype=r))
        dof = self.atom.Natoms * 3 - 3
        T = t * (1.0 / dof)
        return float(T), float(T * dof * (1.0 / 3 / p.xprd / p.yprd / p.zprd))

    def geti(self, k):
        if k in ("Natoms", "Nlocal", "Nghost", "Nclusters", "Nclusters_local", "Nclusters_ghost", "Nclusters_max", "dummy_cj"):
            return getattr(self.atom, k)
        if k == "maxneighs":
            return self.neighbor.maxneighs
        if k in ("reneigh_every", "nstat", "half_neigh"):
            return getattr(self.param, k)
        return C.c_int.in_dll(self.lib, k).value   # file-scope statics of neighbor.c are exported symbols? (see neigh_params)

    def _arr(self, k, n, dtype):
        src = {"numneigh": self.neighbor.numneigh, "numneigh_masked": self.neighbor.numneigh_masked,
               "neighbors": self.neighbor.neighbors}.get(k)
        if src is None:
            src = getattr(self.atom, k)
        if n == 0:
            return np.zeros(0, dtype=dtype)
        return np.ctypeslib.as_array(src, shape=(n,))

    def _clusters(self, name, n):
        arr = getattr(self.atom, name)
        nat = np.array([arr[i].natoms for i in range(n)], np.int32)
        bb = np.array([[arr[i].bbminx, arr[i].bbmaxx, arr[i].bbminy, arr[i].bbmaxy, arr[i].bbminz, arr[i].bbmaxz]
                       for i in range(n)], self.np_real).reshape(n, 6)
        return nat, bb

    def _atoms(self, what, n):
        a = self.atom
        if what == "x":   # -DAOS: atom_x(i) = x[3i] (clusterpair/atom.h:66-70)
            return np.ctypeslib.as_array(a.x, shape=(n * 3,)).reshape(n, 3).copy()
        return np.stack([np.ctypeslib.as_array(q, shape=(n,)) for q in (a.vx, a.vy, a.vz)], axis=1).copy()

    def icluster_bin(self):
        return np.ctypeslib.as_array(self.atom.icluster_bin, shape=(self.atom.Nclusters_local,)).copy()
