"""md-bench_b200 -- host-side mirror of MD-Bench's operator interface over libmdb200 (C ABI,
include/mdb200.h), for tests and bench.py.

The method names follow the reference's function pointers / driver functions
(computeForce, buildNeighbor, initialIntegrate, finalIntegrate, updatePbc, updateAtomsPbc, setupPbc,
setupNeighbor, reneighbour, computeThermo, adjustThermo; reference src/verletlist/{force,neighbor,
integrate,pbc}.h and main.c) so parity tests read like the reference's own driver.

This package never touches the parity checker under the repo's oracle directory and has no CPU path: constructing a Simulation
without a CUDA device raises MdbError.  (The directory name contains a hyphen; import it with
importlib.import_module("md-bench_b200") -- see tests/conftest.py.)
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

HERE = os.path.dirname(os.path.abspath(__file__))
SP, DP = 1, 2
AOS, SOA = 0, 1
FF_LJ, FF_EAM = 0, 1


class MdbError(RuntimeError):
    pass


class Params(C.Structure):
    """struct mdb_params (include/mdb200.h) <- reference Parameter, common/parameter.h:27-61"""
    _fields_ = [
        ("precision", C.c_int), ("layout", C.c_int), ("force_field", C.c_int),
        ("epsilon", C.c_double), ("sigma", C.c_double), ("temp", C.c_double), ("rho", C.c_double),
        ("mass", C.c_double),
        ("ntypes", C.c_int), ("ntimes", C.c_int), ("nstat", C.c_int), ("reneigh_every", C.c_int),
        ("half_neigh", C.c_int),
        ("dt", C.c_double), ("skin", C.c_double), ("cutforce", C.c_double),
        ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
        ("pbc_x", C.c_int), ("pbc_y", C.c_int), ("pbc_z", C.c_int),
        ("from_input", C.c_int),
        ("xlo", C.c_double), ("xhi", C.c_double), ("ylo", C.c_double), ("yhi", C.c_double),
        ("zlo", C.c_double), ("zhi", C.c_double),
    ]


EXPORTS = [
    "mdb_abi_version", "mdb_last_error", "mdb_default_params", "mdb_create", "mdb_destroy",
    "mdb_setOption", "mdb_setStream", "mdb_sync", "mdb_createAtom", "mdb_setAtoms", "mdb_setAtomsDevice",
    "mdb_getAtoms", "mdb_getTypes", "mdb_getCounts", "mdb_saveState", "mdb_restoreState", "mdb_setupThermo",
    "mdb_adjustThermo", "mdb_computeThermo", "mdb_setupNeighbor", "mdb_setupPbc", "mdb_updatePbc",
    "mdb_updateAtomsPbc", "mdb_buildNeighbor", "mdb_computeForce", "mdb_computeForceLJFullNeigh",
    "mdb_computeForceLJHalfNeigh", "mdb_computeForceEam", "mdb_initialIntegrate",
    "mdb_finalIntegrate", "mdb_setup", "mdb_reneighbour", "mdb_run", "mdb_setTiming",
    "mdb_getKernelStats", "mdb_resetKernelStats", "mdb_setEam", "mdb_setEamSplines",
    "mdb_getEamSplines", "mdb_getNeighbors", "mdb_getGhostMap", "mdb_getNeighborParams",
    "mdb_getStencil", "mdb_getBinCounts", "mdb_getEamFp", "mdb_countPairs", "mdb_measureFmaPeak", "mdb_stubNeighbors",
    "mdb_dd_uniqueIdBytes", "mdb_dd_getUniqueId", "mdb_dd_plan", "mdb_dd_schedule", "mdb_dd_create", "mdb_dd_create_cp",
    "mdb_dd_destroy",
    "mdb_dd_setStream", "mdb_dd_sync", "mdb_dd_createAtom", "mdb_dd_setAtoms", "mdb_dd_setEam", "mdb_dd_setup",
    "mdb_dd_reneighbour", "mdb_dd_run", "mdb_dd_computeThermo", "mdb_dd_getCounts", "mdb_dd_getAtoms",
    "mdb_dd_getNeighborTags", "mdb_dd_saveState", "mdb_dd_restoreState", "mdb_dd_setOption",
    "mdb_dd_setTiming", "mdb_dd_getKernelStats", "mdb_dd_resetKernelStats",
    "mdb_cp_create", "mdb_cp_destroy", "mdb_cp_setStream", "mdb_cp_sync", "mdb_cp_setOption", "mdb_cp_createAtom",
    "mdb_cp_setAtoms", "mdb_cp_getAtoms", "mdb_cp_getCounts", "mdb_cp_setupThermo", "mdb_cp_adjustThermo",
    "mdb_cp_computeThermo", "mdb_cp_setupNeighbor", "mdb_cp_buildClusters", "mdb_cp_defineJClusters", "mdb_cp_setupPbc",
    "mdb_cp_binClusters", "mdb_cp_buildNeighbor", "mdb_cp_pruneNeighbor", "mdb_cp_updateSingleAtoms",
    "mdb_cp_updateAtomsPbc", "mdb_cp_updatePbc", "mdb_cp_computeForce", "mdb_cp_initialIntegrate",
    "mdb_cp_finalIntegrate", "mdb_cp_setup", "mdb_cp_reneighbour", "mdb_cp_run", "mdb_cp_saveState",
    "mdb_cp_restoreState", "mdb_cp_setTiming", "mdb_cp_getKernelStats", "mdb_cp_resetKernelStats", "mdb_cp_countPairs",
    "mdb_cp_getClusters", "mdb_cp_getClusterData", "mdb_cp_getClusterTags", "mdb_cp_getClusterBins", "mdb_cp_getLists",
    "mdb_cp_getGhostMap", "mdb_cp_getNeighborParams", "mdb_cp_stub",
]

_lib = None


def lib_path():
    return os.path.join(HERE, "libmdb200.so")


def load_library(build=True):
    """dlopen libmdb200.so (building it first if sources are newer). Fails loudly if missing."""
    global _lib
    if _lib is not None:
        return _lib
    if build and _build.needs_build():
        _build.build()
    if not os.path.exists(lib_path()):
        raise MdbError("libmdb200.so is missing: run `python md-bench_b200/build.py` (no CPU fallback exists)")
    L = C.CDLL(lib_path())
    L.mdb_last_error.restype = C.c_char_p
    L.mdb_create.restype = C.c_void_p
    L.mdb_create.argtypes = [C.POINTER(Params), C.c_int]
    L.mdb_createAtom.restype = C.c_longlong
    for f in ("mdb_computeForce", "mdb_computeForceLJFullNeigh", "mdb_computeForceLJHalfNeigh",
              "mdb_computeForceEam"):
        getattr(L, f).restype = C.c_double
        getattr(L, f).argtypes = [C.c_void_p]
    L.mdb_dd_create.restype = C.c_void_p
    L.mdb_dd_create.argtypes = [C.POINTER(Params), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    L.mdb_dd_create_cp.restype = C.c_void_p
    L.mdb_dd_create_cp.argtypes = [C.POINTER(Params), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    L.mdb_dd_createAtom.restype = C.c_longlong
    L.mdb_dd_createAtom.argtypes = [C.c_void_p]
    L.mdb_cp_create.restype = C.c_void_p
    L.mdb_cp_create.argtypes = [C.POINTER(Params), C.c_int, C.c_int]
    L.mdb_cp_createAtom.restype = C.c_longlong
    L.mdb_cp_createAtom.argtypes = [C.c_void_p]
    L.mdb_cp_computeForce.restype = C.c_double
    L.mdb_cp_computeForce.argtypes = [C.c_void_p]
    _lib = L
    return L


def default_params(**kw):
    L = load_library()
    p = Params()
    L.mdb_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


def measure_fma_peak(precision=DP, device=0):
    """FP32/FP64 vector FMA peak of `device` in TFLOP/s (micro-benchmark in csrc/peaks.cu)"""
    L = load_library()
    t = C.c_double()
    if L.mdb_measureFmaPeak(precision, device, C.byref(t)) != 0:
        raise MdbError("mdb_measureFmaPeak failed")
    return t.value


def _vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Simulation:
    """One simulation domain on one GPU (opaque mdb_ctx)."""

    def __init__(self, params=None, device=0, **kw):
        self.L = load_library()
        self.params = params if params is not None else default_params(**kw)
        self.dp = self.params.precision == DP
        self.aos = self.params.layout == AOS
        self.np_real = np.float64 if self.dp else np.float32
        h = self.L.mdb_create(C.byref(self.params), device)
        if not h:
            raise MdbError(self.L.mdb_last_error().decode())
        self.h = C.c_void_p(h)

    def close(self):
        if getattr(self, "h", None):
            self.L.mdb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise MdbError(self.L.mdb_last_error().decode())

    # ---- atoms ----
    def createAtom(self):
        n = self.L.mdb_createAtom(self.h)
        if n < 0:
            raise MdbError(self.L.mdb_last_error().decode())
        return n

    def _pack(self, a):
        """(n,3) array -> host buffers in this ctx's layout"""
        a = np.ascontiguousarray(a, dtype=self.np_real)
        if self.aos:
            return [a, None, None]
        return [np.ascontiguousarray(a[:, k]) for k in range(3)]

    def setAtoms(self, x, v=None, type=None):
        n = x.shape[0]
        bx = self._pack(x)
        bv = self._pack(v) if v is not None else [None] * 3
        t = None if type is None else np.ascontiguousarray(type, dtype=np.int32)
        self._ck(self.L.mdb_setAtoms(self.h, C.c_longlong(n), *[_vp(b) for b in bx],
                                     *[_vp(b) for b in bv], _vp(t)))

    def setAtomsDevice(self, n, ptrs_x, ptrs_v=(None, None, None), type_ptr=None):
        cv = lambda p: None if p is None else C.c_void_p(p)
        self._ck(self.L.mdb_setAtomsDevice(self.h, C.c_longlong(n), *[cv(p) for p in ptrs_x],
                                           *[cv(p) for p in ptrs_v], cv(type_ptr)))

    def types(self):
        t = np.empty(self.counts()["Nlocal"], np.int32)
        self._ck(self.L.mdb_getTypes(self.h, _vp(t)))
        return t

    def counts(self):
        v = [C.c_longlong() for _ in range(4)]
        mn = C.c_int()
        self._ck(self.L.mdb_getCounts(self.h, *[C.byref(q) for q in v], C.byref(mn)))
        return dict(Natoms=v[0].value, Nlocal=v[1].value, Nghost=v[2].value, Nmax=v[3].value,
                    maxneighs=mn.value)

    def get(self, what, ghosts=False, out=None):
        """x / v / f as an (n,3) array in the reference's index order.  `out` (AOS layout only):
        a preallocated C-contiguous (n,3) array, e.g. a view of pinned host memory."""
        c = self.counts()
        n = c["Nlocal"] + (c["Nghost"] if (ghosts and what == "x") else 0)
        if self.aos:
            a = out if out is not None else np.empty((n, 3), dtype=self.np_real)
            assert a.shape == (n, 3) and a.dtype == self.np_real and a.flags["C_CONTIGUOUS"]
            self._ck(self.L.mdb_getAtoms(self.h, ord(what), int(ghosts), _vp(a), None, None))
            return a
        cols = [np.empty(n, dtype=self.np_real) for _ in range(3)]
        self._ck(self.L.mdb_getAtoms(self.h, ord(what), int(ghosts), *[_vp(q) for q in cols]))
        return np.stack(cols, axis=1)

    # rows of the neighbor list come back in the reference's ORDER (stencil order x ascending index)
    # only when the internal spatial sort is off; as sorted index SETS they always agree
    row_order_exact = True

    def setOption(self, name, value):
        self._ck(self.L.mdb_setOption(self.h, name.encode(), C.c_double(value)))
        if name == "sort_atoms":
            self.row_order_exact = not value

    def saveState(self): self._ck(self.L.mdb_saveState(self.h))
    def restoreState(self): self._ck(self.L.mdb_restoreState(self.h))
    def setStream(self, stream_ptr): self._ck(self.L.mdb_setStream(self.h, C.c_void_p(stream_ptr)))
    def sync(self): self._ck(self.L.mdb_sync(self.h))

    # ---- operators (names of the reference's function pointers) ----
    def setupNeighbor(self): self._ck(self.L.mdb_setupNeighbor(self.h))
    def setupThermo(self): self._ck(self.L.mdb_setupThermo(self.h))
    def adjustThermo(self): self._ck(self.L.mdb_adjustThermo(self.h))
    def setupPbc(self): self._ck(self.L.mdb_setupPbc(self.h))
    def updatePbc(self, reneigh=False): self._ck(self.L.mdb_updatePbc(self.h, int(reneigh)))
    def updateAtomsPbc(self, reneigh=True): self._ck(self.L.mdb_updateAtomsPbc(self.h, int(reneigh)))
    def buildNeighbor(self): self._ck(self.L.mdb_buildNeighbor(self.h))
    def initialIntegrate(self, reneigh=False): self._ck(self.L.mdb_initialIntegrate(self.h, int(reneigh)))
    def finalIntegrate(self, reneigh=False): self._ck(self.L.mdb_finalIntegrate(self.h, int(reneigh)))
    def reneighbour(self): self._ck(self.L.mdb_reneighbour(self.h))
    def setup(self, adjust=True): self._ck(self.L.mdb_setup(self.h, int(adjust)))

    def _force(self, fn):
        t = fn(self.h)
        if t < 0:
            raise MdbError(self.L.mdb_last_error().decode())
        return t

    def stubNeighbors(self, pattern, nneighs=76, nreps=1, seed=12345):
        """synthetic list of the reference's kernel micro-benchmark (main-stub.c): pattern 'seq' | 'fix' | 'rand'"""
        self._ck(self.L.mdb_stubNeighbors(self.h, {"seq": 0, "fix": 1, "rand": 2, "local": 3, "localbank": 4}[pattern], nneighs, nreps, C.c_uint(seed)))

    def computeForce(self): return self._force(self.L.mdb_computeForce)
    def computeForceLJFullNeigh(self): return self._force(self.L.mdb_computeForceLJFullNeigh)
    def computeForceLJHalfNeigh(self): return self._force(self.L.mdb_computeForceLJHalfNeigh)
    def computeForceEam(self): return self._force(self.L.mdb_computeForceEam)

    def computeThermo(self):
        T, P = C.c_double(), C.c_double()
        self._ck(self.L.mdb_computeThermo(self.h, C.byref(T), C.byref(P)))
        return T.value, P.value

    thermo = computeThermo

    def step(self, n):
        """one iteration of the reference time loop (verletlist/main.c:258-273), operator by operator"""
        reneigh = (n + 1) % self.params.reneigh_every == 0
        self.initialIntegrate(reneigh)
        if reneigh:
            self.reneighbour()
        else:
            self.updatePbc(False)
        self.computeForce()
        self.finalIntegrate(reneigh)
        return reneigh

    def run(self, nsteps):
        """mdb_run: whole time loop on the device. Returns (thermo records (k,3), timers dict)."""
        nstat = max(1, self.params.nstat)
        out = np.zeros(3 * (nsteps // nstat + 4))
        nrec = C.c_int()
        tm = (C.c_double * 3)()
        self._ck(self.L.mdb_run(self.h, nsteps, _vp(out), len(out) // 3, C.byref(nrec), tm))
        return out[:3 * nrec.value].reshape(-1, 3), dict(TOTAL=tm[0], FORCE=tm[1], NEIGH=tm[2])

    def setTiming(self, on): self._ck(self.L.mdb_setTiming(self.h, int(on)))

    def kernelStats(self):
        fm, nm = C.c_double(), C.c_double()
        fl, nl, tl = C.c_longlong(), C.c_longlong(), C.c_longlong()
        self._ck(self.L.mdb_getKernelStats(self.h, C.byref(fm), C.byref(fl), C.byref(nm), C.byref(nl),
                                           C.byref(tl)))
        return dict(force_ms=fm.value, force_launches=fl.value, neigh_ms=nm.value,
                    neigh_launches=nl.value, launches=tl.value)

    def resetKernelStats(self): self._ck(self.L.mdb_resetKernelStats(self.h))

    # ---- EAM ----
    def setEam(self, nrho, drho, nr, dr, cut, mass, frho, zr, rhor):
        a = [np.ascontiguousarray(t, dtype=np.float64) for t in (frho, zr, rhor)]
        d = C.c_double
        self._ck(self.L.mdb_setEam(self.h, nrho, d(drho), nr, d(dr), d(cut), d(mass), *[_vp(t) for t in a]))

    def setEamSplines(self, nr, nrho, nr_tot, nrho_tot, rdr, rdrho, rhor, frho, z2r):
        a = [np.ascontiguousarray(t, dtype=self.np_real) for t in (rhor, frho, z2r)]
        self._ck(self.L.mdb_setEamSplines(self.h, nr, nrho, nr_tot, nrho_tot, C.c_double(rdr),
                                          C.c_double(rdrho), *[_vp(t) for t in a]))

    def getEamSplines(self):
        iv = [C.c_int() for _ in range(4)]
        dv = [C.c_double() for _ in range(2)]
        args = [C.byref(q) for q in iv] + [C.byref(q) for q in dv]
        self._ck(self.L.mdb_getEamSplines(self.h, *args, None, None, None))
        nr, nrho, nr_tot, nrho_tot = [q.value for q in iv]
        rh, z2 = np.empty(nr_tot, self.np_real), np.empty(nr_tot, self.np_real)
        fr = np.empty(nrho_tot, self.np_real)
        self._ck(self.L.mdb_getEamSplines(self.h, *args, _vp(rh), _vp(fr), _vp(z2)))
        return dict(nr=nr, nrho=nrho, nr_tot=nr_tot, nrho_tot=nrho_tot, rdr=dv[0].value,
                    rdrho=dv[1].value, rhor_spline=rh, frho_spline=fr, z2r_spline=z2)

    def getEamFp(self, ghosts=False):
        c = self.counts()
        n = c["Nlocal"] + (c["Nghost"] if ghosts else 0)
        a = np.empty(n, self.np_real)
        self._ck(self.L.mdb_getEamFp(self.h, _vp(a), int(ghosts)))
        return a

    # ---- parity accessors ----
    def neighbors(self):
        c = self.counts()
        nn = np.empty(c["Nlocal"], np.int32)
        nb = np.empty((c["Nlocal"], c["maxneighs"]), np.int32)
        self._ck(self.L.mdb_getNeighbors(self.h, _vp(nn), _vp(nb), c["maxneighs"]))
        return nn, nb

    def numneigh(self):
        c = self.counts()
        nn = np.empty(c["Nlocal"], np.int32)
        self._ck(self.L.mdb_getNeighbors(self.h, _vp(nn), None, 0))
        return nn

    def sorted_neighbor_sets(self):
        nn, nb = self.neighbors()
        return nn, [np.sort(nb[i, :nn[i]]) for i in range(len(nn))]

    def ghostMap(self):
        ng = self.counts()["Nghost"]
        a = [np.empty(ng, np.int32) for _ in range(4)]
        self._ck(self.L.mdb_getGhostMap(self.h, *[_vp(q) for q in a]))
        return dict(border_map=a[0], PBCx=a[1], PBCy=a[2], PBCz=a[3])

    def neighborParams(self):
        iv = (C.c_int * 12)()
        rv = (C.c_double * 14)()
        self._ck(self.L.mdb_getNeighborParams(self.h, iv, rv))
        ik = ["nbinx", "nbiny", "nbinz", "mbinx", "mbiny", "mbinz", "mbinxlo", "mbinylo", "mbinzlo",
              "mbins", "nstencil", "max_bin_count"]
        rk = ["bininvx", "bininvy", "bininvz", "binsizex", "binsizey", "binsizez", "cutneighsq",
              "cutneigh", "xprd", "yprd", "zprd", "lattice", "dtforce", "cutforce"]
        d = dict(zip(ik, list(iv)))
        d.update(zip(rk, list(rv)))
        st = np.empty(d["nstencil"], np.int32)
        self._ck(self.L.mdb_getStencil(self.h, _vp(st)))
        d["stencil"] = st
        return d

    def binCounts(self):
        m = self.neighborParams()["mbins"]
        a = np.empty(m, np.int32)
        self._ck(self.L.mdb_getBinCounts(self.h, _vp(a)))
        return a

    def countPairs(self):
        a, b = C.c_longlong(), C.c_longlong()
        self._ck(self.L.mdb_countPairs(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value


# ---- multi-GPU: spatial decomposition -----------------------------------------------------------
def dd_plan(grid, brick, send=True, nprocs=1, periodic=(1, 1, 1)):
    """host-only: slots of `brick` as (direction, peer brick, owning process) triples (mdb_dd_plan)"""
    L = load_library()
    d, p, o = (C.c_int * 26)(), (C.c_int * 26)(), (C.c_int * 26)()
    n = L.mdb_dd_plan(grid[0], grid[1], grid[2], periodic[0], periodic[1], periodic[2], nprocs, brick, int(send), d, p, o)
    if n < 0:
        raise MdbError(L.mdb_last_error().decode())
    return [(d[k], p[k], o[k]) for k in range(n)]


def dd_schedule(grid, nprocs, proc, cnt, periodic=(1, 1, 1)):
    """host-only: transfers of one exchange as `proc` executes them (mdb_dd_schedule); cnt[nbricks, 26]"""
    L = load_library()
    cnt = np.ascontiguousarray(cnt, dtype=np.int32)
    nb = grid[0] * grid[1] * grid[2]
    assert cnt.shape == (nb, 26)
    ops = np.zeros((nb * 27, 7), np.int32)
    n = L.mdb_dd_schedule(grid[0], grid[1], grid[2], periodic[0], periodic[1], periodic[2], nprocs, proc,
                          _vp(cnt), len(ops), _vp(ops))
    if n < 0:
        raise MdbError(L.mdb_last_error().decode())
    keys = ("kind", "src", "dst", "src_start", "dst_start", "len", "peer_proc")
    return [dict(zip(keys, map(int, ops[k]))) for k in range(n)]


def dd_grid(nprocs):
    """brick grid used by bench.py for N GPUs: 1 -> 1x1x1, 2 -> 2x1x1, 4 -> 2x2x1, 8 -> 2x2x2"""
    g = [1, 1, 1]
    a = 0
    while g[0] * g[1] * g[2] < nprocs:
        g[a % 3] *= 2
        a += 1
    if g[0] * g[1] * g[2] != nprocs:
        raise ValueError("number of processes must be a power of two")
    return tuple(g)


def dd_unique_id():
    L = load_library()
    buf = (C.c_char * L.mdb_dd_uniqueIdBytes())()
    if L.mdb_dd_getUniqueId(buf) != 0:
        raise MdbError(L.mdb_last_error().decode())
    return bytes(buf)


class Decomposition:
    """The bricks of a spatially decomposed box owned by this process (opaque mdb_dd).  `params`
    describe the WHOLE box; same driver-level method names as Simulation.  cluster_n = 4 or 8 selects the CLUSTERPAIR scheme
    (mdb_dd_create_cp: ghost clusters from the neighbor bricks), 0 the verletlist scheme."""

    def __init__(self, params, grid, nprocs=1, proc=0, nccl_id=None, device=0, cluster_n=0):
        self.L = load_library()
        self.params = params
        self.dp = params.precision == DP
        self.np_real = np.float64 if self.dp else np.float32
        self.grid = tuple(grid)
        idbuf = None if nccl_id is None else C.create_string_buffer(nccl_id, len(nccl_id))
        idp = None if idbuf is None else C.cast(idbuf, C.c_void_p)
        if cluster_n:
            h = self.L.mdb_dd_create_cp(C.byref(params), cluster_n, grid[0], grid[1], grid[2], nprocs, proc, idp, device)
        else:
            h = self.L.mdb_dd_create(C.byref(params), grid[0], grid[1], grid[2], nprocs, proc, idp, device)
        if not h:
            raise MdbError(self.L.mdb_last_error().decode())
        self.h = C.c_void_p(h)

    def close(self):
        if getattr(self, "h", None):
            self.L.mdb_dd_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise MdbError(self.L.mdb_last_error().decode())

    def createAtom(self):
        n = self.L.mdb_dd_createAtom(self.h)
        if n < 0:
            raise MdbError(self.L.mdb_last_error().decode())
        return n

    def setAtoms(self, tags, x, v=None):
        """x, v: (3, n) arrays (one row per coordinate, e.g. views of pinned memory) in the global frame"""
        tags = np.ascontiguousarray(tags, dtype=np.int32)
        n = len(tags)
        assert x.shape == (3, n) and x.dtype == self.np_real and x.flags["C_CONTIGUOUS"]
        cols = [x[k] for k in range(3)] + ([v[k] for k in range(3)] if v is not None else [None] * 3)
        self._ck(self.L.mdb_dd_setAtoms(self.h, C.c_longlong(n), _vp(tags), *[_vp(q) for q in cols]))

    def get_into(self, what, tags, out):
        """like get(), into preallocated (3, n) C-contiguous `out` and (n,) int32 `tags` (may be None)"""
        self._ck(self.L.mdb_dd_getAtoms(self.h, ord(what), _vp(tags), *[_vp(out[k]) for k in range(3)]))

    def setEam(self, nrho, drho, nr, dr, cut, mass, frho, zr, rhor):
        a = [np.ascontiguousarray(t, dtype=np.float64) for t in (frho, zr, rhor)]
        d = C.c_double
        self._ck(self.L.mdb_dd_setEam(self.h, nrho, d(drho), nr, d(dr), d(cut), d(mass), *[_vp(t) for t in a]))

    def setStream(self, stream_ptr): self._ck(self.L.mdb_dd_setStream(self.h, C.c_void_p(stream_ptr)))
    def sync(self): self._ck(self.L.mdb_dd_sync(self.h))
    def setup(self, adjust=True): self._ck(self.L.mdb_dd_setup(self.h, int(adjust)))
    def reneighbour(self): self._ck(self.L.mdb_dd_reneighbour(self.h))
    def saveState(self): self._ck(self.L.mdb_dd_saveState(self.h))
    def restoreState(self): self._ck(self.L.mdb_dd_restoreState(self.h))
    def setOption(self, name, value): self._ck(self.L.mdb_dd_setOption(self.h, name.encode(), C.c_double(value)))
    def setTiming(self, on): self._ck(self.L.mdb_dd_setTiming(self.h, int(on)))
    def resetKernelStats(self): self._ck(self.L.mdb_dd_resetKernelStats(self.h))

    def computeThermo(self):
        T, P = C.c_double(), C.c_double()
        self._ck(self.L.mdb_dd_computeThermo(self.h, C.byref(T), C.byref(P)))
        return T.value, P.value

    def run(self, nsteps):
        nstat = max(1, self.params.nstat)
        out = np.zeros(3 * (nsteps // nstat + 4))
        nrec = C.c_int()
        tm = (C.c_double * 3)()
        self._ck(self.L.mdb_dd_run(self.h, nsteps, _vp(out), len(out) // 3, C.byref(nrec), tm))
        return out[:3 * nrec.value].reshape(-1, 3), dict(TOTAL=tm[0], FORCE=tm[1], NEIGH=tm[2])

    def counts(self):
        v = (C.c_longlong * 5)()
        self._ck(self.L.mdb_dd_getCounts(self.h, v))
        return dict(Natoms=v[0], Nlocal=v[1], Nghost=v[2], maxneighs=int(v[3]), bricks=int(v[4]))

    def get(self, what):
        """(tags, (n,3) array) of this process's local atoms; positions in the global frame"""
        n = self.counts()["Nlocal"]
        tags = np.empty(n, np.int32)
        cols = [np.empty(n, self.np_real) for _ in range(3)]
        self._ck(self.L.mdb_dd_getAtoms(self.h, ord(what), _vp(tags), *[_vp(q) for q in cols]))
        return tags, np.stack(cols, axis=1)

    def neighborTags(self):
        c = self.counts()
        n, st = c["Nlocal"], c["maxneighs"]
        tags, nn = np.empty(n, np.int32), np.empty(n, np.int32)
        rows = np.empty((n, st), np.int32)
        self._ck(self.L.mdb_dd_getNeighborTags(self.h, _vp(tags), _vp(nn), _vp(rows), st))
        return tags, nn, rows

    def kernelStats(self):
        fm, nm, cm = C.c_double(), C.c_double(), C.c_double()
        fl, nl, tl = C.c_longlong(), C.c_longlong(), C.c_longlong()
        self._ck(self.L.mdb_dd_getKernelStats(self.h, C.byref(fm), C.byref(fl), C.byref(nm), C.byref(nl),
                                              C.byref(tl), C.byref(cm)))
        return dict(force_ms=fm.value, force_launches=fl.value, neigh_ms=nm.value, neigh_launches=nl.value,
                    launches=tl.value, comm_ms=cm.value)


class ClusterSimulation:
    """One CLUSTERPAIR simulation domain on one GPU (opaque mdb_cp): the reference's OPT_SCHEME=clusterpair with
    M = 4 and cluster_n = 4 | 8 (src/clusterpair/).  Method names are the reference driver's
    (clusterpair/main.c:40-93, neighbor.h:42-50); the accessors return what the reference keeps in
    Atom / Neighbor (cluster tiles as (tiles, 3, N) arrays)."""

    M = 4

    def __init__(self, params=None, cluster_n=4, device=0, **kw):
        self.L = load_library()
        self.params = params if params is not None else default_params(**kw)
        self.dp = self.params.precision == DP
        self.aos = self.params.layout == AOS
        self.np_real = np.float64 if self.dp else np.float32
        self.N = cluster_n
        h = self.L.mdb_cp_create(C.byref(self.params), cluster_n, device)
        if not h:
            raise MdbError(self.L.mdb_last_error().decode())
        self.h = C.c_void_p(h)

    def close(self):
        if getattr(self, "h", None):
            self.L.mdb_cp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise MdbError(self.L.mdb_last_error().decode())

    # ---- atoms ----
    def createAtom(self):
        n = self.L.mdb_cp_createAtom(self.h)
        if n < 0:
            raise MdbError(self.L.mdb_last_error().decode())
        return n

    def setAtoms(self, x, v=None):
        """x, v: (n,3) arrays; positions go over in this ctx's layout, velocities always SoA"""
        n = x.shape[0]
        x = np.ascontiguousarray(x, dtype=self.np_real)
        bx = [x, None, None] if self.aos else [np.ascontiguousarray(x[:, k]) for k in range(3)]
        bv = [None] * 3 if v is None else [np.ascontiguousarray(np.asarray(v, dtype=self.np_real)[:, k]) for k in range(3)]
        self._ck(self.L.mdb_cp_setAtoms(self.h, C.c_longlong(n), *[_vp(b) for b in bx], *[_vp(b) for b in bv]))

    set_atoms = setAtoms

    def counts(self):
        v = (C.c_longlong * 8)()
        self._ck(self.L.mdb_cp_getCounts(self.h, v))
        k = ("Natoms", "Nlocal", "Nghost", "Nclusters_local", "Nclusters_ghost", "dummy_cj", "maxneighs", "ncj")
        return dict(zip(k, [int(q) for q in v]))

    def geti(self, k):
        if k in ("reneigh_every", "nstat", "half_neigh"):
            return getattr(self.params, k)
        return self.counts()[k]

    def atoms(self, what, tags=False):
        n = self.counts()["Nlocal"]
        t = np.empty(n, np.int32) if tags else None
        if what == "x" and self.aos:
            a = np.empty((n, 3), self.np_real)
            self._ck(self.L.mdb_cp_getAtoms(self.h, ord("x"), _vp(a), None, None, _vp(t)))
        else:
            cols = [np.empty(n, self.np_real) for _ in range(3)]
            self._ck(self.L.mdb_cp_getAtoms(self.h, ord(what), *[_vp(q) for q in cols], _vp(t)))
            a = np.stack(cols, axis=1)
        return (a, t) if tags else a

    def get(self, what, out=None):
        """x / v of the atom arrays as (n,3); `out`: preallocated C-contiguous (n,3) array for 'x' in the AOS layout
        (e.g. a view of pinned host memory), or a (3,n) array for SoA data"""
        n = self.counts()["Nlocal"]
        if what == "x" and self.aos:
            a = out if out is not None else np.empty((n, 3), self.np_real)
            assert a.shape == (n, 3) and a.dtype == self.np_real and a.flags["C_CONTIGUOUS"]
            self._ck(self.L.mdb_cp_getAtoms(self.h, ord("x"), _vp(a), None, None, None))
            return a
        a = out if out is not None else np.empty((3, n), self.np_real)
        assert a.shape == (3, n) and a.dtype == self.np_real and a.flags["C_CONTIGUOUS"]
        self._ck(self.L.mdb_cp_getAtoms(self.h, ord(what), _vp(a[0]), _vp(a[1]), _vp(a[2]), None))
        return a

    def setAtomsRaw(self, x_aos, v_soa):
        """x: (n,3) C-contiguous (AOS layout ctx), v: (3,n) C-contiguous -- no host-side repacking (bench e2e leg)"""
        n = x_aos.shape[0]
        self._ck(self.L.mdb_cp_setAtoms(self.h, C.c_longlong(n), _vp(x_aos), None, None, _vp(v_soa[0]), _vp(v_soa[1]),
                                        _vp(v_soa[2])))

    def setOption(self, name, value): self._ck(self.L.mdb_cp_setOption(self.h, name.encode(), C.c_double(value)))
    def saveState(self): self._ck(self.L.mdb_cp_saveState(self.h))
    def restoreState(self): self._ck(self.L.mdb_cp_restoreState(self.h))
    def setStream(self, stream_ptr): self._ck(self.L.mdb_cp_setStream(self.h, C.c_void_p(stream_ptr)))
    def sync(self): self._ck(self.L.mdb_cp_sync(self.h))

    # ---- operators ----
    def setupNeighbor(self): self._ck(self.L.mdb_cp_setupNeighbor(self.h))
    def setupThermo(self): self._ck(self.L.mdb_cp_setupThermo(self.h))
    def adjustThermo(self): self._ck(self.L.mdb_cp_adjustThermo(self.h))
    def buildClusters(self): self._ck(self.L.mdb_cp_buildClusters(self.h))
    def defineJClusters(self): self._ck(self.L.mdb_cp_defineJClusters(self.h))
    def setupPbc(self): self._ck(self.L.mdb_cp_setupPbc(self.h))
    def binClusters(self): self._ck(self.L.mdb_cp_binClusters(self.h))
    def buildNeighbor(self): self._ck(self.L.mdb_cp_buildNeighbor(self.h))
    def pruneNeighbor(self): self._ck(self.L.mdb_cp_pruneNeighbor(self.h))
    def updateSingleAtoms(self): self._ck(self.L.mdb_cp_updateSingleAtoms(self.h))
    def updateAtomsPbc(self): self._ck(self.L.mdb_cp_updateAtomsPbc(self.h))
    def updatePbc(self, first=False): self._ck(self.L.mdb_cp_updatePbc(self.h, int(first)))
    def initialIntegrate(self): self._ck(self.L.mdb_cp_initialIntegrate(self.h))
    def finalIntegrate(self): self._ck(self.L.mdb_cp_finalIntegrate(self.h))
    def reneighbour(self): self._ck(self.L.mdb_cp_reneighbour(self.h))
    def setup(self, adjust=False): self._ck(self.L.mdb_cp_setup(self.h, int(adjust)))

    def computeForce(self):
        t = self.L.mdb_cp_computeForce(self.h)
        if t < 0:
            raise MdbError(self.L.mdb_last_error().decode())
        return t

    def thermo(self):
        T, P = C.c_double(), C.c_double()
        self._ck(self.L.mdb_cp_computeThermo(self.h, C.byref(T), C.byref(P)))
        return T.value, P.value

    computeThermo = thermo

    def step(self, n):
        """one iteration of the reference loop, clusterpair/main.c:246-266 (operator by operator)"""
        reneigh = (n + 1) % self.params.reneigh_every == 0
        self.initialIntegrate()
        if reneigh:
            self.reneighbour()
        else:
            self.updatePbc(False)
        self.computeForce()
        self.finalIntegrate()
        return reneigh

    def run(self, nsteps):
        nstat = max(1, self.params.nstat)
        out = np.zeros(3 * (nsteps // nstat + 4))
        nrec = C.c_int()
        tm = (C.c_double * 3)()
        self._ck(self.L.mdb_cp_run(self.h, nsteps, _vp(out), len(out) // 3, C.byref(nrec), tm))
        return out[:3 * nrec.value].reshape(-1, 3), dict(TOTAL=tm[0], FORCE=tm[1], NEIGH=tm[2])

    def setTiming(self, on): self._ck(self.L.mdb_cp_setTiming(self.h, int(on)))

    def kernelStats(self):
        fm, nm = C.c_double(), C.c_double()
        fl, nl, tl = C.c_longlong(), C.c_longlong(), C.c_longlong()
        self._ck(self.L.mdb_cp_getKernelStats(self.h, C.byref(fm), C.byref(fl), C.byref(nm), C.byref(nl), C.byref(tl)))
        return dict(force_ms=fm.value, force_launches=fl.value, neigh_ms=nm.value, neigh_launches=nl.value,
                    launches=tl.value)

    def resetKernelStats(self): self._ck(self.L.mdb_cp_resetKernelStats(self.h))

    def countPairs(self):
        a, b = C.c_longlong(), C.c_longlong()
        self._ck(self.L.mdb_cp_countPairs(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def stub(self, niclusters=256, natoms=4, pattern="seq", nneighs=9, nreps=1, masked=0, seed=12345):
        """synthetic clusters + lists of the reference's kernel micro-benchmark (clusterpair/main-stub.c)"""
        self._ck(self.L.mdb_cp_stub(self.h, niclusters, natoms, {"seq": 0, "fix": 1, "rand": 2, "local": 3, "localbank": 4}[pattern], nneighs, nreps, masked,
                                    C.c_uint(seed)))

    # ---- parity accessors (same names as the checker's bindings use for the reference) ----
    def tiles(self):
        c = self.counts()
        return c["ncj"], c["Nclusters_ghost"]

    def _clusters(self, which, n):
        nat = np.empty(n, np.int32)
        bb = np.empty((n, 6), self.np_real)
        self._ck(self.L.mdb_cp_getClusters(self.h, ord(which), _vp(nat), _vp(bb)))
        return nat, bb

    def iclusters(self): return self._clusters("i", self.counts()["Nclusters_local"])

    def jclusters(self):
        ncj, ngh = self.tiles()
        return self._clusters("j", ncj + ngh)

    def cl(self, what):
        ncj, ngh = self.tiles()
        nt = ncj + (ngh if what == "x" else 0)
        a = np.empty((nt, 3, self.N), self.np_real)
        self._ck(self.L.mdb_cp_getClusterData(self.h, ord(what), _vp(a)))
        return a

    def cluster_tags(self):
        ncj, ngh = self.tiles()
        a = np.empty((ncj + ngh, self.N), np.int32)
        self._ck(self.L.mdb_cp_getClusterTags(self.h, _vp(a)))
        return a

    def icluster_bin(self):
        a = np.empty(self.counts()["Nclusters_local"], np.int32)
        self._ck(self.L.mdb_cp_getClusterBins(self.h, _vp(a)))
        return a

    def cluster_lists(self, strip_dummy=True):
        """numneigh, numneigh_masked, [sorted j-cluster ids per i-cluster]"""
        c = self.counts()
        ncl, mx = c["Nclusters_local"], c["maxneighs"]
        nn, nm = np.empty(ncl, np.int32), np.empty(ncl, np.int32)
        nb = np.empty((ncl, mx), np.int32)
        self._ck(self.L.mdb_cp_getLists(self.h, _vp(nn), _vp(nm), _vp(nb), mx))
        return nn, nm, [np.sort(nb[ci, :nn[ci]]) for ci in range(ncl)]

    def raw_lists(self):
        c = self.counts()
        ncl, mx = c["Nclusters_local"], c["maxneighs"]
        nn, nm = np.empty(ncl, np.int32), np.empty(ncl, np.int32)
        nb = np.empty((ncl, mx), np.int32)
        self._ck(self.L.mdb_cp_getLists(self.h, _vp(nn), _vp(nm), _vp(nb), mx))
        return nn, nm, nb

    def ghost_map(self):
        ng = self.counts()["Nclusters_ghost"]
        a = [np.empty(ng, np.int32) for _ in range(4)]
        self._ck(self.L.mdb_cp_getGhostMap(self.h, *[_vp(q) for q in a]))
        return dict(border_map=a[0], PBCx=a[1], PBCy=a[2], PBCz=a[3])

    def neigh_params(self):
        iv = (C.c_int * 8)()
        rv = (C.c_double * 10)()
        self._ck(self.L.mdb_cp_getNeighborParams(self.h, iv, rv, None))
        d = dict(zip(("nbinx", "nbiny", "mbinx", "mbiny", "mbins", "mbinxlo", "mbinylo", "nstencil"), [int(q) for q in iv]))
        d.update(zip(("binsizex", "binsizey", "bininvx", "bininvy", "cutneighsq", "cutneigh", "xprd", "yprd", "zprd",
                      "rbb_sq"), [float(q) for q in rv]))
        st = np.empty(d["nstencil"], np.int32)
        self._ck(self.L.mdb_cp_getNeighborParams(self.h, iv, rv, _vp(st)))
        d["stencil"] = st
        return d
