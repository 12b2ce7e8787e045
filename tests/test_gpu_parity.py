"""GPU parity tests proper: the CUDA path (through the C ABI, via the ctypes mirror) against the
oracle on the same inputs, against the golden fixtures produced by the reference, and at full size
through size-independent properties."""
import json
import os

import numpy as np
import pytest

from conftest import load_pkg, ref_usable
from parity import TOL, check_snapshot, csr_sets, rel_err, run_lj_fixture
from portbind import OracleVL

pytestmark = pytest.mark.gpu


def make_sim(dp=True, aos=True, sort=True, **kw):
    m = load_pkg()
    p = m.default_params(precision=m.DP if dp else m.SP, layout=m.AOS if aos else m.SOA, **kw)
    s = m.Simulation(p)
    s.setOption("sort_atoms", 1 if sort else 0)
    return s


@pytest.mark.parametrize("sort", [True, False])
@pytest.mark.parametrize("name,dp,half,aos", [("lj_dp_full_nx6", True, 0, True), ("lj_dp_full_nx6", True, 0, False),
                                              ("lj_dp_half_nx6", True, 1, True), ("lj_sp_full_nx6", False, 0, False),
                                              ("lj_sp_full_nx6", False, 0, True)])
def test_cuda_matches_golden_fixture(golden_dir, name, dp, half, aos, sort):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    nx = int(g["nx"])
    s = make_sim(dp, aos, sort, nx=nx, ny=nx, nz=nx, half_neigh=half)
    run_lj_fixture(s, g, dp, feed=lambda x, v: s.setAtoms(x, v), setup_noadjust=lambda: s.setup(adjust=False))
    npar = s.neighborParams()
    for k in ("nbinx", "mbinx", "mbinxlo", "mbins", "nstencil"):
        assert npar[k] == int(g["ng_" + k]), k
    for k in ("bininvx", "binsizex", "cutneighsq"):
        assert npar[k] == float(g["ng_" + k]), k
    assert np.array_equal(npar["stencil"], g["ng_stencil"])
    s.close()


@pytest.mark.parametrize("dp", [True, False])
def test_createAtom_on_device_matches_oracle(dp):
    """device-side createAtom (closed-form emission index) vs the oracle's walk, incl. box sizes
    that are not multiples of the 8x8x8 sub-box"""
    for (nx, ny, nz) in [(6, 6, 6), (5, 7, 3), (9, 4, 11), (16, 16, 16)]:
        s = make_sim(dp, True, nx=nx, ny=ny, nz=nz)
        n = s.createAtom()
        o = OracleVL(dp)
        o.configure(nx=nx, ny=ny, nz=nz)
        o.derive()
        o.create_atoms()
        assert n == o.geti("Nlocal") == 4 * nx * ny * nz
        assert np.array_equal(s.get("x"), o.get("x"))
        assert np.array_equal(s.get("v"), o.get("v"))
        s.close()


@pytest.mark.parametrize("sort", [True, False])
@pytest.mark.parametrize("dp,half", [(True, 0), (True, 1), (False, 0), (False, 1)])
def test_operators_bit_exact_vs_oracle_on_jittered_box(dp, half, sort):
    """Every list-defining operator on identical input bits: wrap, ghosts, bins, lists."""
    rng = np.random.default_rng(11)
    nx, ny, nz = 7, 5, 6
    o = OracleVL(dp)
    o.configure(nx=nx, ny=ny, nz=nz, half_neigh=half)
    o.derive(); o.create_atoms(); o.setup_neighbor(); o.setup_thermo(); o.adjust_thermo()
    r = o.np_real
    x = (o.get("x") + rng.normal(0, 0.15, (o.geti("Nlocal"), 3))).astype(r)
    v = o.get("v")
    o.set_atoms(x, v)
    o.reneighbour()
    o.computeForce()
    s = make_sim(dp, True, sort, nx=nx, ny=ny, nz=nz, half_neigh=half)
    s.setAtoms(x, v)
    s.setupNeighbor(); s.setupThermo()
    s.reneighbour()
    s.computeForce()
    assert s.counts()["Nghost"] == o.geti("Nghost")
    assert np.array_equal(s.get("x", ghosts=True), o.get("x", ghosts=True))
    gm = s.ghostMap()
    for k in ("border_map", "PBCx", "PBCy", "PBCz"):
        assert np.array_equal(gm[k], o.get(k)), k
    assert np.array_equal(s.binCounts(), o.get("bincount"))
    nn, nb = s.neighbors()
    assert np.array_equal(nn, o.get("numneigh"))
    onb = o.get("neighbors")
    for i in range(len(nn)):
        if sort:   # internal spatial sort: same SET (north_star's parity object)
            assert np.array_equal(np.sort(nb[i, :nn[i]]), np.sort(onb[i, :nn[i]])), i
        else:      # reference atom order kept internally: even the row ORDER is the reference's
            assert np.array_equal(nb[i, :nn[i]], onb[i, :nn[i]]), i
    f, fo = s.get("f"), o.get("f")
    assert np.abs(f - fo).max() <= TOL[dp] * np.abs(fo).max()
    T, P = s.thermo(); To, Po = o.thermo()
    assert abs(T - To) <= TOL[dp] * To and abs(P - Po) <= TOL[dp] * Po
    s.close()


def test_maxneighs_resize_and_dense_bins():
    """neighbor.c:247-262: rows longer than maxneighs force a rebuild with 1.2x the longest row;
    a dense cluster also overflows the reference's 8-atom bins (atoms_per_bin doubling)."""
    rng = np.random.default_rng(3)
    o = OracleVL(True)
    o.configure(nx=6, cutforce=2.5, skin=1.3)      # cutneigh 3.8 -> ~190 neighbors > 100
    o.derive(); o.create_atoms(); o.setup_neighbor(); o.setup_thermo()
    x = o.get("x") + rng.normal(0, 0.05, (o.geti("Nlocal"), 3))
    o.set_atoms(x, None)
    o.reneighbour()
    s = make_sim(True, True, nx=6, ny=6, nz=6, skin=1.3)
    s.setAtoms(x, None)
    s.setupNeighbor()
    s.reneighbour()
    assert o.geti("maxneighs") > 100
    assert s.counts()["maxneighs"] == o.geti("maxneighs")
    nn, nb = s.neighbors()
    assert np.array_equal(nn, o.get("numneigh"))
    off, flat = csr_sets(nn, nb)
    off2, flat2 = csr_sets(o.get("numneigh"), o.get("neighbors"))
    assert np.array_equal(flat, flat2)
    s.close()


def test_run_loop_equals_operator_by_operator():
    """mdb_run (device-resident loop) must give exactly the operator-by-operator result."""
    a = make_sim(True, True, nx=6, ny=6, nz=6)
    b = make_sim(True, True, nx=6, ny=6, nz=6)
    for s in (a, b):
        s.createAtom(); s.setup(adjust=True)
    rec, tm = a.run(45)
    b.computeForce()
    for n in range(45):
        b.step(n)
    assert np.array_equal(a.get("x"), b.get("x"))
    assert np.array_equal(a.get("v"), b.get("v"))
    assert rec[0][0] == 0 and rec[-1][0] == 45
    T, P = b.thermo()
    assert rec[-1][1] == T and rec[-1][2] == P
    a.close(); b.close()


@pytest.mark.parametrize("dp,half,nx,key", [(True, 0, 32, "vl_dp_aos"), (False, 0, 32, "vl_sp_soa"), (True, 1, 8, "vl_dp_aos")])
def test_200_step_thermo_goldens(golden_dir, dp, half, nx, key):
    """BASELINE config 1 (Cu FCC 32^3, 200 steps): the `step temp pressure` lines of the reference."""
    th = json.load(open(os.path.join(golden_dir, "thermo_lj.json")))
    t = [q for q in th if q["variant"] == key and q["nx"] == nx and q["half"] == half][0]
    s = make_sim(dp, True, nx=nx, ny=nx, nz=nx, half_neigh=half, ntimes=200)
    s.createAtom(); s.setup(adjust=True)
    rec, tm = s.run(200)
    tol = 1e-9 if dp else 2e-4
    assert len(rec) == len(t["records"])
    for (st, T, P), (gs, gT, gP) in zip(rec, t["records"]):
        assert int(st) == gs
        assert abs(T - gT) <= max(tol * gT, 6e-7 * gT), (st, T, gT)
        assert abs(P - gP) <= max(tol * gP, 6e-7 * gP)
    assert abs(rec[-1][1] - t["T_full"]) <= tol * t["T_full"]
    # the ghost count is a bit-level function of positions: exact for DP; in SP one atom within
    # ~1e-5 of the ghost boundary may fall on the other side after 200 steps of rounding drift
    assert abs(s.counts()["Nghost"] - t["nghost"]) <= (0 if dp else 3)
    s.close()


def test_200_step_trajectory_vs_oracle_dp():
    """x and v after 200 steps (10 rebuilds) vs the oracle: DP rel 1e-10 (north_star)."""
    nx = 10
    s = make_sim(True, True, nx=nx, ny=nx, nz=nx)
    s.createAtom(); s.setup(adjust=True)
    o = OracleVL(True)
    o.configure(nx=nx)
    o.setup(create=True)
    o.set_atoms(s.get("x"), s.get("v"))          # identical initial bits (adjustThermo sums differ in order)
    o.setup(create=False)
    s.run(200)
    o.run(200)
    assert rel_err(s.get("v"), o.get("v")) < 1e-10
    L = s.neighborParams()["xprd"]
    d = np.abs(s.get("x") - o.get("x"))
    d = np.minimum(d, L - d)                      # an atom sitting on the wrap boundary may differ by one box
    assert d.max() < 1e-10 * L
    nn, nb = s.neighbors()
    off, flat = csr_sets(nn, nb)
    off2, flat2 = csr_sets(o.get("numneigh"), o.get("neighbors"))
    assert np.array_equal(nn, o.get("numneigh")) and np.array_equal(flat, flat2)
    s.close()


def test_full_size_properties_128cubed():
    """BASELINE full size per GPU (Cu FCC 128^3 = 8.39M atoms): properties that need no oracle run.
    - perfect lattice: every atom has exactly 78 listed neighbors (cutneigh 2.8) and 54 inside 2.5
    - net force on the lattice is zero to rounding; momentum is conserved over steps
    - state save/restore reproduces the run bit for bit (determinism)."""
    nx = 128
    s = make_sim(True, True, nx=nx, ny=nx, nz=nx)
    n = s.createAtom()
    assert n == 4 * nx ** 3
    s.setup(adjust=True)
    s.saveState()
    nn = s.numneigh()
    assert nn.min() == 78 and nn.max() == 78
    listed, inside = s.countPairs()
    assert listed == 78 * n and inside == 54 * n
    s.computeForce()
    f = s.get("f")
    assert np.abs(f).max() < 1e-10
    T0, _ = s.thermo()
    assert abs(T0 - 1.44) < 1e-12
    rec, _ = s.run(20)
    v = s.get("v")
    assert np.abs(v.sum(axis=0)).max() < 1e-7          # zero total momentum (adjustThermo) is conserved
    x1 = s.get("x")
    s.restoreState(); s.setup(adjust=False)
    rec2, _ = s.run(20)
    assert np.array_equal(rec, rec2)
    assert np.array_equal(s.get("x"), x1)
    s.close()


def test_reference_library_direct_if_runnable():
    """Where the prebuilt reference can run on this host, compare one rebuild + force with it directly."""
    if not ref_usable("vl_dp_aos"):
        pytest.skip("prebuilt reference library not runnable on this host")
    from refbind import RefVL
    r = RefVL("vl_dp_aos")
    r.param.nx = r.param.ny = r.param.nz = 12
    r.setup()
    for n in range(25):
        r.step(n)
    x, v = r.get("x"), r.get("v")
    r.reneighbour(); r.computeForce()
    s = make_sim(True, True, nx=12, ny=12, nz=12)
    s.setAtoms(x, v)
    s.setupNeighbor(); s.setupThermo()
    s.reneighbour(); s.computeForce()
    assert np.array_equal(s.get("x", ghosts=True), r.get("x", ghosts=True))
    nn, nb = s.neighbors()
    off, flat = csr_sets(nn, nb)
    off2, flat2 = csr_sets(r.get("numneigh"), r.get("neighbors"))
    assert np.array_equal(nn, r.get("numneigh")) and np.array_equal(flat, flat2)
    f, fr = s.get("f"), r.get("f")
    assert np.abs(f - fr).max() <= 1e-10 * np.abs(fr).max()
    s.close()
