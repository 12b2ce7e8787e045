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


@pytest.mark.parametrize("dp", [True, False])
@pytest.mark.parametrize("box,keep", [((2, 2, 2), 1.0), ((3, 2, 6), 1.0), ((2, 7, 2), 0.6), ((6, 6, 6), 0.35)])
def test_tiny_thin_and_ragged_boxes_bit_exact_vs_oracle(dp, box, keep):
    """Edge cases of the list-defining operators on identical input bits: boxes smaller than 2 * cutneigh (a single bin per axis,
    every atom has all 26 images), thin slabs, and ragged inputs (a random 40-65 % of the atoms removed: empty and sparse bins,
    atoms with short rows) with jitter -- the per-atom stencil of the list build must give the reference's rows,
    entry by entry, and 25 steps (one rebuild) must follow the oracle."""
    rng = np.random.default_rng(7)
    nx, ny, nz = box
    o = OracleVL(dp)
    o.configure(nx=nx, ny=ny, nz=nz)
    o.derive(); o.create_atoms(); o.setup_neighbor(); o.setup_thermo(); o.adjust_thermo()
    r = o.np_real
    n0 = o.geti("Nlocal")
    sel = np.sort(rng.choice(n0, max(2, int(keep * n0)), replace=False))
    x = (o.get("x")[sel] + rng.normal(0, 0.08, (len(sel), 3))).astype(r)   # larger jitters blow the lattice up (the reference
    v = (0.3 * o.get("v")[sel]).astype(r)                                  # then leaves its bin grid and crashes)
    o.set_atoms(x, v)
    o.reneighbour()
    o.computeForce()
    s = make_sim(dp, True, False, nx=nx, ny=ny, nz=nz)
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
        assert np.array_equal(nb[i, :nn[i]], onb[i, :nn[i]]), i
    f, fo = s.get("f"), o.get("f")
    assert np.abs(f - fo).max() <= TOL[dp] * max(np.abs(fo).max(), 1.0)
    for n in range(25):
        s.step(n); o.step(n)
    assert np.abs(s.get("x") - o.get("x")).max() <= (1e-10 if dp else 1e-4) * max(np.abs(o.get("x")).max(), 1.0)
    nn, nb = s.neighbors()
    assert np.array_equal(nn, o.get("numneigh")) or not dp   # SP trajectories may flip a membership at the skin after 25 steps
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


@pytest.mark.parametrize("dp,sort,fuse_force", [(True, True, 1), (True, False, 1), (False, False, 1), (True, True, 0),
                                                (True, True, 2), (False, False, 2)])
def test_run_loop_equals_operator_by_operator(dp, sort, fuse_force):
    """mdb_run (device-resident loop: integrate halves fused into the force kernel's epilogue, positions double-buffered;
    or, fuse_force=0, the separate final+initial integrate pass) must give exactly the operator-by-operator result,
    across rebuilds (steps 20, 40) and a thermo record in the middle (nstat 30)."""
    a = make_sim(dp, True, sort, nx=6, ny=6, nz=6, nstat=30)
    b = make_sim(dp, True, sort, nx=6, ny=6, nz=6, nstat=30)
    a.setOption("fuse_force", min(fuse_force, 1))
    a.setOption("xy_gather", 1 if fuse_force == 2 else 0)   # 2: fused kernel with the packed (x, y) vector gathers
    for s in (a, b):
        s.createAtom(); s.setup(adjust=True)
    rec, tm = a.run(45)
    b.computeForce()
    for n in range(45):
        b.step(n)
    assert np.array_equal(a.get("x"), b.get("x"))
    assert np.array_equal(a.get("v"), b.get("v"))
    assert rec[0][0] == 0 and rec[1][0] == 30 and rec[-1][0] == 45
    T, P = b.thermo()
    assert rec[-1][1] == T and rec[-1][2] == P
    assert np.array_equal(a.get("f"), b.get("f"))  # the last step is never fused: f is the force of step 45
    # and the loop can be re-entered with the buffers swapped an odd number of times
    a.run(7)
    for n in range(45, 52):
        b.step(n)
    assert np.array_equal(a.get("x"), b.get("x")) and np.array_equal(a.get("v"), b.get("v"))
    a.close(); b.close()


@pytest.mark.parametrize("dp,sort", [(True, True), (False, False)])
def test_lazy_operators_equal_run_loop(dp, sort):
    """option lazy_ops: the reference's operator-by-operator loop (computeForce, finalIntegrate, initialIntegrate through
    the C ABI) must give mdb_run's result bit for bit -- with the fused kernel doing the work (fewer launches) -- across
    rebuilds, a thermo read in the middle (which forces the separate kernels) and reads of x / v / f at the end"""
    a = make_sim(dp, True, sort, nx=6, ny=6, nz=6, nstat=30)
    b = make_sim(dp, True, sort, nx=6, ny=6, nz=6, nstat=30)
    c = make_sim(dp, True, sort, nx=6, ny=6, nz=6, nstat=30)
    b.setOption("lazy_ops", 1)
    for s in (a, b, c):
        s.createAtom(); s.setup(adjust=True)
    rec, _ = a.run(45)
    for s in (b, c):
        s.resetKernelStats()
        s.computeForce()
        for n in range(45):
            s.step(n)
            if n + 1 == 30:
                assert s.thermo() == (rec[1][1], rec[1][2])
    for s in (b, c):
        assert np.array_equal(a.get("x"), s.get("x")) and np.array_equal(a.get("v"), s.get("v"))
        assert np.array_equal(a.get("f"), s.get("f"))
    assert b.kernelStats()["launches"] < c.kernelStats()["launches"] - 35   # one launch less on ~40 of the 45 steps
    for s in (a, b, c):
        s.close()


@pytest.mark.parametrize("dp,half,nx,key", [(True, 0, 32, "vl_dp_aos"), (False, 0, 32, "vl_sp_soa"), (True, 1, 8, "vl_dp_aos")])
def test_200_step_thermo_goldens(golden_dir, dp, half, nx, key):
    """BASELINE config 1 (Cu FCC 32^3, 200 steps): the `step temp pressure` lines of the reference."""
    th = json.load(open(os.path.join(golden_dir, "thermo_lj.json")))
    t = [q for q in th if q["variant"] == key and q["nx"] == nx and q["half"] == half][0]
    s = make_sim(dp, True, nx=nx, ny=nx, nz=nx, half_neigh=half, ntimes=200)
    s.createAtom(); s.setup(adjust=True)
    rec, tm = s.run(200)
    tol = 1e-10 if dp else 1e-4   # north_star: DP rel 1e-10, SP rel 1e-4 over 200 steps
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


# ---- BASELINE config 3: argon, half neighbor lists -------------------------------------------------
@pytest.mark.parametrize("sort", [False, True])
def test_cuda_argon_half_lists(golden_dir, sort):
    from cases import argon_cuda, argon_oracle, row_checksums
    g = np.load(os.path.join(golden_dir, "argon_half.npz"))
    s = argon_cuda(g, sort)
    c = s.counts()
    assert c["Nghost"] == int(g["nghost0"]) and c["maxneighs"] == int(g["maxneighs0"])
    nn, nb = s.neighbors()
    assert np.array_equal(nn, g["numneigh0"])
    s1, s2 = row_checksums(nn, nb)
    assert np.array_equal(s1, g["rowsum0"]) and np.array_equal(s2, g["rowsq0"])
    o = argon_oracle(g)
    off, flat = csr_sets(nn, nb)
    off2, flat2 = csr_sets(o.get("numneigh"), o.get("neighbors"))
    assert np.array_equal(flat, flat2)                       # full sets against the oracle
    assert np.array_equal(s.get("x", ghosts=True), o.get("x", ghosts=True))
    s.computeForce()
    assert np.abs(s.get("f") - g["f0"]).max() <= 1e-10 * np.abs(g["f0"]).max()
    T, P = s.thermo()
    assert abs(T - g["thermo0"][0]) <= 1e-12 * T
    rec, _ = s.run(200)
    assert rel_err(s.get("x"), g["xN"]) < 1e-10 and rel_err(s.get("v"), g["vN"]) < 1e-10
    assert np.abs(s.get("f") - g["fN"]).max() <= 1e-10 * np.abs(g["fN"]).max()
    nn, nb = s.neighbors()
    s1, s2 = row_checksums(nn, nb)
    assert np.array_equal(nn, g["numneighN"]) and np.array_equal(s1, g["rowsumN"]) and np.array_equal(s2, g["rowsqN"])
    assert s.counts()["Nghost"] == int(g["nghostN"])
    assert abs(rec[-1][1] - g["thermoN"][0]) <= 1e-10 * g["thermoN"][0]
    s.close()


# ---- EAM (verletlist): tables, three passes, BASELINE config 4 ---------------------------------------
@pytest.mark.parametrize("sort", [False, True])
def test_cuda_eam_lattice(golden_dir, sort):
    from cases import eam_cuda, eam_oracle, row_checksums
    g = np.load(os.path.join(golden_dir, "eam_cu_nx5.npz"))
    s = eam_cuda(g, sort=sort)
    t = s.getEamSplines()
    for k in ("nr", "nrho", "nr_tot", "nrho_tot"):
        assert t[k] == int(g["eam_" + k]), k
    assert t["rdr"] == float(g["eam_rdr"]) and t["rdrho"] == float(g["eam_rdrho"])
    for k in ("rhor_spline", "frho_spline", "z2r_spline"):
        ref = g["eam_" + k]
        assert np.abs(t[k][7:] - ref[7:]).max() <= 1e-13 * np.abs(ref).max(), k
    npar = s.neighborParams()
    assert abs(npar["cutforce"] - g["params"][0]) < 1e-15 and abs(npar["cutneigh"] - g["params"][1]) < 1e-15
    assert abs(npar["dtforce"] - g["params"][3]) <= 1e-15 * g["params"][3]
    assert s.counts()["Nghost"] == int(g["nghost0"])
    s.setAtoms(g["x0"], g["v0"]); s.setup(adjust=False)            # identical velocity bits
    nn, nb = s.neighbors()
    s1, s2 = row_checksums(nn, nb)
    assert np.array_equal(nn, g["numneigh0"]) and np.array_equal(s1, g["rowsum0"]) and np.array_equal(s2, g["rowsq0"])
    s.computeForceEam()
    assert np.abs(s.getEamFp(ghosts=True) - g["fp0"]).max() <= 1e-11 * np.abs(g["fp0"]).max()
    assert np.abs(s.get("f") - g["f0"]).max() <= 1e-12   # lattice: net forces are cancellation noise, absolute bound
    rec, _ = s.run(int(g["nsteps"]))
    assert abs(rec[-1][1] - g["thermoN"][0]) <= 1e-10 * g["thermoN"][0]
    assert rel_err(s.get("v"), g["vN"]) < 1e-10 and rel_err(s.get("x"), g["xN"]) < 1e-10
    assert np.abs(s.get("f") - g["fN"]).max() <= 1e-10 * np.abs(g["fN"]).max()
    assert s.counts()["Nghost"] == int(g["nghostN"])
    s.close()


@pytest.mark.parametrize("dp", [True, False])
def test_eam_generation3_matches_generation2(golden_dir, dp):
    """eam_variant 2 (default: packed (x, y) / (z, fp) gathers, (value, slope) tables with the cubic's coefficients derived
    in registers) against generation 2 (what a brick runs): forces after setup and a 60-step run, to rounding (the derivative
    is evaluated as ((3 c3 p + 2 c4) p + c5) * rdr instead of with pre-divided coefficients)"""
    from cases import funcfl_args
    g = np.load(os.path.join(golden_dir, "eam_cu_nx5.npz"))
    m = load_pkg()
    sims = []
    for variant in (1, 2):
        s = make_sim(dp, True, True, force_field=m.FF_EAM, nx=6, ny=6, nz=6, ntimes=60)
        s.setOption("eam_variant", variant)
        s.setEam(*funcfl_args(g))
        s.createAtom(); s.setup(adjust=True)
        s.computeForce()
        sims.append(s)
    a, b = sims
    fa, fb = a.get("f"), b.get("f")
    # the t = 0 lattice forces are cancellation noise: compare against the scale of a thermalised step below
    ra, _ = a.run(60)
    rb, _ = b.run(60)
    tol = 1e-10 if dp else 1e-4
    fa, fb = a.get("f"), b.get("f")
    assert np.abs(fa - fb).max() <= tol * np.abs(fa).max()
    assert np.abs(ra[:, 1:] - rb[:, 1:]).max() <= tol * np.abs(ra[:, 1:]).max()
    assert np.abs(a.get("v") - b.get("v")).max() <= (1e-10 if dp else 1e-3) * np.abs(a.get("v")).max()
    a.close(); b.close()


@pytest.mark.gpu
@pytest.mark.parametrize("dp", [True, False])
def test_eam_run_loop_fused_equals_separate_kernels(golden_dir, dp):
    """mdb_run with EAM: the integrate halves in the epilogue of the force pass (k_eam_force_v3<.., FI>, positions updated in
    place) against the separate final+initial integrate kernel (option fuse_force = 0): same operation sequence, so
    positions, velocities and thermo records are equal bit for bit over rebuilds and a thermo record in between."""
    from cases import funcfl_args
    g = np.load(os.path.join(golden_dir, "eam_cu_nx5.npz"))
    m = load_pkg()
    out = []
    for fuse in (1, 0):
        s = make_sim(dp, True, True, force_field=m.FF_EAM, nx=6, ny=6, nz=6, ntimes=130)
        s.setOption("fuse_force", fuse)
        s.setEam(*funcfl_args(g))
        s.createAtom(); s.setup(adjust=True)
        rec, _ = s.run(130)
        out.append((np.array(rec), s.get("x"), s.get("v")))
        s.close()
    (ra, xa, va), (rb, xb, vb) = out
    assert np.array_equal(ra, rb) and np.array_equal(xa, xb) and np.array_equal(va, vb)


def test_cuda_eam_copper_melting_200_steps(golden_dir):
    """BASELINE config 4: copper_melting (32 000 atoms from the LAMMPS dump), Cu_u3 funcfl, 200 steps:
    the reference's `step temp pressure` lines and the oracle on the first 25 steps."""
    from cases import eam_cuda, eam_oracle
    g = np.load(os.path.join(golden_dir, "eam_cu_melting.npz"))
    s = eam_cuda(g, from_dump=True)
    assert s.counts()["Nghost"] == int(g["nghost0"])
    assert np.array_equal(s.numneigh(), g["numneigh0"])
    s.computeForceEam()
    assert np.abs(s.getEamFp(ghosts=True) - g["fp0"]).max() <= 1e-11 * np.abs(g["fp0"]).max()
    # the dump starts on the perfect lattice: net forces are cancellation noise
    assert np.abs(s.get("f") - g["f0"]).max() <= max(1e-10 * np.abs(g["f0"]).max(), 1e-12)
    rec, _ = s.run(200)
    assert len(rec) == len(g["records"])
    for (st, T, P), (gs, gT, gP) in zip(rec, g["records"]):
        assert int(st) == int(gs) and abs(T - gT) <= 6e-7 * gT and abs(P - gP) <= 6e-7 * gP, (st, T, gT)
    assert abs(rec[-1][1] - g["thermoN"][0]) <= 1e-10 * g["thermoN"][0]
    assert s.counts()["Nghost"] == int(g["nghostN"])
    s.close()


# ---- the C driver: reference command line, report format, file readers ------------------------------
def _driver():
    from conftest import ROOT
    exe = os.path.join(ROOT, "md-bench_b200", "driver", "MDBench-VL-B200")
    if not os.path.exists(exe):
        import subprocess
        subprocess.check_call(["make", "-s", "-C", os.path.dirname(exe)])
    return exe


def _thermo_lines(out):
    import re
    return [(int(a), b, c) for a, b, c in re.findall(r"^(-?\d+)\t(\S+)\t(\S+)$", out, flags=re.M)]


@pytest.mark.parametrize("extra", [[], ["--operators"], ["--layout", "soa"], ["--sort"]])
def test_driver_default_run_prints_reference_report(golden_dir, extra):
    import subprocess
    th = json.load(open(os.path.join(golden_dir, "thermo_lj.json")))
    t = [q for q in th if q["variant"] == "vl_dp_aos" and q["nx"] == 32 and q["half"] == 0][0]
    out = subprocess.run([_driver()] + extra, capture_output=True, text=True, timeout=300).stdout
    lines = _thermo_lines(out)
    assert [l[0] for l in lines] == [0, 100, 200]
    for (st, T, P), (gs, gT, gP) in zip(lines, t["records"]):
        assert "%e" % gT == T and "%e" % gP == P, (st, T, P, gT, gP)      # the printed 7 digits
    assert "System: 131072 atoms %d ghost atoms, Steps: 200" % t["nghost"] in out
    assert "million atom updates per second" in out and "TOTAL" in out and "Kernel: CUDA-sm_100a" in out


@pytest.mark.parametrize("variant,env", [("vl_dp_aos", {}), ("vl_dp_aos", {"MDB_LAZY_OPS": "1"}), ("vl_sp_soa", {})])
def test_reference_main_c_drives_libmdb200(golden_dir, variant, env):
    """The boundary proven with the reference's own driver: oracle/_ref/MDBench-<variant>-b200 is the reference's UNMODIFIED
    main.c / atom.c / thermo.c / parameter.c (compiled from /root/reference where they lie, -DCUDA_TARGET) linked against
    md-bench_b200/driver/b200_shim.c + libmdb200.so (oracle/Makefile ref-shim).  Its default run (BASELINE config 1) must
    print the reference CPU build's thermo lines and ghost count."""
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "oracle", "_ref", "MDBench-%s-b200" % variant)
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/MDBench-%s-b200 not built (needs /root/reference: make -C oracle ref-shim)" % variant)
    th = json.load(open(os.path.join(golden_dir, "thermo_lj.json")))
    t = [q for q in th if q["variant"] == variant and q["nx"] == 32 and q["half"] == 0][0]
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300, env=dict(os.environ, **env)).stdout
    lines = _thermo_lines(out)
    assert [l[0] for l in lines] == [0, 100, 200], out[-2000:]
    for (st, T, P), (gs, gT, gP) in zip(lines, t["records"]):
        if variant.startswith("vl_dp"):
            assert "%e" % gT == T and "%e" % gP == P, (st, T, P, gT, gP)      # the printed 7 digits
        else:
            assert abs(float(T) - gT) <= 1e-4 * gT and abs(float(P) - gP) <= 1e-4 * gP, (st, T, P, gT, gP)
    assert "System: 131072 atoms %d ghost atoms, Steps: 200" % t["nghost"] in out
    assert "million atom updates per second" in out


def test_driver_reads_gro_and_param_file(golden_dir, tmp_path):
    """-p params -i input.gro -half 1 (BASELINE config 3): inputs rebuilt from the fixture"""
    import subprocess
    g = np.load(os.path.join(golden_dir, "argon_half.npz"))
    gro = tmp_path / "input.gro"
    with open(gro, "w") as f:
        f.write("Liquid Argon t=   0.00000 step= 0\n %d\n" % len(g["x0"]))
        for i, (x, v) in enumerate(zip(g["x0"], g["v0"])):
            f.write("%5dAr      Ar%5d%8.3f%8.3f%8.3f%8.4f%8.4f%8.4f\n" % (i + 1, i + 1, x[0], x[1], x[2], v[0], v[1], v[2]))
        f.write("   %.5f   %.5f   %.5f\n" % (g["box"][1], g["box"][3], g["box"][5]))
    conf = tmp_path / "params.conf"
    eps, sig, cutf, skin, dt, temp, rho, mass = [float(v) for v in g["params"]]
    conf.write_text("mass %.17g\nsigma %.17g\nepsilon %.17g\nntimes 250000\ndt %.17g\ntemp %.17g\nx_out_freq 500\n"
                    "v_out_freq 5\ncutforce %.17g\nskin %.17g\nreneigh_every %d\nnstat %d   # comment\n"
                    % (mass, sig, eps, dt, temp, cutf, skin, int(g["ints"][0]), int(g["ints"][1])))
    out = subprocess.run([_driver(), "-p", str(conf), "-i", str(gro), "-half", "1", "-n", "200"],
                         capture_output=True, text=True, timeout=300).stdout
    lines = _thermo_lines(out)
    assert lines[0][0] == 0 and lines[-1][0] == 200
    assert lines[0][1] == "%e" % g["thermo0"][0] and lines[0][2] == "%e" % g["thermo0"][1]
    assert lines[-1][1] == "%e" % g["thermoN"][0] and lines[-1][2] == "%e" % g["thermoN"][1]
    assert "System: 1000 atoms %d ghost atoms" % int(g["nghostN"]) in out
    assert "Read 1000 atoms from" in out


def test_driver_eam_funcfl_file(golden_dir, tmp_path):
    """-f eam -e <funcfl> -nx 5 -n 45: the potential file is rebuilt from the fixture's tables"""
    import subprocess
    g = np.load(os.path.join(golden_dir, "eam_cu_nx5.npz"))
    pot = tmp_path / "Cu_test.eam"
    with open(pot, "w") as f:
        f.write("funcfl table rebuilt from tests/golden/eam_cu_nx5.npz\n")
        f.write("   29 %.17g 3.6150 FCC\n" % float(g["funcfl_mass"]))
        f.write("%d %.17g %d %.17g %.17g\n" % (int(g["funcfl_nrho"]), float(g["funcfl_drho"]), int(g["funcfl_nr"]),
                                               float(g["funcfl_dr"]), float(g["funcfl_cut"])))
        for arr in (g["funcfl_frho"], g["funcfl_zr"], g["funcfl_rhor"]):
            for k in range(0, len(arr), 5):
                f.write(" ".join("%.17g" % v for v in arr[k:k + 5]) + "\n")
    out = subprocess.run([_driver(), "-f", "eam", "-e", str(pot), "-nx", "5", "-ny", "5", "-nz", "5", "-n", "45"],
                         capture_output=True, text=True, timeout=300).stdout
    lines = _thermo_lines(out)
    assert [l[0] for l in lines] == [0, 45]
    for (st, T, P), (gs, gT, gP) in zip(lines, g["records"]):
        assert abs(float(T) - gT) <= 2e-6 * gT and abs(float(P) - gP) <= 2e-6 * gP, (st, T, gT)
    assert "Force field: eam" in out


@pytest.mark.parametrize("pbc", [(1, 0, 1), (0, 0, 1), (0, 0, 0)])
def test_open_boundaries_pbc_flags(pbc):
    """param pbc_x/pbc_y/pbc_z = 0 (parameter.c:102-104; honoured by setupPbc, pbc.c:107-224): no images across an open
    face -- ghost set, ghost coordinates and lists bit-exact against the checker, forces to tolerance"""
    rng = np.random.default_rng(5)
    nx, ny, nz = 6, 5, 7
    o = OracleVL(True)
    o.configure(nx=nx, ny=ny, nz=nz, pbc=pbc)
    o.derive(); o.create_atoms(); o.setup_neighbor(); o.setup_thermo(); o.adjust_thermo()
    x = (o.get("x") + rng.normal(0, 0.1, (o.geti("Nlocal"), 3)))
    # keep every atom inside the box: an open face does not wrap (updateAtomsPbc wraps regardless of the flags)
    prd = np.array([o.getr("xprd"), o.getr("yprd"), o.getr("zprd")])
    x = np.clip(x, 1e-3, prd - 1e-3)
    v = o.get("v")
    o.set_atoms(x, v)
    o.reneighbour()
    o.computeForce()
    s = make_sim(True, True, False, nx=nx, ny=ny, nz=nz, pbc_x=pbc[0], pbc_y=pbc[1], pbc_z=pbc[2])
    s.setAtoms(x, v)
    s.setupNeighbor(); s.setupThermo()
    s.reneighbour()
    s.computeForce()
    assert s.counts()["Nghost"] == o.geti("Nghost")
    if pbc == (0, 0, 0):
        assert s.counts()["Nghost"] == 0
    assert np.array_equal(s.get("x", ghosts=True), o.get("x", ghosts=True))
    gm = s.ghostMap()
    for k in ("border_map", "PBCx", "PBCy", "PBCz"):
        assert np.array_equal(gm[k], o.get(k)), k
    nn, nb = s.neighbors()
    assert np.array_equal(nn, o.get("numneigh"))
    onb = o.get("neighbors")
    for i in range(len(nn)):
        assert np.array_equal(nb[i, :nn[i]], onb[i, :nn[i]]), i
    f, fo = s.get("f"), o.get("f")
    assert np.abs(f - fo).max() <= 1e-10 * np.abs(fo).max()
    s.close()


@pytest.mark.parametrize("pattern", ["seq", "fix", "rand"])
def test_stub_neighbor_patterns_and_force(pattern):
    """the kernel micro-benchmark's synthetic lists (main-stub.c:62-106) and the force they give, against numpy"""
    n, nn, nr = 300, 76, 2
    m = load_pkg()
    s = m.Simulation(m.default_params(layout=m.SOA, nx=1, ny=1, nz=1, cutforce=1.0e6, skin=0.0))
    x = np.repeat((np.arange(n) * 1e-5)[:, None], 3, axis=1)
    s.setAtoms(x, None)
    s.stubNeighbors(pattern, nn, nr)
    cnt, nb = s.neighbors()
    assert np.all(cnt == nn * nr)
    assert np.array_equal(nb[:, :nn], nb[:, nn:2 * nn])          # replicated nreps times
    if pattern == "seq":
        assert np.array_equal(nb[:, :nn], (np.arange(n)[:, None] + 1 + np.arange(nn)[None, :]) % n)
    elif pattern == "fix":
        assert np.array_equal(nb[:, :nn], np.repeat(np.arange(nn)[None, :], n, axis=0))
    else:
        assert nb.min() >= 0 and nb.max() < n and not np.any(nb == np.arange(n)[:, None])
        assert len(np.unique(nb[:, :nn])) > n // 2               # spread over the atoms
    if pattern != "fix":                                         # "fix" lists atom i itself for i < nneighs (r = 0)
        s.computeForceLJFullNeigh()
        f = s.get("f")
        d = x[:, None, :] - x[nb]                                # (n, nn*nr, 3)
        rsq = (d * d).sum(axis=2)
        sr2 = 1.0 / rsq
        sr6 = sr2 ** 3
        ref = (d * (48.0 * sr6 * (sr6 - 0.5) * sr2)[:, :, None]).sum(axis=1)
        assert np.abs(f - ref).max() <= 1e-10 * np.abs(ref).max()
    s.close()


def test_stub_driver_report():
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "md-bench_b200", "driver", "MDBench-VL-B200-stub")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.dirname(exe)])
    out = subprocess.run([exe, "-p", "rand", "-na", "20000", "-nn", "76", "-n", "20"], capture_output=True, text=True, timeout=120).stdout
    assert "Pattern: rand" in out and "Number of atoms: 20000" in out and "Mega atom updates/s" in out
    # displayStatistics of the reference's default build (stats.c:22-68): 20000 atoms x 76 neighbors x 20 calls
    assert "Statistics:" in out and "Average neighbors per atom: 76.0000" in out
    assert "Total number of computed pair interactions: 30400000" in out and "Total number of SIMD iterations: 950000" in out
    out = subprocess.run([exe, "-p", "seq", "-na", "4096", "--csv", "-n", "5"], capture_output=True, text=True, timeout=120).stdout
    assert out.splitlines()[0].startswith("steps,pattern,natoms,nneighs,nreps") and out.splitlines()[1].startswith("5,seq,4096,76,1,")


def test_ntypes_explicit_types():
    """ntypes > 1 (the reference's EXPLICIT_TYPES builds): createAtom draws type = rand() % ntypes per atom from the host's rand()
    sequence in emission order (atom.c:159); the pair parameters of every type pair are the same (atom.c:84-89), so the run equals
    the ntypes = 1 run bit for bit; types survive the spatial sort and saveState / restoreState"""
    import ctypes
    m = load_pkg()
    libc = ctypes.CDLL("libc.so.6")
    a = m.Simulation(m.default_params(nx=8, ny=8, nz=8))
    a.createAtom(); a.setup(adjust=True)
    ra, _ = a.run(45)
    libc.srand(1)
    b = m.Simulation(m.default_params(nx=8, ny=8, nz=8, ntypes=3))
    n = b.createAtom()
    libc.srand(1)
    expect = np.array([libc.rand() % 3 for _ in range(n)], np.int32)
    assert np.array_equal(b.types(), expect) and len(set(expect.tolist())) == 3
    b.setup(adjust=True)
    b.saveState()
    rb, _ = b.run(45)
    assert np.array_equal(ra, rb) and np.array_equal(a.get("x"), b.get("x")) and np.array_equal(a.get("v"), b.get("v"))
    assert np.array_equal(b.types(), expect)
    b.restoreState(); b.setup(adjust=False)
    assert np.array_equal(b.types(), expect)
    rc, _ = b.run(45)
    assert np.array_equal(rb, rc)
    libc.srand(1)
    c = m.Simulation(m.default_params(nx=8, ny=8, nz=8, ntypes=3))
    c.setOption("sort_atoms", 1)   # slots are permuted at every rebuild: the accessor still speaks the reference's numbering
    c.createAtom(); c.setup(adjust=True)
    c.run(45)
    assert np.array_equal(c.types(), expect)
    a.close(); b.close(); c.close()
