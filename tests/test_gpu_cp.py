"""GPU parity tests of the CLUSTERPAIR scheme (BASELINE config 2): the CUDA path through the C ABI (mdb_cp_*, via
the ctypes mirror ClusterSimulation) against the checker (cpbind.OracleCP) on the same inputs, against the golden
fixtures produced by the reference's clusterpair builds, and at full size through size-independent properties.

Tolerances: clusters (membership, order, bounding boxes), ghost clusters, ghost maps, cluster bins and cluster-pair
lists (as sorted index sets) bit-exact; forces / x / v / thermo DP rel 1e-10, SP rel 1e-4 (relative to the largest
magnitude of the compared array; the t = 0 lattice forces are cancellation noise)."""
import json
import os

import numpy as np
import pytest

from conftest import load_pkg
from cpbind import OracleCP, initial_atoms
from test_cp_oracle_pinned import _Fixture, assert_same_structure

pytestmark = pytest.mark.gpu

TOL = {True: 1e-10, False: 1e-4}


def make_cp(dp, N, nx, ny=None, nz=None, half=0, aos=True, **kw):
    m = load_pkg()
    p = m.default_params(precision=m.DP if dp else m.SP, layout=m.AOS if aos else m.SOA, nx=nx, ny=ny or nx, nz=nz or nx,
                         half_neigh=half, **kw)
    return m.ClusterSimulation(p, cluster_n=N)


def make_oracle(dp, N, nx, ny=None, nz=None, half=0, **kw):
    o = OracleCP(dp, N, N)   # vector width = N: no dummy padding, like the device lists
    o.configure(nx=nx, ny=ny, nz=nz, half_neigh=half, **kw)
    return o


def jittered(dp, nx, ny, nz, amp=0.08, seed=3):
    """lattice atoms displaced by up to `amp` and wrapped: no equal z inside a column (rank-sort path), partially filled
    clusters at the column ends, atoms on both sides of every periodic face"""
    x, v = initial_atoms(dp, nx, ny, nz)
    r = np.random.default_rng(seed)
    real = np.float64 if dp else np.float32
    lat = real((4.0 / 0.8442) ** (1.0 / 3.0))
    prd = np.array([real(nx * lat), real(ny * lat), real(nz * lat)], real)
    x = (x + r.uniform(-amp, amp, x.shape)).astype(real)
    x = np.where(x < 0, x + prd, x)
    x = np.where(x >= prd, x - prd, x).astype(real)
    return x, v


def max_rel(a, b):
    a, b = np.nan_to_num(np.asarray(a, np.float64), posinf=0, neginf=0), np.nan_to_num(np.asarray(b, np.float64), posinf=0, neginf=0)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def check_forces(s, o, dp, floor=1.0):
    fs, fo = np.nan_to_num(s.cl("f")), np.nan_to_num(o.cl("f"))
    real_lane = np.isfinite(o.cl("x")[:len(fo)])
    d = np.abs(fs - fo)[real_lane].max()
    assert d <= TOL[dp] * max(floor, np.abs(fo[real_lane]).max()), d


@pytest.mark.parametrize("half", [0, 1])
@pytest.mark.parametrize("N", [4, 8])
@pytest.mark.parametrize("dp", [True, False])
def test_structures_forces_trajectory_vs_oracle(dp, N, half):
    """lattice input (every column is full of z ties: the selection-sort replay decides the cluster contents),
    operator by operator through two rebuilds"""
    x, v = initial_atoms(dp, 6)
    s, o = make_cp(dp, N, 6, half=half), make_oracle(dp, N, 6, half=half)
    s.setAtoms(x, v); o.set_atoms(x, v)
    s.setupNeighbor(); s.setupThermo()
    s.buildClusters(); s.defineJClusters(); s.setupPbc(); s.binClusters(); s.buildNeighbor()
    o.setup()
    a, b = s.neigh_params(), o.neigh_params()
    for k in ("nbinx", "nbiny", "mbinx", "mbiny", "mbins", "mbinxlo", "mbinylo", "nstencil"):
        assert a[k] == b[k], k
    for k in ("binsizex", "binsizey", "bininvx", "bininvy", "cutneighsq", "xprd"):
        assert a[k] == b[k], k
    assert np.array_equal(a["stencil"], b["stencil"])
    assert_same_structure(o, s)
    real_lane = np.isfinite(o.cl("x")[:len(o.cl("v"))])
    assert np.array_equal(o.cl("v")[real_lane], s.cl("v")[real_lane])
    assert np.array_equal(o.tags()[1].reshape(-1, N)[:len(s.cluster_tags())], s.cluster_tags())
    s.computeForce(); o.computeForce()
    check_forces(s, o, dp)
    for n in range(45):
        ra, rb = s.step(n), o.step(n)
        assert ra == rb
        if ra:   # independent trajectories: coordinates to rounding, all integer structure still exact
            assert_same_structure(o, s, tol=TOL[dp])
    check_forces(s, o, dp, floor=0.0)
    s.updateSingleAtoms(); o.updateSingleAtoms()
    xs, ts = s.atoms("x", tags=True)
    assert np.array_equal(ts, o.tags()[0])
    assert max_rel(xs, o.atoms("x")) <= TOL[dp]
    assert max_rel(s.atoms("v"), o.atoms("v")) <= (1e-10 if dp else 1e-3)
    (Ts, Ps), (To, Po) = s.thermo(), o.thermo()
    assert abs(Ts - To) <= TOL[dp] * To and abs(Ps - Po) <= TOL[dp] * Po
    s.close()


@pytest.mark.parametrize("aos", [True, False])
@pytest.mark.parametrize("N", [4, 8])
@pytest.mark.parametrize("dp", [True, False])
def test_jittered_noncubic_box_bit_exact(dp, N, aos):
    """non-cubic box, no z ties, ragged columns (partially filled clusters, odd cluster counts padded for N = 8)"""
    nx, ny, nz = 5, 7, 6
    x, v = jittered(dp, nx, ny, nz)
    s, o = make_cp(dp, N, nx, ny, nz, aos=aos), make_oracle(dp, N, nx, ny, nz)
    s.setAtoms(x, v); o.set_atoms(x, v)
    s.setup(adjust=False); o.setup()
    assert_same_structure(o, s)
    nat, _ = s.iclusters()
    assert nat.min() < 4, "the case is meant to contain partially filled clusters"
    s.computeForce(); o.computeForce()
    check_forces(s, o, dp, floor=0.0)
    # wrap + rebuild from moved atoms
    for n in range(20):
        s.step(n); o.step(n)
    assert_same_structure(o, s, tol=TOL[dp])
    s.close()


@pytest.mark.parametrize("name", ["cp44_sp_nx6", "cp44_dp_nx6", "cp44_dp_half_nx6", "cp48ref_dp_nx6", "cp48_dp_nx6", "cp48_sp_nx6"])
def test_cuda_matches_golden_fixture(golden_dir, name):
    """the committed snapshots of the REFERENCE's clusterpair builds (tests/golden/make_golden_cp.py)"""
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    dp, N, nx, half = g["x0"].dtype == np.float64, int(g["N"]), int(g["nx"]), int(g["half"])
    s = make_cp(dp, N, nx, half=half)
    s.setAtoms(g["x0"], g["v0"])
    s.setup(adjust=False)
    f = _Fixture(g)
    if name == "cp48_sp_nx6":   # AVX-512 SP build: rows are padded with dummy entries to its vector width (16 / 8)
        nn, nm, rows = s.cluster_lists()
        assert np.array_equal(nn, g["t0_nnz"]) and np.array_equal(nm, g["t0_numneigh_masked"])
        assert all(np.array_equal(a, b) for a, b in zip(f.cluster_lists()[2], rows))
        assert_same_structure(f, s, lists=False)
    else:
        assert_same_structure(f, s)
    real_lane = np.isfinite(g["t0_clx"][:len(g["t0_clv"])])
    assert np.array_equal(g["t0_clv"][real_lane], s.cl("v")[real_lane])
    if "t0_clf" in g:
        s.computeForce()
        fr = np.nan_to_num(g["t0_clf"])
        assert np.abs(fr - np.nan_to_num(s.cl("f")))[real_lane].max() <= TOL[dp]
        for n in range(int(g["nsteps"])):
            s.step(n)
        s.updateSingleAtoms()
        assert max_rel(s.atoms("x"), g["tN_x"]) <= TOL[dp]
        T, P = s.thermo()
        assert abs(T - g["tN_thermo"][0]) <= TOL[dp] * T and abs(P - g["tN_thermo"][1]) <= TOL[dp] * P
    s.close()


@pytest.mark.parametrize("dp", [True, False])
def test_createAtom_adjustThermo_on_device(dp):
    for (nx, ny, nz) in [(6, 6, 6), (5, 7, 3)]:
        s = make_cp(dp, 4, nx, ny, nz)
        assert s.createAtom() == 4 * nx * ny * nz
        x, v = initial_atoms(dp, nx, ny, nz)
        assert np.array_equal(s.atoms("x"), x)
        s.setupThermo(); s.adjustThermo()
        assert max_rel(s.atoms("v"), v) <= (1e-12 if dp else 1e-5)
        s.close()


@pytest.mark.parametrize("dp,N,half,fuse_force", [(True, 4, 0, 1), (True, 8, 0, 1), (True, 4, 1, 1), (False, 4, 0, 1), (False, 8, 0, 1),
                                                  (False, 4, 0, 0)])
def test_run_loop_equals_operator_by_operator(dp, N, half, fuse_force):
    """mdb_cp_run (device-resident loop; full lists: integrate halves in the force kernel's epilogue with a second cluster
    position array, else / fuse_force=0 the fused final+initial integrate pass) == the same operators called one by one"""
    x, v = jittered(dp, 6, 6, 6, amp=0.1)
    a, b = make_cp(dp, N, 6, half=half, nstat=25), make_cp(dp, N, 6, half=half, nstat=25)
    a.setOption("fuse_force", fuse_force)
    for s in (a, b):
        s.setAtoms(x, v)
        s.setup(adjust=False)
    rec, _ = a.run(60)
    recs = [(0,) + b.thermo()]
    b.computeForce()
    for n in range(60):
        b.step(n)
        if (n + 1) % 25 == 0 and n + 1 < 60:
            recs.append((n + 1,) + b.thermo())
    b.updateSingleAtoms()
    recs.append((60,) + b.thermo())
    assert np.array_equal(rec[:, 0], [q[0] for q in recs])
    if half:   # reaction forces are accumulated with atomics: summation order is not fixed
        assert np.allclose(rec, np.array(recs), rtol=1e-12)
        assert max_rel(a.atoms("x"), b.atoms("x")) < 1e-12
    else:
        assert np.array_equal(rec, np.array(recs))
        assert np.array_equal(a.atoms("x"), b.atoms("x")) and np.array_equal(a.atoms("v"), b.atoms("v"))
    a.close(); b.close()


@pytest.mark.parametrize("N", [4, 8])
def test_sp_kernel_variants(N):
    """SP full lists: two lanes per i-cluster (sp_kernel 1, default) == lane per i atom (sp_kernel 0) bit for bit, on
    ragged clusters, separate and fused (mdb_cp_run); without the Newton step (sp_kernel 2) within the SP tolerance"""
    x, v = jittered(False, 6, 7, 5, amp=0.1)
    sims = []
    for k in (0, 1, 2):
        s = make_cp(False, N, 6, 7, 5, nstat=25)
        s.setOption("sp_kernel", k)
        s.setAtoms(x, v)
        s.setup(adjust=False)
        s.computeForce()
        sims.append(s)
    f0, f1, f2 = (np.nan_to_num(s.cl("f")) for s in sims)
    assert np.array_equal(f0, f1)
    assert np.abs(f2 - f0).max() <= 2e-6 * np.abs(f0).max()
    recs = [s.run(60)[0] for s in sims]
    assert np.array_equal(recs[0], recs[1])
    assert np.array_equal(sims[0].atoms("x"), sims[1].atoms("x")) and np.array_equal(sims[0].atoms("v"), sims[1].atoms("v"))
    assert np.allclose(recs[2], recs[0], rtol=1e-4)
    assert max_rel(sims[2].atoms("x"), sims[0].atoms("x")) <= 1e-4
    for s in sims:
        s.close()


@pytest.mark.parametrize("dp,N", [(True, 4), (False, 8)])
def test_prune_neighbor_vs_oracle(dp, N):
    """pruneNeighbor 15 steps after the build: the device's list (its row order) and cluster positions are handed to
    the checker, both prune, rows agree entry by entry; forces from the pruned list equal those from the full list"""
    x, v = initial_atoms(dp, 6)
    s, o = make_cp(dp, N, 6), make_oracle(dp, N, 6)
    s.setAtoms(x, v); o.set_atoms(x, v)
    s.setup(adjust=False); o.setup()
    s.computeForce()
    for n in range(15):
        s.step(n)
    nn, nm, nb = s.raw_lists()
    assert o.geti("maxneighs") == s.counts()["maxneighs"]
    o._arr("numneigh", len(nn), np.int32)[:] = nn
    o._arr("numneigh_masked", len(nn), np.int32)[:] = nm
    o._arr("neighbors", nb.size, np.int32)[:] = nb.reshape(-1)
    xs = s.cl("x")
    o._arr("cl_x", xs.size, o.np_real)[:] = xs.reshape(-1)
    s.computeForce()
    f_full = s.cl("f").copy()
    s.pruneNeighbor(); o.pruneNeighbor()
    a, b = s.raw_lists(), o.raw_lists()
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert a[0].sum() < nn.sum()
    for ci in range(len(nn)):
        assert np.array_equal(a[2][ci, :a[0][ci]], b[2][ci, :b[0][ci]])
    s.computeForce()
    assert max_rel(s.cl("f"), f_full) <= (1e-12 if dp else 1e-5)
    s.close()


def test_config2_structural_goldens_and_thermo(golden_dir):
    """BASELINE config 2: Cu FCC 32^3, clusterpair 4x4 SP, 200 steps.  Structure at t = 0 (32x32 columns,
    32768 full i-clusters, 13900 ghost clusters, 1 849 584 cluster pairs, min 47, max 89, 32768 masked) and the thermo
    records of the reference's scalar 4x4 SP build (tests/golden/thermo_cp.json)."""
    s = make_cp(False, 4, 32)
    s.createAtom()
    s.setup(adjust=True)
    c = s.counts()
    p = s.neigh_params()
    assert (p["nbinx"], p["nbiny"]) == (32, 32)
    assert (c["Nclusters_local"], c["Nclusters_ghost"], c["Nghost"], c["dummy_cj"]) == (32768, 13900, 55600, 46668)
    nn, nm, _ = s.raw_lists()
    assert int(nn.sum()) == 1849584 and nn.min() == 47 and nn.max() == 89 and int(nm.sum()) == 32768
    assert np.all(s.iclusters()[0] == 4)
    rec, tm = s.run(200)
    gold = [q for q in json.load(open(os.path.join(golden_dir, "thermo_cp.json"))) if q["variant"] == "cpref44_sp"][0]
    assert np.array_equal(rec[:, 0], [q[0] for q in gold["records"]])
    for got, want in zip(rec, gold["records"]):
        assert abs(got[1] - want[1]) <= 1e-4 * want[1] and abs(got[2] - want[2]) <= 1e-4 * want[2], (got, want)
    # SP trajectories agree to 1e-4 only: a bounding box may end up on the other side of the ghost cutoff
    assert abs(s.counts()["Nclusters_ghost"] - gold["nclusters_ghost"]) <= 0.005 * gold["nclusters_ghost"]
    assert tm["TOTAL"] > 0
    s.close()


@pytest.mark.parametrize("dp,N,half", [(True, 4, 0), (False, 4, 1), (True, 8, 0)])
def test_config2_dp_thermo_and_properties(golden_dir, dp, N, half):
    """size-independent properties at 32^3: every atom in exactly one cluster slot, zero net force at the end (the
    last step follows a rebuild), momentum conserved over 200 steps for M = N.  (With N = 2M the lists are not
    symmetric -- a 4-atom i-cluster against an 8-atom j-cluster -- so an atom that outruns the skin between two
    rebuilds leaves a one-sided pair and the reference's 4x8 scheme itself does not conserve momentum at T = 1.44:
    the checker shows the same drift, profiles/dbg_cp.py.)"""
    s = make_cp(dp, N, 32, half=half)
    s.createAtom()
    s.setup(adjust=True)
    tags = s.cluster_tags()[:s.counts()["ncj"]].reshape(-1)
    assert np.array_equal(np.sort(tags[tags >= 0]), np.arange(131072))
    v0 = s.atoms("v").astype(np.float64).sum(axis=0)
    rec, _ = s.run(200)
    f = np.nan_to_num(s.cl("f").astype(np.float64))
    fmax = np.abs(f).max()
    assert fmax > 1.0
    assert np.all(np.abs(f.sum(axis=(0, 2))) <= (1e-10 if dp else 2e-2) * fmax)
    v1 = s.atoms("v").astype(np.float64).sum(axis=0)
    if N == 4:
        assert np.all(np.abs(v1 - v0) <= (1e-10 if dp else 5e-2))
    if dp and N == 4 and not half:
        gold = [q for q in json.load(open(os.path.join(golden_dir, "thermo_cp.json"))) if q["variant"] == "cpref44_dp"][0]
        for got, want in zip(rec, gold["records"]):
            assert abs(got[1] - want[1]) <= 1e-10 * want[1] and abs(got[2] - want[2]) <= 1e-10 * want[2], (got, want)
    s.close()


def test_maxneighs_overflow_resize():
    """rows longer than maxneighs (100): the build is repeated with 1.2 x the longest row (neighbor.c:412-428)"""
    x, v = initial_atoms(True, 6)
    s, o = make_cp(True, 4, 6, skin=1.6), make_oracle(True, 4, 6, skin=1.6)
    s.setAtoms(x, v); o.set_atoms(x, v)
    s.setup(adjust=False); o.setup()
    assert s.counts()["maxneighs"] == o.geti("maxneighs") > 100
    assert_same_structure(o, s)
    s.close()


# ---- the C driver of the clusterpair scheme: reference command line and report ---------------------------------------
def _cp_driver():
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "md-bench_b200", "driver", "MDBench-CP-B200")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.dirname(exe)])
    return exe


@pytest.mark.parametrize("extra,variant,tol", [([], "cpref44_sp", 1e-4), (["--operators"], "cpref44_sp", 1e-4),
                                               (["--precision", "dp"], "cpref44_dp", 1e-6)])
def test_cp_driver_default_run_prints_reference_report(golden_dir, extra, variant, tol):
    """MDBench-CP-B200 with the reference's defaults = BASELINE config 2 (Cu FCC 32^3, 4x4, 200 steps): the
    step/temp/pressure lines of the reference's scalar 4x4 build (thermo_cp.json; the printed 7 digits in DP)"""
    import re
    import subprocess
    gold = [q for q in json.load(open(os.path.join(golden_dir, "thermo_cp.json"))) if q["variant"] == variant][0]
    out = subprocess.run([_cp_driver()] + extra, capture_output=True, text=True, timeout=300).stdout
    lines = [(int(a), float(b), float(c)) for a, b, c in re.findall(r"^(-?\d+)\t(\S+)\t(\S+)$", out, flags=re.M)]
    assert [l[0] for l in lines] == [0, 100, 200], out
    for (st, T, P), (gs, gT, gP) in zip(lines, gold["records"]):
        assert abs(T - gT) <= tol * gT and abs(P - gP) <= tol * gP, (st, T, P, gT, gP)
    assert re.search(r"System: 131072 atoms \d+ ghost atoms, Steps: 200", out)
    assert "million atom updates per second" in out and "TOTAL" in out and "Kernel: CUDA sm_100a 4x4" in out


def _argon_cp(g, N=4, dp=True, half=0):
    m = load_pkg()
    eps, sig, cutf, skin, dt, temp, rho, mass = [float(v) for v in g["params"]]
    reneigh, nstat = [int(v) for v in g["ints"]]
    b = [float(v) for v in g["box"]]
    p = m.default_params(precision=m.DP if dp else m.SP, epsilon=eps, sigma=sig, cutforce=cutf, skin=skin, dt=dt, temp=temp, rho=rho,
                         mass=mass, reneigh_every=reneigh, nstat=nstat, half_neigh=half, nx=1, ny=1, nz=1, from_input=1,
                         xlo=0.0, xhi=b[0], ylo=0.0, yhi=b[1], zlo=0.0, zhi=b[2])
    return m.ClusterSimulation(p, cluster_n=N)


def test_input_file_box_argon_fixture(golden_dir):
    """clusterpair with a reader's box instead of nx * lattice (neighbor.c:78-82): the reference's scalar 4x4 DP build on
    data/argon (1000 atoms, box 3.6014^3, cutneigh 1.9 > L/2, rows of up to 330 entries -> maxneighs resize), fixture
    cp44_dp_argon.npz: structures and lists bit for bit, forces, atoms after one rebuild"""
    g = np.load(os.path.join(golden_dir, "cp44_dp_argon.npz"))
    s = _argon_cp(g)
    s.setAtoms(g["x0"], g["v0"])
    s.setup(adjust=False)
    assert_same_structure(_Fixture(g), s)
    assert s.counts()["maxneighs"] > 100
    s.computeForce()
    real_lane = np.isfinite(g["t0_clx"][:len(g["t0_clf"])])
    fr = np.nan_to_num(g["t0_clf"])
    assert np.abs(fr - np.nan_to_num(s.cl("f")))[real_lane].max() <= 1e-10 * max(np.abs(fr).max(), 1e-300)
    for n in range(int(g["nsteps"])):
        s.step(n)
    s.updateSingleAtoms()
    assert max_rel(s.atoms("x"), g["tN_x"]) <= 1e-10
    T, P = s.thermo()
    assert abs(T - g["tN_thermo"][0]) <= 1e-10 * T and abs(P - g["tN_thermo"][1]) <= 1e-10 * P
    c = s.counts()
    assert (c["Nclusters_local"], c["Nclusters_ghost"]) == tuple(int(v) for v in g["tN_counts"])
    s.close()


@pytest.mark.parametrize("N,half,dp", [(8, 0, True), (4, 1, True), (4, 0, False)])
def test_input_file_box_argon_vs_oracle(golden_dir, N, half, dp):
    """the other variants on the same input against the checker (pinned to the reference on this input in
    tests/test_cp_oracle_pinned.py::test_input_file_box_argon)"""
    g = np.load(os.path.join(golden_dir, "cp44_dp_argon.npz"))
    eps, sig, cutf, skin, dt, temp, rho, mass = [float(v) for v in g["params"]]
    reneigh, nstat = [int(v) for v in g["ints"]]
    s = _argon_cp(g, N=N, dp=dp, half=half)
    o = OracleCP(dp, N, N)
    o.configure(nx=1, half_neigh=half, epsilon=eps, sigma=sig, cutforce=cutf, skin=skin, dt=dt, temp=temp, rho=rho, mass=mass,
                reneigh_every=reneigh, nstat=nstat)
    o.set_box(*[float(v) for v in g["box"]])
    real = np.float64 if dp else np.float32
    x0, v0 = g["x0"].astype(real), g["v0"].astype(real)
    s.setAtoms(x0, v0); o.set_atoms(x0, v0)
    s.setup(adjust=False); o.setup()
    assert_same_structure(o, s)
    for n in range(101):
        s.step(n); o.step(n)
    assert_same_structure(o, s, tol=TOL[dp])
    s.close()


def test_cp_driver_reads_gro_and_param_file(golden_dir, tmp_path):
    """MDBench-CP-B200 --param <file> -i <input.gro>: inputs rebuilt from the fixture"""
    import re
    import subprocess
    g = np.load(os.path.join(golden_dir, "cp44_dp_argon.npz"))
    eps, sig, cutf, skin, dt, temp, rho, mass = [float(v) for v in g["params"]]
    reneigh, nstat = [int(v) for v in g["ints"]]
    conf = tmp_path / "params.conf"
    conf.write_text("epsilon %.17g\nsigma %.17g\ncutforce %.17g\nskin %.17g\ndt %.17g\ntemp %.17g\nrho %.17g\nmass %.17g\n"
                    "reneigh_every %d\nnstat %d\n" % (eps, sig, cutf, skin, dt, temp, rho, mass, reneigh, nstat))
    gro = tmp_path / "input.gro"
    x, v = g["x0"], g["v0"]
    with open(gro, "w") as f:
        f.write("Liquid Argon t=   0.00000 step= 0\n %d\n" % len(x))
        for i in range(len(x)):
            f.write("%5dAr      Ar%5d%8.3f%8.3f%8.3f%8.4f%8.4f%8.4f\n" % (i + 1, i + 1, x[i, 0], x[i, 1], x[i, 2], v[i, 0], v[i, 1], v[i, 2]))
        f.write("   %.5f   %.5f   %.5f\n" % tuple(float(q) for q in g["box"]))
    out = subprocess.run([_cp_driver(), "--param", str(conf), "-i", str(gro), "--precision", "dp", "-n", "105"], capture_output=True,
                         text=True, timeout=120).stdout
    lines = [(int(a), float(b), float(c)) for a, b, c in re.findall(r"^(-?\d+)\t(\S+)\t(\S+)$", out, flags=re.M)]
    assert [l[0] for l in lines] == [0, 105], out
    # positions / velocities went through the .gro's 3 / 4 decimals: thermo to that precision only
    assert abs(lines[-1][1] - g["tN_thermo"][0]) <= 1e-3 * g["tN_thermo"][0]
    assert re.search(r"System: 1000 atoms \d+ ghost atoms, Steps: 105", out)


@pytest.mark.parametrize("N,pattern", [(4, "seq"), (8, "seq"), (4, "rand"), (4, "fix")])
def test_stub_clusters_lists_and_force(N, pattern):
    """the clusterpair kernel micro-benchmark's synthetic clusters and lists (main-stub.c:227-272, 61-122), and the force
    they give against numpy (all listed pairs except the self pair of the own tile)"""
    m = load_pkg()
    ni, nat, nn, nr = 300, 4, 9, 2
    s = m.ClusterSimulation(m.default_params(nx=1, ny=1, nz=1, cutforce=1.0e6, skin=0.0), cluster_n=N)
    s.stub(ni, nat, pattern, nn, nr)
    c = s.counts()
    assert c["Nclusters_local"] == ni and c["Nlocal"] == ni * nat and c["Nclusters_ghost"] == 0
    x = s.cl("x")
    ncj = c["ncj"]
    atoms = x.transpose(0, 2, 1).reshape(-1, 3)                       # (tile, lane) -> atom, both N: index order
    assert np.allclose(atoms[:ni * nat, 0], np.arange(ni * nat) * 1e-5, rtol=0, atol=1e-18)
    cnt, cm, nb = s.raw_lists()
    assert np.all(cnt == nn * nr) and np.array_equal(nb[:, :nn], nb[:, nn:2 * nn])
    self_tile = np.arange(ni) // (N // 4)
    if pattern == "seq":
        assert np.array_equal(nb[:, :nn], (self_tile[:, None] + np.arange(nn)[None, :]) % ncj)
    elif pattern == "fix":
        assert np.array_equal(nb[:, :nn], np.repeat(np.arange(nn)[None, :], ni, axis=0))
    else:
        assert nb.min() >= 0 and nb.max() < ncj
    s.computeForce()
    f = s.cl("f").transpose(0, 2, 1).reshape(-1, 3)[:ni * nat]
    ref = np.zeros_like(f)
    for ci in range(ni):
        for cj in nb[ci]:
            ja = np.arange(cj * N, (cj + 1) * N)
            for a in range(ci * nat, (ci + 1) * nat):
                d = atoms[a][None, :] - atoms[ja]
                keep = ja != a                                         # exclusion on the own tile only (force_lj.c:99-113)
                rsq = (d * d).sum(axis=1)
                with np.errstate(divide="ignore", invalid="ignore"):
                    sr2 = 1.0 / rsq
                    sr6 = sr2 ** 3
                    ff = np.where(keep, 48.0 * sr6 * (sr6 - 0.5) * sr2, 0.0)
                ref[a] += (d * ff[:, None]).sum(axis=0)
    assert np.abs(f - ref).max() <= 1e-10 * np.abs(ref).max()
    s.close()


def test_cp_stub_driver_report():
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "md-bench_b200", "driver", "MDBench-CP-B200-stub")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.dirname(exe)])
    out = subprocess.run([exe, "-p", "rand", "-ni", "20000", "-nn", "50", "-n", "20"], capture_output=True, text=True, timeout=120).stdout
    assert "Pattern: rand" in out and "Number of i-clusters: 20000" in out and "Mega atom updates/s" in out and "MxN: 4x4" in out
    out = subprocess.run([exe, "-p", "seq", "-ni", "4096", "--csv", "-n", "5", "--cluster-n", "8"], capture_output=True, text=True,
                         timeout=120).stdout
    assert out.splitlines()[0].startswith("steps,pattern,niclusters,iclusters_natoms") and out.splitlines()[1].startswith("5,seq,4096,4,9,1,")


@pytest.mark.parametrize("variant,dp", [("cp44_sp", False), ("cp44_dp", True)])
def test_reference_clusterpair_main_c_drives_libmdb200(golden_dir, variant, dp):
    """The boundary proven with the reference's own clusterpair driver: oracle/_ref/MDBench-<variant>-b200 is the reference's
    UNMODIFIED clusterpair/main.c / atom.c / thermo.c / parameter.c (compiled from /root/reference where they lie) linked against
    md-bench_b200/driver/b200_shim_cp.c + libmdb200.so (oracle/Makefile ref-shim).  Its default run (BASELINE config 2 physics,
    4x4 clusters) must print the thermo lines and the ghost count of the reference's scalar 4x4 build (tests/golden/thermo_cp.json)."""
    import re
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "oracle", "_ref", "MDBench-%s-b200" % variant)
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/MDBench-%s-b200 not built (needs /root/reference: make -C oracle ref-shim)" % variant)
    th = json.load(open(os.path.join(golden_dir, "thermo_cp.json")))
    t = [q for q in th if q["variant"] == ("cpref44_dp" if dp else "cpref44_sp")][0]
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300).stdout
    lines = [(int(a), b, c) for a, b, c in re.findall(r"^(-?\d+)\t(\S+)\t(\S+)$", out, flags=re.M)]
    assert [l[0] for l in lines] == [0, 100, 200], out[-2000:]
    for (st, T, P), (gs, gT, gP) in zip(lines, t["records"]):
        if dp:
            assert "%e" % gT == T and "%e" % gP == P, (st, T, P, gT, gP)      # the printed 7 digits
        else:
            assert abs(float(T) - gT) <= 1e-4 * gT and abs(float(P) - gP) <= 1e-4 * gP, (st, T, P, gT, gP)
    m = re.search(r"System: 131072 atoms (\d+) ghost atoms, Steps: 200", out)
    assert m, out[-2000:]
    if dp:
        assert int(m.group(1)) == t["nghost_atoms"]
    else:   # SP trajectories differ in the last digits: an atom or two may sit on the other side of a ghost threshold
        assert abs(int(m.group(1)) - t["nghost_atoms"]) <= 20
    assert "million atom updates per second" in out
