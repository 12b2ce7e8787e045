"""Spatial decomposition: the reference runs one domain, so parity is against the
single-domain path (itself pinned to the oracle/reference) atom by atom through GLOBAL tags.

CPU part: the host-side plan every rank derives (slots, peers, transfer schedule), including a
world_size-2 gloo run that executes the schedule with real sends/receives.
GPU part (-m gpu): a decomposed box on ONE GPU (bricks exchange by device copies; the multi-process
NCCL transport runs the same schedule) against the single-domain run and the oracle.
"""
import json
import os

import numpy as np
import pytest

from conftest import load_pkg

IMG = [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1), (1, 1, 1), (1, -1, 1), (1, 1, -1),
       (1, -1, -1), (-1, 1, 1), (-1, -1, 1), (-1, 1, -1), (-1, -1, -1), (1, 0, 1), (1, 0, -1), (-1, 0, 1),
       (-1, 0, -1), (0, 1, 1), (0, 1, -1), (0, -1, 1), (0, -1, -1), (1, 1, 0), (-1, 1, 0), (1, -1, 0), (-1, -1, 0)]


def brick_at(grid, b, shift, periodic=(1, 1, 1)):
    c = [b % grid[0], (b // grid[0]) % grid[1], b // (grid[0] * grid[1])]
    for a in range(3):
        v = c[a] + shift[a]
        if v < 0 or v >= grid[a]:
            if not periodic[a]:
                return -1
            v %= grid[a]
        c[a] = v
    return c[0] + grid[0] * (c[1] + grid[1] * c[2])


@pytest.mark.parametrize("grid,periodic", [((1, 1, 1), (1, 1, 1)), ((2, 1, 1), (1, 1, 1)), ((2, 2, 2), (1, 1, 1)),
                                           ((3, 2, 4), (1, 1, 1)), ((3, 2, 2), (0, 1, 1)), ((2, 2, 1), (1, 0, 0))])
def test_plan_is_symmetric_and_ordered(grid, periodic):
    m = load_pkg()
    nb = grid[0] * grid[1] * grid[2]
    for b in range(nb):
        send = m.dd_plan(grid, b, True, periodic=periodic)
        recv = m.dd_plan(grid, b, False, periodic=periodic)
        # image d of brick b (shift s_d) is seen by the brick at coords - s_d, and arrives from coords + s_d
        assert sorted(d for d, _, _ in send) == [d for d in range(26) if brick_at(grid, b, [-s for s in IMG[d]], periodic) >= 0]
        for d, p, _ in send:
            assert p == brick_at(grid, b, [-s for s in IMG[d]], periodic)
            assert (d, b) in [(dd, pp) for dd, pp, _ in m.dd_plan(grid, p, False, periodic=periodic)]
        for d, p, _ in recv:
            assert p == brick_at(grid, b, IMG[d], periodic)
        for lst in (send, recv):
            assert [(p, d) for d, p, _ in lst] == sorted((p, d) for d, p, _ in lst)
    if all(periodic):
        assert len(m.dd_plan(grid, 0, True)) == 26


def expected_incoming(m, grid, cnt, r):
    """(sender, direction, n) in the order brick r lays out what it receives"""
    return [(p, d, int(cnt[p, d])) for d, p, _ in m.dd_plan(grid, r, False)]


@pytest.mark.parametrize("grid,nprocs", [((2, 2, 2), 1), ((2, 2, 2), 2), ((2, 2, 2), 8), ((2, 1, 1), 2), ((4, 2, 1), 4)])
def test_schedule_covers_every_segment_once(grid, nprocs):
    m = load_pkg()
    rng = np.random.default_rng(5)
    nb = grid[0] * grid[1] * grid[2]
    cnt = rng.integers(0, 40, (nb, 26)).astype(np.int32)
    cnt[rng.random((nb, 26)) < 0.2] = 0
    sends, recvs = {}, {}
    for proc in range(nprocs):
        for op in m.dd_schedule(grid, nprocs, proc, cnt):
            key = (op["src"], op["dst"])
            if op["kind"] in (0, 1):
                assert key not in sends
                sends[key] = (proc, op)
            if op["kind"] in (0, 2):
                assert key not in recvs
                recvs[key] = (proc, op)
    assert sends.keys() == recvs.keys()
    per = nb // nprocs
    for (s, r), (proc, op) in sends.items():
        assert proc == s // per and recvs[(s, r)][0] == r // per
        inc = expected_incoming(m, grid, cnt, r)
        start = sum(n for p, d, n in inc if p < s)
        length = sum(n for p, d, n in inc if p == s)
        assert (op["dst_start"], op["len"]) == (start, length)
        out = [(p, d, int(cnt[s, d])) for d, p, _ in m.dd_plan(grid, s, True)]
        assert op["src_start"] == sum(n for p, d, n in out if p < r)
    # every non-empty (sender, receiver) pair is scheduled
    for r in range(nb):
        for p in set(p for p, d, n in expected_incoming(m, grid, cnt, r) if n):
            assert (p, r) in sends


def _gloo_worker(rank, world, port, grid, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = load_pkg()
        nb = grid[0] * grid[1] * grid[2]
        per = nb // world
        rng = np.random.default_rng(100 + rank)
        mine = range(rank * per, (rank + 1) * per)
        # slot counts of my bricks -> allgather (what DomainGroup::gather_offsets does over NCCL)
        local = torch.from_numpy(rng.integers(0, 30, (per, 26)).astype(np.int32))
        parts = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(parts, local)
        cnt = torch.cat(parts).numpy()
        # send list of brick s = entries (s, direction, k) in slot order, encoded as one int64 each
        def send_list(s):
            out = []
            for d, p, _ in m.dd_plan(grid, s, True, nprocs=world):
                out += [s * 10**6 + d * 10**4 + k for k in range(cnt[s, d])]
            return torch.tensor(out, dtype=torch.int64)
        sbuf = {s: send_list(s) for s in mine}
        rbuf = {r: torch.full((int(sum(cnt[p, d] for d, p, _ in m.dd_plan(grid, r, False, nprocs=world))),), -1,
                              dtype=torch.int64) for r in mine}
        reqs = []
        for op in m.dd_schedule(grid, world, rank, cnt):
            if op["kind"] == 0:
                rbuf[op["dst"]][op["dst_start"]:op["dst_start"] + op["len"]] = \
                    sbuf[op["src"]][op["src_start"]:op["src_start"] + op["len"]]
            elif op["kind"] == 1:
                reqs.append(dist.isend(sbuf[op["src"]][op["src_start"]:op["src_start"] + op["len"]].clone(), op["peer_proc"]))
            else:
                reqs.append(dist.irecv(rbuf[op["dst"]][op["dst_start"]:op["dst_start"] + op["len"]], op["peer_proc"]))
        for rq in reqs:
            rq.wait()
        for r in mine:
            want = []
            for d, p, _ in m.dd_plan(grid, r, False, nprocs=world):
                want += [p * 10**6 + d * 10**4 + k for k in range(cnt[p, d])]
            assert rbuf[r].tolist() == want, "brick %d received the wrong entries" % r
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("grid", [(2, 2, 2), (2, 1, 1)])
def test_schedule_executes_over_gloo_world2(grid):
    """two processes run the schedule with real point-to-point transfers (gloo, CPU): every brick ends up
    with exactly the entries its ghost area expects, in order, and nothing deadlocks"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + (7 if grid[1] == 1 else 0)
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, grid, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


# ---------------------------------------------------------------------------------------------------
# GPU: decomposed box on one device vs the single-domain path
def single(m, dp=True, **kw):
    s = m.Simulation(m.default_params(precision=m.DP if dp else m.SP, layout=m.SOA, **kw))
    s.setOption("sort_atoms", 0)
    return s


def neighbor_tag_sets_single(s):
    """single-domain lists as sorted lists of global tags (ghost -> source atom; index == tag there)"""
    nn, nb = s.neighbors()
    nl = len(nn)
    bm = s.ghostMap()["border_map"]
    full = np.concatenate([np.arange(nl, dtype=np.int64), bm.astype(np.int64)])
    return [np.sort(full[nb[i, :nn[i]]]) for i in range(nl)]


def min_image(d, box):
    return d - box * np.round(d / box)


@pytest.mark.gpu
@pytest.mark.parametrize("grid,dp,half", [((2, 2, 2), True, 0), ((2, 1, 1), True, 0), ((1, 2, 2), True, 1),
                                          ((1, 1, 1), True, 0), ((2, 2, 2), False, 0), ((2, 2, 1), False, 1)])
def test_dd_setup_matches_single_domain(grid, dp, half):
    m = load_pkg()
    nx = 12
    kw = dict(nx=nx, ny=nx, nz=nx, half_neigh=half)
    s = single(m, dp, **kw)
    s.createAtom(); s.setup(adjust=True); s.computeForce()
    d = m.Decomposition(m.default_params(precision=m.DP if dp else m.SP, **kw), grid)
    assert d.createAtom() == 4 * nx ** 3
    d.setup(adjust=True)
    c = d.counts()
    assert c["Nlocal"] == 4 * nx ** 3 and c["bricks"] == grid[0] * grid[1] * grid[2]
    tags, x = d.get("x")
    assert np.array_equal(np.sort(tags), np.arange(4 * nx ** 3))
    eps = 1e-13 if dp else 1e-5
    box = s.neighborParams()["xprd"]
    assert np.abs(x - s.get("x")[tags]).max() <= eps * box
    _, v = d.get("v")
    assert np.abs(v - s.get("v")[tags]).max() <= (1e-12 if dp else 1e-4) * np.abs(v).max()
    # neighbor lists as multisets of global tags
    ref_sets = neighbor_tag_sets_single(s)
    t2, nn, rows = d.neighborTags()
    assert np.array_equal(t2, tags)
    if not half:
        for i in range(len(t2)):
            assert np.array_equal(np.sort(rows[i, :nn[i]]), ref_sets[t2[i]]), "atom tag %d" % t2[i]
    else:
        # half lists keep local pairs once (which side depends on the internal order) and every ghost pair:
        # compare the symmetrised pair multiset
        def pairs(tg, sets):
            p = np.concatenate([np.stack([np.full(len(q), t), q], 1) for t, q in zip(tg, sets)])
            return p
        a = pairs(t2, [rows[i, :nn[i]] for i in range(len(t2))])
        b = pairs(np.arange(len(ref_sets)), ref_sets)
        # ghost pairs appear from both sides, local pairs once: symmetrise and count
        def sym(p):
            lo, hi = np.minimum(p[:, 0], p[:, 1]), np.maximum(p[:, 0], p[:, 1])
            return np.sort(lo.astype(np.int64) * 10**7 + hi)
        sa, sb = sym(a), sym(b)
        assert np.array_equal(np.unique(sa), np.unique(sb))
    s.close(); d.close()


@pytest.mark.gpu
@pytest.mark.parametrize("grid,dp,half", [((2, 2, 2), True, 0), ((2, 2, 1), True, 1), ((2, 1, 2), False, 0)])
def test_dd_200_steps_match_single_domain_and_oracle(grid, dp, half):
    """200 steps incl. 10 rebuilds with atoms migrating between bricks: x, v, thermo vs the single-domain
    run (tolerance of north_star: DP rel 1e-10, SP rel 1e-4); DP full also vs the oracle"""
    from portbind import OracleVL
    m = load_pkg()
    nx = 12
    kw = dict(nx=nx, ny=nx, nz=nx, half_neigh=half)
    s = single(m, dp, **kw)
    s.createAtom(); s.setup(adjust=True)
    rec_s, _ = s.run(200)
    d = m.Decomposition(m.default_params(precision=m.DP if dp else m.SP, **kw), grid)
    d.createAtom(); d.setup(adjust=True)
    tags0, x0 = d.get("x")
    rec_d, _ = d.run(200)
    tol = 1e-10 if dp else 1e-4
    assert rec_d.shape == rec_s.shape and np.array_equal(rec_d[:, 0], rec_s[:, 0])
    assert np.abs(rec_d[:, 1:] - rec_s[:, 1:]).max() <= tol * np.abs(rec_s[:, 1:]).max()
    tags, x = d.get("x")
    _, v = d.get("v")
    box = s.neighborParams()["xprd"]
    sx, sv = s.get("x"), s.get("v")
    assert np.array_equal(np.sort(tags), np.arange(4 * nx ** 3)), "atoms lost or duplicated in migration"
    # chaotic growth of rounding differences over 200 steps stays far below these bounds (same as the
    # single-domain vs oracle test)
    xtol = 1e-10 if dp else 2e-3
    assert np.abs(min_image(x - sx[tags], box)).max() <= xtol * box
    assert np.abs(v - sv[tags]).max() <= (1e-10 if dp else 2e-2) * np.abs(sv).max()
    # atoms did change bricks: tags owned by the first brick (slot order is brick by brick) before / after
    if grid != (1, 1, 1):
        lo = np.zeros(3)
        ext = box / np.array(grid)
        def in_brick0(xx):
            w = np.mod(xx, box)
            return np.all((w >= lo) & (w < ext), axis=1)
        before, after = set(tags0[in_brick0(x0)].tolist()), set(tags[in_brick0(x)].tolist())
        assert before != after, "no atom crossed a brick face in 200 steps?"
    if dp and not half:
        o = OracleVL(True)
        o.configure(nx=nx, ny=nx, nz=nx)
        o.setup(create=True)
        orec = o.run(200)
        assert abs(rec_d[-1][1] - orec[-1][1]) <= 1e-10 * orec[-1][1]
        assert np.abs(min_image(x - o.get("x")[tags], box)).max() <= 1e-10 * box
    s.close(); d.close()


@pytest.mark.gpu
def test_dd_thermo_golden_config1(golden_dir):
    """BASELINE config 1 (Cu FCC 32^3, 200 steps) cut into 2x2x2 bricks prints the reference's thermo lines"""
    m = load_pkg()
    g = [e for e in json.load(open(os.path.join(golden_dir, "thermo_lj.json")))
         if e["variant"] == "vl_dp_aos" and e["nx"] == 32 and e["half"] == 0][0]
    d = m.Decomposition(m.default_params(), (2, 2, 2))
    d.createAtom(); d.setup(adjust=True)
    rec, _ = d.run(200)
    want = {int(r[0]): (r[1], r[2]) for r in g["records"]}
    assert len(rec) == len(want)
    for step, T, P in rec:   # the reference prints %e (7 significant digits)
        wt, wp = want[int(step)]
        assert float("%e" % T) == wt and float("%e" % P) == wp, (step, T, P, wt, wp)
    assert abs(rec[-1][1] - g["T_full"]) <= 1e-10 * g["T_full"] and abs(rec[-1][2] - g["P_full"]) <= 1e-10 * g["P_full"]
    d.close()


@pytest.mark.gpu
def test_dd_eam_matches_single_domain(golden_dir):
    from cases import funcfl_args
    m = load_pkg()
    g = np.load(os.path.join(golden_dir, "eam_cu_nx5.npz"))
    nx = 8
    kw = dict(force_field=m.FF_EAM, nx=nx, ny=nx, nz=nx, ntimes=60)
    s = single(m, True, **kw)
    s.setEam(*funcfl_args(g)); s.createAtom(); s.setup(adjust=True)
    rec_s, _ = s.run(60)
    d = m.Decomposition(m.default_params(**kw), (2, 2, 1))
    d.setEam(*funcfl_args(g)); d.createAtom(); d.setup(adjust=True)
    rec_d, _ = d.run(60)
    assert np.abs(rec_d[:, 1:] - rec_s[:, 1:]).max() <= 1e-10 * np.abs(rec_s[:, 1:]).max()
    tags, v = d.get("v")
    assert np.abs(v - s.get("v")[tags]).max() <= 1e-10 * np.abs(v).max()
    s.close(); d.close()


@pytest.mark.gpu
@pytest.mark.parametrize("grid,dp", [((2, 2, 2), True), ((1, 2, 1), False), ((1, 1, 1), True)])
def test_dd_fused_force_step_is_bit_identical(grid, dp):
    """decomposed mdb_run with the integrate halves in the force kernel's epilogue (x, y, z updated in place, gathers from
    the double-buffered packed copies whose ghost range is refreshed after every halo) == the separate kernels, bit for bit:
    65 steps with rebuilds + migration at 20/40/60 and a thermo record at 25 and 50"""
    m = load_pkg()
    out = []
    for fuse in (1, 0):
        d = m.Decomposition(m.default_params(precision=m.DP if dp else m.SP, nx=12, ny=12, nz=12, nstat=25), grid)
        d.setOption("fuse_force", fuse)
        d.createAtom(); d.setup(adjust=True)
        rec, _ = d.run(65)
        rec2, _ = d.run(9)      # re-entry with the gather copies swapped an odd number of times
        out.append((rec, rec2, d.get("x"), d.get("v")))
        d.close()
    a, b = out
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert np.array_equal(a[2][0], b[2][0]) and np.array_equal(a[2][1], b[2][1])
    assert np.array_equal(a[3][1], b[3][1])


@pytest.mark.gpu
def test_dd_save_restore_is_bit_reproducible():
    m = load_pkg()
    d = m.Decomposition(m.default_params(nx=12, ny=12, nz=12), (2, 2, 2))
    d.createAtom(); d.setup(adjust=True); d.saveState()
    r1, _ = d.run(60)
    t1, x1 = d.get("x")
    d.restoreState(); d.setup(adjust=False)
    r2, _ = d.run(60)
    t2, x2 = d.get("x")
    assert np.array_equal(r1, r2) and np.array_equal(t1, t2) and np.array_equal(x1, x2)
    d.close()


@pytest.mark.gpu
@pytest.mark.parametrize("grid", [(2, 1, 1), (2, 2, 2)])
def test_dd_nccl_processes(grid):
    """one process per GPU, bricks exchange by NCCL send/recv (needs >= 2 GPUs; skipped on a 1-GPU box)"""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29650 + grid[1]), os.path.join(here, "dd_nccl_worker.py")] + [str(g) for g in grid] + ["12", "60"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DD_NCCL_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


# ---- clusterpair scheme on the same brick grid (cp_dd.cuh) ------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("grid,dp,N,skin", [((2, 2, 2), True, 4, 0.9), ((2, 1, 1), True, 8, 0.9), ((1, 1, 1), True, 4, 0.9),
                                            ((2, 2, 2), True, 4, 0.3), ((1, 2, 2), False, 4, 0.3), ((2, 2, 1), False, 8, 0.3)])
def test_cp_dd_matches_single_domain(grid, dp, N, skin):
    """clusterpair decomposition (ghost clusters from the neighbor bricks, atoms migrating at the rebuilds) against the single
    clusterpair domain atom by atom through global tags: 100 steps = 5 rebuilds, thermo records incl. the intermediate one.
    DP at rel 1e-10 needs a skin no atom outruns between two rebuilds (0.9): a cluster-pair list covers more atom pairs than
    the cutoff sphere (bounding boxes), so WHICH late pairs are still caught when an atom outruns the default skin (at T = 1.44
    the fastest do) depends on the clustering, which differs between a brick and the whole box; with the default skin the two
    runs then agree to ~1e-6, asserted at 1e-5."""
    m = load_pkg()
    nx = 12
    kw = dict(nx=nx, ny=nx, nz=nx, nstat=50, skin=skin)
    P = lambda: m.default_params(precision=m.DP if dp else m.SP, **kw)
    s = m.ClusterSimulation(P(), cluster_n=N)
    s.createAtom(); s.setup(adjust=True)
    rec_s, _ = s.run(100)
    d = m.Decomposition(P(), grid, cluster_n=N)
    assert d.createAtom() == 4 * nx ** 3
    d.setup(adjust=True)
    c = d.counts()
    assert c["Nlocal"] == 4 * nx ** 3 and c["bricks"] == grid[0] * grid[1] * grid[2] and c["Nghost"] > 0
    tags0, x0 = d.get("x")
    rec_d, _ = d.run(100)
    tol = (1e-10 if skin > 0.5 else 1e-5) if dp else 1e-4
    assert rec_d.shape == rec_s.shape and np.array_equal(rec_d[:, 0], rec_s[:, 0])
    assert np.abs(rec_d[:, 1:] - rec_s[:, 1:]).max() <= tol * np.abs(rec_s[:, 1:]).max()
    tags, x = d.get("x")
    _, v = d.get("v")
    assert np.array_equal(np.sort(tags), np.arange(4 * nx ** 3)), "atoms lost or duplicated in migration"
    xs, ts = s.atoms("x", tags=True)
    vs = s.atoms("v")
    sx, sv = np.empty_like(xs), np.empty_like(vs)
    sx[ts], sv[ts] = xs, vs
    box = (4.0 / 0.8442) ** (1.0 / 3.0) * nx
    strict = dp and skin > 0.5
    assert np.abs(min_image(x - sx[tags], box)).max() <= (1e-10 if strict else 2e-3) * box
    assert np.abs(v - sv[tags]).max() <= (1e-10 if strict else 2e-2) * np.abs(sv).max()
    if grid != (1, 1, 1):
        ext = box / np.array(grid)
        def in_brick0(xx):
            w = np.mod(xx, box)
            return np.all((w >= 0) & (w < ext), axis=1)
        assert set(tags0[in_brick0(x0)].tolist()) != set(tags[in_brick0(x)].tolist()), "no atom crossed a brick face in 100 steps?"
    s.close(); d.close()


@pytest.mark.gpu
def test_cp_dd_save_restore_and_set_atoms():
    """restoreState + setup reproduces a run bit for bit; atoms handed over from host buffers (what the bench's end-to-end leg
    does) give the same run as the generated ones"""
    m = load_pkg()
    P = m.default_params(precision=m.SP, nx=12, ny=12, nz=12)
    d = m.Decomposition(P, (2, 1, 2), cluster_n=4)
    d.createAtom(); d.setup(adjust=True)
    d.saveState()
    tags, x = d.get("x")
    _, v = d.get("v")
    r1, _ = d.run(45)
    t1, x1 = d.get("x")
    d.restoreState(); d.setup(adjust=False)
    r2, _ = d.run(45)
    t2, x2 = d.get("x")
    assert np.array_equal(r1, r2) and np.array_equal(t1, t2) and np.array_equal(x1, x2)
    d.setAtoms(tags, np.ascontiguousarray(x.T), np.ascontiguousarray(v.T)); d.setup(adjust=False)
    r3, _ = d.run(45)
    assert np.allclose(r3, r1, rtol=1e-5)
    d.close()


@pytest.mark.gpu
@pytest.mark.parametrize("grid", [(2, 1, 1), (2, 2, 2)])
def test_cp_dd_nccl_processes(grid):
    """clusterpair bricks, one process per GPU, NCCL send/recv (needs >= 2 GPUs; skipped on a 1-GPU box)"""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29670 + grid[1]), os.path.join(here, "dd_nccl_worker.py")] + [str(g) for g in grid] + ["12", "60", "4"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DD_NCCL_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
