"""Shared checkers: drive any implementation of the verletlist operators (oracle port, reference
binding, CUDA Simulation) through the reference's driver flow and compare with a golden fixture.

Tolerances (north_star): neighbor lists equal as sorted index sets bit-exactly; ghost maps and
ghost coordinates bit-exact; forces / positions / velocities / thermo within DP rel 1e-10,
SP rel 1e-4 (relative to the largest magnitude of the compared array, because the t=0 lattice
forces are pure cancellation noise).
"""
import numpy as np

TOL = {True: 1e-10, False: 1e-4}


def csr_sets(nn, nb):
    off = np.zeros(len(nn) + 1, np.int64)
    off[1:] = np.cumsum(nn)
    flat = np.concatenate([np.sort(nb[i, :nn[i]]) for i in range(len(nn))]).astype(np.int32) \
        if len(nn) else np.zeros(0, np.int32)
    return off, flat


def rel_err(a, b):
    s = max(np.abs(b).max(), 1e-300)
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / s)


def check_snapshot(impl, g, tag, dp, exact_inputs, f_floor=0.0):
    """impl exposes get('x', ghosts=True), get('v'), get('f'), ghost map, neighbor rows."""
    tol = TOL[dp]
    x = impl.get("x", ghosts=True)
    gx = g[tag + "_x"]
    assert x.shape == gx.shape, (x.shape, gx.shape)
    if hasattr(impl, "ghostMap"):
        gm = impl.ghostMap()
    else:
        gm = {k: impl.get(k) for k in ("border_map", "PBCx", "PBCy", "PBCz")}
    if exact_inputs:
        # same input bits -> same ghost set, same ghost coordinates, same list membership
        for k in ("border_map", "PBCx", "PBCy", "PBCz"):
            assert np.array_equal(gm[k], g[tag + "_" + k]), k
        assert np.array_equal(x, gx), "positions incl. ghosts must be bit-identical"
    else:
        assert rel_err(x, gx) < tol
    if hasattr(impl, "neighbors"):
        nn, nb = impl.neighbors()
    else:
        nn, nb = impl.get("numneigh"), impl.get("neighbors")
    off, flat = csr_sets(nn, nb)
    same_lists = np.array_equal(nn, g[tag + "_numneigh"]) and np.array_equal(flat, g[tag + "_nbr_flat"])
    if exact_inputs:
        assert same_lists, "neighbor lists differ as sorted index sets"
        if getattr(impl, "row_order_exact", True):
            assert np.array_equal(nb[0, :nn[0]], g[tag + "_row0_raw"]), "row order differs"
    f = impl.get("f")
    gf = g[tag + "_f"]
    fs = max(np.abs(gf).max(), f_floor)
    assert np.abs(f - gf).max() <= tol * fs, (np.abs(f - gf).max(), fs)
    assert rel_err(impl.get("v"), g[tag + "_v"]) < tol
    T, P = impl.thermo()
    assert abs(T - g[tag + "_thermo"][0]) <= tol * abs(g[tag + "_thermo"][0])
    assert abs(P - g[tag + "_thermo"][1]) <= tol * abs(g[tag + "_thermo"][1])
    return same_lists


def run_lj_fixture(impl, g, dp, feed, setup_noadjust):
    """feed(x, v): hand the fixture's t=0 local atoms to impl; setup_noadjust(): setup w/o adjustThermo"""
    nl = g["t0_v"].shape[0]
    feed(g["t0_x"][:nl], g["t0_v"])
    setup_noadjust()
    impl.computeForce()
    # largest single pair force magnitude ~ the scale against which t=0 cancellation noise is judged
    check_snapshot(impl, g, "t0", dp, exact_inputs=True, f_floor=1.0)
    for n in range(int(g["nsteps"])):
        impl.step(n)
    same = check_snapshot(impl, g, "tN", dp, exact_inputs=False)
    return same
