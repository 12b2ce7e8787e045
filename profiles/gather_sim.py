"""CPU model of the verletlist force kernel's gathers (no GPU): for a melted LJ box taken from the oracle, count the 32-byte
sectors and 128-byte lines one warp request touches (32 lanes = 32 consecutive atoms, the k-th list entry of each) for the
8-byte z gather and the 16-byte (x, y) gather, under different atom orders and different orders of the entries inside a row.
Used to decide which orderings are worth a GPU A/B (the L1 data pipe of k_force_lj_full_fi is at 88 %).

    python profiles/gather_sim.py --nx 24 --steps 100
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from portbind import OracleVL  # noqa: E402


def lists_from_oracle(nx, steps):
    o = OracleVL(True)
    o.configure(nx=nx)
    o.setup(create=True)
    if steps:
        o.run(steps)
        o.reneighbour()
    x = o.get("x", ghosts=True)
    nn = o.get("numneigh")
    nb = o.get("neighbors")
    nlocal = len(nn)
    return x, nn, nb, nlocal


def request_stats(rows_nn, rows, addr_of, elem_bytes):
    """rows[i] = list of neighbor slots (memory slots) of the atom in memory slot i.  Returns mean sectors / lines per warp
    request and the mean number of requests per warp (= longest row of the warp)."""
    n = len(rows)
    sect = lines = reqs = 0
    for w in range(0, n - 31, 32):
        L = max(rows_nn[w:w + 32])
        for k in range(L):
            a = [addr_of(rows[i][k]) for i in range(w, w + 32) if k < rows_nn[i]]
            a = np.asarray(a, np.int64) * elem_bytes
            sect += len(np.unique(a // 32))
            lines += len(np.unique(a // 128))
            reqs += 1
    return sect / reqs, lines / reqs, reqs / (n // 32)


def evaluate(name, perm, x, nn, nb, nlocal, row_order):
    """perm: memory slot -> original local atom (locals only; ghosts keep their slots behind the locals)"""
    inv = np.empty(nlocal, np.int64)
    inv[perm] = np.arange(nlocal)
    slot = np.concatenate([inv, np.arange(nlocal, len(x))])  # original index -> memory slot
    rows, rnn = [], []
    for s in range(nlocal):
        i = perm[s]
        r = slot[nb[i, :nn[i]]]
        if row_order == "index":
            r = np.sort(r)
        rows.append(r)
        rnn.append(nn[i])
    out = []
    for eb in (8, 16):
        out.append(request_stats(rnn, rows, lambda j: j, eb))
    print("%-44s z(8B): %5.2f sectors %5.2f lines | xy(16B): %5.2f sectors %5.2f lines | req/warp %.1f"
          % (name + " / rows " + row_order, out[0][0], out[0][1], out[1][0], out[1][1], out[0][2]))


def morton(ix, iy, iz):
    def spread(v):
        v = v.astype(np.uint64) & 0x1fffff
        v = (v | v << 32) & 0x1f00000000ffff
        v = (v | v << 16) & 0x1f0000ff0000ff
        v = (v | v << 8) & 0x100f00f00f00f00f
        v = (v | v << 4) & 0x10c30c30c30c30c3
        v = (v | v << 2) & 0x1249249249249249
        return v
    return spread(ix) | spread(iy) << 1 | spread(iz) << 2


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=24)
    ap.add_argument("--steps", type=int, default=100)
    a = ap.parse_args()
    x, nn, nb, nlocal = lists_from_oracle(a.nx, a.steps)
    print("atoms %d (+%d ghosts), mean listed %.1f" % (nlocal, len(x) - nlocal, nn.mean()))
    ident = np.arange(nlocal)
    evaluate("generator order", ident, x, nn, nb, nlocal, "stencil")
    evaluate("generator order", ident, x, nn, nb, nlocal, "index")
    xl = x[:nlocal]
    L = xl.max() + 1e-9
    for cell in (1.4, 1.0, 0.7):
        g = np.floor(xl / cell).astype(np.int64)
        nxg = int(L / cell) + 2
        key = (g[:, 2] * nxg + g[:, 1]) * nxg + g[:, 0]
        p = np.argsort(key, kind="stable")
        evaluate("x-fastest cells of %.1f" % cell, p, x, nn, nb, nlocal, "index")
        p = np.argsort(morton(g[:, 0], g[:, 1], g[:, 2]), kind="stable")
        evaluate("Morton cells of %.1f" % cell, p, x, nn, nb, nlocal, "index")




def duo_stats(nx=24, steps=100):
    """Pair rows ("duo"): lane t owns atoms a(t), b(t); entries listed by both are gathered once.  Prints gathers and
    warp iterations relative to the one-atom-per-lane kernel."""
    x, nn, nb, nlocal = lists_from_oracle(nx, steps)
    sets = [set(nb[i, :nn[i]].tolist()) for i in range(nlocal)]
    base_iter = sum(max(nn[w:w + 32]) for w in range(0, nlocal - 31, 32))  # warp iterations now (1 gather, 1 pair each)
    for name, S in (("(2t, 2t+1)", 1), ("(m, m+4) generator rows", 4), ("(m, m+2)", 2), ("(m,m+16)", 16)):
        npairs = nlocal // 2
        a = np.array([2 * S * (t // S) + t % S for t in range(npairs)])
        b = a + S
        both = np.array([len(sets[i] & sets[j]) for i, j in zip(a, b)])
        d = x[a] - x[b]
        L = x[:nlocal].max()
        d -= np.round(d / L) * L
        dist = np.sqrt((d * d).sum(1))
        g_ideal = (nn[a] + nn[b] - both).sum()
        it1 = it23 = g = 0
        for w in range(0, npairs - 31, 32):
            n1 = both[w:w + 32].min()
            ra = nn[a[w:w + 32]] - n1
            rb = nn[b[w:w + 32]] - n1
            m = np.maximum(ra, rb).max()
            it1 += n1
            it23 += m
            g += 32 * (n1 + 2 * m)
        print("pairs %-26s dist %.2f  both mean %.1f min %d  | gathers/pairs: ideal %.3f, warp-uniform %.3f of now | "
              "pair evaluations issued %.3f of now" % (name, dist.mean(), both.mean(), both.min(),
              g_ideal / nn.sum(), g / (32.0 * base_iter), (2 * it1 + 2 * it23) / float(base_iter)))


if __name__ == "__main__":
    if os.environ.get("DUO"):
        duo_stats(int(os.environ.get("NX", 16)), int(os.environ.get("STEPS", 100)))
    else:
        main()
