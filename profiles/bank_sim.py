"""CPU model of L1 data-pipe wavefronts of the force kernel's gathers (no GPU).  Model (fits the ncu counters of round 1:
6.4 wavefronts per 8-byte gather, 9.6 per 16-byte gather): a warp load is served in passes of 128 bytes of lanes (16 lanes
x 8 B, 8 lanes x 16 B); inside a pass lanes that hit the same 4-byte banks of DIFFERENT lines serialise, so a pass costs the
largest number of distinct addresses that share a bank = (address / elem) mod lanes_per_pass.
    python profiles/bank_sim.py [nx] [steps]
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "profiles"))
from gather_sim import lists_from_oracle


def pass_cost(js, nbanks):
    """js: indices of the active lanes of one pass"""
    js = np.unique(js)
    if len(js) == 0:
        return 0
    return np.bincount(js % nbanks, minlength=nbanks).max()


def warp_cost(rows, w):
    """rows: list of 32 arrays (one per lane, -1 = idle slot).  Returns (z wavefronts, xy wavefronts, iterations)"""
    L = max(len(r) for r in rows)
    cz = cxy = 0
    for k in range(L):
        col = np.array([r[k] if k < len(r) else -1 for r in rows])
        for h in range(2):
            a = col[16 * h:16 * h + 16]
            cz += pass_cost(a[a >= 0], 16)
        for q in range(4):
            a = col[8 * q:8 * q + 8]
            cxy += pass_cost(a[a >= 0], 8)
    return cz, cxy, L


def order_rotate(row, lane, nb=16):
    """entry k should have j % nb == (lane + k) % nb; leftovers fill the holes in order"""
    buckets = [[] for _ in range(nb)]
    for j in row:
        buckets[j % nb].append(j)
    out = []
    n = len(row)
    for k in range(n):
        b = buckets[(lane + k) % nb]
        out.append(b.pop() if b else -2)
    left = [j for b in buckets for j in b]
    for k in range(n):
        if out[k] == -2:
            out[k] = left.pop()
    return np.array(out)


def order_rotate_pad(row, lane, nb=16, maxlen=96):
    """same, but holes stay idle (-1) while the row is shorter than maxlen; leftovers go to the end / remaining holes"""
    buckets = [[] for _ in range(nb)]
    for j in row:
        buckets[j % nb].append(j)
    out = []
    k = 0
    while any(buckets) and k < maxlen:
        b = buckets[(lane + k) % nb]
        out.append(b.pop() if b else -1)
        k += 1
    left = [j for b in buckets for j in b]
    for k in range(len(out)):
        if out[k] == -1 and left:
            out[k] = left.pop()
    out += left
    while out and out[-1] == -1:
        out.pop()
    return np.array(out)


def main():
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    x, nn, nb, nlocal = lists_from_oracle(nx, steps)
    nw = min(nlocal // 32, 120)
    res = {}
    for name, fn in (("reference order", lambda r, l: r), ("rotate mod 16", order_rotate), ("rotate mod 8", lambda r, l: order_rotate(r, l, 8)),
                     ("rotate mod 16, idle holes <= 88", lambda r, l: order_rotate_pad(r, l, 16, 88)),
                     ("rotate mod 16, idle holes <= 96", lambda r, l: order_rotate_pad(r, l, 16, 96))):
        tz = txy = it = 0
        for w in range(nw):
            rows = [fn(nb[i, :nn[i]], i % 32) for i in range(32 * w, 32 * w + 32)]
            a, b, c = warp_cost(rows, w)
            tz += a; txy += b; it += c
        base = res.setdefault("it", it)
        print("%-36s per warp iteration: z %.2f  xy %.2f wavefronts | iterations %.3f of reference | per listed pair-warp: %.2f"
              % (name, tz / it, txy / it, it / base, (tz + txy + it) / base))




def order_warp_repair(rows32, prefer_big=False):
    """cooperative schedule of one warp (32 rows): entry k of lane l prefers class (l + k) % 16; lanes whose preferred bucket
    is empty take a class that is free in their half warp (mod 16, z gather) and quarter warp (mod 8, xy gather)."""
    buckets = [[[] for _ in range(16)] for _ in range(32)]
    for l, row in enumerate(rows32):
        for j in row:
            buckets[l][j % 16].append(int(j))
    left = [len(r) for r in rows32]
    out = [[] for _ in range(32)]
    k = 0
    while any(left):
        for h in range(2):
            lanes = [l for l in range(16 * h, 16 * h + 16) if left[l] > 0]
            taken16 = set()
            taken8 = {0: set(), 1: set()}
            unmatched = []
            for l in lanes:
                r = (l + k) % 16
                if buckets[l][r]:
                    out[l].append(buckets[l][r].pop()); left[l] -= 1
                    taken16.add(r); taken8[(l % 16) // 8].add(r % 8)
                else:
                    unmatched.append(l)
            for l in unmatched:
                q = (l % 16) // 8
                cand = [r for r in range(16) if buckets[l][r]]
                def score(r):
                    return ((r not in taken16) * 2 + (r % 8 not in taken8[q]), len(buckets[l][r]))
                r = max(cand, key=score)
                out[l].append(buckets[l][r].pop()); left[l] -= 1
                taken16.add(r); taken8[q].add(r % 8)
        k += 1
    return [np.array(o) for o in out]


def order_warp_matching(rows32):
    """greedy per step: lanes in order of fewest remaining options pick the largest own bucket among the free classes"""
    buckets = [[[] for _ in range(16)] for _ in range(32)]
    for l, row in enumerate(rows32):
        for j in row:
            buckets[l][j % 16].append(int(j))
    left = [len(r) for r in rows32]
    out = [[] for _ in range(32)]
    k = 0
    while any(left):
        for h in range(2):
            lanes = [l for l in range(16 * h, 16 * h + 16) if left[l] > 0]
            taken16 = set()
            taken8 = {0: set(), 1: set()}
            # longest rows first: they define the trip count and must not be left with conflicts at the end
            for l in sorted(lanes, key=lambda l: -left[l]):
                q = (l % 16) // 8
                cand = [r for r in range(16) if buckets[l][r]]
                def score(r):
                    return ((r not in taken16) * 2 + (r % 8 not in taken8[q]), len(buckets[l][r]), (r - l - k) % 16 == 0)
                r = max(cand, key=score)
                out[l].append(buckets[l][r].pop()); left[l] -= 1
                taken16.add(r); taken8[q].add(r % 8)
        k += 1
    return [np.array(o) for o in out]


def main2():
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    x, nn, nb, nlocal = lists_from_oracle(nx, steps)
    nw = min(nlocal // 32, 60)
    for name, fn in (("reference order", None), ("rotate + repair", order_warp_repair), ("greedy matching", order_warp_matching), ("rotate + parallel repair", lambda r: order_warp_parallel_repair(r))):
        tz = txy = it = 0
        for w in range(nw):
            rows = [nb[i, :nn[i]] for i in range(32 * w, 32 * w + 32)]
            if fn:
                rows = fn(rows)
            a, b, c = warp_cost(rows, w)
            tz += a; txy += b; it += c
        print("%-36s per warp iteration: z %.2f  xy %.2f wavefronts (conflict-free: 2 and 4)" % (name, tz / it, txy / it))




def order_warp_parallel_repair(rows32):
    """like order_warp_repair, but the unmatched lanes of a step choose at the same time, seeing only what the matched lanes
    took; candidate = first free class counted from the preferred one (what the GPU kernel does)"""
    buckets = [[[] for _ in range(16)] for _ in range(32)]
    for l, row in enumerate(rows32):
        for j in row:
            buckets[l][j % 16].append(int(j))
    left = [len(r) for r in rows32]
    out = [[] for _ in range(32)]
    k = 0
    while any(left):
        for h in range(2):
            lanes = [l for l in range(16 * h, 16 * h + 16) if left[l] > 0]
            taken16 = set()
            taken8 = {0: set(), 1: set()}
            unmatched = []
            for l in lanes:
                r = (l + k) % 16
                if buckets[l][r]:
                    out[l].append(buckets[l][r].pop()); left[l] -= 1
                    taken16.add(r); taken8[(l % 16) // 8].add(r % 8)
                else:
                    unmatched.append(l)
            for l in unmatched:
                q = (l % 16) // 8
                order = [(l + k + d) % 16 for d in range(16)]
                cand = [r for r in order if buckets[l][r]]
                best = None
                for test in (lambda r: r not in taken16 and r % 8 not in taken8[q], lambda r: r not in taken16,
                             lambda r: r % 8 not in taken8[q], lambda r: True):
                    c = [r for r in cand if test(r)]
                    if c:
                        best = c[0]
                        break
                out[l].append(buckets[l][best].pop()); left[l] -= 1
        k += 1
    return [np.array(o) for o in out]


if __name__ == "__main__":
    main2() if os.environ.get("COOP") else main()
