// dd_topo.h -- host-only bookkeeping of the spatial decomposition (no CUDA): the brick grid, which
// brick receives / sends each of the 26 periodic-image directions, and the order in which a brick
// lays out what it sends and what it receives.  Shared by the device code (dd_group.cuh) and the
// C ABI's mdb_dd_plan(), which the CPU tests use to check that every rank derives the same plan.
//
// The reference has no decomposition ("replicas only"); what is generalised here is its
// setupPbc/updatePbc pair (verletlist/pbc.c:98-227, 42-55): image b of the ADDGHOST ladder with
// shift s_b is an atom that the brick at coords - s_b sees at x + s_b*ext.  With one brick per axis
// the receiver is the brick itself and the scheme is the reference's.
#pragma once
#include <algorithm>
#include <utility>
#include <vector>

namespace mdb {

static const signed char DD_IMG[26][3] = { { +1, 0, 0 }, { -1, 0, 0 }, { 0, +1, 0 }, { 0, -1, 0 }, { 0, 0, +1 },
    { 0, 0, -1 }, { +1, +1, +1 }, { +1, -1, +1 }, { +1, +1, -1 }, { +1, -1, -1 }, { -1, +1, +1 }, { -1, -1, +1 },
    { -1, +1, -1 }, { -1, -1, -1 }, { +1, 0, +1 }, { +1, 0, -1 }, { -1, 0, +1 }, { -1, 0, -1 }, { 0, +1, +1 },
    { 0, +1, -1 }, { 0, -1, +1 }, { 0, -1, -1 }, { +1, +1, 0 }, { -1, +1, 0 }, { +1, -1, 0 }, { -1, -1, 0 } };

struct Topo {
    int g[3]        = { 1, 1, 1 };
    int periodic[3] = { 1, 1, 1 };
    int nbricks = 1, nprocs = 1;

    void coords(int b, int c[3]) const
    {
        c[0] = b % g[0];
        c[1] = (b / g[0]) % g[1];
        c[2] = b / (g[0] * g[1]);
    }
    int id(const int c[3]) const { return c[0] + g[0] * (c[1] + g[1] * c[2]); }
    int owner(int b) const { return b / (nbricks / nprocs); } // consecutive bricks per process
    // brick at coords(b) + sign*s_d, or -1 across a non-periodic face
    int displaced(int b, int d, int sign) const
    {
        int c[3];
        coords(b, c);
        for (int a = 0; a < 3; a++) {
            const int s = sign * DD_IMG[d][a];
            if (!s) continue;
            int v = c[a] + s;
            if (v < 0 || v >= g[a]) {
                if (!periodic[a]) return -1;
                v = (v + g[a]) % g[a];
            }
            c[a] = v;
        }
        return id(c);
    }
    int receiver(int b, int d) const { return displaced(b, d, -1); } // who sees b's image d
    int sender(int b, int d) const { return displaced(b, d, +1); }   // whose atoms arrive at b as image d

    // slots: valid directions ordered by (peer brick, direction); returns their number
    int slots(int b, bool send, int dir[26], int peer[26]) const
    {
        std::pair<int, int> v[26];
        int n = 0;
        for (int d = 0; d < 26; d++) {
            const int p = send ? receiver(b, d) : sender(b, d);
            if (p >= 0) v[n++] = std::make_pair(p, d);
        }
        std::sort(v, v + n);
        for (int k = 0; k < n; k++) {
            peer[k] = v[k].first;
            dir[k]  = v[k].second;
        }
        return n;
    }
    unsigned valid_mask(int b) const
    {
        unsigned m = 0;
        for (int d = 0; d < 26; d++)
            if (receiver(b, d) >= 0) m |= 1u << d;
        return m;
    }
};

// One contiguous transfer of an exchange: `len` entries from entry `src_start` of brick `src`'s send list
// to entry `dst_start` of what brick `dst` receives.  kind: 0 = both bricks in this process (device copy),
// 1 = this process sends, 2 = this process receives.
struct Xfer {
    int kind, src, dst, src_start, dst_start, len, peer_proc;
};
// The schedule of one exchange as process `proc` executes it, from cnt[brick*26 + direction] = number of
// entries brick sends in that direction (known to every process after the allgather).  Receiving bricks
// ascending, then sending bricks ascending: both ends of a pair of processes enumerate their common
// transfers in the same order, which is what NCCL's in-order matching of send/recv needs.
inline void dd_schedule(const Topo& t, int proc, const int* cnt, std::vector<Xfer>& out)
{
    out.clear();
    // first entry of the segment S sends to R: S's send slots are ordered by (receiver, direction)
    auto send_start = [&](int S, int R) {
        int dir[26], peer[26];
        const int ns = t.slots(S, true, dir, peer);
        int off      = 0;
        for (int k = 0; k < ns && peer[k] < R; k++) off += cnt[S * 26 + dir[k]];
        return off;
    };
    for (int R = 0; R < t.nbricks; R++) {
        int dir[26], peer[26];
        const int ns = t.slots(R, false, dir, peer);
        int total    = 0;
        for (int k = 0; k < ns;) {
            int e = k, len = 0;
            while (e < ns && peer[e] == peer[k]) len += cnt[peer[e] * 26 + dir[e]], e++;
            const int S   = peer[k];
            const bool ms = t.owner(S) == proc, mr = t.owner(R) == proc;
            if (len && (ms || mr)) {
                Xfer x;
                x.kind      = ms && mr ? 0 : (ms ? 1 : 2);
                x.src       = S;
                x.dst       = R;
                x.src_start = send_start(S, R);
                x.dst_start = total;
                x.len       = len;
                x.peer_proc = ms && mr ? proc : (ms ? t.owner(R) : t.owner(S));
                out.push_back(x);
            }
            total += len;
            k = e;
        }
    }
}

} // namespace mdb
