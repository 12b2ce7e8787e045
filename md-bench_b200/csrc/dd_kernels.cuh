// dd_kernels.cuh -- spatial decomposition: each GPU owns a brick of the box plus a ghost shell.
// Inside a brick, coordinates are BRICK-LOCAL ([0, ext) per axis), so a brick looks exactly like the
// reference's periodic box of size ext (pbc.c:98-227 conditions, updateAtomsPbc's single +-prd wrap,
// setupNeighbor's bin grid) -- the only difference is where the periodic images come from: the
// neighbor brick in that direction instead of the box itself.  With one brick per axis the neighbor
// is the brick itself and the scheme reduces to the reference's setupPbc/updatePbc.
// Image index b in [0,26) uses the reference's ADDGHOST ladder (c_img in vl_kernels.cuh):
// s_b = c_img[b] is the image shift; the atom is SENT to the brick at coords - s_b and appears there
// at x + s_b*ext.  See dd_topo.h for the slot / peer-segment order both sides agree on.
#pragma once
#include "mdb_util.cuh"

namespace mdb {

// What a brick sends, by send slot (valid directions ordered by (receiving brick, direction)).
// Entries [off[s], off[s+1]) of the send list belong to slot s; consecutive slots with the same
// receiver form a peer segment [pstart, pstart+plen), stored in the send buffer as W arrays of plen
// (SoA per peer), so the receiver takes each array with one contiguous transfer straight into place.
struct SendTable {
    int nslots;
    int dir[26];
    int off[27];
    int pstart[26];
    int plen[26];
};

// createAtom (atom.c:67-187) restricted to one brick: one thread per FCC site of the brick, global
// emission index (closed form, see k_create_atoms) kept as the atom's tag, position brick-local.
// Inside the brick the atoms are stored in the same 8x8x8 sub-box order the reference's generator uses
// for the whole box: 32 consecutive atoms = one lattice plane of a sub-box, a compact patch whose
// neighbor gathers share cache lines (x-fastest rows were 1.4x slower in the force kernel, profiles/r1_ab2.txt).
__device__ __forceinline__ long long fcc_emission_index(int i, int j, int k, int nx, int ny, int nz)
{
    const int ox = i >> 3, sx = i & 7, oy = j >> 3, sy = j & 7, oz = k >> 3, sz = k & 7;
    const int bx = min(8, 2 * nx - 8 * ox), by = min(8, 2 * ny - 8 * oy), bz = min(8, 2 * nz - 8 * oz);
    const long long sites = 8LL * oz * (2LL * nx) * (2LL * ny) + (2LL * nx) * (8LL * oy) * bz + (8LL * ox) * by * bz;
    return sites / 2 + (long long)sz * (bx * by / 2) + (long long)sy * (bx / 2) + (((sy + sz) & 1) ? sx / 2 : (sx + 1) / 2);
}
template <class real>
__global__ void k_dd_create_atoms(int gnx, int gny, int gnz, int lnx, int lny, int lnz, int cx, int cy, int cz,
    real alat, real* __restrict__ x, real* __restrict__ y, real* __restrict__ z, real* __restrict__ vx,
    real* __restrict__ vy, real* __restrict__ vz, int* __restrict__ type, int* __restrict__ tag)
{
    const long long t     = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = 4LL * lnx * lny * lnz;
    if (t >= total) return;
    const int ihl = (int)(t % lnx);
    const int jl  = (int)((t / lnx) % (2 * lny));
    const int kl  = (int)(t / ((long long)lnx * 2 * lny));
    const int j = jl + 2 * lny * cy, k = kl + 2 * lnz * cz;
    const int il = 2 * ihl + ((j + k) & 1);
    const int i  = il + 2 * lnx * cx;
    const long long a = fcc_emission_index(i, j, k, gnx, gny, gnz);     // global tag
    const long long q = fcc_emission_index(il, jl, kl, lnx, lny, lnz); // slot inside the brick
    int n = k * (2 * gny) * (2 * gnx) + j * (2 * gnx) + i + 1;
    double v[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        for (int m = 0; m < 5; m++) park_miller(n);
        v[c] = park_miller(n);
    }
    x[q]    = (real)(0.5 * (double)alat * il);
    y[q]    = (real)(0.5 * (double)alat * jl);
    z[q]    = (real)(0.5 * (double)alat * kl);
    vx[q]   = (real)v[0];
    vy[q]   = (real)v[1];
    vz[q]   = (real)v[2];
    type[q] = 0;
    tag[q]  = (int)a;
}

// ---- migration: atoms that left the brick (the decomposed updateAtomsPbc, pbc.c:59-84) -------------
// dest[i] = ladder index of the image shift that brings the atom into the receiving brick's local
// frame (x + s*ext), or -1 if it stays.  Axes with a single brick wrap in place like the reference.
// s2b[27]: shift triple -> ladder index (-1 for no shift or a direction without a neighbor brick).
struct ShiftMap {
    signed char b[27];
};
template <class real>
__global__ void k_dd_dest(int nlocal, real ex, real ey, real ez, int px, int py, int pz, int perx, int pery, int perz,
    ShiftMap s2b, real* __restrict__ x, real* __restrict__ y, real* __restrict__ z, int* __restrict__ dest,
    int* __restrict__ leave)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nlocal) return;
    int s[3]        = { 0, 0, 0 };
    real* c[3]      = { x, y, z };
    const real e[3] = { ex, ey, ez };
    const int p[3]  = { px, py, pz };
    const int per[3] = { perx, pery, perz };
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const real v = c[a][i];
        if (p[a] == 1) {
            if (per[a]) c[a][i] = wrap1(v, e[a]);
        } else {
            if (v < (real)0.0) s[a] = +1;  // receiver sees x + ext
            else if (v >= e[a]) s[a] = -1; // receiver sees x - ext
        }
    }
    const int b = s2b.b[(s[0] + 1) + 3 * (s[1] + 1) + 9 * (s[2] + 1)];
    dest[i]     = b;
    leave[i]    = b >= 0;
}
// compact: stayers keep their relative order at the front, leavers are listed in ascending order
static __global__ void k_dd_split(int nlocal, const int* __restrict__ leave, const int* __restrict__ leave_scan,
    int* __restrict__ stay_src, int* __restrict__ leavers)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nlocal) return;
    const int l = leave_scan[i];
    if (leave[i]) leavers[l] = i;
    else stay_src[i - l] = i;
}
// flags[s*nb + q] = 1 if border atom list[q] has image dir[s]
static __global__ void k_dd_flags_mask(int nb, SendTable T, const int* __restrict__ list, const unsigned* __restrict__ mask,
    int* __restrict__ flags)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)T.nslots * nb) return;
    const int s = (int)(t / nb), q = (int)(t % nb);
    flags[t]    = (mask[list[q]] >> T.dir[s]) & 1u;
}
static __global__ void k_dd_flags_dest(int nb, SendTable T, const int* __restrict__ list, const int* __restrict__ dest,
    int* __restrict__ flags)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)T.nslots * nb) return;
    const int s = (int)(t / nb), q = (int)(t % nb);
    flags[t]    = dest[list[q]] == T.dir[s];
}
// sendlist[scan[t]] = list[q] for set flags: slot blocks come out contiguous, ascending inside
static __global__ void k_dd_fill_sendlist(int nb, int nslots, const int* __restrict__ list, const int* __restrict__ flags,
    const int* __restrict__ scan, int* __restrict__ sendlist)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)nslots * nb) return;
    if (flags[t]) sendlist[scan[t]] = list[(int)(t % nb)];
}
// off[s] = first entry of slot s (off[nslots] = total is written by the scan itself)
static __global__ void k_dd_block_offsets(int nb, int nslots, const int* __restrict__ scan, int* __restrict__ off)
{
    const int s = threadIdx.x;
    if (s < nslots) off[s] = scan[(long long)s * nb];
}
static __global__ void k_dd_flag_border(int n, unsigned valid, unsigned* __restrict__ mask, int* __restrict__ flag)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned m = mask[i] & valid;
    mask[i]          = m;
    flag[i]          = m != 0;
}
static __global__ void k_compact(int n, const int* __restrict__ flag, const int* __restrict__ scan, int* __restrict__ list)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) list[scan[i]] = i;
}

__device__ __forceinline__ int dd_slot_of(const SendTable& T, int t)
{
    int s = 0;
#pragma unroll
    for (int k = 1; k < 26; k++) s += (k < T.nslots && t >= T.off[k]);
    return s;
}
// pack W arrays of the listed atoms into per-peer SoA segments; with SHIFT the first three arrays
// (positions) get the image shift x + s_b*ext as ONE fma like updatePbc
template <class real, int W, bool SHIFT>
__global__ void k_dd_pack(int total, SendTable T, const int* __restrict__ sendlist, real ex, real ey, real ez,
    const real* __restrict__ a0, const real* __restrict__ a1, const real* __restrict__ a2, const real* __restrict__ a3,
    const real* __restrict__ a4, const real* __restrict__ a5, real* __restrict__ out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int s = dd_slot_of(T, t), b = T.dir[s];
    const size_t pl = T.plen[s];
    const int i     = sendlist[t];
    real* o         = out + (size_t)W * T.pstart[s] + (t - T.pstart[s]);
    if (SHIFT) {
        o[0] = fma_rn((real)c_img[b][0], ex, a0[i]);
        if (W > 1) o[pl] = fma_rn((real)c_img[b][1], ey, a1[i]);
        if (W > 2) o[2 * pl] = fma_rn((real)c_img[b][2], ez, a2[i]);
    } else {
        o[0] = a0[i];
        if (W > 1) o[pl] = a1[i];
        if (W > 2) o[2 * pl] = a2[i];
    }
    if (W > 3) {
        o[3 * pl] = a3[i];
        o[4 * pl] = a4[i];
        o[5 * pl] = a5[i];
    }
}
// ---- halo exchange by PEER STORES (one process per GPU, NVLink / NVSwitch peer memory) -------------------------------
// updatePbc across bricks without NCCL on the per-step path: the kernel that computes the image positions stores them
// straight into the ghost range of the RECEIVER's x, y, z arrays (mapped through CUDA IPC), segment by segment of the
// transfer schedule; a brick that is its own neighbor along an axis is just another segment whose destination is local.
// Ordering between the GPUs is a two-phase handshake on flag words in peer memory (k_dd_signal / k_dd_wait):
//   ready:  every process tells its senders that it has finished reading the previous ghost positions, and waits for the
//           same word from its receivers before it overwrites their ghosts;
//   done:   after the push kernel (kernel boundary + system fence) every process tells its receivers that this epoch's
//           positions have landed, and waits for the same word from its senders before anything reads its own ghosts.
struct PushSeg {
    int src_start, len; // entries [src_start, src_start + len) of the brick's send list
    void *dx, *dy, *dz; // destination of entry src_start in the receiver's arrays
};
struct PushTable {
    int nseg;
    PushSeg seg[32];
};
template <class real>
__global__ void k_dd_push(int total, SendTable T, PushTable P, const int* __restrict__ sendlist, real ex, real ey, real ez,
    const real* __restrict__ x, const real* __restrict__ y, const real* __restrict__ z)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int b = T.dir[dd_slot_of(T, t)];
    int g       = 0;
#pragma unroll 4
    for (int k = 1; k < 32; k++) g += (k < P.nseg && t >= P.seg[k].src_start);
    const PushSeg sg = P.seg[g];
    const int o      = t - sg.src_start;
    const int i      = sendlist[t];
    // the image shift as ONE fma like updatePbc
    ((real*)sg.dx)[o] = fma_rn((real)c_img[b][0], ex, x[i]);
    ((real*)sg.dy)[o] = fma_rn((real)c_img[b][1], ey, y[i]);
    ((real*)sg.dz)[o] = fma_rn((real)c_img[b][2], ez, z[i]);
}
struct PeerFlags {
    int* p[32]; // flag buffer of every process (own entry = own buffer)
};
// word `slot * 32 + me` of every process in `mask` := value (release at system scope: everything this GPU wrote before
// is visible to the reader that acquires the word)
static __global__ void k_dd_signal(PeerFlags F, unsigned mask, int slot, int me, int value)
{
    const int p = threadIdx.x;
    if (p >= 32 || !((mask >> p) & 1u)) return;
    __threadfence_system();
    int* w = F.p[p] + slot * 32 + me;
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(w), "r"(value) : "memory");
}
// spin until word `slot * 32 + p` of the own buffer reaches `value` for every p in `mask`; gives up after `timeout_ns`
// and raises the error word (own[96]) instead of hanging the GPU (the host checks it at every rebuild and after a run)
static __global__ void k_dd_wait(int* own, unsigned mask, int slot, int value, unsigned long long timeout_ns)
{
    const int p = threadIdx.x;
    if (p >= 32 || !((mask >> p) & 1u)) return;
    const int* w = own + slot * 32 + p;
    if (*((volatile int*)own + 96)) return; // an earlier wait already gave up: do not stall every following step as well
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        int v;
        asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(w) : "memory");
        if (v >= value) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > timeout_ns) {
            atomicExch(own + 96, 1);
            break;
        }
        __nanosleep(200);
    }
}

// two int arrays (type, tag) in the same per-peer segment layout
static __global__ void k_dd_pack_int2(int total, SendTable T, const int* __restrict__ sendlist, const int* __restrict__ a0,
    const int* __restrict__ a1, int* __restrict__ out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int s = dd_slot_of(T, t);
    const int i = sendlist[t];
    int* o      = out + (size_t)2 * T.pstart[s] + (t - T.pstart[s]);
    o[0]            = a0[i];
    o[T.plen[s]]    = a1[i];
}
// gather the stayers to the front (like k_permute_atoms but without forces)
template <class real>
__global__ void k_dd_gather_stay(int n, const int* __restrict__ src, const real* __restrict__ x, const real* __restrict__ y,
    const real* __restrict__ z, const real* __restrict__ vx, const real* __restrict__ vy, const real* __restrict__ vz,
    const int* __restrict__ type, const int* __restrict__ tag, real* __restrict__ nx, real* __restrict__ ny,
    real* __restrict__ nz, real* __restrict__ nvx, real* __restrict__ nvy, real* __restrict__ nvz, int* __restrict__ ntype,
    int* __restrict__ ntag)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const int o = src[q];
    nx[q] = x[o]; ny[q] = y[o]; nz[q] = z[o];
    nvx[q] = vx[o]; nvy[q] = vy[o]; nvz[q] = vz[o];
    ntype[q] = type[o];
    ntag[q]  = tag[o];
}
// neighbor rows as global tags (parity read-back): out[i*stride + k] = tag[neighbor k of i]
static __global__ void k_dd_rows_as_tags(int nlocal, int stride, NbLayout L, const int* __restrict__ numneigh,
    const int* __restrict__ neighbors, const int* __restrict__ tag, int* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nlocal) return;
    const int n = min(numneigh[i], stride);
    for (int k = 0; k < n; k++) out[(size_t)i * stride + k] = tag[neighbors[L.base(i) + (size_t)k * L.sk]];
}
// brick-local -> global coordinates for read-back
template <class real>
__global__ void k_dd_to_global(int n, real ox, real oy, real oz, const real* __restrict__ x, const real* __restrict__ y,
    const real* __restrict__ z, real* __restrict__ gx, real* __restrict__ gy, real* __restrict__ gz)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    gx[i] = x[i] + ox; gy[i] = y[i] + oy; gz[i] = z[i] + oz;
}

// ---- scatter by brick: atoms handed over in the global frame (input readers, bench restarts) ------
// flag[i] = atom i (wrapped into the box once, like updateAtomsPbc) lies in brick (cx,cy,cz)
template <class real>
__global__ void k_dd_select(int n, real ex, real ey, real ez, int gx, int gy, int gz, int cx, int cy, int cz,
    const real* __restrict__ x, const real* __restrict__ y, const real* __restrict__ z, int* __restrict__ flag)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const real X = wrap1(x[i], ex * gx), Y = wrap1(y[i], ey * gy), Z = wrap1(z[i], ez * gz);
    const int a = min(gx - 1, max(0, (int)floor(X / ex))), b = min(gy - 1, max(0, (int)floor(Y / ey))),
              c = min(gz - 1, max(0, (int)floor(Z / ez)));
    flag[i] = a == cx && b == cy && c == cz;
}
template <class real>
__global__ void k_dd_take(int n, const int* __restrict__ flag, const int* __restrict__ scan, real ex, real ey, real ez,
    int gx, int gy, int gz, int cx, int cy, int cz, const real* __restrict__ x, const real* __restrict__ y,
    const real* __restrict__ z, const real* __restrict__ vx, const real* __restrict__ vy, const real* __restrict__ vz,
    const int* __restrict__ tag, real* __restrict__ ox, real* __restrict__ oy, real* __restrict__ oz, real* __restrict__ ovx,
    real* __restrict__ ovy, real* __restrict__ ovz, int* __restrict__ otype, int* __restrict__ otag)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !flag[i]) return;
    const int q = scan[i];
    ox[q]    = wrap1(x[i], ex * gx) - cx * ex;
    oy[q]    = wrap1(y[i], ey * gy) - cy * ey;
    oz[q]    = wrap1(z[i], ez * gz) - cz * ez;
    ovx[q]   = vx ? vx[i] : (real)0;
    ovy[q]   = vy ? vy[i] : (real)0;
    ovz[q]   = vz ? vz[i] : (real)0;
    otype[q] = 0;
    otag[q]  = tag[i];
}

} // namespace mdb
