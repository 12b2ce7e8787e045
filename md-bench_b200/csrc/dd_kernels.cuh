// dd_kernels.cuh -- spatial decomposition: each GPU owns a brick of the box plus a ghost shell.
// Inside a brick, coordinates are BRICK-LOCAL ([0, ext) per axis), so a brick looks exactly like the
// reference's periodic box of size ext (pbc.c:98-227 conditions, updateAtomsPbc's single +-prd wrap,
// setupNeighbor's bin grid) -- the only difference is where the periodic images come from: the
// neighbor brick in that direction instead of the box itself.  With one brick per axis the neighbor
// is the brick itself and the scheme reduces to the reference's setupPbc/updatePbc.
// Image index b in [0,26) uses the reference's ADDGHOST ladder (c_img in vl_kernels.cuh):
// s_b = c_img[b] is the image shift; the atom is SENT in direction -s_b and appears at x + s_b*ext.
#pragma once
#include "mdb_util.cuh"

namespace mdb {

struct DirTable {
    int off[27]; // off[b] .. off[b+1]: entries of direction b
};

// createAtom (atom.c:67-187) restricted to one brick: one thread per FCC site of the brick, global
// emission index (closed form, see k_create_atoms) kept as the atom's tag, position brick-local.
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
    const int ox = i >> 3, sx = i & 7, oy = j >> 3, sy = j & 7, oz = k >> 3, sz = k & 7;
    const int bx = min(8, 2 * gnx - 8 * ox), by = min(8, 2 * gny - 8 * oy), bz = min(8, 2 * gnz - 8 * oz);
    const long long sites = 8LL * oz * (2LL * gnx) * (2LL * gny) + (2LL * gnx) * (8LL * oy) * bz + (8LL * ox) * by * bz;
    const long long a = sites / 2 + (long long)sz * (bx * by / 2) + (long long)sy * (bx / 2) +
                        (((sy + sz) & 1) ? sx / 2 : (sx + 1) / 2);
    int n = k * (2 * gny) * (2 * gnx) + j * (2 * gnx) + i + 1;
    double v[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        for (int m = 0; m < 5; m++) park_miller(n);
        v[c] = park_miller(n);
    }
    x[t]    = (real)(0.5 * (double)alat * il);
    y[t]    = (real)(0.5 * (double)alat * jl);
    z[t]    = (real)(0.5 * (double)alat * kl);
    vx[t]   = (real)v[0];
    vy[t]   = (real)v[1];
    vz[t]   = (real)v[2];
    type[t] = 0;
    tag[t]  = (int)a;
}

// ---- migration: atoms that left the brick (the decomposed updateAtomsPbc, pbc.c:59-84) -------------
// dest[i] = ladder index of the image shift that brings the atom into the receiving brick's local
// frame (x + s*ext), or -1 if it stays.  Axes with a single brick wrap in place like the reference.
template <class real>
__global__ void k_dd_dest(int nlocal, real ex, real ey, real ez, int px, int py, int pz, const signed char* __restrict__ s2b,
    real* __restrict__ x, real* __restrict__ y, real* __restrict__ z, int* __restrict__ dest, int* __restrict__ leave)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nlocal) return;
    int s[3]   = { 0, 0, 0 };
    real* c[3] = { x, y, z };
    const real e[3] = { ex, ey, ez };
    const int p[3]  = { px, py, pz };
#pragma unroll
    for (int a = 0; a < 3; a++) {
        real v = c[a][i];
        if (p[a] == 1) {
            c[a][i] = wrap1(v, e[a]);
        } else {
            if (v < (real)0.0) s[a] = +1;       // receiver sees x + ext
            else if (v >= e[a]) s[a] = -1;      // receiver sees x - ext
        }
    }
    const int b = s2b[(s[0] + 1) + 3 * (s[1] + 1) + 9 * (s[2] + 1)];
    dest[i]  = b;
    leave[i] = b >= 0;
}
// compact: stayers keep their relative order at the front, leavers are listed in ascending order
__global__ void k_dd_split(int nlocal, const int* __restrict__ leave, const int* __restrict__ leave_scan,
    int* __restrict__ stay_src, int* __restrict__ leavers)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nlocal) return;
    const int l = leave_scan[i];
    if (leave[i]) leavers[l] = i;
    else stay_src[i - l] = i;
}
// flags[b*nb + q] = 1 if border/leaver atom q goes to direction b
__global__ void k_dd_flags_mask(int nb, const int* __restrict__ list, const unsigned* __restrict__ mask, int* __restrict__ flags)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 26LL * nb) return;
    const int b = (int)(t / nb), q = (int)(t % nb);
    flags[t]    = (mask[list[q]] >> b) & 1u;
}
__global__ void k_dd_flags_dest(int nb, const int* __restrict__ list, const int* __restrict__ dest, int* __restrict__ flags)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 26LL * nb) return;
    const int b = (int)(t / nb), q = (int)(t % nb);
    flags[t]    = dest[list[q]] == b;
}
// sendlist[scan[t]] = list[q] for set flags: direction blocks come out contiguous and ascending
__global__ void k_dd_fill_sendlist(int nb, const int* __restrict__ list, const int* __restrict__ flags,
    const int* __restrict__ scan, int* __restrict__ sendlist)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 26LL * nb) return;
    if (flags[t]) sendlist[scan[t]] = list[(int)(t % nb)];
}
__global__ void k_dd_block_offsets(int nb, const int* __restrict__ scan, int* __restrict__ off)
{
    const int b = threadIdx.x;
    if (b < 26) off[b] = scan[(long long)b * nb];
}
__global__ void k_flag_nonzero(int n, const unsigned* __restrict__ mask, int* __restrict__ flag)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = mask[i] != 0;
}
__global__ void k_compact(int n, const int* __restrict__ flag, const int* __restrict__ scan, int* __restrict__ list)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) list[scan[i]] = i;
}

// pack W arrays of the listed atoms, direction segment b laid out as [a0 | a1 | ...], each cnt_b long;
// the first three arrays (positions) get the image shift x + s_b*ext as ONE fma like updatePbc (F11)
template <class real, int W>
__global__ void k_dd_pack(int total, DirTable T, const int* __restrict__ sendlist, real ex, real ey, real ez,
    const real* __restrict__ a0, const real* __restrict__ a1, const real* __restrict__ a2, const real* __restrict__ a3,
    const real* __restrict__ a4, const real* __restrict__ a5, real* __restrict__ out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    int b = 0;
#pragma unroll
    for (int k = 1; k < 26; k++) b += (t >= T.off[k]);
    const int cnt = T.off[b + 1] - T.off[b], q = t - T.off[b], i = sendlist[t];
    real* o = out + (size_t)W * T.off[b] + q;
    o[0]               = fma_rn((real)c_img[b][0], ex, a0[i]);
    o[cnt]             = fma_rn((real)c_img[b][1], ey, a1[i]);
    o[2 * (size_t)cnt] = fma_rn((real)c_img[b][2], ez, a2[i]);
    if (W > 3) {
        o[3 * (size_t)cnt] = a3[i];
        o[4 * (size_t)cnt] = a4[i];
        o[5 * (size_t)cnt] = a5[i];
    }
}
__global__ void k_dd_pack_int(int total, const int* __restrict__ sendlist, const int* __restrict__ a, int* __restrict__ out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < total) out[t] = a[sendlist[t]];
}
// unpack W arrays received for block b into dst arrays starting at `base`
template <class real, int W>
__global__ void k_dd_unpack(int total, DirTable T, int base, const real* __restrict__ in, real* __restrict__ a0,
    real* __restrict__ a1, real* __restrict__ a2, real* __restrict__ a3, real* __restrict__ a4, real* __restrict__ a5)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    int b = 0;
#pragma unroll
    for (int k = 1; k < 26; k++) b += (t >= T.off[k]);
    const int cnt = T.off[b + 1] - T.off[b], q = t - T.off[b];
    const real* s = in + (size_t)W * T.off[b] + q;
    a0[base + t] = s[0];
    a1[base + t] = s[cnt];
    a2[base + t] = s[2 * (size_t)cnt];
    if (W > 3) {
        a3[base + t] = s[3 * (size_t)cnt];
        a4[base + t] = s[4 * (size_t)cnt];
        a5[base + t] = s[5 * (size_t)cnt];
    }
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

} // namespace mdb
