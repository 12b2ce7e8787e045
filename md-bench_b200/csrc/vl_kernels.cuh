// vl_kernels.cuh -- sm_100a kernels of the verletlist hot path.
// Data layout in HBM: SoA (x[],y[],z[], vx.., fx..), locals first then ghosts; the neighbor list
// is addressed as neighbors[i*si + k*sk]: row-major rows (si = row length, sk = 1; what the
// lanes-per-atom force kernel wants) or transposed (si = 1, sk = row stride; coalesced for one
// thread per atom).  The reference's own numbering/rows are rebuilt only for parity read-back.
#pragma once
#include "mdb_util.cuh"

namespace mdb {

// ---------------------------------------------------------------------------------------------
// geometry of setupNeighbor (reference verletlist/neighbor.c:24-38)
template <class real> struct BinGeom {
    real xprd, yprd, zprd;
    real bininvx, bininvy, bininvz;
    int nbinx, nbiny, nbinz;
    int mbinxlo, mbinylo, mbinzlo;
    int mbinx, mbiny, mbinz, mbins;
};

// coord2bin, verletlist/neighbor.c:298-327, including the stray "+ 1".  Each product is a single
// rounded multiply (no contraction), the cast truncates toward zero like C's (int).
template <class real> __device__ __forceinline__ int axis2bin(real v, real prd, real bininv, int nbin, int mlo)
{
    if (v >= prd) return (int)mul_rn(sub_rn(v, prd), bininv) + nbin - mlo;
    if (v >= (real)0.0) return (int)mul_rn(v, bininv) - mlo;
    return (int)mul_rn(v, bininv) - mlo - 1;
}
template <class real> __device__ __forceinline__ int coord2bin(const BinGeom<real>& g, real x, real y, real z)
{
    const int ix = axis2bin(x, g.xprd, g.bininvx, g.nbinx, g.mbinxlo);
    const int iy = axis2bin(y, g.yprd, g.bininvy, g.nbiny, g.mbinylo);
    const int iz = axis2bin(z, g.zprd, g.bininvz, g.nbinz, g.mbinzlo);
    int b        = iz * g.mbiny * g.mbinx + iy * g.mbinx + ix + 1;
    // a blown-up simulation must not turn into an out-of-bounds write
    return b < 0 ? 0 : (b > g.mbins ? g.mbins : b);
}

// ---------------------------------------------------------------------------------------------
// createAtom, verletlist/atom.c:67-187.  One thread per FCC site (i+j+k even).  The reference
// emits atoms while walking 8x8x8 sub-boxes of the half-lattice; because 2*n{x,y,z} and 8 are even
// every (sub-)box holds exactly half of its sites, which gives the emission index in closed form.
__device__ __forceinline__ double park_miller(int& seed) // common/util.c:24-33
{
    const int IA = 16807, IM = 2147483647, IQ = 127773, IR = 2836;
    const int k  = seed / IQ;
    seed         = IA * (seed - k * IQ) - IR * k;
    if (seed < 0) seed += IM;
    return (1.0 / IM) * seed;
}

template <class real>
__global__ void k_create_atoms(int nx, int ny, int nz, real alat, real* __restrict__ x,
    real* __restrict__ y, real* __restrict__ z, real* __restrict__ vx, real* __restrict__ vy,
    real* __restrict__ vz, int* __restrict__ type)
{
    const long long t     = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = 4LL * nx * ny * nz;
    if (t >= total) return;
    const int ih = (int)(t % nx);
    const int j  = (int)((t / nx) % (2 * ny));
    const int k  = (int)(t / ((long long)nx * 2 * ny));
    const int i  = 2 * ih + ((j + k) & 1);
    const int ox = i >> 3, sx = i & 7, oy = j >> 3, sy = j & 7, oz = k >> 3, sz = k & 7;
    const int bx = min(8, 2 * nx - 8 * ox), by = min(8, 2 * ny - 8 * oy), bz = min(8, 2 * nz - 8 * oz);
    long long sites = 8LL * oz * (2LL * nx) * (2LL * ny) // complete layers of boxes
                      + (2LL * nx) * (8LL * oy) * bz      // complete rows of boxes in this layer
                      + (8LL * ox) * by * bz;             // boxes before this one in the row
    long long a = sites / 2 + (long long)sz * (bx * by / 2) + (long long)sy * (bx / 2) +
                  (((sy + sz) & 1) ? sx / 2 : (sx + 1) / 2);
    int n = k * (2 * ny) * (2 * nx) + j * (2 * nx) + i + 1;
    double v[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        for (int m = 0; m < 5; m++) park_miller(n);
        v[c] = park_miller(n);
    }
    // 0.5 * alat * i is evaluated in double and narrowed on assignment (atom.c:126-128)
    x[a]    = (real)(0.5 * (double)alat * i);
    y[a]    = (real)(0.5 * (double)alat * j);
    z[a]    = (real)(0.5 * (double)alat * k);
    vx[a]   = (real)v[0];
    vy[a]   = (real)v[1];
    vz[a]   = (real)v[2];
    type[a] = 0; // rand() % ntypes with ntypes == 1 (atom.c:158)
}

// ---------------------------------------------------------------------------------------------
// deterministic two-stage sum reductions (thermo, adjustThermo).  Accumulation in double.
constexpr int RED_THREADS = 256;
constexpr int RED_BLOCKS  = 1184; // 148 SMs x 8 resident blocks

__device__ __forceinline__ double block_sum(double v)
{
    __shared__ double sh[RED_THREADS / 32];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        v = threadIdx.x < RED_THREADS / 32 ? sh[threadIdx.x] : 0.0;
#pragma unroll
        for (int d = 4; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    }
    __syncthreads();
    return v; // valid in thread 0
}

// partial[b*4 + {0,1,2,3}] = sum vx, sum vy, sum vz, sum (vx^2+vy^2+vz^2)*mass over the block's atoms
template <class real>
__global__ void __launch_bounds__(RED_THREADS) k_vel_partial(int n, const real* __restrict__ vx,
    const real* __restrict__ vy, const real* __restrict__ vz, real mass, double* __restrict__ partial)
{
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const real a = vx[i], b = vy[i], c = vz[i];
        s0 += a;
        s1 += b;
        s2 += c;
        s3 += (double)((a * a + b * b + c * c) * mass);
    }
    s0 = block_sum(s0);
    s1 = block_sum(s1);
    s2 = block_sum(s2);
    s3 = block_sum(s3);
    if (threadIdx.x == 0) {
        partial[blockIdx.x * 4 + 0] = s0;
        partial[blockIdx.x * 4 + 1] = s1;
        partial[blockIdx.x * 4 + 2] = s2;
        partial[blockIdx.x * 4 + 3] = s3;
    }
}
// out[0..3] = sums over blocks
static __global__ void __launch_bounds__(RED_THREADS) k_vel_final(int nblocks, const double* __restrict__ partial, double* __restrict__ out)
{
    for (int c = 0; c < 4; c++) {
        double s = 0;
        for (int b = threadIdx.x; b < nblocks; b += blockDim.x) s += partial[b * 4 + c];
        s = block_sum(s);
        if (threadIdx.x == 0) out[c] = s;
    }
}

// adjustThermo steps, common/thermo.c:98-102 and 117-121
template <class real>
__global__ void k_vel_shift(int n, real* vx, real* vy, real* vz, real sx, real sy, real sz)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { vx[i] -= sx; vy[i] -= sy; vz[i] -= sz; }
}
template <class real> __global__ void k_vel_scale(int n, real* vx, real* vy, real* vz, real f)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { vx[i] *= f; vy[i] *= f; vz[i] *= f; }
}

template <class real> struct Vec2Of;
template <> struct Vec2Of<double> { typedef double2 type; };
template <> struct Vec2Of<float> { typedef float2 type; };

// ---------------------------------------------------------------------------------------------
// velocity-Verlet halves, verletlist/integrate.c:21-31 / 33-40
template <class real>
__global__ void k_initial_integrate(int n, real dtforce, real dt, real* __restrict__ x,
    real* __restrict__ y, real* __restrict__ z, real* __restrict__ vx, real* __restrict__ vy,
    real* __restrict__ vz, const real* __restrict__ fx, const real* __restrict__ fy,
    const real* __restrict__ fz)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const real a = vx[i] + dtforce * fx[i], b = vy[i] + dtforce * fy[i], c = vz[i] + dtforce * fz[i];
    vx[i] = a; vy[i] = b; vz[i] = c;
    x[i]  = x[i] + dt * a;
    y[i]  = y[i] + dt * b;
    z[i]  = z[i] + dt * c;
}
template <class real>
__global__ void k_final_integrate(int n, real dtforce, real* __restrict__ vx, real* __restrict__ vy,
    real* __restrict__ vz, const real* __restrict__ fx, const real* __restrict__ fy,
    const real* __restrict__ fz)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    vx[i] += dtforce * fx[i];
    vy[i] += dtforce * fy[i];
    vz[i] += dtforce * fz[i];
}

// finalIntegrate of step n fused with initialIntegrate of step n+1 (used by the device-resident loop
// when no thermo record falls between them): one pass over x, v, f instead of two.  The arithmetic
// is the same sequence of operations as the two kernels above, so the result is bit-identical.
template <class real>
__global__ void k_final_initial_integrate(int n, real dtforce, real dt, real* __restrict__ x,
    real* __restrict__ y, real* __restrict__ z, real* __restrict__ vx, real* __restrict__ vy,
    real* __restrict__ vz, const real* __restrict__ fx, const real* __restrict__ fy,
    const real* __restrict__ fz)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const real f0 = fx[i], f1 = fy[i], f2 = fz[i];
    real a = vx[i] + dtforce * f0, b = vy[i] + dtforce * f1, c = vz[i] + dtforce * f2; // final(n)
    a = a + dtforce * f0; b = b + dtforce * f1; c = c + dtforce * f2;                    // initial(n+1)
    vx[i] = a; vy[i] = b; vz[i] = c;
    x[i]  = x[i] + dt * a;
    y[i]  = y[i] + dt * b;
    z[i]  = z[i] + dt * c;
}

// ---------------------------------------------------------------------------------------------
// PBC.  Ghost image codes in the order of the reference's ADDGHOST ladder (verletlist/pbc.c:
// 107-224): 6 faces, 8 corners, 12 edges (x-z, y-z, x-y).  need[axis] = +1: requires coordinate
// < cutneigh (image shifted by +prd); -1: requires coordinate >= prd - cutneigh.
__constant__ signed char c_img[26][3] = {
    { +1, 0, 0 }, { -1, 0, 0 }, { 0, +1, 0 }, { 0, -1, 0 }, { 0, 0, +1 }, { 0, 0, -1 },
    { +1, +1, +1 }, { +1, -1, +1 }, { +1, +1, -1 }, { +1, -1, -1 },
    { -1, +1, +1 }, { -1, -1, +1 }, { -1, +1, -1 }, { -1, -1, -1 },
    { +1, 0, +1 }, { +1, 0, -1 }, { -1, 0, +1 }, { -1, 0, -1 },
    { 0, +1, +1 }, { 0, +1, -1 }, { 0, -1, +1 }, { 0, -1, -1 },
    { +1, +1, 0 }, { -1, +1, 0 }, { +1, -1, 0 }, { -1, -1, 0 } };

template <class real> struct PbcGeom {
    real xprd, yprd, zprd, cutneigh;
    real xhi_cut, yhi_cut, zhi_cut; // prd - cutneigh, rounded once like the reference's (xprd - cutneigh)
    int pbc_x, pbc_y, pbc_z;
};

template <class real> __device__ __forceinline__ unsigned ghost_mask(const PbcGeom<real>& g, real x, real y, real z)
{
    const bool lo[3] = { x < g.cutneigh, y < g.cutneigh, z < g.cutneigh };
    const bool hi[3] = { x >= g.xhi_cut, y >= g.yhi_cut, z >= g.zhi_cut };
    const bool en[3] = { g.pbc_x != 0, g.pbc_y != 0, g.pbc_z != 0 };
    unsigned m       = 0;
#pragma unroll
    for (int b = 0; b < 26; b++) {
        bool ok = true;
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const int d = c_img[b][a];
            if (d > 0) ok = ok && en[a] && lo[a];
            if (d < 0) ok = ok && en[a] && hi[a];
        }
        if (ok) m |= 1u << b;
    }
    return m;
}

// setupPbc pass 1 (pbc.c:98-227): per local atom, which of the 26 images exist
template <class real>
__global__ void k_ghost_count(int nlocal, PbcGeom<real> g, const real* __restrict__ x,
    const real* __restrict__ y, const real* __restrict__ z, unsigned* __restrict__ mask,
    int* __restrict__ count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nlocal) return;
    const unsigned m = ghost_mask(g, x[i], y[i], z[i]);
    mask[i]          = m;
    count[i]         = __popc(m);
}
// setupPbc pass 2: ghost index = exclusive scan of the counts in atom order + rank of the image in
// the ladder, i.e. exactly the reference's Nghost++ order.  code = (dx+1) | (dy+1)<<2 | (dz+1)<<4.
static __global__ void k_ghost_fill(int nlocal, const unsigned* __restrict__ mask,
    const int* __restrict__ offset, int* __restrict__ border_map, int* __restrict__ code,
    int* __restrict__ type)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nlocal) return;
    unsigned m = mask[i];
    int g      = offset[i];
    const int t = type[i];
    while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        border_map[g]    = i;
        code[g]          = (c_img[b][0] + 1) | ((c_img[b][1] + 1) << 2) | ((c_img[b][2] + 1) << 4);
        type[nlocal + g] = t;
        g++;
    }
}
// updatePbc, pbc.c:42-55: x[nlocal+g] = x[border_map[g]] + PBCx[g]*xprd as ONE fma (the
// reference's -Ofast build contracts it; ghost coordinates feed the list build).
template <class real>
__global__ void k_update_pbc(int nlocal, int nghost, real xprd, real yprd, real zprd,
    const int* __restrict__ border_map, const int* __restrict__ code, real* __restrict__ x,
    real* __restrict__ y, real* __restrict__ z, typename Vec2Of<real>::type* __restrict__ xy)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nghost) return;
    const int s = border_map[g], c = code[g];
    const real a = fma_rn((real)((c & 3) - 1), xprd, x[s]);
    const real b = fma_rn((real)(((c >> 2) & 3) - 1), yprd, y[s]);
    x[nlocal + g] = a;
    y[nlocal + g] = b;
    z[nlocal + g] = fma_rn((real)(((c >> 4) & 3) - 1), zprd, z[s]);
    if (xy) { // the packed (x, y) copy the xy-gather force kernel reads (k_force_lj_full_fi<.., XY>)
        typename Vec2Of<real>::type v;
        v.x = a; v.y = b;
        xy[nlocal + g] = v;
    }
}
template <class real>
__global__ void k_pack_xy(int n, const real* __restrict__ x, const real* __restrict__ y, typename Vec2Of<real>::type* __restrict__ xy)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    typename Vec2Of<real>::type v;
    v.x = x[i]; v.y = y[i];
    xy[i] = v;
}
// updateAtomsPbc, pbc.c:59-84
template <class real> __device__ __forceinline__ real wrap1(real v, real prd)
{
    if (v < (real)0.0) return add_rn(v, prd);
    if (v >= prd) return sub_rn(v, prd);
    return v;
}
template <class real>
__global__ void k_update_atoms_pbc(int nlocal, real xprd, real yprd, real zprd, real* __restrict__ x,
    real* __restrict__ y, real* __restrict__ z)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nlocal) return;
    x[i] = wrap1(x[i], xprd);
    y[i] = wrap1(y[i], yprd);
    z[i] = wrap1(z[i], zprd);
}

// ---------------------------------------------------------------------------------------------
// binatoms, verletlist/neighbor.c:329-358, as a counting sort: histogram -> scan -> fill -> sort
// each bin ascending (= the reference's insertion order, so rows come out in the same order).
template <class real>
__global__ void k_bin_count(int nall, BinGeom<real> g, const real* __restrict__ x,
    const real* __restrict__ y, const real* __restrict__ z, const int* __restrict__ rank,
    int* __restrict__ atom_bin, int* __restrict__ bincount)
{
    // rank == nullptr: key = the reference's bin index; else key = rank[bin] (sort order of the bins, used only by sort_atoms)
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nall) return;
    int b = coord2bin(g, x[i], y[i], z[i]);
    if (rank) b = rank[b];
    atom_bin[i] = b;
    atomicAdd(&bincount[b], 1);
}
static __global__ void k_bin_fill(int nall, const int* __restrict__ atom_bin, const int* __restrict__ binstart,
    int* __restrict__ cursor, int* __restrict__ binatoms)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nall) return;
    const int b                          = atom_bin[i];
    binatoms[binstart[b] + atomicAdd(&cursor[b], 1)] = i;
}
static __global__ void k_bin_sort(int nbins, const int* __restrict__ binstart, int* __restrict__ binatoms,
    const int* __restrict__ key, int* __restrict__ maxcount)
{
    // key == nullptr: ascending atom index; else ascending key[atom] (the atom's reference index)
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    int cnt     = 0;
    if (b < nbins) {
        const int s = binstart[b];
        cnt         = binstart[b + 1] - s;
        int* a      = binatoms + s;
        for (int i = 1; i < cnt; i++) { // insertion sort; bins hold ~8 atoms
            const int v  = a[i];
            const int kv = key ? key[v] : v;
            int j        = i - 1;
            while (j >= 0 && (key ? key[a[j]] : a[j]) > kv) { a[j + 1] = a[j]; j--; }
            a[j + 1] = v;
        }
    }
    cnt = __reduce_max_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt > 0) atomicMax(maxcount, cnt);
}

// rank of bin b in the blocked order of sort_atoms (Sim::build_bin_rank): bins are numbered 1 + (iz * my + iy) * mx + ix
// (coord2bin's stray "+ 1"; bin 0 keeps rank 0).  Blocks of B^3 bins, clipped at the upper faces of the grid.
static __global__ void k_block_rank(int nb, int mx, int my, int mz, int B, int* __restrict__ rank)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    if (b == 0) { rank[0] = 0; return; }
    const int l = b - 1, ix = l % mx, iy = (l / mx) % my, iz = l / (mx * my);
    const int bx = ix / B, by = iy / B, bz = iz / B;
    const int sx = min(B, mx - bx * B), sy = min(B, my - by * B), sz = min(B, mz - bz * B);
    const long long r = (long long)mx * my * (bz * B)         // layers of blocks below
        + (long long)mx * (by * B) * sz                       // rows of blocks in front, in this layer
        + (long long)(bx * B) * sy * sz                       // blocks to the left, in this row
        + ((long long)(iz - bz * B) * sy + (iy - by * B)) * sx + (ix - bx * B);
    rank[b] = (int)(1 + r);
}

// sortAtom (verletlist/neighbor.c:360-426, the reference's optional SORT_ATOMS step): permute the
// local atoms into bin order.  perm[q] = old index of the atom that moves to slot q.  Unlike the
// reference the permutation is tracked (orig[] = reference index of each slot), so everything the
// caller sees keeps the reference's atom numbering.
template <class real>
__global__ void k_permute_atoms(int n, const int* __restrict__ perm, const real* __restrict__ x,
    const real* __restrict__ y, const real* __restrict__ z, const real* __restrict__ vx,
    const real* __restrict__ vy, const real* __restrict__ vz, const real* __restrict__ fx,
    const real* __restrict__ fy, const real* __restrict__ fz, const int* __restrict__ type,
    const int* __restrict__ orig, real* __restrict__ nx, real* __restrict__ ny, real* __restrict__ nz,
    real* __restrict__ nvx, real* __restrict__ nvy, real* __restrict__ nvz, real* __restrict__ nfx,
    real* __restrict__ nfy, real* __restrict__ nfz, int* __restrict__ ntype, int* __restrict__ norig)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const int o = perm[q];
    nx[q] = x[o]; ny[q] = y[o]; nz[q] = z[o];
    nvx[q] = vx[o]; nvy[q] = vy[o]; nvz[q] = vz[o];
    nfx[q] = fx[o]; nfy[q] = fy[o]; nfz[q] = fz[o];
    ntype[q] = type[o];
    norig[q] = orig[o];
}
static __global__ void k_iota(int n, int* __restrict__ a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = i;
}
// out[orig[p]] = in[p] for locals: back to the reference's atom order
template <class real>
__global__ void k_scatter_orig(int n, const int* __restrict__ orig, const real* __restrict__ a,
    const real* __restrict__ b, const real* __restrict__ c, real* __restrict__ oa, real* __restrict__ ob,
    real* __restrict__ oc)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int o = orig[p];
    oa[o] = a[p]; ob[o] = b[p]; oc[o] = c[p];
}

static __global__ void k_scatter_orig_int(int n, const int* __restrict__ orig, const int* __restrict__ a, int* __restrict__ out)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) out[orig[p]] = a[p];
}

// buildNeighbor, verletlist/neighbor.c:186-264.  One thread per local atom walks the stencil of its bin as 21 runs of
// x-adjacent bins (contiguous in the CSR).  Membership is the reference's: rsq = fma(dx,dx,fma(dy,dy,dz*dz)) (the contraction
// GCC -Ofast makes) <= cutneighsq (neighbor.c:240).  For DP a float pre-test on float copies of the candidates
// sorts every candidate into certainly inside / certainly outside / uncertain (margin = Sim::list_margin) and only the
// uncertain band evaluates the exact FP64 expression; for SP the float expression is the reference's own.  Candidates are
// SoA float arrays in CSR order: one 128-bit load fetches a coordinate of FOUR consecutive candidates, the distance of two
// candidates is one FADD2/FMUL2/FFMA2 sequence (bit-identical to the scalar round-to-nearest operations), and the
// threshold tests are two more packed subtractions whose SIGN bits are shifted into the pass / maybe masks with one funnel
// shift each (rs < T  <=>  sign(rs - T)).  Each flush of up to 32 candidates is split into that branch-free phase and an
// append phase whose trip count is the number of hits.  Entries beyond maxneighs are counted but not stored and the host
// re-runs with a larger maxneighs (neighbor.c:247-262).  (Generations 1-4 of this kernel: profiles/r1_ab*.txt.)
template <class real>
__global__ void k_pack_binned_soa(int nall, int npad, const int* __restrict__ binatoms, const real* __restrict__ x,
    const real* __restrict__ y, const real* __restrict__ z, float* __restrict__ cx, float* __restrict__ cy, float* __restrict__ cz,
    int* __restrict__ cid)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= npad) return;
    if (m < nall) {
        const int j = binatoms[m];
        cx[m] = (float)x[j]; cy[m] = (float)y[j]; cz[m] = (float)z[j];
        cid[m] = j;
    } else { // alignment padding behind the last candidate: never in range
        cx[m] = cy[m] = cz[m] = 1.0e30f;
        cid[m] = -1;
    }
}
__device__ __forceinline__ void ld2x2(const float* p, f32x2& a, f32x2& b)
{
    asm("ld.global.nc.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
}
__device__ __forceinline__ int bfind(unsigned v) // index of the highest set bit (FLO.U32)
{
    int p;
    asm("bfind.u32 %0, %1;" : "=r"(p) : "r"(v));
    return p;
}
template <int BYTES, class T> __device__ __forceinline__ T* mad_wide_s(T* base, int idx) // base + idx * BYTES as one IMAD.WIDE
{
    T* r;
    asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(r) : "r"(idx), "n"(BYTES), "l"(base));
    return r;
}
template <int BYTES, class T> __device__ __forceinline__ T* mad_wide_u(T* base, unsigned idx)
{
    T* r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(idx), "n"(BYTES), "l"(base));
    return r;
}
__device__ __forceinline__ void st_global(int* p, int v) { asm volatile("st.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_nc(const int* p)
{
    int v;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
// ---- per-ATOM stencil ----------------------------------------------------------------------------------------------------
// The reference's stencil is per BIN (every bin any atom of the bin could reach, neighbor.c:160-183): 81 bins = 21 runs,
// ~600 candidates for ~75 hits.  An atom only needs the bins its own sphere of radius cutneigh touches (~37 of 81): a
// run whose (y, z) bin row lies farther than cutneigh from the atom is skipped, and the x range of a run shrinks to
// sqrt(cutneigh^2 - gap_yz^2) around the atom.  Both tests are conservative (float arithmetic with a margin of 1e-3 bin
// widths, far above the rounding of coord2bin's bin edges), so only candidates that would fail the distance test
// disappear and the rows come out identical, entry by entry.  A narrowed run (~20 candidates) fits one flush.
// One run of x-adjacent stencil bins (a (dy, dz) row of the stencil) as the build kernel reads it: two 128-bit loads.
// gap between an atom at (uy, uz) inside its bin and the row: max(sy * uy + ay, 0) with (ay, sy) = (dy * bs - margin, -1) above,
// (-(dy + 1) * bs - margin, +1) below, (-margin, 0) for the own row -- conservative by `margin`, like the x range.
struct alignas(16) RunRow {
    int base, i0, i1, pad; // bin offset of the row's x offset 0 (dz * mbiny * mbinx + dy * mbinx), first / last x offset of the run
    float ay, sy, az, sz;
};
__device__ __forceinline__ float sqrt_approx(float v) // MUFU.SQRT (2 ulp; the caller adds a margin of 1e-3 bin widths)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
struct RunGeom {
    float bsx, bsy, bsz;    // bin widths
    float binvx;            // 1 / bsx
    float cutsq_hi, margin; // cutneigh^2 * (1 + 1e-4), 1e-3 * min bin width
};
template <class real, bool HALF>
__global__ void __launch_bounds__(128) k_build_neighbor_v6(int nlocal, BinGeom<real> g, RunGeom rg, real cutneighsq, float lo,
    float hi, const real* __restrict__ x, const real* __restrict__ y, const real* __restrict__ z, const float* __restrict__ cx,
    const float* __restrict__ cy, const float* __restrict__ cz, const int* __restrict__ cid, const int* __restrict__ binstart,
    const RunRow* __restrict__ runs, int nruns, int maxneighs,
    NbLayout L, const int* __restrict__ orig, int* __restrict__ numneigh, int* __restrict__ neighbors, int* __restrict__ max_n)
{
    const int i     = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < nlocal;
    const int ii    = live ? i : 0;
    int n           = 0;
    const real xt = x[ii], yt = y[ii], zt = z[ii];
    const float xs = (float)xt, ys = (float)yt, zs = (float)zt;
    const f32x2 xs2 = pk2(xs, xs), ys2 = pk2(ys, ys), zs2 = pk2(zs, zs);
    const float tpass = sizeof(real) == 4 ? nextafterf((float)cutneighsq, INFINITY) : lo;
    const float tmay  = sizeof(real) == 4 ? tpass : nextafterf(hi, INFINITY);
    const f32x2 tp2 = pk2(tpass, tpass), tm2 = pk2(tmay, tmay);
    const int oi = HALF ? orig[ii] : 0;
    const int ix = axis2bin(xt, g.xprd, g.bininvx, g.nbinx, g.mbinxlo);
    const int iy = axis2bin(yt, g.yprd, g.bininvy, g.nbiny, g.mbinylo);
    const int iz = axis2bin(zt, g.zprd, g.bininvz, g.nbinz, g.mbinzlo);
    int ibin     = iz * g.mbiny * g.mbinx + iy * g.mbinx + ix + 1;
    ibin         = ibin < 0 ? 0 : (ibin > g.mbins ? g.mbins : ibin);
    // position relative to the lower edge of the own bin (absolute bin = index + mbinlo)
    const float ux = xs - (float)(ix + g.mbinxlo) * rg.bsx, uy = ys - (float)(iy + g.mbinylo) * rg.bsy,
                uz = zs - (float)(iz + g.mbinzlo) * rg.bsz;
    // the row as a 32-bit element offset from its first entry: one IMAD.WIDE per store instead of a 64-bit pointer bump
    int* const row     = neighbors + L.base(ii);
    int r    = 0;
    for (;;) {
        int s = 0, e = 0;
        // all lanes of the warp look at the SAME stencil row r: warp-uniform control flow, and the lanes -- atoms adjacent in
        // space -- find similar numbers of candidates and hits in it, which keeps the trip counts of the two divergent loops
        // below (distance tests, appends) close to their means.  A lane whose sphere misses the row idles for that round.
        // (Per-lane run sequences, each lane skipping ahead to the next row it needs: 4.42 vs 4.12 ms, profiles/r2_s3_call5.sh.)
        if (r >= nruns) break;
        {
            const int4 ri   = __ldg(reinterpret_cast<const int4*>(runs) + 2 * r);       // bin offset of x offset 0, first / last x offset
            const float4 rf = __ldg(reinterpret_cast<const float4*>(runs) + 2 * r + 1); // gap(y) = max(sy * uy + ay, 0), same for z
            r++;
            const float gym = fmaxf(fmaf(rf.y, uy, rf.x), 0.0f), gzm = fmaxf(fmaf(rf.w, uz, rf.z), 0.0f);
            const float rem = fmaf(-gzm, gzm, fmaf(-gym, gym, rg.cutsq_hi));
            const float rx  = sqrt_approx(fmaxf(rem, 0.0f)) + rg.margin;
            // x bins (relative to the own bin) that intersect [ux - rx, ux + rx]
            const int first = max((int)floorf((ux - rx) * rg.binvx), ri.y), last = min((int)floorf((ux + rx) * rg.binvx), ri.z);
            const int b0 = max(ibin + ri.x + first, 0), b1 = min(ibin + ri.x + last + 1, g.mbins + 1);
            if (live && rem >= 0.0f && last >= first && b1 > b0) {
                s = __ldg(&binstart[b0]);
                e = __ldg(&binstart[b1]);
            }
            if (!__any_sync(0xffffffffu, e > s)) continue;
        }
        for (int c0 = s & ~3; c0 < e;) {
            // groups of 4 candidates in this flush, rounded up to PAIRS of groups: the loop body is two groups with no
            // remainder code behind it; what lies beyond e is padding or the next bin's candidates, masked by vmask below
            const int ng = min(8, (((e - c0 + 3) >> 2) + 1) & ~1);
            unsigned mp = 0, mm = 0;
            auto group = [&](int q) {
                f32x2 X0, X1, Y0, Y1, Z0, Z1;
                ld2x2(cx + c0 + 4 * q, X0, X1);
                ld2x2(cy + c0 + 4 * q, Y0, Y1);
                ld2x2(cz + c0 + 4 * q, Z0, Z1);
                f32x2 dx = sub2(xs2, X0), dy = sub2(ys2, Y0), dz = sub2(zs2, Z0);
                const f32x2 r0 = fma2(dx, dx, fma2(dy, dy, mul2(dz, dz)));
                dx = sub2(xs2, X1); dy = sub2(ys2, Y1); dz = sub2(zs2, Z1);
                const f32x2 r1 = fma2(dx, dx, fma2(dy, dy, mul2(dz, dz)));
                float a, b;
                upk2(sub2(r0, tm2), a, b);
                mm = __funnelshift_l(__float_as_uint(a), mm, 1); mm = __funnelshift_l(__float_as_uint(b), mm, 1);
                upk2(sub2(r1, tm2), a, b);
                mm = __funnelshift_l(__float_as_uint(a), mm, 1); mm = __funnelshift_l(__float_as_uint(b), mm, 1);
                if (sizeof(real) == 8) {
                    upk2(sub2(r0, tp2), a, b);
                    mp = __funnelshift_l(__float_as_uint(a), mp, 1); mp = __funnelshift_l(__float_as_uint(b), mp, 1);
                    upk2(sub2(r1, tp2), a, b);
                    mp = __funnelshift_l(__float_as_uint(a), mp, 1); mp = __funnelshift_l(__float_as_uint(b), mp, 1);
                }
            };
#pragma unroll 1
            for (int q = 0; q < ng; q += 2) {
                group(q);
                group(q + 1);
            }
            if (sizeof(real) == 4) mp = mm;
            const int k = 4 * ng, tlo = max(s - c0, 0), thi = min(e - c0, k);
            const unsigned vmask = (unsigned)((1ull << (k - tlo)) - 1ull) & ~(unsigned)((1ull << (k - thi)) - 1ull);
            unsigned todo = mm & vmask;
            // candidate at bit p is cl[-p] (the first candidate of the flush sits in the highest bit).  The append loop is
            // branch-free: the rare uncertain hits are settled first, both addresses come from 64-bit bases held in registers
            // (mad.wide with an immediate stride), self / overflow are predicates: 30 -> 19 instructions per hit (SASS), list
            // build 5.63 -> 5.07 ms DP, 4.90 -> 4.69 ms SP per rebuild at 128^3 (profiles/r2_s3_call1.sh, r2_s3_call2.sh).
            const int* const cl = cid + c0 + (k - 1);
            unsigned unc = todo & ~mp; // hits of the uncertain band (~1e-4 of the candidates): settled before the append loop
            while (unc) {
                const int p        = bfind(unc);
                const unsigned bit = 1u << p;
                unc ^= bit;
                const int j = __ldg(cl - p);
                const real dx = sub_rn(xt, x[j]), dy = sub_rn(yt, y[j]), dz = sub_rn(zt, z[j]); // the reference's exact expression
                if (!(fma_rn(dx, dx, fma_rn(dy, dy, mul_rn(dz, dz))) <= cutneighsq)) todo ^= bit;
            }
            while (todo) {
                const int p = bfind(todo);
                todo ^= 1u << p;
                const int j = ld_nc(mad_wide_s<-4>(cl, p));
                bool ok = j != i;
                if (HALF) ok = ok && !(j < nlocal && orig[j] < oi); // neighbor.c:224 on reference indices
                if (ok && n < maxneighs) st_global(mad_wide_u<128>(row, (unsigned)n), j); // L.sk == 32 entries (host checks)
                asm("{ .reg .pred q; setp.ne.s32 q, %1, 0; @q add.s32 %0, %0, 1; }" : "+r"(n) : "r"((int)ok)); // one predicated IADD
            }
            c0 += k;
        }
    }
    if (live) numneigh[i] = n;
    n = __reduce_max_sync(0xffffffffu, n);
    if ((threadIdx.x & 31) == 0) atomicMax(max_n, n);
}

// parity read-back: transposed list in internal numbering -> the reference's row-major rows in the
// reference's numbering (extmap: internal index -> reference index, locals and ghosts)
static __global__ void k_untranspose(int nlocal, int row_stride, NbLayout L, const int* __restrict__ numneigh,
    const int* __restrict__ nbT, const int* __restrict__ extmap, int* __restrict__ rows,
    int* __restrict__ numneigh_ext)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nlocal) return;
    const int e     = extmap[i];
    numneigh_ext[e] = numneigh[i];
    if (!rows) return;
    const int n = min(numneigh[i], row_stride);
    for (int k = 0; k < n; k++) rows[(size_t)e * row_stride + k] = extmap[nbT[L.base(i) + (size_t)k * L.sk]];
}

// ---------------------------------------------------------------------------------------------
// LJ 12-6 force, full neighbor lists: verletlist/force_lj.c:14-105.  One thread per local atom, k-major list tiles,
// 4 neighbors in flight per thread.
// ---- v2: same contract, tuned for the FP64 pipe --------------------------------------------------
// * reciprocal by rcp.approx.ftz.f64 (MUFU, ~20 bits) + two Newton steps (4 DFMA) instead of the
//   IEEE division sequence (~9 FP64-pipe instructions + slow-path call); result within ~1 ulp.
// * force = s*s^3*(A*s^3 - B) with A = 48 eps sigma6^2, B = 24 eps sigma6 (5 instead of 7 multiplies)
// * neighbor indices of the NEXT group of 4 are requested before the current group is evaluated.
__device__ __forceinline__ double rcp_nr(double a)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double e = fma(-a, y, 1.0);
    y        = fma(y, e, y);
    e        = fma(-a, y, 1.0);
    y        = fma(y, e, y);
    return y;
}
__device__ __forceinline__ float rcp_nr(float a) { return __frcp_rn(a); }

template <class real> struct LJConst2 {
    real cutforcesq, A, B;
};
template <class real> __device__ __forceinline__ real lj_pair2(real rsq, const LJConst2<real>& c)
{
    const real s  = rcp_nr(rsq);
    const real s3 = s * s * s;
    return (s * s3) * (c.A * s3 - c.B);
}

template <class real, int U>
__device__ __forceinline__ void force_lj_full_v2_body(int nlocal, const LJConst2<real>& c,
    const real* __restrict__ x, const real* __restrict__ y, const real* __restrict__ z,
    const int* __restrict__ numneigh, const int* __restrict__ nbT, const NbLayout& L,
    real* __restrict__ fx, real* __restrict__ fy, real* __restrict__ fz)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nlocal) return;
    const real xt = x[i], yt = y[i], zt = z[i];
    const int nn  = numneigh[i];
    real fix = 0, fiy = 0, fiz = 0;
    const int* nb   = nbT + L.base(i);
    const int nfull = nn - nn % U;
    int j[U], jn[U];
    if (nfull > 0) {
#pragma unroll
        for (int u = 0; u < U; u++) j[u] = __ldg(nb + (size_t)u * L.sk);
    }
    for (int k = 0; k < nfull; k += U) {
        real dx[U], dy[U], dz[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            dx[u] = xt - __ldg(x + j[u]);
            dy[u] = yt - __ldg(y + j[u]);
            dz[u] = zt - __ldg(z + j[u]);
        }
        nb += (size_t)U * L.sk;
        if (k + U < nfull) {
#pragma unroll
            for (int u = 0; u < U; u++) jn[u] = __ldg(nb + (size_t)u * L.sk);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const real rsq = dx[u] * dx[u] + dy[u] * dy[u] + dz[u] * dz[u];
            if (rsq < c.cutforcesq) {
                const real f = lj_pair2(rsq, c);
                fix += dx[u] * f;
                fiy += dy[u] * f;
                fiz += dz[u] * f;
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) j[u] = jn[u];
    }
    for (int k = nfull; k < nn; k++) {
        const int jj  = __ldg(nb);
        nb += L.sk;
        const real dx = xt - __ldg(x + jj), dy = yt - __ldg(y + jj), dz = zt - __ldg(z + jj);
        const real rsq = dx * dx + dy * dy + dz * dz;
        if (rsq < c.cutforcesq) {
            const real f = lj_pair2(rsq, c);
            fix += dx * f;
            fiy += dy * f;
            fiz += dz * f;
        }
    }
    fx[i] = fix;
    fy[i] = fiy;
    fz[i] = fiz;
}

template <class real, int U>
__global__ void __launch_bounds__(128) k_force_lj_full_v2(int nlocal, LJConst2<real> c,
    const real* __restrict__ x, const real* __restrict__ y, const real* __restrict__ z,
    const int* __restrict__ numneigh, const int* __restrict__ nbT, NbLayout L,
    real* __restrict__ fx, real* __restrict__ fy, real* __restrict__ fz)
{
    force_lj_full_v2_body<real, U>(nlocal, c, x, y, z, numneigh, nbT, L, fx, fy, fz);
}
// ---- v6: v2 with the in-cutoff block made branch-free, so that the U pairs in flight interleave ---------------------------
// SASS of v2: every pair's force block is a divergent region holding a 12-deep dependent DFMA/DMUL chain, executed pair
// after pair; the micro-benchmark with perfectly coalesced lists ("seq") therefore reaches only 54 % of the FP64 issue
// rate.  Here the force of all U pairs is computed unconditionally and selected to zero outside the cutoff (a listed pair
// is never at distance 0, so the reciprocal is finite), which lets the scheduler interleave the U chains.
template <class real, int U>
__global__ void __launch_bounds__(128) k_force_lj_full_v6(int nlocal, LJConst2<real> c, const real* __restrict__ x,
    const real* __restrict__ y, const real* __restrict__ z, const int* __restrict__ numneigh, const int* __restrict__ nbT,
    NbLayout L, real* __restrict__ fx, real* __restrict__ fy, real* __restrict__ fz)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nlocal) return;
    const real xt = x[i], yt = y[i], zt = z[i];
    const int nn  = numneigh[i];
    real fix = 0, fiy = 0, fiz = 0;
    const int* nb   = nbT + L.base(i);
    const int nfull = nn - nn % U;
    int j[U], jn[U];
    if (nfull > 0) {
#pragma unroll
        for (int u = 0; u < U; u++) j[u] = __ldg(nb + (size_t)u * L.sk);
    }
    for (int k = 0; k < nfull; k += U) {
        real dx[U], dy[U], dz[U], rsq[U], f[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            dx[u] = xt - __ldg(x + j[u]);
            dy[u] = yt - __ldg(y + j[u]);
            dz[u] = zt - __ldg(z + j[u]);
        }
        nb += (size_t)U * L.sk;
        if (k + U < nfull) {
#pragma unroll
            for (int u = 0; u < U; u++) jn[u] = __ldg(nb + (size_t)u * L.sk);
        }
#pragma unroll
        for (int u = 0; u < U; u++) rsq[u] = dx[u] * dx[u] + dy[u] * dy[u] + dz[u] * dz[u];
#pragma unroll
        for (int u = 0; u < U; u++) f[u] = lj_pair2(rsq[u], c);
#pragma unroll
        for (int u = 0; u < U; u++) {
            const real g = rsq[u] < c.cutforcesq ? f[u] : (real)0;
            fix = fma(dx[u], g, fix); fiy = fma(dy[u], g, fiy); fiz = fma(dz[u], g, fiz);
        }
#pragma unroll
        for (int u = 0; u < U; u++) j[u] = jn[u];
    }
    for (int k = nfull; k < nn; k++) {
        const int jj  = __ldg(nb);
        nb += L.sk;
        const real dx = xt - __ldg(x + jj), dy = yt - __ldg(y + jj), dz = zt - __ldg(z + jj);
        const real rsq = dx * dx + dy * dy + dz * dz;
        const real g   = rsq < c.cutforcesq ? lj_pair2(rsq, c) : (real)0;
        fix = fma(dx, g, fix); fiy = fma(dy, g, fiy); fiz = fma(dz, g, fiz);
    }
    fx[i] = fix;
    fy[i] = fiy;
    fz[i] = fiz;
}

// ---- fi: the force kernel with the velocity-Verlet halves in its epilogue -------------------------------------------------
// Inside the device-resident loop finalIntegrate(n) and initialIntegrate(n+1) follow the force of step n with nothing
// in between (no thermo record, no rebuild before the NEXT force), and both only need atom i's own f, v, x
// (verletlist/integrate.c:21-40).  This kernel therefore keeps f in registers, applies both halves in the same sequence
// of operations as k_final_initial_integrate (bit-identical v and x) and writes the new positions into a SECOND
// coordinate set (other threads still gather the old x[j]); the caller swaps the two sets afterwards.  Saved per step:
// one launch and the 15T-byte integrate pass (f is neither written nor re-read).  BF = branch-free force block (v6,
// SP default), else the divergent block of v2 (DP default).
// XY: the neighbors' x and y come from a packed (x, y) array with ONE 2-element vector gather instead of two scalar
// gathers (the kernel is bound by L1 wavefronts = cache lines touched per request: a warp's k-th neighbors sit in ~4 bins
// of ~7.5 consecutive atoms, which is ~3.3 lines per 8-byte gather but only ~5 per 16-byte one); z stays scalar.  The
// epilogue then also writes the packed copy of the new positions; k_update_pbc refreshes its ghost range.
template <class real> struct FusedIntegrate {
    real *vx, *vy, *vz, *xn, *yn, *zn;
    real dtforce, dt;
    const typename Vec2Of<real>::type* xy;
    typename Vec2Of<real>::type* xyn;
    const real* zg; // ZG: gather copy of z, see below
    real* zgn;
    // single domain: the periodic images of the atom (updatePbc, pbc.c:42-55) are written by the same epilogue -- image b
    // of atom i is ghost nlocal + goff[i] + rank of bit b in gmask[i] (setupPbc's order, k_ghost_fill) -- so that a step
    // between two rebuilds is ONE launch.  gmask == nullptr: ghosts are somebody else's business (bricks, or no ghosts).
    const unsigned* gmask;
    const int* goff;
    real xprd, yprd, zprd;
};
// ZG (with XY; decomposed runs): ALL gathers go to copies -- (x, y) packed and a copy of z -- which are double-buffered,
// while x, y, z themselves are only read and written by the thread that owns the atom and are therefore updated IN
// PLACE (fi.xn/yn/zn = x/y/z).  The neighbor GPUs keep pushing their halo into the ghost range of the same x, y, z
// allocations (no pointer exchange per step); k_pack_gather copies that ghost range into the gather copies.
template <class real, bool XY, bool ZG>
__device__ __forceinline__ void gather_pos(const real* __restrict__ x, const real* __restrict__ y, const real* __restrict__ z,
    const typename Vec2Of<real>::type* __restrict__ xy, const real* __restrict__ zg, int j, real& a, real& b, real& c)
{
    if (XY) {
        const typename Vec2Of<real>::type p = __ldg(xy + j);
        a = p.x; b = p.y;
    } else {
        a = __ldg(x + j); b = __ldg(y + j);
    }
    c = ZG ? __ldg(zg + j) : __ldg(z + j);
}
template <class real>
__global__ void k_pack_gather(int first, int n, const real* __restrict__ x, const real* __restrict__ y, const real* __restrict__ z,
    typename Vec2Of<real>::type* __restrict__ xy, real* __restrict__ zg)
{
    const int i = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= first + n) return;
    typename Vec2Of<real>::type v;
    v.x = x[i]; v.y = y[i];
    xy[i] = v;
    zg[i] = z[i];
}
template <class real, int U, bool BF, bool XY = false, bool ZG = false>
__global__ void __launch_bounds__(128, 8) k_force_lj_full_fi(int nlocal, LJConst2<real> c, const real* __restrict__ x,
    const real* __restrict__ y, const real* __restrict__ z, const int* __restrict__ numneigh, const int* __restrict__ nbT,
    NbLayout L, FusedIntegrate<real> fi)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nlocal) return;
    // the own position comes from the gather copies where there are any (same values; the lines are the ones the neighbors'
    // gathers of this block touch anyway, so x[i], y[i] (and z[i] with ZG) need not pass through L1 as well: force 1.491 ->
    // 1.48 ms DP, 0.948 -> 0.940 ms SP, bricks 1.578 -> 1.559 ms; bit-identical, profiles/r2_s2_n2.txt)
    real xt, yt, zt;
    if (XY) {
        const typename Vec2Of<real>::type p = __ldg(fi.xy + i);
        xt = p.x; yt = p.y;
    } else {
        xt = x[i]; yt = y[i];
    }
    zt = ZG ? __ldg(fi.zg + i) : z[i];
    const int nn  = numneigh[i];
    real fix = 0, fiy = 0, fiz = 0;
    const int* nb   = nbT + L.base(i);
    const int nfull = nn - nn % U;
    int j[U], jn[U];
    if (nfull > 0) {
#pragma unroll
        for (int u = 0; u < U; u++) j[u] = __ldg(nb + (size_t)u * L.sk);
    }
    for (int k = 0; k < nfull; k += U) {
        real dx[U], dy[U], dz[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            real a, b, cc;
            gather_pos<real, XY, ZG>(x, y, z, fi.xy, fi.zg, j[u], a, b, cc);
            dx[u] = xt - a;
            dy[u] = yt - b;
            dz[u] = zt - cc;
        }
        nb += (size_t)U * L.sk;
        if (k + U < nfull) {
#pragma unroll
            for (int u = 0; u < U; u++) jn[u] = __ldg(nb + (size_t)u * L.sk);
        }
        if (BF) {
            real rsq[U], f[U];
#pragma unroll
            for (int u = 0; u < U; u++) rsq[u] = dx[u] * dx[u] + dy[u] * dy[u] + dz[u] * dz[u];
#pragma unroll
            for (int u = 0; u < U; u++) f[u] = lj_pair2(rsq[u], c);
#pragma unroll
            for (int u = 0; u < U; u++) {
                const real g = rsq[u] < c.cutforcesq ? f[u] : (real)0;
                fix = fma(dx[u], g, fix); fiy = fma(dy[u], g, fiy); fiz = fma(dz[u], g, fiz);
            }
        } else {
#pragma unroll
            for (int u = 0; u < U; u++) {
                const real rsq = dx[u] * dx[u] + dy[u] * dy[u] + dz[u] * dz[u];
                if (rsq < c.cutforcesq) {
                    const real f = lj_pair2(rsq, c);
                    fix += dx[u] * f;
                    fiy += dy[u] * f;
                    fiz += dz[u] * f;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) j[u] = jn[u];
    }
    for (int k = nfull; k < nn; k++) {
        const int jj  = __ldg(nb);
        nb += L.sk;
        real pa, pb, pc;
        gather_pos<real, XY, ZG>(x, y, z, fi.xy, fi.zg, jj, pa, pb, pc);
        const real dx = xt - pa, dy = yt - pb, dz = zt - pc;
        const real rsq = dx * dx + dy * dy + dz * dz;
        if (BF) {
            const real g = rsq < c.cutforcesq ? lj_pair2(rsq, c) : (real)0;
            fix = fma(dx, g, fix); fiy = fma(dy, g, fiy); fiz = fma(dz, g, fiz);
        } else if (rsq < c.cutforcesq) {
            const real f = lj_pair2(rsq, c);
            fix += dx * f;
            fiy += dy * f;
            fiz += dz * f;
        }
    }
    // the atom index is re-read from the special registers so that it is not live across the pair loop (the DP
    // kernel sits exactly at 64 registers; one more live value spills inside the loop)
    unsigned tid, bid;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
    asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(bid));
    const int e = (int)(bid * 128u + tid);
    real a = fi.vx[e] + fi.dtforce * fix, b = fi.vy[e] + fi.dtforce * fiy, cc = fi.vz[e] + fi.dtforce * fiz; // final(n)
    a = a + fi.dtforce * fix; b = b + fi.dtforce * fiy; cc = cc + fi.dtforce * fiz;                          // initial(n+1)
    fi.vx[e] = a; fi.vy[e] = b; fi.vz[e] = cc;
    const real xe = xt + fi.dt * a, ye = yt + fi.dt * b, ze = zt + fi.dt * cc;
    fi.xn[e] = xe;
    fi.yn[e] = ye;
    fi.zn[e] = ze;
    if (ZG) fi.zgn[e] = ze;
    if (XY) {
        typename Vec2Of<real>::type v;
        v.x = xe; v.y = ye;
        fi.xyn[e] = v;
    }
    if (!ZG && fi.gmask) { // the atom's periodic images, same single fma as k_update_pbc
        unsigned m = fi.gmask[e];
        if (m) {
            int g = nlocal + fi.goff[e];
            do {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                const real gx = fma_rn((real)c_img[b][0], fi.xprd, xe), gy = fma_rn((real)c_img[b][1], fi.yprd, ye),
                           gz = fma_rn((real)c_img[b][2], fi.zprd, ze);
                fi.xn[g] = gx; fi.yn[g] = gy; fi.zn[g] = gz;
                if (XY) {
                    typename Vec2Of<real>::type v;
                    v.x = gx; v.y = gy;
                    fi.xyn[g] = v;
                }
                g++;
            } while (m);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// LJ with half neighbor lists (verletlist/force_lj.c:107-198): reaction force on local j by native RED.ADD, forces zeroed first.
template <class real, int U>
__global__ void __launch_bounds__(128) k_force_lj_half_v2(int nlocal, LJConst2<real> c, const real* __restrict__ x,
    const real* __restrict__ y, const real* __restrict__ z, const int* __restrict__ numneigh, const int* __restrict__ nbT,
    NbLayout L, real* fx, real* fy, real* fz)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nlocal) return;
    const real xt = x[i], yt = y[i], zt = z[i];
    const int nn  = numneigh[i];
    real fix = 0, fiy = 0, fiz = 0;
    const int* nb   = nbT + L.base(i);
    const int nfull = nn - nn % U;
    int j[U], jn[U];
    if (nfull > 0) {
#pragma unroll
        for (int u = 0; u < U; u++) j[u] = __ldg(nb + (size_t)u * L.sk);
    }
    for (int k = 0; k < nfull; k += U) {
        real dx[U], dy[U], dz[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            dx[u] = xt - __ldg(x + j[u]);
            dy[u] = yt - __ldg(y + j[u]);
            dz[u] = zt - __ldg(z + j[u]);
        }
        nb += (size_t)U * L.sk;
        if (k + U < nfull) {
#pragma unroll
            for (int u = 0; u < U; u++) jn[u] = __ldg(nb + (size_t)u * L.sk);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const real rsq = dx[u] * dx[u] + dy[u] * dy[u] + dz[u] * dz[u];
            if (rsq < c.cutforcesq) {
                const real f  = lj_pair2(rsq, c);
                const real px = dx[u] * f, py = dy[u] * f, pz = dz[u] * f;
                fix += px; fiy += py; fiz += pz;
                if (j[u] < nlocal) { // force_lj.c:171-175
                    atomicAdd(&fx[j[u]], -px);
                    atomicAdd(&fy[j[u]], -py);
                    atomicAdd(&fz[j[u]], -pz);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) j[u] = jn[u];
    }
    for (int k = nfull; k < nn; k++) {
        const int jj  = __ldg(nb);
        nb += L.sk;
        const real dx = xt - __ldg(x + jj), dy = yt - __ldg(y + jj), dz = zt - __ldg(z + jj);
        const real rsq = dx * dx + dy * dy + dz * dz;
        if (rsq < c.cutforcesq) {
            const real f  = lj_pair2(rsq, c);
            const real px = dx * f, py = dy * f, pz = dz * f;
            fix += px; fiy += py; fiz += pz;
            if (jj < nlocal) {
                atomicAdd(&fx[jj], -px);
                atomicAdd(&fy[jj], -py);
                atomicAdd(&fz[jj], -pz);
            }
        }
    }
    atomicAdd(&fx[i], fix);
    atomicAdd(&fy[i], fiy);
    atomicAdd(&fz[i], fiz);
}

// workload counters (the reference's Stats, verletlist/stats.h:13-18): listed pairs and pairs
// inside the force cutoff for the current positions
template <class real>
__global__ void k_count_pairs(int nlocal, real cutforcesq, const real* __restrict__ x,
    const real* __restrict__ y, const real* __restrict__ z, const int* __restrict__ numneigh,
    const int* __restrict__ nbT, NbLayout L, unsigned long long* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long listed = 0, inside = 0;
    if (i < nlocal) {
        const real xt = x[i], yt = y[i], zt = z[i];
        const int nn  = numneigh[i];
        listed        = nn;
        for (int k = 0; k < nn; k++) {
            const int j   = nbT[L.base(i) + (size_t)k * L.sk];
            const real dx = xt - x[j], dy = yt - y[j], dz = zt - z[j];
            if (dx * dx + dy * dy + dz * dz < cutforcesq) inside++;
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        listed += __shfl_down_sync(0xffffffffu, listed, d);
        inside += __shfl_down_sync(0xffffffffu, inside, d);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&out[0], listed);
        atomicAdd(&out[1], inside);
    }
}

// ---------------------------------------------------------------------------------------------
// Synthetic neighbor lists of the reference's kernel micro-benchmark (verletlist/main-stub.c:62-106 createNeighbors):
// pattern 0 "seq": neighbors i+1, i+2, ... (mod Nlocal); 1 "fix": 0 .. nneighs-1 for every atom; 2 "rand": nneighs random
// atoms != i (a hash stream per atom instead of rand(); the reference additionally rejects duplicates, which does not
// change the access pattern).  The first nneighs entries are replicated nreps times.
static __global__ void k_stub_neighbors(int nlocal, int pattern, int nneighs, int nreps, unsigned seed, NbLayout L, int* __restrict__ numneigh,
    int* __restrict__ neighbors)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nlocal) return;
    int* out   = neighbors + L.base(i);
    unsigned h = (pattern >= 3 ? 12345u : seed) ^ (0x9e3779b9u * (unsigned)(i + 1));
    int j      = pattern == 0 ? (i + 1) % nlocal : 0;
    const int m = pattern == 0 ? nlocal : nneighs;
    for (int k = 0; k < nneighs; k++) {
        int v;
        if (pattern == 2) {
            do {
                h ^= h << 13; h ^= h >> 17; h ^= h << 5; // xorshift32
                v = (int)(h % (unsigned)nlocal);
            } while (v == i && nlocal > 1);
        } else if (pattern == 3 || pattern == 4) {
            // diagnostic patterns (not in the reference; profiles/r2_vlforce_analysis.txt): random neighbors inside the
            // atom's own window of W atoms (W = `seed`, a power of two), i.e. L1-resident gathers; 4 additionally gives
            // the 16 lanes of a half warp 16 different residues mod 16 (= different L1 data banks for 8-byte elements)
            h ^= h << 13; h ^= h >> 17; h ^= h << 5;
            const int W  = max(16, (int)seed);
            const int w0 = i & ~(W - 1);
            v = pattern == 3 ? w0 + (int)(h % (unsigned)W) : w0 + 16 * (int)(h % (unsigned)(W / 16)) + ((i + k) & 15);
            v = min(v, nlocal - 1);
        } else {
            v = j;
            j = (j + 1) % m;
        }
        for (int r = 0; r < nreps; r++) out[(size_t)(r * nneighs + k) * L.sk] = v;
    }
    numneigh[i] = nneighs * nreps;
}

// ---------------------------------------------------------------------------------------------
// host-layout conversion: AoS {x,y,z}* <-> SoA
template <class real>
__global__ void k_aos_to_soa(size_t n, const real* __restrict__ a, real* __restrict__ x,
    real* __restrict__ y, real* __restrict__ z)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    x[i] = a[3 * i + 0];
    y[i] = a[3 * i + 1];
    z[i] = a[3 * i + 2];
}
template <class real>
__global__ void k_soa_to_aos(size_t n, const real* __restrict__ x, const real* __restrict__ y,
    const real* __restrict__ z, real* __restrict__ a)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    a[3 * i + 0] = x[i];
    a[3 * i + 1] = y[i];
    a[3 * i + 2] = z[i];
}

} // namespace mdb
