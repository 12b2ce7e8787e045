// cp_kernels.cuh -- sm_100a kernels of the CLUSTERPAIR scheme (GROMACS-style M x N cluster pairs;
// reference src/clusterpair/).  M = 4 i-atoms per cluster like every CPU build of the reference
// (force.h:48); N = 4 or 8 is a template parameter.
//
// Data layout in HBM = the reference's cluster layout (force.h:62-91), because it is
// already what a GPU wants: a TILE of W = N atoms stored as [x0..x(W-1) | y0.. | z0..], tiles
// consecutive.  Tile t is j-cluster t; for N = 8 it also holds the two i-clusters 2t (lanes 0-3) and
// 2t+1 (lanes 4-7).  Ghost j-clusters follow the local ones, then one all-infinity dummy tile.
// A j-cluster is fetched with three (N = 4) or six (N = 8) 256-bit loads in DP, three float4 loads in
// SP, all lanes of an i-cluster reading the same addresses (broadcast): no gathers, which is what
// bounds the verletlist kernel.
#pragma once
#include <cmath>

#include "mdb_util.cuh"

namespace mdb {

constexpr int CP_M = 4;
// Padding lanes of partially filled tiles, and the dummy tile: the reference stores +INFINITY there and relies on
// inf/NaN failing `rsq < cutforcesq`.  The device arrays hold a large FINITE sentinel instead, so
// the force kernels can run branch-free (a masked pair contributes 0 * finite, never 0 * inf); the accessor that
// hands cl_x back to the caller turns the sentinel into +INFINITY again.  Real coordinates are < 1e6.
#define CP_PAD ((real)1.0e15)
#define CP_PAD_MIN ((real)1.0e14)

// geometry of clusterpair/neighbor.c:26-45 (2-D bins = columns along z)
template <class real> struct CpGeom {
    real xprd, yprd, zprd, bininvx, bininvy, cutneigh, cutneighsq, rbb_sq;
    int nbinx, nbiny, mbinxlo, mbinylo, mbinx, mbiny, mbins;
};

// coord2bin2D, clusterpair/neighbor.c:601-618: single rounded multiplies, truncation like C's (int)
template <class real> __device__ __forceinline__ void cp_coord2bin2D(const CpGeom<real>& g, real x, real y, int& ix, int& iy)
{
    ix = axis2bin(x, g.xprd, g.bininvx, g.nbinx, g.mbinxlo);
    iy = axis2bin(y, g.yprd, g.bininvy, g.nbiny, g.mbinylo);
}
template <class real> __device__ __forceinline__ int cp_coord2bin(const CpGeom<real>& g, real x, real y)
{
    int ix, iy;
    cp_coord2bin2D(g, x, y, ix, iy);
    const int b = iy * g.mbinx + ix + 1; // neighbor.c:599 (with its "+ 1")
    return b < 0 ? 0 : (b >= g.mbins ? g.mbins - 1 : b);
}

// ---- index helpers: force.h:62-91 with M = 4 ---------------------------------------------------------
template <int N> __device__ __host__ __forceinline__ int cp_cj0(int ci) { return N == CP_M ? ci : ci >> 1; }
template <int N> __device__ __host__ __forceinline__ size_t cp_ci_base3(int ci) // CI_VECTOR_BASE_INDEX
{
    return N == CP_M ? (size_t)ci * N * 3 : (size_t)(ci >> 1) * N * 3 + (ci & 1) * (N >> 1);
}
template <int N> __device__ __host__ __forceinline__ size_t cp_ci_base1(int ci) // CI_SCALAR_BASE_INDEX
{
    return N == CP_M ? (size_t)ci * N : (size_t)(ci >> 1) * N + (ci & 1) * (N >> 1);
}

// ---- binAtoms (neighbor.c:620-651): histogram / fill; bins then sorted inside k_cp_sort_emit ----------
template <class real>
__global__ void k_cp_bin_count(int n, CpGeom<real> g, const real* __restrict__ x, const real* __restrict__ y,
    int* __restrict__ atom_bin, int* __restrict__ bincount)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned act = __ballot_sync(0xffffffffu, i < n);
    if (i >= n) return;
    const int b = cp_coord2bin(g, x[i], y[i]);
    atom_bin[i] = b;
    // a bin is a whole column of the box (hundreds of atoms) and the atoms of a warp are neighbors in space: one atomic per
    // distinct bin of the warp instead of one per atom
    const unsigned m = __match_any_sync(act, b);
    if ((int)(threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(&bincount[b], __popc(m));
}
// k_bin_fill with the same aggregation: the first lane of each group of equal bins reserves the group's slots
static __global__ void k_cp_bin_fill(int n, const int* __restrict__ atom_bin, const int* __restrict__ binstart, int* __restrict__ cursor,
    int* __restrict__ binatoms)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned act = __ballot_sync(0xffffffffu, i < n);
    if (i >= n) return;
    const int b = atom_bin[i], lane = threadIdx.x & 31;
    const unsigned m = __match_any_sync(act, b);
    const int leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(&cursor[b], __popc(m));
    base = __shfl_sync(m, base, leader);
    binatoms[binstart[b] + base + __popc(m & ((1u << lane) - 1u))] = i;
}
// clusters per bin (neighbor.c:687-690): ceil(c / M), made even when N > M
template <int N> __global__ void k_cp_clusters_per_bin(int mbins, const int* __restrict__ bincount, int* __restrict__ ncl,
    int* __restrict__ maxcount)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    int c       = 0;
    if (b < mbins) {
        c     = bincount[b];
        int k = (c + CP_M - 1) / CP_M;
        if (N > CP_M && (k & 1)) k++;
        ncl[b] = k;
    }
    c = __reduce_max_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c > 0) atomicMax(maxcount, c);
}

// One block per bin: order the bin's atoms like binAtoms + sortAtomsByZCoord (neighbor.c:620-683) and emit the
// bin's clusters (buildClusters, neighbor.c:685-775).
//   binAtoms appends atoms in ascending index; sortAtomsByZCoord is a selection sort with strict '<' and
//   SWAP, so among equal z the final order is not that of a stable sort (SURVEY hard part 3).  If all z
//   of the bin are distinct the result is simply the sorted order (rank sort); otherwise the selection sort
//   is replayed: one parallel arg-min (lowest position wins ties) + swap per round.
template <class real, int N>
__global__ void __launch_bounds__(128) k_cp_sort_emit(int mbins, const int* __restrict__ binstart, const int* __restrict__ binatoms_in,
    const int* __restrict__ clbase, const real* __restrict__ x, const real* __restrict__ y, const real* __restrict__ z,
    const real* __restrict__ vx, const real* __restrict__ vy, const real* __restrict__ vz, const int* __restrict__ tag,
    real* __restrict__ cl_x, real* __restrict__ cl_v, int* __restrict__ cl_tag, int* __restrict__ inat,
    real* __restrict__ ibb, int* __restrict__ ibin)
{
    extern __shared__ unsigned char smem_raw[];
    const int bin = blockIdx.x;
    const int s0 = binstart[bin], c = binstart[bin + 1] - s0;
    if (c == 0) return;
    const int c4 = (c + 3) & ~3;
    real* uz  = reinterpret_cast<real*>(smem_raw); // z of the atoms as the (unordered) fill left them, padded to c4 with +inf
    real* sz  = uz + c4;                           // tie path: z of the atoms in ascending atom index
    int* sid  = reinterpret_cast<int*>(sz + c4);   // tie path: the atoms in ascending atom index
    int* sout = sid + c;                           // result order
    int* uid  = sout + c;                          // the atoms as the fill left them
    __shared__ int s_ties;
    const int t = threadIdx.x;
    if (t == 0) s_ties = 0;
    for (int k = t; k < c4; k += blockDim.x) {
        const int v = k < c ? binatoms_in[s0 + k] : 0;
        if (k < c) { uid[k] = v; sout[k] = -1; }
        uz[k] = k < c ? z[v] : (real)INFINITY;
    }
    __syncthreads();
    // binAtoms fills a bin in ascending atom index and a selection sort of DISTINCT keys ends in sorted order, so without
    // ties the result is the rank by z -- ONE pass over the pairs of the bin, straight from the unordered fill, four z per
    // shared-memory load.  Equal z give equal ranks, i.e. a slot of the result stays empty: that is how ties are found
    // (no equality count in the hot loop), and any tie sends the whole bin through the replay below.
    for (int k = t; k < c; k += blockDim.x) {
        const real zk = uz[k];
        int r = 0;
        for (int q = 0; q < c4; q += 4) {
            real v0, v1, v2, v3;
            if constexpr (sizeof(real) == 4) {
                const float4 v = *reinterpret_cast<const float4*>(uz + q);
                v0 = v.x; v1 = v.y; v2 = v.z; v3 = v.w;
            } else {
                const double2 a = *reinterpret_cast<const double2*>(uz + q), b = *reinterpret_cast<const double2*>(uz + q + 2);
                v0 = a.x; v1 = a.y; v2 = b.x; v3 = b.y;
            }
            r += (v0 < zk) + (v1 < zk) + (v2 < zk) + (v3 < zk);
        }
        sout[r] = uid[k]; // r < c: at most c - 1 of the c finite z are smaller
    }
    __syncthreads();
    for (int k = t; k < c; k += blockDim.x)
        if (sout[k] < 0) s_ties = 1;
    __syncthreads();
    if (s_ties) { // equal z in the bin (every bin of the initial lattice): replay the selection sort on (sz, sid)
        for (int k = t; k < c; k += blockDim.x) { // ascending atom index = the order binAtoms filled the bin
            const int v = uid[k];
            int r       = 0;
            for (int q = 0; q < c; q++) r += uid[q] < v;
            sid[r] = v;
            sz[r]  = uz[k];
        }
        __syncthreads();
        // the rounds depend on one another, so ONE warp walks them with shuffles only (no block barrier per round:
        // 5.7 -> 4.3 ms per build of the initial lattice at 128^3, profiles/r2_s3_call15.sh); the other warps wait below
        if (t < 32) {
            for (int a = 0; a < c - 1; a++) {
                real mz = INFINITY;
                int mp  = 0x7fffffff;
                for (int k = a + t; k < c; k += 32) {
                    const real zk = sz[k];
                    if (zk < mz) { mz = zk; mp = k; } // ascending k: first position wins
                }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    const real oz = __shfl_xor_sync(0xffffffffu, mz, d);
                    const int op  = __shfl_xor_sync(0xffffffffu, mp, d);
                    if (oz < mz || (oz == mz && op < mp)) { mz = oz; mp = op; }
                }
                if (mp >= c) mp = a; // non-finite z (blown-up run): keep the element in place
                if (t == 0) { // neighbor.c:676-677: bin_ptr[ac_i] = min_idx; bin_ptr[min_ac] = i
                    const int ia = sid[a];
                    const real za = sz[a];
                    sid[a] = sid[mp]; sz[a] = sz[mp];
                    sid[mp] = ia; sz[mp] = za;
                }
                __syncwarp();
            }
        }
        __syncthreads();
        for (int k = t; k < c; k += blockDim.x) sout[k] = sid[k];
        __syncthreads();
    }
    // emit clusters of this bin
    int ncl = (c + CP_M - 1) / CP_M;
    if (N > CP_M && (ncl & 1)) ncl++;
    const int base = clbase[bin];
    for (int k = t; k < ncl; k += blockDim.x) {
        const int ci = base + k;
        real* cx = cl_x + cp_ci_base3<N>(ci);
        real* cv = cl_v + cp_ci_base3<N>(ci);
        int* ct  = cl_tag + cp_ci_base1<N>(ci);
        real lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
        int nat = 0;
#pragma unroll
        for (int cii = 0; cii < CP_M; cii++) {
            const int ac = k * CP_M + cii;
            if (ac < c) {
                const int i = sout[ac];
                const real p[3] = { x[i], y[i], z[i] };
                cx[cii] = p[0]; cx[N + cii] = p[1]; cx[2 * N + cii] = p[2];
                cv[cii] = vx[i]; cv[N + cii] = vy[i]; cv[2 * N + cii] = vz[i];
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    if (lo[d] > p[d]) lo[d] = p[d];
                    if (hi[d] < p[d]) hi[d] = p[d];
                }
                ct[cii] = tag[i];
                nat++;
            } else {
                cx[cii] = CP_PAD; cx[N + cii] = CP_PAD; cx[2 * N + cii] = CP_PAD;
                cv[cii] = 0; cv[N + cii] = 0; cv[2 * N + cii] = 0;
                ct[cii] = -1;
            }
        }
        inat[ci] = nat;
        ibin[ci] = bin;
        real* bb = ibb + (size_t)ci * 6;
        bb[0] = lo[0]; bb[1] = hi[0]; bb[2] = lo[1]; bb[3] = hi[1]; bb[4] = lo[2]; bb[5] = hi[2];
    }
}

// defineJClusters, neighbor.c:777-895 (M == N: copy; 2M == N: union of the two i-clusters of the tile)
template <class real, int N>
__global__ void k_cp_define_j(int ncj, const int* __restrict__ inat, const real* __restrict__ ibb, int* __restrict__ jnat,
    real* __restrict__ jbb)
{
    const int cj = blockIdx.x * blockDim.x + threadIdx.x;
    if (cj >= ncj) return;
    real* o = jbb + (size_t)cj * 6;
    if (N == CP_M) {
        const real* a = ibb + (size_t)cj * 6;
#pragma unroll
        for (int k = 0; k < 6; k++) o[k] = a[k];
        jnat[cj] = inat[cj];
    } else {
        const real *a = ibb + (size_t)(2 * cj) * 6, *b = ibb + (size_t)(2 * cj + 1) * 6;
#pragma unroll
        for (int d = 0; d < 3; d++) {
            o[2 * d]     = a[2 * d] < b[2 * d] ? a[2 * d] : b[2 * d];             // MIN
            o[2 * d + 1] = a[2 * d + 1] > b[2 * d + 1] ? a[2 * d + 1] : b[2 * d + 1]; // MAX
        }
        jnat[cj] = inat[2 * cj] + inat[2 * cj + 1];
    }
}

// ---- setupPbc (pbc.c:183-323): ghost j-clusters by bounding box, ADDGHOST ladder order ---------------------
template <class real>
__global__ void k_cp_ghost_count(int ncj, PbcGeom<real> g, const int* __restrict__ jnat, const real* __restrict__ jbb,
    unsigned* __restrict__ mask, int* __restrict__ count)
{
    const int cj = blockIdx.x * blockDim.x + threadIdx.x;
    if (cj >= ncj) return;
    unsigned m = 0;
    if (jnat[cj] > 0) {
        const real* b = jbb + (size_t)cj * 6;
        const bool lo[3] = { b[0] < g.cutneigh, b[2] < g.cutneigh, b[4] < g.cutneigh };
        const bool hi[3] = { b[1] >= g.xhi_cut, b[3] >= g.yhi_cut, b[5] >= g.zhi_cut };
#pragma unroll
        for (int q = 0; q < 26; q++) {
            bool ok = true;
#pragma unroll
            for (int a = 0; a < 3; a++) {
                const int d = c_img[q][a];
                if (d > 0) ok = ok && lo[a];
                if (d < 0) ok = ok && hi[a];
            }
            if (ok) m |= 1u << q;
        }
    }
    mask[cj]  = m;
    count[cj] = __popc(m);
}
template <int N>
__global__ void k_cp_ghost_fill(int ncj, const unsigned* __restrict__ mask, const int* __restrict__ offset,
    const int* __restrict__ cl_tag_in, int* __restrict__ border_map, int* __restrict__ code, int* __restrict__ jnat,
    int* __restrict__ cl_tag)
{
    const int cj = blockIdx.x * blockDim.x + threadIdx.x;
    if (cj >= ncj) return;
    unsigned m = mask[cj];
    int g      = offset[cj];
    const int nat = jnat[cj];
    while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        border_map[g]  = cj;
        code[g]        = (c_img[b][0] + 1) | ((c_img[b][1] + 1) << 2) | ((c_img[b][2] + 1) << 4);
        jnat[ncj + g]  = nat;
        for (int q = 0; q < N; q++) cl_tag[(size_t)(ncj + g) * N + q] = q < nat ? cl_tag_in[(size_t)cj * N + q] : -1;
        g++;
    }
}
// updatePbcCPU, pbc.c:45-114: one thread per ghost tile lane; image = fma(PBC, prd, source) (ONE fma);
// first != 0 also pads the tile with infinity and computes the bounding box (one thread per tile then)
template <class real, int N>
__global__ void k_cp_update_pbc(int ncj, int nghost, real xprd, real yprd, real zprd, const int* __restrict__ border_map,
    const int* __restrict__ code, const int* __restrict__ jnat, real* __restrict__ cl_x)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nghost * N) return;
    const int g = t / N, q = t % N;
    const int cj = ncj + g;
    if (q >= jnat[cj]) return;
    const int c = code[g];
    const real* s = cl_x + (size_t)border_map[g] * N * 3;
    real* d       = cl_x + (size_t)cj * N * 3;
    d[q]         = fma_rn((real)((c & 3) - 1), xprd, s[q]);
    d[N + q]     = fma_rn((real)(((c >> 2) & 3) - 1), yprd, s[N + q]);
    d[2 * N + q] = fma_rn((real)(((c >> 4) & 3) - 1), zprd, s[2 * N + q]);
}
template <class real, int N>
__global__ void k_cp_update_pbc_first(int ncj, int nghost, real xprd, real yprd, real zprd, const int* __restrict__ border_map,
    const int* __restrict__ code, const int* __restrict__ jnat, real* __restrict__ cl_x, real* __restrict__ jbb)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > nghost) return;
    const int cj = ncj + g;
    real* d = cl_x + (size_t)cj * N * 3;
    if (g == nghost) { // the dummy cluster at the end, pbc.c:304-311
        for (int q = 0; q < 3 * N; q++) d[q] = CP_PAD;
        return;
    }
    const int c = code[g], nat = jnat[cj];
    const real* s = cl_x + (size_t)border_map[g] * N * 3;
    const real sh[3] = { (real)((c & 3) - 1), (real)(((c >> 2) & 3) - 1), (real)(((c >> 4) & 3) - 1) };
    const real prd[3] = { xprd, yprd, zprd };
    real* bb = jbb + (size_t)cj * 6;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        real lo = INFINITY, hi = -INFINITY;
        for (int q = 0; q < N; q++) {
            if (q < nat) {
                const real v = fma_rn(sh[a], prd[a], s[a * N + q]);
                d[a * N + q] = v;
                if (lo > v) lo = v;
                if (hi < v) hi = v;
            } else d[a * N + q] = CP_PAD;
        }
        bb[2 * a] = lo; bb[2 * a + 1] = hi;
    }
}

// ---- binClusters (neighbor.c:897-1043): bin of every j-cluster ------------------------------------------------
// local j-cluster: the bin of its i-cluster(s); ghost: the bin of its INNERMOST atom (neighbor.c:957-985)
template <class real, int N>
__global__ void k_cp_cluster_bin(int ncj, int nghost, CpGeom<real> g, const int* __restrict__ ibin, const int* __restrict__ code,
    const int* __restrict__ jnat, const real* __restrict__ cl_x, int* __restrict__ cbin, int* __restrict__ cbincount)
{
    const int cj = blockIdx.x * blockDim.x + threadIdx.x;
    if (cj >= ncj + nghost) return;
    int bin = -1;
    if (cj < ncj) {
        bin = ibin[N == CP_M ? cj : 2 * cj];
    } else if (jnat[cj] > 0) {
        const int c = code[cj - ncj];
        const int px = (c & 3) - 1, py = ((c >> 2) & 3) - 1;
        const real* p = cl_x + (size_t)cj * N * 3;
        int ix, iy;
        cp_coord2bin2D(g, p[0], p[N], ix, iy);
        ix = max(min(ix, g.mbinx - 1), 0);
        iy = max(min(iy, g.mbiny - 1), 0);
        for (int q = 1; q < jnat[cj]; q++) {
            int nix, niy;
            cp_coord2bin2D(g, p[q], p[N + q], nix, niy);
            nix = max(min(nix, g.mbinx - 1), 0);
            niy = max(min(niy, g.mbiny - 1), 0);
            if (px > 0 && ix > nix) ix = nix;
            if (px < 0 && ix < nix) ix = nix;
            if (py > 0 && iy > niy) iy = niy;
            if (py < 0 && iy < niy) iy = niy;
        }
        bin = iy * g.mbinx + ix + 1;
        bin = bin >= g.mbins ? g.mbins - 1 : bin;
    }
    cbin[cj] = bin;
    if (bin >= 0) atomicAdd(&cbincount[bin], 1);
}
static __global__ void k_cp_cluster_fill(int n, const int* __restrict__ cbin, const int* __restrict__ cbinstart, int* __restrict__ cursor,
    int* __restrict__ cbinlist)
{
    const int cj = blockIdx.x * blockDim.x + threadIdx.x;
    if (cj >= n) return;
    const int b = cbin[cj];
    if (b >= 0) cbinlist[cbinstart[b] + atomicAdd(&cursor[b], 1)] = cj;
}
// per bin: clusters ascending in (bbminz, id) -- the reference keeps its bins z-sorted too (neighbor.c:989-1012; the
// ORDER inside a bin does not change the neighbor SETS) -- plus the running maximum of bbmaxz, which lets the list
// build binary-search the first cluster that can be in z-range.  One warp per bin, rank sort.
template <class real>
__global__ void __launch_bounds__(128) k_cp_cluster_sort(int mbins, const int* __restrict__ cbinstart,
    const int* __restrict__ cbinlist_in, int* cbinlist, const real* __restrict__ jbb, real* jbbs, real* __restrict__ pmaxz)
{
    typedef typename Vec2Of<real>::type vec2;
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (b >= mbins) return;
    const int s = cbinstart[b], c = cbinstart[b + 1] - s;
    for (int k = lane; k < c; k += 32) {
        const int v   = cbinlist_in[s + k];
        const real zv = jbb[(size_t)v * 6 + 4];
        int r         = 0;
        for (int q = 0; q < c; q++) {
            const int u   = cbinlist_in[s + q];
            const real zu = jbb[(size_t)u * 6 + 4];
            r += (zu < zv) || (zu == zv && u < v);
        }
        cbinlist[s + r] = v;
        // the bounding boxes once more IN BIN ORDER: the list build reads them by position (consecutive clusters of a bin
        // are consecutive in memory) instead of through the cluster index
        const vec2* src = reinterpret_cast<const vec2*>(jbb + (size_t)v * 6);
        vec2* dst       = reinterpret_cast<vec2*>(jbbs + (size_t)(s + r) * 6);
        dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
    }
    __syncwarp();
    for (int k = lane; k < c; k += 32) {
        real m = -INFINITY;
        for (int q = 0; q <= k; q++) {
            const real zmax = jbbs[(size_t)(s + q) * 6 + 5];
            m               = zmax > m ? zmax : m;
        }
        pmaxz[s + k] = m;
    }
}

// ---- buildNeighborCPU (neighbor.c:262-481): cluster-pair list -------------------------------------------------------
// One thread per i-cluster.  Set semantics of the reference: every j-cluster of the stencil bins with
// d_bb_sq < cutneighsq and (d_bb_sq < rbb_sq or some atom pair closer than cutneigh); half lists keep cj >= CJ1(ci);
// the diagonal entry (cj == CJ0(ci)) is moved to the front and counted in numneigh_masked (neighbor.c:374-385).
// d_bb_sq accumulates z, y, x with the contraction of the reference build; atomDistanceInRange is
// fma(dz,dz,fma(dx,dx,dy*dy)) < cutneighsq (disassembly of the reference build).
template <class real, int N>
__global__ void __launch_bounds__(128) k_cp_build_neighbor(int ncl, int half, CpGeom<real> g, const int* __restrict__ stencil,
    int nstencil, const int* __restrict__ ibin, const int* __restrict__ inat, const real* __restrict__ ibb,
    const int* __restrict__ jnat, const real* __restrict__ jbbs, const real* __restrict__ cl_x, const int* __restrict__ cbinstart,
    const int* __restrict__ cbinlist, const real* __restrict__ pmaxz, int maxneighs, int* __restrict__ numneigh,
    int* __restrict__ numneigh_masked, int* __restrict__ neighbors, int* __restrict__ max_n)
{
    const int ci = blockIdx.x * blockDim.x + threadIdx.x;
    int n = 0;
    if (ci < ncl) {
        const real* I = ibb + (size_t)ci * 6;
        const real ixlo = I[0], ixhi = I[1], iylo = I[2], iyhi = I[3], izlo = I[4], izhi = I[5];
        const int bin = ibin[ci], nati = inat[ci], self = cp_cj0<N>(ci);
        const real* xi = cl_x + cp_ci_base3<N>(ci);
        // the i atoms; slots beyond nati are moved to -CP_PAD (the padding slots of the j tiles sit at +CP_PAD), so that the atom
        // test below needs no `a < nati` / `b < natj`: such pairs fail the test like the reference's skipped pairs (neighbor.c:216-234)
        real px[CP_M], py[CP_M], pz[CP_M];
#pragma unroll
        for (int q = 0; q < CP_M; q++) {
            const bool v = q < nati;
            px[q] = v ? xi[q] : -CP_PAD; py[q] = v ? xi[N + q] : -CP_PAD; pz[q] = v ? xi[2 * N + q] : -CP_PAD;
        }
        int* row = neighbors + (size_t)ci * maxneighs;
        int nmasked = 0;
        const real zlo_need = izlo - g.cutneigh; // clusters entirely below this cannot be in range
        typedef typename Vec2Of<real>::type vec2;
        for (int k = 0; k < nstencil; k++) {
            const int jb = bin + __ldg(stencil + k);
            if (jb < 0 || jb >= g.mbins) continue;
            const int s = __ldg(cbinstart + jb), e = __ldg(cbinstart + jb + 1);
            // first position whose running max of bbmaxz reaches zlo_need (conservative: the exact test follows)
            int lo = s, hi = e;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (__ldg(pmaxz + mid) < zlo_need - (real)1e-3) lo = mid + 1;
                else hi = mid;
            }
            for (int m = lo; m < e; m++) {
                const int cj = __ldg(cbinlist + m);
                const vec2* J = reinterpret_cast<const vec2*>(jbbs + (size_t)m * 6); // (xlo, xhi) (ylo, yhi) (zlo, zhi), bin order
                const vec2 jz = __ldg(J + 2);
                if (jz.x - izhi > g.cutneigh + (real)1e-3) break; // sorted by bbminz: nothing further can be in range
                if (half && self > cj) continue;                  // neighbor.c:318: ci_cj1 > cj
                const vec2 jy = __ldg(J + 1), jx = __ldg(J);
                real dl, dh, dm, d2;
                dl = izlo - jz.y; dh = jz.x - izhi; dm = fmax(fmax(dl, dh), (real)0);
                d2 = mul_rn(dm, dm);
                dl = iylo - jy.y; dh = jy.x - iyhi; dm = fmax(fmax(dl, dh), (real)0);
                d2 = fma_rn(dm, dm, d2);
                dl = ixlo - jx.y; dh = jx.x - ixhi; dm = fmax(fmax(dl, dh), (real)0);
                d2 = fma_rn(dm, dm, d2);
                if (!(d2 < g.cutneighsq)) continue;
                bool in = d2 < g.rbb_sq;
                if (!in) { // atomDistanceInRange, neighbor.c:216-234, as straight-line code: the whole j tile arrives by 128-bit
                    // loads, padding slots of the tile sit at +CP_PAD and invalid i slots at -CP_PAD (see above), so no pair of
                    // them passes the test and neither nati nor natj is consulted; float: two j atoms per packed operation
                    const real* xj = cl_x + (size_t)cj * N * 3;
                    if constexpr (sizeof(real) == 4) {
                        f32x2 X[N / 2], Y[N / 2], Z[N / 2];
#pragma unroll
                        for (int q = 0; q < N / 4; q++) {
                            ld2x2((const float*)xj + 4 * q, X[2 * q], X[2 * q + 1]);
                            ld2x2((const float*)xj + N + 4 * q, Y[2 * q], Y[2 * q + 1]);
                            ld2x2((const float*)xj + 2 * N + 4 * q, Z[2 * q], Z[2 * q + 1]);
                        }
#pragma unroll
                        for (int a = 0; a < CP_M; a++) {
                            const f32x2 xa = pk2((float)px[a], (float)px[a]), ya = pk2((float)py[a], (float)py[a]), za = pk2((float)pz[a], (float)pz[a]);
#pragma unroll
                            for (int q = 0; q < N / 2; q++) {
                                const f32x2 dx = sub2(xa, X[q]), dy = sub2(ya, Y[q]), dz = sub2(za, Z[q]);
                                float r0, r1;
                                upk2(fma2(dz, dz, fma2(dx, dx, mul2(dy, dy))), r0, r1);
                                in = in | (r0 < (float)g.cutneighsq) | (r1 < (float)g.cutneighsq);
                            }
                        }
                    } else {
                        typedef typename Vec2Of<real>::type v2;
                        v2 X[N / 2], Y[N / 2], Z[N / 2];
#pragma unroll
                        for (int q = 0; q < N / 2; q++) {
                            X[q] = __ldg(reinterpret_cast<const v2*>(xj) + q);
                            Y[q] = __ldg(reinterpret_cast<const v2*>(xj + N) + q);
                            Z[q] = __ldg(reinterpret_cast<const v2*>(xj + 2 * N) + q);
                        }
#pragma unroll
                        for (int a = 0; a < CP_M; a++) {
#pragma unroll
                            for (int q = 0; q < N / 2; q++) {
                                real dx = sub_rn(px[a], X[q].x), dy = sub_rn(py[a], Y[q].x), dz = sub_rn(pz[a], Z[q].x);
                                in = in | (fma_rn(dz, dz, fma_rn(dx, dx, mul_rn(dy, dy))) < g.cutneighsq);
                                dx = sub_rn(px[a], X[q].y); dy = sub_rn(py[a], Y[q].y); dz = sub_rn(pz[a], Z[q].y);
                                in = in | (fma_rn(dz, dz, fma_rn(dx, dx, mul_rn(dy, dy))) < g.cutneighsq);
                            }
                        }
                    }
                }
                if (in) {
                    if (n < maxneighs) {
                        if (cj != self) row[n] = cj;
                        else { row[n] = row[nmasked]; row[nmasked] = cj; nmasked++; }
                    }
                    n++;
                }
            }
        }
        numneigh[ci]        = n;
        numneigh_masked[ci] = nmasked;
    }
    n = __reduce_max_sync(0xffffffffu, n);
    if ((threadIdx.x & 31) == 0) atomicMax(max_n, n);
}

// pruneNeighbor (neighbor.c:483-531): drop listed cluster pairs without any atom pair inside cutneigh (the pairs
// the bounding-box shortcut d_bb_sq < rbb_sq admitted), compacting each row exactly like the reference (the last
// entry moves into the hole; numneigh_masked shrinks when the hole is in the masked prefix).
template <class real, int N>
__global__ void __launch_bounds__(128) k_cp_prune(int ncl, real cutsq, const int* __restrict__ inat, const int* __restrict__ jnat,
    const real* __restrict__ cl_x, int maxneighs, int* __restrict__ numneigh, int* __restrict__ numneigh_masked,
    int* __restrict__ neighbors)
{
    const int ci = blockIdx.x * blockDim.x + threadIdx.x;
    if (ci >= ncl) return;
    const real* xi = cl_x + cp_ci_base3<N>(ci);
    const int nati = inat[ci];
    real px[CP_M], py[CP_M], pz[CP_M];
#pragma unroll
    for (int q = 0; q < CP_M; q++) { px[q] = xi[q]; py[q] = xi[N + q]; pz[q] = xi[2 * N + q]; }
    int* row = neighbors + (size_t)ci * maxneighs;
    int n = numneigh[ci], nm = numneigh_masked[ci], k = 0;
    while (k < n) {
        const int cj   = row[k];
        const real* xj = cl_x + (size_t)cj * N * 3;
        const int natj = jnat[cj];
        bool in        = false;
        for (int b = 0; b < natj && !in; b++) {
            const real xb = xj[b], yb = xj[N + b], zb = xj[2 * N + b];
#pragma unroll
            for (int a = 0; a < CP_M; a++)
                if (a < nati) {
                    const real dx = sub_rn(px[a], xb), dy = sub_rn(py[a], yb), dz = sub_rn(pz[a], zb);
                    in = in || (fma_rn(dz, dz, fma_rn(dx, dx, mul_rn(dy, dy))) < cutsq);
                }
        }
        if (in) k++;
        else {
            n--;
            if (k < nm) nm--;
            row[k] = row[n];
        }
    }
    numneigh[ci]        = n;
    numneigh_masked[ci] = nm;
}

// ---- force: computeForceLJRef semantics (force_lj.c:47-164) ---------------------------------------------------------
// One lane per i-atom slot; the 4 lanes of an i-cluster walk the same list row and fetch each j-cluster tile
// with 256-bit (DP) / 128-bit (SP) loads from identical addresses (broadcast).  Each lane evaluates its atom against
// the N atoms of the tile in registers.  Exclusion on the diagonal tile: full lists skip j == own lane, half lists
// keep j > own lane (force_lj.c:99-113).  Padding lanes sit at +infinity: rsq is inf / NaN and fails the cutoff
// test (no fast-math anywhere in this library).
template <class real> struct CpTileLoad;
__device__ __forceinline__ float rcp_fast(float a);
__device__ __forceinline__ double rcp_fast(double a);
template <class real> __device__ __forceinline__ real lj_pair_fast(real rsq, const LJConst2<real>& c)
{
    const real s  = rcp_fast(rsq);
    const real s3 = s * s * s;
    return (s * s3) * (c.A * s3 - c.B);
}
template <> struct CpTileLoad<double> {
    template <int N> static __device__ __forceinline__ void load(const double* p, double (&v)[N])
    {
#pragma unroll
        for (int q = 0; q < N; q += 4)
            asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[q]), "=d"(v[q + 1]), "=d"(v[q + 2]), "=d"(v[q + 3]) : "l"(p + q));
    }
    // positions change between force calls inside one graph of kernels: plain (coherent) loads for the half kernel's
    // force tiles are not needed; positions are read-only during a force kernel, so .nc is valid.
};
template <> struct CpTileLoad<float> {
    template <int N> static __device__ __forceinline__ void load(const float* p, float (&v)[N])
    {
#pragma unroll
        for (int q = 0; q < N; q += 4) {
            const float4 f = __ldg(reinterpret_cast<const float4*>(p + q));
            v[q] = f.x; v[q + 1] = f.y; v[q + 2] = f.z; v[q + 3] = f.w;
        }
    }
};

// reciprocal for the pair kernels: SP = MUFU.RCP + one Newton step (2 FFMA, ~1 ulp), DP = rcp_nr (vl_kernels.cuh)
__device__ __forceinline__ float rcp_fast(float a)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(a));
    const float e = fmaf(-a, y, 1.0f);
    return fmaf(y, e, y);
}
__device__ __forceinline__ double rcp_fast(double a) { return rcp_nr(a); }

// one j tile against the lane's i-atom, branch-free.  DIAG: the tile is the i-cluster's own j-cluster.
template <class real, int N, bool HALF, bool DIAG>
__device__ __forceinline__ void cp_tile(const real* __restrict__ t, const LJConst2<real>& c, real xt, real yt, real zt, int ii,
    int cii, bool react, real* __restrict__ fj, real& fix, real& fiy, real& fiz)
{
    real xj[N], yj[N], zj[N];
    CpTileLoad<real>::template load<N>(t, xj);
    CpTileLoad<real>::template load<N>(t + N, yj);
    CpTileLoad<real>::template load<N>(t + 2 * N, zj);
    real rx[N], ry[N], rz[N];
#pragma unroll
    for (int q = 0; q < N; q++) {
        const real dx = xt - xj[q], dy = yt - yj[q], dz = zt - zj[q];
        const real rsq = dx * dx + dy * dy + dz * dz;
        bool in = rsq < c.cutforcesq;
        if (DIAG) in = in && (HALF ? (ii < q) : (ii != q)); // force_lj.c:99-113
        real f = lj_pair_fast(rsq, c);
        f      = in ? f : (real)0; // rsq == 0 on the excluded self pair gives inf/NaN: discarded here, never multiplied
        if (HALF) {
            rx[q] = dx * f; ry[q] = dy * f; rz[q] = dz * f;
            fix += rx[q]; fiy += ry[q]; fiz += rz[q];
        } else {
            fix = fma(dx, f, fix); fiy = fma(dy, f, fiy); fiz = fma(dz, f, fiz);
        }
    }
    if (HALF) { // reaction on the j tile: sum over the 4 lanes of the i-cluster, then one atomic per j atom and lane
#pragma unroll
        for (int q = 0; q < N; q++) {
            real sx = rx[q], sy = ry[q], sz = rz[q];
            sx += __shfl_xor_sync(0xffffffffu, sx, 1); sy += __shfl_xor_sync(0xffffffffu, sy, 1); sz += __shfl_xor_sync(0xffffffffu, sz, 1);
            sx += __shfl_xor_sync(0xffffffffu, sx, 2); sy += __shfl_xor_sync(0xffffffffu, sy, 2); sz += __shfl_xor_sync(0xffffffffu, sz, 2);
            if (react && cii == (q & 3) && (sx != 0 || sy != 0 || sz != 0)) {
                atomicAdd(fj + q, -sx); atomicAdd(fj + N + q, -sy); atomicAdd(fj + 2 * N + q, -sz);
            }
        }
    }
}

// ---- SP, full lists: the same tile with PACKED FP32 arithmetic (sm_100 add/mul/fma.f32x2 -> FADD2 / FMUL2 / FFMA2).
// The scalar kernel is bound by instruction issue (ncu: issue 81 %, FMA pipe 55 %, profiles/r1_s3_cp_force_raw.txt).
// A packed instruction does two lanes' worth of work for one issue slot (measured: FFMA2 reaches the same 72.7 TFLOP/s
// as FFMA at half the issue rate and keeps it with integer work mixed in, profiles/ubench_ffma2.cu), so evaluating
// j atoms (q, q+1) together moves the bound from issue to the FMA pipe itself.
struct CpPackedConst {
    f32x2 A, negB, one;
    float cutforcesq;
};
template <int N> struct CpTileRegs {
    f32x2 x[N / 2], y[N / 2], z[N / 2];
    __device__ __forceinline__ void load(const float* __restrict__ t)
    {
#pragma unroll
        for (int q = 0; q < N / 2; q += 2) { // one 128-bit load = two packed pairs
            asm volatile("ld.global.nc.v2.u64 {%0, %1}, [%2];" : "=l"(x[q]), "=l"(x[q + 1]) : "l"(t + 2 * q));
            asm volatile("ld.global.nc.v2.u64 {%0, %1}, [%2];" : "=l"(y[q]), "=l"(y[q + 1]) : "l"(t + N + 2 * q));
            asm volatile("ld.global.nc.v2.u64 {%0, %1}, [%2];" : "=l"(z[q]), "=l"(z[q + 1]) : "l"(t + 2 * N + 2 * q));
        }
    }
};
// iq = the lane's own slot inside this tile if the tile is the i-cluster's own j-cluster, else -1
template <int N>
__device__ __forceinline__ void cp_tile_packed(const CpTileRegs<N>& T, const CpPackedConst& c, f32x2 xt, f32x2 yt, f32x2 zt, int iq,
    f32x2& fx, f32x2& fy, f32x2& fz)
{
#pragma unroll
    for (int q = 0; q < N / 2; q++) {
        const f32x2 dx = sub2(xt, T.x[q]), dy = sub2(yt, T.y[q]), dz = sub2(zt, T.z[q]);
        const f32x2 rsq = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
        float r0, r1, y0, y1;
        upk2(rsq, r0, r1);
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(r0));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y1) : "f"(r1));
        const bool in0 = r0 < c.cutforcesq && iq != 2 * q, in1 = r1 < c.cutforcesq && iq != 2 * q + 1; // force_lj.c:99-113
        f32x2 y        = pk2(y0, y1);
        const f32x2 nr = sub2(c.one, mul2(rsq, y)); // Newton step: y += y * (1 - rsq * y)
        y              = fma2(y, nr, y);
        const f32x2 s3 = mul2(mul2(y, y), y);
        f32x2 f        = mul2(mul2(y, s3), fma2(c.A, s3, c.negB));
        float f0, f1;
        upk2(f, f0, f1);
        f  = pk2(in0 ? f0 : 0.0f, in1 ? f1 : 0.0f); // excluded self pair: rsq = 0 gives inf/NaN, discarded here
        fx = fma2(dx, f, fx); fy = fma2(dy, f, fy); fz = fma2(dz, f, fz);
    }
}
// FI (full lists, inside mdb_cp_run): finalIntegrate(n) + initialIntegrate(n+1) in the force kernel's epilogue, on the
// force still in registers -- the same operations as k_cp_integrate<MODE 2> (integrate.c:23-63), so bit-identical.  The new
// positions go to a second cluster array (other lanes still read the old tiles); the caller keeps its padding slots,
// ghost tiles and dummy tile in step and swaps the two arrays afterwards.  cl_f is not written.
template <class real> struct CpFused {
    real *cl_v, *cl_xn;
    real dtforce, dt;
    // single domain, small systems: the ghost tiles' lanes of the atom (updatePbcCPU, pbc.c:45-114) are written by the same
    // epilogue -- image b of tile t is ghost tile ncj + goff[t] + rank of bit b in gmask[t] (setupPbc's order, k_cp_ghost_fill)
    // -- so that a step between two rebuilds is ONE launch.  gmask == nullptr: somebody else updates the ghosts.
    const unsigned* gmask;
    const int* goff;
    int ncj;
    real xprd, yprd, zprd;
};
template <class real>
__device__ __forceinline__ real cp_fused_integrate(const CpFused<real>& fi, size_t s, real xold, real f)
{
    real v = fi.cl_v[s] + fi.dtforce * f;
    v      = v + fi.dtforce * f;
    fi.cl_v[s]  = v;
    const real xn = xold + fi.dt * v;
    fi.cl_xn[s]   = xn;
    return xn;
}
// one atom: both integrate halves, then its lane in the ghost images of its tile (the same single fma as k_cp_update_pbc).
// e = slot of the atom's x in the cluster arrays (tile * 3N + lane)
template <class real, int N>
__device__ __forceinline__ void cp_fused_atom(const CpFused<real>& fi, size_t e, real xo, real yo, real zo, real fx, real fy, real fz)
{
    const real xn = cp_fused_integrate(fi, e, xo, fx), yn = cp_fused_integrate(fi, e + N, yo, fy), zn = cp_fused_integrate(fi, e + 2 * N, zo, fz);
    if (fi.gmask) {
        const size_t t = e / (3 * N);
        unsigned m = fi.gmask[t];
        if (m) {
            const int q = (int)(e - t * 3 * N);
            size_t d = (size_t)(fi.ncj + fi.goff[t]) * 3 * N + q;
            do {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                fi.cl_xn[d]         = fma_rn((real)c_img[b][0], fi.xprd, xn);
                fi.cl_xn[d + N]     = fma_rn((real)c_img[b][1], fi.yprd, yn);
                fi.cl_xn[d + 2 * N] = fma_rn((real)c_img[b][2], fi.zprd, zn);
                d += 3 * N;
            } while (m);
        }
    }
}
// slot of this lane's i atom, re-derived from the special registers after the pair loop: nothing of the epilogue
// (addresses, the loads of v and x) can then be hoisted above the loop, where it would cost registers.  Blocks of 128.
template <int N> __device__ __forceinline__ size_t cp_epilogue_slot()
{
    unsigned t, b;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t));
    asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(b));
    const unsigned tid = b * 128u + t;
    return cp_ci_base3<N>((int)(tid >> 2)) + (tid & 3u);
}
template <int N, bool FI = false>
__global__ void __launch_bounds__(128) k_cp_force_lj_sp_packed(int ncl, int dummy_cj, LJConst2<float> c, const float* __restrict__ cl_x,
    const int* __restrict__ numneigh, const int* __restrict__ neighbors, int maxneighs, float* __restrict__ cl_f, CpFused<float> fi)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int ci  = tid >> 2, cii = tid & 3;
    const bool valid = ci < ncl;
    const int cic = valid ? ci : ncl - 1;
    const size_t ib = cp_ci_base3<N>(cic);
    float x0 = cl_x[ib + cii], y0 = cl_x[ib + N + cii], z0 = cl_x[ib + 2 * N + cii];
    const bool pad_i = x0 >= 1.0e14f;
    if (pad_i) x0 = y0 = z0 = -1.0e15f;
    const f32x2 xt = pk2(x0, x0), yt = pk2(y0, y0), zt = pk2(z0, z0);
    CpPackedConst pc { pk2(c.A, c.A), pk2(-c.B, -c.B), pk2(1.0f, 1.0f), c.cutforcesq };
    const int self = cp_cj0<N>(cic);
    const int ii   = N == CP_M ? cii : cii + CP_M * (cic & 1);
    const int* row = neighbors + (size_t)cic * maxneighs;
    const int nn   = valid ? numneigh[cic] : 0;
    f32x2 fx = pk2(0.f, 0.f), fy = fx, fz = fx;
    // software pipeline: list entry two tiles ahead, tile positions one tile ahead (the dummy tile past the end)
    int cj = nn > 0 ? __ldg(row) : dummy_cj, cj1 = nn > 1 ? __ldg(row + 1) : dummy_cj;
    CpTileRegs<N> A, B;
    A.load(cl_x + (size_t)cj * N * 3);
    for (int k = 0; k < nn; k++) {
        const int cj2 = k + 2 < nn ? __ldg(row + k + 2) : dummy_cj;
        B.load(cl_x + (size_t)cj1 * N * 3);
        cp_tile_packed<N>(A, pc, xt, yt, zt, cj == self ? ii : -1, fx, fy, fz);
        A   = B;
        cj  = cj1;
        cj1 = cj2;
    }
    if (!valid) return;
    float a, b;
    upk2(fx, a, b); const float fix = pad_i ? 0.f : a + b;
    upk2(fy, a, b); const float fiy = pad_i ? 0.f : a + b;
    upk2(fz, a, b); const float fiz = pad_i ? 0.f : a + b;
    if (FI) {
        if (pad_i) return;
        // own coordinates re-read (a real atom's x0/y0/z0 are unmodified, but keeping them live costs registers)
        const size_t e = cp_epilogue_slot<N>();
        cp_fused_atom<float, N>(fi, e, cl_x[e], cl_x[e + N], cl_x[e + 2 * N], fix, fiy, fiz);
        return;
    }
    cl_f[ib + cii] = fix; cl_f[ib + N + cii] = fiy; cl_f[ib + 2 * N + cii] = fiz;
}

// ---- SP, full lists, generation 3: TWO lanes per i-cluster, two i atoms per lane ------------------------------------------
// ncu of the lane-per-i-atom kernel above: FMA pipe and L1 data stage both near 60-70 %, half of the stall samples waiting
// for tile loads.  Each lane fetched a whole j tile (3 x 128 bit) to evaluate 4 pairs; here a lane keeps TWO i atoms of
// its cluster and evaluates 8 pairs per fetched tile, so a warp covers 16 i-clusters and the tile loads (L1 wavefronts,
// issue slots, address arithmetic, list-entry loads) per evaluated pair halve; the packed FP32 pair arithmetic is
// unchanged (j atoms (q, q+1) share an instruction) and there are four independent dependency chains per tile instead of two.
// NR = Newton step on the MUFU reciprocal (as above); without it the reciprocal is the MUFU result itself (max. 1 ulp).
template <int N, bool NR>
__device__ __forceinline__ void cp_tile_packed2(const CpTileRegs<N>& T, const CpPackedConst& c, f32x2 xt, f32x2 yt, f32x2 zt, int iq,
    f32x2& fx, f32x2& fy, f32x2& fz)
{
#pragma unroll
    for (int q = 0; q < N / 2; q++) {
        const f32x2 dx = sub2(xt, T.x[q]), dy = sub2(yt, T.y[q]), dz = sub2(zt, T.z[q]);
        const f32x2 rsq = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
        float r0, r1, y0, y1;
        upk2(rsq, r0, r1);
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(r0));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y1) : "f"(r1));
        const bool in0 = r0 < c.cutforcesq && iq != 2 * q, in1 = r1 < c.cutforcesq && iq != 2 * q + 1; // force_lj.c:99-113
        f32x2 y = pk2(y0, y1);
        if (NR) {
            const f32x2 nr = sub2(c.one, mul2(rsq, y));
            y              = fma2(y, nr, y);
        }
        const f32x2 s3 = mul2(mul2(y, y), y);
        f32x2 f        = mul2(mul2(y, s3), fma2(c.A, s3, c.negB));
        float f0, f1;
        upk2(f, f0, f1);
        f  = pk2(in0 ? f0 : 0.0f, in1 ? f1 : 0.0f);
        fx = fma2(dx, f, fx); fy = fma2(dy, f, fy); fz = fma2(dz, f, fz);
    }
}
template <int N, bool FI, bool NR>
__global__ void __launch_bounds__(128) k_cp_force_lj_sp_duo(int ncl, int dummy_cj, LJConst2<float> c, const float* __restrict__ cl_x,
    const int* __restrict__ numneigh, const int* __restrict__ neighbors, int maxneighs, float* __restrict__ cl_f, CpFused<float> fi)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int ci  = tid >> 1, h = tid & 1; // i atoms 2h and 2h + 1 of cluster ci
    const bool valid = ci < ncl;
    const int cic = valid ? ci : ncl - 1;
    const size_t ib = cp_ci_base3<N>(cic) + 2 * h;
    const float2 px = *reinterpret_cast<const float2*>(cl_x + ib), py = *reinterpret_cast<const float2*>(cl_x + ib + N),
                 pz = *reinterpret_cast<const float2*>(cl_x + ib + 2 * N);
    const bool pad0 = px.x >= 1.0e14f, pad1 = px.y >= 1.0e14f;
    const float ax = pad0 ? -1.0e15f : px.x, ay = pad0 ? -1.0e15f : py.x, az = pad0 ? -1.0e15f : pz.x;
    const float bx = pad1 ? -1.0e15f : px.y, by = pad1 ? -1.0e15f : py.y, bz = pad1 ? -1.0e15f : pz.y;
    const f32x2 xa = pk2(ax, ax), ya = pk2(ay, ay), za = pk2(az, az), xb = pk2(bx, bx), yb = pk2(by, by), zb = pk2(bz, bz);
    CpPackedConst pc { pk2(c.A, c.A), pk2(-c.B, -c.B), pk2(1.0f, 1.0f), c.cutforcesq };
    const int self = cp_cj0<N>(cic);
    const int ia   = (N == CP_M ? 0 : CP_M * (cic & 1)) + 2 * h; // slot of the first i atom inside the diagonal tile
    const int* row = neighbors + (size_t)cic * maxneighs;
    const int nn   = valid ? numneigh[cic] : 0;
    f32x2 fxa = pk2(0.f, 0.f), fya = fxa, fza = fxa, fxb = fxa, fyb = fxa, fzb = fxa;
    int cj = nn > 0 ? __ldg(row) : dummy_cj, cj1 = nn > 1 ? __ldg(row + 1) : dummy_cj;
    CpTileRegs<N> A, B;
    A.load(cl_x + (size_t)cj * N * 3);
    for (int k = 0; k < nn; k++) { // list entry two tiles ahead, tile positions one tile ahead (the dummy tile past the end)
        const int cj2 = k + 2 < nn ? __ldg(row + k + 2) : dummy_cj;
        B.load(cl_x + (size_t)cj1 * N * 3);
        const bool diag = cj == self;
        cp_tile_packed2<N, NR>(A, pc, xa, ya, za, diag ? ia : -1, fxa, fya, fza);
        cp_tile_packed2<N, NR>(A, pc, xb, yb, zb, diag ? ia + 1 : -1, fxb, fyb, fzb);
        A   = B;
        cj  = cj1;
        cj1 = cj2;
    }
    if (!valid) return;
    float u, v;
    upk2(fxa, u, v); const float f0x = pad0 ? 0.f : u + v;
    upk2(fya, u, v); const float f0y = pad0 ? 0.f : u + v;
    upk2(fza, u, v); const float f0z = pad0 ? 0.f : u + v;
    upk2(fxb, u, v); const float f1x = pad1 ? 0.f : u + v;
    upk2(fyb, u, v); const float f1y = pad1 ? 0.f : u + v;
    upk2(fzb, u, v); const float f1z = pad1 ? 0.f : u + v;
    if (FI) {
        // slot re-derived from the special registers (see cp_epilogue_slot): blocks of 128 threads = 64 i-clusters
        unsigned t, b;
        asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t));
        asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(b));
        const unsigned g = b * 128u + t;
        const size_t e   = cp_ci_base3<N>((int)(g >> 1)) + 2 * (g & 1u);
        if (!pad0) cp_fused_atom<float, N>(fi, e, cl_x[e], cl_x[e + N], cl_x[e + 2 * N], f0x, f0y, f0z);
        if (!pad1) cp_fused_atom<float, N>(fi, e + 1, cl_x[e + 1], cl_x[e + N + 1], cl_x[e + 2 * N + 1], f1x, f1y, f1z);
        return;
    }
    *reinterpret_cast<float2*>(cl_f + ib)         = make_float2(f0x, f1x);
    *reinterpret_cast<float2*>(cl_f + ib + N)     = make_float2(f0y, f1y);
    *reinterpret_cast<float2*>(cl_f + ib + 2 * N) = make_float2(f0z, f1z);
}

template <class real, int N, bool HALF, bool FI = false>
__global__ void __launch_bounds__(128) k_cp_force_lj(int ncl, int ncj, LJConst2<real> c, const real* __restrict__ cl_x,
    const int* __restrict__ numneigh, const int* __restrict__ numneigh_masked, const int* __restrict__ neighbors, int maxneighs,
    real* __restrict__ cl_f, CpFused<real> fi)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int ci  = tid >> 2, cii = tid & 3;
    const bool valid = ci < ncl;
    const int cic = valid ? ci : ncl - 1;
    const size_t ib = cp_ci_base3<N>(cic);
    real xt = cl_x[ib + cii], yt = cl_x[ib + N + cii], zt = cl_x[ib + 2 * N + cii];
    const bool pad_i = xt >= CP_PAD_MIN;
    if (pad_i) xt = yt = zt = -CP_PAD; // far from every real atom AND from the +CP_PAD padding lanes of the j tiles
    const int self = cp_cj0<N>(cic);
    const int ii   = N == CP_M ? cii : cii + CP_M * (cic & 1); // own lane inside the diagonal tile
    const int nn   = valid ? numneigh[cic] : 0;
    const int nm   = valid ? numneigh_masked[cic] : 0; // the list build keeps the diagonal entries in front (neighbor.c:374-385)
    const int* row = neighbors + (size_t)cic * maxneighs;
    real fix = 0, fiy = 0, fiz = 0;
    int cjn = nn > 0 ? __ldg(row) : 0;
    int k   = 0;
    for (; k < nm; k++) {
        const int cj = cjn;
        if (k + 1 < nn) cjn = __ldg(row + k + 1);
        const real* t = cl_x + (size_t)cj * N * 3;
        real* fj      = cl_f + (size_t)cj * N * 3;
        // the reference also subtracts from ghost tiles (its HALF_NEIGHBOR_LISTS_CHECK_CJ guard is ineffective, SURVEY 8a
        // a16) but never reads them back; skipping cj >= ncj leaves every local force unchanged.
        if (cj == self) cp_tile<real, N, HALF, true>(t, c, xt, yt, zt, ii, cii, cj < ncj, fj, fix, fiy, fiz);
        else cp_tile<real, N, HALF, false>(t, c, xt, yt, zt, ii, cii, cj < ncj, fj, fix, fiy, fiz);
    }
    for (; k < nn; k++) {
        const int cj = cjn;
        if (k + 1 < nn) cjn = __ldg(row + k + 1);
        cp_tile<real, N, HALF, false>(cl_x + (size_t)cj * N * 3, c, xt, yt, zt, ii, cii, cj < ncj, cl_f + (size_t)cj * N * 3, fix,
            fiy, fiz);
    }
    if (!valid) return;
    if (FI && !HALF) {
        if (pad_i) return; // pad_i <=> cii >= inat[ci]: padding slots are not integrated (integrate.c:27)
        // xt/yt/zt of a real atom are its unmodified coordinates
        const size_t e = cp_epilogue_slot<N>();
        cp_fused_atom<real, N>(fi, e, xt, yt, zt, fix, fiy, fiz);
        return;
    }
    if (pad_i) fix = fiy = fiz = 0;
    if (HALF) {
        if (!pad_i) { atomicAdd(cl_f + ib + cii, fix); atomicAdd(cl_f + ib + N + cii, fiy); atomicAdd(cl_f + ib + 2 * N + cii, fiz); }
    } else {
        cl_f[ib + cii] = fix; cl_f[ib + N + cii] = fiy; cl_f[ib + 2 * N + cii] = fiz;
    }
}

// ---- force, generation 2: one WARP per i-cluster, one LANE per j atom ---------------------------------------------------
// ncu of the lane-per-i-atom kernels above: every j tile costs ~3 L1 wavefronts per cluster pair (three 128-bit row loads,
// eight different tiles per warp instruction), as many cycles as the arithmetic of its 16 atom pairs -- and with packed
// FP32 math halving the issue slots the kernel simply became L1-bound (0.211 vs 0.220 ms, profiles/r1_ab3.txt).
// Here a warp walks ONE list row, 32 / N tiles per iteration: lane (t, q) fetches atom q of tile t with a single 128-bit
// (SP) / 256-bit (DP) load from an {x, y, z, -} copy of the cluster positions (k_cp_pack_j) -- one wavefront per cluster
// pair -- and evaluates it against the four i atoms, which are warp-uniform registers.  The 12 i-force sums are
// reduced over the warp once per row with a halving butterfly (12 -> 6 -> 3 values, then 3 full steps).
// one 32-byte {x, y, z, -} record per cluster slot (DP; 16 bytes for SP), fetched with a single 256- / 128-bit load
struct __align__(32) PosD {
    double x, y, z, w;
};
template <class real> struct PosOf;
template <> struct PosOf<double> { typedef PosD type; };
template <> struct PosOf<float> { typedef float4 type; };
__device__ __forceinline__ void ld_pos(const PosD* p, double& x, double& y, double& z)
{
    double w;
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x), "=d"(y), "=d"(z), "=d"(w) : "l"(p));
}
__device__ __forceinline__ void ld_pos(const float4* p, float& x, float& y, float& z)
{
    const float4 v = __ldg(p);
    x = v.x; y = v.y; z = v.z;
}
__device__ __forceinline__ void st_pos(PosD* p, double x, double y, double z)
{
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(x), "d"(y), "d"(z), "d"(0.0) : "memory");
}
__device__ __forceinline__ void st_pos(float4* p, float x, float y, float z) { *p = make_float4(x, y, z, 0.f); }

template <class real, int N>
__global__ void k_cp_pack_j(size_t nslots, const real* __restrict__ cl_x, typename PosOf<real>::type* __restrict__ out)
{
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslots) return;
    const size_t t = s / N, q = s % N;
    const real* p  = cl_x + t * N * 3 + q;
    st_pos(out + s, p[0], p[N], p[2 * N]);
}

template <class real> struct CpPairMath;
template <> struct CpPairMath<double> { // scalar FP64
    double xi[4], yi[4], zi[4], fx[4], fy[4], fz[4];
    double cut, A, B;
    __device__ __forceinline__ void init(const double* px, const double* py, const double* pz, const LJConst2<double>& c)
    {
#pragma unroll
        for (int m = 0; m < 4; m++) { xi[m] = px[m]; yi[m] = py[m]; zi[m] = pz[m]; fx[m] = fy[m] = fz[m] = 0; }
        cut = c.cutforcesq; A = c.A; B = c.B;
    }
    // pairs (i atom m, this lane's j atom) for m = 0..3; ok[m] = pair allowed (exclusions, list padding);
    // returns the reaction sum for the j atom in (rx, ry, rz)
    template <bool HALF>
    __device__ __forceinline__ void tile(double xj, double yj, double zj, const bool (&ok)[4], double& rx, double& ry, double& rz)
    {
#pragma unroll
        for (int m = 0; m < 4; m++) {
            const double dx = xi[m] - xj, dy = yi[m] - yj, dz = zi[m] - zj;
            const double rsq = dx * dx + dy * dy + dz * dz;
            const double s = rcp_nr(rsq), s3 = s * s * s;
            double f = (s * s3) * (A * s3 - B);
            f        = (rsq < cut && ok[m]) ? f : 0.0;
            fx[m] = fma(dx, f, fx[m]); fy[m] = fma(dy, f, fy[m]); fz[m] = fma(dz, f, fz[m]);
            if (HALF) { rx = fma(dx, f, rx); ry = fma(dy, f, ry); rz = fma(dz, f, rz); }
        }
    }
    __device__ __forceinline__ void sums(double (&v)[12])
    {
#pragma unroll
        for (int m = 0; m < 4; m++) { v[3 * m] = fx[m]; v[3 * m + 1] = fy[m]; v[3 * m + 2] = fz[m]; }
    }
};
template <> struct CpPairMath<float> { // packed FP32: i atoms (0,1) and (2,3) share an instruction
    f32x2 xi[2], yi[2], zi[2], fx[2], fy[2], fz[2], A, negB, one;
    float cut;
    __device__ __forceinline__ void init(const float* px, const float* py, const float* pz, const LJConst2<float>& c)
    {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            xi[h] = pk2(px[2 * h], px[2 * h + 1]); yi[h] = pk2(py[2 * h], py[2 * h + 1]); zi[h] = pk2(pz[2 * h], pz[2 * h + 1]);
            fx[h] = fy[h] = fz[h] = pk2(0.f, 0.f);
        }
        A = pk2(c.A, c.A); negB = pk2(-c.B, -c.B); one = pk2(1.f, 1.f);
        cut = c.cutforcesq;
    }
    template <bool HALF>
    __device__ __forceinline__ void tile(float xj, float yj, float zj, const bool (&ok)[4], float& rx, float& ry, float& rz)
    {
        const f32x2 xj2 = pk2(xj, xj), yj2 = pk2(yj, yj), zj2 = pk2(zj, zj);
        f32x2 r2x = pk2(0.f, 0.f), r2y = r2x, r2z = r2x;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const f32x2 dx = sub2(xi[h], xj2), dy = sub2(yi[h], yj2), dz = sub2(zi[h], zj2);
            const f32x2 rsq = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
            float r0, r1, y0, y1;
            upk2(rsq, r0, r1);
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(r0));
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y1) : "f"(r1));
            const bool in0 = r0 < cut && ok[2 * h], in1 = r1 < cut && ok[2 * h + 1];
            f32x2 y        = pk2(y0, y1);
            y              = fma2(y, sub2(one, mul2(rsq, y)), y); // Newton step
            const f32x2 s3 = mul2(mul2(y, y), y);
            f32x2 f        = mul2(mul2(y, s3), fma2(A, s3, negB));
            float f0, f1;
            upk2(f, f0, f1);
            f     = pk2(in0 ? f0 : 0.0f, in1 ? f1 : 0.0f);
            fx[h] = fma2(dx, f, fx[h]); fy[h] = fma2(dy, f, fy[h]); fz[h] = fma2(dz, f, fz[h]);
            if (HALF) { r2x = fma2(dx, f, r2x); r2y = fma2(dy, f, r2y); r2z = fma2(dz, f, r2z); }
        }
        if (HALF) {
            float a, b;
            upk2(r2x, a, b); rx += a + b;
            upk2(r2y, a, b); ry += a + b;
            upk2(r2z, a, b); rz += a + b;
        }
    }
    __device__ __forceinline__ void sums(float (&v)[12])
    {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            upk2(fx[h], v[6 * h], v[6 * h + 3]);
            upk2(fy[h], v[6 * h + 1], v[6 * h + 4]);
            upk2(fz[h], v[6 * h + 2], v[6 * h + 5]);
        }
    }
};

template <class real, int N, bool HALF>
__global__ void __launch_bounds__(128) k_cp_force_jl(int ncl, int ncj, int dummy_cj, LJConst2<real> c, const real* __restrict__ cl_x,
    const typename PosOf<real>::type* __restrict__ pos, const int* __restrict__ numneigh, const int* __restrict__ neighbors,
    int maxneighs, real* __restrict__ cl_f)
{
    constexpr int TPW = 32 / N; // tiles per warp iteration
    const int ci = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (ci >= ncl) return; // whole warp
    const int lane = threadIdx.x & 31, t = lane / N, q = lane % N;
    const size_t ib = cp_ci_base3<N>(ci);
    real px[4], py[4], pz[4];
#pragma unroll
    for (int m = 0; m < 4; m++) {
        px[m] = cl_x[ib + m]; py[m] = cl_x[ib + N + m]; pz[m] = cl_x[ib + 2 * N + m];
        if (px[m] >= CP_PAD_MIN) px[m] = py[m] = pz[m] = -CP_PAD; // padding i lane: far from everything, j padding included
    }
    CpPairMath<real> M;
    M.init(px, py, pz, c);
    const int self = cp_cj0<N>(ci);
    const int ioff = N == CP_M ? 0 : CP_M * (ci & 1); // lane of i atom 0 inside the diagonal tile
    const int* row = neighbors + (size_t)ci * maxneighs;
    const int nn   = numneigh[ci];
    // software pipeline: list entries two iterations ahead, positions one iteration ahead
    // rows are maxneighs (>= 2 TPW) wide: the first entries are fetched without waiting for the row length
    int cj_next = __ldg(row + t), cj_nn = __ldg(row + TPW + t);
    if (t >= nn) cj_next = dummy_cj;
    if (TPW + t >= nn) cj_nn = dummy_cj;
    real xj, yj, zj;
    ld_pos(pos + (size_t)cj_next * N + q, xj, yj, zj);
    for (int k0 = 0; k0 < nn; k0 += TPW) {
        const int cj = cj_next;
        const real x = xj, y = yj, z = zj;
        cj_next = cj_nn;
        cj_nn   = k0 + 2 * TPW + t < nn ? __ldg(row + k0 + 2 * TPW + t) : dummy_cj;
        ld_pos(pos + (size_t)cj_next * N + q, xj, yj, zj);
        const bool diag = cj == self;
        bool ok[4];
#pragma unroll
        for (int m = 0; m < 4; m++) ok[m] = !diag || (HALF ? (ioff + m < q) : (ioff + m != q)); // force_lj.c:99-113
        real rx = 0, ry = 0, rz = 0;
        M.template tile<HALF>(x, y, z, ok, rx, ry, rz);
        // reaction on local j atoms.  The reference also subtracts from ghost tiles (its HALF_NEIGHBOR_LISTS_CHECK_CJ guard is
        // ineffective) but never reads them back; skipping cj >= ncj leaves every local force unchanged.
        if (HALF && cj < ncj && (rx != 0 || ry != 0 || rz != 0)) {
            real* fj = cl_f + (size_t)cj * N * 3 + q;
            atomicAdd(fj, -rx); atomicAdd(fj + N, -ry); atomicAdd(fj + 2 * N, -rz);
        }
    }
    // warp reduction of v[3 m + comp]: 12 -> 6 (xor 16) -> 3 (xor 8) -> full sums over xor 4, 2, 1
    real v[12];
    M.sums(v);
    const bool h1 = lane & 16, h2 = lane & 8;
    real w[6];
#pragma unroll
    for (int j = 0; j < 6; j++) {
        const real send = h1 ? v[j] : v[j + 6];
        const real keep = h1 ? v[j + 6] : v[j];
        w[j]            = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    real u[3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const real send = h2 ? w[j] : w[j + 3];
        const real keep = h2 ? w[j + 3] : w[j];
        u[j]            = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int d = 4; d > 0; d >>= 1)
#pragma unroll
        for (int j = 0; j < 3; j++) u[j] += __shfl_xor_sync(0xffffffffu, u[j], d);
    if ((lane & 7) == 0) {
        const int m = (h1 ? 2 : 0) + (h2 ? 1 : 0); // i atom whose three components this lane group holds
        if (cl_x[ib + m] < CP_PAD_MIN) { // padding i lanes keep their zero
            real* f = cl_f + ib + m;
            if (HALF) { atomicAdd(f, u[0]); atomicAdd(f + N, u[1]); atomicAdd(f + 2 * N, u[2]); }
            else { f[0] = u[0]; f[N] = u[1]; f[2 * N] = u[2]; }
        } else if (!HALF) { cl_f[ib + m] = 0; cl_f[ib + N + m] = 0; cl_f[ib + 2 * N + m] = 0; }
    }
}

// ---- integrate (clusterpair/integrate.c:23-63): one thread per i-atom slot, padding lanes untouched ------------------
template <class real, int N, int MODE> // MODE 0: initial, 1: final, 2: final(n) + initial(n+1)
__global__ void k_cp_integrate(int ncl, real dtforce, real dt, const int* __restrict__ inat, real* __restrict__ cl_x,
    real* __restrict__ cl_v, const real* __restrict__ cl_f)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int ci  = tid >> 2, cii = tid & 3;
    if (ci >= ncl || cii >= inat[ci]) return;
    const size_t b = cp_ci_base3<N>(ci) + cii;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const size_t s = b + (size_t)a * N;
        const real f = cl_f[s];
        real v = cl_v[s] + dtforce * f;
        if (MODE == 2) v = v + dtforce * f;
        cl_v[s] = v;
        if (MODE != 1) cl_x[s] = cl_x[s] + dt * v;
    }
}
// updateSingleAtoms (neighbor.c:1045-1071): cluster data back to the atom arrays, compacted in cluster order
template <class real, int N>
__global__ void k_cp_update_single_atoms(int ncl, const int* __restrict__ inat, const int* __restrict__ atom_off,
    const real* __restrict__ cl_x, const real* __restrict__ cl_v, const int* __restrict__ cl_tag, real* __restrict__ x,
    real* __restrict__ y, real* __restrict__ z, real* __restrict__ vx, real* __restrict__ vy, real* __restrict__ vz, int* __restrict__ tag)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int ci  = tid >> 2, cii = tid & 3;
    if (ci >= ncl || cii >= inat[ci]) return;
    const size_t b = cp_ci_base3<N>(ci) + cii;
    const int i    = atom_off[ci] + cii;
    x[i] = cl_x[b]; y[i] = cl_x[b + N]; z[i] = cl_x[b + 2 * N];
    vx[i] = cl_v[b]; vy[i] = cl_v[b + N]; vz[i] = cl_v[b + 2 * N];
    tag[i] = cl_tag[cp_ci_base1<N>(ci) + cii];
}
// ---- kernel micro-benchmark (clusterpair/main-stub.c:227-272 synthetic clusters, 61-122 createNeighbors) -----------------
// i-cluster ci holds `nat` atoms at x = y = z = (ci * nat + cii) * 1e-5, the rest of the cluster is padding; tile t is
// j-cluster t (defineJClusters); velocities and forces zero
template <class real, int N>
__global__ void k_cp_stub_clusters(int ncl, int nat, real* __restrict__ cl_x, real* __restrict__ cl_v, real* __restrict__ cl_f,
    int* __restrict__ cl_tag, int* __restrict__ inat, int* __restrict__ jnat, real* __restrict__ ibb, int* __restrict__ ibin)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int ci  = tid >> 2, cii = tid & 3;
    if (ci > ncl) return;
    const size_t b = cp_ci_base3<N>(ci) + cii;
    if (ci == ncl) { // the dummy tile behind the last one
        if (N == CP_M || (ci & 1) == 0)
            for (int q = cii; q < 3 * N; q += CP_M) cl_x[(size_t)cp_cj0<N>(ci) * N * 3 + q] = CP_PAD;
        return;
    }
    const real p = cii < nat ? (real)(ci * nat + cii) * (real)0.00001 : CP_PAD;
    cl_x[b] = p; cl_x[b + N] = p; cl_x[b + 2 * N] = p;
    cl_v[b] = 0; cl_v[b + N] = 0; cl_v[b + 2 * N] = 0;
    cl_f[b] = 0; cl_f[b + N] = 0; cl_f[b + 2 * N] = 0;
    cl_tag[cp_ci_base1<N>(ci) + cii] = cii < nat ? ci * nat + cii : -1;
    if (cii == 0) {
        inat[ci] = nat;
        ibin[ci] = 0;
        const real lo = (real)(ci * nat) * (real)0.00001, hi = (real)(ci * nat + nat - 1) * (real)0.00001;
        for (int d = 0; d < 3; d++) { ibb[(size_t)ci * 6 + 2 * d] = lo; ibb[(size_t)ci * 6 + 2 * d + 1] = hi; }
        if (N == CP_M) jnat[ci] = nat;
        else if ((ci & 1) == 0) jnat[ci >> 1] = 2 * nat;
    }
}
// pattern 0 "seq": j-clusters CJ0(ci), CJ0(ci)+1, ... (mod ncj); 1 "fix": 0 .. nneighs-1; 2 "rand"; replicated nreps times
template <int N>
__global__ void k_cp_stub_neighbors(int ncl, int ncj, int pattern, int nneighs, int nreps, int nmasked, unsigned seed, int maxneighs,
    int* __restrict__ numneigh, int* __restrict__ numneigh_masked, int* __restrict__ neighbors)
{
    const int ci = blockIdx.x * blockDim.x + threadIdx.x;
    if (ci >= ncl) return;
    int* row   = neighbors + (size_t)ci * maxneighs;
    unsigned h = seed ^ (0x9e3779b9u * (unsigned)(ci + 1));
    int j      = pattern == 0 ? cp_cj0<N>(ci) : 0;
    const int m = pattern == 0 ? ncj : nneighs;
    for (int k = 0; k < nneighs; k++) {
        int v;
        if (pattern == 2) { // never the own tile: its self pairs would need the masked loop (the reference does not check)
            do {
                h ^= h << 13; h ^= h >> 17; h ^= h << 5;
                v = (int)(h % (unsigned)ncj);
            } while (v == cp_cj0<N>(ci) && ncj > 1);
        } else {
            v = j;
            j = (j + 1) % m;
        }
        for (int r = 0; r < nreps; r++) row[r * nneighs + k] = v;
    }
    numneigh[ci]        = nneighs * nreps;
    numneigh_masked[ci] = nmasked;
}

template <class real> __global__ void k_cp_zero(size_t n, real* __restrict__ a)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = 0;
}
// pairs for the roofline: listed cluster pairs x M x N and atom pairs inside the cutoff
template <class real, int N>
__global__ void k_cp_count_pairs(int ncl, real cutforcesq, const real* __restrict__ cl_x, const int* __restrict__ numneigh,
    const int* __restrict__ neighbors, int maxneighs, unsigned long long* __restrict__ out)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int ci  = tid >> 2, cii = tid & 3;
    unsigned long long listed = 0, inside = 0;
    if (ci < ncl) {
        const size_t ib = cp_ci_base3<N>(ci);
        real xt = cl_x[ib + cii], yt = cl_x[ib + N + cii], zt = cl_x[ib + 2 * N + cii];
        if (xt >= CP_PAD_MIN) xt = yt = zt = -CP_PAD;
        const int nn = numneigh[ci];
        if (cii == 0) listed = nn;
        for (int k = 0; k < nn; k++) {
            const real* t = cl_x + (size_t)neighbors[(size_t)ci * maxneighs + k] * N * 3;
            for (int q = 0; q < N; q++) {
                const real dx = xt - t[q], dy = yt - t[N + q], dz = zt - t[2 * N + q];
                inside += (dx * dx + dy * dy + dz * dz) < cutforcesq;
            }
        }
    }
    atomicAdd(out, listed);
    atomicAdd(out + 1, inside);
}

} // namespace mdb
