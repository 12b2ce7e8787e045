// eam_kernels.cuh -- EAM: host-side table builder (runs once) and the three device passes.
// Reference: common/eam_utils.c:95-284 (file2array, array2spline, interpolate) and
// verletlist/force_eam.c:19-231 (density + embedding derivative, ghost fp copy, pair force).
#pragma once
#include <vector>

#include "mdb_util.cuh"

namespace mdb {

template <class real> struct EamTables {
    int nr = 0, nrho = 0, nr_tot = 0, nrho_tot = 0;
    real rdr = 0, rdrho = 0;
    bool ready = false;
};

// 4-point Lagrange regrid, common/eam_utils.c:125-140 (used for frho, rhor and zr)
template <class real> static double eam_lagrange(const std::vector<real>& tab, int ntab, double dtab, double r)
{
    const double sixth = 1.0 / 6.0;
    double p = r / dtab + 1.0;
    int k    = (int)(p);
    k        = k < ntab - 2 ? k : ntab - 2;
    k        = k > 2 ? k : 2;
    p -= k;
    p = p < 2.0 ? p : 2.0;
    const double cof1 = -sixth * p * (p - 1.0) * (p - 2.0);
    const double cof2 = 0.5 * (p * p - 1.0) * (p - 2.0);
    const double cof3 = -0.5 * p * (p + 1.0) * (p - 2.0);
    const double cof4 = sixth * p * (p * p - 1.0);
    return cof1 * tab[k - 1] + cof2 * tab[k] + cof3 * tab[k + 1] + cof4 * tab[k + 2];
}

// cubic spline coefficients, 7 per knot: common/eam_utils.c:253-284
template <class real> static void eam_interpolate(int n, real delta, const std::vector<real>& f, std::vector<real>& s)
{
    for (int m = 1; m <= n; m++) s[m * 7 + 6] = f[m];
    s[1 * 7 + 5]       = s[2 * 7 + 6] - s[1 * 7 + 6];
    s[2 * 7 + 5]       = 0.5 * (s[3 * 7 + 6] - s[1 * 7 + 6]);
    s[(n - 1) * 7 + 5] = 0.5 * (s[n * 7 + 6] - s[(n - 2) * 7 + 6]);
    s[n * 7 + 5]       = s[n * 7 + 6] - s[(n - 1) * 7 + 6];
    for (int m = 3; m <= n - 2; m++)
        s[m * 7 + 5] = ((s[(m - 2) * 7 + 6] - s[(m + 2) * 7 + 6]) + 8.0 * (s[(m + 1) * 7 + 6] - s[(m - 1) * 7 + 6])) / 12.0;
    for (int m = 1; m <= n - 1; m++) {
        s[m * 7 + 4] = 3.0 * (s[(m + 1) * 7 + 6] - s[m * 7 + 6]) - 2.0 * s[m * 7 + 5] - s[(m + 1) * 7 + 5];
        s[m * 7 + 3] = s[m * 7 + 5] + s[(m + 1) * 7 + 5] - 2.0 * (s[(m + 1) * 7 + 6] - s[m * 7 + 6]);
    }
    s[n * 7 + 4] = 0.0;
    s[n * 7 + 3] = 0.0;
    for (int m = 1; m <= n; m++) {
        s[m * 7 + 2] = s[m * 7 + 5] / delta;
        s[m * 7 + 1] = 2.0 * s[m * 7 + 4] / delta;
        s[m * 7 + 0] = 3.0 * s[m * 7 + 3] / delta;
    }
}

// funcfl tables (0-based, as read from the file) -> spline tables.  One element type only
// (ntypes == 1), as in the reference's funcfl path.
template <class real>
static void build_eam_tables(int nrho, real fdrho, int nr, real fdr, const double* frho0, const double* zr0,
    const double* rhor0, EamTables<real>& t, std::vector<real>& rhor_spline, std::vector<real>& frho_spline,
    std::vector<real>& z2r_spline)
{
    // readEamFile shifts the tables to 1-based, eam_utils.c:85-90
    std::vector<real> frho(nrho + 1, 0), rhor(nr + 1, 0), zr(nr + 1, 0);
    for (int i = nrho; i > 0; i--) frho[i] = (real)frho0[i - 1];
    for (int i = nr; i > 0; i--) { rhor[i] = (real)rhor0[i - 1]; zr[i] = (real)zr0[i - 1]; }
    const real edr = fdr, edrho = fdrho; // MAX(0, file value), eam_utils.c:107-108
    const double rmax = (nr - 1) * fdr, rhomax = (nrho - 1) * fdrho;
    const int enr = (int)(rmax / edr + 0.5), enrho = (int)(rhomax / edrho + 0.5);
    std::vector<real> afrho(enrho + 1, 0), arhor(enr + 1, 0), az2r(enr + 1, 0);
    for (int m = 1; m <= enrho; m++) afrho[m] = (real)eam_lagrange(frho, nrho, fdrho, (m - 1) * edrho);
    for (int m = 1; m <= enr; m++) arhor[m] = (real)eam_lagrange(rhor, nr, fdr, (m - 1) * edr);
    for (int m = 1; m <= enr; m++) {
        const double r = (m - 1) * edr;
        const double zri = eam_lagrange(zr, nr, fdr, r), zrj = eam_lagrange(zr, nr, fdr, r);
        az2r[m] = (real)(27.2 * 0.529 * zri * zrj); // eam_utils.c:218
    }
    t.rdr = (real)(1.0 / edr);
    t.rdrho = (real)(1.0 / edrho);
    t.nr = enr;
    t.nrho = enrho;
    t.nrho_tot = (enrho + 1) * 7 + 64;
    t.nr_tot = (enr + 1) * 7 + 64;
    t.nrho_tot -= t.nrho_tot % 64; // eam_utils.c:226-229
    t.nr_tot -= t.nr_tot % 64;
    frho_spline.assign(t.nrho_tot, 0);
    rhor_spline.assign(t.nr_tot, 0);
    z2r_spline.assign(t.nr_tot, 0);
    eam_interpolate(enrho, edrho, afrho, frho_spline);
    eam_interpolate(enr, edr, arhor, rhor_spline);
    eam_interpolate(enr, edr, az2r, z2r_spline);
}

// ---------------------------------------------------------------------------------------------
// pass 1: rho_i and the embedding derivative fp[i]; force_eam.c:49-112
template <class real>
__global__ void __launch_bounds__(128) k_eam_density(int nlocal, real cutforcesq, EamTables<real> t,
    const real* __restrict__ rhor_spline, const real* __restrict__ frho_spline,
    const real* __restrict__ x, const real* __restrict__ y, const real* __restrict__ z,
    const int* __restrict__ numneigh, const int* __restrict__ nbT, NbLayout L, real* __restrict__ fp)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nlocal) return;
    const real xt = x[i], yt = y[i], zt = z[i];
    const int nn  = numneigh[i];
    real rhoi     = 0;
    for (int k = 0; k < nn; k++) {
        const int j   = __ldg(nbT + L.base(i) + (size_t)k * L.sk);
        const real dx = xt - x[j], dy = yt - y[j], dz = zt - z[j];
        const real rsq = dx * dx + dy * dy + dz * dz;
        if (rsq < cutforcesq) {
            real p = sqrt(rsq) * t.rdr + (real)1.0;
            int m  = (int)(p);
            m      = m < t.nr - 1 ? m : t.nr - 1;
            p -= m;
            p = p < (real)1.0 ? p : (real)1.0;
            const real* s = rhor_spline + m * 7;
            rhoi += ((__ldg(s + 3) * p + __ldg(s + 4)) * p + __ldg(s + 5)) * p + __ldg(s + 6);
        }
    }
    real p = (real)1.0 * rhoi * t.rdrho + (real)1.0;
    int m  = (int)(p);
    m      = max(1, min(m, t.nrho - 1));
    p -= m;
    p = min(p, (real)1.0);
    const real* s = frho_spline + m * 7;
    fp[i]         = (__ldg(s + 0) * p + __ldg(s + 1)) * p + __ldg(s + 2);
}

// fp of ghosts = fp of their source atom; force_eam.c:118-120 (the second halo exchange when decomposed)
template <class real>
__global__ void k_eam_ghost_fp(int nlocal, int nghost, const int* __restrict__ border_map, real* __restrict__ fp)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < nghost) fp[nlocal + g] = fp[border_map[g]];
}

// pass 2: pair force; force_eam.c:127-224
template <class real>
__global__ void __launch_bounds__(128) k_eam_force(int nlocal, real cutforcesq, EamTables<real> t,
    const real* __restrict__ rhor_spline, const real* __restrict__ z2r_spline, const real* __restrict__ x,
    const real* __restrict__ y, const real* __restrict__ z, const real* __restrict__ fp,
    const int* __restrict__ numneigh, const int* __restrict__ nbT, NbLayout L, real* __restrict__ fx,
    real* __restrict__ fy, real* __restrict__ fz)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nlocal) return;
    const real xt = x[i], yt = y[i], zt = z[i], fpi = fp[i];
    const int nn = numneigh[i];
    real fix = 0, fiy = 0, fiz = 0;
    for (int k = 0; k < nn; k++) {
        const int j   = __ldg(nbT + L.base(i) + (size_t)k * L.sk);
        const real dx = xt - x[j], dy = yt - y[j], dz = zt - z[j];
        const real rsq = dx * dx + dy * dy + dz * dz;
        if (rsq < cutforcesq) {
            const real r = sqrt(rsq);
            real p       = r * t.rdr + (real)1.0;
            int m        = (int)(p);
            m            = m < t.nr - 1 ? m : t.nr - 1;
            p -= m;
            p = p < (real)1.0 ? p : (real)1.0;
            const real* rs = rhor_spline + m * 7;
            const real* zs = z2r_spline + m * 7;
            const real rhoip = (__ldg(rs + 0) * p + __ldg(rs + 1)) * p + __ldg(rs + 2);
            const real z2p   = (__ldg(zs + 0) * p + __ldg(zs + 1)) * p + __ldg(zs + 2);
            const real z2    = ((__ldg(zs + 3) * p + __ldg(zs + 4)) * p + __ldg(zs + 5)) * p + __ldg(zs + 6);
            const real recip = (real)1.0 / r;
            const real phi   = z2 * recip;
            const real phip  = z2p * recip - phi * recip;
            const real psip  = fpi * rhoip + fp[j] * rhoip + phip;
            const real fpair = -psip * recip;
            fix += dx * fpair;
            fiy += dy * fpair;
            fiz += dz * fpair;
        }
    }
    fx[i] = fix;
    fy[i] = fiy;
    fz[i] = fiz;
}

// ---- generation 2 of the EAM passes ----------------------------------------------------------------------------------
// The first kernels above spend their time in 4 (density) / 10 (force) scalar table gathers per pair, an IEEE square root
// and a division.  Here the spline coefficients a pair needs are repacked into rows of 4 / 12 reals (k_eam_pack_tables)
// and fetched with one / three vector loads; r and 1/r come from rsqrt.approx + Newton steps; two neighbors are in flight.
// The arithmetic on the coefficients is unchanged (same Horner forms as force_eam.c:86-88, 170-180).
template <class real>
__global__ void k_eam_pack_tables(int rows, const real* __restrict__ rhor_spline, const real* __restrict__ z2r_spline,
    real* __restrict__ rho4, real* __restrict__ frc12)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= rows) return;
    const real *rs = rhor_spline + (size_t)m * 7, *zs = z2r_spline + (size_t)m * 7;
    real* a = rho4 + (size_t)m * 4;
    a[0] = rs[3]; a[1] = rs[4]; a[2] = rs[5]; a[3] = rs[6];
    real* b = frc12 + (size_t)m * 12;
    b[0] = rs[0]; b[1] = rs[1]; b[2] = rs[2]; b[3] = zs[0];
    b[4] = zs[1]; b[5] = zs[2]; b[6] = zs[3]; b[7] = zs[4];
    b[8] = zs[5]; b[9] = zs[6]; b[10] = 0; b[11] = 0;
}
__device__ __forceinline__ void ld4(const double* p, double& a, double& b, double& c, double& d)
{
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__device__ __forceinline__ void ld4(const float* p, float& a, float& b, float& c, float& d)
{
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    a = v.x; b = v.y; c = v.z; d = v.w;
}
// 1 / sqrt(a) to working precision: approximation + Newton steps y <- y (1.5 - 0.5 a y^2)
__device__ __forceinline__ double rsqrt_nr(double a)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    const double h = 0.5 * a;
    y = y * fma(-h * y, y, 1.5);
    y = y * fma(-h * y, y, 1.5);
    y = y * fma(-h * y, y, 1.5);
    return y;
}
__device__ __forceinline__ float rsqrt_nr(float a)
{
    float y = rsqrtf(a);
    return y * fmaf(-0.5f * a * y, y, 1.5f);
}

template <class real, int U>
__global__ void __launch_bounds__(128) k_eam_density_v2(int nlocal, real cutforcesq, EamTables<real> t, const real* __restrict__ rho4,
    const real* __restrict__ frho_spline, const real* __restrict__ x, const real* __restrict__ y, const real* __restrict__ z,
    const int* __restrict__ numneigh, const int* __restrict__ nbT, NbLayout L, real* __restrict__ fp)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nlocal) return;
    const real xt = x[i], yt = y[i], zt = z[i];
    const int nn  = numneigh[i];
    const int* nb = nbT + L.base(i);
    real rhoi     = 0;
    for (int k = 0; k < nn; k += U) {
        int j[U];
        real rsq[U];
#pragma unroll
        for (int u = 0; u < U; u++) j[u] = k + u < nn ? __ldg(nb + (size_t)(k + u) * L.sk) : i;
#pragma unroll
        for (int u = 0; u < U; u++) {
            const real dx = xt - __ldg(x + j[u]), dy = yt - __ldg(y + j[u]), dz = zt - __ldg(z + j[u]);
            rsq[u] = dx * dx + dy * dy + dz * dz;
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (rsq[u] < cutforcesq && k + u < nn) {
                const real r = rsq[u] * rsqrt_nr(rsq[u]);
                real p = r * t.rdr + (real)1.0;
                int m  = (int)(p);
                m      = m < t.nr - 1 ? m : t.nr - 1;
                p -= m;
                p = p < (real)1.0 ? p : (real)1.0;
                real s3, s4, s5, s6;
                ld4(rho4 + (size_t)m * 4, s3, s4, s5, s6);
                rhoi += ((s3 * p + s4) * p + s5) * p + s6;
            }
        }
    }
    real p = (real)1.0 * rhoi * t.rdrho + (real)1.0;
    int m  = (int)(p);
    m      = max(1, min(m, t.nrho - 1));
    p -= m;
    p = min(p, (real)1.0);
    const real* s = frho_spline + m * 7;
    fp[i]         = (__ldg(s + 0) * p + __ldg(s + 1)) * p + __ldg(s + 2);
}

template <class real, int U>
__global__ void __launch_bounds__(128) k_eam_force_v2(int nlocal, real cutforcesq, EamTables<real> t, const real* __restrict__ frc12,
    const real* __restrict__ x, const real* __restrict__ y, const real* __restrict__ z, const real* __restrict__ fp,
    const int* __restrict__ numneigh, const int* __restrict__ nbT, NbLayout L, real* __restrict__ fx, real* __restrict__ fy,
    real* __restrict__ fz)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nlocal) return;
    const real xt = x[i], yt = y[i], zt = z[i], fpi = fp[i];
    const int nn  = numneigh[i];
    const int* nb = nbT + L.base(i);
    real fix = 0, fiy = 0, fiz = 0;
    for (int k = 0; k < nn; k += U) {
        int j[U];
        real dx[U], dy[U], dz[U], rsq[U], fpj[U];
#pragma unroll
        for (int u = 0; u < U; u++) j[u] = k + u < nn ? __ldg(nb + (size_t)(k + u) * L.sk) : i;
#pragma unroll
        for (int u = 0; u < U; u++) {
            dx[u] = xt - __ldg(x + j[u]); dy[u] = yt - __ldg(y + j[u]); dz[u] = zt - __ldg(z + j[u]);
            fpj[u] = __ldg(fp + j[u]);
            rsq[u] = dx[u] * dx[u] + dy[u] * dy[u] + dz[u] * dz[u];
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (rsq[u] < cutforcesq && k + u < nn) {
                const real recip = rsqrt_nr(rsq[u]);
                const real r     = rsq[u] * recip;
                real p           = r * t.rdr + (real)1.0;
                int m            = (int)(p);
                m                = m < t.nr - 1 ? m : t.nr - 1;
                p -= m;
                p = p < (real)1.0 ? p : (real)1.0;
                const real* row = frc12 + (size_t)m * 12;
                real rs0, rs1, rs2, zs0, zs1, zs2, zs3, zs4, zs5, zs6, pad0, pad1;
                ld4(row, rs0, rs1, rs2, zs0);
                ld4(row + 4, zs1, zs2, zs3, zs4);
                ld4(row + 8, zs5, zs6, pad0, pad1);
                const real rhoip = (rs0 * p + rs1) * p + rs2;
                const real z2p   = (zs0 * p + zs1) * p + zs2;
                const real z2    = ((zs3 * p + zs4) * p + zs5) * p + zs6;
                const real phi   = z2 * recip;
                const real phip  = z2p * recip - phi * recip;
                const real psip  = fpi * rhoip + fpj[u] * rhoip + phip;
                const real fpair = -psip * recip;
                fix += dx[u] * fpair;
                fiy += dy[u] * fpair;
                fiz += dz[u] * fpair;
            }
        }
    }
    fx[i] = fix;
    fy[i] = fiy;
    fz[i] = fiz;
}

// ---- generation 3 of the EAM passes (option eam_variant = 2; A/B, written after round 1's GPU budget was spent) -----------
// ncu of generation 2: both passes sit at 89 % of the L1 data pipe; per listed pair the force pass issues 1 index load,
// 4 scalar gathers (x, y, z, fp) and 3 vector loads of its own table row (96 B DP).  Here
//  * x, y come from the packed (x, y) copy (k_pack_xy) and z, fp from a packed (z, fp) array that the density pass and
//    the ghost-fp kernel write: 2 vector gathers per pair instead of 4 scalar ones (3 -> 2 in the density pass);
//  * the force pass reads (value, slope) of rhor and z2r at the knots m and m+1 -- two adjacent 4-real rows -- and
//    derives the cubic's coefficients in registers exactly as interpolate() does (eam_utils.c:269-274:
//    c4 = 3 (f1 - f0) - 2 s0 - s1, c3 = s0 + s1 - 2 (f1 - f0)); the derivative is ((3 c3 p + 2 c4) p + c5) * rdr instead
//    of the pre-divided coefficients (eam_utils.c:279-283), a last-bit difference.  64 B of table per pair instead of 96.
template <class real>
__global__ void k_eam_pack_vs(int rows, const real* __restrict__ rhor_spline, const real* __restrict__ z2r_spline, real* __restrict__ vs4)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= rows) return;
    real* a = vs4 + (size_t)m * 4;
    a[0] = rhor_spline[(size_t)m * 7 + 6]; a[1] = rhor_spline[(size_t)m * 7 + 5];
    a[2] = z2r_spline[(size_t)m * 7 + 6];  a[3] = z2r_spline[(size_t)m * 7 + 5];
}
template <class real, int U, bool PF>
__global__ void __launch_bounds__(128) k_eam_density_v3(int nlocal, real cutforcesq, EamTables<real> t, const real* __restrict__ rho4,
    const real* __restrict__ frho_spline, const real* __restrict__ x, const real* __restrict__ y, const real* __restrict__ z,
    const typename Vec2Of<real>::type* __restrict__ xy, const int* __restrict__ numneigh, const int* __restrict__ nbT, NbLayout L,
    real* __restrict__ fp, typename Vec2Of<real>::type* __restrict__ zf)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nlocal) return;
    const real xt = x[i], yt = y[i], zt = z[i];
    const int nn  = numneigh[i];
    const int* nb = nbT + L.base(i);
    real rhoi     = 0;
    int j[U], jn[U];
#pragma unroll
    for (int u = 0; u < U; u++) j[u] = u < nn ? __ldg(nb + (size_t)u * L.sk) : i;
    for (int k = 0; k < nn; k += U) {
        real rsq[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const typename Vec2Of<real>::type p = __ldg(xy + j[u]);
            const real dx = xt - p.x, dy = yt - p.y, dz = zt - __ldg(z + j[u]);
            rsq[u] = dx * dx + dy * dy + dz * dz;
        }
        if (PF) { // the indices of the next group are requested before this group is evaluated
#pragma unroll
            for (int u = 0; u < U; u++) jn[u] = k + U + u < nn ? __ldg(nb + (size_t)(k + U + u) * L.sk) : i;
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (rsq[u] < cutforcesq && k + u < nn) {
                const real r = rsq[u] * rsqrt_nr(rsq[u]);
                real p = r * t.rdr + (real)1.0;
                int m  = (int)(p);
                m      = m < t.nr - 1 ? m : t.nr - 1;
                p -= m;
                p = p < (real)1.0 ? p : (real)1.0;
                real s3, s4, s5, s6;
                ld4(rho4 + (size_t)m * 4, s3, s4, s5, s6);
                rhoi += ((s3 * p + s4) * p + s5) * p + s6;
            }
        }
        if (!PF) {
#pragma unroll
            for (int u = 0; u < U; u++) jn[u] = k + U + u < nn ? __ldg(nb + (size_t)(k + U + u) * L.sk) : i;
        }
#pragma unroll
        for (int u = 0; u < U; u++) j[u] = jn[u];
    }
    real p = (real)1.0 * rhoi * t.rdrho + (real)1.0;
    int m  = (int)(p);
    m      = max(1, min(m, t.nrho - 1));
    p -= m;
    p = min(p, (real)1.0);
    const real* s = frho_spline + m * 7;
    const real f  = (__ldg(s + 0) * p + __ldg(s + 1)) * p + __ldg(s + 2);
    fp[i]         = f;
    typename Vec2Of<real>::type v;
    v.x = zt; v.y = f;
    zf[i] = v;
}
template <class real>
__global__ void k_eam_ghost_fp_v3(int nlocal, int nghost, const int* __restrict__ border_map, const real* __restrict__ z,
    real* __restrict__ fp, typename Vec2Of<real>::type* __restrict__ zf)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nghost) return;
    const real f = fp[border_map[g]];
    fp[nlocal + g] = f;
    typename Vec2Of<real>::type v;
    v.x = z[nlocal + g]; v.y = f;
    zf[nlocal + g] = v;
}
// FI: finalIntegrate(n) + initialIntegrate(n+1) in the epilogue (integrate.c:21-40; the operation sequence of
// k_final_initial_integrate, so a run equals the separate kernels bit for bit).  Every gather of this pass and of the next
// density pass goes to the copies (xy, zf), which are rebuilt per force call, so x, y, z are updated IN PLACE and f is
// not stored.
template <class real> struct EamIntegrate {
    real *vx, *vy, *vz;
    real dtforce, dt;
};
template <class real, int U, bool PF, bool FI = false>
__global__ void __launch_bounds__(128, FI ? 7 : 1) k_eam_force_v3(int nlocal, real cutforcesq, EamTables<real> t, const real* __restrict__ vs4,
    real* x, real* y, real* z, const real* __restrict__ fp,
    const typename Vec2Of<real>::type* __restrict__ xy, const typename Vec2Of<real>::type* __restrict__ zf,
    const int* __restrict__ numneigh, const int* __restrict__ nbT, NbLayout L, real* __restrict__ fx, real* __restrict__ fy,
    real* __restrict__ fz, EamIntegrate<real> fi)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nlocal) return;
    const real xt = x[i], yt = y[i], zt = z[i], fpi = fp[i];
    const int nn  = numneigh[i];
    const int* nb = nbT + L.base(i);
    real fix = 0, fiy = 0, fiz = 0;
    int j[U], jn[U];
#pragma unroll
    for (int u = 0; u < U; u++) j[u] = u < nn ? __ldg(nb + (size_t)u * L.sk) : i;
    for (int k = 0; k < nn; k += U) {
        real dx[U], dy[U], dz[U], rsq[U], fpj[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const typename Vec2Of<real>::type p = __ldg(xy + j[u]), q = __ldg(zf + j[u]);
            dx[u] = xt - p.x; dy[u] = yt - p.y; dz[u] = zt - q.x;
            fpj[u] = q.y;
            rsq[u] = dx[u] * dx[u] + dy[u] * dy[u] + dz[u] * dz[u];
        }
        if (PF) {
#pragma unroll
            for (int u = 0; u < U; u++) jn[u] = k + U + u < nn ? __ldg(nb + (size_t)(k + U + u) * L.sk) : i;
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (rsq[u] < cutforcesq && k + u < nn) {
                const real recip = rsqrt_nr(rsq[u]);
                const real r     = rsq[u] * recip;
                real p           = r * t.rdr + (real)1.0;
                int m            = (int)(p);
                m                = m < t.nr - 1 ? m : t.nr - 1;
                p -= m;
                p = p < (real)1.0 ? p : (real)1.0;
                real rf0, rs0, zf0, zs0, rf1, rs1, zf1, zs1;
                ld4(vs4 + (size_t)m * 4, rf0, rs0, zf0, zs0);
                ld4(vs4 + (size_t)m * 4 + 4, rf1, rs1, zf1, zs1);
                const real dr_ = rf1 - rf0, dz_ = zf1 - zf0;
                const real rc4 = (real)3.0 * dr_ - (real)2.0 * rs0 - rs1, rc3 = rs0 + rs1 - (real)2.0 * dr_;
                const real zc4 = (real)3.0 * dz_ - (real)2.0 * zs0 - zs1, zc3 = zs0 + zs1 - (real)2.0 * dz_;
                const real rhoip = (((real)3.0 * rc3 * p + (real)2.0 * rc4) * p + rs0) * t.rdr;
                const real z2p   = (((real)3.0 * zc3 * p + (real)2.0 * zc4) * p + zs0) * t.rdr;
                const real z2    = ((zc3 * p + zc4) * p + zs0) * p + zf0;
                const real phi   = z2 * recip;
                const real phip  = z2p * recip - phi * recip;
                const real psip  = fpi * rhoip + fpj[u] * rhoip + phip;
                const real fpair = -psip * recip;
                fix += dx[u] * fpair;
                fiy += dy[u] * fpair;
                fiz += dz[u] * fpair;
            }
        }
        if (!PF) {
#pragma unroll
            for (int u = 0; u < U; u++) jn[u] = k + U + u < nn ? __ldg(nb + (size_t)(k + U + u) * L.sk) : i;
        }
#pragma unroll
        for (int u = 0; u < U; u++) j[u] = jn[u];
    }
    if (FI) {
        real a = fi.vx[i] + fi.dtforce * fix, b = fi.vy[i] + fi.dtforce * fiy, c = fi.vz[i] + fi.dtforce * fiz; // final(n)
        a = a + fi.dtforce * fix; b = b + fi.dtforce * fiy; c = c + fi.dtforce * fiz;                          // initial(n+1)
        fi.vx[i] = a; fi.vy[i] = b; fi.vz[i] = c;
        x[i] = xt + fi.dt * a;
        y[i] = yt + fi.dt * b;
        z[i] = zt + fi.dt * c;
    } else {
        fx[i] = fix;
        fy[i] = fiy;
        fz[i] = fiz;
    }
}

} // namespace mdb
