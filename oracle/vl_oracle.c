/*
 * oracle/vl_oracle.c -- TEST INFRASTRUCTURE ONLY (the parity oracle, never shipped, never timed as
 * the product).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
 *
 * A plain-C, single-threaded restatement of MD-Bench's verletlist hot path, written from the
 * reference's behaviour (file:line cited per function, paths relative to the reference root).
 * It is compiled with -ffp-contract=off and every fused multiply-add the reference *build*
 * performs where bits matter is written out as an explicit fma() call, so the list
 * membership and ghost coordinates are a pure function of the inputs.
 *
 * PINNING: this restatement is checked against the unmodified reference compiled from
 * /root/reference (oracle/_ref, see oracle/Makefile) in tests/test_oracle_pinned.py -- neighbor
 * sets / ghost maps / ghost coordinates / bins bit-exact, forces and trajectories to rounding --
 * and against fixtures generated from that reference (tests/golden/make_golden.py).  The reference
 * itself ships no tests or golden vectors.
 *
 * One build per precision: -DPRECISION=2 (double) or 1 (float), like the reference's MD_FLOAT.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#if PRECISION == 1
typedef float real;
#define RFMA fmaf
#define RSQRT sqrtf
#else
typedef double real;
#define RFMA fma
#define RSQRT sqrt
#endif

#define DELTA 20000 /* growth quantum of atom and ghost arrays: verletlist/atom.c:17, pbc.c:14 */

typedef struct {
    /* ---- parameters: common/parameter.h:27-61, defaults common/parameter.c:16-51 ---- */
    int force_field; /* 0 = lj, 1 = eam */
    real epsilon, sigma, sigma6, temp, rho, mass;
    int ntypes, ntimes, nstat, reneigh_every, half_neigh;
    real dt, dtforce, skin, cutforce, cutneigh;
    int nx, ny, nz, pbc_x, pbc_y, pbc_z;
    real lattice, xprd, yprd, zprd;
    int from_input; /* 1: box came from an input file (setupNeighbor's other branch) */
    real xlo, xhi, ylo, yhi, zlo, zhi;
    /* ---- atoms: verletlist/atom.h:24-39 (SoA here; layout is not part of the arithmetic) ---- */
    int Natoms, Nlocal, Nghost, Nmax;
    real *x, *y, *z, *vx, *vy, *vz, *fx, *fy, *fz;
    int* type;
    int *border_map, *PBCx, *PBCy, *PBCz;
    int NmaxGhost;
    /* ---- neighbor state: verletlist/neighbor.c:24-38 ---- */
    real bininvx, bininvy, bininvz, binsizex, binsizey, binsizez, cutneighsq;
    int mbinxlo, mbinylo, mbinzlo, nbinx, nbiny, nbinz, mbinx, mbiny, mbinz, mbins;
    int atoms_per_bin, nstencil, nmax, maxneighs;
    int *bincount, *bins, *stencil, *numneigh, *neighbors;
    /* ---- thermo: common/thermo.c:14-26 ---- */
    real mvv2e, dof_boltz, t_scale, p_scale;
    /* ---- eam: common/eam.h:20-29 ---- */
    int nrho, nr, nrho_tot, nr_tot, fp_nmax;
    real rdr, rdrho, *rhor_spline, *frho_spline, *z2r_spline, *fp;
    /* counters for the roofline's algorithmic-work model (verletlist/stats.h) */
    long long pairs_listed, pairs_in_cutoff, force_calls;
} OVL;

#define EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* common/parameter.c:16-51 initParameter; main.c:233 cutneigh; main.c:42-45 lattice/box       */
EXPORT OVL* ovl_new(void)
{
    OVL* o           = (OVL*)calloc(1, sizeof(OVL));
    o->force_field   = 0;
    o->epsilon       = 1.0;
    o->sigma         = 1.0;
    o->sigma6        = 1.0;
    o->rho           = 0.8442;
    o->ntypes        = 1;
    o->ntimes        = 200;
    o->dt            = 0.005;
    o->nx = o->ny = o->nz = 32;
    o->pbc_x = o->pbc_y = o->pbc_z = 1;
    o->cutforce      = 2.5;
    o->skin          = 0.3;
    o->cutneigh      = o->cutforce + o->skin;
    o->temp          = 1.44;
    o->nstat         = 100;
    o->mass          = 1.0;
    o->dtforce       = 0.5 * o->dt;
    o->reneigh_every = 20;
    o->half_neigh    = 0;
    o->atoms_per_bin = 8;   /* neighbor.c:54 */
    o->maxneighs     = 100; /* neighbor.c:58 */
    return o;
}

EXPORT void ovl_free(OVL* o)
{
    real** r[] = { &o->x, &o->y, &o->z, &o->vx, &o->vy, &o->vz, &o->fx, &o->fy, &o->fz,
        &o->rhor_spline, &o->frho_spline, &o->z2r_spline, &o->fp };
    for (unsigned i = 0; i < sizeof(r) / sizeof(r[0]); i++) free(*r[i]);
    int** q[] = { &o->type, &o->border_map, &o->PBCx, &o->PBCy, &o->PBCz, &o->bincount, &o->bins,
        &o->stencil, &o->numneigh, &o->neighbors };
    for (unsigned i = 0; i < sizeof(q) / sizeof(q[0]); i++) free(*q[i]);
    free(o);
}

/* Setters taking doubles; narrowing to `real` happens exactly where the reference narrows
 * (assignment of a double expression to an MD_FLOAT field). */
EXPORT void ovl_set_lj(OVL* o, double epsilon, double sigma, double cutforce, double skin,
    double dt, double temp, double rho, double mass)
{
    o->epsilon  = epsilon;
    o->sigma    = sigma;
    real s2     = o->sigma * o->sigma; /* parameter.c:118-119 */
    o->sigma6   = s2 * s2 * s2;
    o->cutforce = cutforce;
    o->skin     = skin;
    o->dt       = dt;
    o->dtforce  = 0.5 * o->dt; /* parameter.c:115 */
    o->temp     = temp;
    o->rho      = rho;
    o->mass     = mass;
}
EXPORT void ovl_set_run(OVL* o, int nx, int ny, int nz, int ntimes, int nstat, int reneigh_every,
    int half_neigh, int pbc_x, int pbc_y, int pbc_z)
{
    o->nx = nx; o->ny = ny; o->nz = nz;
    o->ntimes = ntimes; o->nstat = nstat; o->reneigh_every = reneigh_every;
    o->half_neigh = half_neigh;
    o->pbc_x = pbc_x; o->pbc_y = pbc_y; o->pbc_z = pbc_z;
}
EXPORT void ovl_set_force_field(OVL* o, int ff) { o->force_field = ff; }

/* main.c:233 and main.c:42-45 (evaluated in double by pow(), narrowed on assignment) */
EXPORT void ovl_derive(OVL* o)
{
    if (o->force_field == 0) o->cutneigh = o->cutforce + o->skin;
    o->lattice = pow((4.0 / o->rho), (1.0 / 3.0));
    if (!o->from_input) {
        o->xprd = o->nx * o->lattice;
        o->yprd = o->ny * o->lattice;
        o->zprd = o->nz * o->lattice;
    }
}

/* box taken from an input file: atom.c readers set param->xlo..zhi, xprd.. */
EXPORT void ovl_set_box(OVL* o, double xlo, double xhi, double ylo, double yhi, double zlo, double zhi)
{
    o->from_input = 1;
    o->xlo = xlo; o->xhi = xhi; o->ylo = ylo; o->yhi = yhi; o->zlo = zlo; o->zhi = zhi;
    o->xprd = o->xhi - o->xlo;
    o->yprd = o->yhi - o->ylo;
    o->zprd = o->zhi - o->zlo;
}

/* verletlist/atom.c:590-618 growAtom */
static void grow_atom(OVL* o)
{
    int nold = o->Nmax;
    o->Nmax += DELTA;
    real** r[] = { &o->x, &o->y, &o->z, &o->vx, &o->vy, &o->vz, &o->fx, &o->fy, &o->fz };
    for (int i = 0; i < 9; i++) {
        real* p = (real*)calloc(o->Nmax, sizeof(real));
        if (*r[i]) { memcpy(p, *r[i], nold * sizeof(real)); free(*r[i]); }
        *r[i] = p;
    }
    int* t = (int*)calloc(o->Nmax, sizeof(int));
    if (o->type) { memcpy(t, o->type, nold * sizeof(int)); free(o->type); }
    o->type = t;
}

/* common/util.c:24-33 Park-Miller minimal standard, Schrage factorisation */
static double lcg(int* seed)
{
    const int IA = 16807, IM = 2147483647, IQ = 127773, IR = 2836;
    int k = (*seed) / IQ;
    *seed = IA * (*seed - k * IQ) - IR * k;
    if (*seed < 0) *seed += IM;
    return (1.0 / IM) * (*seed);
}

/* verletlist/atom.c:67-187 createAtom: FCC sites (i+j+k even) on the half-lattice, emitted by
 * walking 8x8x8 sub-boxes; velocities from the LCG seeded by the site number. */
EXPORT void ovl_create_atoms(OVL* o)
{
    real xlo = 0.0, xhi = o->xprd, ylo = 0.0, yhi = o->yprd, zlo = 0.0, zhi = o->zprd;
    o->Natoms = 4 * o->nx * o->ny * o->nz;
    o->Nlocal = 0;
    real alat = pow((4.0 / o->rho), (1.0 / 3.0));
    int ilo = (int)(xlo / (0.5 * alat) - 1), ihi = (int)(xhi / (0.5 * alat) + 1);
    int jlo = (int)(ylo / (0.5 * alat) - 1), jhi = (int)(yhi / (0.5 * alat) + 1);
    int klo = (int)(zlo / (0.5 * alat) - 1), khi = (int)(zhi / (0.5 * alat) + 1);
    if (ilo < 0) ilo = 0;
    if (ihi > 2 * o->nx - 1) ihi = 2 * o->nx - 1;
    if (jlo < 0) jlo = 0;
    if (jhi > 2 * o->ny - 1) jhi = 2 * o->ny - 1;
    if (klo < 0) klo = 0;
    if (khi > 2 * o->nz - 1) khi = 2 * o->nz - 1;
    int sx = 0, sy = 0, sz = 0, ox = 0, oy = 0, oz = 0;
    const int sub = 8;
    while (oz * sub <= khi) {
        int k = oz * sub + sz, j = oy * sub + sy, i = ox * sub + sx;
        if (((i + j + k) % 2 == 0) && i >= ilo && i <= ihi && j >= jlo && j <= jhi && k >= klo &&
            k <= khi) {
            real xt = 0.5 * alat * i, yt = 0.5 * alat * j, zt = 0.5 * alat * k;
            if (xt >= xlo && xt < xhi && yt >= ylo && yt < yhi && zt >= zlo && zt < zhi) {
                int n = k * (2 * o->ny) * (2 * o->nx) + j * (2 * o->nx) + i + 1;
                real v[3];
                for (int c = 0; c < 3; c++) {
                    for (int m = 0; m < 5; m++) lcg(&n);
                    v[c] = lcg(&n);
                }
                if (o->Nlocal == o->Nmax) grow_atom(o);
                int a = o->Nlocal++;
                o->x[a] = xt; o->y[a] = yt; o->z[a] = zt;
                o->vx[a] = v[0]; o->vy[a] = v[1]; o->vz[a] = v[2];
                o->type[a] = 0; /* rand() % ntypes with ntypes == 1 */
            }
        }
        sx++;
        if (sx == sub) { sx = 0; sy++; }
        if (sy == sub) { sy = 0; sz++; }
        if (sz == sub) { sz = 0; ox++; }
        if (ox * sub > ihi) { ox = 0; oy++; }
        if (oy * sub > jhi) { oy = 0; oz++; }
    }
}

/* feed arbitrary atoms (what the file readers atom.c:199-562 end up doing) */
EXPORT void ovl_set_atoms(OVL* o, int n, const real* x, const real* y, const real* z,
    const real* vx, const real* vy, const real* vz)
{
    while (o->Nmax < n) grow_atom(o);
    o->Natoms = o->Nlocal = n;
    o->Nghost = 0;
    for (int i = 0; i < n; i++) {
        o->x[i] = x[i]; o->y[i] = y[i]; o->z[i] = z[i];
        o->vx[i] = vx ? vx[i] : 0; o->vy[i] = vy ? vy[i] : 0; o->vz[i] = vz ? vz[i] : 0;
        o->type[i] = 0;
    }
}

/* common/thermo.c:30-53 setupThermo */
EXPORT void ovl_setup_thermo(OVL* o)
{
    int natoms = o->Natoms;
    if (o->force_field == 0) {
        o->mvv2e     = 1.0;
        o->dof_boltz = (natoms * 3 - 3);
        o->t_scale   = o->mvv2e / o->dof_boltz;
        o->p_scale   = 1.0 / 3 / o->xprd / o->yprd / o->zprd;
    } else {
        o->mvv2e     = 1.036427e-04;
        o->dof_boltz = (natoms * 3 - 3) * 8.617343e-05;
        o->t_scale   = o->mvv2e / o->dof_boltz;
        o->p_scale   = 1.602176e+06 / 3 / o->xprd / o->yprd / o->zprd;
        o->dtforce = 0.5 * o->dt / o->mass; /* initEam, eam_utils.c:35 (kept idempotent: setup may run twice) */
        o->dtforce /= o->mvv2e;             /* thermo.c:51 */
    }
}

/* common/thermo.c:55-66 computeThermo (values only; the driver prints "%i\t%e\t%e") */
EXPORT void ovl_thermo(OVL* o, double* T, double* P)
{
    real t = 0.0;
    for (int i = 0; i < o->Nlocal; i++)
        t += (o->vx[i] * o->vx[i] + o->vy[i] * o->vy[i] + o->vz[i] * o->vz[i]) * o->mass;
    t  = t * o->t_scale;
    *T = t;
    *P = (t * o->dof_boltz) * o->p_scale;
}

/* common/thermo.c:82-122 adjustThermo: zero centre-of-mass velocity, rescale to param->temp */
EXPORT void ovl_adjust_thermo(OVL* o)
{
    real vxtot = 0.0, vytot = 0.0, vztot = 0.0;
    for (int i = 0; i < o->Nlocal; i++) { vxtot += o->vx[i]; vytot += o->vy[i]; vztot += o->vz[i]; }
    vxtot = vxtot / o->Natoms; vytot = vytot / o->Natoms; vztot = vztot / o->Natoms;
    for (int i = 0; i < o->Nlocal; i++) { o->vx[i] -= vxtot; o->vy[i] -= vytot; o->vz[i] -= vztot; }
    real t = 0.0;
    for (int i = 0; i < o->Nlocal; i++)
        t += (o->vx[i] * o->vx[i] + o->vy[i] * o->vy[i] + o->vz[i] * o->vz[i]) * o->mass;
    t *= o->t_scale;
    real factor = sqrt(o->temp / t);
    for (int i = 0; i < o->Nlocal; i++) { o->vx[i] *= factor; o->vy[i] *= factor; o->vz[i] *= factor; }
}

/* ------------------------------------------------------------------------------------------ */
/* verletlist/neighbor.c:267-296 bindist */
static real bindist(OVL* o, int i, int j, int k)
{
    real delx, dely, delz;
    delx = i > 0 ? (i - 1) * o->binsizex : (i == 0 ? 0.0 : (i + 1) * o->binsizex);
    dely = j > 0 ? (j - 1) * o->binsizey : (j == 0 ? 0.0 : (j + 1) * o->binsizey);
    delz = k > 0 ? (k - 1) * o->binsizez : (k == 0 ? 0.0 : (k + 1) * o->binsizez);
    return (delx * delx + dely * dely + delz * delz);
}

/* verletlist/neighbor.c:43-62 initNeighbor + 64-184 setupNeighbor */
EXPORT void ovl_setup_neighbor(OVL* o)
{
    const real SMALL = 1.0e-6, FACTOR = 0.999;
    real neighscale = 5.0 / 6.0;
    real xprd = o->nx * o->lattice, yprd = o->ny * o->lattice, zprd = o->nz * o->lattice;
    real cutneigh = o->cutneigh;
    o->nbinx = neighscale * o->nx;
    o->nbiny = neighscale * o->ny;
    o->nbinz = neighscale * o->nz;
    if (o->from_input) { xprd = o->xprd; yprd = o->yprd; zprd = o->zprd; }
    real xlo = 0.0, xhi = xprd, ylo = 0.0, yhi = yprd, zlo = 0.0, zhi = zprd;
    o->cutneighsq = cutneigh * cutneigh;
    if (o->from_input) {
        o->binsizex = o->binsizey = o->binsizez = cutneigh * 0.5;
        o->nbinx = (int)((o->xhi - o->xlo) / o->binsizex);
        o->nbiny = (int)((o->yhi - o->ylo) / o->binsizey);
        o->nbinz = (int)((o->zhi - o->zlo) / o->binsizez);
        if (o->nbinx == 0) o->nbinx = 1;
        if (o->nbiny == 0) o->nbiny = 1;
        if (o->nbinz == 0) o->nbinz = 1;
        o->bininvx = o->nbinx / (o->xhi - o->xlo);
        o->bininvy = o->nbiny / (o->yhi - o->ylo);
        o->bininvz = o->nbinz / (o->zhi - o->zlo);
    } else {
        o->binsizex = xprd / o->nbinx;
        o->binsizey = yprd / o->nbiny;
        o->binsizez = zprd / o->nbinz;
        o->bininvx  = 1.0 / o->binsizex;
        o->bininvy  = 1.0 / o->binsizey;
        o->bininvz  = 1.0 / o->binsizez;
    }
    real coord;
    int mbinxhi, mbinyhi, mbinzhi;
    coord = xlo - cutneigh - SMALL * xprd;
    o->mbinxlo = (int)(coord * o->bininvx);
    if (coord < 0.0) o->mbinxlo -= 1;
    coord   = xhi + cutneigh + SMALL * xprd;
    mbinxhi = (int)(coord * o->bininvx);
    coord = ylo - cutneigh - SMALL * yprd;
    o->mbinylo = (int)(coord * o->bininvy);
    if (coord < 0.0) o->mbinylo -= 1;
    coord   = yhi + cutneigh + SMALL * yprd;
    mbinyhi = (int)(coord * o->bininvy);
    coord = zlo - cutneigh - SMALL * zprd;
    o->mbinzlo = (int)(coord * o->bininvz);
    if (coord < 0.0) o->mbinzlo -= 1;
    coord   = zhi + cutneigh + SMALL * zprd;
    mbinzhi = (int)(coord * o->bininvz);
    o->mbinxlo -= 1; mbinxhi += 1; o->mbinx = mbinxhi - o->mbinxlo + 1;
    o->mbinylo -= 1; mbinyhi += 1; o->mbiny = mbinyhi - o->mbinylo + 1;
    o->mbinzlo -= 1; mbinzhi += 1; o->mbinz = mbinzhi - o->mbinzlo + 1;
    int nextx = (int)(cutneigh * o->bininvx);
    if (nextx * o->binsizex < FACTOR * cutneigh) nextx++;
    int nexty = (int)(cutneigh * o->bininvy);
    if (nexty * o->binsizey < FACTOR * cutneigh) nexty++;
    int nextz = (int)(cutneigh * o->bininvz);
    if (nextz * o->binsizez < FACTOR * cutneigh) nextz++;
    free(o->stencil);
    o->stencil  = (int*)malloc((2 * nextz + 1) * (2 * nexty + 1) * (2 * nextx + 1) * sizeof(int));
    o->nstencil = 0;
    for (int k = -nextz; k <= nextz; k++)
        for (int j = -nexty; j <= nexty; j++)
            for (int i = -nextx; i <= nextx; i++)
                if (bindist(o, i, j, k) < o->cutneighsq)
                    o->stencil[o->nstencil++] = k * o->mbiny * o->mbinx + j * o->mbinx + i;
    o->mbins = o->mbinx * o->mbiny * o->mbinz;
    free(o->bincount);
    free(o->bins);
    o->bincount = (int*)malloc(o->mbins * sizeof(int));
    o->bins     = (int*)malloc((size_t)o->mbins * o->atoms_per_bin * sizeof(int));
    /* the bin functions below use the box extents stored here */
    if (!o->from_input) { o->xprd = xprd; o->yprd = yprd; o->zprd = zprd; }
}

/* verletlist/neighbor.c:298-327 coord2bin (including the stray "+ 1") */
static int coord2bin(OVL* o, real xin, real yin, real zin)
{
    int ix, iy, iz;
    if (xin >= o->xprd) ix = (int)((xin - o->xprd) * o->bininvx) + o->nbinx - o->mbinxlo;
    else if (xin >= 0.0) ix = (int)(xin * o->bininvx) - o->mbinxlo;
    else ix = (int)(xin * o->bininvx) - o->mbinxlo - 1;
    if (yin >= o->yprd) iy = (int)((yin - o->yprd) * o->bininvy) + o->nbiny - o->mbinylo;
    else if (yin >= 0.0) iy = (int)(yin * o->bininvy) - o->mbinylo;
    else iy = (int)(yin * o->bininvy) - o->mbinylo - 1;
    if (zin >= o->zprd) iz = (int)((zin - o->zprd) * o->bininvz) + o->nbinz - o->mbinzlo;
    else if (zin >= 0.0) iz = (int)(zin * o->bininvz) - o->mbinzlo;
    else iz = (int)(zin * o->bininvz) - o->mbinzlo - 1;
    return (iz * o->mbiny * o->mbinx + iy * o->mbinx + ix + 1);
}
EXPORT int ovl_coord2bin(OVL* o, double x, double y, double z) { return coord2bin(o, x, y, z); }

/* verletlist/neighbor.c:329-358 binatoms */
EXPORT void ovl_binatoms(OVL* o)
{
    int nall = o->Nlocal + o->Nghost, resize = 1;
    while (resize > 0) {
        resize = 0;
        for (int i = 0; i < o->mbins; i++) o->bincount[i] = 0;
        for (int i = 0; i < nall; i++) {
            int ibin = coord2bin(o, o->x[i], o->y[i], o->z[i]);
            if (o->bincount[ibin] < o->atoms_per_bin) {
                int ac = o->bincount[ibin]++;
                o->bins[(size_t)ibin * o->atoms_per_bin + ac] = i;
            } else {
                resize = 1;
            }
        }
        if (resize) {
            free(o->bins);
            o->atoms_per_bin *= 2;
            o->bins = (int*)malloc((size_t)o->mbins * o->atoms_per_bin * sizeof(int));
        }
    }
}

/* verletlist/neighbor.c:186-264 buildNeighborCPU.  Distance uses the FMA nesting of the
 * reference's GCC -Ofast build, fma(dx,dx,fma(dy,dy,dz*dz)) (verified by objdump of
 * oracle/_ref/libmdref_vl_dp_aos.so); inclusion test is "<=" (neighbor.c:240).  Rows never
 * overrun: the count continues but stores stop at maxneighs, then the list is re-made. */
EXPORT void ovl_build_neighbor(OVL* o)
{
    int nall = o->Nlocal + o->Nghost;
    if (nall > o->nmax) {
        o->nmax = nall;
        free(o->numneigh);
        free(o->neighbors);
        o->numneigh  = (int*)malloc(o->nmax * sizeof(int));
        o->neighbors = (int*)malloc((size_t)o->nmax * o->maxneighs * sizeof(int));
    }
    ovl_binatoms(o);
    int resize = 1;
    while (resize) {
        int new_maxneighs = o->maxneighs;
        resize            = 0;
        for (int i = 0; i < o->Nlocal; i++) {
            int* neighptr = &o->neighbors[(size_t)i * o->maxneighs];
            int n         = 0;
            real xtmp = o->x[i], ytmp = o->y[i], ztmp = o->z[i];
            int ibin = coord2bin(o, xtmp, ytmp, ztmp);
            for (int k = 0; k < o->nstencil; k++) {
                int jbin     = ibin + o->stencil[k];
                int* loc_bin = &o->bins[(size_t)jbin * o->atoms_per_bin];
                for (int m = 0; m < o->bincount[jbin]; m++) {
                    int j = loc_bin[m];
                    if ((j == i) || (o->half_neigh && (j < i))) continue;
                    real delx = xtmp - o->x[j], dely = ytmp - o->y[j], delz = ztmp - o->z[j];
                    real rsq  = RFMA(delx, delx, RFMA(dely, dely, delz * delz));
                    if (rsq <= o->cutneighsq) {
                        if (n < o->maxneighs) neighptr[n] = j;
                        n++;
                    }
                }
            }
            o->numneigh[i] = n;
            if (n >= o->maxneighs) {
                resize = 1;
                if (n >= new_maxneighs) new_maxneighs = n;
            }
        }
        if (resize) {
            o->maxneighs = new_maxneighs * 1.2;
            free(o->neighbors);
            o->neighbors = (int*)malloc((size_t)o->nmax * o->maxneighs * sizeof(int));
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* verletlist/pbc.c:230-242 growPbc */
static void grow_pbc(OVL* o)
{
    int nold = o->NmaxGhost;
    o->NmaxGhost += DELTA;
    int** q[] = { &o->border_map, &o->PBCx, &o->PBCy, &o->PBCz };
    for (int i = 0; i < 4; i++) {
        int* p = (int*)calloc(o->NmaxGhost, sizeof(int));
        if (*q[i]) { memcpy(p, *q[i], nold * sizeof(int)); free(*q[i]); }
        *q[i] = p;
    }
}

/* verletlist/pbc.c:98-227 setupPbc: ghost creation in the fixed order
 * 6 faces -> 8 corners -> 12 edges (x-z, y-z, x-y) per local atom. */
EXPORT void ovl_setup_pbc(OVL* o)
{
    real xprd = o->xprd, yprd = o->yprd, zprd = o->zprd, cut = o->cutneigh;
    int Nghost = -1;
#define ADDGHOST(dx, dy, dz)                                                                     \
    do {                                                                                         \
        Nghost++;                                                                                \
        o->border_map[Nghost] = i;                                                               \
        o->PBCx[Nghost] = dx; o->PBCy[Nghost] = dy; o->PBCz[Nghost] = dz;                        \
        o->type[o->Nlocal + Nghost] = o->type[i];                                                \
    } while (0)
    for (int i = 0; i < o->Nlocal; i++) {
        if (o->Nlocal + Nghost + 7 >= o->Nmax) grow_atom(o);
        if (Nghost + 7 >= o->NmaxGhost) grow_pbc(o);
        real x = o->x[i], y = o->y[i], z = o->z[i];
        int xl = x < cut, xh = x >= (xprd - cut);
        int yl = y < cut, yh = y >= (yprd - cut);
        int zl = z < cut, zh = z >= (zprd - cut);
        if (o->pbc_x != 0) { if (xl) ADDGHOST(+1, 0, 0); if (xh) ADDGHOST(-1, 0, 0); }
        if (o->pbc_y != 0) { if (yl) ADDGHOST(0, +1, 0); if (yh) ADDGHOST(0, -1, 0); }
        if (o->pbc_z != 0) { if (zl) ADDGHOST(0, 0, +1); if (zh) ADDGHOST(0, 0, -1); }
        if (o->pbc_x != 0 && o->pbc_y != 0 && o->pbc_z != 0) {
            if (xl && yl && zl) ADDGHOST(+1, +1, +1);
            if (xl && yh && zl) ADDGHOST(+1, -1, +1);
            if (xl && yl && zh) ADDGHOST(+1, +1, -1);
            if (xl && yh && zh) ADDGHOST(+1, -1, -1);
            if (xh && yl && zl) ADDGHOST(-1, +1, +1);
            if (xh && yh && zl) ADDGHOST(-1, -1, +1);
            if (xh && yl && zh) ADDGHOST(-1, +1, -1);
            if (xh && yh && zh) ADDGHOST(-1, -1, -1);
        }
        if (o->pbc_x != 0 && o->pbc_z != 0) {
            if (xl && zl) ADDGHOST(+1, 0, +1);
            if (xl && zh) ADDGHOST(+1, 0, -1);
            if (xh && zl) ADDGHOST(-1, 0, +1);
            if (xh && zh) ADDGHOST(-1, 0, -1);
        }
        if (o->pbc_y != 0 && o->pbc_z != 0) {
            if (yl && zl) ADDGHOST(0, +1, +1);
            if (yl && zh) ADDGHOST(0, +1, -1);
            if (yh && zl) ADDGHOST(0, -1, +1);
            if (yh && zh) ADDGHOST(0, -1, -1);
        }
        if (o->pbc_x != 0 && o->pbc_y != 0) {
            if (yl && xl) ADDGHOST(+1, +1, 0);
            if (yl && xh) ADDGHOST(-1, +1, 0);
            if (yh && xl) ADDGHOST(+1, -1, 0);
            if (yh && xh) ADDGHOST(-1, -1, 0);
        }
    }
#undef ADDGHOST
    o->Nghost = Nghost + 1;
}

/* verletlist/pbc.c:42-55 updatePbcCPU.  x + PBC*prd is a single fma in the reference build
 * (SURVEY F11); ghost coordinates feed the list build, so the contraction is written out. */
EXPORT void ovl_update_pbc(OVL* o)
{
    int nl = o->Nlocal;
    for (int i = 0; i < o->Nghost; i++) {
        int s        = o->border_map[i];
        o->x[nl + i] = RFMA((real)o->PBCx[i], o->xprd, o->x[s]);
        o->y[nl + i] = RFMA((real)o->PBCy[i], o->yprd, o->y[s]);
        o->z[nl + i] = RFMA((real)o->PBCz[i], o->zprd, o->z[s]);
    }
}

/* verletlist/pbc.c:59-84 updateAtomsPbcCPU */
EXPORT void ovl_update_atoms_pbc(OVL* o)
{
    for (int i = 0; i < o->Nlocal; i++) {
        if (o->x[i] < 0.0) o->x[i] += o->xprd; else if (o->x[i] >= o->xprd) o->x[i] -= o->xprd;
        if (o->y[i] < 0.0) o->y[i] += o->yprd; else if (o->y[i] >= o->yprd) o->y[i] -= o->yprd;
        if (o->z[i] < 0.0) o->z[i] += o->zprd; else if (o->z[i] >= o->zprd) o->z[i] -= o->zprd;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* verletlist/integrate.c:21-31 and 33-40 */
EXPORT void ovl_initial_integrate(OVL* o)
{
    for (int i = 0; i < o->Nlocal; i++) {
        o->vx[i] += o->dtforce * o->fx[i];
        o->vy[i] += o->dtforce * o->fy[i];
        o->vz[i] += o->dtforce * o->fz[i];
        o->x[i] = o->x[i] + o->dt * o->vx[i];
        o->y[i] = o->y[i] + o->dt * o->vy[i];
        o->z[i] = o->z[i] + o->dt * o->vz[i];
    }
}
EXPORT void ovl_final_integrate(OVL* o)
{
    for (int i = 0; i < o->Nlocal; i++) {
        o->vx[i] += o->dtforce * o->fx[i];
        o->vy[i] += o->dtforce * o->fy[i];
        o->vz[i] += o->dtforce * o->fz[i];
    }
}

/* ------------------------------------------------------------------------------------------ */
/* verletlist/force_lj.c:14-105 computeForceLJFullNeigh */
EXPORT void ovl_force_lj_full(OVL* o)
{
    real cutforcesq = o->cutforce * o->cutforce, sigma6 = o->sigma6, epsilon = o->epsilon;
    const real num1 = 1.0, num48 = 48.0, num05 = 0.5;
    for (int i = 0; i < o->Nlocal; i++) { o->fx[i] = 0.0; o->fy[i] = 0.0; o->fz[i] = 0.0; }
    for (int i = 0; i < o->Nlocal; i++) {
        int* neighs   = &o->neighbors[(size_t)i * o->maxneighs];
        int numneighs = o->numneigh[i];
        real xtmp = o->x[i], ytmp = o->y[i], ztmp = o->z[i], fix = 0, fiy = 0, fiz = 0;
        for (int k = 0; k < numneighs; k++) {
            int j     = neighs[k];
            real delx = xtmp - o->x[j], dely = ytmp - o->y[j], delz = ztmp - o->z[j];
            real rsq  = delx * delx + dely * dely + delz * delz;
            if (rsq < cutforcesq) {
                real sr2   = num1 / rsq;
                real sr6   = sr2 * sr2 * sr2 * sigma6;
                real force = num48 * sr6 * (sr6 - num05) * sr2 * epsilon;
                fix += delx * force; fiy += dely * force; fiz += delz * force;
                o->pairs_in_cutoff++;
            }
        }
        o->fx[i] += fix; o->fy[i] += fiy; o->fz[i] += fiz;
        o->pairs_listed += numneighs;
    }
    o->force_calls++;
}

/* verletlist/force_lj.c:107-198 computeForceLJHalfNeigh (reaction only on local j, 176-180) */
EXPORT void ovl_force_lj_half(OVL* o)
{
    real cutforcesq = o->cutforce * o->cutforce, sigma6 = o->sigma6, epsilon = o->epsilon;
    const real num1 = 1.0, num48 = 48.0, num05 = 0.5;
    int nlocal = o->Nlocal;
    for (int i = 0; i < nlocal; i++) { o->fx[i] = 0.0; o->fy[i] = 0.0; o->fz[i] = 0.0; }
    for (int i = 0; i < nlocal; i++) {
        int* neighs   = &o->neighbors[(size_t)i * o->maxneighs];
        int numneighs = o->numneigh[i];
        real xtmp = o->x[i], ytmp = o->y[i], ztmp = o->z[i], fix = 0, fiy = 0, fiz = 0;
        for (int k = 0; k < numneighs; k++) {
            int j     = neighs[k];
            real delx = xtmp - o->x[j], dely = ytmp - o->y[j], delz = ztmp - o->z[j];
            real rsq  = delx * delx + dely * dely + delz * delz;
            if (rsq < cutforcesq) {
                real sr2   = num1 / rsq;
                real sr6   = sr2 * sr2 * sr2 * sigma6;
                real force = num48 * sr6 * (sr6 - num05) * sr2 * epsilon;
                fix += delx * force; fiy += dely * force; fiz += delz * force;
                if (j < nlocal) {
                    o->fx[j] -= delx * force; o->fy[j] -= dely * force; o->fz[j] -= delz * force;
                }
                o->pairs_in_cutoff++;
            }
        }
        o->fx[i] += fix; o->fy[i] += fiy; o->fz[i] += fiz;
        o->pairs_listed += numneighs;
    }
    o->force_calls++;
}

/* ------------------------------------------------------------------------------------------ */
/* EAM tables.  The reference builds them once on the host (common/eam_utils.c:22-284); the
 * tables are INPUTS to the force passes, so the oracle takes them either from its own
 * restatement of that builder (ovl_eam_from_funcfl) or verbatim from a fixture (ovl_eam_set). */
EXPORT void ovl_eam_set(OVL* o, int nr, int nrho, int nr_tot, int nrho_tot, double rdr,
    double rdrho, const real* rhor_spline, const real* frho_spline, const real* z2r_spline)
{
    o->nr = nr; o->nrho = nrho; o->nr_tot = nr_tot; o->nrho_tot = nrho_tot;
    o->rdr = rdr; o->rdrho = rdrho;
    free(o->rhor_spline); free(o->frho_spline); free(o->z2r_spline);
    o->rhor_spline = (real*)malloc(nr_tot * sizeof(real));
    o->frho_spline = (real*)malloc(nrho_tot * sizeof(real));
    o->z2r_spline  = (real*)malloc(nr_tot * sizeof(real));
    memcpy(o->rhor_spline, rhor_spline, nr_tot * sizeof(real));
    memcpy(o->frho_spline, frho_spline, nrho_tot * sizeof(real));
    memcpy(o->z2r_spline, z2r_spline, nr_tot * sizeof(real));
}

/* common/eam_utils.c:253-284 interpolate */
static void eam_interpolate(int n, real delta, const real* f, real* spline)
{
    for (int m = 1; m <= n; m++) spline[m * 7 + 6] = f[m];
    spline[1 * 7 + 5]       = spline[2 * 7 + 6] - spline[1 * 7 + 6];
    spline[2 * 7 + 5]       = 0.5 * (spline[3 * 7 + 6] - spline[1 * 7 + 6]);
    spline[(n - 1) * 7 + 5] = 0.5 * (spline[n * 7 + 6] - spline[(n - 2) * 7 + 6]);
    spline[n * 7 + 5]       = spline[n * 7 + 6] - spline[(n - 1) * 7 + 6];
    for (int m = 3; m <= n - 2; m++)
        spline[m * 7 + 5] = ((spline[(m - 2) * 7 + 6] - spline[(m + 2) * 7 + 6]) +
                                8.0 * (spline[(m + 1) * 7 + 6] - spline[(m - 1) * 7 + 6])) /
                            12.0;
    for (int m = 1; m <= n - 1; m++) {
        spline[m * 7 + 4] = 3.0 * (spline[(m + 1) * 7 + 6] - spline[m * 7 + 6]) -
                            2.0 * spline[m * 7 + 5] - spline[(m + 1) * 7 + 5];
        spline[m * 7 + 3] = spline[m * 7 + 5] + spline[(m + 1) * 7 + 5] -
                            2.0 * (spline[(m + 1) * 7 + 6] - spline[m * 7 + 6]);
    }
    spline[n * 7 + 4] = 0.0;
    spline[n * 7 + 3] = 0.0;
    for (int m = 1; m <= n; m++) {
        spline[m * 7 + 2] = spline[m * 7 + 5] / delta;
        spline[m * 7 + 1] = 2.0 * spline[m * 7 + 4] / delta;
        spline[m * 7 + 0] = 3.0 * spline[m * 7 + 3] / delta;
    }
}

/* 4-point Lagrange regrid used three times in common/eam_utils.c:95-220 file2array */
static double eam_lagrange(const real* tab, int ntab, double dtab, double r)
{
    const double sixth = 1.0 / 6.0;
    double p = r / dtab + 1.0;
    int k    = (int)(p);
    if (k > ntab - 2) k = ntab - 2;
    if (k < 2) k = 2;
    p -= k;
    if (p > 2.0) p = 2.0;
    double cof1 = -sixth * p * (p - 1.0) * (p - 2.0);
    double cof2 = 0.5 * (p * p - 1.0) * (p - 2.0);
    double cof3 = -0.5 * p * (p + 1.0) * (p - 2.0);
    double cof4 = sixth * p * (p * p - 1.0);
    return cof1 * tab[k - 1] + cof2 * tab[k] + cof3 * tab[k + 1] + cof4 * tab[k + 2];
}

/* funcfl tables (0-based as in the file: frho[nrho], zr[nr], rhor[nr]) -> splines.
 * Restates readEamFile's 1-shift (eam_utils.c:85-90), file2array (95-220), array2spline (222-251)
 * and initEam's parameter overrides (22-40). */
EXPORT void ovl_eam_from_funcfl(OVL* o, int nrho, double drho_, int nr, double dr_, double cut_,
    double mass_, const real* frho0, const real* zr0, const real* rhor0)
{
    real fdrho = drho_, fdr = dr_, fcut = cut_, fmass = mass_; /* Funcfl fields are MD_FLOAT */
    o->mass     = fmass;
    o->cutforce = fcut;
    o->cutneigh = o->cutforce + 1.0;
    o->temp     = 600.0;
    o->dt       = 0.001;
    o->rho      = 0.07041125;
    o->dtforce  = 0.5 * o->dt / o->mass;
    real* frho = (real*)calloc(nrho + 1, sizeof(real));
    real* rhor = (real*)calloc(nr + 1, sizeof(real));
    real* zr   = (real*)calloc(nr + 1, sizeof(real));
    for (int i = nrho; i > 0; i--) frho[i] = frho0[i - 1];
    for (int i = nr; i > 0; i--) { rhor[i] = rhor0[i - 1]; zr[i] = zr0[i - 1]; }
    real edr = 0.0, edrho = 0.0;
    double rmax = 0.0, rhomax = 0.0;
    if (fdr > edr) edr = fdr;
    if (fdrho > edrho) edrho = fdrho;
    if ((nr - 1) * fdr > rmax) rmax = (nr - 1) * fdr;
    if ((nrho - 1) * fdrho > rhomax) rhomax = (nrho - 1) * fdrho;
    int enr   = (int)(rmax / edr + 0.5);
    int enrho = (int)(rhomax / edrho + 0.5);
    real* afrho = (real*)calloc(enrho + 1, sizeof(real));
    real* arhor = (real*)calloc(enr + 1, sizeof(real));
    real* az2r  = (real*)calloc(enr + 1, sizeof(real));
    for (int m = 1; m <= enrho; m++) afrho[m] = eam_lagrange(frho, nrho, fdrho, (m - 1) * edrho);
    for (int m = 1; m <= enr; m++) arhor[m] = eam_lagrange(rhor, nr, fdr, (m - 1) * edr);
    for (int m = 1; m <= enr; m++) {
        double r   = (m - 1) * edr;
        double zri = eam_lagrange(zr, nr, fdr, r), zrj = eam_lagrange(zr, nr, fdr, r);
        az2r[m]    = 27.2 * 0.529 * zri * zrj;
    }
    o->rdr = 1.0 / edr; o->rdrho = 1.0 / edrho;
    o->nr = enr; o->nrho = enrho;
    o->nrho_tot = (enrho + 1) * 7 + 64; o->nr_tot = (enr + 1) * 7 + 64;
    o->nrho_tot -= o->nrho_tot % 64; o->nr_tot -= o->nr_tot % 64;
    free(o->rhor_spline); free(o->frho_spline); free(o->z2r_spline);
    o->frho_spline = (real*)calloc(o->nrho_tot, sizeof(real));
    o->rhor_spline = (real*)calloc(o->nr_tot, sizeof(real));
    o->z2r_spline  = (real*)calloc(o->nr_tot, sizeof(real));
    eam_interpolate(enrho, edrho, afrho, o->frho_spline);
    eam_interpolate(enr, edr, arhor, o->rhor_spline);
    eam_interpolate(enr, edr, az2r, o->z2r_spline);
    free(frho); free(rhor); free(zr); free(afrho); free(arhor); free(az2r);
}

/* verletlist/force_eam.c:19-231 computeForceEam: density + embedding derivative, ghost fp copy,
 * pair force.  Full neighbor lists only (force.c:16-18 ignores half_neigh for EAM). */
EXPORT void ovl_force_eam(OVL* o)
{
    if (o->fp_nmax < o->Nmax) {
        o->fp_nmax = o->Nmax;
        free(o->fp);
        o->fp = (real*)calloc(o->Nmax, sizeof(real));
    }
    int Nlocal = o->Nlocal, nr = o->nr, nrho = o->nrho;
    real rdr = o->rdr, rdrho = o->rdrho, cutforcesq = o->cutforce * o->cutforce;
    const real *rs = o->rhor_spline, *fs = o->frho_spline, *zs = o->z2r_spline;
    real* fp = o->fp;
    for (int i = 0; i < Nlocal; i++) {
        int* neighs   = &o->neighbors[(size_t)i * o->maxneighs];
        int numneighs = o->numneigh[i];
        real xtmp = o->x[i], ytmp = o->y[i], ztmp = o->z[i], rhoi = 0;
        for (int k = 0; k < numneighs; k++) {
            int j     = neighs[k];
            real delx = xtmp - o->x[j], dely = ytmp - o->y[j], delz = ztmp - o->z[j];
            real rsq  = delx * delx + dely * dely + delz * delz;
            if (rsq < cutforcesq) {
                real p = RSQRT(rsq) * rdr + 1.0;
                int m  = (int)(p);
                m      = m < nr - 1 ? m : nr - 1;
                p -= m;
                p = p < 1.0 ? p : 1.0;
                rhoi += ((rs[m * 7 + 3] * p + rs[m * 7 + 4]) * p + rs[m * 7 + 5]) * p + rs[m * 7 + 6];
            }
        }
        real p = 1.0 * rhoi * rdrho + 1.0;
        int m  = (int)(p);
        m      = m < nrho - 1 ? m : nrho - 1;
        m      = m > 1 ? m : 1;
        p -= m;
        p     = p < 1.0 ? p : 1.0;
        fp[i] = (fs[m * 7 + 0] * p + fs[m * 7 + 1]) * p + fs[m * 7 + 2];
    }
    for (int i = 0; i < o->Nghost; i++) fp[Nlocal + i] = fp[o->border_map[i]];
    for (int i = 0; i < Nlocal; i++) {
        int* neighs   = &o->neighbors[(size_t)i * o->maxneighs];
        int numneighs = o->numneigh[i];
        real xtmp = o->x[i], ytmp = o->y[i], ztmp = o->z[i], fix = 0, fiy = 0, fiz = 0;
        for (int k = 0; k < numneighs; k++) {
            int j     = neighs[k];
            real delx = xtmp - o->x[j], dely = ytmp - o->y[j], delz = ztmp - o->z[j];
            real rsq  = delx * delx + dely * dely + delz * delz;
            if (rsq < cutforcesq) {
                real r = RSQRT(rsq);
                real p = r * rdr + 1.0;
                int m  = (int)(p);
                m      = m < nr - 1 ? m : nr - 1;
                p -= m;
                p          = p < 1.0 ? p : 1.0;
                real rhoip = (rs[m * 7 + 0] * p + rs[m * 7 + 1]) * p + rs[m * 7 + 2];
                real z2p   = (zs[m * 7 + 0] * p + zs[m * 7 + 1]) * p + zs[m * 7 + 2];
                real z2 = ((zs[m * 7 + 3] * p + zs[m * 7 + 4]) * p + zs[m * 7 + 5]) * p + zs[m * 7 + 6];
                real recip = 1.0 / r;
                real phi   = z2 * recip;
                real phip  = z2p * recip - phi * recip;
                real psip  = fp[i] * rhoip + fp[j] * rhoip + phip;
                real fpair = -psip * recip;
                fix += delx * fpair; fiy += dely * fpair; fiz += delz * fpair;
                o->pairs_in_cutoff++;
            }
        }
        o->fx[i] = fix; o->fy[i] = fiy; o->fz[i] = fiz;
        o->pairs_listed += numneighs;
    }
    o->force_calls++;
}

/* verletlist/force.c:13-34 initForce dispatch */
EXPORT void ovl_compute_force(OVL* o)
{
    if (o->force_field == 1) ovl_force_eam(o);
    else if (o->half_neigh) ovl_force_lj_half(o);
    else ovl_force_lj_full(o);
}

/* ------------------------------------------------------------------------------------------ */
/* verletlist/main.c:46-73 setup() after parameters are final (atoms created by the caller when
 * they come from a file) */
EXPORT void ovl_setup(OVL* o, int create)
{
    ovl_derive(o);
    if (create) ovl_create_atoms(o);
    ovl_setup_neighbor(o);
    ovl_setup_thermo(o);
    if (create) ovl_adjust_thermo(o);
    ovl_setup_pbc(o);
    ovl_update_pbc(o);
    ovl_build_neighbor(o);
}

/* verletlist/main.c:76-95 reneighbour (SORT_ATOMS off) */
EXPORT void ovl_reneighbour(OVL* o)
{
    ovl_update_atoms_pbc(o);
    ovl_setup_pbc(o);
    ovl_update_pbc(o);
    ovl_build_neighbor(o);
}

/* verletlist/main.c:244-288: thermo(0), first force, time loop, thermo(-1).
 * thermo_out receives (step, T, P) triples; returns the number of records. */
EXPORT int ovl_run(OVL* o, int nsteps, double* thermo_out, int max_records)
{
    int nrec = 0;
    double T, P;
    ovl_thermo(o, &T, &P);
    if (nrec < max_records) { thermo_out[3 * nrec] = 0; thermo_out[3 * nrec + 1] = T; thermo_out[3 * nrec + 2] = P; nrec++; }
    ovl_compute_force(o);
    for (int n = 0; n < nsteps; n++) {
        int reneigh = (n + 1) % o->reneigh_every == 0;
        ovl_initial_integrate(o);
        if (reneigh) ovl_reneighbour(o); else ovl_update_pbc(o);
        ovl_compute_force(o);
        ovl_final_integrate(o);
        if (!((n + 1) % o->nstat) && (n + 1) < nsteps) {
            ovl_thermo(o, &T, &P);
            if (nrec < max_records) { thermo_out[3 * nrec] = n + 1; thermo_out[3 * nrec + 1] = T; thermo_out[3 * nrec + 2] = P; nrec++; }
        }
    }
    ovl_thermo(o, &T, &P);
    if (nrec < max_records) { thermo_out[3 * nrec] = nsteps; thermo_out[3 * nrec + 1] = T; thermo_out[3 * nrec + 2] = P; nrec++; }
    return nrec;
}

/* ------------------------------------------------------------------------------------------ */
/* accessors for ctypes */
EXPORT int ovl_get_int(OVL* o, const char* k)
{
#define K(n) if (!strcmp(k, #n)) return o->n;
    K(Natoms) K(Nlocal) K(Nghost) K(Nmax) K(maxneighs) K(nstencil) K(mbins) K(atoms_per_bin)
    K(nbinx) K(nbiny) K(nbinz) K(mbinx) K(mbiny) K(mbinz) K(mbinxlo) K(mbinylo) K(mbinzlo)
    K(nr) K(nrho) K(nr_tot) K(nrho_tot) K(ntimes) K(nstat) K(reneigh_every) K(half_neigh)
#undef K
    return -1;
}
EXPORT double ovl_get_real(OVL* o, const char* k)
{
#define K(n) if (!strcmp(k, #n)) return o->n;
    K(xprd) K(yprd) K(zprd) K(lattice) K(cutneigh) K(cutneighsq) K(cutforce) K(bininvx) K(bininvy)
    K(bininvz) K(binsizex) K(binsizey) K(binsizez) K(dt) K(dtforce) K(mass) K(temp) K(rho) K(rdr)
    K(rdrho) K(sigma6) K(epsilon) K(t_scale) K(p_scale) K(dof_boltz)
#undef K
    return NAN;
}
EXPORT long long ovl_get_counter(OVL* o, int which)
{
    return which == 0 ? o->pairs_listed : which == 1 ? o->pairs_in_cutoff : o->force_calls;
}
EXPORT void* ovl_ptr(OVL* o, const char* k)
{
#define K(n) if (!strcmp(k, #n)) return (void*)o->n;
    K(x) K(y) K(z) K(vx) K(vy) K(vz) K(fx) K(fy) K(fz) K(type) K(border_map) K(PBCx) K(PBCy) K(PBCz)
    K(bincount) K(bins) K(stencil) K(numneigh) K(neighbors) K(rhor_spline) K(frho_spline)
    K(z2r_spline) K(fp)
#undef K
    return NULL;
}
EXPORT int ovl_precision(void) { return PRECISION; }
