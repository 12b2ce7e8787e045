/*
 * oracle/cp_oracle.c -- TEST INFRASTRUCTURE ONLY (parity oracle; never shipped, never timed as the
 * product).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
 *
 * Plain-C, single-threaded restatement of MD-Bench's CLUSTERPAIR scheme (GROMACS-style MxN cluster
 * pairs, reference src/clusterpair/), written from the reference's behaviour with file:line cited per
 * function (paths relative to the reference root).  M = 4 i-atoms per cluster like every CPU build of
 * the reference (force.h:48); N = 4 or 8 is a RUN-TIME field here (the reference fixes it at compile
 * time through VECTOR_WIDTH, force.h:50-58), so the index macros of force.h:62-91 are functions below.
 *
 * PINNING (tests/test_cp_oracle_pinned.py): cluster membership and order, bounding boxes, ghost
 * clusters, bins and cluster-pair lists are compared bit for bit with
 *   - the UNMODIFIED reference clusterpair libraries (AVX-512 builds, 4x8; oracle/_ref/libmdref_cp_*),
 *   - the reference with its scalar kernel computeForceLJRef enabled for M = 4 (USE_REFERENCE_VERSION
 *     plus the one-line CLUSTER_M 1 -> 4 change SURVEY 8c describes; oracle/Makefile target ref-cpref),
 *     which also pins forces and trajectories (exact division; the SIMD kernels use rcp14).
 * Compiled with -ffp-contract=off; fused multiply-adds of the reference build are explicit where bits
 * decide membership (objdump of oracle/_ref: atomDistanceInRange = fma(dz,dz,fma(dx,dx,dy*dy))).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#if PRECISION == 1
typedef float real;
#define RFMA fmaf
#define RSQRT sqrtf
#define RCBRT cbrtf
#define RCEIL ceilf
#else
typedef double real;
#define RFMA fma
#define RSQRT sqrt
#define RCBRT cbrt
#define RCEIL ceil
#endif

#define EXPORT __attribute__((visibility("default")))
#define CLUSTER_M 4
#define MAXV(a, b) ((a) > (b) ? (a) : (b))
#define MINV(a, b) ((a) < (b) ? (a) : (b))
#define DELTA 20000

typedef struct {
    int natoms;
    real bbminx, bbmaxx, bbminy, bbmaxy, bbminz, bbmaxz;
} Cluster; /* clusterpair/atom.h:19-24 */

typedef struct {
    int N, vector_width; /* CLUSTER_N, VECTOR_WIDTH (only used for the dummy padding of the lists) */
    /* parameters (common/parameter.h) */
    real epsilon, sigma, sigma6, temp, rho, mass, dt, dtforce, skin, cutforce, cutneigh, lattice;
    int ntimes, nstat, reneigh_every, half_neigh, nx, ny, nz;
    real xprd, yprd, zprd;
    int from_input; /* box lengths handed over by a reader (param->input_file != NULL, neighbor.c:78-82) */
    /* atoms (clusterpair/atom.h:26-60) */
    int Natoms, Nlocal, Nghost, Nmax;
    int Nclusters, Nclusters_local, Nclusters_ghost, Nclusters_max;
    real *x, *y, *z, *vx, *vy, *vz;
    int* tag; /* not in the reference: original index of each atom, carried through the permutations */
    int *border_map, *PBCx, *PBCy, *PBCz, NmaxGhost;
    real *cl_x, *cl_v, *cl_f;
    int* cl_tag;
    Cluster *iclusters, *jclusters;
    int* icluster_bin;
    int dummy_cj;
    /* neighbor statics (clusterpair/neighbor.c:26-45) */
    real bininvx, bininvy, binsizex, binsizey, cutneighsq;
    int mbinxlo, mbinylo, nbinx, nbiny, mbinx, mbiny, mbins, atoms_per_bin, clusters_per_bin, nstencil, nmax;
    int *bincount, *bins, *bin_nclusters, *bin_clusters, *stencil;
    int maxneighs, *numneigh, *numneigh_masked, *neighbors;
    /* thermo (common/thermo.c) */
    real t_scale, p_scale, dof_boltz;
} OCP;

/* ---- index macros of clusterpair/force.h:62-91 with CLUSTER_N at run time (M = 4, N = 4 or 8) ---- */
static inline int cj0_from_ci(const OCP* o, int ci) { return o->N == CLUSTER_M ? ci : ci >> 1; }
static inline int cj1_from_ci(const OCP* o, int ci) { return o->N == CLUSTER_M ? ci : ci >> 1; }
static inline int ci_base(const OCP* o, int ci, int b)
{
    return o->N == CLUSTER_M ? ci * o->N * b : (ci >> 1) * o->N * b + (ci & 1) * (o->N >> 1);
}
static inline int cj_base(const OCP* o, int cj, int b) { return cj * o->N * b; }
#define W(o) ((o)->N) /* CL_{X,Y,Z}_OFFSET = {0,1,2} * max(M,N) = N here */

static void* xrealloc(void* p, size_t n)
{
    void* q = realloc(p, n ? n : 1);
    if (!q) { fprintf(stderr, "cp_oracle: out of memory\n"); exit(1); }
    return q;
}

EXPORT OCP* ocp_new(int N, int vector_width)
{
    OCP* o = (OCP*)calloc(1, sizeof(OCP));
    o->N = N; o->vector_width = vector_width > 0 ? vector_width : N;
    o->epsilon = 1.0; o->sigma = 1.0; o->sigma6 = 1.0; o->rho = 0.8442; o->ntimes = 200; o->dt = 0.005;
    o->nx = o->ny = o->nz = 32; o->cutforce = 2.5; o->skin = 0.3; o->cutneigh = o->cutforce + o->skin;
    o->temp = 1.44; o->nstat = 100; o->mass = 1.0; o->dtforce = 0.5 * o->dt; o->reneigh_every = 20;
    o->atoms_per_bin = 8;                              /* neighbor.c:57 */
    o->clusters_per_bin = (8 / CLUSTER_M) + 10;        /* neighbor.c:58 */
    o->maxneighs = 100;                                /* neighbor.c:65 */
    return o;
}
EXPORT void ocp_free(OCP* o)
{
    void* p[] = { o->x, o->y, o->z, o->vx, o->vy, o->vz, o->tag, o->border_map, o->PBCx, o->PBCy, o->PBCz, o->cl_x,
        o->cl_v, o->cl_f, o->cl_tag, o->iclusters, o->jclusters, o->icluster_bin, o->bincount, o->bins,
        o->bin_nclusters, o->bin_clusters, o->stencil, o->numneigh, o->numneigh_masked, o->neighbors };
    for (unsigned i = 0; i < sizeof p / sizeof p[0]; i++) free(p[i]);
    free(o);
}
EXPORT void ocp_set_lj(OCP* o, double epsilon, double sigma, double cutforce, double skin, double dt, double temp,
    double rho, double mass)
{
    o->epsilon = epsilon; o->sigma = sigma;
    real s2 = o->sigma * o->sigma; /* parameter.c:118-119 */
    o->sigma6 = s2 * s2 * s2;
    o->cutforce = cutforce; o->skin = skin; o->dt = dt; o->dtforce = 0.5 * o->dt;
    o->temp = temp; o->rho = rho; o->mass = mass;
    o->cutneigh = o->cutforce + o->skin; /* clusterpair/main.c:216 */
}
EXPORT void ocp_set_run(OCP* o, int nx, int ny, int nz, int ntimes, int nstat, int reneigh_every, int half_neigh)
{
    o->nx = nx; o->ny = ny; o->nz = nz; o->ntimes = ntimes; o->nstat = nstat; o->reneigh_every = reneigh_every;
    o->half_neigh = half_neigh;
}
/* box of an input file: readAtom sets param->xprd/yprd/zprd = hi - lo (atom.c readers); setupNeighbor then takes them
 * instead of nx * lattice (neighbor.c:78-82) and treats the box as [0, prd) */
EXPORT void ocp_set_box(OCP* o, double xprd, double yprd, double zprd)
{
    o->from_input = 1;
    o->xprd = xprd; o->yprd = yprd; o->zprd = zprd;
}
/* the atoms createAtom + adjustThermo produced (identical code in both schemes, atom.c:49-180; generated by
 * the verletlist oracle in the tests) */
EXPORT void ocp_set_atoms(OCP* o, int n, const real* x, const real* y, const real* z, const real* vx, const real* vy,
    const real* vz)
{
    o->Natoms = o->Nlocal = n; o->Nmax = n;
    real** a[] = { &o->x, &o->y, &o->z, &o->vx, &o->vy, &o->vz };
    const real* s[] = { x, y, z, vx, vy, vz };
    for (int k = 0; k < 6; k++) {
        *a[k] = (real*)xrealloc(*a[k], n * sizeof(real));
        memcpy(*a[k], s[k], n * sizeof(real));
    }
    o->tag = (int*)xrealloc(o->tag, n * sizeof(int));
    for (int i = 0; i < n; i++) o->tag[i] = i;
}

/* clusterpair/atom.c:643-677 growClusters */
static void grow_clusters(OCP* o)
{
    const int nold = o->Nclusters_max;
    o->Nclusters_max += DELTA;
    const size_t nc = o->Nclusters_max;
    o->iclusters    = (Cluster*)xrealloc(o->iclusters, nc * sizeof(Cluster));
    o->jclusters    = (Cluster*)xrealloc(o->jclusters, nc * sizeof(Cluster));
    o->icluster_bin = (int*)xrealloc(o->icluster_bin, nc * sizeof(int));
    o->cl_x   = (real*)xrealloc(o->cl_x, nc * CLUSTER_M * 3 * sizeof(real));
    o->cl_f   = (real*)xrealloc(o->cl_f, nc * CLUSTER_M * 3 * sizeof(real));
    o->cl_v   = (real*)xrealloc(o->cl_v, nc * CLUSTER_M * 3 * sizeof(real));
    o->cl_tag = (int*)xrealloc(o->cl_tag, nc * CLUSTER_M * sizeof(int));
    for (size_t k = (size_t)nold * CLUSTER_M; k < nc * CLUSTER_M; k++) o->cl_tag[k] = -1;
    memset(o->cl_f + (size_t)nold * CLUSTER_M * 3, 0, (nc - nold) * CLUSTER_M * 3 * sizeof(real));
}

/* ------------------------------------------------------------------------------------------ */
/* clusterpair/neighbor.c:565-580 bindist, 582-618 coord2bin / coord2bin2D                     */
static real bindist(const OCP* o, int i, int j)
{
    real delx = i > 0 ? (i - 1) * o->binsizex : (i == 0 ? (real)0.0 : (i + 1) * o->binsizex);
    real dely = j > 0 ? (j - 1) * o->binsizey : (j == 0 ? (real)0.0 : (j + 1) * o->binsizey);
    return (delx * delx + dely * dely);
}
static void coord2bin2D(const OCP* o, real xin, real yin, int* ix, int* iy)
{
    if (xin >= o->xprd) *ix = (int)((xin - o->xprd) * o->bininvx) + o->nbinx - o->mbinxlo;
    else if (xin >= (real)0.0) *ix = (int)(xin * o->bininvx) - o->mbinxlo;
    else *ix = (int)(xin * o->bininvx) - o->mbinxlo - 1;
    if (yin >= o->yprd) *iy = (int)((yin - o->yprd) * o->bininvy) + o->nbiny - o->mbinylo;
    else if (yin >= (real)0.0) *iy = (int)(yin * o->bininvy) - o->mbinylo;
    else *iy = (int)(yin * o->bininvy) - o->mbinylo - 1;
}
static int coord2bin(const OCP* o, real xin, real yin)
{
    int ix, iy;
    coord2bin2D(o, xin, yin, &ix, &iy);
    return iy * o->mbinx + ix + 1;
}

/* clusterpair/main.c:46-49 + neighbor.c:70-172 setupNeighbor */
EXPORT void ocp_setup_neighbor(OCP* o)
{
    const real SMALL = 1.0e-6, FACTOR = 0.999;
    o->lattice = pow((4.0 / o->rho), (1.0 / 3.0));
    if (!o->from_input) { o->xprd = o->nx * o->lattice; o->yprd = o->ny * o->lattice; o->zprd = o->nz * o->lattice; }
    const real xlo = 0.0, xhi = o->xprd, ylo = 0.0, yhi = o->yprd, zlo = 0.0, zhi = o->zprd;
    /* neighbor.c:93-98 as the reference BUILD evaluates it (objdump of oracle/_ref, -Ofast): the two divisions
     * and the cbrt are folded into  nbin = ceil(prd * cbrt(density * (1/atoms_in_cell)))  with everything in
     * MD_FLOAT (cbrtf and a float ceil in the SP build).  On the generated lattices prd / targetsize is an
     * INTEGER mathematically (4x4: exactly nx), so the rounding of this very expression decides the bin count. */
    real atom_density   = ((real)(o->Nlocal)) / ((o->xprd * o->yprd) * o->zprd);
    real inv_targetsize = RCBRT(atom_density * ((real)1.0 / (real)MAXV(CLUSTER_M, o->N)));
    o->nbinx = MAXV(1, (int)RCEIL(o->xprd * inv_targetsize));
    o->nbiny = MAXV(1, (int)RCEIL(o->yprd * inv_targetsize));
    o->binsizex = (xhi - xlo) / o->nbinx;
    o->binsizey = (yhi - ylo) / o->nbiny;
    o->bininvx  = 1.0 / o->binsizex;
    o->bininvy  = 1.0 / o->binsizey;
    o->cutneighsq = o->cutneigh * o->cutneigh;
    real coord;
    int mbinxhi, mbinyhi;
    coord = xlo - o->cutneigh - SMALL * o->xprd;
    o->mbinxlo = (int)(coord * o->bininvx);
    if (coord < 0.0) o->mbinxlo -= 1;
    coord = xhi + o->cutneigh + SMALL * o->xprd;
    mbinxhi = (int)(coord * o->bininvx);
    coord = ylo - o->cutneigh - SMALL * o->yprd;
    o->mbinylo = (int)(coord * o->bininvy);
    if (coord < 0.0) o->mbinylo -= 1;
    coord = yhi + o->cutneigh + SMALL * o->yprd;
    mbinyhi = (int)(coord * o->bininvy);
    o->mbinxlo -= 1; mbinxhi += 1; o->mbinx = mbinxhi - o->mbinxlo + 1;
    o->mbinylo -= 1; mbinyhi += 1; o->mbiny = mbinyhi - o->mbinylo + 1;
    int nextx = (int)(o->cutneigh * o->bininvx), nexty = (int)(o->cutneigh * o->bininvy);
    if (nextx * o->binsizex < FACTOR * o->cutneigh) nextx++;
    if (nexty * o->binsizey < FACTOR * o->cutneigh) nexty++;
    o->stencil  = (int*)xrealloc(o->stencil, (2 * nexty + 1) * (2 * nextx + 1) * sizeof(int));
    o->nstencil = 0;
    for (int j = -nexty; j <= nexty; j++)
        for (int i = -nextx; i <= nextx; i++)
            if (bindist(o, i, j) < o->cutneighsq) o->stencil[o->nstencil++] = j * o->mbinx + i;
    o->mbins         = o->mbinx * o->mbiny;
    o->bincount      = (int*)xrealloc(o->bincount, o->mbins * sizeof(int));
    o->bins          = (int*)xrealloc(o->bins, (size_t)o->mbins * o->atoms_per_bin * sizeof(int));
    o->bin_nclusters = (int*)xrealloc(o->bin_nclusters, o->mbins * sizeof(int));
    o->bin_clusters  = (int*)xrealloc(o->bin_clusters, (size_t)o->mbins * o->clusters_per_bin * sizeof(int));
    /* common/thermo.c:30-53 setupThermo (LJ) */
    o->dof_boltz = (real)(o->Natoms * 3 - 3);
    o->t_scale   = (real)1.0 / o->dof_boltz;
    o->p_scale   = 1.0 / 3 / o->xprd / o->yprd / o->zprd;
}

/* neighbor.c:599-630 binAtoms, 632-661 sortAtomsByZCoord (selection sort with strict < and swap: the
 * permutation among equal z is NOT that of a stable sort -- reproduced literally) */
static void bin_atoms(OCP* o)
{
    int resize = 1;
    while (resize > 0) {
        resize = 0;
        for (int i = 0; i < o->mbins; i++) o->bincount[i] = 0;
        for (int i = 0; i < o->Nlocal; i++) {
            int ibin = coord2bin(o, o->x[i], o->y[i]);
            if (o->bincount[ibin] < o->atoms_per_bin) {
                int ac = o->bincount[ibin]++;
                o->bins[(size_t)ibin * o->atoms_per_bin + ac] = i;
            } else resize = 1;
        }
        if (resize) {
            o->atoms_per_bin *= 2;
            o->bins = (int*)xrealloc(o->bins, (size_t)o->mbins * o->atoms_per_bin * sizeof(int));
        }
    }
}
static void sort_atoms_by_z(OCP* o)
{
    for (int bin = 0; bin < o->mbins; bin++) {
        int c = o->bincount[bin];
        int* bin_ptr = &o->bins[(size_t)bin * o->atoms_per_bin];
        for (int ac_i = 0; ac_i < c; ac_i++) {
            int i = bin_ptr[ac_i], min_ac = ac_i, min_idx = i;
            real min_z = o->z[i];
            for (int ac_j = ac_i + 1; ac_j < c; ac_j++) {
                int j = bin_ptr[ac_j];
                real zj = o->z[j];
                if (zj < min_z) { min_ac = ac_j; min_idx = j; min_z = zj; }
            }
            bin_ptr[ac_i]   = min_idx;
            bin_ptr[min_ac] = i;
        }
    }
}
/* neighbor.c:663-753 buildClusters */
EXPORT void ocp_build_clusters(OCP* o)
{
    const int N = o->N;
    o->Nclusters_local = 0;
    bin_atoms(o);
    sort_atoms_by_z(o);
    for (int bin = 0; bin < o->mbins; bin++) {
        int c = o->bincount[bin], ac = 0;
        int nclusters = ((c + CLUSTER_M - 1) / CLUSTER_M);
        if (N > CLUSTER_M && nclusters % 2) nclusters++;
        for (int cl = 0; cl < nclusters; cl++) {
            const int ci = o->Nclusters_local;
            if (ci >= o->Nclusters_max) grow_clusters(o);
            real* ci_x = &o->cl_x[ci_base(o, ci, 3)];
            real* ci_v = &o->cl_v[ci_base(o, ci, 3)];
            int* ci_t  = &o->cl_tag[ci_base(o, ci, 1)];
            real bbminx = INFINITY, bbmaxx = -INFINITY, bbminy = INFINITY, bbmaxy = -INFINITY, bbminz = INFINITY,
                 bbmaxz = -INFINITY;
            o->iclusters[ci].natoms = 0;
            for (int cii = 0; cii < CLUSTER_M; cii++) {
                if (ac < c) {
                    int i = o->bins[(size_t)bin * o->atoms_per_bin + ac];
                    real xtmp = o->x[i], ytmp = o->y[i], ztmp = o->z[i];
                    ci_x[0 * W(o) + cii] = xtmp; ci_x[1 * W(o) + cii] = ytmp; ci_x[2 * W(o) + cii] = ztmp;
                    ci_v[0 * W(o) + cii] = o->vx[i]; ci_v[1 * W(o) + cii] = o->vy[i]; ci_v[2 * W(o) + cii] = o->vz[i];
                    if (bbminx > xtmp) bbminx = xtmp;
                    if (bbmaxx < xtmp) bbmaxx = xtmp;
                    if (bbminy > ytmp) bbminy = ytmp;
                    if (bbmaxy < ytmp) bbmaxy = ytmp;
                    if (bbminz > ztmp) bbminz = ztmp;
                    if (bbmaxz < ztmp) bbmaxz = ztmp;
                    ci_t[cii] = o->tag[i];
                    o->iclusters[ci].natoms++;
                } else {
                    ci_x[0 * W(o) + cii] = INFINITY; ci_x[1 * W(o) + cii] = INFINITY; ci_x[2 * W(o) + cii] = INFINITY;
                    ci_t[cii] = -1;
                }
                ac++;
            }
            o->icluster_bin[ci] = bin;
            Cluster* q = &o->iclusters[ci];
            q->bbminx = bbminx; q->bbmaxx = bbmaxx; q->bbminy = bbminy; q->bbmaxy = bbmaxy; q->bbminz = bbminz; q->bbmaxz = bbmaxz;
            o->Nclusters_local++;
        }
    }
}
/* neighbor.c:755-873 defineJClusters (M == N and 2M == N branches) */
EXPORT void ocp_define_jclusters(OCP* o)
{
    for (int ci = 0; ci < o->Nclusters_local; ci++) {
        int cj0 = cj0_from_ci(o, ci);
        if (o->N == CLUSTER_M) {
            o->jclusters[cj0] = o->iclusters[ci];
        } else if (ci % 2 == 0) {
            const Cluster *a = &o->iclusters[ci], *b = &o->iclusters[ci + 1];
            Cluster* j = &o->jclusters[cj0];
            j->bbminx = MINV(a->bbminx, b->bbminx); j->bbmaxx = MAXV(a->bbmaxx, b->bbmaxx);
            j->bbminy = MINV(a->bbminy, b->bbminy); j->bbmaxy = MAXV(a->bbmaxy, b->bbmaxy);
            j->bbminz = MINV(a->bbminz, b->bbminz); j->bbmaxz = MAXV(a->bbmaxz, b->bbmaxz);
            j->natoms = a->natoms + b->natoms;
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* clusterpair/pbc.c:45-114 updatePbcCPU: ghost j-cluster = image of a local j-cluster.  The shifted
 * coordinate is ONE fma in the reference build (like verletlist). */
EXPORT void ocp_update_pbc(OCP* o, int firstUpdate)
{
    const int N = o->N, jfac = MAXV(1, N / CLUSTER_M), ncj = o->Nclusters_local / jfac;
    for (int cg = 0; cg < o->Nclusters_ghost; cg++) {
        const int cj = ncj + cg;
        real* cjX   = &o->cl_x[cj_base(o, cj, 3)];
        real* bmapX = &o->cl_x[cj_base(o, o->border_map[cg], 3)];
        real bbminx = INFINITY, bbmaxx = -INFINITY, bbminy = INFINITY, bbmaxy = -INFINITY, bbminz = INFINITY,
             bbmaxz = -INFINITY;
        for (int cjj = 0; cjj < o->jclusters[cj].natoms; cjj++) {
            real xtmp = RFMA((real)o->PBCx[cg], o->xprd, bmapX[0 * W(o) + cjj]);
            real ytmp = RFMA((real)o->PBCy[cg], o->yprd, bmapX[1 * W(o) + cjj]);
            real ztmp = RFMA((real)o->PBCz[cg], o->zprd, bmapX[2 * W(o) + cjj]);
            cjX[0 * W(o) + cjj] = xtmp; cjX[1 * W(o) + cjj] = ytmp; cjX[2 * W(o) + cjj] = ztmp;
            if (firstUpdate) {
                if (bbminx > xtmp) bbminx = xtmp;
                if (bbmaxx < xtmp) bbmaxx = xtmp;
                if (bbminy > ytmp) bbminy = ytmp;
                if (bbmaxy < ytmp) bbmaxy = ytmp;
                if (bbminz > ztmp) bbminz = ztmp;
                if (bbmaxz < ztmp) bbmaxz = ztmp;
            }
        }
        if (firstUpdate) {
            for (int cjj = o->jclusters[cj].natoms; cjj < N; cjj++) {
                cjX[0 * W(o) + cjj] = INFINITY; cjX[1 * W(o) + cjj] = INFINITY; cjX[2 * W(o) + cjj] = INFINITY;
            }
            Cluster* q = &o->jclusters[cj];
            q->bbminx = bbminx; q->bbmaxx = bbmaxx; q->bbminy = bbminy; q->bbmaxy = bbmaxy; q->bbminz = bbminz; q->bbmaxz = bbmaxz;
        }
    }
}
/* pbc.c:117-144 updateAtomsPbcCPU */
EXPORT void ocp_update_atoms_pbc(OCP* o)
{
    for (int i = 0; i < o->Nlocal; i++) {
        if (o->x[i] < 0.0) o->x[i] += o->xprd; else if (o->x[i] >= o->xprd) o->x[i] -= o->xprd;
        if (o->y[i] < 0.0) o->y[i] += o->yprd; else if (o->y[i] >= o->yprd) o->y[i] -= o->yprd;
        if (o->z[i] < 0.0) o->z[i] += o->zprd; else if (o->z[i] >= o->zprd) o->z[i] -= o->zprd;
    }
}
/* pbc.c:150-323 setupPbc (ADDGHOST 150-166): ghost clusters by bounding box, 6 faces -> 8 corners -> 12 edges */
EXPORT void ocp_setup_pbc(OCP* o)
{
    static const signed char img[26][3] = { { +1, 0, 0 }, { -1, 0, 0 }, { 0, +1, 0 }, { 0, -1, 0 }, { 0, 0, +1 },
        { 0, 0, -1 }, { +1, +1, +1 }, { +1, -1, +1 }, { +1, +1, -1 }, { +1, -1, -1 }, { -1, +1, +1 }, { -1, -1, +1 },
        { -1, +1, -1 }, { -1, -1, -1 }, { +1, 0, +1 }, { +1, 0, -1 }, { -1, 0, +1 }, { -1, 0, -1 }, { 0, +1, +1 },
        { 0, +1, -1 }, { 0, -1, +1 }, { 0, -1, -1 }, { +1, +1, 0 }, { -1, +1, 0 }, { +1, -1, 0 }, { -1, -1, 0 } };
    const int N = o->N, jfac = MAXV(1, N / CLUSTER_M), ncj = o->Nclusters_local / jfac;
    const real cutNeigh = o->cutneigh;
    int Nghost = -1, Nghost_atoms = 0;
    for (int cj = 0; cj < ncj; cj++) {
        if (o->jclusters[cj].natoms <= 0) continue;
        while (o->Nclusters_local + (Nghost + 7 + 26) * jfac >= o->Nclusters_max) grow_clusters(o);
        if ((Nghost + 7 + 26) * jfac >= o->NmaxGhost) {
            o->NmaxGhost += DELTA;
            o->border_map = (int*)xrealloc(o->border_map, o->NmaxGhost * sizeof(int));
            o->PBCx = (int*)xrealloc(o->PBCx, o->NmaxGhost * sizeof(int));
            o->PBCy = (int*)xrealloc(o->PBCy, o->NmaxGhost * sizeof(int));
            o->PBCz = (int*)xrealloc(o->PBCz, o->NmaxGhost * sizeof(int));
        }
        const Cluster* q = &o->jclusters[cj];
        const int lo[3] = { q->bbminx < cutNeigh, q->bbminy < cutNeigh, q->bbminz < cutNeigh };
        const int hi[3] = { q->bbmaxx >= (o->xprd - cutNeigh), q->bbmaxy >= (o->yprd - cutNeigh),
            q->bbmaxz >= (o->zprd - cutNeigh) };
        for (int b = 0; b < 26; b++) {
            int ok = 1;
            for (int a = 0; a < 3; a++) {
                if (img[b][a] > 0) ok = ok && lo[a];
                if (img[b][a] < 0) ok = ok && hi[a];
            }
            if (!ok) continue;
            Nghost++;
            const int cg = ncj + Nghost;
            o->border_map[Nghost] = cj;
            o->PBCx[Nghost] = img[b][0]; o->PBCy[Nghost] = img[b][1]; o->PBCz[Nghost] = img[b][2];
            o->jclusters[cg].natoms = q->natoms;
            Nghost_atoms += q->natoms;
            for (int cjj = 0; cjj < N; cjj++)
                o->cl_tag[cj_base(o, cg, 1) + cjj] = cjj < q->natoms ? o->cl_tag[cj_base(o, cj, 1) + cjj] : -1;
        }
    }
    while (ncj + (Nghost + 1) * jfac + jfac >= o->Nclusters_max) grow_clusters(o);
    real* cjX = &o->cl_x[cj_base(o, ncj + Nghost + 1, 3)]; /* dummy cluster at the end, pbc.c:304-311 */
    for (int cjj = 0; cjj < N; cjj++) { cjX[0 * W(o) + cjj] = INFINITY; cjX[1 * W(o) + cjj] = INFINITY; cjX[2 * W(o) + cjj] = INFINITY; }
    o->dummy_cj        = ncj + Nghost + 1;
    o->Nghost          = Nghost_atoms;
    o->Nclusters_ghost = Nghost + 1;
    o->Nclusters       = o->Nclusters_local + Nghost + 1;
    ocp_update_pbc(o, 1);
}

/* neighbor.c:875-1021 binClusters: local j-clusters per bin in cluster order, ghost j-clusters into the bin of
 * their innermost atom, kept sorted by bbminz */
EXPORT void ocp_bin_clusters(OCP* o)
{
    const int N = o->N, nlocal = o->Nclusters_local, jfac = MAXV(1, N / CLUSTER_M), ncj = nlocal / jfac;
    int resize = 1;
    while (resize > 0) {
        resize = 0;
        for (int bin = 0; bin < o->mbins; bin++) o->bin_nclusters[bin] = 0;
        for (int ci = 0; ci < nlocal && !resize; ci++) {
            if (!(CLUSTER_M < N && ci % 2)) {
                int bin = o->icluster_bin[ci], c = o->bin_nclusters[bin];
                if (c + 1 < o->clusters_per_bin) {
                    o->bin_clusters[(size_t)bin * o->clusters_per_bin + c] = cj0_from_ci(o, ci);
                    o->bin_nclusters[bin]++;
                } else resize = 1;
            }
        }
        for (int cg = 0; cg < o->Nclusters_ghost && !resize; cg++) {
            const int cj = ncj + cg;
            int ix = -1, iy = -1;
            if (o->jclusters[cj].natoms > 0) {
                real* cj_x   = &o->cl_x[cj_base(o, cj, 3)];
                real cj_minz = o->jclusters[cj].bbminz;
                coord2bin2D(o, cj_x[0], cj_x[W(o)], &ix, &iy);
                ix = MAXV(MINV(ix, o->mbinx - 1), 0);
                iy = MAXV(MINV(iy, o->mbiny - 1), 0);
                for (int cjj = 1; cjj < o->jclusters[cj].natoms; cjj++) {
                    int nix, niy;
                    coord2bin2D(o, cj_x[cjj], cj_x[W(o) + cjj], &nix, &niy);
                    nix = MAXV(MINV(nix, o->mbinx - 1), 0);
                    niy = MAXV(MINV(niy, o->mbiny - 1), 0);
                    if (o->PBCx[cg] > 0 && ix > nix) ix = nix;
                    if (o->PBCx[cg] < 0 && ix < nix) ix = nix;
                    if (o->PBCy[cg] > 0 && iy > niy) iy = niy;
                    if (o->PBCy[cg] < 0 && iy < niy) iy = niy;
                }
                int bin = iy * o->mbinx + ix + 1, c = o->bin_nclusters[bin];
                if (c < o->clusters_per_bin) {
                    int inserted = 0;
                    int* bc = &o->bin_clusters[(size_t)bin * o->clusters_per_bin];
                    for (int i = 0; i < c; i++) {
                        int last_cl = bc[i];
                        if (o->jclusters[last_cl].bbminz > cj_minz) {
                            bc[i] = cj;
                            for (int j = i + 1; j <= c; j++) { int tmp = bc[j]; bc[j] = last_cl; last_cl = tmp; }
                            inserted = 1;
                            break;
                        }
                    }
                    if (!inserted) bc[c] = cj;
                    o->bin_nclusters[bin]++;
                } else resize = 1;
            }
        }
        if (resize) {
            o->clusters_per_bin *= 2;
            o->bin_clusters = (int*)xrealloc(o->bin_clusters, (size_t)o->mbins * o->clusters_per_bin * sizeof(int));
        }
    }
}

/* neighbor.c:216-234 atomDistanceInRange; the reference build evaluates fma(dz,dz,fma(dx,dx,dy*dy)) */
static int atom_distance_in_range(const OCP* o, int ci, int cj, real rsq)
{
    const real* ci_x = &o->cl_x[ci_base(o, ci, 3)];
    const real* cj_x = &o->cl_x[cj_base(o, cj, 3)];
    for (int cii = 0; cii < o->iclusters[ci].natoms; cii++)
        for (int cjj = 0; cjj < o->jclusters[cj].natoms; cjj++) {
            real delx = ci_x[cii] - cj_x[cjj], dely = ci_x[W(o) + cii] - cj_x[W(o) + cjj],
                 delz = ci_x[2 * W(o) + cii] - cj_x[2 * W(o) + cjj];
            if (RFMA(delz, delz, RFMA(delx, delx, dely * dely)) < rsq) return 1;
        }
    return 0;
}
/* neighbor.c:262-481 buildNeighborCPU.  The do/while that skips leading clusters out of z-range (317-330) only
 * prunes candidates the full test below rejects anyway, so the set is: stencil bins x clusters with
 * d_bb_sq < cutneighsq and (d_bb_sq < rbb_sq or an atom pair in range); diagonal entries moved to the front. */
EXPORT void ocp_build_neighbor(OCP* o)
{
    const int N = o->N;
    if (o->Nclusters_local > o->nmax) {
        o->nmax            = o->Nclusters_local;
        o->numneigh        = (int*)xrealloc(o->numneigh, o->nmax * sizeof(int));
        o->numneigh_masked = (int*)xrealloc(o->numneigh_masked, o->nmax * sizeof(int));
        o->neighbors       = (int*)xrealloc(o->neighbors, (size_t)o->nmax * o->maxneighs * sizeof(int));
    }
    real bbx = 0.5 * (o->binsizex + o->binsizex), bby = 0.5 * (o->binsizey + o->binsizey);
    real rbb_sq = MAXV(0.0, o->cutneigh - 0.5 * sqrt(bbx * bbx + bby * bby));
    rbb_sq      = rbb_sq * rbb_sq;
    int resize  = 1;
    while (resize) {
        int new_maxneighs = o->maxneighs;
        resize = 0;
        for (int ci = 0; ci < o->Nclusters_local; ci++) {
            const int ci_cj1 = cj1_from_ci(o, ci);
            int* neighptr = &o->neighbors[(size_t)ci * o->maxneighs];
            int n = 0, nmasked = 0;
            const int ibin = o->icluster_bin[ci];
            const Cluster* I = &o->iclusters[ci];
            for (int k = 0; k < o->nstencil; k++) {
                const int jbin = ibin + o->stencil[k];
                const int* loc_bin = &o->bin_clusters[(size_t)jbin * o->clusters_per_bin];
                const int c = o->bin_nclusters[jbin];
                for (int m = 0; m < c; m++) {
                    const int cj = loc_bin[m];
                    if (o->half_neigh && ci_cj1 > cj) continue;
                    const Cluster* J = &o->jclusters[cj];
                    real dl, dh, dm, dm0, d_bb_sq;
                    dl = I->bbminz - J->bbmaxz; dh = J->bbminz - I->bbmaxz; dm = MAXV(dl, dh); dm0 = MAXV(dm, (real)0.0);
                    d_bb_sq = dm0 * dm0;
                    dl = I->bbminy - J->bbmaxy; dh = J->bbminy - I->bbmaxy; dm = MAXV(dl, dh); dm0 = MAXV(dm, (real)0.0);
                    d_bb_sq = RFMA(dm0, dm0, d_bb_sq);
                    dl = I->bbminx - J->bbmaxx; dh = J->bbminx - I->bbmaxx; dm = MAXV(dl, dh); dm0 = MAXV(dm, (real)0.0);
                    d_bb_sq = RFMA(dm0, dm0, d_bb_sq);
                    if (d_bb_sq < o->cutneighsq) {
                        if (d_bb_sq < rbb_sq || atom_distance_in_range(o, ci, cj, o->cutneighsq)) {
                            const int masked = cj == cj0_from_ci(o, ci); /* get_imask_simd_*(1, ci, cj) != MASK_ALL */
                            if (n < o->maxneighs) {
                                if (!masked) neighptr[n] = cj;
                                else { neighptr[n] = neighptr[nmasked]; neighptr[nmasked] = cj; nmasked++; }
                            }
                            n++;
                        }
                    }
                }
            }
            if (N < o->vector_width) /* dummy padding to the vector width, neighbor.c:395-403 */
                while (n % (o->vector_width / N)) { if (n < o->maxneighs) neighptr[n] = o->dummy_cj; n++; }
            o->numneigh[ci] = n;
            o->numneigh_masked[ci] = nmasked;
            if (n >= o->maxneighs) { resize = 1; if (n >= new_maxneighs) new_maxneighs = n; }
        }
        if (resize) {
            o->maxneighs = new_maxneighs * 1.2;
            o->neighbors = (int*)xrealloc(o->neighbors, (size_t)o->nmax * o->maxneighs * sizeof(int));
        }
    }
}
/* neighbor.c:483-531 pruneNeighbor: drops listed cluster pairs without an atom pair inside cutneigh (NOT cutforce:
 * the reference's own comment line 486-487); the hole is filled with the row's last entry */
EXPORT void ocp_prune_neighbor(OCP* o)
{
    const int N = o->N;
    for (int ci = 0; ci < o->Nclusters_local; ci++) {
        int* neighs = &o->neighbors[(size_t)ci * o->maxneighs];
        int numneighs = o->numneigh[ci], numneighs_masked = o->numneigh_masked[ci], k = 0;
        if (N < o->vector_width)
            while (numneighs > 0 && neighs[numneighs - 1] == o->dummy_cj) numneighs--;
        while (k < numneighs) {
            if (atom_distance_in_range(o, ci, neighs[k], o->cutneighsq)) k++;
            else {
                numneighs--;
                if (k < numneighs_masked) numneighs_masked--;
                neighs[k] = neighs[numneighs];
            }
        }
        if (N < o->vector_width)
            while (numneighs % (o->vector_width / N)) neighs[numneighs++] = o->dummy_cj;
        o->numneigh[ci] = numneighs;
        o->numneigh_masked[ci] = numneighs_masked;
    }
}
/* neighbor.c:1023-1049 updateSingleAtoms: cluster data back to the atom arrays, compacted in cluster order */
EXPORT void ocp_update_single_atoms(OCP* o)
{
    int Natom = 0;
    for (int ci = 0; ci < o->Nclusters_local; ci++) {
        const real* ci_x = &o->cl_x[ci_base(o, ci, 3)];
        const real* ci_v = &o->cl_v[ci_base(o, ci, 3)];
        for (int cii = 0; cii < o->iclusters[ci].natoms; cii++) {
            o->x[Natom] = ci_x[cii]; o->y[Natom] = ci_x[W(o) + cii]; o->z[Natom] = ci_x[2 * W(o) + cii];
            o->vx[Natom] = ci_v[cii]; o->vy[Natom] = ci_v[W(o) + cii]; o->vz[Natom] = ci_v[2 * W(o) + cii];
            o->tag[Natom] = o->cl_tag[ci_base(o, ci, 1) + cii];
            Natom++;
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* clusterpair/force_lj.c:47-164 computeForceLJRef (the scalar kernel, exact division) */
EXPORT void ocp_compute_force(OCP* o)
{
    const int N = o->N;
    const real cutforcesq = o->cutforce * o->cutforce, sigma6 = o->sigma6, epsilon = o->epsilon;
    for (int ci = 0; ci < o->Nclusters_local; ci++) {
        real* ci_f = &o->cl_f[ci_base(o, ci, 3)];
        for (int cii = 0; cii < o->iclusters[ci].natoms; cii++) { ci_f[cii] = 0.0; ci_f[W(o) + cii] = 0.0; ci_f[2 * W(o) + cii] = 0.0; }
    }
    for (int ci = 0; ci < o->Nclusters_local; ci++) {
        const int ci_cj0 = cj0_from_ci(o, ci);
        const real* ci_x = &o->cl_x[ci_base(o, ci, 3)];
        real* ci_f       = &o->cl_f[ci_base(o, ci, 3)];
        const int* neighs = &o->neighbors[(size_t)ci * o->maxneighs];
        for (int k = 0; k < o->numneigh[ci]; k++) {
            const int cj = neighs[k];
            const real* cj_x = &o->cl_x[cj_base(o, cj, 3)];
            real* cj_f       = &o->cl_f[cj_base(o, cj, 3)];
            for (int cii = 0; cii < CLUSTER_M; cii++) {
                const real xtmp = ci_x[cii], ytmp = ci_x[W(o) + cii], ztmp = ci_x[2 * W(o) + cii];
                real fix = 0, fiy = 0, fiz = 0;
                for (int cjj = 0; cjj < N; cjj++) {
                    const int ii = N == CLUSTER_M ? cii : cii + CLUSTER_M * (ci & 0x1);
                    const int cond = o->half_neigh ? (ci_cj0 != cj || ii < cjj) : (ci_cj0 != cj || ii != cjj);
                    if (cond) {
                        const real delx = xtmp - cj_x[cjj], dely = ytmp - cj_x[W(o) + cjj], delz = ztmp - cj_x[2 * W(o) + cjj];
                        const real rsq = delx * delx + dely * dely + delz * delz;
                        if (rsq < cutforcesq) {
                            const real sr2 = (real)1.0 / rsq, sr6 = sr2 * sr2 * sr2 * sigma6;
                            const real force = (real)48.0 * sr6 * (sr6 - (real)0.5) * sr2 * epsilon;
                            if (o->half_neigh) { cj_f[cjj] -= delx * force; cj_f[W(o) + cjj] -= dely * force; cj_f[2 * W(o) + cjj] -= delz * force; }
                            fix += delx * force; fiy += dely * force; fiz += delz * force;
                        }
                    }
                }
                ci_f[cii] += fix; ci_f[W(o) + cii] += fiy; ci_f[2 * W(o) + cii] += fiz;
            }
        }
    }
}
/* clusterpair/integrate.c:23-44, 46-63 */
EXPORT void ocp_initial_integrate(OCP* o)
{
    for (int ci = 0; ci < o->Nclusters_local; ci++) {
        real *X = &o->cl_x[ci_base(o, ci, 3)], *V = &o->cl_v[ci_base(o, ci, 3)], *F = &o->cl_f[ci_base(o, ci, 3)];
        for (int cii = 0; cii < o->iclusters[ci].natoms; cii++)
            for (int a = 0; a < 3; a++) {
                V[a * W(o) + cii] += o->dtforce * F[a * W(o) + cii];
                X[a * W(o) + cii] += o->dt * V[a * W(o) + cii];
            }
    }
}
EXPORT void ocp_final_integrate(OCP* o)
{
    for (int ci = 0; ci < o->Nclusters_local; ci++) {
        real *V = &o->cl_v[ci_base(o, ci, 3)], *F = &o->cl_f[ci_base(o, ci, 3)];
        for (int cii = 0; cii < o->iclusters[ci].natoms; cii++)
            for (int a = 0; a < 3; a++) V[a * W(o) + cii] += o->dtforce * F[a * W(o) + cii];
    }
}
/* common/thermo.c:55-80 computeThermo -- reads the ATOM arrays (stale between updateSingleAtoms calls) */
EXPORT void ocp_thermo(OCP* o, double* T, double* P)
{
    real t = 0.0;
    for (int i = 0; i < o->Nlocal; i++) t += (o->vx[i] * o->vx[i] + o->vy[i] * o->vy[i] + o->vz[i] * o->vz[i]) * o->mass;
    t = t * o->t_scale;
    real p = (t * o->dof_boltz) * o->p_scale;
    *T = t; *P = p;
}
/* clusterpair/main.c:40-76 setup (after the atoms exist), 78-93 reneighbour */
EXPORT void ocp_setup(OCP* o)
{
    ocp_setup_neighbor(o);
    ocp_build_clusters(o);
    ocp_define_jclusters(o);
    ocp_setup_pbc(o);
    ocp_bin_clusters(o);
    ocp_build_neighbor(o);
}
EXPORT void ocp_reneighbour(OCP* o)
{
    ocp_update_single_atoms(o);
    ocp_update_atoms_pbc(o);
    ocp_build_clusters(o);
    ocp_define_jclusters(o);
    ocp_setup_pbc(o);
    ocp_bin_clusters(o);
    ocp_build_neighbor(o);
}
/* clusterpair/main.c:225-300 time loop; out receives (step, T, P) triples */
EXPORT int ocp_run(OCP* o, int nsteps, double* out, int maxrec)
{
    int nrec = 0;
    double T, P;
    ocp_thermo(o, &T, &P);
    if (nrec < maxrec) { out[3 * nrec] = 0; out[3 * nrec + 1] = T; out[3 * nrec + 2] = P; nrec++; }
    ocp_compute_force(o);
    for (int n = 0; n < nsteps; n++) {
        ocp_initial_integrate(o);
        if ((n + 1) % o->reneigh_every) ocp_update_pbc(o, 0);
        else ocp_reneighbour(o);
        ocp_compute_force(o);
        ocp_final_integrate(o);
        if (!((n + 1) % o->nstat) && (n + 1) < nsteps) {
            ocp_thermo(o, &T, &P);
            if (nrec < maxrec) { out[3 * nrec] = n + 1; out[3 * nrec + 1] = T; out[3 * nrec + 2] = P; nrec++; }
        }
    }
    ocp_update_single_atoms(o);
    ocp_thermo(o, &T, &P);
    if (nrec < maxrec) { out[3 * nrec] = nsteps; out[3 * nrec + 1] = T; out[3 * nrec + 2] = P; nrec++; }
    return nrec;
}

/* ---- accessors ---- */
EXPORT int ocp_get_int(OCP* o, const char* k)
{
#define F(n) if (!strcmp(k, #n)) return o->n;
    F(N) F(Natoms) F(Nlocal) F(Nghost) F(Nclusters) F(Nclusters_local) F(Nclusters_ghost) F(Nclusters_max) F(dummy_cj)
    F(nbinx) F(nbiny) F(mbinx) F(mbiny) F(mbins) F(mbinxlo) F(mbinylo) F(nstencil) F(maxneighs) F(atoms_per_bin)
    F(clusters_per_bin) F(nstat) F(reneigh_every) F(half_neigh)
#undef F
    return -1;
}
EXPORT double ocp_get_real(OCP* o, const char* k)
{
#define F(n) if (!strcmp(k, #n)) return (double)o->n;
    F(xprd) F(yprd) F(zprd) F(binsizex) F(binsizey) F(bininvx) F(bininvy) F(cutneigh) F(cutneighsq) F(lattice) F(dtforce)
#undef F
    return NAN;
}
EXPORT void* ocp_ptr(OCP* o, const char* k)
{
#define F(n) if (!strcmp(k, #n)) return (void*)o->n;
    F(x) F(y) F(z) F(vx) F(vy) F(vz) F(tag) F(border_map) F(PBCx) F(PBCy) F(PBCz) F(cl_x) F(cl_v) F(cl_f) F(cl_tag)
    F(iclusters) F(jclusters) F(icluster_bin) F(bincount) F(bin_nclusters) F(bin_clusters) F(stencil) F(numneigh)
    F(numneigh_masked) F(neighbors)
#undef F
    return NULL;
}
EXPORT int ocp_sizeof_cluster(void) { return (int)sizeof(Cluster); }
