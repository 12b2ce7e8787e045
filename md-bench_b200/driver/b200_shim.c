/* b200_shim.c -- libmdb200 bound into the reference's OWN, UNMODIFIED driver.
 *
 * This is the file INTEGRATION.md tells an MD-Bench maintainer to add (src/verletlist/b200_shim.c).  It is compiled
 * against the reference's real headers (atom.h, neighbor.h, pbc.h, integrate.h, force.h, device.h, eam.h, parameter.h)
 * and linked with the reference's unmodified main.c, atom.c, stats.c, vtk.c and src/common/{parameter,thermo,eam_utils,
 * util,allocate,timing}.c, built with -DCUDA_TARGET exactly like the reference's NVCC variant (so that main.c:276-278
 * fetches the device state before it prints thermo).  It REPLACES the reference's neighbor.c, pbc.c, integrate.c,
 * force*.c, device.c and device_spec.c: every function pointer and plain function main.c calls from those files is
 * defined here and forwards to ONE entry point of include/mdb200.h.
 *
 *   recipe:  target ref-shim of the checker's Makefile (the one that compiles the reference)  ->  MDBench-vl_{dp_aos,sp_soa}-b200
 *   test:    tests/test_gpu_parity.py::test_reference_main_c_drives_libmdb200
 *
 * Flow of the reference's setup() (verletlist/main.c:36-74) with this shim:
 *   initAtom, createAtom / readAtom, setupThermo, adjustThermo   reference code, on the host
 *   initNeighbor                                                 remembers the Parameter*
 *   setupNeighbor, setupPbc (before initDevice)                  nothing yet: there is no device context
 *   initDevice                                                   mdb_create + mdb_setEam + mdb_setAtoms (host arrays in the
 *                                                                driver's precision and AoS/SoA layout) + mdb_setupNeighbor
 *                                                                + mdb_setupPbc; atom->Nghost is reported back
 *   updatePbc, buildNeighbor, computeForce, integrate            function pointers -> mdb_*
 * Like the reference's CUDA operators (forceCuda.cu:136-156) the integrate functions copy the velocities back to the
 * host when `reneigh` is set, which is what the host computeThermo (common/thermo.c:55-80) reads.
 */
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atom.h>
#include <device.h>
#include <eam.h>
#include <force.h>
#include <integrate.h>
#include <neighbor.h>
#include <parameter.h>
#include <pbc.h>

#include "mdb200.h"

static mdb_ctx* ctx;           /* one domain */
static Parameter* b200_param;  /* main()'s Parameter, seen first by initNeighbor() */
static Atom* b200_atom;

static void die(const char* where)
{
    printf("[CUDA Error]: %s: %s\r\n", where, mdb_last_error());
    exit(-1);
}
#define CK(call, where)                                                                          \
    do {                                                                                         \
        if ((call) != 0) die(where);                                                             \
    } while (0)

/* opaque stand-ins for the reference's device pointers: main.c passes atom.d_atom.x to memcpyFromGPU */
static MD_FLOAT tag_x, tag_v;

/* ---- device.h:20-25 -------------------------------------------------------------------------- */
void* allocateGPU(size_t bytesize) { return NULL; }              /* the library owns every device array */
void* reallocateGPU(void* ptr, size_t new_bytesize) { return ptr; }
void memcpyToGPU(void* d_ptr, void* h_ptr, size_t bytesize) {}
void memsetGPU(void* d_ptr, int value, size_t bytesize) {}
void memcpyFromGPU(void* h_ptr, void* d_ptr, size_t bytesize)
{
    Atom* atom = b200_atom;
    if (d_ptr == (void*)&tag_x) CK(mdb_getAtoms(ctx, 'x', 0, atom->x, atom->y, atom->z), "memcpyFromGPU");
    else if (d_ptr == (void*)&tag_v) CK(mdb_getAtoms(ctx, 'v', 0, atom->vx, atom->vy, atom->vz), "memcpyFromGPU");
}

void initDevice(Atom* atom, Neighbor* neighbor)                  /* device_spec.c:11 */
{
    Parameter* p = b200_param;
    mdb_params q;
    mdb_default_params(&q);
    q.precision = PRECISION;                                     /* config.mk: -DPRECISION=1|2 */
#ifdef AOS
    q.layout = MDB_AOS;
#else
    q.layout = MDB_SOA;
#endif
    q.force_field = p->force_field == FF_EAM ? MDB_FF_EAM : MDB_FF_LJ;
    q.epsilon = p->epsilon; q.sigma = p->sigma; q.temp = p->temp; q.rho = p->rho; q.mass = p->mass;
    q.ntypes = p->ntypes; q.ntimes = p->ntimes; q.nstat = p->nstat; q.reneigh_every = p->reneigh_every;
    q.half_neigh = p->half_neigh; q.dt = p->dt; q.skin = p->skin; q.cutforce = p->cutforce;
    q.nx = p->nx; q.ny = p->ny; q.nz = p->nz; q.pbc_x = p->pbc_x; q.pbc_y = p->pbc_y; q.pbc_z = p->pbc_z;
    q.from_input = p->input_file != NULL;
    q.xlo = p->xlo; q.xhi = p->xhi; q.ylo = p->ylo; q.yhi = p->yhi; q.zlo = p->zlo; q.zhi = p->zhi;
    if (p->force_field == FF_EAM) q.skin = p->cutneigh - p->cutforce; /* initEam: cutneigh = cutforce + 1 */
    if (!(ctx = mdb_create(&q, 0))) die("initDevice");
    if (getenv("MDB_LAZY_OPS") != NULL) CK(mdb_setOption(ctx, "lazy_ops", 1.0), "initDevice");
    if (p->force_field == FF_EAM) {                              /* the funcfl tables initEam has read */
        Funcfl* f = &eam.file;
        double *frho = malloc(sizeof(double) * f->nrho), *zr = malloc(sizeof(double) * f->nr), *rhor = malloc(sizeof(double) * f->nr);
        for (int i = 0; i < f->nrho; i++) frho[i] = f->frho[i + 1]; /* readEamFile shifts the tables to 1-based, eam_utils.c:86-91 */
        for (int i = 0; i < f->nr; i++) { zr[i] = f->zr[i + 1]; rhor[i] = f->rhor[i + 1]; }
        CK(mdb_setEam(ctx, f->nrho, f->drho, f->nr, f->dr, f->cut, f->mass, frho, zr, rhor), "initEam");
        free(frho); free(zr); free(rhor);
    }
    /* what createAtom / readAtom and the host adjustThermo produced (atom.h:51-73 layout macros) */
    CK(mdb_setAtoms(ctx, atom->Nlocal, atom->x, atom->y, atom->z, atom->vx, atom->vy, atom->vz, atom->type), "initDevice");
    CK(mdb_setupNeighbor(ctx), "setupNeighbor");
    CK(mdb_setupThermo(ctx), "setupThermo");
    b200_atom      = atom;
    atom->d_atom.x = &tag_x;
    atom->d_atom.vx = &tag_v;
    setupPbc(atom, p);                                           /* main.c:67 ran before the context existed */
}

/* ---- neighbor.h:55-59 ------------------------------------------------------------------------ */
void initNeighbor(Neighbor* neighbor, Parameter* param)          /* neighbor.c:43-62 */
{
    b200_param            = param;
    neighbor->ncalls      = 0;
    neighbor->maxneighs   = 100;
    neighbor->half_neigh  = param->half_neigh;
    neighbor->numneigh    = NULL;
    neighbor->neighbors   = NULL;
}
void setupNeighbor(Parameter* param) {}                          /* bin geometry: mdb_setupNeighbor in initDevice */
static void buildNeighborB200(Atom* atom, Neighbor* neighbor)
{
    CK(mdb_buildNeighbor(ctx), "buildNeighbor");
    CK(mdb_getCounts(ctx, NULL, NULL, NULL, NULL, &neighbor->maxneighs), "buildNeighbor");
    neighbor->ncalls++;
}
BuildNeighborFunction buildNeighbor = buildNeighborB200;

/* ---- pbc.h:15-22 ----------------------------------------------------------------------------- */
void initPbc(Atom* atom) {}
void setupPbc(Atom* atom, Parameter* param)                      /* pbc.c:98-227, on the device */
{
    if (!ctx) return;
    long long ng;
    CK(mdb_setupPbc(ctx), "setupPbc");
    CK(mdb_getCounts(ctx, NULL, NULL, &ng, NULL, NULL), "setupPbc");
    atom->Nghost = (int)ng;
}
static void updatePbcB200(Atom* atom, Parameter* param, bool reneigh) { CK(mdb_updatePbc(ctx, reneigh), "updatePbc"); }
static void updateAtomsPbcB200(Atom* atom, Parameter* param, bool reneigh) { CK(mdb_updateAtomsPbc(ctx, reneigh), "updateAtomsPbc"); }
UpdatePbcFunction updatePbc      = updatePbcB200;
UpdatePbcFunction updateAtomsPbc = updateAtomsPbcB200;

/* ---- integrate.h:12-14 ----------------------------------------------------------------------- */
static void initialIntegrateB200(bool reneigh, Parameter* param, Atom* atom)
{
    CK(mdb_initialIntegrate(ctx, reneigh), "initialIntegrate");
    if (reneigh) memcpyFromGPU(atom->vx, atom->d_atom.vx, 0);    /* forceCuda.cu:154-156 */
}
static void finalIntegrateB200(bool reneigh, Parameter* param, Atom* atom)
{
    CK(mdb_finalIntegrate(ctx, reneigh), "finalIntegrate");
    if (reneigh) memcpyFromGPU(atom->vx, atom->d_atom.vx, 0);    /* forceCuda.cu:136-138 */
}
IntegrationFunction initialIntegrate = initialIntegrateB200;
IntegrationFunction finalIntegrate   = finalIntegrateB200;

/* ---- force.h:16-17, force.c:13-34 ------------------------------------------------------------ */
static double forceLJFullB200(Parameter* p, Atom* a, Neighbor* n, Stats* s) { double t = mdb_computeForceLJFullNeigh(ctx); if (t < 0) die("computeForceLJFullNeigh"); return t; }
static double forceLJHalfB200(Parameter* p, Atom* a, Neighbor* n, Stats* s) { double t = mdb_computeForceLJHalfNeigh(ctx); if (t < 0) die("computeForceLJHalfNeigh"); return t; }
static double forceEamB200(Parameter* p, Atom* a, Neighbor* n, Stats* s) { double t = mdb_computeForceEam(ctx); if (t < 0) die("computeForceEam"); return t; }
ComputeForceFunction computeForce;
void initForce(Parameter* param)
{
    switch (param->force_field) {
    case FF_EAM: computeForce = forceEamB200; break;
    case FF_LJ: computeForce = param->half_neigh ? forceLJHalfB200 : forceLJFullB200; break;
    default: fprintf(stderr, "Error: Unknown force field!\n"); exit(EXIT_FAILURE);
    }
}
