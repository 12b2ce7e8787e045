/* b200_shim_cp.c -- libmdb200 bound into the reference's OWN, UNMODIFIED clusterpair driver.
 *
 * The clusterpair counterpart of b200_shim.c: compiled against the reference's real headers
 * (clusterpair/atom.h, neighbor.h, pbc.h, integrate.h, force.h, common/device.h, parameter.h) and linked with the
 * reference's unmodified clusterpair/main.c, atom.c, stats.c, vtk.c and src/common/{parameter,thermo,eam_utils,util,
 * allocate,timing}.c.  It REPLACES the reference's clusterpair/neighbor.c, pbc.c, integrate.c, force*.c and
 * common/device.c: every function pointer and plain function main.c calls from those files (main.c:40-93, 225-300) is
 * defined here and forwards to ONE mdb_cp_* entry point of include/mdb200.h.  CLUSTER_N comes from the reference's own
 * force.h (VECTOR_WIDTH), as in its builds.
 *
 *   recipe:  target ref-shim of the checker's Makefile (the one that compiles the reference)  ->  MDBench-cp44_{sp,dp}-b200
 *   test:    tests/test_gpu_cp.py::test_reference_clusterpair_main_c_drives_libmdb200
 *
 * Flow of the reference's setup() (clusterpair/main.c:40-76) with this shim:
 *   initAtom, createAtom / readAtom, setupThermo, adjustThermo    reference code, on the host
 *   initNeighbor                                                  remembers the Parameter*
 *   setupNeighbor(param, atom)                                    mdb_cp_create (the atoms exist, their final velocities not yet)
 *   buildClusters (first call)                                    mdb_cp_setAtoms with what adjustThermo left on the host +
 *                                                                 mdb_cp_setupNeighbor, then mdb_cp_buildClusters
 *   defineJClusters, setupPbc, binClusters, buildNeighbor         mdb_cp_*
 * The host computeThermo (common/thermo.c) reads the atom arrays, which the reference refreshes in updateSingleAtoms only
 * (neighbor.c:1023-1049): the shim's updateSingleAtoms copies the device's atom arrays back at the same point.
 */
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atom.h>
#include <device.h>
#include <force.h>
#include <integrate.h>
#include <neighbor.h>
#include <parameter.h>
#include <pbc.h>
#include <stats.h>

#include "mdb200.h"

static mdb_cp* ctx;
static Parameter* b200_param;
static int uploaded;

static void die(const char* where)
{
    printf("[CUDA Error]: %s: %s\r\n", where, mdb_last_error());
    exit(-1);
}
#define CK(call, where)                                                                          \
    do {                                                                                         \
        if ((call) != 0) die(where);                                                             \
    } while (0)

void initDevice(Atom* atom, Neighbor* neighbor) {}               /* common/device.c:70: everything is on the device already */

/* ---- neighbor.h:40-50 ------------------------------------------------------------------------ */
void initNeighbor(Neighbor* neighbor, Parameter* param)          /* neighbor.c:47-68 */
{
    b200_param                = param;
    neighbor->ncalls          = 0;
    neighbor->maxneighs       = 100;
    neighbor->half_neigh      = param->half_neigh;
    neighbor->numneigh        = NULL;
    neighbor->numneigh_masked = NULL;
    neighbor->neighbors       = NULL;
    neighbor->neighbors_imask = NULL;
}
void setupNeighbor(Parameter* p, Atom* atom)                     /* neighbor.c:70-172 */
{
    mdb_params q;
    mdb_default_params(&q);
    q.precision = PRECISION;                                     /* config.mk: -DPRECISION=1|2 */
#ifdef AOS
    q.layout = MDB_AOS;
#else
    q.layout = MDB_SOA;
#endif
    q.force_field = MDB_FF_LJ;
    q.epsilon = p->epsilon; q.sigma = p->sigma; q.temp = p->temp; q.rho = p->rho; q.mass = p->mass;
    q.ntypes = p->ntypes; q.ntimes = p->ntimes; q.nstat = p->nstat; q.reneigh_every = p->reneigh_every;
    q.half_neigh = p->half_neigh; q.dt = p->dt; q.skin = p->skin; q.cutforce = p->cutforce;
    q.nx = p->nx; q.ny = p->ny; q.nz = p->nz; q.pbc_x = p->pbc_x; q.pbc_y = p->pbc_y; q.pbc_z = p->pbc_z;
    q.from_input = p->input_file != NULL;
    q.xlo = p->xlo; q.xhi = p->xhi; q.ylo = p->ylo; q.yhi = p->yhi; q.zlo = p->zlo; q.zhi = p->zhi;
    if (!(ctx = mdb_cp_create(&q, CLUSTER_N, 0))) die("setupNeighbor");
    CK(mdb_cp_setOption(ctx, "prune_every", (double)p->prune_every), "setupNeighbor");
}
void buildClusters(Atom* atom)                                   /* neighbor.c:599-753 */
{
    if (!uploaded) { /* what createAtom / readAtom and the host adjustThermo produced (atom.h:66-92 layout macros) */
        CK(mdb_cp_setAtoms(ctx, atom->Nlocal, atom->x, atom->y, atom->z, atom->vx, atom->vy, atom->vz), "buildClusters");
        CK(mdb_cp_setupNeighbor(ctx), "setupNeighbor");
        uploaded = 1;
    }
    CK(mdb_cp_buildClusters(ctx), "buildClusters");
}
void defineJClusters(Atom* atom) { CK(mdb_cp_defineJClusters(ctx), "defineJClusters"); }   /* neighbor.c:755-873 */
void binClusters(Atom* atom) { CK(mdb_cp_binClusters(ctx), "binClusters"); }               /* neighbor.c:875-1021 */
void updateSingleAtoms(Atom* atom)                                                          /* neighbor.c:1023-1049 */
{
    CK(mdb_cp_updateSingleAtoms(ctx), "updateSingleAtoms");
    CK(mdb_cp_getAtoms(ctx, 'x', atom->x, atom->y, atom->z, NULL), "updateSingleAtoms");
    CK(mdb_cp_getAtoms(ctx, 'v', atom->vx, atom->vy, atom->vz, NULL), "updateSingleAtoms");
}
void pruneNeighbor(Parameter* p, Atom* atom, Neighbor* neighbor) { CK(mdb_cp_pruneNeighbor(ctx), "pruneNeighbor"); } /* neighbor.c:483-531 */
static void buildNeighborB200(Atom* atom, Neighbor* neighbor)                               /* neighbor.c:262-481 */
{
    long long v[8];
    CK(mdb_cp_buildNeighbor(ctx), "buildNeighbor");
    CK(mdb_cp_getCounts(ctx, v), "buildNeighbor");
    neighbor->maxneighs = (int)v[6];
    neighbor->ncalls++;
}
BuildNeighborFunction buildNeighbor = buildNeighborB200;

/* ---- pbc.h ------------------------------------------------------------------------------------ */
void initPbc(Atom* atom) {}
void setupPbc(Atom* atom, Parameter* param)                      /* pbc.c:183-323 */
{
    long long v[8];
    CK(mdb_cp_setupPbc(ctx), "setupPbc");
    CK(mdb_cp_getCounts(ctx, v), "setupPbc");
    atom->Nghost          = (int)v[2];
    atom->Nclusters_ghost = (int)v[4];
}
static void updatePbcB200(Atom* atom, Parameter* param, bool first) { CK(mdb_cp_updatePbc(ctx, first), "updatePbc"); }
static void updateAtomsPbcB200(Atom* atom, Parameter* param, bool first) { CK(mdb_cp_updateAtomsPbc(ctx), "updateAtomsPbc"); }
UpdatePbcFunction updatePbc      = updatePbcB200;
UpdatePbcFunction updateAtomsPbc = updateAtomsPbcB200;

/* ---- integrate.h:14 --------------------------------------------------------------------------- */
static void initialIntegrateB200(Parameter* param, Atom* atom) { CK(mdb_cp_initialIntegrate(ctx), "initialIntegrate"); }
static void finalIntegrateB200(Parameter* param, Atom* atom) { CK(mdb_cp_finalIntegrate(ctx), "finalIntegrate"); }
IntegrationFunction initialIntegrate = initialIntegrateB200;
IntegrationFunction finalIntegrate   = finalIntegrateB200;

/* ---- force.h:16-21, force.c -------------------------------------------------------------------- */
static double forceB200(Parameter* p, Atom* a, Neighbor* n, Stats* s)
{
    const double t = mdb_cp_computeForce(ctx);
    if (t < 0) die("computeForce");
    return t;
}
ComputeForceFunction computeForce = forceB200;
void initForce(Parameter* param)
{
    if (param->force_field != FF_LJ) {
        fprintf(stderr, "Error: the clusterpair scheme has only the LJ kernels!\n");
        exit(EXIT_FAILURE);
    }
    computeForce = forceB200;
}
