/* operators.c -- the reference's operator API (global function pointers + plain setup functions)
 * bound to libmdb200.  This file is the in-tree example of the shim INTEGRATION.md describes:
 * each operator has the reference's name and signature and forwards to one mdb_* entry point.
 * Failures print the library's message and exit(-1), like cuda_assert (common/device.c:15-21). */
#include <stdlib.h>
#include <string.h>

#include "mdbench.h"

void mdb_die(const char* where)
{
    printf("[CUDA Error]: %s: %s\r\n", where, mdb_last_error());
    exit(-1);
}
#define CK(call, where)                                                                          \
    do {                                                                                         \
        if ((call) != 0) mdb_die(where);                                                         \
    } while (0)

/* ---- force.h:16-35 ---- */
static double computeForceLJFullNeighB200(Parameter* p, Atom* a, Neighbor* n, Stats* s)
{
    double t = mdb_computeForceLJFullNeigh(a->d_atom);
    if (t < 0) mdb_die("computeForceLJFullNeigh");
    return t;
}
static double computeForceLJHalfNeighB200(Parameter* p, Atom* a, Neighbor* n, Stats* s)
{
    double t = mdb_computeForceLJHalfNeigh(a->d_atom);
    if (t < 0) mdb_die("computeForceLJHalfNeigh");
    return t;
}
static double computeForceEamB200(Parameter* p, Atom* a, Neighbor* n, Stats* s)
{
    double t = mdb_computeForceEam(a->d_atom);
    if (t < 0) mdb_die("computeForceEam");
    return t;
}
ComputeForceFunction computeForce;
void initForce(Parameter* param) /* force.c:13-34 */
{
    switch (param->force_field) {
    case FF_EAM: computeForce = computeForceEamB200; break;
    case FF_LJ: computeForce = param->half_neigh ? computeForceLJHalfNeighB200 : computeForceLJFullNeighB200; break;
    default: fprintf(stderr, "Error: Unknown force field!\n"); exit(EXIT_FAILURE);
    }
}

/* ---- neighbor.h:55-59 ---- */
static void buildNeighborB200(Atom* a, Neighbor* n)
{
    CK(mdb_buildNeighbor(a->d_atom), "buildNeighbor");
    long long nl, ng;
    CK(mdb_getCounts(a->d_atom, NULL, &nl, &ng, NULL, &n->maxneighs), "buildNeighbor");
    n->ncalls++;
}
BuildNeighborFunction buildNeighbor = buildNeighborB200;
void setupNeighbor(Parameter* p, Atom* a) { CK(mdb_setupNeighbor(a->d_atom), "setupNeighbor"); }

/* ---- integrate.h:12-14 ---- */
static void initialIntegrateB200(bool reneigh, Parameter* p, Atom* a) { CK(mdb_initialIntegrate(a->d_atom, reneigh), "initialIntegrate"); }
static void finalIntegrateB200(bool reneigh, Parameter* p, Atom* a) { CK(mdb_finalIntegrate(a->d_atom, reneigh), "finalIntegrate"); }
IntegrationFunction initialIntegrate = initialIntegrateB200;
IntegrationFunction finalIntegrate   = finalIntegrateB200;

/* ---- pbc.h:15-22 ---- */
static void updatePbcB200(Atom* a, Parameter* p, bool reneigh) { CK(mdb_updatePbc(a->d_atom, reneigh), "updatePbc"); }
static void updateAtomsPbcB200(Atom* a, Parameter* p, bool reneigh) { CK(mdb_updateAtomsPbc(a->d_atom, reneigh), "updateAtomsPbc"); }
UpdatePbcFunction updatePbc      = updatePbcB200;
UpdatePbcFunction updateAtomsPbc = updateAtomsPbcB200;
void setupPbc(Atom* a, Parameter* p)
{
    CK(mdb_setupPbc(a->d_atom), "setupPbc");
    long long ng;
    CK(mdb_getCounts(a->d_atom, NULL, NULL, &ng, NULL, NULL), "setupPbc");
    a->Nghost = (int)ng;
}

/* ---- thermo.h ---- */
void setupThermo(Parameter* p, Atom* a) { CK(mdb_setupThermo(a->d_atom), "setupThermo"); }
void adjustThermo(Parameter* p, Atom* a) { CK(mdb_adjustThermo(a->d_atom), "adjustThermo"); }
void computeThermo(int iflag, Parameter* p, Atom* a) /* thermo.c:55-80 */
{
    double T, P;
    CK(mdb_computeThermo(a->d_atom, &T, &P), "computeThermo");
    fprintf(stdout, "%i\t%e\t%e\n", iflag == -1 ? p->ntimes : iflag, T, P);
}

/* ---- device.h:20 initDevice: create the device context from the final parameters ---- */
void initDevice(Atom* a, Parameter* p)
{
    mdb_params q;
    mdb_default_params(&q);
    q.precision = p->precision; q.layout = p->layout; q.force_field = p->force_field;
    q.epsilon = p->epsilon; q.sigma = p->sigma; q.temp = p->temp; q.rho = p->rho; q.mass = p->mass;
    q.ntypes = 1; q.ntimes = p->ntimes; q.nstat = p->nstat; q.reneigh_every = p->reneigh_every;
    q.half_neigh = p->half_neigh; q.dt = p->dt; q.skin = p->skin; q.cutforce = p->cutforce;
    q.nx = p->nx; q.ny = p->ny; q.nz = p->nz; q.pbc_x = p->pbc_x; q.pbc_y = p->pbc_y; q.pbc_z = p->pbc_z;
    q.from_input = p->input_file != NULL;
    q.xlo = p->xlo; q.xhi = p->xhi; q.ylo = p->ylo; q.yhi = p->yhi; q.zlo = p->zlo; q.zhi = p->zhi;
    a->d_atom = mdb_create(&q, p->device);
    if (!a->d_atom) mdb_die("initDevice");
    if (p->sort_atoms) CK(mdb_setOption(a->d_atom, "sort_atoms", 1.0), "initDevice");
}

/* ---- atom.c:67-187 createAtom: generated on the device ---- */
void createAtom(Atom* a, Parameter* p)
{
    long long n = mdb_createAtom(a->d_atom);
    if (n < 0) mdb_die("createAtom");
    a->Natoms = a->Nlocal = (int)n;
    a->ntypes = p->ntypes;
}
