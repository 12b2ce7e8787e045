// mdb200.cu -- the extern "C" boundary declared in include/mdb200.h.
// Each entry point forwards to the precision-specific Sim<real>; C++ exceptions become a non-zero
// return + mdb_last_error().  There is no CPU path: make_sim() fails without a CUDA device.
#include "sim.cuh"

using namespace mdb;

struct mdb_ctx {
    SimBase* sim;
};

static thread_local std::string g_err;
namespace mdb {
void set_last_error(const char* msg) { g_err = msg; } // for the other translation units of the C ABI (cp_sim.cu)
}

// Every entry point first launches what the lazy operators still owe (sim.cuh) and marks the gather copies stale;
// MDB_TRY_KEEP (updatePbc: it refreshes the copies' ghost range itself) and MDB_TRY_LAZY (the three lazy operators) differ.
#define MDB_TRY_(pre, body)                                                                      \
    try {                                                                                        \
        mdb::NvtxRange nvtx_range_(__func__);                                                    \
        if (!c || !c->sim) throw Error("null mdb_ctx");                                          \
        pre;                                                                                     \
        body;                                                                                    \
        return 0;                                                                                \
    } catch (const std::exception& e) {                                                          \
        g_err = e.what();                                                                        \
        return -1;                                                                               \
    }
#define MDB_TRY(body) MDB_TRY_(c->sim->flush_lazy(); c->sim->invalidate_copies(), body)
#define MDB_TRY_KEEP(body) MDB_TRY_(c->sim->flush_lazy(), body)
#define MDB_TRY_LAZY(body) MDB_TRY_((void)0, body)

extern "C" {

int mdb_abi_version(void) { return MDB200_ABI_VERSION; }
const char* mdb_last_error(void) { return g_err.c_str(); }

void mdb_default_params(mdb_params* p) // common/parameter.c:16-51
{
    memset(p, 0, sizeof *p);
    p->precision     = MDB_DP;
    p->layout        = MDB_AOS;
    p->force_field   = MDB_FF_LJ;
    p->epsilon       = 1.0;
    p->sigma         = 1.0;
    p->rho           = 0.8442;
    p->ntypes        = 1;
    p->ntimes        = 200;
    p->dt            = 0.005;
    p->nx = p->ny = p->nz = 32;
    p->pbc_x = p->pbc_y = p->pbc_z = 1;
    p->cutforce      = 2.5;
    p->skin          = 0.3;
    p->temp          = 1.44;
    p->nstat         = 100;
    p->mass          = 1.0;
    p->reneigh_every = 20;
    p->half_neigh    = 0;
}

mdb_ctx* mdb_create(const mdb_params* p, int device)
{
    try {
        if (!p) throw Error("mdb_create: null params");
        if (p->ntypes < 1) throw Error("mdb_create: ntypes must be >= 1");
        mdb_ctx* c = new mdb_ctx;
        c->sim     = make_sim(*p, device);
        return c;
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
void mdb_destroy(mdb_ctx* c)
{
    if (!c) return;
    delete c->sim;
    delete c;
}
int mdb_setStream(mdb_ctx* c, void* s) { MDB_TRY(c->sim->setStream((cudaStream_t)s)) }
int mdb_sync(mdb_ctx* c) { MDB_TRY(c->sim->sync()) }

long long mdb_createAtom(mdb_ctx* c)
{
    try {
        if (!c || !c->sim) throw Error("null mdb_ctx");
        c->sim->drop_lazy(); // pending force / finalIntegrate of the OLD atoms must not run against the new ones
        c->sim->invalidate_copies();
        return c->sim->createAtom();
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}
int mdb_setAtoms(mdb_ctx* c, long long n, const void* x, const void* y, const void* z, const void* vx,
    const void* vy, const void* vz, const int* type)
{
    MDB_TRY(c->sim->setAtoms(n, x, y, z, vx, vy, vz, type, false))
}
int mdb_setAtomsDevice(mdb_ctx* c, long long n, const void* x, const void* y, const void* z, const void* vx,
    const void* vy, const void* vz, const int* type)
{
    MDB_TRY(c->sim->setAtoms(n, x, y, z, vx, vy, vz, type, true))
}
int mdb_getTypes(mdb_ctx* c, int* types) { MDB_TRY(c->sim->getTypes(types)) }
int mdb_getAtoms(mdb_ctx* c, int which, int with_ghosts, void* x, void* y, void* z)
{
    MDB_TRY(c->sim->getAtoms(which, with_ghosts != 0, x, y, z))
}
int mdb_getCounts(mdb_ctx* c, long long* Natoms, long long* Nlocal, long long* Nghost, long long* Nmax, int* maxneighs)
{
    MDB_TRY({
        long long v[4];
        int mn;
        c->sim->getCounts(v, &mn);
        if (Natoms) *Natoms = v[0];
        if (Nlocal) *Nlocal = v[1];
        if (Nghost) *Nghost = v[2];
        if (Nmax) *Nmax = v[3];
        if (maxneighs) *maxneighs = mn;
    })
}
int mdb_saveState(mdb_ctx* c) { MDB_TRY(c->sim->saveState()) }
int mdb_restoreState(mdb_ctx* c) { MDB_TRY(c->sim->restoreState()) }

int mdb_setupThermo(mdb_ctx* c) { MDB_TRY(c->sim->setupThermo()) }
int mdb_adjustThermo(mdb_ctx* c) { MDB_TRY(c->sim->adjustThermo()) }
int mdb_computeThermo(mdb_ctx* c, double* T, double* P) { MDB_TRY(c->sim->computeThermo(T, P)) }

int mdb_setupNeighbor(mdb_ctx* c) { MDB_TRY(c->sim->setupNeighbor()) }
int mdb_setupPbc(mdb_ctx* c) { MDB_TRY(c->sim->setupPbc()) }
int mdb_updatePbc(mdb_ctx* c, int) { MDB_TRY_KEEP(c->sim->updatePbc()) }
int mdb_updateAtomsPbc(mdb_ctx* c, int) { MDB_TRY(c->sim->updateAtomsPbc()) }
int mdb_buildNeighbor(mdb_ctx* c) { MDB_TRY(c->sim->buildNeighbor()) }

static double force_call(mdb_ctx* c, int which)
{
    try {
        if (!c || !c->sim) throw Error("null mdb_ctx");
        return c->sim->abi_computeForce(which);
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1.0;
    }
}
double mdb_computeForce(mdb_ctx* c) { return force_call(c, FORCE_DISPATCH); }
double mdb_computeForceLJFullNeigh(mdb_ctx* c) { return force_call(c, FORCE_LJ_FULL); }
double mdb_computeForceLJHalfNeigh(mdb_ctx* c) { return force_call(c, FORCE_LJ_HALF); }
double mdb_computeForceEam(mdb_ctx* c) { return force_call(c, FORCE_EAM); }
int mdb_initialIntegrate(mdb_ctx* c, int) { MDB_TRY_LAZY(c->sim->abi_initialIntegrate()) }
int mdb_finalIntegrate(mdb_ctx* c, int) { MDB_TRY_LAZY(c->sim->abi_finalIntegrate()) }

int mdb_setup(mdb_ctx* c, int adjust) { MDB_TRY(c->sim->setup(adjust != 0)) }
int mdb_reneighbour(mdb_ctx* c) { MDB_TRY(c->sim->reneighbour()) }
int mdb_run(mdb_ctx* c, int nsteps, double* thermo_out, int max_records, int* nrecords, double* timers)
{
    MDB_TRY(c->sim->run(nsteps, thermo_out, max_records, nrecords, timers))
}
int mdb_setTiming(mdb_ctx* c, int on) { MDB_TRY(c->sim->timing = on != 0) }
int mdb_getKernelStats(mdb_ctx* c, double* force_ms, long long* force_launches, double* neigh_ms,
    long long* neigh_launches, long long* total_launches)
{
    MDB_TRY({
        if (force_ms) *force_ms = c->sim->force_ms;
        if (force_launches) *force_launches = c->sim->force_launches;
        if (neigh_ms) *neigh_ms = c->sim->neigh_ms;
        if (neigh_launches) *neigh_launches = c->sim->neigh_launches;
        if (total_launches) *total_launches = c->sim->launches;
    })
}
int mdb_resetKernelStats(mdb_ctx* c)
{
    MDB_TRY({
        c->sim->force_ms = c->sim->neigh_ms = 0;
        c->sim->force_launches = c->sim->neigh_launches = c->sim->launches = 0;
    })
}

int mdb_setEam(mdb_ctx* c, int nrho, double drho, int nr, double dr, double cut, double mass, const double* frho,
    const double* zr, const double* rhor)
{
    MDB_TRY(c->sim->setEam(nrho, drho, nr, dr, cut, mass, frho, zr, rhor))
}
int mdb_setEamSplines(mdb_ctx* c, int nr, int nrho, int nr_tot, int nrho_tot, double rdr, double rdrho,
    const void* rhor_spline, const void* frho_spline, const void* z2r_spline)
{
    MDB_TRY(c->sim->setEamSplines(nr, nrho, nr_tot, nrho_tot, rdr, rdrho, rhor_spline, frho_spline, z2r_spline))
}
int mdb_getEamSplines(mdb_ctx* c, int* nr, int* nrho, int* nr_tot, int* nrho_tot, double* rdr, double* rdrho,
    void* rhor_spline, void* frho_spline, void* z2r_spline)
{
    MDB_TRY(c->sim->getEamSplines(nr, nrho, nr_tot, nrho_tot, rdr, rdrho, rhor_spline, frho_spline, z2r_spline))
}

int mdb_getNeighbors(mdb_ctx* c, int* numneigh, int* neighbors, int row_stride)
{
    MDB_TRY(c->sim->getNeighbors(numneigh, neighbors, row_stride))
}
int mdb_getGhostMap(mdb_ctx* c, int* bm, int* px, int* py, int* pz) { MDB_TRY(c->sim->getGhostMap(bm, px, py, pz)) }
int mdb_getNeighborParams(mdb_ctx* c, int ints[12], double reals[14]) { MDB_TRY(c->sim->getNeighborParams(ints, reals)) }
int mdb_getStencil(mdb_ctx* c, int* st) { MDB_TRY(c->sim->getStencil(st)) }
int mdb_getBinCounts(mdb_ctx* c, int* bc) { MDB_TRY(c->sim->getBinCounts(bc)) }
int mdb_getEamFp(mdb_ctx* c, void* fp, int with_ghosts) { MDB_TRY(c->sim->getEamFp(fp, with_ghosts != 0)) }
int mdb_setOption(mdb_ctx* c, const char* name, double value) { MDB_TRY(c->sim->setOption(name, value)) }
int mdb_countPairs(mdb_ctx* c, long long* listed, long long* in_cutoff) { MDB_TRY(c->sim->countPairs(listed, in_cutoff)) }
int mdb_stubNeighbors(mdb_ctx* c, int pattern, int nneighs, int nreps, unsigned seed)
{
    MDB_TRY(c->sim->stubNeighbors(pattern, nneighs, nreps, seed))
}

} // extern "C"

// ---- spatial decomposition (include/mdb200.h, "multi-GPU") -----------------------------------------
#include "dd_topo.h"
#include "nccl_dl.h"

struct mdb_dd {
    DDBase* g;
};

#define MDB_DD_TRY(body)                                                                         \
    try {                                                                                        \
        if (!d || !d->g) throw Error("null mdb_dd");                                             \
        body;                                                                                    \
        return 0;                                                                                \
    } catch (const std::exception& e) {                                                          \
        g_err = e.what();                                                                        \
        return -1;                                                                               \
    }

extern "C" {

int mdb_dd_uniqueIdBytes(void) { return (int)sizeof(NcclApi::unique_id); }
int mdb_dd_getUniqueId(void* id)
{
    try {
        NcclApi& N = nccl_api();
        N.load();
        MDB_NCCL(N.GetUniqueId((NcclApi::unique_id*)id));
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}
int mdb_dd_plan(int gx, int gy, int gz, int perx, int pery, int perz, int nprocs, int brick, int send, int dir[26],
    int peer[26], int owner[26])
{
    Topo t;
    t.g[0] = gx; t.g[1] = gy; t.g[2] = gz;
    t.periodic[0] = perx; t.periodic[1] = pery; t.periodic[2] = perz;
    t.nbricks = gx * gy * gz;
    t.nprocs  = nprocs;
    if (t.nbricks < 1 || nprocs < 1 || t.nbricks % nprocs || brick < 0 || brick >= t.nbricks) {
        g_err = "mdb_dd_plan: bad brick grid / process count";
        return -1;
    }
    const int n = t.slots(brick, send != 0, dir, peer);
    for (int k = 0; k < n; k++) owner[k] = t.owner(peer[k]);
    return n;
}
int mdb_dd_schedule(int gx, int gy, int gz, int perx, int pery, int perz, int nprocs, int proc, const int* cnt,
    int max_ops, int* ops)
{
    Topo t;
    t.g[0] = gx; t.g[1] = gy; t.g[2] = gz;
    t.periodic[0] = perx; t.periodic[1] = pery; t.periodic[2] = perz;
    t.nbricks = gx * gy * gz;
    t.nprocs  = nprocs;
    if (t.nbricks < 1 || nprocs < 1 || t.nbricks % nprocs || proc < 0 || proc >= nprocs || !cnt) {
        g_err = "mdb_dd_schedule: bad brick grid / process";
        return -1;
    }
    std::vector<Xfer> plan;
    dd_schedule(t, proc, cnt, plan);
    for (size_t k = 0; k < plan.size() && (int)k < max_ops; k++) {
        const Xfer& x = plan[k];
        const int v[7] = { x.kind, x.src, x.dst, x.src_start, x.dst_start, x.len, x.peer_proc };
        memcpy(ops + 7 * k, v, sizeof v);
    }
    return (int)plan.size();
}
mdb_dd* mdb_dd_create(const mdb_params* p, int gx, int gy, int gz, int nprocs, int proc, const void* nccl_id, int device)
{
    try {
        if (!p) throw Error("mdb_dd_create: null params");
        if (p->ntypes != 1) throw Error("mdb_dd_create: only ntypes == 1 is supported (EXPLICIT_TYPES off)");
        const int grid[3] = { gx, gy, gz };
        mdb_dd* d = new mdb_dd;
        d->g      = make_dd(*p, grid, nprocs, proc, nccl_id, device);
        return d;
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
mdb_dd* mdb_dd_create_cp(const mdb_params* p, int cluster_n, int gx, int gy, int gz, int nprocs, int proc, const void* nccl_id, int device)
{
    try {
        if (!p) throw Error("mdb_dd_create_cp: null params");
        if (p->ntypes != 1) throw Error("mdb_dd_create_cp: only ntypes == 1 is supported (EXPLICIT_TYPES off)");
        const int grid[3] = { gx, gy, gz };
        mdb_dd* d = new mdb_dd;
        d->g      = make_cp_dd(*p, cluster_n, grid, nprocs, proc, nccl_id, device);
        return d;
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
void mdb_dd_destroy(mdb_dd* d)
{
    if (!d) return;
    delete d->g;
    delete d;
}
int mdb_dd_setStream(mdb_dd* d, void* s) { MDB_DD_TRY(d->g->setStream((cudaStream_t)s)) }
int mdb_dd_sync(mdb_dd* d) { MDB_DD_TRY(d->g->sync()) }
long long mdb_dd_createAtom(mdb_dd* d)
{
    try {
        if (!d || !d->g) throw Error("null mdb_dd");
        return d->g->createAtom();
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}
int mdb_dd_setAtoms(mdb_dd* d, long long n, const int* tags, const void* x, const void* y, const void* z, const void* vx,
    const void* vy, const void* vz)
{
    MDB_DD_TRY(d->g->setAtoms(n, tags, x, y, z, vx, vy, vz))
}
int mdb_dd_setEam(mdb_dd* d, int nrho, double drho, int nr, double dr, double cut, double mass, const double* frho,
    const double* zr, const double* rhor)
{
    MDB_DD_TRY(d->g->setEam(nrho, drho, nr, dr, cut, mass, frho, zr, rhor))
}
int mdb_dd_setup(mdb_dd* d, int adjust) { MDB_DD_TRY(d->g->setup(adjust != 0)) }
int mdb_dd_reneighbour(mdb_dd* d) { MDB_DD_TRY(d->g->reneighbour()) }
int mdb_dd_run(mdb_dd* d, int nsteps, double* thermo_out, int max_records, int* nrecords, double* timers)
{
    MDB_DD_TRY(d->g->run(nsteps, thermo_out, max_records, nrecords, timers))
}
int mdb_dd_computeThermo(mdb_dd* d, double* T, double* P) { MDB_DD_TRY(d->g->computeThermo(T, P)) }
int mdb_dd_getCounts(mdb_dd* d, long long v[5]) { MDB_DD_TRY(d->g->getCounts(v)) }
int mdb_dd_getAtoms(mdb_dd* d, int which, int* tags, void* x, void* y, void* z) { MDB_DD_TRY(d->g->getAtoms(which, tags, x, y, z)) }
int mdb_dd_getNeighborTags(mdb_dd* d, int* tags, int* numneigh, int* rows, int stride)
{
    MDB_DD_TRY(d->g->getNeighborTags(tags, numneigh, rows, stride))
}
int mdb_dd_saveState(mdb_dd* d) { MDB_DD_TRY(d->g->saveState()) }
int mdb_dd_restoreState(mdb_dd* d) { MDB_DD_TRY(d->g->restoreState()) }
int mdb_dd_setOption(mdb_dd* d, const char* name, double v) { MDB_DD_TRY(d->g->setOption(name, v)) }
int mdb_dd_setTiming(mdb_dd* d, int on) { MDB_DD_TRY(d->g->setTiming(on != 0)) }
int mdb_dd_getKernelStats(mdb_dd* d, double* force_ms, long long* force_launches, double* neigh_ms,
    long long* neigh_launches, long long* total_launches, double* comm_ms)
{
    MDB_DD_TRY(d->g->stats(force_ms, force_launches, neigh_ms, neigh_launches, total_launches, comm_ms, false))
}
int mdb_dd_resetKernelStats(mdb_dd* d) { MDB_DD_TRY(d->g->stats(nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, true)) }

} // extern "C"
