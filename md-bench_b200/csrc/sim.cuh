// sim.cuh -- host-side state and orchestration of one simulation domain on one GPU.
// SimBase is the precision-erased interface the C ABI (mdb200.cu) talks to; Sim<real> implements it.
#pragma once
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/mdb200.h"
#include "mdb_util.cuh"

namespace mdb {

struct SimBase {
    virtual ~SimBase() {}
    virtual void setStream(cudaStream_t s)                                                   = 0;
    virtual void sync()                                                                      = 0;
    virtual long long createAtom()                                                           = 0;
    virtual void setAtoms(long long n, const void* x, const void* y, const void* z, const void* vx,
        const void* vy, const void* vz, const int* type, bool on_device)                     = 0;
    virtual void getAtoms(int which, bool ghosts, void* x, void* y, void* z)                 = 0;
    virtual void getTypes(int* types)                                                        = 0;
    virtual void getCounts(long long* c, int* maxneighs)                                     = 0;
    virtual void saveState()                                                                 = 0;
    virtual void restoreState()                                                              = 0;
    virtual void setupThermo()                                                               = 0;
    virtual void adjustThermo()                                                              = 0;
    virtual void computeThermo(double* T, double* P)                                         = 0;
    virtual void setupNeighbor()                                                             = 0;
    virtual void setupPbc()                                                                  = 0;
    virtual void updatePbc()                                                                 = 0;
    virtual void updateAtomsPbc()                                                            = 0;
    virtual void buildNeighbor()                                                             = 0;
    virtual double computeForce(int which)                                                   = 0;
    virtual void initialIntegrate()                                                          = 0;
    virtual void finalIntegrate()                                                            = 0;
    virtual void setup(bool adjust)                                                          = 0;
    virtual void reneighbour()                                                               = 0;
    virtual void run(int nsteps, double* thermo_out, int max_records, int* nrecords, double* timers) = 0;
    virtual void getNeighbors(int* numneigh, int* neighbors, int row_stride)                 = 0;
    virtual void getGhostMap(int* bm, int* px, int* py, int* pz)                             = 0;
    virtual void getNeighborParams(int* ints, double* reals)                                 = 0;
    virtual void getStencil(int* st)                                                         = 0;
    virtual void getBinCounts(int* bc)                                                       = 0;
    virtual void countPairs(long long* listed, long long* inside)                            = 0;
    virtual void setEam(int nrho, double drho, int nr, double dr, double cut, double mass,
        const double* frho, const double* zr, const double* rhor)                            = 0;
    virtual void setEamSplines(int nr, int nrho, int nr_tot, int nrho_tot, double rdr, double rdrho,
        const void* rhor, const void* frho, const void* z2r)                                 = 0;
    virtual void getEamSplines(int* nr, int* nrho, int* nr_tot, int* nrho_tot, double* rdr,
        double* rdrho, void* rhor, void* frho, void* z2r)                                    = 0;
    virtual void getEamFp(void* fp, bool ghosts)                                             = 0;
    virtual void setOption(const char* name, double value)                                   = 0;
    virtual void stubNeighbors(int pattern, int nneighs, int nreps, unsigned seed)           = 0;
    // Lazy operators (option "lazy_ops"): the reference's loop calls computeForce, finalIntegrate and initialIntegrate one
    // after the other through its function pointers (verletlist/main.c:258-273).  With lazy_ops the first two only record
    // that they are due; if the next call is initialIntegrate the three run as the ONE fused kernel mdb_run uses, else
    // flush_lazy() (called by the C ABI before every other entry point) launches them separately, in order.  Same results
    // bit for bit; computeForce then returns 0 s (nothing has been launched yet).
    virtual double abi_computeForce(int which) { return computeForce(which); }
    virtual void abi_finalIntegrate() { finalIntegrate(); }
    virtual void abi_initialIntegrate() { initialIntegrate(); }
    virtual void flush_lazy() {}
    virtual void drop_lazy() {} // forget pending work (the atoms it refers to are being replaced)
    virtual void invalidate_copies() {} // a call from outside may have changed positions: gather copies are stale

    bool timing             = false;
    double force_ms         = 0, neigh_ms = 0;
    long long force_launches = 0, neigh_launches = 0, launches = 0;
};

enum { FORCE_DISPATCH = 0, FORCE_LJ_FULL = 1, FORCE_LJ_HALF = 2, FORCE_EAM = 3 };

SimBase* make_sim(const mdb_params& p, int device);

// a spatially decomposed box: the bricks of this process + the transport between bricks (dd_group.cuh)
struct DDBase {
    virtual ~DDBase() {}
    virtual void setStream(cudaStream_t s)                                                                   = 0;
    virtual void sync()                                                                                      = 0;
    virtual long long createAtom()                                                                           = 0;
    virtual void setAtoms(long long n, const int* tags, const void* x, const void* y, const void* z, const void* vx,
        const void* vy, const void* vz)                                                                      = 0;
    virtual void setup(bool adjust)                                                                          = 0;
    virtual void reneighbour()                                                                               = 0;
    virtual void run(int nsteps, double* thermo_out, int max_records, int* nrecords, double* timers)         = 0;
    virtual void computeThermo(double* T, double* P)                                                         = 0;
    virtual void getCounts(long long* v)                                                                     = 0;
    virtual void getAtoms(int which, int* tags, void* x, void* y, void* z)                                   = 0;
    virtual void getNeighborTags(int* tags, int* numneigh, int* rows, int stride)                            = 0;
    virtual void saveState()                                                                                 = 0;
    virtual void restoreState()                                                                              = 0;
    virtual void setOption(const char* name, double v)                                                       = 0;
    virtual void setTiming(bool on)                                                                          = 0;
    virtual void stats(double* force_ms, long long* force_launches, double* neigh_ms, long long* neigh_launches,
        long long* launches, double* comm_ms, bool reset)                                                    = 0;
    virtual void setEam(int nrho, double drho, int nr, double dr, double cut, double mass, const double* frho,
        const double* zr, const double* rhor)                                                                = 0;
};

DDBase* make_dd(const mdb_params& global, const int grid[3], int nprocs, int proc, const void* nccl_id, int device);
// the clusterpair scheme on the same brick grid (cp_dd.cuh)
DDBase* make_cp_dd(const mdb_params& global, int cluster_n, const int grid[3], int nprocs, int proc, const void* nccl_id, int device);

} // namespace mdb
