// cp_sim.cu -- CpSim<real, N>: one CLUSTERPAIR simulation domain (GROMACS-style 4 x N cluster pairs,
// reference src/clusterpair/) resident on one B200, and its extern "C" boundary (mdb_cp_*, include/mdb200.h).
// Host code only orchestrates; every per-atom / per-cluster operation is a kernel in cp_kernels.cuh.  The
// device->host traffic inside the time loop is three small flag reads per rebuild (cluster count, ghost count,
// longest list row).
#include <algorithm>
#include <chrono>
#include <initializer_list>
#include <vector>

#include "../../include/mdb200.h"
#include "mdb_util.cuh"
#include "scan.cuh"
#include "vl_kernels.cuh"
#include "cp_kernels.cuh"
#include "sim.cuh"
#include "dd_topo.h"
#include "dd_kernels.cuh"
#include "nccl_dl.h"

namespace mdb {

struct CpBase {
    virtual ~CpBase() {}
    virtual void setStream(cudaStream_t s)                                                           = 0;
    virtual void sync()                                                                              = 0;
    virtual void setOption(const char* name, double v)                                               = 0;
    virtual long long createAtom()                                                                   = 0;
    virtual void setAtoms(long long n, const void* x, const void* y, const void* z, const void* vx, const void* vy,
        const void* vz)                                                                              = 0;
    virtual void getAtoms(int which, void* x, void* y, void* z, int* tag)                            = 0;
    virtual void getCounts(long long* v)                                                             = 0;
    virtual void setupThermo()                                                                       = 0;
    virtual void adjustThermo()                                                                      = 0;
    virtual void computeThermo(double* T, double* P)                                                 = 0;
    virtual void setupNeighbor()                                                                     = 0;
    virtual void buildClusters()                                                                     = 0;
    virtual void defineJClusters()                                                                   = 0;
    virtual void setupPbc()                                                                          = 0;
    virtual void binClusters()                                                                       = 0;
    virtual void buildNeighbor()                                                                     = 0;
    virtual void pruneNeighbor()                                                                     = 0;
    virtual void updateSingleAtoms()                                                                 = 0;
    virtual void updateAtomsPbc()                                                                    = 0;
    virtual void updatePbc(bool first)                                                               = 0;
    virtual double computeForce()                                                                    = 0;
    virtual void initialIntegrate()                                                                  = 0;
    virtual void finalIntegrate()                                                                    = 0;
    virtual void setup(bool adjust)                                                                  = 0;
    virtual void reneighbour()                                                                       = 0;
    virtual void run(int nsteps, double* thermo_out, int max_records, int* nrecords, double* timers) = 0;
    virtual void saveState()                                                                         = 0;
    virtual void restoreState()                                                                      = 0;
    virtual void countPairs(long long* cluster_pairs, long long* inside)                             = 0;
    virtual void getClusters(int which, int* natoms, void* bbox)                                     = 0;
    virtual void getClusterData(int which, void* out)                                                = 0;
    virtual void getClusterTags(int* tags)                                                           = 0;
    virtual void getClusterBins(int* b)                                                              = 0;
    virtual void getLists(int* nn, int* nm, int* nb, int stride)                                     = 0;
    virtual void getGhostMap(int* bm, int* px, int* py, int* pz)                                     = 0;
    virtual void getNeighborParams(int* I, double* R, int* st)                                       = 0;
    virtual void stub(int niclusters, int nat, int pattern, int nneighs, int nreps, int masked, unsigned seed) = 0;

    bool timing              = false;
    double force_ms          = 0, neigh_ms = 0;
    long long force_launches = 0, neigh_launches = 0, launches = 0;
};

template <class real, int N> struct CpSim final : CpBase {
    static constexpr int JFAC = N / CP_M; // i-clusters per tile (max(1, N / M))
    mdb_params P;
    int device;
    cudaStream_t stream = nullptr, own_stream = nullptr;
    // ---- Parameter fields in working precision (common/parameter.h:27-61) ----
    real epsilon, sigma, sigma6, temp, rho, mass, dt, dtforce, skin, cutforce, cutneigh, lattice, xprd, yprd, zprd;
    real dof_boltz = 1, t_scale = 1, p_scale = 1;
    bool thermo_ready = false, neigh_ready = false;
    int prune_every = 1000; // common/parameter.c:40
    int force_variant = 0;
    int sp_kernel = 2; // SP full lists: see launch_packed
    int ghost_epilogue = -1; // fused step writes the ghost tiles too: -1 = where it pays (launch_force), 0 / 1 = off / on
    bool ghosts_current = false, brick_mode = false; // brick_mode: a brick of a decomposed box (cp_dd.cuh): ghosts come from the neighbors
    bool fuse_force = true; // mdb_cp_run, full lists: integrate halves in the force kernel's epilogue (CpFused)
    // ---- atoms (clusterpair/atom.h:26-60) ----
    long long Natoms = 0;
    int Nlocal = 0;
    DBuf<real> x, y, z, vx, vy, vz, sx, sy, sz, svx, svy, svz, stage;
    DBuf<int> tag, type;
    int saved_n = 0;
    // ---- clusters ----
    int ncl = 0, ncj = 0, nghost = 0, dummy_cj = 0; // Nclusters_local, local tiles, Nclusters_ghost
    DBuf<real> cl_x, cl_v, cl_f, ibb, jbb, jbbs, pmaxz; // jbbs: the j bounding boxes in bin order (k_cp_cluster_sort)
    DBuf<real> cl_xn; // second cluster position array of the fused force + integrate step (run())
    DBuf<int> cl_tag, inat, ibin, jnat, atom_off;
    // ---- bins (clusterpair/neighbor.c:26-45) ----
    CpGeom<real> g {};
    real binsizex = 0, binsizey = 0;
    int nstencil  = 0;
    std::vector<int> h_stencil;
    DBuf<int> stencil, atom_bin, bincount, binstart, cursor, binatoms, nclbin, clbase;
    DBuf<int> cbin, cbincount, cbinstart, ccursor, cbinlist, cbinlist2;
    // ---- ghosts (clusterpair/pbc.c) ----
    DBuf<unsigned> gmask;
    DBuf<int> gcnt, goff, border_map, code;
    // ---- lists (clusterpair/neighbor.h:31-40) ----
    int maxneighs = 100; // neighbor.c:65
    DBuf<int> numneigh, numneigh_masked, neighbors;
    DBuf<char> pos4; // {x, y, z, -} per cluster slot for the force kernel (k_cp_pack_j)
    bool lists_ready = false;
    // ---- scratch ----
    Scanner scanner;
    int* h_flags  = nullptr;
    double* h_red = nullptr;
    DBuf<int> d_flags;
    DBuf<double> d_partial, d_red, d_thermo;
    DBuf<unsigned long long> d_cnt;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evA = nullptr, evB = nullptr, evR0 = nullptr, evR1 = nullptr;

    CpSim(const mdb_params& p, int dev) : P(p), device(dev)
    {
        MDB_CUDA(cudaSetDevice(device));
        MDB_CUDA(cudaStreamCreateWithFlags(&own_stream, cudaStreamNonBlocking));
        stream = own_stream;
        MDB_CUDA(cudaMallocHost(&h_flags, 16 * sizeof(int)));
        MDB_CUDA(cudaMallocHost(&h_red, 16 * sizeof(double)));
        for (cudaEvent_t* e : { &ev0, &ev1, &evA, &evB, &evR0, &evR1 }) MDB_CUDA(cudaEventCreate(e));
        scanner.launches = &launches;
        d_flags.ensure(16, false, stream);
        d_red.ensure(16, false, stream);
        d_partial.ensure(RED_BLOCKS * 4, false, stream);
        d_cnt.ensure(4, false, stream);
        derive();
        preload_kernels();
    }
    void preload_kernels() // the kernels of the time loop, loaded now instead of lazily inside a short timed run (see Sim::preload_kernels)
    {
        cudaFuncAttributes a;
        const void* ks[] = { (const void*)k_cp_force_lj<real, N, false>, (const void*)k_cp_force_lj<real, N, false, true>,
            (const void*)k_cp_force_lj<real, N, true>, (const void*)k_cp_force_jl<real, N, true>, (const void*)k_cp_integrate<real, N, 0>,
            (const void*)k_cp_integrate<real, N, 1>, (const void*)k_cp_integrate<real, N, 2>, (const void*)k_cp_update_pbc<real, N>,
            (const void*)k_cp_update_pbc_first<real, N>, (const void*)k_cp_update_single_atoms<real, N>, (const void*)k_cp_bin_count<real>,
            (const void*)k_cp_sort_emit<real, N>, (const void*)k_cp_define_j<real, N>, (const void*)k_cp_ghost_count<real>,
            (const void*)k_cp_ghost_fill<N>, (const void*)k_cp_cluster_bin<real, N>, (const void*)k_cp_cluster_fill,
            (const void*)k_cp_cluster_sort<real>, (const void*)k_cp_build_neighbor<real, N>, (const void*)k_cp_clusters_per_bin<N>,
            (const void*)k_update_atoms_pbc<real>, (const void*)k_vel_partial<real>, (const void*)k_vel_final, (const void*)k_cp_bin_fill };
        for (const void* k : ks) MDB_CUDA(cudaFuncGetAttributes(&a, k));
        if (sizeof(real) == 4) {
            MDB_CUDA(cudaFuncGetAttributes(&a, (const void*)k_cp_force_lj_sp_duo<N, true, false>));
            MDB_CUDA(cudaFuncGetAttributes(&a, (const void*)k_cp_force_lj_sp_duo<N, false, false>));
        }
        scanner.preload();
    }
    ~CpSim() override
    {
        cudaSetDevice(device);
        cudaStreamSynchronize(stream);
        if (phases && phase_calls)
            fprintf(stderr, "[mdb_cp] reneighbour x%d, ms per call: updateSingleAtoms %.3f updateAtomsPbc %.3f buildClusters %.3f "
                            "defineJClusters %.3f setupPbc %.3f binClusters %.3f buildNeighbor %.3f\n", phase_calls,
                phase_ms[0] / phase_calls, phase_ms[1] / phase_calls, phase_ms[2] / phase_calls, phase_ms[3] / phase_calls,
                phase_ms[4] / phase_calls, phase_ms[5] / phase_calls, phase_ms[6] / phase_calls);
        for (DBuf<real>* b : { &x, &y, &z, &vx, &vy, &vz, &sx, &sy, &sz, &svx, &svy, &svz, &stage, &cl_x, &cl_xn, &cl_v, &cl_f, &ibb,
                 &jbb, &jbbs, &pmaxz })
            b->release();
        for (DBuf<int>* b : { &tag, &type, &cl_tag, &inat, &ibin, &jnat, &atom_off, &stencil, &atom_bin, &bincount, &binstart,
                 &cursor, &binatoms, &nclbin, &clbase, &cbin, &cbincount, &cbinstart, &ccursor, &cbinlist, &cbinlist2, &gcnt,
                 &goff, &border_map, &code, &numneigh, &numneigh_masked, &neighbors, &d_flags })
            b->release();
        gmask.release();
        pos4.release();
        d_partial.release(); d_red.release(); d_thermo.release(); d_cnt.release();
        scanner.release();
        cudaFreeHost(h_flags);
        cudaFreeHost(h_red);
        for (cudaEvent_t e : { ev0, ev1, evA, evB, evR0, evR1 }) cudaEventDestroy(e);
        cudaStreamDestroy(own_stream);
    }

    // initParameter + command line (common/parameter.c:16-51, clusterpair/main.c:216, 46-49)
    void derive()
    {
        epsilon  = (real)P.epsilon;
        sigma    = (real)P.sigma;
        real s2  = sigma * sigma;
        sigma6   = s2 * s2 * s2;
        temp     = (real)P.temp;
        rho      = (real)P.rho;
        mass     = (real)P.mass;
        dt       = (real)P.dt;
        dtforce  = (real)(0.5 * (double)dt);
        skin     = (real)P.skin;
        cutforce = (real)P.cutforce;
        cutneigh = cutforce + skin;
        lattice  = (real)pow((4.0 / (double)rho), (1.0 / 3.0));
        if (P.from_input) { // box of an input file: the readers set param->xprd = xhi - xlo, and setupNeighbor takes it
            xprd = (real)P.xhi - (real)P.xlo; // (neighbor.c:78-82); the box is then treated as [0, prd) like the reference does
            yprd = (real)P.yhi - (real)P.ylo;
            zprd = (real)P.zhi - (real)P.zlo;
        } else {
            xprd = P.nx * lattice; yprd = P.ny * lattice; zprd = P.nz * lattice;
        }
    }
    void setStream(cudaStream_t s) override
    {
        MDB_CUDA(cudaStreamSynchronize(stream));
        stream = s ? s : own_stream;
    }
    void sync() override { MDB_CUDA(cudaStreamSynchronize(stream)); }
    void setOption(const char* name, double v) override
    {
        if (!strcmp(name, "prune_every")) prune_every = (int)v;
        else if (!strcmp(name, "force_variant")) force_variant = (int)v;
        else if (!strcmp(name, "sp_kernel")) sp_kernel = (int)v;
        else if (!strcmp(name, "ghost_epilogue")) ghost_epilogue = (int)v;
        else if (!strcmp(name, "fuse_force")) fuse_force = v != 0;
        else throw Error(fmt("mdb_cp_setOption: unknown option '%s'", name));
    }

    // ------------------------------------------------------------------ atoms
    void ensure_atoms(size_t n)
    {
        for (DBuf<real>* b : { &x, &y, &z, &vx, &vy, &vz }) b->ensure(n, false, stream);
        tag.ensure(n, false, stream);
        type.ensure(n, false, stream);
    }
    long long createAtom() override // clusterpair/atom.c:49-180 (same generator as verletlist/atom.c:67-187)
    {
        if (P.from_input) throw Error("mdb_cp_createAtom: from_input is set; hand the atoms over with mdb_cp_setAtoms");
        derive();
        Natoms = 4LL * P.nx * P.ny * P.nz;
        if (Natoms > 1500000000LL) throw Error("createAtom: too many atoms for one domain");
        Nlocal = (int)Natoms;
        ensure_atoms(Nlocal);
        MDB_LAUNCH(launches, k_create_atoms<real>, grid_for(Natoms, 256), 256, 0, stream, P.nx, P.ny, P.nz, lattice, x.p,
            y.p, z.p, vx.p, vy.p, vz.p, type.p);
        MDB_LAUNCH(launches, k_iota, grid_for(Nlocal, 256), 256, 0, stream, Nlocal, tag.p);
        neigh_ready = lists_ready = false;
        return Natoms;
    }
    void setAtoms(long long n, const void* ax, const void* ay, const void* az, const void* avx, const void* avy,
        const void* avz) override
    {
        if (n <= 0 || n > 1500000000LL) throw Error("setAtoms: bad atom count");
        derive();
        Natoms = n;
        Nlocal = (int)n;
        ensure_atoms(n);
        if (P.layout == MDB_AOS) { // atom_x(i) = x[3i] (clusterpair/atom.h:66-70)
            stage.ensure(3 * n, false, stream);
            MDB_CUDA(cudaMemcpyAsync(stage.p, ax, 3 * n * sizeof(real), cudaMemcpyHostToDevice, stream));
            MDB_LAUNCH(launches, k_aos_to_soa<real>, grid_for(n, 256), 256, 0, stream, (size_t)n, (const real*)stage.p, x.p,
                y.p, z.p);
        } else {
            MDB_CUDA(cudaMemcpyAsync(x.p, ax, n * sizeof(real), cudaMemcpyHostToDevice, stream));
            MDB_CUDA(cudaMemcpyAsync(y.p, ay, n * sizeof(real), cudaMemcpyHostToDevice, stream));
            MDB_CUDA(cudaMemcpyAsync(z.p, az, n * sizeof(real), cudaMemcpyHostToDevice, stream));
        }
        if (avx) { // velocities are always SoA (clusterpair/atom.h:72-92)
            MDB_CUDA(cudaMemcpyAsync(vx.p, avx, n * sizeof(real), cudaMemcpyHostToDevice, stream));
            MDB_CUDA(cudaMemcpyAsync(vy.p, avy, n * sizeof(real), cudaMemcpyHostToDevice, stream));
            MDB_CUDA(cudaMemcpyAsync(vz.p, avz, n * sizeof(real), cudaMemcpyHostToDevice, stream));
        } else {
            for (DBuf<real>* b : { &vx, &vy, &vz }) MDB_CUDA(cudaMemsetAsync(b->p, 0, n * sizeof(real), stream));
        }
        MDB_CUDA(cudaMemsetAsync(type.p, 0, n * sizeof(int), stream));
        MDB_LAUNCH(launches, k_iota, grid_for(Nlocal, 256), 256, 0, stream, Nlocal, tag.p);
        MDB_CUDA(cudaStreamSynchronize(stream));
        neigh_ready = lists_ready = false;
    }
    void getAtoms(int which, void* ax, void* ay, void* az, int* atag) override
    {
        const size_t n = Nlocal;
        if (which == 'x') {
            if (P.layout == MDB_AOS) {
                stage.ensure(3 * n, false, stream);
                MDB_LAUNCH(launches, k_soa_to_aos<real>, grid_for(n, 256), 256, 0, stream, n, x.p, y.p, z.p, stage.p);
                MDB_CUDA(cudaMemcpyAsync(ax, stage.p, 3 * n * sizeof(real), cudaMemcpyDeviceToHost, stream));
            } else {
                MDB_CUDA(cudaMemcpyAsync(ax, x.p, n * sizeof(real), cudaMemcpyDeviceToHost, stream));
                MDB_CUDA(cudaMemcpyAsync(ay, y.p, n * sizeof(real), cudaMemcpyDeviceToHost, stream));
                MDB_CUDA(cudaMemcpyAsync(az, z.p, n * sizeof(real), cudaMemcpyDeviceToHost, stream));
            }
        } else if (which == 'v') {
            MDB_CUDA(cudaMemcpyAsync(ax, vx.p, n * sizeof(real), cudaMemcpyDeviceToHost, stream));
            MDB_CUDA(cudaMemcpyAsync(ay, vy.p, n * sizeof(real), cudaMemcpyDeviceToHost, stream));
            MDB_CUDA(cudaMemcpyAsync(az, vz.p, n * sizeof(real), cudaMemcpyDeviceToHost, stream));
        } else throw Error("mdb_cp_getAtoms: which must be 'x' or 'v'");
        if (atag) MDB_CUDA(cudaMemcpyAsync(atag, tag.p, n * sizeof(int), cudaMemcpyDeviceToHost, stream));
        MDB_CUDA(cudaStreamSynchronize(stream));
    }
    void getCounts(long long* v) override
    {
        long long ghost_atoms = 0;
        if (nghost > 0) {
            std::vector<int> h(nghost);
            MDB_CUDA(cudaMemcpyAsync(h.data(), jnat.p + ncj, nghost * sizeof(int), cudaMemcpyDeviceToHost, stream));
            MDB_CUDA(cudaStreamSynchronize(stream));
            for (int q : h) ghost_atoms += q;
        }
        v[0] = Natoms; v[1] = Nlocal; v[2] = ghost_atoms; v[3] = ncl; v[4] = nghost; v[5] = dummy_cj; v[6] = maxneighs;
        v[7] = ncj;
    }
    void saveState() override
    {
        const size_t n = Nlocal;
        DBuf<real>* src[] = { &x, &y, &z, &vx, &vy, &vz };
        DBuf<real>* dst[] = { &sx, &sy, &sz, &svx, &svy, &svz };
        for (int k = 0; k < 6; k++) {
            dst[k]->ensure(n, false, stream);
            MDB_CUDA(cudaMemcpyAsync(dst[k]->p, src[k]->p, n * sizeof(real), cudaMemcpyDeviceToDevice, stream));
        }
        saved_n = Nlocal;
    }
    void restoreState() override
    {
        if (!saved_n) throw Error("restoreState: nothing saved");
        const size_t n = saved_n;
        DBuf<real>* dst[] = { &x, &y, &z, &vx, &vy, &vz };
        DBuf<real>* src[] = { &sx, &sy, &sz, &svx, &svy, &svz };
        for (int k = 0; k < 6; k++)
            MDB_CUDA(cudaMemcpyAsync(dst[k]->p, src[k]->p, n * sizeof(real), cudaMemcpyDeviceToDevice, stream));
        MDB_LAUNCH(launches, k_iota, grid_for(Nlocal, 256), 256, 0, stream, Nlocal, tag.p);
        lists_ready = false;
    }

    // ------------------------------------------------------------------ thermo (common/thermo.c)
    void setupThermo() override // thermo.c:30-53 (LJ)
    {
        dof_boltz    = (real)(Natoms * 3 - 3);
        t_scale      = (real)1.0 / dof_boltz;
        p_scale      = (real)(1.0 / 3 / (double)xprd / (double)yprd / (double)zprd);
        thermo_ready = true;
    }
    void vel_sums(double* out)
    {
        const int nb = (int)std::min<size_t>(RED_BLOCKS, grid_for(Nlocal, RED_THREADS));
        MDB_LAUNCH(launches, k_vel_partial<real>, nb, RED_THREADS, 0, stream, Nlocal, vx.p, vy.p, vz.p, mass, d_partial.p);
        MDB_LAUNCH(launches, k_vel_final, 1, RED_THREADS, 0, stream, nb, d_partial.p, out);
    }
    void read_red()
    {
        MDB_CUDA(cudaMemcpyAsync(h_red, d_red.p, 4 * sizeof(double), cudaMemcpyDeviceToHost, stream));
        MDB_CUDA(cudaStreamSynchronize(stream));
    }
    void thermo_from_sum(double msum, double* T, double* Pr)
    {
        real t = (real)msum;
        t      = t * t_scale;
        real p = (t * dof_boltz) * p_scale;
        *T     = t;
        *Pr    = p;
    }
    void computeThermo(double* T, double* Pr) override // thermo.c:55-80, from the ATOM arrays
    {
        if (!thermo_ready) setupThermo();
        vel_sums(d_red.p);
        read_red();
        thermo_from_sum(h_red[3], T, Pr);
    }
    void adjustThermo() override // thermo.c:82-122
    {
        if (!thermo_ready) setupThermo();
        vel_sums(d_red.p);
        read_red();
        const real vxtot = (real)h_red[0] / (real)Natoms, vytot = (real)h_red[1] / (real)Natoms,
                   vztot = (real)h_red[2] / (real)Natoms;
        MDB_LAUNCH(launches, k_vel_shift<real>, grid_for(Nlocal, 256), 256, 0, stream, Nlocal, vx.p, vy.p, vz.p, vxtot, vytot,
            vztot);
        vel_sums(d_red.p);
        read_red();
        real t = (real)h_red[3];
        t *= t_scale;
        const real factor = (real)sqrt((double)(temp / t));
        MDB_LAUNCH(launches, k_vel_scale<real>, grid_for(Nlocal, 256), 256, 0, stream, Nlocal, vx.p, vy.p, vz.p, factor);
    }

    // ------------------------------------------------------------------ neighbor geometry
    real bindist(int i, int j) const // clusterpair/neighbor.c:565-580
    {
        real delx = i > 0 ? (i - 1) * binsizex : (i == 0 ? (real)0.0 : (i + 1) * binsizex);
        real dely = j > 0 ? (j - 1) * binsizey : (j == 0 ? (real)0.0 : (j + 1) * binsizey);
        return (delx * delx + dely * dely);
    }
    static real rcbrt(real v) { return sizeof(real) == 4 ? (real)cbrtf((float)v) : (real)cbrt((double)v); }
    static real rceil(real v) { return sizeof(real) == 4 ? (real)ceilf((float)v) : (real)ceil((double)v); }
    void setupNeighbor() override // clusterpair/neighbor.c:70-172
    {
        if (Nlocal <= 0) throw Error("mdb_cp_setupNeighbor: no atoms (the bin size depends on the atom density)");
        const real SMALL = (real)1.0e-6, FACTOR = (real)0.999;
        const real xlo = 0, xhi = xprd, ylo = 0, yhi = yprd;
        // neighbor.c:93-98 as the reference build evaluates it (-Ofast folds the divisions):
        // nbin = ceil(prd * cbrt(density * (1 / atoms_in_cell))), all in MD_FLOAT
        const real atom_density   = ((real)Nlocal) / ((xprd * yprd) * zprd);
        const real inv_targetsize = rcbrt(atom_density * ((real)1.0 / (real)(N > CP_M ? N : CP_M)));
        g.nbinx  = std::max(1, (int)rceil(xprd * inv_targetsize));
        g.nbiny  = std::max(1, (int)rceil(yprd * inv_targetsize));
        binsizex = (xhi - xlo) / g.nbinx;
        binsizey = (yhi - ylo) / g.nbiny;
        g.bininvx = (real)(1.0 / (double)binsizex);
        g.bininvy = (real)(1.0 / (double)binsizey);
        g.cutneigh   = cutneigh;
        g.cutneighsq = cutneigh * cutneigh;
        real coord;
        int mhix, mhiy;
        coord     = xlo - cutneigh - SMALL * xprd;
        g.mbinxlo = (int)(coord * g.bininvx);
        if (coord < (real)0.0) g.mbinxlo -= 1;
        coord = xhi + cutneigh + SMALL * xprd;
        mhix  = (int)(coord * g.bininvx);
        coord     = ylo - cutneigh - SMALL * yprd;
        g.mbinylo = (int)(coord * g.bininvy);
        if (coord < (real)0.0) g.mbinylo -= 1;
        coord = yhi + cutneigh + SMALL * yprd;
        mhiy  = (int)(coord * g.bininvy);
        g.mbinxlo -= 1; mhix += 1; g.mbinx = mhix - g.mbinxlo + 1;
        g.mbinylo -= 1; mhiy += 1; g.mbiny = mhiy - g.mbinylo + 1;
        int nextx = (int)(cutneigh * g.bininvx), nexty = (int)(cutneigh * g.bininvy);
        if (nextx * binsizex < FACTOR * cutneigh) nextx++;
        if (nexty * binsizey < FACTOR * cutneigh) nexty++;
        h_stencil.clear();
        for (int j = -nexty; j <= nexty; j++)
            for (int i = -nextx; i <= nextx; i++)
                if (bindist(i, j) < g.cutneighsq) h_stencil.push_back(j * g.mbinx + i);
        nstencil = (int)h_stencil.size();
        g.mbins  = g.mbinx * g.mbiny;
        g.xprd = xprd; g.yprd = yprd; g.zprd = zprd;
        // buildNeighborCPU's rbb_sq (neighbor.c:276-279), evaluated like the C expression (double intermediates)
        const real bbx = (real)(0.5 * (double)(binsizex + binsizex)), bby = (real)(0.5 * (double)(binsizey + binsizey));
        real r         = (real)std::max(0.0, (double)cutneigh - 0.5 * sqrt((double)(bbx * bbx + bby * bby)));
        g.rbb_sq       = r * r;
        stencil.ensure(nstencil, false, stream);
        MDB_CUDA(cudaMemcpyAsync(stencil.p, h_stencil.data(), nstencil * sizeof(int), cudaMemcpyHostToDevice, stream));
        MDB_CUDA(cudaStreamSynchronize(stream));
        for (DBuf<int>* b : { &bincount, &binstart, &cursor, &nclbin, &clbase, &cbincount, &cbinstart, &ccursor })
            b->ensure(g.mbins + 2, false, stream);
        neigh_ready = true;
    }

    // ------------------------------------------------------------------ clusters
    void ensure_tiles(size_t tiles, bool keep)
    {
        const size_t nt = tiles + tiles / 8 + 64;
        if (tiles * 3 * N <= cl_x.cap) return;
        const size_t oldf = cl_f.cap;
        cl_x.ensure(nt * 3 * N, keep, stream);
        cl_v.ensure(nt * 3 * N, keep, stream);
        cl_f.ensure(nt * 3 * N, keep, stream);
        MDB_CUDA(cudaMemsetAsync(cl_f.p + (keep ? oldf : 0), 0, (cl_f.cap - (keep ? oldf : 0)) * sizeof(real), stream));
        cl_tag.ensure(nt * N, keep, stream);
        jnat.ensure(nt, keep, stream);
        jbb.ensure(nt * 6, keep, stream);
    }
    void buildClusters() override // neighbor.c:599-753: binAtoms, sortAtomsByZCoord, buildClusters
    {
        if (!neigh_ready) setupNeighbor();
        const int n = Nlocal;
        atom_bin.ensure(n, false, stream);
        binatoms.ensure(n, false, stream);
        MDB_CUDA(cudaMemsetAsync(bincount.p, 0, (g.mbins + 1) * sizeof(int), stream));
        MDB_CUDA(cudaMemsetAsync(cursor.p, 0, (g.mbins + 1) * sizeof(int), stream));
        MDB_CUDA(cudaMemsetAsync(d_flags.p, 0, 4 * sizeof(int), stream));
        MDB_LAUNCH(launches, k_cp_bin_count<real>, grid_for(n, 256), 256, 0, stream, n, g, x.p, y.p, atom_bin.p, bincount.p);
        scanner.exclusive(bincount.p, binstart.p, g.mbins, binstart.p + g.mbins, stream);
        MDB_LAUNCH(launches, k_cp_bin_fill, grid_for(n, 256), 256, 0, stream, n, atom_bin.p, binstart.p, cursor.p, binatoms.p);
        MDB_LAUNCH(launches, k_cp_clusters_per_bin<N>, grid_for(g.mbins, 256), 256, 0, stream, g.mbins, bincount.p, nclbin.p,
            d_flags.p + 0);
        scanner.exclusive(nclbin.p, clbase.p, g.mbins, d_flags.p + 1, stream);
        MDB_CUDA(cudaMemcpyAsync(h_flags, d_flags.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream));
        MDB_CUDA(cudaStreamSynchronize(stream));
        const int maxcount = h_flags[0];
        ncl    = h_flags[1];
        ncj    = ncl / JFAC;
        nghost = 0;
        dummy_cj = ncj;
        ensure_tiles((size_t)ncj + ncj / 3 + 1, false);
        inat.ensure(ncl, false, stream);
        ibin.ensure(ncl, false, stream);
        ibb.ensure((size_t)ncl * 6, false, stream);
        const size_t smem = (size_t)(maxcount + 4) * (2 * sizeof(real) + 3 * sizeof(int));
        if (smem > 200 * 1024) throw Error("buildClusters: a bin column holds too many atoms for one thread block");
        if (smem > 48 * 1024)
            MDB_CUDA(cudaFuncSetAttribute(k_cp_sort_emit<real, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MDB_LAUNCH(launches, (k_cp_sort_emit<real, N>), g.mbins, 128, smem, stream, g.mbins, binstart.p, binatoms.p, clbase.p,
            x.p, y.p, z.p, vx.p, vy.p, vz.p, tag.p, cl_x.p, cl_v.p, cl_tag.p, inat.p, ibb.p, ibin.p);
        lists_ready = false;
    }
    void defineJClusters() override // neighbor.c:755-873
    {
        MDB_LAUNCH(launches, (k_cp_define_j<real, N>), grid_for(ncj, 256), 256, 0, stream, ncj, inat.p, ibb.p, jnat.p, jbb.p);
    }
    PbcGeom<real> pbc_geom() const
    {
        PbcGeom<real> q;
        q.xprd = xprd; q.yprd = yprd; q.zprd = zprd; q.cutneigh = cutneigh;
        q.xhi_cut = xprd - cutneigh; q.yhi_cut = yprd - cutneigh; q.zhi_cut = zprd - cutneigh;
        q.pbc_x = q.pbc_y = q.pbc_z = 1; // clusterpair/pbc.c ignores pbc_x/y/z
        return q;
    }
    void setupPbc() override // pbc.c:183-323
    {
        gmask.ensure(ncj, false, stream);
        gcnt.ensure(ncj, false, stream);
        goff.ensure(ncj, false, stream);
        MDB_LAUNCH(launches, k_cp_ghost_count<real>, grid_for(ncj, 256), 256, 0, stream, ncj, pbc_geom(), jnat.p, jbb.p,
            gmask.p, gcnt.p);
        scanner.exclusive(gcnt.p, goff.p, ncj, d_flags.p + 0, stream);
        MDB_CUDA(cudaMemcpyAsync(h_flags, d_flags.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
        MDB_CUDA(cudaStreamSynchronize(stream));
        nghost   = h_flags[0];
        dummy_cj = ncj + nghost;
        ghosts_current = false;
        ensure_tiles((size_t)ncj + nghost + 1, true);
        border_map.ensure(nghost + 1, false, stream);
        code.ensure(nghost + 1, false, stream);
        MDB_LAUNCH(launches, k_cp_ghost_fill<N>, grid_for(ncj, 256), 256, 0, stream, ncj, gmask.p, goff.p, cl_tag.p,
            border_map.p, code.p, jnat.p, cl_tag.p);
        updatePbc(true);
    }
    void updatePbc(bool first) override // pbc.c:45-114
    {
        if (first) {
            MDB_LAUNCH(launches, (k_cp_update_pbc_first<real, N>), grid_for(nghost + 1, 128), 128, 0, stream, ncj, nghost, xprd,
                yprd, zprd, border_map.p, code.p, jnat.p, cl_x.p, jbb.p);
        } else if (nghost > 0 && !ghosts_current) { // (the fused kernel may have written the images of the new positions already)
            MDB_LAUNCH(launches, (k_cp_update_pbc<real, N>), grid_for((size_t)nghost * N, 256), 256, 0, stream, ncj, nghost, xprd,
                yprd, zprd, border_map.p, code.p, jnat.p, cl_x.p);
        }
    }
    void updateAtomsPbc() override // pbc.c:117-144 (same wrap as verletlist/pbc.c:59-84)
    {
        MDB_LAUNCH(launches, k_update_atoms_pbc<real>, grid_for(Nlocal, 256), 256, 0, stream, Nlocal, xprd, yprd, zprd, x.p, y.p,
            z.p);
    }
    void binClusters() override // neighbor.c:875-1021
    {
        const int nt = ncj + nghost;
        cbin.ensure(nt, false, stream);
        cbinlist.ensure(nt, false, stream);
        cbinlist2.ensure(nt, false, stream);
        pmaxz.ensure(nt, false, stream);
        jbbs.ensure((size_t)nt * 6, false, stream);
        MDB_CUDA(cudaMemsetAsync(cbincount.p, 0, (g.mbins + 1) * sizeof(int), stream));
        MDB_CUDA(cudaMemsetAsync(ccursor.p, 0, (g.mbins + 1) * sizeof(int), stream));
        MDB_LAUNCH(launches, (k_cp_cluster_bin<real, N>), grid_for(nt, 256), 256, 0, stream, ncj, nghost, g, ibin.p, code.p,
            jnat.p, cl_x.p, cbin.p, cbincount.p);
        scanner.exclusive(cbincount.p, cbinstart.p, g.mbins, cbinstart.p + g.mbins, stream);
        MDB_LAUNCH(launches, k_cp_cluster_fill, grid_for(nt, 256), 256, 0, stream, nt, cbin.p, cbinstart.p, ccursor.p,
            cbinlist2.p);
        MDB_LAUNCH(launches, k_cp_cluster_sort<real>, grid_for((size_t)g.mbins * 32, 128), 128, 0, stream, g.mbins, cbinstart.p,
            cbinlist2.p, cbinlist.p, jbb.p, jbbs.p, pmaxz.p);
    }
    void buildNeighbor() override // buildNeighborCPU, neighbor.c:262-481
    {
        if (timing) MDB_CUDA(cudaEventRecord(evA, stream));
        numneigh.ensure(ncl, false, stream);
        numneigh_masked.ensure(ncl, false, stream);
        for (;;) {
            neighbors.ensure((size_t)ncl * maxneighs, false, stream);
            MDB_CUDA(cudaMemsetAsync(d_flags.p + 2, 0, sizeof(int), stream));
            MDB_LAUNCH(launches, (k_cp_build_neighbor<real, N>), grid_for(ncl, 128), 128, 0, stream, ncl, P.half_neigh, g,
                stencil.p, nstencil, ibin.p, inat.p, ibb.p, jnat.p, jbbs.p, cl_x.p, cbinstart.p, cbinlist.p, pmaxz.p, maxneighs,
                numneigh.p, numneigh_masked.p, neighbors.p, d_flags.p + 2);
            neigh_launches++;
            MDB_CUDA(cudaMemcpyAsync(h_flags + 2, d_flags.p + 2, sizeof(int), cudaMemcpyDeviceToHost, stream));
            MDB_CUDA(cudaStreamSynchronize(stream));
            if (h_flags[2] >= maxneighs) { // neighbor.c:412-428
                maxneighs = (int)(h_flags[2] * 1.2);
                continue;
            }
            break;
        }
        lists_ready = true;
        if (timing) {
            float ms = 0;
            MDB_CUDA(cudaEventRecord(evB, stream));
            MDB_CUDA(cudaEventSynchronize(evB));
            MDB_CUDA(cudaEventElapsedTime(&ms, evA, evB));
            neigh_ms += ms;
        }
    }
    void pruneNeighbor() override // neighbor.c:483-531
    {
        if (!lists_ready) throw Error("pruneNeighbor: no cluster-pair list");
        MDB_LAUNCH(launches, (k_cp_prune<real, N>), grid_for(ncl, 128), 128, 0, stream, ncl, g.cutneighsq, inat.p, jnat.p,
            cl_x.p, maxneighs, numneigh.p, numneigh_masked.p, neighbors.p);
    }
    void updateSingleAtoms() override // neighbor.c:1023-1049
    {
        if (ncl == 0) return;
        atom_off.ensure(ncl, false, stream);
        scanner.exclusive(inat.p, atom_off.p, ncl, nullptr, stream);
        MDB_LAUNCH(launches, (k_cp_update_single_atoms<real, N>), grid_for((size_t)ncl * CP_M, 256), 256, 0, stream, ncl, inat.p,
            atom_off.p, cl_x.p, cl_v.p, cl_tag.p, x.p, y.p, z.p, vx.p, vy.p, vz.p, tag.p);
    }

    // ------------------------------------------------------------------ force / integrate
    // fused = computeForce(n) + finalIntegrate(n) + initialIntegrate(n+1) in one launch (run() only): the new positions
    // go to cl_xn, which must already mirror cl_x's padding slots, ghost tiles and dummy tile (sync_second_array), and the
    // two arrays are swapped afterwards.
    bool can_fuse_force() const
    {
        if (!fuse_force || P.half_neigh) return false;
        if (sizeof(real) == 8 && N == 8) return false; // the DP 4x8 kernel would spill with the epilogue (ptxas: 64 regs + 12 B)
        int fv = force_variant;
        if (fv == 0) fv = sizeof(real) == 4 ? 2 : 1;
        return fv == 1 || fv == 2;
    }
    void sync_second_array()
    {
        cl_xn.ensure(cl_x.cap, false, stream);
        MDB_CUDA(cudaMemcpyAsync(cl_xn.p, cl_x.p, (size_t)(ncj + nghost + 1) * 3 * N * sizeof(real), cudaMemcpyDeviceToDevice, stream));
    }
    void launch_force(bool fused = false)
    {
        NvtxRange nvtx_range_("force");
        if (!lists_ready) throw Error("computeForce: no cluster-pair list (call mdb_cp_buildNeighbor first)");
        if (timing) MDB_CUDA(cudaEventRecord(ev0, stream));
        LJConst2<real> c2 { cutforce * cutforce, (real)48.0 * epsilon * sigma6 * sigma6, (real)24.0 * epsilon * sigma6 };
        const unsigned grid = grid_for((size_t)ncl * CP_M, 128);
        // small single domains: the ghost tiles of the next step are written by the same launch (one launch per step; at 128^3
        // the extra mask loads cost more than the kernel they replace, as for the verletlist kernel)
        const bool want_ghosts = ghost_epilogue < 0 ? Nlocal <= (1 << 19) : ghost_epilogue != 0;
        const bool own_ghosts  = fused && want_ghosts && !brick_mode && nghost > 0;
        const CpFused<real> fi { cl_v.p, cl_xn.p, dtforce, dt, own_ghosts ? gmask.p : (const unsigned*)nullptr,
            own_ghosts ? goff.p : (const int*)nullptr, ncj, xprd, yprd, zprd };
        if (fused) {
            int fv = force_variant;
            if (fv == 0) fv = sizeof(real) == 4 ? 2 : 1;
            if (fv == 2) launch_packed_fused(grid, c2, fi);
            else
                MDB_LAUNCH(launches, (k_cp_force_lj<real, N, false, true>), grid, 128, 0, stream, ncl, ncj, c2, cl_x.p, numneigh.p,
                    numneigh_masked.p, neighbors.p, maxneighs, cl_f.p, fi);
            std::swap(cl_x, cl_xn);
            ghosts_current = own_ghosts;
            force_launches++;
            if (timing) {
                float ms = 0;
                MDB_CUDA(cudaEventRecord(ev1, stream));
                MDB_CUDA(cudaEventSynchronize(ev1));
                MDB_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
                force_ms += ms;
            }
            return;
        }
        // force_variant 0 (default) picks the fastest measured kernel per case (profiles/r1_ab3.txt): full lists -> lane per
        // i atom (packed FP32 in SP), half lists -> warp per i-cluster / lane per j atom (its reaction forces need no
        // shuffles).  1 = lane per i atom, scalar; 2 = lane per i atom, packed FP32 (SP full only); 3 = warp per i-cluster.
        int fv = force_variant;
        if (fv == 0) fv = P.half_neigh ? 3 : (sizeof(real) == 4 ? 2 : 1);
        if (fv == 2 && (sizeof(real) != 4 || P.half_neigh)) fv = 1;
        if (fv == 3) {
            typedef typename PosOf<real>::type P4;
            const size_t nslots = (size_t)(ncj + nghost + 1) * N;
            pos4.ensure(nslots * sizeof(P4), false, stream);
            MDB_LAUNCH(launches, (k_cp_pack_j<real, N>), grid_for(nslots, 256), 256, 0, stream, nslots, cl_x.p, (P4*)pos4.p);
            if (P.half_neigh) {
                MDB_CUDA(cudaMemsetAsync(cl_f.p, 0, (size_t)ncj * 3 * N * sizeof(real), stream));
                MDB_LAUNCH(launches, (k_cp_force_jl<real, N, true>), grid_for(ncl, 4), 128, 0, stream, ncl, ncj, dummy_cj, c2,
                    cl_x.p, (const P4*)pos4.p, numneigh.p, neighbors.p, maxneighs, cl_f.p);
            } else {
                MDB_LAUNCH(launches, (k_cp_force_jl<real, N, false>), grid_for(ncl, 4), 128, 0, stream, ncl, ncj, dummy_cj, c2,
                    cl_x.p, (const P4*)pos4.p, numneigh.p, neighbors.p, maxneighs, cl_f.p);
            }
        } else if (P.half_neigh) {
            MDB_CUDA(cudaMemsetAsync(cl_f.p, 0, (size_t)ncj * 3 * N * sizeof(real), stream));
            MDB_LAUNCH(launches, (k_cp_force_lj<real, N, true>), grid, 128, 0, stream, ncl, ncj, c2, cl_x.p, numneigh.p,
                numneigh_masked.p, neighbors.p, maxneighs, cl_f.p, fi);
        } else if (fv == 2) {
            launch_packed(grid, c2);
        } else {
            MDB_LAUNCH(launches, (k_cp_force_lj<real, N, false>), grid, 128, 0, stream, ncl, ncj, c2, cl_x.p, numneigh.p,
                numneigh_masked.p, neighbors.p, maxneighs, cl_f.p, fi);
        }
        force_launches++;
        if (timing) {
            float ms = 0;
            MDB_CUDA(cudaEventRecord(ev1, stream));
            MDB_CUDA(cudaEventSynchronize(ev1));
            MDB_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
            force_ms += ms;
        }
    }
    // sp_kernel: 0 = lane per i atom (k_cp_force_lj_sp_packed), 1 = two lanes per i-cluster (k_cp_force_lj_sp_duo),
    // 2 (default) = the same with the MUFU reciprocal as it is (max. error 1 ulp = 2^-23) instead of MUFU + Newton step: the
    // reference's own SP kernel takes _mm512_rcp14_ps, relative error 2^-14, unrefined (common/simd/avx512_float.h:55-58;
    // clusterpair/force_lj.c:325-326,536), so this is still 9 bits closer to the exact quotient than what it is compared with
    void launch_packed(unsigned grid, const LJConst2<float>& c2)
    {
        const CpFused<float> nofi { nullptr, nullptr, 0.f, 0.f, nullptr, nullptr, 0, 0.f, 0.f, 0.f };
        const unsigned g2 = grid_for((size_t)ncl * 2, 128);
        if (sp_kernel == 1)
            MDB_LAUNCH(launches, (k_cp_force_lj_sp_duo<N, false, true>), g2, 128, 0, stream, ncl, dummy_cj, c2, (const float*)cl_x.p,
                numneigh.p, neighbors.p, maxneighs, (float*)cl_f.p, nofi);
        else if (sp_kernel == 2)
            MDB_LAUNCH(launches, (k_cp_force_lj_sp_duo<N, false, false>), g2, 128, 0, stream, ncl, dummy_cj, c2, (const float*)cl_x.p,
                numneigh.p, neighbors.p, maxneighs, (float*)cl_f.p, nofi);
        else
            MDB_LAUNCH(launches, k_cp_force_lj_sp_packed<N>, grid, 128, 0, stream, ncl, dummy_cj, c2, (const float*)cl_x.p, numneigh.p,
                neighbors.p, maxneighs, (float*)cl_f.p, nofi);
    }
    void launch_packed(unsigned, const LJConst2<double>&) {}
    void launch_packed_fused(unsigned grid, const LJConst2<float>& c2, const CpFused<float>& fi)
    {
        const unsigned g2 = grid_for((size_t)ncl * 2, 128);
        if (sp_kernel == 1)
            MDB_LAUNCH(launches, (k_cp_force_lj_sp_duo<N, true, true>), g2, 128, 0, stream, ncl, dummy_cj, c2, (const float*)cl_x.p,
                numneigh.p, neighbors.p, maxneighs, (float*)cl_f.p, fi);
        else if (sp_kernel == 2)
            MDB_LAUNCH(launches, (k_cp_force_lj_sp_duo<N, true, false>), g2, 128, 0, stream, ncl, dummy_cj, c2, (const float*)cl_x.p,
                numneigh.p, neighbors.p, maxneighs, (float*)cl_f.p, fi);
        else
            MDB_LAUNCH(launches, (k_cp_force_lj_sp_packed<N, true>), grid, 128, 0, stream, ncl, dummy_cj, c2, (const float*)cl_x.p,
                numneigh.p, neighbors.p, maxneighs, (float*)cl_f.p, fi);
    }
    void launch_packed_fused(unsigned, const LJConst2<double>&, const CpFused<double>&) {}
    double computeForce() override // returns elapsed seconds like the reference's ComputeForceFunction
    {
        const bool t        = timing;
        timing              = true;
        const double before = force_ms;
        launch_force();
        timing          = t;
        const double el = (force_ms - before) * 1e-3;
        if (!t) force_ms = before;
        return el;
    }
    void integrate(int mode)
    {
        if (mode != 1) ghosts_current = false; // positions move
        const unsigned grid = grid_for((size_t)ncl * CP_M, 256);
        if (mode == 0)
            MDB_LAUNCH(launches, (k_cp_integrate<real, N, 0>), grid, 256, 0, stream, ncl, dtforce, dt, inat.p, cl_x.p, cl_v.p, cl_f.p);
        else if (mode == 1)
            MDB_LAUNCH(launches, (k_cp_integrate<real, N, 1>), grid, 256, 0, stream, ncl, dtforce, dt, inat.p, cl_x.p, cl_v.p, cl_f.p);
        else
            MDB_LAUNCH(launches, (k_cp_integrate<real, N, 2>), grid, 256, 0, stream, ncl, dtforce, dt, inat.p, cl_x.p, cl_v.p, cl_f.p);
    }
    void initialIntegrate() override { integrate(0); } // integrate.c:23-44
    void finalIntegrate() override { integrate(1); }   // integrate.c:46-63

    // ------------------------------------------------------------------ driver flow
    void setup(bool adjust) override // clusterpair/main.c:40-76 after the atoms exist
    {
        derive();
        setupNeighbor();
        setupThermo();
        if (adjust) adjustThermo();
        buildClusters();
        defineJClusters();
        setupPbc();
        binClusters();
        buildNeighbor();
    }
    // MDB_CP_PHASES=1 in the environment: wall-clock per operator of reneighbour() (with a sync after each),
    // printed when the ctx is destroyed.  Diagnostics only.
    double phase_ms[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    int phase_calls    = 0;
    const bool phases  = getenv("MDB_CP_PHASES") != nullptr;
    template <class F> void phase(int k, F f)
    {
        if (!phases) { f(); return; }
        MDB_CUDA(cudaStreamSynchronize(stream));
        const auto t0 = std::chrono::steady_clock::now();
        f();
        MDB_CUDA(cudaStreamSynchronize(stream));
        phase_ms[k] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    void reneighbour() override // clusterpair/main.c:78-93
    {
        NvtxRange nvtx_range_("reneighbour");
        phase(0, [&] { updateSingleAtoms(); });
        phase(1, [&] { updateAtomsPbc(); });
        phase(2, [&] { buildClusters(); });
        phase(3, [&] { defineJClusters(); });
        phase(4, [&] { setupPbc(); });
        phase(5, [&] { binClusters(); });
        phase(6, [&] { buildNeighbor(); });
        phase_calls++;
    }
    void run(int nsteps, double* thermo_out, int max_records, int* nrecords, double* timers) override
    {
        if (!thermo_ready) setupThermo();
        const int nstat  = P.nstat > 0 ? P.nstat : nsteps + 1;
        const int every  = P.reneigh_every > 0 ? P.reneigh_every : nsteps + 1;
        const int pevery = prune_every > 0 ? prune_every : nsteps + 1;
        d_thermo.ensure((size_t)4 * (nsteps / nstat + 3), false, stream);
        std::vector<int> rec_step;
        auto record = [&](int step) {
            vel_sums(d_thermo.p + 4 * rec_step.size());
            rec_step.push_back(step);
        };
        const double f0 = force_ms, n0 = neigh_ms;
        record(0);
        launch_force();
        MDB_CUDA(cudaEventRecord(evR0, stream)); // timer[TOTAL] starts after the first force call, main.c:236-239
        bool initial_done = false;
        const bool fuse = can_fuse_force();
        bool second_ok  = false; // cl_xn mirrors cl_x's padding slots, ghost tiles and dummy tile (they change per rebuild)
        for (int n = 0; n < nsteps; n++) {
            if (!initial_done) initialIntegrate();
            if ((n + 1) % every) {
                if (!((n + 1) % pevery)) pruneNeighbor();
                updatePbc(false);
            } else {
                reneighbour();
                second_ok = false;
            }
            const bool rec  = !((n + 1) % nstat) && (n + 1) < nsteps;
            const bool last = n + 1 == nsteps;
            if (fuse && !rec && !last) { // force(n) + final(n) + initial(n+1) in one launch
                if (!second_ok) sync_second_array();
                second_ok = true;
                launch_force(true);
                initial_done = true;
                continue;
            }
            launch_force();
            // final(n) + initial(n+1) fuse into one pass unless something reads the state in between
            if (rec || last) {
                finalIntegrate();
                initial_done = false;
                if (rec) record(n + 1);
            } else {
                integrate(2);
                initial_done = true;
            }
        }
        MDB_CUDA(cudaEventRecord(evR1, stream));
        updateSingleAtoms(); // main.c:300
        record(nsteps);
        std::vector<double> h(4 * rec_step.size());
        MDB_CUDA(cudaMemcpyAsync(h.data(), d_thermo.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, stream));
        MDB_CUDA(cudaStreamSynchronize(stream));
        float ms = 0;
        MDB_CUDA(cudaEventElapsedTime(&ms, evR0, evR1));
        int nr = 0;
        for (size_t r = 0; r < rec_step.size(); r++)
            if (thermo_out && nr < max_records) {
                thermo_out[3 * nr] = rec_step[r];
                thermo_from_sum(h[4 * r + 3], &thermo_out[3 * nr + 1], &thermo_out[3 * nr + 2]);
                nr++;
            }
        if (nrecords) *nrecords = nr;
        if (timers) {
            timers[0] = ms * 1e-3;
            timers[1] = (force_ms - f0) * 1e-3;
            timers[2] = (neigh_ms - n0) * 1e-3;
        }
    }

    // ------------------------------------------------------------------ counters / parity accessors
    void countPairs(long long* cluster_pairs, long long* inside) override
    {
        MDB_CUDA(cudaMemsetAsync(d_cnt.p, 0, 2 * sizeof(unsigned long long), stream));
        MDB_LAUNCH(launches, (k_cp_count_pairs<real, N>), grid_for((size_t)ncl * CP_M, 128), 128, 0, stream, ncl,
            cutforce * cutforce, cl_x.p, numneigh.p, neighbors.p, maxneighs, d_cnt.p);
        unsigned long long h[2];
        MDB_CUDA(cudaMemcpyAsync(h, d_cnt.p, sizeof h, cudaMemcpyDeviceToHost, stream));
        MDB_CUDA(cudaStreamSynchronize(stream));
        *cluster_pairs = (long long)h[0];
        *inside        = (long long)h[1];
    }
    void d2h(void* dst, const void* src, size_t bytes)
    {
        if (bytes) MDB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream));
    }
    void getClusters(int which, int* natoms, void* bbox) override
    {
        if (which == 'i') {
            d2h(natoms, inat.p, ncl * sizeof(int));
            d2h(bbox, ibb.p, (size_t)ncl * 6 * sizeof(real));
        } else if (which == 'j') {
            d2h(natoms, jnat.p, (size_t)(ncj + nghost) * sizeof(int));
            d2h(bbox, jbb.p, (size_t)(ncj + nghost) * 6 * sizeof(real));
        } else throw Error("mdb_cp_getClusters: which must be 'i' or 'j'");
        MDB_CUDA(cudaStreamSynchronize(stream));
    }
    void getClusterData(int which, void* out) override
    {
        const size_t nx = (size_t)(ncj + nghost) * 3 * N;
        if (which == 'x') d2h(out, cl_x.p, nx * sizeof(real));
        else if (which == 'v') d2h(out, cl_v.p, (size_t)ncj * 3 * N * sizeof(real));
        else if (which == 'f') d2h(out, cl_f.p, (size_t)ncj * 3 * N * sizeof(real));
        else throw Error("mdb_cp_getClusterData: which must be 'x', 'v' or 'f'");
        MDB_CUDA(cudaStreamSynchronize(stream));
        if (which == 'x') { // padding lanes: device sentinel -> the reference's +INFINITY (cp_kernels.cuh, CP_PAD)
            real* o = (real*)out;
            for (size_t k = 0; k < nx; k++)
                if (o[k] >= CP_PAD_MIN) o[k] = INFINITY;
        }
    }
    void getClusterTags(int* tags) override
    {
        d2h(tags, cl_tag.p, (size_t)(ncj + nghost) * N * sizeof(int));
        MDB_CUDA(cudaStreamSynchronize(stream));
    }
    void getClusterBins(int* b) override
    {
        d2h(b, ibin.p, ncl * sizeof(int));
        MDB_CUDA(cudaStreamSynchronize(stream));
    }
    void getLists(int* nn, int* nm, int* nb, int stride) override
    {
        if (!lists_ready) throw Error("mdb_cp_getLists: no cluster-pair list");
        d2h(nn, numneigh.p, ncl * sizeof(int));
        d2h(nm, numneigh_masked.p, ncl * sizeof(int));
        if (nb)
            MDB_CUDA(cudaMemcpy2DAsync(nb, (size_t)stride * sizeof(int), neighbors.p, (size_t)maxneighs * sizeof(int),
                (size_t)std::min(stride, maxneighs) * sizeof(int), ncl, cudaMemcpyDeviceToHost, stream));
        MDB_CUDA(cudaStreamSynchronize(stream));
    }
    void getGhostMap(int* bm, int* px, int* py, int* pz) override
    {
        std::vector<int> c(nghost);
        d2h(bm, border_map.p, nghost * sizeof(int));
        d2h(c.data(), code.p, nghost * sizeof(int));
        MDB_CUDA(cudaStreamSynchronize(stream));
        for (int k = 0; k < nghost; k++) {
            px[k] = (c[k] & 3) - 1;
            py[k] = ((c[k] >> 2) & 3) - 1;
            pz[k] = ((c[k] >> 4) & 3) - 1;
        }
    }
    // synthetic clusters and lists of the reference's kernel micro-benchmark (clusterpair/main-stub.c)
    void stub(int niclusters, int nat, int pattern, int nneighs, int nreps, int masked, unsigned seed) override
    {
        if (niclusters < 1 || nat < 1 || nat > CP_M || pattern < 0 || pattern > 2 || nneighs < 1 || nreps < 1)
            throw Error("mdb_cp_stub: bad arguments");
        if (JFAC > 1 && (niclusters & 1)) niclusters++; // two i-clusters per tile
        ncl = niclusters; ncj = ncl / JFAC; nghost = 0; dummy_cj = ncj;
        if (pattern == 2 && ncj <= nneighs)
            throw Error("P_RAND: Number of j-clusters should be higher than number of j-cluster neighbors per i-cluster!");
        Natoms = (long long)ncl * nat;
        Nlocal = (int)Natoms;
        ensure_atoms(Nlocal);
        ensure_tiles((size_t)ncj + 2, false);
        inat.ensure(ncl, false, stream);
        ibin.ensure(ncl, false, stream);
        ibb.ensure((size_t)ncl * 6, false, stream);
        MDB_LAUNCH(launches, (k_cp_stub_clusters<real, N>), grid_for((size_t)(ncl + 1) * CP_M, 256), 256, 0, stream, ncl, nat, cl_x.p,
            cl_v.p, cl_f.p, cl_tag.p, inat.p, jnat.p, ibb.p, ibin.p);
        updateSingleAtoms();
        maxneighs = nneighs * nreps;
        numneigh.ensure(ncl, false, stream);
        numneigh_masked.ensure(ncl, false, stream);
        neighbors.ensure((size_t)ncl * maxneighs, false, stream);
        // this library's kernels need the diagonal entries inside the masked prefix (the list build puts them there); the
        // synthetic patterns repeat them (nreps) or place them anywhere ("fix"), so only "rand" / a single "seq" block may be
        // left unmasked when the caller asks for the unmasked loop
        const int nm = masked ? maxneighs : (pattern == 2 ? 0 : (pattern == 0 && nreps == 1 ? 1 : maxneighs));
        MDB_LAUNCH(launches, k_cp_stub_neighbors<N>, grid_for(ncl, 128), 128, 0, stream, ncl, ncj, pattern, nneighs, nreps, nm, seed,
            maxneighs, numneigh.p, numneigh_masked.p, neighbors.p);
        lists_ready = true;
        neigh_ready = false;
    }
    void getNeighborParams(int* I, double* R, int* st) override
    {
        if (!neigh_ready) setupNeighbor();
        const int iv[8]     = { g.nbinx, g.nbiny, g.mbinx, g.mbiny, g.mbins, g.mbinxlo, g.mbinylo, nstencil };
        const double rv[10] = { (double)binsizex, (double)binsizey, (double)g.bininvx, (double)g.bininvy, (double)g.cutneighsq,
            (double)cutneigh, (double)xprd, (double)yprd, (double)zprd, (double)g.rbb_sq };
        memcpy(I, iv, sizeof iv);
        memcpy(R, rv, sizeof rv);
        if (st) memcpy(st, h_stencil.data(), nstencil * sizeof(int));
    }
};

static CpBase* make_cp(const mdb_params& p, int cluster_n, int device)
{
    int ndev      = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) throw Error("mdb_cp_create: no CUDA device (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) throw Error(fmt("mdb_cp_create: device %d out of range (%d devices)", device, ndev));
    if (p.force_field != MDB_FF_LJ) throw Error("mdb_cp_create: the clusterpair scheme has only the LJ kernels (force.c)");
    if (p.ntypes != 1) throw Error("mdb_cp_create: only ntypes == 1 is supported (EXPLICIT_TYPES off)");
    if (cluster_n != 4 && cluster_n != 8) throw Error("mdb_cp_create: cluster_n must be 4 or 8 (M = 4)");
    if (p.precision == MDB_DP) return cluster_n == 4 ? (CpBase*)new CpSim<double, 4>(p, device) : new CpSim<double, 8>(p, device);
    if (p.precision == MDB_SP) return cluster_n == 4 ? (CpBase*)new CpSim<float, 4>(p, device) : new CpSim<float, 8>(p, device);
    throw Error("mdb_cp_create: precision must be MDB_SP or MDB_DP");
}

#include "cp_dd.cuh"

void set_last_error(const char* msg); // mdb200.cu

} // namespace mdb

using namespace mdb;

struct mdb_cp {
    CpBase* s;
};

#define MDB_CP_TRY(body)                                                                         \
    try {                                                                                        \
        mdb::NvtxRange nvtx_range_(__func__);                                                    \
        if (!c || !c->s) throw Error("null mdb_cp");                                             \
        body;                                                                                    \
        return 0;                                                                                \
    } catch (const std::exception& e) {                                                          \
        set_last_error(e.what());                                                                \
        return -1;                                                                               \
    }

extern "C" {

mdb_cp* mdb_cp_create(const mdb_params* p, int cluster_n, int device)
{
    try {
        if (!p) throw Error("mdb_cp_create: null params");
        mdb_cp* c = new mdb_cp;
        c->s      = make_cp(*p, cluster_n, device);
        return c;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return nullptr;
    }
}
void mdb_cp_destroy(mdb_cp* c)
{
    if (!c) return;
    delete c->s;
    delete c;
}
int mdb_cp_setStream(mdb_cp* c, void* s) { MDB_CP_TRY(c->s->setStream((cudaStream_t)s)) }
int mdb_cp_sync(mdb_cp* c) { MDB_CP_TRY(c->s->sync()) }
int mdb_cp_setOption(mdb_cp* c, const char* name, double v) { MDB_CP_TRY(c->s->setOption(name, v)) }
long long mdb_cp_createAtom(mdb_cp* c)
{
    try {
        if (!c || !c->s) throw Error("null mdb_cp");
        return c->s->createAtom();
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return -1;
    }
}
int mdb_cp_setAtoms(mdb_cp* c, long long n, const void* x, const void* y, const void* z, const void* vx, const void* vy,
    const void* vz)
{
    MDB_CP_TRY(c->s->setAtoms(n, x, y, z, vx, vy, vz))
}
int mdb_cp_getAtoms(mdb_cp* c, int which, void* x, void* y, void* z, int* tag) { MDB_CP_TRY(c->s->getAtoms(which, x, y, z, tag)) }
int mdb_cp_getCounts(mdb_cp* c, long long v[8]) { MDB_CP_TRY(c->s->getCounts(v)) }
int mdb_cp_setupThermo(mdb_cp* c) { MDB_CP_TRY(c->s->setupThermo()) }
int mdb_cp_adjustThermo(mdb_cp* c) { MDB_CP_TRY(c->s->adjustThermo()) }
int mdb_cp_computeThermo(mdb_cp* c, double* T, double* P) { MDB_CP_TRY(c->s->computeThermo(T, P)) }
int mdb_cp_setupNeighbor(mdb_cp* c) { MDB_CP_TRY(c->s->setupNeighbor()) }
int mdb_cp_buildClusters(mdb_cp* c) { MDB_CP_TRY(c->s->buildClusters()) }
int mdb_cp_defineJClusters(mdb_cp* c) { MDB_CP_TRY(c->s->defineJClusters()) }
int mdb_cp_setupPbc(mdb_cp* c) { MDB_CP_TRY(c->s->setupPbc()) }
int mdb_cp_binClusters(mdb_cp* c) { MDB_CP_TRY(c->s->binClusters()) }
int mdb_cp_buildNeighbor(mdb_cp* c) { MDB_CP_TRY(c->s->buildNeighbor()) }
int mdb_cp_pruneNeighbor(mdb_cp* c) { MDB_CP_TRY(c->s->pruneNeighbor()) }
int mdb_cp_updateSingleAtoms(mdb_cp* c) { MDB_CP_TRY(c->s->updateSingleAtoms()) }
int mdb_cp_updateAtomsPbc(mdb_cp* c) { MDB_CP_TRY(c->s->updateAtomsPbc()) }
int mdb_cp_updatePbc(mdb_cp* c, int first) { MDB_CP_TRY(c->s->updatePbc(first != 0)) }
double mdb_cp_computeForce(mdb_cp* c)
{
    try {
        if (!c || !c->s) throw Error("null mdb_cp");
        return c->s->computeForce();
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return -1.0;
    }
}
int mdb_cp_initialIntegrate(mdb_cp* c) { MDB_CP_TRY(c->s->initialIntegrate()) }
int mdb_cp_finalIntegrate(mdb_cp* c) { MDB_CP_TRY(c->s->finalIntegrate()) }
int mdb_cp_setup(mdb_cp* c, int adjust) { MDB_CP_TRY(c->s->setup(adjust != 0)) }
int mdb_cp_reneighbour(mdb_cp* c) { MDB_CP_TRY(c->s->reneighbour()) }
int mdb_cp_run(mdb_cp* c, int nsteps, double* thermo_out, int max_records, int* nrecords, double* timers)
{
    MDB_CP_TRY(c->s->run(nsteps, thermo_out, max_records, nrecords, timers))
}
int mdb_cp_saveState(mdb_cp* c) { MDB_CP_TRY(c->s->saveState()) }
int mdb_cp_restoreState(mdb_cp* c) { MDB_CP_TRY(c->s->restoreState()) }
int mdb_cp_setTiming(mdb_cp* c, int on) { MDB_CP_TRY(c->s->timing = on != 0) }
int mdb_cp_getKernelStats(mdb_cp* c, double* force_ms, long long* force_launches, double* neigh_ms, long long* neigh_launches,
    long long* total_launches)
{
    MDB_CP_TRY({
        if (force_ms) *force_ms = c->s->force_ms;
        if (force_launches) *force_launches = c->s->force_launches;
        if (neigh_ms) *neigh_ms = c->s->neigh_ms;
        if (neigh_launches) *neigh_launches = c->s->neigh_launches;
        if (total_launches) *total_launches = c->s->launches;
    })
}
int mdb_cp_resetKernelStats(mdb_cp* c)
{
    MDB_CP_TRY({
        c->s->force_ms = c->s->neigh_ms = 0;
        c->s->force_launches = c->s->neigh_launches = c->s->launches = 0;
    })
}
int mdb_cp_countPairs(mdb_cp* c, long long* cluster_pairs, long long* inside) { MDB_CP_TRY(c->s->countPairs(cluster_pairs, inside)) }
int mdb_cp_getClusters(mdb_cp* c, int which, int* natoms, void* bbox) { MDB_CP_TRY(c->s->getClusters(which, natoms, bbox)) }
int mdb_cp_getClusterData(mdb_cp* c, int which, void* out) { MDB_CP_TRY(c->s->getClusterData(which, out)) }
int mdb_cp_getClusterTags(mdb_cp* c, int* tags) { MDB_CP_TRY(c->s->getClusterTags(tags)) }
int mdb_cp_getClusterBins(mdb_cp* c, int* b) { MDB_CP_TRY(c->s->getClusterBins(b)) }
int mdb_cp_getLists(mdb_cp* c, int* nn, int* nm, int* nb, int stride) { MDB_CP_TRY(c->s->getLists(nn, nm, nb, stride)) }
int mdb_cp_getGhostMap(mdb_cp* c, int* bm, int* px, int* py, int* pz) { MDB_CP_TRY(c->s->getGhostMap(bm, px, py, pz)) }
int mdb_cp_stub(mdb_cp* c, int niclusters, int iclusters_natoms, int pattern, int nneighs, int nreps, int masked, unsigned seed)
{
    MDB_CP_TRY(c->s->stub(niclusters, iclusters_natoms, pattern, nneighs, nreps, masked, seed))
}
int mdb_cp_getNeighborParams(mdb_cp* c, int ints[8], double reals[10], int* stencil)
{
    MDB_CP_TRY(c->s->getNeighborParams(ints, reals, stencil))
}

} // extern "C"
