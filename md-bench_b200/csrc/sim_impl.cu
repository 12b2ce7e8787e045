// sim_impl.cu -- Sim<real>: one verletlist simulation domain resident on one B200.
// Host code here only orchestrates: every per-atom operation is a kernel in vl_kernels.cuh /
// eam_kernels.cuh; the only device->host traffic inside the time loop is one small flag read per
// neighbor rebuild (ghost total, bin / list overflow).
#include <algorithm>
#include <initializer_list>

#include "sim.cuh"
#include "scan.cuh"
#include "vl_kernels.cuh"
#include "eam_kernels.cuh"
#include "dd_topo.h"
#include "dd_kernels.cuh"
#include "nccl_dl.h"

namespace mdb {

template <class real> struct Sim final : SimBase {
    mdb_params P;
    int device;
    cudaStream_t stream = nullptr, own_stream = nullptr;
    // ---- Parameter fields in working precision (reference common/parameter.h:27-61) ----
    real epsilon, sigma, sigma6, temp, rho, mass, dt, dtforce, skin, cutforce, cutneigh, lattice;
    real xprd, yprd, zprd, xlo, xhi, ylo, yhi, zlo, zhi;
    // ---- thermo scales (common/thermo.c:14-26) ----
    real mvv2e = 1, dof_boltz = 1, t_scale = 1, p_scale = 1;
    bool thermo_ready = false;
    // ---- atoms ----
    long long Natoms = 0;
    int Nlocal = 0, Nghost = 0;
    // Internal order: local atoms are kept sorted by neighbor bin (sort_atoms(), re-done at every
    // rebuild) so that the gathers of the force and list kernels hit few cache lines per warp;
    // orig[p] is the reference's index of the atom in slot p, and every accessor of the C ABI
    // translates back, so callers only ever see the reference's numbering.
    DBuf<real> x, y, z, vx, vy, vz, fx, fy, fz, sx, sy, sz, svx, svy, svz, stage;
    DBuf<real> x2, y2, z2, vx2, vy2, vz2, fx2, fy2, fz2, tx, ty, tz;
    DBuf<int> type, border_map, ghost_code, ghost_cnt, ghost_off, orig, orig2, type2, extmap, nn_ext;
    // Spatial sort of the local atoms (the reference's SORT_ATOMS build option, off by default there
    // too, config.mk:23).  Off by default: for the benchmark lattices the generator's order is already
    // spatially compact and its lattice regularity makes the gathers of neighbouring lanes fall on
    // consecutive atoms (A/B in profiles/r1_ab.txt: force 1.49 ms unsorted vs 1.67 ms sorted at 8.4M
    // atoms).  Turn it on (mdb_setOption "sort_atoms") for long runs of diffusing systems.
    bool sort_enabled = false, extmap_valid = false;
    bool fuse_integrate = true;
    // ghosts_current: the ghost positions belong to the local positions as they are now (set by the fused kernel's epilogue,
    // cleared by everything else that moves atoms or renumbers them)
    int ghost_epilogue = -1; // -1: where it pays (see forceFinalInitialIntegrate), 0 / 1: off / on
    bool ghosts_current = false, ghost_tables_valid = false;
    bool fuse_force = true; // mdb_run: integrate halves in the force kernel's epilogue (k_force_lj_full_fi)
    // packed (x, y) copy of the positions for the fused kernel's vector gathers ("xy_gather"); valid only between fused
    // steps of one mdb_run: the fused epilogue writes the locals, updatePbc the ghosts, everything else invalidates it
    typedef typename Vec2Of<real>::type vec2;
    int xy_gather = 1;
    bool xy_valid = ghosts_current = false;
    DBuf<vec2> xy, xy2;
    DBuf<real> zg, zg2; // gather copy of z (in-place variant of the fused step, decomposed runs)
    int sort_block = 0; // sort_atoms: 0 = the reference's x-fastest bin order, B > 0 = blocks of B^3 bins (x-fastest inside)
    long long bin_rank_key = -1; // geometry the rank table was built for
    DBuf<int> bin_rank;
    NbLayout LL { 0, 0, 0 };                                    // element (i,k) at neighbors[LL.base(i) + k*LL.sk]
    DBuf<float> cxs, cys, czs; // candidates in CSR order, SoA (k_build_neighbor_v6)
    DBuf<int> cids;
    DBuf<int> s_type1; // saveState: types in the reference's atom order
    DBuf<RunRow> runs; // runs of x-adjacent stencil bins as k_build_neighbor_v6 reads them
    RunGeom rg;
    int nruns = 0;
    std::vector<int> h_ghost_order, h_orig, h_bm, h_code;
    DBuf<unsigned> ghost_msk;
    int saved_n = 0;
    // ---- neighbor (verletlist/neighbor.c:24-38) ----
    BinGeom<real> bg {};
    real binsizex, binsizey, binsizez, cutneighsq;
    int nstencil = 0, maxneighs = 100, max_bin_count = 0;
    size_t nstride = 0;
    bool neigh_ready = false;
    std::vector<int> h_stencil;
    DBuf<int> stencil, atom_bin, bincount, binstart, cursor, binatoms, numneigh, neighbors, rows;
    Scanner scanner;
    // ---- eam ----
    EamTables<real> eam;
    DBuf<real> fp, rhor_spline, frho_spline, z2r_spline, eam_rho4, eam_frc12; // + repacked rows for the v2 kernels
    DBuf<real> eam_vs4; // (value, slope) of rhor and z2r per knot, generation-3 force pass (eam_variant 2)
    DBuf<vec2> eam_zf;  // packed (z, fp) of locals and ghosts, generation 3
    int eam_variant = 2; // option "eam_variant": 0 first kernels (IEEE sqrt / division), 1 generation 2, 2 generation 3 (default; a brick runs 2)
    // ---- scratch ----
    int* h_flags      = nullptr; // pinned: [0] ghost total, [1] max neighbors, [2] max bin count
    DBuf<int> d_flags;
    double* h_red = nullptr; // pinned
    DBuf<double> d_partial, d_red, d_thermo;
    DBuf<unsigned long long> d_cnt;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evA = nullptr, evB = nullptr;

    Sim(const mdb_params& p, int dev) : P(p), device(dev)
    {
        MDB_CUDA(cudaSetDevice(device));
        MDB_CUDA(cudaStreamCreateWithFlags(&own_stream, cudaStreamNonBlocking));
        stream = own_stream;
        MDB_CUDA(cudaMallocHost(&h_flags, 16 * sizeof(int)));
        MDB_CUDA(cudaMallocHost(&h_red, 16 * sizeof(double)));
        MDB_CUDA(cudaEventCreate(&ev0));
        MDB_CUDA(cudaEventCreate(&ev1));
        MDB_CUDA(cudaEventCreate(&evA));
        MDB_CUDA(cudaEventCreate(&evB));
        scanner.launches = &launches;
        d_flags.ensure(16, false, stream);
        d_red.ensure(16, false, stream);
        d_partial.ensure(RED_BLOCKS * 4, false, stream);
        d_cnt.ensure(4, false, stream);
        derive();
        preload_kernels();
    }
    // CUDA loads a kernel lazily at its first launch (milliseconds each).  A short run in a fresh process -- the C driver's
    // default run is BASELINE config 1: 131 072 atoms x 200 steps = ~10 ms of kernel time -- would meet every kernel of the
    // time loop for the first time INSIDE the timed loop (measured: 29 ms instead of 11 ms); load them here instead.
    void preload_kernels()
    {
        cudaFuncAttributes a;
        const void* ks[] = { (const void*)k_force_lj_full_fi<real, 4, sizeof(real) == 4, true>,
            (const void*)k_force_lj_full_fi<real, 4, sizeof(real) == 4>, (const void*)k_force_lj_full_fi<real, 4, sizeof(real) == 4, true, true>,
            (const void*)k_force_lj_full_v2<real, 4>, (const void*)k_force_lj_full_v6<real, 4>, (const void*)k_force_lj_half_v2<real, 4>,
            (const void*)k_build_neighbor_v6<real, false>, (const void*)k_build_neighbor_v6<real, true>, (const void*)k_pack_binned_soa<real>, (const void*)k_bin_count<real>,
            (const void*)k_bin_fill, (const void*)k_bin_sort, (const void*)k_ghost_count<real>, (const void*)k_ghost_fill,
            (const void*)k_update_pbc<real>, (const void*)k_update_atoms_pbc<real>, (const void*)k_pack_xy<real>,
            (const void*)k_initial_integrate<real>, (const void*)k_final_integrate<real>, (const void*)k_final_initial_integrate<real>,
            (const void*)k_vel_partial<real>, (const void*)k_vel_final, (const void*)k_permute_atoms<real>,
            (const void*)k_scatter_orig<real> };
        for (const void* k : ks) MDB_CUDA(cudaFuncGetAttributes(&a, k));
        scanner.preload();
    }
    ~Sim() override
    {
        cudaSetDevice(device);
        cudaStreamSynchronize(stream);
        for (DBuf<real>* b : { &x, &y, &z, &vx, &vy, &vz, &fx, &fy, &fz, &sx, &sy, &sz, &svx, &svy,
                 &svz, &stage, &fp, &rhor_spline, &frho_spline, &z2r_spline, &eam_rho4, &eam_frc12, &x2, &y2, &z2, &vx2, &vy2,
                 &vz2, &fx2, &fy2, &fz2, &tx, &ty, &tz })
            b->release();
        for (DBuf<real>* b : { &zg, &zg2, &eam_vs4 }) b->release();
        for (DBuf<vec2>* b : { &xy, &xy2, &eam_zf }) b->release();
        for (DBuf<int>* b : { &orig, &orig2, &type2, &extmap, &nn_ext, &bin_rank }) b->release();
        for (DBuf<int>* b : { &type, &border_map, &ghost_code, &ghost_cnt, &ghost_off, &stencil,
                 &atom_bin, &bincount, &binstart, &cursor, &binatoms, &numneigh, &neighbors, &rows, &d_flags })
            b->release();
        ghost_msk.release();
        cxs.release(); cys.release(); czs.release(); cids.release(); s_type1.release();
        runs.release();
        d_partial.release();
        d_red.release();
        d_thermo.release();
        d_cnt.release();
        scanner.release();
        cudaFreeHost(h_flags);
        cudaFreeHost(h_red);
        cudaEventDestroy(ev0);
        cudaEventDestroy(ev1);
        cudaEventDestroy(evA);
        cudaEventDestroy(evB);
        cudaStreamDestroy(own_stream);
    }

    // initParameter + command line (common/parameter.c:16-51, verletlist/main.c:233, 42-45).
    // Each value is narrowed to `real` where the reference assigns to an MD_FLOAT field.
    void derive()
    {
        epsilon  = (real)P.epsilon;
        sigma    = (real)P.sigma;
        real s2  = sigma * sigma; // parameter.c:118-119
        sigma6   = s2 * s2 * s2;
        temp     = (real)P.temp;
        rho      = (real)P.rho;
        mass     = (real)P.mass;
        dt       = (real)P.dt;
        derive_dtforce();                    // parameter.c:115 / eam_utils.c:35
        thermo_ready = false;                // setupThermo re-applies its EAM scaling (thermo.c:51)
        skin     = (real)P.skin;
        cutforce = (real)P.cutforce;
        cutneigh = cutforce + skin; // main.c:233
        if (P.force_field == MDB_FF_EAM) cutneigh = (real)((double)cutforce + 1.0); // eam_utils.c:30
        lattice  = (real)pow((4.0 / (double)rho), (1.0 / 3.0));
        if (P.from_input) {
            xlo = (real)P.xlo; xhi = (real)P.xhi; ylo = (real)P.ylo; yhi = (real)P.yhi;
            zlo = (real)P.zlo; zhi = (real)P.zhi;
            xprd = xhi - xlo; yprd = yhi - ylo; zprd = zhi - zlo;
        } else {
            xprd = P.nx * lattice; yprd = P.ny * lattice; zprd = P.nz * lattice;
            xlo = ylo = zlo = 0; xhi = xprd; yhi = yprd; zhi = zprd;
        }
    }

    void setStream(cudaStream_t s) override
    {
        MDB_CUDA(cudaStreamSynchronize(stream));
        stream = s ? s : own_stream;
    }
    void sync() override { MDB_CUDA(cudaStreamSynchronize(stream)); }

    void zero3(real* a, real* b, real* c, size_t n)
    {
        MDB_CUDA(cudaMemsetAsync(a, 0, n * sizeof(real), stream));
        MDB_CUDA(cudaMemsetAsync(b, 0, n * sizeof(real), stream));
        MDB_CUDA(cudaMemsetAsync(c, 0, n * sizeof(real), stream));
    }
    void ensure_atoms(size_t n, bool keep)
    {
        for (DBuf<real>* b : { &x, &y, &z, &vx, &vy, &vz, &fx, &fy, &fz }) b->ensure(n, keep, stream);
        type.ensure(n, keep, stream);
        orig.ensure(n, keep, stream);
    }
    void reset_order() // internal order := the caller's (reference) order
    {
        MDB_LAUNCH(launches, k_iota, grid_for(Nlocal, 256), 256, 0, stream, Nlocal, orig.p);
        extmap_valid = false;
    }

    // ------------------------------------------------------------------ atoms
    long long createAtom() override
    {
        derive();
        Natoms = 4LL * P.nx * P.ny * P.nz;
        if (Natoms > 2000000000LL) throw Error("createAtom: more than 2e9 atoms per domain");
        Nlocal = (int)Natoms;
        Nghost = 0;
        ensure_atoms((size_t)Nlocal + Nlocal / 4 + 1024, false);
        MDB_LAUNCH(launches, k_create_atoms<real>, grid_for(Natoms, 256), 256, 0, stream, P.nx, P.ny,
            P.nz, lattice, x.p, y.p, z.p, vx.p, vy.p, vz.p, type.p);
        if (P.ntypes > 1) { // atom.c:159: type = rand() % ntypes, one draw per atom in emission order (the host's rand() sequence)
            std::vector<int> t((size_t)Nlocal);
            for (int i = 0; i < Nlocal; i++) t[i] = rand() % P.ntypes;
            MDB_CUDA(cudaMemcpyAsync(type.p, t.data(), (size_t)Nlocal * sizeof(int), cudaMemcpyHostToDevice, stream));
            MDB_CUDA(cudaStreamSynchronize(stream));
        }
        zero3(fx.p, fy.p, fz.p, Nlocal);
        reset_order();
        neigh_ready = false;
        nstride = 0; // no list for these atoms yet
        xy_valid = ghosts_current = false;
        return Natoms;
    }
    // types of the local atoms in the reference's numbering.  EXPLICIT_TYPES (force_lj.c:61-67) looks the pair parameters up
    // by type pair, but every entry of those tables holds the same value (atom.c:84-89), so the kernels take the scalars.
    void getTypes(int* out) override
    {
        std::vector<int> t((size_t)Nlocal), o((size_t)Nlocal);
        MDB_CUDA(cudaMemcpyAsync(t.data(), type.p, (size_t)Nlocal * sizeof(int), cudaMemcpyDeviceToHost, stream));
        MDB_CUDA(cudaMemcpyAsync(o.data(), orig.p, (size_t)Nlocal * sizeof(int), cudaMemcpyDeviceToHost, stream));
        MDB_CUDA(cudaStreamSynchronize(stream));
        for (int p = 0; p < Nlocal; p++) out[o[p]] = t[p];
    }

    void load3(long long n, const void* a, const void* b, const void* c, real* dx, real* dy, real* dz,
        bool on_device)
    {
        const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
        if (P.layout == MDB_AOS) {
            if (on_device) {
                MDB_LAUNCH(launches, k_aos_to_soa<real>, grid_for(n, 256), 256, 0, stream, (size_t)n,
                    (const real*)a, dx, dy, dz);
            } else {
                stage.ensure(3 * n, false, stream);
                MDB_CUDA(cudaMemcpyAsync(stage.p, a, 3 * n * sizeof(real), kind, stream));
                MDB_LAUNCH(launches, k_aos_to_soa<real>, grid_for(n, 256), 256, 0, stream, (size_t)n,
                    (const real*)stage.p, dx, dy, dz);
            }
        } else {
            MDB_CUDA(cudaMemcpyAsync(dx, a, n * sizeof(real), kind, stream));
            MDB_CUDA(cudaMemcpyAsync(dy, b, n * sizeof(real), kind, stream));
            MDB_CUDA(cudaMemcpyAsync(dz, c, n * sizeof(real), kind, stream));
        }
    }

    void setAtoms(long long n, const void* ax, const void* ay, const void* az, const void* avx,
        const void* avy, const void* avz, const int* atype, bool on_device) override
    {
        if (n <= 0 || n > 2000000000LL) throw Error("setAtoms: bad atom count");
        derive();
        Natoms = n;
        Nlocal = (int)n;
        Nghost = 0;
        ensure_atoms((size_t)n + n / 4 + 1024, false);
        load3(n, ax, ay, az, x.p, y.p, z.p, on_device);
        if (avx) load3(n, avx, avy, avz, vx.p, vy.p, vz.p, on_device);
        else zero3(vx.p, vy.p, vz.p, n);
        if (atype)
            MDB_CUDA(cudaMemcpyAsync(type.p, atype, n * sizeof(int),
                on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, stream));
        else MDB_CUDA(cudaMemsetAsync(type.p, 0, n * sizeof(int), stream));
        zero3(fx.p, fy.p, fz.p, n);
        reset_order();
        MDB_CUDA(cudaStreamSynchronize(stream)); // host buffers may be reused by the caller
        neigh_ready = false;
        nstride = 0;
        xy_valid = ghosts_current = false;
        pending_force = pending_final = false;
    }

    void getAtoms(int which, bool ghosts, void* ax, void* ay, void* az) override
    {
        const real *a, *b, *c;
        if (which == 'x') { a = x.p; b = y.p; c = z.p; }
        else if (which == 'v') { a = vx.p; b = vy.p; c = vz.p; }
        else if (which == 'f') { a = fx.p; b = fy.p; c = fz.p; }
        else throw Error("getAtoms: which must be 'x', 'v' or 'f'");
        const bool wg  = ghosts && which == 'x' && Nghost > 0;
        const size_t n = (size_t)Nlocal + (wg ? Nghost : 0);
        // back to the reference's numbering: slot p -> orig[p] (ghosts: reference ghost order)
        const int* map = orig.p;
        if (wg) { build_extmap(); map = extmap.p; }
        tx.ensure(n, false, stream); ty.ensure(n, false, stream); tz.ensure(n, false, stream);
        MDB_LAUNCH(launches, k_scatter_orig<real>, grid_for(n, 256), 256, 0, stream, (int)n, map, a, b, c,
            tx.p, ty.p, tz.p);
        if (P.layout == MDB_AOS) {
            stage.ensure(3 * n, false, stream);
            MDB_LAUNCH(launches, k_soa_to_aos<real>, grid_for(n, 256), 256, 0, stream, n, tx.p, ty.p, tz.p, stage.p);
            MDB_CUDA(cudaMemcpyAsync(ax, stage.p, 3 * n * sizeof(real), cudaMemcpyDeviceToHost, stream));
        } else {
            MDB_CUDA(cudaMemcpyAsync(ax, tx.p, n * sizeof(real), cudaMemcpyDeviceToHost, stream));
            MDB_CUDA(cudaMemcpyAsync(ay, ty.p, n * sizeof(real), cudaMemcpyDeviceToHost, stream));
            MDB_CUDA(cudaMemcpyAsync(az, tz.p, n * sizeof(real), cudaMemcpyDeviceToHost, stream));
        }
        MDB_CUDA(cudaStreamSynchronize(stream));
    }

    void getCounts(long long* c, int* mn) override
    {
        c[0] = Natoms; c[1] = Nlocal; c[2] = Nghost; c[3] = (long long)x.cap;
        *mn  = maxneighs;
    }

    void saveState() override // stored in the reference's atom order
    {
        const size_t n = Nlocal;
        for (DBuf<real>* b : { &sx, &sy, &sz, &svx, &svy, &svz }) b->ensure(n, false, stream);
        MDB_LAUNCH(launches, k_scatter_orig<real>, grid_for(n, 256), 256, 0, stream, (int)n, orig.p, x.p, y.p,
            z.p, sx.p, sy.p, sz.p);
        MDB_LAUNCH(launches, k_scatter_orig<real>, grid_for(n, 256), 256, 0, stream, (int)n, orig.p, vx.p, vy.p,
            vz.p, svx.p, svy.p, svz.p);
        s_type1.ensure(n, false, stream);
        MDB_LAUNCH(launches, k_scatter_orig_int, grid_for(n, 256), 256, 0, stream, (int)n, orig.p, type.p, s_type1.p);
        saved_n = Nlocal;
    }
    void restoreState() override
    {
        if (!saved_n) throw Error("restoreState: nothing saved");
        const size_t n = saved_n;
        DBuf<real>* src[] = { &sx, &sy, &sz, &svx, &svy, &svz };
        DBuf<real>* dst[] = { &x, &y, &z, &vx, &vy, &vz };
        for (int k = 0; k < 6; k++)
            MDB_CUDA(cudaMemcpyAsync(dst[k]->p, src[k]->p, n * sizeof(real), cudaMemcpyDeviceToDevice, stream));
        MDB_CUDA(cudaMemcpyAsync(type.p, s_type1.p, n * sizeof(int), cudaMemcpyDeviceToDevice, stream));
        zero3(fx.p, fy.p, fz.p, n);
        Nlocal = saved_n;
        Nghost = 0;
        reset_order();
        nstride = 0; // the lists belong to the state that was just replaced
        xy_valid = ghosts_current = false;
    }

    // ------------------------------------------------------------------ thermo
    void setupThermo() override // common/thermo.c:30-53
    {
        const long long natoms = brick ? gNatoms : Natoms;
        // brick mode: the thermo scales are those of the whole box (bgrid bricks of xprd x yprd x zprd)
        const double gx = (double)xprd * bgrid[0], gy = (double)yprd * bgrid[1], gz = (double)zprd * bgrid[2];
        if (P.force_field == MDB_FF_LJ) {
            mvv2e     = (real)1.0;
            dof_boltz = (real)(natoms * 3 - 3);
            t_scale   = mvv2e / dof_boltz;
            p_scale   = (real)(1.0 / 3 / gx / gy / gz);
        } else {
            mvv2e     = (real)1.036427e-04;
            dof_boltz = (real)((double)(natoms * 3 - 3) * 8.617343e-05);
            t_scale   = mvv2e / dof_boltz;
            p_scale   = (real)(1.602176e+06 / 3 / gx / gy / gz);
            if (!thermo_ready) dtforce = dtforce / mvv2e; // thermo.c:51 (once per setup)
        }
        thermo_ready = true;
    }

    // sums {vx, vy, vz, m v^2} over local atoms -> out (device, 4 doubles)
    void vel_sums(double* out)
    {
        const int nb = (int)std::min<size_t>(RED_BLOCKS, grid_for(Nlocal, RED_THREADS));
        MDB_LAUNCH(launches, k_vel_partial<real>, nb, RED_THREADS, 0, stream, Nlocal, vx.p, vy.p, vz.p,
            mass, d_partial.p);
        MDB_LAUNCH(launches, k_vel_final, 1, RED_THREADS, 0, stream, nb, d_partial.p, out);
    }
    void read_red()
    {
        MDB_CUDA(cudaMemcpyAsync(h_red, d_red.p, 4 * sizeof(double), cudaMemcpyDeviceToHost, stream));
        MDB_CUDA(cudaStreamSynchronize(stream));
    }
    void thermo_from_sum(double msum, double* T, double* Pr) // thermo.c:64-65
    {
        real t = (real)msum;
        t      = t * t_scale;
        real p = (t * dof_boltz) * p_scale;
        *T     = t;
        *Pr    = p;
    }
    void computeThermo(double* T, double* Pr) override // common/thermo.c:55-80
    {
        if (!thermo_ready) setupThermo();
        vel_sums(d_red.p);
        read_red();
        thermo_from_sum(h_red[3], T, Pr);
    }
    void adjustThermo() override // common/thermo.c:82-122
    {
        if (!thermo_ready) setupThermo();
        vel_sums(d_red.p);
        read_red();
        const real vxtot = (real)h_red[0] / (real)Natoms, vytot = (real)h_red[1] / (real)Natoms,
                   vztot = (real)h_red[2] / (real)Natoms;
        MDB_LAUNCH(launches, k_vel_shift<real>, grid_for(Nlocal, 256), 256, 0, stream, Nlocal, vx.p, vy.p,
            vz.p, vxtot, vytot, vztot);
        vel_sums(d_red.p);
        read_red();
        real t = (real)h_red[3];
        t *= t_scale;
        const real factor = (real)sqrt((double)(temp / t));
        MDB_LAUNCH(launches, k_vel_scale<real>, grid_for(Nlocal, 256), 256, 0, stream, Nlocal, vx.p, vy.p,
            vz.p, factor);
    }

    // ------------------------------------------------------------------ neighbor geometry
    real bindist(int i, int j, int k) const // verletlist/neighbor.c:267-296
    {
        real delx = i > 0 ? (i - 1) * binsizex : (i == 0 ? (real)0.0 : (i + 1) * binsizex);
        real dely = j > 0 ? (j - 1) * binsizey : (j == 0 ? (real)0.0 : (j + 1) * binsizey);
        real delz = k > 0 ? (k - 1) * binsizez : (k == 0 ? (real)0.0 : (k + 1) * binsizez);
        return (delx * delx + dely * dely + delz * delz);
    }
    void setupNeighbor() override // verletlist/neighbor.c:43-62 + 64-184
    {
        const real SMALL = (real)1.0e-6, FACTOR = (real)0.999;
        const real neighscale = (real)(5.0 / 6.0);
        real bx = P.nx * lattice, by = P.ny * lattice, bz = P.nz * lattice;
        bg.nbinx = (int)(neighscale * P.nx);
        bg.nbiny = (int)(neighscale * P.ny);
        bg.nbinz = (int)(neighscale * P.nz);
        if (P.from_input) { bx = xprd; by = yprd; bz = zprd; }
        const real lox = 0, hix = bx, loy = 0, hiy = by, loz = 0, hiz = bz;
        cutneighsq = cutneigh * cutneigh;
        if (P.from_input) {
            binsizex = binsizey = binsizez = (real)((double)cutneigh * 0.5);
            bg.nbinx = (int)((xhi - xlo) / binsizex);
            bg.nbiny = (int)((yhi - ylo) / binsizey);
            bg.nbinz = (int)((zhi - zlo) / binsizez);
            if (bg.nbinx == 0) bg.nbinx = 1;
            if (bg.nbiny == 0) bg.nbiny = 1;
            if (bg.nbinz == 0) bg.nbinz = 1;
            bg.bininvx = bg.nbinx / (xhi - xlo);
            bg.bininvy = bg.nbiny / (yhi - ylo);
            bg.bininvz = bg.nbinz / (zhi - zlo);
        } else {
            binsizex   = bx / bg.nbinx;
            binsizey   = by / bg.nbiny;
            binsizez   = bz / bg.nbinz;
            bg.bininvx = (real)(1.0 / (double)binsizex);
            bg.bininvy = (real)(1.0 / (double)binsizey);
            bg.bininvz = (real)(1.0 / (double)binsizez);
        }
        int mhix, mhiy, mhiz;
        real coord;
        coord      = lox - cutneigh - SMALL * bx;
        bg.mbinxlo = (int)(coord * bg.bininvx);
        if (coord < (real)0.0) bg.mbinxlo -= 1;
        coord = hix + cutneigh + SMALL * bx;
        mhix  = (int)(coord * bg.bininvx);
        coord      = loy - cutneigh - SMALL * by;
        bg.mbinylo = (int)(coord * bg.bininvy);
        if (coord < (real)0.0) bg.mbinylo -= 1;
        coord = hiy + cutneigh + SMALL * by;
        mhiy  = (int)(coord * bg.bininvy);
        coord      = loz - cutneigh - SMALL * bz;
        bg.mbinzlo = (int)(coord * bg.bininvz);
        if (coord < (real)0.0) bg.mbinzlo -= 1;
        coord = hiz + cutneigh + SMALL * bz;
        mhiz  = (int)(coord * bg.bininvz);
        bg.mbinxlo -= 1; mhix += 1; bg.mbinx = mhix - bg.mbinxlo + 1;
        bg.mbinylo -= 1; mhiy += 1; bg.mbiny = mhiy - bg.mbinylo + 1;
        bg.mbinzlo -= 1; mhiz += 1; bg.mbinz = mhiz - bg.mbinzlo + 1;
        int nextx = (int)(cutneigh * bg.bininvx);
        if (nextx * binsizex < FACTOR * cutneigh) nextx++;
        int nexty = (int)(cutneigh * bg.bininvy);
        if (nexty * binsizey < FACTOR * cutneigh) nexty++;
        int nextz = (int)(cutneigh * bg.bininvz);
        if (nextz * binsizez < FACTOR * cutneigh) nextz++;
        h_stencil.clear();
        for (int k = -nextz; k <= nextz; k++)
            for (int j = -nexty; j <= nexty; j++)
                for (int i = -nextx; i <= nextx; i++)
                    if (bindist(i, j, k) < cutneighsq)
                        h_stencil.push_back(k * bg.mbiny * bg.mbinx + j * bg.mbinx + i);
        nstencil       = (int)h_stencil.size();
        const long long mb = (long long)bg.mbinx * bg.mbiny * bg.mbinz;
        if (mb > 2000000000LL) throw Error("setupNeighbor: too many bins");
        bg.mbins = (int)mb;
        bg.xprd = bx; bg.yprd = by; bg.zprd = bz;
        if (!P.from_input) { xprd = bx; yprd = by; zprd = bz; }
        stencil.ensure(nstencil, false, stream);
        MDB_CUDA(cudaMemcpyAsync(stencil.p, h_stencil.data(), nstencil * sizeof(int), cudaMemcpyHostToDevice, stream));
        // runs of consecutive offsets (x-adjacent bins are adjacent in the CSR): the list build walks 21 runs instead of 81 bins
        // the bins' real widths are 1 / bininv (binsize itself differs for from_input, neighbor.c:78-92)
        rg.bsx = (float)(1.0 / (double)bg.bininvx); rg.bsy = (float)(1.0 / (double)bg.bininvy);
        rg.bsz = (float)(1.0 / (double)bg.bininvz);
        rg.binvx    = (float)bg.bininvx;
        rg.cutsq_hi = (float)((double)cutneighsq * (1.0 + 1e-4));
        rg.margin   = 1e-3f * std::min({ rg.bsx, rg.bsy, rg.bsz });
        std::vector<RunRow> rr;
        for (int k = -nextz; k <= nextz; k++)
            for (int j = -nexty; j <= nexty; j++) {
                int i0 = 0, len = 0;
                for (int i = -nextx; i <= nextx; i++)
                    if (bindist(i, j, k) < cutneighsq) {
                        if (len == 0) i0 = i;
                        else if (i != i0 + len) throw Error("setupNeighbor: stencil row is not one run");
                        len++;
                    }
                if (len == 0) continue;
                RunRow q;
                q.base = k * bg.mbiny * bg.mbinx + j * bg.mbinx;
                q.i0 = i0; q.i1 = i0 + len - 1; q.pad = 0;
                const double m = rg.margin;
                q.ay = (float)(j > 0 ? j * (double)rg.bsy - m : (j < 0 ? -(j + 1) * (double)rg.bsy - m : -m));
                q.sy = j > 0 ? -1.0f : (j < 0 ? 1.0f : 0.0f);
                q.az = (float)(k > 0 ? k * (double)rg.bsz - m : (k < 0 ? -(k + 1) * (double)rg.bsz - m : -m));
                q.sz = k > 0 ? -1.0f : (k < 0 ? 1.0f : 0.0f);
                rr.push_back(q);
            }
        nruns = (int)rr.size();
        runs.ensure(nruns, false, stream);
        MDB_CUDA(cudaMemcpyAsync(runs.p, rr.data(), nruns * sizeof(RunRow), cudaMemcpyHostToDevice, stream));
        MDB_CUDA(cudaStreamSynchronize(stream));
        bincount.ensure(bg.mbins + 2, false, stream);
        binstart.ensure(bg.mbins + 3, false, stream);
        cursor.ensure(bg.mbins + 2, false, stream);
        neigh_ready = true;
    }

    // ------------------------------------------------------------------ spatial sort
    // sortAtom (verletlist/neighbor.c:360-426; main.c:63-66,82-88) -- always on here, because on
    // the GPU it is what makes the neighbor gathers coalesce: counting sort of the local atoms by
    // bin, ties broken by reference index (so the order is a pure function of positions and ids).
    void sort_atoms()
    {
        if (!sort_enabled || Nlocal == 0) return;
        if (!neigh_ready) setupNeighbor();
        const int nb = bg.mbins + 1;
        atom_bin.ensure(Nlocal, false, stream);
        binatoms.ensure(Nlocal, false, stream);
        MDB_CUDA(cudaMemsetAsync(bincount.p, 0, (nb + 1) * sizeof(int), stream));
        MDB_CUDA(cudaMemsetAsync(cursor.p, 0, (nb + 1) * sizeof(int), stream));
        if (sort_block > 0) build_bin_rank();
        MDB_LAUNCH(launches, k_bin_count<real>, grid_for(Nlocal, 256), 256, 0, stream, Nlocal, bg, x.p, y.p, z.p,
            sort_block > 0 ? bin_rank.p : (const int*)nullptr, atom_bin.p, bincount.p);
        scanner.exclusive(bincount.p, binstart.p, nb, binstart.p + nb, stream);
        MDB_LAUNCH(launches, k_bin_fill, grid_for(Nlocal, 256), 256, 0, stream, Nlocal, atom_bin.p, binstart.p,
            cursor.p, binatoms.p);
        MDB_LAUNCH(launches, k_bin_sort, grid_for(nb, 128), 128, 0, stream, nb, binstart.p, binatoms.p, orig.p,
            d_flags.p + 3);
        const size_t cap = x.cap;
        for (DBuf<real>* b : { &x2, &y2, &z2, &vx2, &vy2, &vz2, &fx2, &fy2, &fz2 }) b->ensure(cap, false, stream);
        type2.ensure(cap, false, stream);
        orig2.ensure(cap, false, stream);
        MDB_LAUNCH(launches, k_permute_atoms<real>, grid_for(Nlocal, 256), 256, 0, stream, Nlocal, binatoms.p, x.p,
            y.p, z.p, vx.p, vy.p, vz.p, fx.p, fy.p, fz.p, type.p, orig.p, x2.p, y2.p, z2.p, vx2.p, vy2.p, vz2.p,
            fx2.p, fy2.p, fz2.p, type2.p, orig2.p);
        std::swap(x, x2); std::swap(y, y2); std::swap(z, z2);
        std::swap(vx, vx2); std::swap(vy, vy2); std::swap(vz, vz2);
        std::swap(fx, fx2); std::swap(fy, fy2); std::swap(fz, fz2);
        std::swap(type, type2); std::swap(orig, orig2);
        extmap_valid = false;
    }

    // Sort order of the bins for sort_atoms (option sort_block = B > 0): blocks of B x B x B bins, x-fastest inside a block and
    // between blocks, so that 128 consecutive atoms (one thread block of the force kernel) fill a compact box like the
    // generator's 8x8x8 sub-boxes do instead of a stick of ~18 bins along x.  The rank of every bin is a closed form
    // (k_block_rank); the table is rebuilt only when the bin grid or B changes.
    void build_bin_rank()
    {
        const long long key = (((long long)bg.mbinx * 4096 + bg.mbiny) * 4096 + bg.mbinz) * 64 + sort_block;
        if (key == bin_rank_key) return;
        const int nb = bg.mbins + 1;
        bin_rank.ensure(nb, false, stream);
        MDB_LAUNCH(launches, k_block_rank, grid_for(nb, 256), 256, 0, stream, nb, bg.mbinx, bg.mbiny, bg.mbinz, sort_block, bin_rank.p);
        bin_rank_key = key;
    }

    // internal index -> reference index for locals AND ghosts.  The reference numbers ghosts in the
    // order of its serial setupPbc loop: by source atom (reference index), then by position in the
    // ADDGHOST ladder (pbc.c:107-224).  Only needed for parity read-back, so done lazily on the host.
    void build_extmap()
    {
        if (extmap_valid) return;
        static const signed char img[26][3] = { { +1, 0, 0 }, { -1, 0, 0 }, { 0, +1, 0 }, { 0, -1, 0 },
            { 0, 0, +1 }, { 0, 0, -1 }, { +1, +1, +1 }, { +1, -1, +1 }, { +1, +1, -1 }, { +1, -1, -1 },
            { -1, +1, +1 }, { -1, -1, +1 }, { -1, +1, -1 }, { -1, -1, -1 }, { +1, 0, +1 }, { +1, 0, -1 },
            { -1, 0, +1 }, { -1, 0, -1 }, { 0, +1, +1 }, { 0, +1, -1 }, { 0, -1, +1 }, { 0, -1, -1 },
            { +1, +1, 0 }, { -1, +1, 0 }, { +1, -1, 0 }, { -1, -1, 0 } };
        int rank_of_code[64];
        for (int b = 0; b < 26; b++)
            rank_of_code[(img[b][0] + 1) | ((img[b][1] + 1) << 2) | ((img[b][2] + 1) << 4)] = b;
        h_orig.resize(Nlocal);
        h_bm.resize(Nghost);
        h_code.resize(Nghost);
        MDB_CUDA(cudaMemcpyAsync(h_orig.data(), orig.p, Nlocal * sizeof(int), cudaMemcpyDeviceToHost, stream));
        if (Nghost) {
            MDB_CUDA(cudaMemcpyAsync(h_bm.data(), border_map.p, Nghost * sizeof(int), cudaMemcpyDeviceToHost, stream));
            MDB_CUDA(cudaMemcpyAsync(h_code.data(), ghost_code.p, Nghost * sizeof(int), cudaMemcpyDeviceToHost, stream));
        }
        MDB_CUDA(cudaStreamSynchronize(stream));
        std::vector<long long> key(Nghost);
        h_ghost_order.resize(Nghost);
        for (int g = 0; g < Nghost; g++) {
            key[g]           = (long long)h_orig[h_bm[g]] * 32 + rank_of_code[h_code[g] & 63];
            h_ghost_order[g] = g;
        }
        std::sort(h_ghost_order.begin(), h_ghost_order.end(), [&](int a, int b) { return key[a] < key[b]; });
        std::vector<int> h_ext((size_t)Nlocal + Nghost);
        for (int p = 0; p < Nlocal; p++) h_ext[p] = h_orig[p];
        for (int r = 0; r < Nghost; r++) h_ext[(size_t)Nlocal + h_ghost_order[r]] = Nlocal + r;
        extmap.ensure(h_ext.size(), false, stream);
        MDB_CUDA(cudaMemcpyAsync(extmap.p, h_ext.data(), h_ext.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
        MDB_CUDA(cudaStreamSynchronize(stream));
        extmap_valid = true;
    }

    // ------------------------------------------------------------------ PBC
    PbcGeom<real> pbc_geom() const
    {
        PbcGeom<real> g;
        g.xprd = xprd; g.yprd = yprd; g.zprd = zprd; g.cutneigh = cutneigh;
        g.xhi_cut = xprd - cutneigh; g.yhi_cut = yprd - cutneigh; g.zhi_cut = zprd - cutneigh;
        g.pbc_x = P.pbc_x; g.pbc_y = P.pbc_y; g.pbc_z = P.pbc_z;
        return g;
    }
    void setupPbc() override // verletlist/pbc.c:98-227
    {
        ghost_msk.ensure(Nlocal, false, stream);
        ghost_cnt.ensure(Nlocal, false, stream);
        ghost_off.ensure(Nlocal, false, stream);
        MDB_LAUNCH(launches, k_ghost_count<real>, grid_for(Nlocal, 256), 256, 0, stream, Nlocal, pbc_geom(),
            x.p, y.p, z.p, ghost_msk.p, ghost_cnt.p);
        scanner.exclusive(ghost_cnt.p, ghost_off.p, Nlocal, d_flags.p + 0, stream);
        MDB_CUDA(cudaMemcpyAsync(h_flags, d_flags.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
        MDB_CUDA(cudaStreamSynchronize(stream));
        Nghost = h_flags[0];
        ensure_atoms((size_t)Nlocal + Nghost, true); // growAtom, atom.c:590-618
        border_map.ensure(Nghost, false, stream);    // growPbc, pbc.c:230-242
        ghost_code.ensure(Nghost, false, stream);
        MDB_LAUNCH(launches, k_ghost_fill, grid_for(Nlocal, 256), 256, 0, stream, Nlocal, ghost_msk.p,
            ghost_off.p, border_map.p, ghost_code.p, type.p);
        extmap_valid = false;
        ghost_tables_valid = true; // ghost_msk / ghost_off describe the ghosts of the atoms in their current slots
        ghosts_current     = false;
    }
    void updatePbc() override // verletlist/pbc.c:42-55
    {
        if (Nghost == 0) return;
        if (ghosts_current) return; // the fused force + integrate kernel has already written the images of the new positions
        MDB_LAUNCH(launches, k_update_pbc<real>, grid_for(Nghost, 256), 256, 0, stream, Nlocal, Nghost, xprd,
            yprd, zprd, border_map.p, ghost_code.p, x.p, y.p, z.p, xy_valid ? xy.p : (vec2*)nullptr);
    }
    void updateAtomsPbc() override // verletlist/pbc.c:59-84
    {
        MDB_LAUNCH(launches, k_update_atoms_pbc<real>, grid_for(Nlocal, 256), 256, 0, stream, Nlocal, xprd,
            yprd, zprd, x.p, y.p, z.p);
    }

    // ------------------------------------------------------------------ neighbor build
    void bin_atoms() // binatoms, verletlist/neighbor.c:329-358
    {
        const int nall = Nlocal + Nghost;
        const int nb   = bg.mbins + 1; // coord2bin's "+ 1" can reach index mbins
        atom_bin.ensure(nall, false, stream);
        binatoms.ensure(nall, false, stream);
        MDB_CUDA(cudaMemsetAsync(bincount.p, 0, (nb + 1) * sizeof(int), stream));
        MDB_CUDA(cudaMemsetAsync(cursor.p, 0, (nb + 1) * sizeof(int), stream));
        MDB_CUDA(cudaMemsetAsync(d_flags.p + 1, 0, 2 * sizeof(int), stream));
        MDB_LAUNCH(launches, k_bin_count<real>, grid_for(nall, 256), 256, 0, stream, nall, bg, x.p, y.p, z.p,
            (const int*)nullptr, atom_bin.p, bincount.p);
        scanner.exclusive(bincount.p, binstart.p, nb, binstart.p + nb, stream);
        MDB_LAUNCH(launches, k_bin_fill, grid_for(nall, 256), 256, 0, stream, nall, atom_bin.p, binstart.p,
            cursor.p, binatoms.p);
        MDB_LAUNCH(launches, k_bin_sort, grid_for(nb, 128), 128, 0, stream, nb, binstart.p, binatoms.p,
            (const int*)nullptr, d_flags.p + 2);
    }

    // Bounds of the single-precision pre-test of k_build_neighbor_v2: a candidate whose float rsq is
    // below lo is certainly inside cutneighsq, above hi certainly outside.  Error model: each float
    // coordinate is off by <= M*2^-24 (M = largest |coordinate|), so each difference by
    // E = 2*M*2^-24 + d*2^-23; |rsq_f - rsq| <= 2*sqrt(3)*d*E + 3*E^2 + O(2^-22)*rsq with d <= cutneigh.
    // The margin takes 8x that bound.  For SP the band collapses to the exact test itself.
    void list_margin(float& lo, float& hi) const
    {
        const double M = std::max({ (double)xprd, (double)yprd, (double)zprd }) + 2.0 * (double)cutneigh;
        const double d = (double)cutneigh * 1.01;
        const double E = 2.0 * M * ldexp(1.0, -24) + d * ldexp(1.0, -23);
        const double m = 8.0 * (2.0 * 1.7320508 * d * E + 3.0 * E * E) + (double)cutneighsq * 2e-6;
        lo = (float)((double)cutneighsq - m);
        hi = (float)((double)cutneighsq + m);
        if (sizeof(real) == 4) { lo = -1.0f; } // SP: always run the exact (float) expression
    }
    void buildNeighbor() override // verletlist/neighbor.c:186-264
    {
        if (!neigh_ready) setupNeighbor();
        float ms = 0;
        if (timing) MDB_CUDA(cudaEventRecord(evA, stream));
        bin_atoms();
        // candidates in CSR order as SoA float arrays, padded to a multiple of 4
        const int nall = Nlocal + Nghost, npad = (int)round_up((size_t)nall + 4, 4);
        for (DBuf<float>* b : { &cxs, &cys, &czs }) b->ensure(npad, false, stream);
        cids.ensure(npad, false, stream);
        MDB_LAUNCH(launches, k_pack_binned_soa<real>, grid_for(npad, 256), 256, 0, stream, nall, npad, binatoms.p, x.p, y.p,
            z.p, cxs.p, cys.p, czs.p, cids.p);
        nstride = round_up((size_t)Nlocal, 32);
        numneigh.ensure(nstride, false, stream);
        float lo, hi;
        list_margin(lo, hi);
        for (;;) {
            const size_t rowlen = round_up((size_t)maxneighs, 8);
            LL = NbLayout { 32 * rowlen, 32, 5 };
            neighbors.ensure(rowlen * nstride, false, stream);
            MDB_CUDA(cudaMemsetAsync(d_flags.p + 1, 0, sizeof(int), stream));
#define MDB_BUILD_V6(H)                                                                                                        \
    MDB_LAUNCH(launches, (k_build_neighbor_v6<real, H>), grid_for(Nlocal, 128), 128, 0, stream, Nlocal, bg, rg, cutneighsq, lo,  \
        hi, x.p, y.p, z.p, cxs.p, cys.p, czs.p, cids.p, binstart.p, runs.p, nruns, maxneighs, LL, orig.p, \
        numneigh.p, neighbors.p, d_flags.p + 1)
            if (LL.sk != 32) throw Error("buildNeighbor: the list build stores rows at a stride of 32 entries");
            if (P.half_neigh) MDB_BUILD_V6(true);
            else MDB_BUILD_V6(false);
#undef MDB_BUILD_V6
            neigh_launches++;
            MDB_CUDA(cudaMemcpyAsync(h_flags + 1, d_flags.p + 1, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream));
            MDB_CUDA(cudaStreamSynchronize(stream));
            max_bin_count = h_flags[2];
            if (h_flags[1] >= maxneighs) { // neighbor.c:247-262
                maxneighs = (int)(h_flags[1] * 1.2);
                continue;
            }
            break;
        }
        if (timing) {
            MDB_CUDA(cudaEventRecord(evB, stream));
            MDB_CUDA(cudaEventSynchronize(evB));
            MDB_CUDA(cudaEventElapsedTime(&ms, evA, evB));
            neigh_ms += ms;
        }
    }

    // ------------------------------------------------------------------ force
    void launch_force(int which)
    {
        NvtxRange nvtx_range_("force");
        if (nstride == 0) throw Error("computeForce: no neighbor list (call mdb_setup or mdb_buildNeighbor first)");
        if (which == FORCE_DISPATCH)
            which = P.force_field == MDB_FF_EAM ? FORCE_EAM : (P.half_neigh ? FORCE_LJ_HALF : FORCE_LJ_FULL);
        if (timing) MDB_CUDA(cudaEventRecord(ev0, stream));
        if (which == FORCE_EAM) {
            launch_eam();
        } else {
            LJConst2<real> c2 { cutforce * cutforce, (real)48.0 * epsilon * sigma6 * sigma6, (real)24.0 * epsilon * sigma6 };
            if (which == FORCE_LJ_FULL) {
                if (sizeof(real) == 4) // SP: the branch-free block (1.00 vs 1.13 ms, profiles/r1_ab3.txt)
                    MDB_LAUNCH(launches, (k_force_lj_full_v6<real, 4>), grid_for(Nlocal, 128), 128, 0, stream, Nlocal, c2,
                        x.p, y.p, z.p, numneigh.p, neighbors.p, LL, fx.p, fy.p, fz.p);
                else // DP: the divergent block (the branch-free one pays 1.617 vs 1.537 ms for the FP64 work outside the cutoff)
                    MDB_LAUNCH(launches, (k_force_lj_full_v2<real, 4>), grid_for(Nlocal, 128), 128, 0, stream,
                        Nlocal, c2, x.p, y.p, z.p, numneigh.p, neighbors.p, LL, fx.p, fy.p, fz.p);
            } else {
                zero3(fx.p, fy.p, fz.p, Nlocal);
                MDB_LAUNCH(launches, (k_force_lj_half_v2<real, 4>), grid_for(Nlocal, 128), 128, 0, stream, Nlocal, c2,
                    x.p, y.p, z.p, numneigh.p, neighbors.p, LL, fx.p, fy.p, fz.p);
            }
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
    void eam_density() // force_eam.c:49-112
    {
        if (!eam.ready) throw Error("computeForceEam: no EAM tables (call mdb_setEam first)");
        fp.ensure((size_t)Nlocal + Nghost, false, stream);
        if (eam_variant >= 1) // (generation 3 is single-domain only: a brick runs generation 2)
            MDB_LAUNCH(launches, (k_eam_density_v2<real, 2>), grid_for(Nlocal, 128), 128, 0, stream, Nlocal, cutforce * cutforce, eam,
                eam_rho4.p, frho_spline.p, x.p, y.p, z.p, numneigh.p, neighbors.p, LL, fp.p);
        else
        MDB_LAUNCH(launches, k_eam_density<real>, grid_for(Nlocal, 128), 128, 0, stream, Nlocal, cutforce * cutforce,
            eam, rhor_spline.p, frho_spline.p, x.p, y.p, z.p, numneigh.p, neighbors.p, LL, fp.p);
    }
    void eam_force() // force_eam.c:127-224
    {
        if (eam_variant >= 1)
            MDB_LAUNCH(launches, (k_eam_force_v2<real, 2>), grid_for(Nlocal, 128), 128, 0, stream, Nlocal, cutforce * cutforce, eam,
                eam_frc12.p, x.p, y.p, z.p, fp.p, numneigh.p, neighbors.p, LL, fx.p, fy.p, fz.p);
        else
        MDB_LAUNCH(launches, k_eam_force<real>, grid_for(Nlocal, 128), 128, 0, stream, Nlocal, cutforce * cutforce,
            eam, rhor_spline.p, z2r_spline.p, x.p, y.p, z.p, fp.p, numneigh.p, neighbors.p, LL, fx.p, fy.p, fz.p);
    }
    void launch_eam_v3(bool integrate = false) // generation 3: see eam_kernels.cuh
    {
        if (!eam.ready) throw Error("computeForceEam: no EAM tables (call mdb_setEam first)");
        const int nall = Nlocal + Nghost;
        fp.ensure((size_t)nall, false, stream);
        xy.ensure(x.cap, false, stream);
        eam_zf.ensure(x.cap, false, stream);
        xy_valid = ghosts_current = false; // the packed (x, y) copy is rebuilt for every force call here
        MDB_LAUNCH(launches, k_pack_xy<real>, grid_for(nall, 256), 256, 0, stream, nall, x.p, y.p, xy.p);
        // U neighbors in flight, the next group's indices requested before the current group is evaluated (A/B at 128^3,
        // profiles/r2_s3_call8.sh: density U = 2 -> 3 + prefetch 4.87 -> 4.81 ms per force call, force pass + prefetch 4.87 -> 4.62 ms)
        MDB_LAUNCH(launches, (k_eam_density_v3<real, 3, true>), grid_for(Nlocal, 128), 128, 0, stream, Nlocal, cutforce * cutforce, eam,
            eam_rho4.p, frho_spline.p, x.p, y.p, z.p, xy.p, numneigh.p, neighbors.p, LL, fp.p, eam_zf.p);
        if (Nghost) // force_eam.c:118-120
            MDB_LAUNCH(launches, k_eam_ghost_fp_v3<real>, grid_for(Nghost, 256), 256, 0, stream, Nlocal, Nghost, border_map.p, z.p,
                fp.p, eam_zf.p);
        EamIntegrate<real> ei { vx.p, vy.p, vz.p, dtforce, dt };
        if (integrate)
            MDB_LAUNCH(launches, (k_eam_force_v3<real, 2, true, true>), grid_for(Nlocal, 128), 128, 0, stream, Nlocal, cutforce * cutforce,
                eam, eam_vs4.p, x.p, y.p, z.p, fp.p, xy.p, eam_zf.p, numneigh.p, neighbors.p, LL, fx.p, fy.p, fz.p, ei);
        else
            MDB_LAUNCH(launches, (k_eam_force_v3<real, 2, true>), grid_for(Nlocal, 128), 128, 0, stream, Nlocal, cutforce * cutforce,
                eam, eam_vs4.p, x.p, y.p, z.p, fp.p, xy.p, eam_zf.p, numneigh.p, neighbors.p, LL, fx.p, fy.p, fz.p, ei);
    }
    // EAM: computeForce(n) + finalIntegrate(n) + initialIntegrate(n+1), the integrate halves in the epilogue of the force pass
    bool can_fuse_eam() const
    {
        return fuse_force && fuse_integrate && !brick && P.force_field == MDB_FF_EAM && eam_variant == 2 && !lazy_ops;
    }
    void eamForceFinalInitialIntegrate()
    {
        NvtxRange nvtx_range_("force+integrate");
        if (nstride == 0) throw Error("computeForce: no neighbor list (call mdb_buildNeighbor first)");
        if (timing) MDB_CUDA(cudaEventRecord(ev0, stream));
        launch_eam_v3(true);
        force_launches++;
        if (timing) {
            float ms = 0;
            MDB_CUDA(cudaEventRecord(ev1, stream));
            MDB_CUDA(cudaEventSynchronize(ev1));
            MDB_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
            force_ms += ms;
        }
    }
    void launch_eam()
    {
        if (brick) throw Error("computeForceEam on a brick goes through the decomposition (fp exchange)");
        if (eam_variant == 2) { launch_eam_v3(); return; }
        eam_density();
        if (Nghost) // force_eam.c:118-120
            MDB_LAUNCH(launches, k_eam_ghost_fp<real>, grid_for(Nghost, 256), 256, 0, stream, Nlocal, Nghost,
                border_map.p, fp.p);
        eam_force();
    }
    // ComputeForceFunction: returns elapsed seconds like the reference (force.h:16)
    double computeForce(int which) override
    {
        if (nstride == 0) throw Error("computeForce: no neighbor list (call mdb_buildNeighbor first)");
        const bool t = timing;
        timing       = true;
        const double before = force_ms;
        launch_force(which);
        timing = t;
        const double el = (force_ms - before) * 1e-3;
        if (!t) force_ms = before;
        return el;
    }

    // ------------------------------------------------------------------ integrate
    void initialIntegrate() override // verletlist/integrate.c:21-31
    {
        xy_valid = ghosts_current = false;
        MDB_LAUNCH(launches, k_initial_integrate<real>, grid_for(Nlocal, 256), 256, 0, stream, Nlocal,
            dtforce, dt, x.p, y.p, z.p, vx.p, vy.p, vz.p, fx.p, fy.p, fz.p);
    }
    void finalIntegrate() override // verletlist/integrate.c:33-40
    {
        MDB_LAUNCH(launches, k_final_integrate<real>, grid_for(Nlocal, 256), 256, 0, stream, Nlocal, dtforce,
            vx.p, vy.p, vz.p, fx.p, fy.p, fz.p);
    }

    // computeForce(n) + finalIntegrate(n) + initialIntegrate(n+1) in ONE launch (k_force_lj_full_fi): LJ full lists of a
    // single domain with the default kernels.  The new positions land in the sort buffers x2/y2/z2, which then become
    // x/y/z; their ghost range is rewritten by the updatePbc / setupPbc of the next step before anything reads it.
    bool can_fuse_force() const
    {
        return fuse_force && fuse_integrate && !brick && P.force_field != MDB_FF_EAM && !P.half_neigh;
    }
    void forceFinalInitialIntegrate()
    {
        NvtxRange nvtx_range_("force+integrate");
        if (nstride == 0) throw Error("computeForce: no neighbor list (call mdb_buildNeighbor first)");
        x2.ensure(x.cap, false, stream);
        y2.ensure(y.cap, false, stream);
        z2.ensure(z.cap, false, stream);
        if (timing) MDB_CUDA(cudaEventRecord(ev0, stream));
        LJConst2<real> c2 { cutforce * cutforce, (real)48.0 * epsilon * sigma6 * sigma6, (real)24.0 * epsilon * sigma6 };
        const bool use_xy = xy_gather != 0;
        if (use_xy) {
            xy.ensure(x.cap, false, stream);
            xy2.ensure(x.cap, false, stream);
            if (!xy_valid) { // locals and ghosts as they are now
                const int nall = Nlocal + Nghost;
                MDB_LAUNCH(launches, k_pack_xy<real>, grid_for(nall, 256), 256, 0, stream, nall, x.p, y.p, xy.p);
            }
        }
        // the ghosts of the NEXT step are written by the same launch (single domain, ghost tables of the last setupPbc)
        // worth it where a step is launch-bound (32^3: 2.69 -> 2.93 G atom-steps/s); at 128^3 the extra mask load of every atom
        // costs more than the 10 us kernel it replaces (1.478 -> 1.499 ms), hence the size threshold of the default
        const bool want_ghosts = ghost_epilogue < 0 ? Nlocal <= (1 << 19) : ghost_epilogue != 0;
        const bool own_ghosts  = want_ghosts && !brick && ghost_tables_valid && Nghost > 0;
        FusedIntegrate<real> fi { vx.p, vy.p, vz.p, x2.p, y2.p, z2.p, dtforce, dt, xy.p, xy2.p, nullptr, nullptr,
            own_ghosts ? ghost_msk.p : (const unsigned*)nullptr, own_ghosts ? ghost_off.p : (const int*)nullptr, xprd, yprd, zprd };
        if (use_xy)
            MDB_LAUNCH(launches, (k_force_lj_full_fi<real, 4, sizeof(real) == 4, true>), grid_for(Nlocal, 128), 128, 0, stream,
                Nlocal, c2, x.p, y.p, z.p, numneigh.p, neighbors.p, LL, fi);
        else // SP: branch-free force block (v6), DP: the divergent block of v2
            MDB_LAUNCH(launches, (k_force_lj_full_fi<real, 4, sizeof(real) == 4>), grid_for(Nlocal, 128), 128, 0, stream, Nlocal,
                c2, x.p, y.p, z.p, numneigh.p, neighbors.p, LL, fi);
        std::swap(x, x2); std::swap(y, y2); std::swap(z, z2);
        if (use_xy) std::swap(xy, xy2);
        xy_valid = use_xy; // locals current; the ghost range follows with the next updatePbc
        ghosts_current = own_ghosts; // ... unless this launch has written it already
        force_launches++;
        if (timing) {
            float ms = 0;
            MDB_CUDA(cudaEventRecord(ev1, stream));
            MDB_CUDA(cudaEventSynchronize(ev1));
            MDB_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
            force_ms += ms;
        }
    }
    // The same fused step for a brick of a decomposed run (also valid for a single domain): x, y, z are updated IN PLACE --
    // the neighbor GPUs hold pointers into their ghost range -- and every gather goes to the double-buffered copies
    // xy / zg.  Call after the halo of this step has landed (forward() / reneighbour()).
    bool can_fuse_force_inplace() const
    {
        return fuse_force && fuse_integrate && P.force_field != MDB_FF_EAM && !P.half_neigh;
    }
    void forceFinalInitialIntegrateInPlace()
    {
        NvtxRange nvtx_range_("force+integrate");
        if (nstride == 0) throw Error("computeForce: no neighbor list (call mdb_buildNeighbor first)");
        for (DBuf<vec2>* b : { &xy, &xy2 }) b->ensure(x.cap, false, stream);
        for (DBuf<real>* b : { &zg, &zg2 }) b->ensure(x.cap, false, stream);
        if (timing) MDB_CUDA(cudaEventRecord(ev0, stream));
        // gather copies: everything after a rebuild / an unfused step, else only the ghost range that just arrived
        const int first = xy_valid ? Nlocal : 0, cnt = xy_valid ? Nghost : Nlocal + Nghost;
        if (cnt > 0)
            MDB_LAUNCH(launches, k_pack_gather<real>, grid_for(cnt, 256), 256, 0, stream, first, cnt, x.p, y.p, z.p, xy.p, zg.p);
        LJConst2<real> c2 { cutforce * cutforce, (real)48.0 * epsilon * sigma6 * sigma6, (real)24.0 * epsilon * sigma6 };
        FusedIntegrate<real> fi { vx.p, vy.p, vz.p, x.p, y.p, z.p, dtforce, dt, xy.p, xy2.p, zg.p, zg2.p, nullptr, nullptr, xprd, yprd, zprd };
        MDB_LAUNCH(launches, (k_force_lj_full_fi<real, 4, sizeof(real) == 4, true, true>), grid_for(Nlocal, 128), 128, 0, stream,
            Nlocal, c2, x.p, y.p, z.p, numneigh.p, neighbors.p, LL, fi);
        std::swap(xy, xy2);
        std::swap(zg, zg2);
        xy_valid = true;
        force_launches++;
        if (timing) {
            float ms = 0;
            MDB_CUDA(cudaEventRecord(ev1, stream));
            MDB_CUDA(cudaEventSynchronize(ev1));
            MDB_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
            force_ms += ms;
        }
    }
    // ---- lazy operators (sim.cuh): the function-pointer loop of the reference gets the fused kernel too ----
    bool lazy_ops = false, pending_force = false, pending_final = false;
    double abi_computeForce(int which) override
    {
        flush_lazy();
        const bool lj_full = which == FORCE_LJ_FULL || (which == FORCE_DISPATCH && P.force_field != MDB_FF_EAM && !P.half_neigh);
        if (lazy_ops && lj_full && can_fuse_force() && nstride != 0) {
            pending_force = true; // due; launched by the next initialIntegrate (fused) or by flush_lazy()
            return 0.0;
        }
        return computeForce(which);
    }
    void abi_finalIntegrate() override
    {
        if (pending_force && !pending_final) { pending_final = true; return; }
        flush_lazy();
        finalIntegrate();
    }
    void abi_initialIntegrate() override
    {
        if (pending_force && pending_final) {
            pending_force = pending_final = false;
            forceFinalInitialIntegrate();
            return;
        }
        flush_lazy();
        initialIntegrate();
    }
    void flush_lazy() override
    {
        if (pending_force) { pending_force = false; launch_force(FORCE_DISPATCH); }
        if (pending_final) { pending_final = false; finalIntegrate(); }
    }
    void drop_lazy() override { pending_force = pending_final = false; }
    void invalidate_copies() override { xy_valid = false; ghosts_current = false; }
    void finalInitialIntegrate() // finalIntegrate(n) + initialIntegrate(n+1) in one pass
    {
        xy_valid = ghosts_current = false;
        MDB_LAUNCH(launches, k_final_initial_integrate<real>, grid_for(Nlocal, 256), 256, 0, stream, Nlocal,
            dtforce, dt, x.p, y.p, z.p, vx.p, vy.p, vz.p, fx.p, fy.p, fz.p);
    }

#include "dd_brick.inc"

    // ------------------------------------------------------------------ driver flow
    void setup(bool adjust) override // verletlist/main.c:58-72
    {
        setupNeighbor();
        thermo_ready = false;
        derive_dtforce();
        setupThermo();
        if (adjust) adjustThermo();
        sort_atoms(); // main.c:63-66 (SORT_ATOMS)
        setupPbc();
        updatePbc();
        buildNeighbor();
    }
    void derive_dtforce()
    {
        dtforce = (real)(0.5 * (double)dt);
        if (P.force_field == MDB_FF_EAM) dtforce = (real)(0.5 * (double)dt / (double)mass); // eam_utils.c:35
    }
    void reneighbour() override // verletlist/main.c:76-95
    {
        NvtxRange nvtx_range_("reneighbour");
        xy_valid = ghosts_current = false;
        updateAtomsPbc();
        sort_atoms(); // main.c:82-88 (SORT_ATOMS; here at every rebuild)
        setupPbc();
        updatePbc();
        buildNeighbor();
    }
    void run(int nsteps, double* thermo_out, int max_records, int* nrecords, double* timers) override
    {
        if (!thermo_ready) setupThermo();
        xy_valid = ghosts_current = false; // positions may have been set from outside since the last run
        const int nstat = P.nstat > 0 ? P.nstat : nsteps + 1;
        const int every = P.reneigh_every > 0 ? P.reneigh_every : nsteps + 1;
        const int maxrec = nsteps / nstat + 3;
        d_thermo.ensure((size_t)4 * maxrec, false, stream);
        std::vector<int> rec_step;
        auto record = [&](int step) {
            vel_sums(d_thermo.p + 4 * rec_step.size());
            rec_step.push_back(step);
        };
        const double f0 = force_ms, n0 = neigh_ms;
        record(0);            // main.c:244
        launch_force(FORCE_DISPATCH); // main.c:250
        MDB_CUDA(cudaEventRecord(run_ev(0), stream)); // timer[TOTAL] starts after the first force, main.c:252
        bool initial_done = false; // step n's initialIntegrate already applied by the fused kernel
        for (int n = 0; n < nsteps; n++) {
            const bool reneigh = (n + 1) % every == 0; // main.c:259
            if (!initial_done) initialIntegrate();
            if (reneigh) reneighbour();
            else updatePbc();
            const bool rec = !((n + 1) % nstat) && (n + 1) < nsteps; // main.c:275-280
            const bool split = rec || n + 1 == nsteps || !fuse_integrate;
            if (!split && can_fuse_force()) {
                forceFinalInitialIntegrate();
                initial_done = true;
                continue;
            }
            if (!split && can_fuse_eam()) {
                eamForceFinalInitialIntegrate();
                initial_done = true;
                continue;
            }
            launch_force(FORCE_DISPATCH);
            if (split) {
                finalIntegrate();
                initial_done = false;
                if (rec) record(n + 1);
            } else {
                finalInitialIntegrate();
                initial_done = true;
            }
        }
        MDB_CUDA(cudaEventRecord(run_ev(1), stream));
        record(nsteps); // computeThermo(-1), main.c:288
        std::vector<double> h(4 * rec_step.size());
        MDB_CUDA(cudaMemcpyAsync(h.data(), d_thermo.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, stream));
        MDB_CUDA(cudaStreamSynchronize(stream));
        float ms = 0;
        MDB_CUDA(cudaEventElapsedTime(&ms, run_ev(0), run_ev(1)));
        int nr = 0;
        for (size_t r = 0; r < rec_step.size(); r++) {
            if (thermo_out && nr < max_records) {
                thermo_out[3 * nr] = rec_step[r];
                thermo_from_sum(h[4 * r + 3], &thermo_out[3 * nr + 1], &thermo_out[3 * nr + 2]);
                nr++;
            }
        }
        if (nrecords) *nrecords = nr;
        if (timers) {
            timers[0] = ms * 1e-3;
            timers[1] = (force_ms - f0) * 1e-3;
            timers[2] = (neigh_ms - n0) * 1e-3;
        }
    }
    cudaEvent_t run_events[2] = { nullptr, nullptr };
    cudaEvent_t run_ev(int k)
    {
        if (!run_events[k]) MDB_CUDA(cudaEventCreate(&run_events[k]));
        return run_events[k];
    }

    // ------------------------------------------------------------------ parity accessors
    void getNeighbors(int* nn, int* nb, int row_stride) override
    {
        if (nstride == 0) throw Error("getNeighbors: no neighbor list");
        build_extmap();
        nn_ext.ensure(Nlocal, false, stream);
        if (nb) rows.ensure((size_t)Nlocal * row_stride, false, stream);
        MDB_LAUNCH(launches, k_untranspose, grid_for(Nlocal, 128), 128, 0, stream, Nlocal, row_stride, LL,
            numneigh.p, neighbors.p, extmap.p, nb ? rows.p : (int*)nullptr, nn_ext.p);
        MDB_CUDA(cudaMemcpyAsync(nn, nn_ext.p, Nlocal * sizeof(int), cudaMemcpyDeviceToHost, stream));
        if (nb)
            MDB_CUDA(cudaMemcpyAsync(nb, rows.p, (size_t)Nlocal * row_stride * sizeof(int),
                cudaMemcpyDeviceToHost, stream));
        MDB_CUDA(cudaStreamSynchronize(stream));
    }
    void getGhostMap(int* bm, int* px, int* py, int* pz) override
    {
        build_extmap();
        for (int r = 0; r < Nghost; r++) { // reference ghost order
            const int g = h_ghost_order[r], code = h_code[g];
            bm[r] = h_orig[h_bm[g]];
            px[r] = (code & 3) - 1;
            py[r] = ((code >> 2) & 3) - 1;
            pz[r] = ((code >> 4) & 3) - 1;
        }
    }
    void getNeighborParams(int* I, double* R) override
    {
        if (!neigh_ready) setupNeighbor();
        const int iv[12] = { bg.nbinx, bg.nbiny, bg.nbinz, bg.mbinx, bg.mbiny, bg.mbinz, bg.mbinxlo,
            bg.mbinylo, bg.mbinzlo, bg.mbins, nstencil, max_bin_count };
        const double rv[14] = { (double)bg.bininvx, (double)bg.bininvy, (double)bg.bininvz, (double)binsizex,
            (double)binsizey, (double)binsizez, (double)cutneighsq, (double)cutneigh, (double)xprd,
            (double)yprd, (double)zprd, (double)lattice, (double)dtforce, (double)cutforce };
        memcpy(I, iv, sizeof iv);
        memcpy(R, rv, sizeof rv);
    }
    void getStencil(int* st) override
    {
        if (!neigh_ready) setupNeighbor();
        memcpy(st, h_stencil.data(), nstencil * sizeof(int));
    }
    void getBinCounts(int* bc) override
    {
        // reference bincount[] is indexed by coord2bin's value directly
        MDB_CUDA(cudaMemcpyAsync(bc, bincount.p, bg.mbins * sizeof(int), cudaMemcpyDeviceToHost, stream));
        MDB_CUDA(cudaStreamSynchronize(stream));
    }
    void countPairs(long long* listed, long long* inside) override
    {
        MDB_CUDA(cudaMemsetAsync(d_cnt.p, 0, 2 * sizeof(unsigned long long), stream));
        MDB_LAUNCH(launches, k_count_pairs<real>, grid_for(Nlocal, 128), 128, 0, stream, Nlocal,
            cutforce * cutforce, x.p, y.p, z.p, numneigh.p, neighbors.p, LL, d_cnt.p);
        unsigned long long h[2];
        MDB_CUDA(cudaMemcpyAsync(h, d_cnt.p, sizeof h, cudaMemcpyDeviceToHost, stream));
        MDB_CUDA(cudaStreamSynchronize(stream));
        *listed = (long long)h[0];
        *inside = (long long)h[1];
    }

    // ------------------------------------------------------------------ EAM tables (host, once)
    void upload_splines(const std::vector<real>& rh, const std::vector<real>& fr, const std::vector<real>& z2)
    {
        rhor_spline.ensure(rh.size(), false, stream);
        frho_spline.ensure(fr.size(), false, stream);
        z2r_spline.ensure(z2.size(), false, stream);
        MDB_CUDA(cudaMemcpyAsync(rhor_spline.p, rh.data(), rh.size() * sizeof(real), cudaMemcpyHostToDevice, stream));
        MDB_CUDA(cudaMemcpyAsync(frho_spline.p, fr.data(), fr.size() * sizeof(real), cudaMemcpyHostToDevice, stream));
        MDB_CUDA(cudaMemcpyAsync(z2r_spline.p, z2.data(), z2.size() * sizeof(real), cudaMemcpyHostToDevice, stream));
        const int rows = (int)(rh.size() / 7); // rows the kernels can index: m <= nr - 1 < nr_tot / 7
        eam_rho4.ensure((size_t)rows * 4, false, stream);
        eam_frc12.ensure((size_t)rows * 12, false, stream);
        MDB_LAUNCH(launches, k_eam_pack_tables<real>, grid_for(rows, 128), 128, 0, stream, rows, rhor_spline.p, z2r_spline.p,
            eam_rho4.p, eam_frc12.p);
        eam_vs4.ensure((size_t)rows * 4, false, stream);
        MDB_LAUNCH(launches, k_eam_pack_vs<real>, grid_for(rows, 128), 128, 0, stream, rows, rhor_spline.p, z2r_spline.p, eam_vs4.p);
        MDB_CUDA(cudaStreamSynchronize(stream));
        h_rhor = rh; h_frho = fr; h_z2r = z2;
        eam.ready = true;
    }
    std::vector<real> h_rhor, h_frho, h_z2r;

    void setEam(int nrho, double drho_, int nr, double dr_, double cut_, double mass_, const double* frho0,
        const double* zr0, const double* rhor0) override
    {
        // initEam overrides, common/eam_utils.c:27-35 (Funcfl fields are MD_FLOAT)
        const real fdrho = (real)drho_, fdr = (real)dr_, fcut = (real)cut_, fmass = (real)mass_;
        P.force_field = MDB_FF_EAM;
        P.mass = (double)fmass; P.cutforce = (double)fcut; P.temp = 600.0; P.dt = 0.001; P.rho = 0.07041125;
        derive();
        cutneigh = (real)((double)cutforce + 1.0);
        derive_dtforce();
        std::vector<real> rh, fr, z2;
        build_eam_tables<real>(nrho, fdrho, nr, fdr, frho0, zr0, rhor0, eam, rh, fr, z2);
        upload_splines(rh, fr, z2);
    }
    void setEamSplines(int nr, int nrho, int nr_tot, int nrho_tot, double rdr, double rdrho, const void* rhor,
        const void* frho, const void* z2r) override
    {
        eam.nr = nr; eam.nrho = nrho; eam.nr_tot = nr_tot; eam.nrho_tot = nrho_tot;
        eam.rdr = (real)rdr; eam.rdrho = (real)rdrho;
        std::vector<real> rh((const real*)rhor, (const real*)rhor + nr_tot),
            fr((const real*)frho, (const real*)frho + nrho_tot), z2((const real*)z2r, (const real*)z2r + nr_tot);
        upload_splines(rh, fr, z2);
    }
    void getEamSplines(int* nr, int* nrho, int* nr_tot, int* nrho_tot, double* rdr, double* rdrho, void* rhor,
        void* frho, void* z2r) override
    {
        if (!eam.ready) throw Error("getEamSplines: no EAM tables");
        *nr = eam.nr; *nrho = eam.nrho; *nr_tot = eam.nr_tot; *nrho_tot = eam.nrho_tot;
        *rdr = (double)eam.rdr; *rdrho = (double)eam.rdrho;
        if (rhor) memcpy(rhor, h_rhor.data(), h_rhor.size() * sizeof(real));
        if (frho) memcpy(frho, h_frho.data(), h_frho.size() * sizeof(real));
        if (z2r) memcpy(z2r, h_z2r.data(), h_z2r.size() * sizeof(real));
    }
    void getEamFp(void* out, bool ghosts) override
    {
        const bool wg  = ghosts && Nghost > 0;
        const size_t n = (size_t)Nlocal + (wg ? Nghost : 0);
        const int* map = orig.p;
        if (wg) { build_extmap(); map = extmap.p; }
        tx.ensure(n, false, stream); ty.ensure(n, false, stream); tz.ensure(n, false, stream);
        MDB_LAUNCH(launches, k_scatter_orig<real>, grid_for(n, 256), 256, 0, stream, (int)n, map, fp.p, fp.p, fp.p,
            tx.p, ty.p, tz.p);
        MDB_CUDA(cudaMemcpyAsync(out, tx.p, n * sizeof(real), cudaMemcpyDeviceToHost, stream));
        MDB_CUDA(cudaStreamSynchronize(stream));
    }
    // synthetic lists of the reference's kernel micro-benchmark (main-stub.c:62-106) instead of a list build
    void stubNeighbors(int pattern, int nneighs, int nreps, unsigned seed) override
    {
        if (Nlocal <= 0) throw Error("mdb_stubNeighbors: no atoms");
        if (pattern < 0 || pattern > 4 || nneighs < 1 || nreps < 1) throw Error("mdb_stubNeighbors: bad pattern / counts");
        if (pattern == 2 && Nlocal <= nneighs)
            throw Error("When using random pattern, number of atoms should be higher than number of neighbors per atom!");
        Nghost    = 0;
        maxneighs = nneighs * nreps;
        nstride   = round_up((size_t)Nlocal, 32);
        const size_t rowlen = round_up((size_t)maxneighs, 8);
        LL = NbLayout { 32 * rowlen, 32, 5 };
        numneigh.ensure(nstride, false, stream);
        neighbors.ensure(rowlen * nstride, false, stream);
        MDB_LAUNCH(launches, k_stub_neighbors, grid_for(Nlocal, 128), 128, 0, stream, Nlocal, pattern, nneighs, nreps, seed, LL,
            numneigh.p, neighbors.p);
        extmap_valid = false;
    }
    void setOption(const char* name, double v) override
    {
        if (!strcmp(name, "sort_atoms")) sort_enabled = v != 0;
        else if (!strcmp(name, "sort_block")) sort_block = std::max(0, std::min(63, (int)v));
        else if (!strcmp(name, "fuse_integrate")) fuse_integrate = v != 0;
        else if (!strcmp(name, "fuse_force")) fuse_force = v != 0;
        else if (!strcmp(name, "xy_gather")) xy_gather = (int)v;
        else if (!strcmp(name, "ghost_epilogue")) ghost_epilogue = (int)v;
        else if (!strcmp(name, "lazy_ops")) { flush_lazy(); lazy_ops = v != 0; }
        else if (!strcmp(name, "eam_variant")) eam_variant = (int)v;
        else throw Error(fmt("mdb_setOption: unknown option '%s'", name));
    }
};

SimBase* make_sim(const mdb_params& p, int device)
{
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        throw Error("mdb_create: no CUDA device (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) throw Error(fmt("mdb_create: device %d out of range (%d devices)", device, ndev));
    if (p.precision == MDB_DP) return new Sim<double>(p, device);
    if (p.precision == MDB_SP) return new Sim<float>(p, device);
    throw Error("mdb_create: precision must be MDB_SP or MDB_DP");
}

#include "dd_group.cuh"

} // namespace mdb
