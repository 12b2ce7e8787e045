// cp_dd.cuh -- CpGroup<real, N>: the CLUSTERPAIR scheme on a spatially decomposed box.
// (textually included at the end of cp_sim.cu, inside namespace mdb, after CpSim)
//
// The bricks of this process are ordinary CpSim domains in brick-local coordinates (dd_topo.h gives the brick grid and
// the transfer schedule, as for the verletlist bricks of dd_group.cuh).  What is generalised is the reference's
// setupPbc / updatePbc / updateAtomsPbc trio for ghost CLUSTERS (clusterpair/pbc.c:183-323, 45-114, 117-144):
//   * setupPbc: every brick selects its border j-clusters exactly like the single domain does (bounding box within
//     cutneigh of a face, ADDGHOST ladder, k_cp_ghost_count) -- image d of tile cj is the tile the brick at coords - s_d
//     sees at x + s_d * ext.  The images are grouped by receiving brick (a stable partition done on the host once per
//     rebuild: ~1e5 small records) and travel as whole tiles, so they land in the receiver's ghost range of cl_x without
//     an unpack step; with one brick per axis the receiver is the brick itself and the scheme is the reference's.
//   * updatePbc: the same send list replayed every step (k_cp_dd_pack<false> -> one grouped exchange).
//   * updateAtomsPbc: atoms that left the brick migrate to the neighbor brick (k_dd_dest of the verletlist bricks).
// Transport: device copies between bricks of one process (how one GPU runs a decomposed box for the parity tests),
// NCCL send/recv over NVLink between processes, one group per exchange.  Thermo sums go through ncclAllReduce.
// Results equal the single domain's to rounding: ghost tiles are numbered by sender instead of by the ladder, which only
// changes the order of equal-z clusters inside a bin, i.e. the summation order.

// images of the listed border tiles into the send buffers (send order): one thread per tile lane; FULL (rebuild) also
// pads the tile, counts its atoms, computes its bounding box (pbc.c:262-303) and copies the tags
template <class real, int N, bool FULL>
__global__ void k_cp_dd_pack(int nsend, real xprd, real yprd, real zprd, const int* __restrict__ s_src, const int* __restrict__ s_code,
    const int* __restrict__ jnat, const real* __restrict__ cl_x, const int* __restrict__ cl_tag, real* __restrict__ tiles,
    int* __restrict__ nat_out, real* __restrict__ bb_out, int* __restrict__ tag_out)
{
    if (FULL) {
        const int k = blockIdx.x * blockDim.x + threadIdx.x;
        if (k >= nsend) return;
        const int cj = s_src[k], c = s_code[k], nat = jnat[cj];
        const real* s = cl_x + (size_t)cj * N * 3;
        real* d       = tiles + (size_t)k * N * 3;
        const real sh[3]  = { (real)((c & 3) - 1), (real)(((c >> 2) & 3) - 1), (real)(((c >> 4) & 3) - 1) };
        const real prd[3] = { xprd, yprd, zprd };
#pragma unroll
        for (int a = 0; a < 3; a++) {
            real lo = INFINITY, hi = -INFINITY;
            for (int q = 0; q < N; q++) {
                if (q < nat) {
                    const real v = fma_rn(sh[a], prd[a], s[a * N + q]);
                    d[a * N + q] = v;
                    if (lo > v) lo = v;
                    if (hi < v) hi = v;
                } else d[a * N + q] = CP_PAD;
            }
            bb_out[(size_t)k * 6 + 2 * a] = lo; bb_out[(size_t)k * 6 + 2 * a + 1] = hi;
        }
        nat_out[k] = nat;
        for (int q = 0; q < N; q++) tag_out[(size_t)k * N + q] = q < nat ? cl_tag[(size_t)cj * N + q] : -1;
    } else {
        const int t = blockIdx.x * blockDim.x + threadIdx.x;
        if (t >= nsend * N) return;
        const int k = t / N, q = t % N;
        const int cj = s_src[k];
        if (q >= jnat[cj]) return; // padding lanes keep the sentinel written at the rebuild
        const int c = s_code[k];
        const real* s = cl_x + (size_t)cj * N * 3;
        real* d       = tiles + (size_t)k * N * 3;
        d[q]         = fma_rn((real)((c & 3) - 1), xprd, s[q]);
        d[N + q]     = fma_rn((real)(((c >> 2) & 3) - 1), yprd, s[N + q]);
        d[2 * N + q] = fma_rn((real)(((c >> 4) & 3) - 1), zprd, s[2 * N + q]);
    }
}
// border tiles in ladder order: source tile and image direction of every ghost image (k_cp_ghost_fill without the writes
// into the own ghost range)
static __global__ void k_cp_dd_ghost_list(int ncj, const unsigned* __restrict__ mask, const int* __restrict__ offset, int* __restrict__ src,
    int* __restrict__ dir)
{
    const int cj = blockIdx.x * blockDim.x + threadIdx.x;
    if (cj >= ncj) return;
    unsigned m = mask[cj];
    int g      = offset[cj];
    while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        src[g] = cj;
        dir[g] = b;
        g++;
    }
}
static __global__ void k_gather_int(int n, const int* __restrict__ idx, const int* __restrict__ a, int* __restrict__ out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = a[idx[k]];
}
// leavers (send order) -> records {x + s ext, y + .., z + .., vx, vy, vz} and tags; dest[i] = ladder index of the shift
template <class real>
__global__ void k_cp_dd_pack_atoms(int n, const int* __restrict__ list, const int* __restrict__ dest, real ex, real ey, real ez,
    const real* __restrict__ x, const real* __restrict__ y, const real* __restrict__ z, const real* __restrict__ vx,
    const real* __restrict__ vy, const real* __restrict__ vz, const int* __restrict__ tag, real* __restrict__ rec, int* __restrict__ tag_out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int i = list[k], b = dest[i];
    real* o = rec + (size_t)k * 6;
    o[0] = fma_rn((real)c_img[b][0], ex, x[i]);
    o[1] = fma_rn((real)c_img[b][1], ey, y[i]);
    o[2] = fma_rn((real)c_img[b][2], ez, z[i]);
    o[3] = vx[i]; o[4] = vy[i]; o[5] = vz[i];
    tag_out[k] = tag[i];
}
template <class real>
__global__ void k_cp_dd_unpack_atoms(int n, int first, const real* __restrict__ rec, const int* __restrict__ tag_in, real* __restrict__ x,
    real* __restrict__ y, real* __restrict__ z, real* __restrict__ vx, real* __restrict__ vy, real* __restrict__ vz, int* __restrict__ tag,
    int* __restrict__ type)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const real* r = rec + (size_t)k * 6;
    const int i   = first + k;
    x[i] = r[0]; y[i] = r[1]; z[i] = r[2];
    vx[i] = r[3]; vy[i] = r[4]; vz[i] = r[5];
    tag[i]  = tag_in[k];
    type[i] = 0;
}
template <class real> __global__ void k_cp_fill(size_t n, real v, real* __restrict__ a)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}

template <class real, int N> struct CpGroup final : DDBase {
    typedef CpSim<real, N> Brick;
    Topo topo;
    int proc = 0, first = 0, nlb = 0, device = 0;
    mdb_params G;
    std::vector<Brick*> bricks;
    cudaStream_t stream = nullptr, own_stream = nullptr;
    NcclApi::comm_t comm = nullptr;
    long long gNatoms = 0, launches = 0;
    double comm_ms = 0;
    bool timing = false;
    cudaEvent_t ev[4] = { nullptr, nullptr, nullptr, nullptr };
    int* h_cnt    = nullptr; // pinned: [nbricks * 26] entries per (sending brick, direction) of the current exchange
    double* h_sum = nullptr; // pinned
    DBuf<int> d_cnt;
    DBuf<double> d_sum, d_rec;
    std::vector<int> cnt;
    std::vector<Xfer> plan_g, plan_m; // ghost tiles (replayed every step) / migration

    struct Side { // exchange state of one local brick
        DBuf<int> g_src, g_dir, s_src, s_code, s_nat, s_tag;
        DBuf<real> s_tiles, s_bb;
        int nsend = 0;
        DBuf<int> dest, leave, scan, stay_src, leavers, ldest, m_list, m_tag, m_tag_in, t_tag, t_type, sv_tag;
        DBuf<real> m_rec, m_in, t[6], sv[6];
        int nleave = 0, nstay = 0, narrive = 0, saved_n = 0;
        void release()
        {
            for (DBuf<int>* b : { &g_src, &g_dir, &s_src, &s_code, &s_nat, &s_tag, &dest, &leave, &scan, &stay_src, &leavers, &ldest,
                     &m_list, &m_tag, &m_tag_in, &t_tag, &t_type, &sv_tag })
                b->release();
            for (DBuf<real>* b : { &s_tiles, &s_bb, &m_rec, &m_in }) b->release();
            for (auto& b : t) b.release();
            for (auto& b : sv) b.release();
        }
    };
    std::vector<Side> side;
    std::vector<int> h_a, h_b; // host scratch for the partitions

    CpGroup(const mdb_params& g, const int grid[3], int nprocs, int proc_, const void* nccl_id, int dev) : proc(proc_), device(dev), G(g)
    {
        for (int a = 0; a < 3; a++) topo.g[a] = grid[a];
        topo.periodic[0] = topo.periodic[1] = topo.periodic[2] = 1; // clusterpair/pbc.c ignores pbc_x/y/z
        topo.nbricks = grid[0] * grid[1] * grid[2];
        topo.nprocs  = nprocs;
        if (topo.nbricks < 1 || nprocs < 1 || topo.nbricks % nprocs) throw Error("decomposition: #bricks must be a multiple of #processes");
        if (proc < 0 || proc >= nprocs) throw Error("decomposition: bad process index");
        if (g.from_input) throw Error("decomposition: only generated lattices (nx, ny, nz) are supported");
        if (g.nx % grid[0] || g.ny % grid[1] || g.nz % grid[2]) throw Error("decomposition: nx/ny/nz must be multiples of the brick grid");
        if (g.half_neigh) throw Error("clusterpair decomposition: full neighbor lists only (the owner computes both sides)");
        nlb   = topo.nbricks / nprocs;
        first = proc * nlb;
        MDB_CUDA(cudaSetDevice(device));
        MDB_CUDA(cudaStreamCreateWithFlags(&own_stream, cudaStreamNonBlocking));
        stream = own_stream;
        for (auto& e : ev) MDB_CUDA(cudaEventCreate(&e));
        MDB_CUDA(cudaMallocHost(&h_cnt, (size_t)topo.nbricks * 26 * sizeof(int)));
        MDB_CUDA(cudaMallocHost(&h_sum, 4096 * sizeof(double)));
        d_cnt.ensure((size_t)topo.nbricks * 26, false, stream);
        d_sum.ensure(4096, false, stream);
        cnt.assign((size_t)topo.nbricks * 26, 0);
        gNatoms = 4LL * g.nx * g.ny * g.nz;
        mdb_params L = g;
        L.nx = g.nx / grid[0]; L.ny = g.ny / grid[1]; L.nz = g.nz / grid[2];
        side.resize(nlb);
        for (int k = 0; k < nlb; k++) {
            Brick* b = new Brick(L, device);
            b->setStream(stream);
            b->brick_mode = true;
            bricks.push_back(b);
        }
        if (nprocs > 1) {
            if (!nccl_id) throw Error("decomposition over several processes needs the NCCL unique id of process 0");
            NcclApi& A = nccl_api();
            A.load();
            NcclApi::unique_id id;
            memcpy(&id, nccl_id, sizeof id);
            MDB_NCCL(A.CommInitRank(&comm, nprocs, id, proc));
        }
    }
    ~CpGroup() override
    {
        cudaSetDevice(device);
        cudaStreamSynchronize(stream);
        for (Brick* b : bricks) delete b;
        for (Side& s : side) s.release();
        d_cnt.release(); d_sum.release(); d_rec.release();
        for (auto& b : st) b.release();
        st_tag.release(); st_flag.release(); st_scan.release();
        if (comm) nccl_api().CommDestroy(comm);
        if (h_cnt) cudaFreeHost(h_cnt);
        if (h_sum) cudaFreeHost(h_sum);
        for (auto& e : ev)
            if (e) cudaEventDestroy(e);
        if (own_stream) cudaStreamDestroy(own_stream);
    }
    void setStream(cudaStream_t s) override
    {
        MDB_CUDA(cudaStreamSynchronize(stream));
        stream = s ? s : own_stream;
        for (Brick* b : bricks) b->setStream(stream);
    }
    void sync() override { MDB_CUDA(cudaStreamSynchronize(stream)); }
    void setOption(const char* name, double v) override
    {
        for (Brick* b : bricks) b->setOption(name, v);
    }
    void setTiming(bool on) override
    {
        timing = on;
        for (Brick* b : bricks) b->timing = on;
    }
    void stats(double* force_ms, long long* force_launches, double* neigh_ms, long long* neigh_launches, long long* nl, double* cms,
        bool reset) override
    {
        double f = 0, n = 0;
        long long fl = 0, nn = 0, l = launches;
        for (Brick* b : bricks) { f += b->force_ms; n += b->neigh_ms; fl += b->force_launches; nn += b->neigh_launches; l += b->launches; }
        if (force_ms) *force_ms = f;
        if (force_launches) *force_launches = fl;
        if (neigh_ms) *neigh_ms = n;
        if (neigh_launches) *neigh_launches = nn;
        if (nl) *nl = l;
        if (cms) *cms = comm_ms;
        if (reset) {
            for (Brick* b : bricks) { b->force_ms = b->neigh_ms = 0; b->force_launches = b->neigh_launches = b->launches = 0; }
            launches = 0;
            comm_ms  = 0;
        }
    }
    // n atoms lying in THIS process's bricks from host SoA buffers in the global frame, with their global tags (what the bench's
    // end-to-end leg hands over); each atom is dealt to the brick it lies in (k_dd_select / k_dd_take of the verletlist bricks)
    DBuf<real> st[6];
    DBuf<int> st_tag, st_flag, st_scan;
    void setAtoms(long long n, const int* tags, const void* ax, const void* ay, const void* az, const void* avx, const void* avy,
        const void* avz) override
    {
        if (n < 0 || n > 1500000000LL || !tags || !ax || !ay || !az) throw Error("mdb_dd_setAtoms: bad arguments");
        const void* h[6] = { ax, ay, az, avx, avy, avz };
        const bool has_v = avx && avy && avz;
        for (int k = 0; k < (has_v ? 6 : 3); k++) {
            st[k].ensure(n + 1, false, stream);
            MDB_CUDA(cudaMemcpyAsync(st[k].p, h[k], n * sizeof(real), cudaMemcpyHostToDevice, stream));
        }
        st_tag.ensure(n + 1, false, stream);
        st_flag.ensure(n + 1, false, stream);
        st_scan.ensure(n + 2, false, stream);
        MDB_CUDA(cudaMemcpyAsync(st_tag.p, tags, n * sizeof(int), cudaMemcpyHostToDevice, stream));
        long long taken = 0;
        for (int k = 0; k < nlb; k++) {
            Brick* b = bricks[k];
            int c[3];
            topo.coords(first + k, c);
            b->derive();
            MDB_LAUNCH(launches, k_dd_select<real>, grid_for(n, 256), 256, 0, stream, (int)n, b->xprd, b->yprd, b->zprd, topo.g[0], topo.g[1],
                topo.g[2], c[0], c[1], c[2], st[0].p, st[1].p, st[2].p, st_flag.p);
            b->scanner.exclusive(st_flag.p, st_scan.p, n, st_scan.p + n, stream);
            int cntb = 0;
            MDB_CUDA(cudaMemcpyAsync(&cntb, st_scan.p + n, sizeof(int), cudaMemcpyDeviceToHost, stream));
            MDB_CUDA(cudaStreamSynchronize(stream));
            b->Nlocal = cntb;
            b->Natoms = cntb;
            b->ensure_atoms((size_t)cntb + cntb / 4 + 1024);
            if (cntb)
                MDB_LAUNCH(launches, k_dd_take<real>, grid_for(n, 256), 256, 0, stream, (int)n, st_flag.p, st_scan.p, b->xprd, b->yprd, b->zprd,
                    topo.g[0], topo.g[1], topo.g[2], c[0], c[1], c[2], st[0].p, st[1].p, st[2].p, has_v ? st[3].p : (const real*)nullptr,
                    has_v ? st[4].p : (const real*)nullptr, has_v ? st[5].p : (const real*)nullptr, st_tag.p, b->x.p, b->y.p, b->z.p, b->vx.p,
                    b->vy.p, b->vz.p, b->type.p, b->tag.p);
            b->neigh_ready = b->lists_ready = false;
            taken += cntb;
        }
        MDB_CUDA(cudaStreamSynchronize(stream)); // host buffers may be reused by the caller
        if (taken != n) throw Error(fmt("mdb_dd_setAtoms: %lld of %lld atoms lie outside this process's bricks", n - taken, n));
    }
    void getNeighborTags(int*, int*, int*, int) override { throw Error("clusterpair decomposition: no per-atom neighbor rows (cluster-pair lists)"); }
    void setEam(int, double, int, double, double, double, const double*, const double*, const double*) override
    {
        throw Error("the clusterpair scheme has only the LJ kernels (force.c)");
    }

    // ------------------------------------------------------------------ plumbing
    void sum_over_procs(double* h, int n) // h[0..n) += the other processes' (h is pinned)
    {
        if (topo.nprocs == 1) return;
        MDB_CUDA(cudaMemcpyAsync(d_sum.p, h, n * sizeof(double), cudaMemcpyHostToDevice, stream));
        MDB_NCCL(nccl_api().AllReduce(d_sum.p, d_sum.p, n, NcclApi::Float64, NcclApi::Sum, comm, stream));
        MDB_CUDA(cudaMemcpyAsync(h, d_sum.p, n * sizeof(double), cudaMemcpyDeviceToHost, stream));
        MDB_CUDA(cudaStreamSynchronize(stream));
    }
    // cnt[] of the local bricks is filled in: every process learns every brick's counts
    void share_counts()
    {
        if (topo.nprocs > 1) {
            const size_t n = (size_t)nlb * 26;
            for (size_t k = 0; k < n; k++) h_cnt[(size_t)first * 26 + k] = cnt[(size_t)first * 26 + k];
            MDB_CUDA(cudaMemcpyAsync(d_cnt.p + (size_t)first * 26, h_cnt + (size_t)first * 26, n * sizeof(int), cudaMemcpyHostToDevice, stream));
            MDB_NCCL(nccl_api().AllGather(d_cnt.p + (size_t)first * 26, d_cnt.p, n, NcclApi::Int32, comm, stream));
            MDB_CUDA(cudaMemcpyAsync(h_cnt, d_cnt.p, (size_t)topo.nbricks * 26 * sizeof(int), cudaMemcpyDeviceToHost, stream));
            MDB_CUDA(cudaStreamSynchronize(stream));
            for (size_t k = 0; k < (size_t)topo.nbricks * 26; k++) cnt[k] = h_cnt[k];
        }
    }
    // stable partition of n entries with directions dir[] into brick S's send order (slots by (receiver, direction));
    // order[k] = entry at send position k; cnt[S*26 + d] = entries of direction d
    void partition(int S, int n, const int* dir, std::vector<int>& order)
    {
        int sdir[26], speer[26], slot_of[26], off[27];
        const int ns = topo.slots(S, true, sdir, speer);
        for (int d = 0; d < 26; d++) { slot_of[d] = -1; cnt[(size_t)S * 26 + d] = 0; }
        for (int k = 0; k < ns; k++) slot_of[sdir[k]] = k;
        for (int e = 0; e < n; e++) {
            if (dir[e] < 0 || dir[e] >= 26 || slot_of[dir[e]] < 0) throw Error("decomposition: image direction without a receiving brick");
            cnt[(size_t)S * 26 + dir[e]]++;
        }
        off[0] = 0;
        for (int k = 0; k < ns; k++) off[k + 1] = off[k] + cnt[(size_t)S * 26 + sdir[k]];
        order.resize(n);
        for (int e = 0; e < n; e++) order[off[slot_of[dir[e]]]++] = e;
    }
    int received(int R) const // entries brick R receives in the current exchange
    {
        int n = 0;
        for (int d = 0; d < 26; d++) {
            const int S = topo.sender(R, d);
            if (S >= 0) n += cnt[(size_t)S * 26 + d];
        }
        return n;
    }
    // one exchange: `bpe` bytes per entry, src(local brick) = base of its send buffer, dst(local brick) = base of what it receives
    template <class SrcF, class DstF> void exchange(const std::vector<Xfer>& plan, size_t bpe, SrcF src, DstF dst)
    {
        NcclApi& A = nccl_api();
        if (topo.nprocs > 1) MDB_NCCL(A.GroupStart());
        for (const Xfer& x : plan) {
            if (x.kind == 0)
                MDB_CUDA(cudaMemcpyAsync((char*)dst(x.dst - first) + (size_t)x.dst_start * bpe, (const char*)src(x.src - first) + (size_t)x.src_start * bpe,
                    (size_t)x.len * bpe, cudaMemcpyDeviceToDevice, stream));
            else if (x.kind == 1)
                MDB_NCCL(A.Send((const char*)src(x.src - first) + (size_t)x.src_start * bpe, (size_t)x.len * bpe, NcclApi::Int8, x.peer_proc, comm, stream));
            else
                MDB_NCCL(A.Recv((char*)dst(x.dst - first) + (size_t)x.dst_start * bpe, (size_t)x.len * bpe, NcclApi::Int8, x.peer_proc, comm, stream));
        }
        if (topo.nprocs > 1) MDB_NCCL(A.GroupEnd());
    }

    // ------------------------------------------------------------------ atoms
    long long createAtom() override // clusterpair/atom.c:49-180 restricted to each brick, global tags
    {
        const int gn[3] = { G.nx, G.ny, G.nz };
        for (int k = 0; k < nlb; k++) {
            Brick* b = bricks[k];
            int c[3];
            topo.coords(first + k, c);
            b->derive();
            const long long n = 4LL * b->P.nx * b->P.ny * b->P.nz;
            if (n > 1500000000LL) throw Error("createAtom: too many atoms for one brick");
            b->Natoms = n;
            b->Nlocal = (int)n;
            b->ensure_atoms((size_t)n + n / 4 + 1024);
            MDB_LAUNCH(launches, k_dd_create_atoms<real>, grid_for(n, 256), 256, 0, stream, gn[0], gn[1], gn[2], b->P.nx, b->P.ny, b->P.nz,
                c[0], c[1], c[2], b->lattice, b->x.p, b->y.p, b->z.p, b->vx.p, b->vy.p, b->vz.p, b->type.p, b->tag.p);
            b->neigh_ready = b->lists_ready = false;
        }
        return gNatoms;
    }
    void grow_atoms(Brick* b, size_t n) // keep the contents
    {
        if (n <= b->x.cap) return;
        const size_t m = n + n / 8 + 1024;
        for (DBuf<real>* a : { &b->x, &b->y, &b->z, &b->vx, &b->vy, &b->vz }) a->ensure(m, true, stream);
        b->tag.ensure(m, true, stream);
        b->type.ensure(m, true, stream);
    }
    void vel_sums(double* out) // out[0..4) = sum over all bricks of all processes
    {
        d_rec.ensure((size_t)4 * nlb, false, stream);
        for (int k = 0; k < nlb; k++) bricks[k]->vel_sums(d_rec.p + 4 * k);
        MDB_CUDA(cudaMemcpyAsync(h_sum + 8, d_rec.p, (size_t)4 * nlb * sizeof(double), cudaMemcpyDeviceToHost, stream));
        MDB_CUDA(cudaStreamSynchronize(stream));
        for (int c = 0; c < 4; c++) {
            h_sum[c] = 0;
            for (int k = 0; k < nlb; k++) h_sum[c] += h_sum[8 + 4 * k + c];
        }
        sum_over_procs(h_sum, 4);
        for (int c = 0; c < 4; c++) out[c] = h_sum[c];
    }
    void setupThermo() // thermo.c:30-53 with the GLOBAL atom count and volume
    {
        for (Brick* b : bricks) {
            b->dof_boltz    = (real)(gNatoms * 3 - 3);
            b->t_scale      = (real)1.0 / b->dof_boltz;
            b->p_scale      = (real)(1.0 / 3 / ((double)b->xprd * topo.g[0]) / ((double)b->yprd * topo.g[1]) / ((double)b->zprd * topo.g[2]));
            b->thermo_ready = true;
        }
    }
    void adjustThermo() // thermo.c:82-122 over all bricks
    {
        double s[4];
        vel_sums(s);
        const real vxtot = (real)s[0] / (real)gNatoms, vytot = (real)s[1] / (real)gNatoms, vztot = (real)s[2] / (real)gNatoms;
        for (Brick* b : bricks)
            MDB_LAUNCH(launches, k_vel_shift<real>, grid_for(b->Nlocal, 256), 256, 0, stream, b->Nlocal, b->vx.p, b->vy.p, b->vz.p, vxtot,
                vytot, vztot);
        vel_sums(s);
        real t = (real)s[3];
        t *= bricks[0]->t_scale;
        const real factor = (real)sqrt((double)(bricks[0]->temp / t));
        for (Brick* b : bricks)
            MDB_LAUNCH(launches, k_vel_scale<real>, grid_for(b->Nlocal, 256), 256, 0, stream, b->Nlocal, b->vx.p, b->vy.p, b->vz.p, factor);
    }
    void computeThermo(double* T, double* P) override
    {
        if (!bricks[0]->thermo_ready) setupThermo();
        double s[4];
        vel_sums(s);
        bricks[0]->thermo_from_sum(s[3], T, P);
    }

    // ------------------------------------------------------------------ migration (updateAtomsPbc across bricks, pbc.c:117-144)
    void migrate()
    {
        for (int k = 0; k < nlb; k++) {
            Brick* b = bricks[k];
            Side& s  = side[k];
            const int n = b->Nlocal;
            s.dest.ensure(n, false, stream);
            s.leave.ensure(n, false, stream);
            s.scan.ensure(n + 1, false, stream);
            ShiftMap sm;
            for (int q = 0; q < 27; q++) sm.b[q] = -1;
            for (int d = 0; d < 26; d++)
                if (topo.receiver(first + k, d) >= 0) sm.b[(DD_IMG[d][0] + 1) + 3 * (DD_IMG[d][1] + 1) + 9 * (DD_IMG[d][2] + 1)] = (signed char)d;
            MDB_LAUNCH(launches, k_dd_dest<real>, grid_for(n, 256), 256, 0, stream, n, b->xprd, b->yprd, b->zprd, topo.g[0], topo.g[1],
                topo.g[2], 1, 1, 1, sm, b->x.p, b->y.p, b->z.p, s.dest.p, s.leave.p);
            b->scanner.exclusive(s.leave.p, s.scan.p, n, b->d_flags.p + 8, stream);
            MDB_CUDA(cudaMemcpyAsync(b->h_flags + 8, b->d_flags.p + 8, sizeof(int), cudaMemcpyDeviceToHost, stream));
        }
        MDB_CUDA(cudaStreamSynchronize(stream));
        for (int k = 0; k < nlb; k++) {
            Brick* b = bricks[k];
            Side& s  = side[k];
            const int n = b->Nlocal;
            s.nleave = b->h_flags[8];
            s.nstay  = n - s.nleave;
            s.stay_src.ensure(n, false, stream);
            s.leavers.ensure(s.nleave + 1, false, stream);
            s.ldest.ensure(s.nleave + 1, false, stream);
            MDB_LAUNCH(launches, k_dd_split, grid_for(n, 256), 256, 0, stream, n, s.leave.p, s.scan.p, s.stay_src.p, s.leavers.p);
            h_a.resize(s.nleave);
            if (s.nleave) {
                MDB_LAUNCH(launches, k_gather_int, grid_for(s.nleave, 256), 256, 0, stream, s.nleave, s.leavers.p, s.dest.p, s.ldest.p);
                MDB_CUDA(cudaMemcpyAsync(h_a.data(), s.ldest.p, s.nleave * sizeof(int), cudaMemcpyDeviceToHost, stream));
                MDB_CUDA(cudaStreamSynchronize(stream));
            }
            partition(first + k, s.nleave, h_a.data(), h_b); // h_b[k] = position in the ascending leaver list
            s.m_list.ensure(s.nleave + 1, false, stream);
            s.m_rec.ensure((size_t)6 * s.nleave + 1, false, stream);
            s.m_tag.ensure(s.nleave + 1, false, stream);
            if (s.nleave) {
                // send order as atom indices: m_list[k] = leavers[h_b[k]]
                s.scan.ensure(s.nleave + 1, false, stream); // reuse as the uploaded permutation
                MDB_CUDA(cudaMemcpyAsync(s.scan.p, h_b.data(), s.nleave * sizeof(int), cudaMemcpyHostToDevice, stream));
                MDB_LAUNCH(launches, k_gather_int, grid_for(s.nleave, 256), 256, 0, stream, s.nleave, s.scan.p, s.leavers.p, s.m_list.p);
                MDB_LAUNCH(launches, k_cp_dd_pack_atoms<real>, grid_for(s.nleave, 256), 256, 0, stream, s.nleave, s.m_list.p, s.dest.p, b->xprd,
                    b->yprd, b->zprd, b->x.p, b->y.p, b->z.p, b->vx.p, b->vy.p, b->vz.p, b->tag.p, s.m_rec.p, s.m_tag.p);
                MDB_CUDA(cudaStreamSynchronize(stream)); // h_b is reused by the next brick
            }
        }
        share_counts();
        dd_schedule(topo, proc, cnt.data(), plan_m);
        for (int k = 0; k < nlb; k++) {
            Brick* b = bricks[k];
            Side& s  = side[k];
            s.narrive = received(first + k);
            if (s.nleave) { // stayers to the front, in their old order
                const size_t cap = b->x.cap;
                for (auto& q : s.t) q.ensure(cap, false, stream);
                s.t_tag.ensure(cap, false, stream);
                s.t_type.ensure(cap, false, stream);
                if (s.nstay)
                    MDB_LAUNCH(launches, k_dd_gather_stay<real>, grid_for(s.nstay, 256), 256, 0, stream, s.nstay, s.stay_src.p, b->x.p, b->y.p,
                        b->z.p, b->vx.p, b->vy.p, b->vz.p, b->type.p, b->tag.p, s.t[0].p, s.t[1].p, s.t[2].p, s.t[3].p, s.t[4].p, s.t[5].p,
                        s.t_type.p, s.t_tag.p);
                std::swap(b->x, s.t[0]); std::swap(b->y, s.t[1]); std::swap(b->z, s.t[2]);
                std::swap(b->vx, s.t[3]); std::swap(b->vy, s.t[4]); std::swap(b->vz, s.t[5]);
                std::swap(b->type, s.t_type); std::swap(b->tag, s.t_tag);
            }
            grow_atoms(b, (size_t)s.nstay + s.narrive);
            s.m_in.ensure((size_t)6 * s.narrive + 1, false, stream);
            s.m_tag_in.ensure(s.narrive + 1, false, stream);
        }
        if (timing) MDB_CUDA(cudaEventRecord(ev[0], stream));
        exchange(plan_m, 6 * sizeof(real), [&](int k) { return (void*)side[k].m_rec.p; }, [&](int k) { return (void*)side[k].m_in.p; });
        exchange(plan_m, sizeof(int), [&](int k) { return (void*)side[k].m_tag.p; }, [&](int k) { return (void*)side[k].m_tag_in.p; });
        comm_time();
        for (int k = 0; k < nlb; k++) {
            Brick* b = bricks[k];
            Side& s  = side[k];
            if (s.narrive)
                MDB_LAUNCH(launches, k_cp_dd_unpack_atoms<real>, grid_for(s.narrive, 256), 256, 0, stream, s.narrive, s.nstay, s.m_in.p,
                    s.m_tag_in.p, b->x.p, b->y.p, b->z.p, b->vx.p, b->vy.p, b->vz.p, b->tag.p, b->type.p);
            b->Nlocal = s.nstay + s.narrive;
            b->Natoms = b->Nlocal;
        }
    }
    void comm_time()
    {
        if (!timing) return;
        float ms = 0;
        MDB_CUDA(cudaEventRecord(ev[1], stream));
        MDB_CUDA(cudaEventSynchronize(ev[1]));
        MDB_CUDA(cudaEventElapsedTime(&ms, ev[0], ev[1]));
        comm_ms += ms;
    }

    // ------------------------------------------------------------------ ghost clusters (setupPbc across bricks, pbc.c:183-323)
    void setupPbc()
    {
        for (int k = 0; k < nlb; k++) {
            Brick* b = bricks[k];
            b->gmask.ensure(b->ncj, false, stream);
            b->gcnt.ensure(b->ncj, false, stream);
            b->goff.ensure(b->ncj, false, stream);
            MDB_LAUNCH(launches, k_cp_ghost_count<real>, grid_for(b->ncj, 256), 256, 0, stream, b->ncj, b->pbc_geom(), b->jnat.p, b->jbb.p,
                b->gmask.p, b->gcnt.p);
            b->scanner.exclusive(b->gcnt.p, b->goff.p, b->ncj, b->d_flags.p + 8, stream);
            MDB_CUDA(cudaMemcpyAsync(b->h_flags + 8, b->d_flags.p + 8, sizeof(int), cudaMemcpyDeviceToHost, stream));
        }
        MDB_CUDA(cudaStreamSynchronize(stream));
        int code_of[26];
        for (int d = 0; d < 26; d++) code_of[d] = (DD_IMG[d][0] + 1) | ((DD_IMG[d][1] + 1) << 2) | ((DD_IMG[d][2] + 1) << 4);
        std::vector<int> h_src, h_dir, order, h_ssrc, h_scode;
        for (int k = 0; k < nlb; k++) {
            Brick* b = bricks[k];
            Side& s  = side[k];
            const int n = b->h_flags[8];
            s.nsend     = n;
            s.g_src.ensure(n + 1, false, stream);
            s.g_dir.ensure(n + 1, false, stream);
            MDB_LAUNCH(launches, k_cp_dd_ghost_list, grid_for(b->ncj, 256), 256, 0, stream, b->ncj, b->gmask.p, b->goff.p, s.g_src.p, s.g_dir.p);
            h_src.resize(n); h_dir.resize(n);
            if (n) {
                MDB_CUDA(cudaMemcpyAsync(h_src.data(), s.g_src.p, n * sizeof(int), cudaMemcpyDeviceToHost, stream));
                MDB_CUDA(cudaMemcpyAsync(h_dir.data(), s.g_dir.p, n * sizeof(int), cudaMemcpyDeviceToHost, stream));
                MDB_CUDA(cudaStreamSynchronize(stream));
            }
            partition(first + k, n, h_dir.data(), order);
            h_ssrc.resize(n); h_scode.resize(n);
            for (int q = 0; q < n; q++) { h_ssrc[q] = h_src[order[q]]; h_scode[q] = code_of[h_dir[order[q]]]; }
            s.s_src.ensure(n + 1, false, stream);
            s.s_code.ensure(n + 1, false, stream);
            s.s_tiles.ensure((size_t)n * 3 * N + 1, false, stream);
            s.s_bb.ensure((size_t)n * 6 + 1, false, stream);
            s.s_nat.ensure(n + 1, false, stream);
            s.s_tag.ensure((size_t)n * N + 1, false, stream);
            if (n) {
                MDB_CUDA(cudaMemcpyAsync(s.s_src.p, h_ssrc.data(), n * sizeof(int), cudaMemcpyHostToDevice, stream));
                MDB_CUDA(cudaMemcpyAsync(s.s_code.p, h_scode.data(), n * sizeof(int), cudaMemcpyHostToDevice, stream));
                MDB_LAUNCH(launches, (k_cp_dd_pack<real, N, true>), grid_for(n, 128), 128, 0, stream, n, b->xprd, b->yprd, b->zprd, s.s_src.p,
                    s.s_code.p, b->jnat.p, b->cl_x.p, b->cl_tag.p, s.s_tiles.p, s.s_nat.p, s.s_bb.p, s.s_tag.p);
                MDB_CUDA(cudaStreamSynchronize(stream)); // the host vectors are reused by the next brick
            }
        }
        share_counts();
        dd_schedule(topo, proc, cnt.data(), plan_g);
        std::vector<int> h_rcode;
        for (int k = 0; k < nlb; k++) { // the receiving side: ghost range, image code of every ghost (binClusters needs it)
            Brick* b = bricks[k];
            const int R = first + k;
            int rdir[26], rpeer[26];
            const int ns = topo.slots(R, false, rdir, rpeer);
            h_rcode.clear();
            for (int q = 0; q < ns; q++) h_rcode.insert(h_rcode.end(), cnt[(size_t)rpeer[q] * 26 + rdir[q]], code_of[rdir[q]]);
            b->nghost   = (int)h_rcode.size();
            b->dummy_cj = b->ncj + b->nghost;
            b->ensure_tiles((size_t)b->ncj + b->nghost + 1, true);
            b->border_map.ensure(b->nghost + 1, false, stream);
            b->code.ensure(b->nghost + 1, false, stream);
            if (b->nghost) {
                MDB_CUDA(cudaMemcpyAsync(b->code.p, h_rcode.data(), b->nghost * sizeof(int), cudaMemcpyHostToDevice, stream));
                MDB_CUDA(cudaStreamSynchronize(stream));
            }
            MDB_LAUNCH(launches, k_cp_fill<real>, 1, 3 * N, 0, stream, (size_t)3 * N, (real)CP_PAD, b->cl_x.p + (size_t)b->dummy_cj * 3 * N); // pbc.c:304-311
        }
        if (timing) MDB_CUDA(cudaEventRecord(ev[0], stream));
        exchange(plan_g, 3 * N * sizeof(real), [&](int k) { return (void*)side[k].s_tiles.p; },
            [&](int k) { return (void*)(bricks[k]->cl_x.p + (size_t)bricks[k]->ncj * 3 * N); });
        exchange(plan_g, sizeof(int), [&](int k) { return (void*)side[k].s_nat.p; }, [&](int k) { return (void*)(bricks[k]->jnat.p + bricks[k]->ncj); });
        exchange(plan_g, 6 * sizeof(real), [&](int k) { return (void*)side[k].s_bb.p; },
            [&](int k) { return (void*)(bricks[k]->jbb.p + (size_t)bricks[k]->ncj * 6); });
        exchange(plan_g, N * sizeof(int), [&](int k) { return (void*)side[k].s_tag.p; },
            [&](int k) { return (void*)(bricks[k]->cl_tag.p + (size_t)bricks[k]->ncj * N); });
        comm_time();
    }
    void forward() // updatePbc across bricks (pbc.c:45-114): the same images, new positions
    {
        for (int k = 0; k < nlb; k++) {
            Brick* b = bricks[k];
            Side& s  = side[k];
            if (s.nsend)
                MDB_LAUNCH(launches, (k_cp_dd_pack<real, N, false>), grid_for((size_t)s.nsend * N, 256), 256, 0, stream, s.nsend, b->xprd, b->yprd,
                    b->zprd, s.s_src.p, s.s_code.p, b->jnat.p, b->cl_x.p, b->cl_tag.p, s.s_tiles.p, (int*)nullptr, (real*)nullptr, (int*)nullptr);
        }
        if (timing) MDB_CUDA(cudaEventRecord(ev[0], stream));
        exchange(plan_g, 3 * N * sizeof(real), [&](int k) { return (void*)side[k].s_tiles.p; },
            [&](int k) { return (void*)(bricks[k]->cl_x.p + (size_t)bricks[k]->ncj * 3 * N); });
        comm_time();
    }

    // ------------------------------------------------------------------ driver flow (clusterpair/main.c)
    void setup(bool adjust) override // main.c:40-76 after the atoms exist
    {
        for (Brick* b : bricks) { b->derive(); b->setupNeighbor(); }
        setupThermo();
        if (adjust) adjustThermo();
        for (Brick* b : bricks) { b->buildClusters(); b->defineJClusters(); }
        setupPbc();
        for (Brick* b : bricks) { b->binClusters(); b->buildNeighbor(); }
    }
    void reneighbour() override // main.c:78-93
    {
        NvtxRange nvtx_range_("reneighbour");
        for (Brick* b : bricks) b->updateSingleAtoms();
        migrate();
        for (Brick* b : bricks) { b->buildClusters(); b->defineJClusters(); }
        setupPbc();
        for (Brick* b : bricks) { b->binClusters(); b->buildNeighbor(); }
    }
    void run(int nsteps, double* thermo_out, int max_records, int* nrecords, double* timers) override // CpSim::run over all bricks
    {
        Brick* b0 = bricks[0];
        if (!b0->thermo_ready) setupThermo();
        const int nstat  = G.nstat > 0 ? G.nstat : nsteps + 1;
        const int every  = G.reneigh_every > 0 ? G.reneigh_every : nsteps + 1;
        const int pevery = b0->prune_every > 0 ? b0->prune_every : nsteps + 1;
        const int maxrec = nsteps / nstat + 3;
        d_rec.ensure((size_t)4 * nlb * (maxrec + 1), false, stream);
        std::vector<int> rec_step;
        auto record = [&](int step) {
            for (int k = 0; k < nlb; k++) bricks[k]->vel_sums(d_rec.p + 4 * (rec_step.size() * nlb + k));
            rec_step.push_back(step);
        };
        double f0 = 0, n0 = 0;
        for (Brick* b : bricks) { f0 += b->force_ms; n0 += b->neigh_ms; }
        record(0);
        for (Brick* b : bricks) b->launch_force();
        MDB_CUDA(cudaEventRecord(ev[2], stream));
        bool initial_done = false;
        const bool fuse   = b0->can_fuse_force();
        bool second_ok    = false;
        for (int n = 0; n < nsteps; n++) {
            if (!initial_done)
                for (Brick* b : bricks) b->initialIntegrate();
            if ((n + 1) % every) {
                if (!((n + 1) % pevery))
                    for (Brick* b : bricks) b->pruneNeighbor();
                forward();
            } else {
                reneighbour();
                second_ok = false;
            }
            const bool rec  = !((n + 1) % nstat) && (n + 1) < nsteps;
            const bool last = n + 1 == nsteps;
            if (fuse && !rec && !last) {
                if (!second_ok)
                    for (Brick* b : bricks) b->sync_second_array();
                second_ok = true;
                for (Brick* b : bricks) b->launch_force(true);
                initial_done = true;
                continue;
            }
            for (Brick* b : bricks) b->launch_force();
            if (rec || last) {
                for (Brick* b : bricks) b->finalIntegrate();
                initial_done = false;
                if (rec) record(n + 1);
            } else {
                for (Brick* b : bricks) b->integrate(2);
                initial_done = true;
            }
        }
        MDB_CUDA(cudaEventRecord(ev[3], stream));
        for (Brick* b : bricks) b->updateSingleAtoms(); // main.c:300
        record(nsteps);
        const size_t nr_all = rec_step.size();
        std::vector<double> h(4 * nlb * nr_all);
        MDB_CUDA(cudaMemcpyAsync(h.data(), d_rec.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, stream));
        MDB_CUDA(cudaStreamSynchronize(stream));
        if (nr_all > 1000) throw Error("run: too many thermo records");
        for (size_t r = 0; r < nr_all; r++) {
            h_sum[r] = 0;
            for (int k = 0; k < nlb; k++) h_sum[r] += h[4 * (r * nlb + k) + 3];
        }
        sum_over_procs(h_sum, (int)nr_all);
        float ms = 0;
        MDB_CUDA(cudaEventElapsedTime(&ms, ev[2], ev[3]));
        int nr = 0;
        for (size_t r = 0; r < nr_all; r++)
            if (thermo_out && nr < max_records) {
                thermo_out[3 * nr] = rec_step[r];
                b0->thermo_from_sum(h_sum[r], &thermo_out[3 * nr + 1], &thermo_out[3 * nr + 2]);
                nr++;
            }
        if (nrecords) *nrecords = nr;
        if (timers) {
            double f1 = 0, n1 = 0;
            for (Brick* b : bricks) { f1 += b->force_ms; n1 += b->neigh_ms; }
            timers[0] = ms * 1e-3;
            timers[1] = (f1 - f0) * 1e-3;
            timers[2] = (n1 - n0) * 1e-3;
        }
    }

    // ------------------------------------------------------------------ accessors
    void getCounts(long long* v) override // Natoms (global), local atoms, ghost clusters, maxneighs, bricks of this process
    {
        long long nl = 0, ng = 0;
        int mn = 0;
        for (Brick* b : bricks) { nl += b->Nlocal; ng += b->nghost; mn = std::max(mn, b->maxneighs); }
        v[0] = gNatoms; v[1] = nl; v[2] = ng; v[3] = mn; v[4] = nlb;
    }
    void getAtoms(int which, int* tags, void* ax, void* ay, void* az) override // atom arrays (as of the last updateSingleAtoms), global frame
    {
        if (which != 'x' && which != 'v') throw Error("mdb_dd_getAtoms (clusterpair): which must be 'x' or 'v'");
        size_t off = 0;
        for (int k = 0; k < nlb; k++) {
            Brick* b = bricks[k];
            const size_t n = b->Nlocal;
            DBuf<real>* src[3] = { which == 'x' ? &b->x : &b->vx, which == 'x' ? &b->y : &b->vy, which == 'x' ? &b->z : &b->vz };
            void* dst[3] = { ax, ay, az };
            for (int c = 0; c < 3; c++)
                MDB_CUDA(cudaMemcpyAsync((real*)dst[c] + off, src[c]->p, n * sizeof(real), cudaMemcpyDeviceToHost, stream));
            if (tags) MDB_CUDA(cudaMemcpyAsync(tags + off, b->tag.p, n * sizeof(int), cudaMemcpyDeviceToHost, stream));
            MDB_CUDA(cudaStreamSynchronize(stream));
            if (which == 'x') {
                int c[3];
                topo.coords(first + k, c);
                const real o[3] = { b->xprd * c[0], b->yprd * c[1], b->zprd * c[2] };
                for (int a = 0; a < 3; a++)
                    for (size_t i = 0; i < n; i++) ((real*)dst[a])[off + i] += o[a];
            }
            off += n;
        }
    }
    void saveState() override
    {
        for (int k = 0; k < nlb; k++) {
            Brick* b = bricks[k];
            Side& s  = side[k];
            const size_t n = b->Nlocal;
            DBuf<real>* src[] = { &b->x, &b->y, &b->z, &b->vx, &b->vy, &b->vz };
            for (int q = 0; q < 6; q++) {
                s.sv[q].ensure(n, false, stream);
                MDB_CUDA(cudaMemcpyAsync(s.sv[q].p, src[q]->p, n * sizeof(real), cudaMemcpyDeviceToDevice, stream));
            }
            s.sv_tag.ensure(n, false, stream);
            MDB_CUDA(cudaMemcpyAsync(s.sv_tag.p, b->tag.p, n * sizeof(int), cudaMemcpyDeviceToDevice, stream));
            s.saved_n = b->Nlocal;
        }
    }
    void restoreState() override
    {
        for (int k = 0; k < nlb; k++) {
            Brick* b = bricks[k];
            Side& s  = side[k];
            if (!s.saved_n) throw Error("restoreState: nothing saved");
            const size_t n = s.saved_n;
            grow_atoms(b, n);
            DBuf<real>* dst[] = { &b->x, &b->y, &b->z, &b->vx, &b->vy, &b->vz };
            for (int q = 0; q < 6; q++) MDB_CUDA(cudaMemcpyAsync(dst[q]->p, s.sv[q].p, n * sizeof(real), cudaMemcpyDeviceToDevice, stream));
            MDB_CUDA(cudaMemcpyAsync(b->tag.p, s.sv_tag.p, n * sizeof(int), cudaMemcpyDeviceToDevice, stream));
            MDB_CUDA(cudaMemsetAsync(b->type.p, 0, n * sizeof(int), stream));
            b->Nlocal = b->Natoms = s.saved_n;
            b->lists_ready = false;
        }
    }
};

DDBase* make_cp_dd(const mdb_params& p, int cluster_n, const int grid[3], int nprocs, int proc, const void* nccl_id, int device)
{
    int ndev      = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) throw Error("mdb_dd_create_cp: no CUDA device (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) throw Error(fmt("mdb_dd_create_cp: device %d out of range (%d devices)", device, ndev));
    if (p.force_field != MDB_FF_LJ) throw Error("mdb_dd_create_cp: the clusterpair scheme has only the LJ kernels (force.c)");
    if (cluster_n != 4 && cluster_n != 8) throw Error("mdb_dd_create_cp: cluster_n must be 4 or 8 (M = 4)");
    if (p.precision == MDB_DP)
        return cluster_n == 4 ? (DDBase*)new CpGroup<double, 4>(p, grid, nprocs, proc, nccl_id, device)
                              : new CpGroup<double, 8>(p, grid, nprocs, proc, nccl_id, device);
    if (p.precision == MDB_SP)
        return cluster_n == 4 ? (DDBase*)new CpGroup<float, 4>(p, grid, nprocs, proc, nccl_id, device)
                              : new CpGroup<float, 8>(p, grid, nprocs, proc, nccl_id, device);
    throw Error("mdb_dd_create_cp: precision must be MDB_SP or MDB_DP");
}
