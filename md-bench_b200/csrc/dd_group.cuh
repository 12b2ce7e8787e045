// dd_group.cuh -- DomainGroup<real>: the bricks of a spatially decomposed box that live in THIS process,
// driven in lockstep on one stream, plus the transport between bricks:
//   * bricks of the same process: device-to-device copies (also how a single GPU runs a decomposed box
//     for the parity tests);
//   * bricks of other processes (one process per GPU): NCCL send/recv over NVLink, grouped per phase.
// Every process derives the same transfer schedule from the allgathered slot counts, so both ends of
// a transfer enumerate it at the same position (NCCL matches sends and receives of a pair in order).
// The reference has no decomposition; the per-brick operators are the single-domain ones (sim_impl.cu)
// and only the setupPbc/updatePbc/updateAtomsPbc trio (pbc.c) is generalised (dd_kernels.cuh).
// (textually included at the end of sim_impl.cu, inside namespace mdb)

template <class real> struct DomainGroup final : DDBase {
    typedef Sim<real> Brick;
    Topo topo;
    int proc = 0, first = 0, nlocal_bricks = 0, device = 0;
    mdb_params G; // global parameters
    std::vector<Brick*> bricks;
    cudaStream_t stream = nullptr, own_stream = nullptr;
    NcclApi::comm_t comm = nullptr;
    long long gNatoms = 0, launches = 0;
    bool timing = false;
    double comm_ms = 0;

    DBuf<int> d_off_all;
    DBuf<double> d_sum;
    int* h_off_all = nullptr; // pinned, nbricks*32
    double* h_sum  = nullptr; // pinned
    std::vector<int> cnt;     // [nbricks*26] entries per (sending brick, direction) of the last exchange
    cudaEvent_t ev[4] = { nullptr, nullptr, nullptr, nullptr };

    // ---- halo by peer stores (dd_kernels.cuh, k_dd_push): one brick per process, several processes
    bool push_wanted = true, push_ready = false; // option "halo_push" (0: NCCL send/recv on every step)
    int epoch = 0;
    int* sync_flags = nullptr;            // own flag words: [0,32) "ready" from p, [32,64) "done" from p, [96] error
    PeerFlags peer_flags {};              // every process's flag buffer, mapped
    // peers' position allocations as mapped here.  A brick's x, y, z alternate between two allocations (the spatial sort
    // and the migration swap buffers), so every mapping ever opened is kept: after the first rebuilds nothing is opened
    // any more (cudaIpcOpenMemHandle of an 84 MB array costs milliseconds)
    struct IpcSlot {
        std::vector<std::pair<cudaIpcMemHandle_t, void*>> open;
    };
    std::vector<IpcSlot> peer_xyz;
    // allocations of x, y, z (and of the sort buffers they are swapped with) that this process has outgrown while a peer
    // may still have them mapped; freed in prepare_push() after every process has closed its stale mappings
    std::vector<void*> retired;
    void free_retired()
    {
        for (void* q : retired) cudaFree(q);
        retired.clear();
    }
    DBuf<unsigned char> ipc_stage;        // device staging for the handle allgather
    PushTable push_tab {};
    unsigned push_to = 0, push_from = 0;  // processes this one pushes to / is pushed by
    std::vector<Xfer> pos_plan;

    enum OpKind { COPY, SEND, RECV };
    struct Op {
        OpKind kind;
        const void* src;
        void* dst;
        size_t bytes;
        int peer;
    };
    std::vector<Op> pos_ops, fp_ops;

    DomainGroup(const mdb_params& g, const int grid[3], int nprocs, int proc_, const void* nccl_id, int dev)
        : proc(proc_), device(dev), G(g)
    {
        for (int a = 0; a < 3; a++) topo.g[a] = grid[a];
        topo.periodic[0] = g.pbc_x; topo.periodic[1] = g.pbc_y; topo.periodic[2] = g.pbc_z;
        topo.nbricks = grid[0] * grid[1] * grid[2];
        topo.nprocs  = nprocs;
        if (topo.nbricks < 1 || nprocs < 1 || topo.nbricks % nprocs) throw Error("decomposition: #bricks must be a multiple of #processes");
        if (proc < 0 || proc >= nprocs) throw Error("decomposition: bad process index");
        if (g.from_input) throw Error("decomposition: only generated lattices (nx, ny, nz) are supported");
        if (g.nx % grid[0] || g.ny % grid[1] || g.nz % grid[2]) throw Error("decomposition: nx/ny/nz must be multiples of the brick grid");
        nlocal_bricks = topo.nbricks / nprocs;
        first         = proc * nlocal_bricks;
        MDB_CUDA(cudaSetDevice(device));
        MDB_CUDA(cudaStreamCreateWithFlags(&own_stream, cudaStreamNonBlocking));
        stream = own_stream;
        for (auto& e : ev) MDB_CUDA(cudaEventCreate(&e));
        MDB_CUDA(cudaMallocHost(&h_off_all, (size_t)topo.nbricks * 32 * sizeof(int)));
        MDB_CUDA(cudaMallocHost(&h_sum, 4096 * sizeof(double)));
        d_off_all.ensure((size_t)topo.nbricks * 32, false, stream);
        d_sum.ensure(4096, false, stream);
        cnt.assign((size_t)topo.nbricks * 26, 0);
        gNatoms = 4LL * g.nx * g.ny * g.nz;
        mdb_params L = g;
        L.nx = g.nx / grid[0]; L.ny = g.ny / grid[1]; L.nz = g.nz / grid[2];
        const int gn[3] = { g.nx, g.ny, g.nz };
        for (int k = 0; k < nlocal_bricks; k++) {
            Brick* b = new Brick(L, device);
            b->brick_init(&topo, first + k, gn, gNatoms);
            b->d_off = d_off_all.p + (size_t)(first + k) * 32;
            b->setStream(stream);
            // several bricks: re-sort the atoms by bin at every rebuild.  Unlike the single domain (where the generator's
            // order stays the better one) a brick's force kernel slows down by 20 % within 100 steps without it, and its
            // list build by 27 % (profiles/r1_ab3.txt, dd_case)
            // in blocks of 2^3 bins: 128 consecutive atoms fill a compact box (force 1.635 -> 1.580 ms per launch at N = 2,
            // profiles/r2_s2_n2.txt)
            if (topo.nbricks > 1) { b->sort_enabled = true; b->sort_block = 2; }
            bricks.push_back(b);
        }
        if (nprocs > 1) {
            if (!nccl_id) throw Error("decomposition over several processes needs the NCCL unique id of process 0");
            NcclApi& N = nccl_api();
            N.load();
            NcclApi::unique_id id;
            memcpy(&id, nccl_id, sizeof id);
            MDB_NCCL(N.CommInitRank(&comm, nprocs, id, proc));
            if (nlocal_bricks == 1 && nprocs <= 32) init_push();
        }
    }
    // allgather of `bytes` bytes per process through NCCL (host buffers in, host buffers out)
    void allgather_bytes(const void* mine, void* all, size_t bytes)
    {
        ipc_stage.ensure(bytes * topo.nprocs, false, stream);
        MDB_CUDA(cudaMemcpyAsync(ipc_stage.p + bytes * proc, mine, bytes, cudaMemcpyHostToDevice, stream));
        MDB_NCCL(nccl_api().AllGather(ipc_stage.p + bytes * proc, ipc_stage.p, bytes, NcclApi::Int8, comm, stream));
        MDB_CUDA(cudaMemcpyAsync(all, ipc_stage.p, bytes * topo.nprocs, cudaMemcpyDeviceToHost, stream));
        MDB_CUDA(cudaStreamSynchronize(stream));
    }
    bool all_agree(bool ok) // every process must take the same path
    {
        double v = ok ? 0.0 : 1.0;
        h_sum[0] = v;
        sum_over_procs(h_sum, 1);
        return h_sum[0] == 0.0;
    }
    void init_push() // flag words of every process, mapped once
    {
        bool ok = true;
        cudaIpcMemHandle_t mine;
        memset(&mine, 0, sizeof mine);
        if (cudaMalloc(&sync_flags, 128 * sizeof(int)) != cudaSuccess) { sync_flags = nullptr; ok = false; }
        if (ok) {
            MDB_CUDA(cudaMemsetAsync(sync_flags, 0, 128 * sizeof(int), stream));
            MDB_CUDA(cudaStreamSynchronize(stream));
            ok = cudaIpcGetMemHandle(&mine, sync_flags) == cudaSuccess;
        }
        std::vector<cudaIpcMemHandle_t> all(topo.nprocs);
        allgather_bytes(&mine, all.data(), sizeof mine);
        ok = all_agree(ok);
        for (int p = 0; ok && p < topo.nprocs; p++) {
            if (p == proc) { peer_flags.p[p] = sync_flags; continue; }
            void* q = nullptr;
            if (cudaIpcOpenMemHandle(&q, all[p], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; cudaGetLastError(); break; }
            peer_flags.p[p] = (int*)q;
        }
        push_ready = all_agree(ok);
        peer_xyz.assign(topo.nprocs, IpcSlot {});
        if (push_ready) { // from now on the position arrays are exported: never free them under a peer's mapping
            Brick* b = bricks[0];
            for (DBuf<real>* a : { &b->x, &b->y, &b->z, &b->x2, &b->y2, &b->z2 }) a->retire = &retired;
        }
    }
    // per rebuild: where every process keeps x, y, z now (allocation handles + first ghost entry), then the segment table
    void prepare_push()
    {
        if (!push_ready) return;
        Brick* b = bricks[0];
        struct Pub {
            cudaIpcMemHandle_t h[6]; // x, y, z and the sort buffers they alternate with (zero = not allocated)
            int nlocal, ok;
        } mine;
        memset(&mine, 0, sizeof mine);
        mine.nlocal = b->Nlocal;
        mine.ok     = 1;
        real* arr[6] = { b->x.p, b->y.p, b->z.p, b->x2.p, b->y2.p, b->z2.p };
        for (int k = 0; k < 6; k++)
            if (arr[k] && cudaIpcGetMemHandle(&mine.h[k], arr[k]) != cudaSuccess) {
                if (k < 3) mine.ok = 0;
                memset(&mine.h[k], 0, sizeof mine.h[k]);
                cudaGetLastError();
            }
        std::vector<Pub> all(topo.nprocs);
        allgather_bytes(&mine, all.data(), sizeof mine);
        bool ok = true;
        for (int p = 0; p < topo.nprocs; p++) ok = ok && all[p].ok;
        // close every mapping of an allocation its owner no longer publishes (it has been outgrown and is waiting in the
        // owner's `retired` list); after the barrier below nobody maps a retired allocation any more and the owners free them
        for (int p = 0; p < topo.nprocs; p++) {
            auto& open = peer_xyz[p].open;
            for (size_t e = 0; e < open.size();) {
                bool live = false;
                for (int k = 0; k < 6; k++) live = live || memcmp(&open[e].first, &all[p].h[k], sizeof(cudaIpcMemHandle_t)) == 0;
                if (live) { e++; continue; }
                cudaIpcCloseMemHandle(open[e].second);
                open.erase(open.begin() + e);
            }
        }
        all_agree(true); // barrier: every process has closed its stale mappings
        free_retired();
        push_tab.nseg = 0;
        push_to = push_from = 0;
        for (const Xfer& t : pos_plan) {
            if (!ok) break;
            if (t.kind == 2) { push_from |= 1u << t.peer_proc; continue; }
            if (t.len == 0) continue;
            if (push_tab.nseg >= 32) { ok = false; break; }
            PushSeg& sg = push_tab.seg[push_tab.nseg];
            sg.src_start = t.src_start;
            sg.len       = t.len;
            void* base[3];
            int nl;
            if (t.kind == 0) { // this brick is its own neighbor along an axis with one brick
                for (int k = 0; k < 3; k++) base[k] = arr[k];
                nl = b->Nlocal;
            } else {
                const int p = t.peer_proc;
                IpcSlot& ps = peer_xyz[p];
                for (int k = 0; k < 3 && ok; k++) {
                    base[k] = nullptr;
                    for (auto& e : ps.open)
                        if (memcmp(&e.first, &all[p].h[k], sizeof(cudaIpcMemHandle_t)) == 0) base[k] = e.second;
                    if (base[k]) continue;
                    void* q = nullptr;
                    if (cudaIpcOpenMemHandle(&q, all[p].h[k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; cudaGetLastError(); break; }
                    ps.open.push_back(std::make_pair(all[p].h[k], q));
                    base[k] = q;
                }
                if (!ok) break;
                nl = all[p].nlocal;
                push_to |= 1u << p;
            }
            sg.dx = (real*)base[0] + nl + t.dst_start;
            sg.dy = (real*)base[1] + nl + t.dst_start;
            sg.dz = (real*)base[2] + nl + t.dst_start;
            push_tab.nseg++;
        }
        // segments ascending in src_start (k_dd_push locates an entry by counting the starts below it)
        std::sort(push_tab.seg, push_tab.seg + push_tab.nseg, [](const PushSeg& a, const PushSeg& c) { return a.src_start < c.src_start; });
        if (!all_agree(ok)) push_ready = false; // fall back to NCCL send/recv for good, on every process alike
    }
    void forward_push() // updatePbc (pbc.c:42-55) across GPUs by peer stores, see dd_kernels.cuh
    {
        Brick* b = bricks[0];
        epoch++;
        const unsigned long long timeout = 30ull * 1000 * 1000 * 1000;
        // ready: my senders may overwrite my ghosts; wait until my receivers have released theirs
        if (push_from) MDB_LAUNCH(launches, k_dd_signal, 1, 32, 0, stream, peer_flags, push_from, 0, proc, epoch);
        if (push_to) MDB_LAUNCH(launches, k_dd_wait, 1, 32, 0, stream, sync_flags, push_to, 0, epoch, timeout);
        if (b->n_gsend)
            MDB_LAUNCH(launches, k_dd_push<real>, grid_for(b->n_gsend, 256), 256, 0, stream, b->n_gsend, b->gtab, push_tab, b->dd_gsend.p,
                b->xprd, b->yprd, b->zprd, b->x.p, b->y.p, b->z.p);
        // done: this epoch's positions have landed at my receivers; wait for the same from my senders
        if (push_to) MDB_LAUNCH(launches, k_dd_signal, 1, 32, 0, stream, peer_flags, push_to, 1, proc, epoch);
        if (push_from) MDB_LAUNCH(launches, k_dd_wait, 1, 32, 0, stream, sync_flags, push_from, 1, epoch, timeout);
    }
    void check_push_error()
    {
        if (!push_ready || !sync_flags) return;
        int e = 0;
        MDB_CUDA(cudaMemcpyAsync(&e, sync_flags + 96, sizeof(int), cudaMemcpyDeviceToHost, stream));
        MDB_CUDA(cudaStreamSynchronize(stream));
        if (e) throw Error("decomposition: a peer GPU did not answer the halo handshake within 30 s");
    }
    ~DomainGroup() override
    {
        cudaSetDevice(device);
        cudaStreamSynchronize(stream);
        for (size_t p = 0; p < peer_xyz.size(); p++)
            for (auto& e : peer_xyz[p].open) cudaIpcCloseMemHandle(e.second);
        if (push_ready && comm) { // nobody may free an exported array while a peer still maps it or pushes into it
            try { all_agree(true); } catch (...) {}
        }
        if (sync_flags) {
            for (int p = 0; p < topo.nprocs && p < 32; p++)
                if (p != proc && peer_flags.p[p]) cudaIpcCloseMemHandle(peer_flags.p[p]);
            cudaFree(sync_flags);
        }
        ipc_stage.release();
        if (comm) nccl_api().CommDestroy(comm);
        for (Brick* b : bricks) { b->brick_release(); delete b; }
        free_retired();
        d_off_all.release();
        d_sum.release();
        for (auto& b : st) b.release();
        st_tag.release(); st_flag.release(); st_scan.release();
        cudaFreeHost(h_off_all);
        cudaFreeHost(h_sum);
        for (auto& e : ev) cudaEventDestroy(e);
        cudaStreamDestroy(own_stream);
    }
    void setStream(cudaStream_t s) override
    {
        MDB_CUDA(cudaStreamSynchronize(stream));
        stream = s ? s : own_stream;
        for (Brick* b : bricks) b->setStream(stream);
    }
    void sync() override { MDB_CUDA(cudaStreamSynchronize(stream)); }
    bool mine(int b) const { return topo.owner(b) == proc; }
    Brick* local(int b) const { return bricks[b - first]; }

    // ------------------------------------------------------------------ transport
    // after phase B of every local brick: all slot offsets to every process and to the host
    void gather_offsets()
    {
        if (topo.nprocs > 1) {
            const size_t n = (size_t)nlocal_bricks * 32;
            MDB_NCCL(nccl_api().AllGather(d_off_all.p + (size_t)first * 32, d_off_all.p, n, NcclApi::Int32, comm, stream));
        }
        MDB_CUDA(cudaMemcpyAsync(h_off_all, d_off_all.p, (size_t)topo.nbricks * 32 * sizeof(int), cudaMemcpyDeviceToHost, stream));
        MDB_CUDA(cudaStreamSynchronize(stream));
        for (int b = 0; b < topo.nbricks; b++) {
            int dir[26], peer[26];
            const int ns = topo.slots(b, true, dir, peer);
            for (int d = 0; d < 26; d++) cnt[(size_t)b * 26 + d] = 0;
            for (int k = 0; k < ns; k++) cnt[(size_t)b * 26 + dir[k]] = h_off_all[b * 32 + k + 1] - h_off_all[b * 32 + k];
        }
    }
    // what brick r receives: segments (sender, first entry, length) in (sender, direction) order
    struct Seg {
        int sender, start, len;
    };
    int incoming(int r, std::vector<Seg>& segs) const
    {
        int dir[26], peer[26];
        const int ns = topo.slots(r, false, dir, peer);
        segs.clear();
        int total = 0;
        for (int k = 0; k < ns; k++) {
            const int n = cnt[(size_t)peer[k] * 26 + dir[k]];
            if (segs.empty() || segs.back().sender != peer[k]) segs.push_back(Seg { peer[k], total, 0 });
            segs.back().len += n;
            total += n;
        }
        return total;
    }
    // schedule of one exchange (dd_schedule) with pointers: W arrays of `elem` bytes per entry.
    // src(S) = send buffer of local brick S (per-peer SoA segments), dst(R, k) = k-th destination array of
    // local brick R at the first entry to be written.
    template <class SrcFn, class DstFn> void build_ops(std::vector<Op>& ops, int W, size_t elem, bool migrate, SrcFn src, DstFn dst)
    {
        ops.clear();
        std::vector<Xfer> plan;
        dd_schedule(topo, proc, cnt.data(), plan);
        for (const Xfer& t : plan) {
            if (t.kind != 2) { // the sender's own table must agree with the plan derived from the counts
                int ps = 0, pl = 0;
                local(t.src)->brick_segment(migrate, t.dst, &ps, &pl);
                if (pl != t.len || ps != t.src_start) throw Error("decomposition: send/receive segment mismatch");
            }
            for (int k = 0; k < W; k++) {
                const char* s = t.kind != 2 ? (const char*)src(local(t.src)) + ((size_t)W * t.src_start + (size_t)k * t.len) * elem : nullptr;
                char* d       = t.kind != 1 ? (char*)dst(local(t.dst), k) + (size_t)t.dst_start * elem : nullptr;
                const size_t bytes = (size_t)t.len * elem;
                ops.push_back(Op { t.kind == 0 ? COPY : (t.kind == 1 ? SEND : RECV), s, d, bytes, t.peer_proc });
            }
        }
    }
    void run_ops(const std::vector<Op>& ops) { run_ops(ops, stream); }
    void run_ops(const std::vector<Op>& ops, cudaStream_t st)
    {
        NcclApi& N = nccl_api();
        bool grouped = false;
        for (const Op& o : ops) {
            if (o.kind == COPY) {
                MDB_CUDA(cudaMemcpyAsync(o.dst, o.src, o.bytes, cudaMemcpyDeviceToDevice, st));
            } else {
                if (!grouped) { MDB_NCCL(N.GroupStart()); grouped = true; }
                if (o.kind == SEND) MDB_NCCL(N.Send(o.src, o.bytes, NcclApi::Int8, o.peer, comm, st));
                else MDB_NCCL(N.Recv(o.dst, o.bytes, NcclApi::Int8, o.peer, comm, st));
            }
        }
        if (grouped) MDB_NCCL(N.GroupEnd());
    }
    void sum_over_procs(double* h, int n) // h[0..n) += the other processes' (h is pinned)
    {
        if (topo.nprocs == 1) return;
        MDB_CUDA(cudaMemcpyAsync(d_sum.p, h, n * sizeof(double), cudaMemcpyHostToDevice, stream));
        MDB_NCCL(nccl_api().AllReduce(d_sum.p, d_sum.p, n, NcclApi::Float64, NcclApi::Sum, comm, stream));
        MDB_CUDA(cudaMemcpyAsync(h, d_sum.p, n * sizeof(double), cudaMemcpyDeviceToHost, stream));
        MDB_CUDA(cudaStreamSynchronize(stream));
    }

    // ------------------------------------------------------------------ atoms
    long long createAtom() override
    {
        for (Brick* b : bricks) b->brick_createAtom();
        return gNatoms;
    }
    // what an input reader hands over (atom.c:199-562), decomposed: n atoms of THIS process's bricks from
    // host SoA buffers in the global frame, with their global tags; each is dealt to the brick it lies in
    DBuf<real> st[6];
    DBuf<int> st_tag, st_flag, st_scan;
    void setAtoms(long long n, const int* tags, const void* ax, const void* ay, const void* az, const void* avx,
        const void* avy, const void* avz) override
    {
        if (n < 0 || n > 2000000000LL || !tags || !ax || !ay || !az) throw Error("mdb_dd_setAtoms: bad arguments");
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
        for (Brick* b : bricks) {
            b->derive();
            MDB_LAUNCH(launches, k_dd_select<real>, grid_for(n, 256), 256, 0, stream, (int)n, b->xprd, b->yprd, b->zprd,
                topo.g[0], topo.g[1], topo.g[2], b->bcoord[0], b->bcoord[1], b->bcoord[2], st[0].p, st[1].p, st[2].p,
                st_flag.p);
            b->scanner.exclusive(st_flag.p, st_scan.p, n, st_scan.p + n, stream);
            int cntb = 0;
            MDB_CUDA(cudaMemcpyAsync(&cntb, st_scan.p + n, sizeof(int), cudaMemcpyDeviceToHost, stream));
            MDB_CUDA(cudaStreamSynchronize(stream));
            b->Nlocal = cntb;
            b->Natoms = cntb;
            b->Nghost = 0;
            b->ensure_atoms((size_t)cntb + cntb / 4 + 1024, false);
            if (cntb)
                MDB_LAUNCH(launches, k_dd_take<real>, grid_for(n, 256), 256, 0, stream, (int)n, st_flag.p, st_scan.p, b->xprd,
                    b->yprd, b->zprd, topo.g[0], topo.g[1], topo.g[2], b->bcoord[0], b->bcoord[1], b->bcoord[2], st[0].p,
                    st[1].p, st[2].p, has_v ? st[3].p : (const real*)nullptr, has_v ? st[4].p : (const real*)nullptr,
                    has_v ? st[5].p : (const real*)nullptr, st_tag.p, b->x.p, b->y.p, b->z.p, b->vx.p, b->vy.p, b->vz.p,
                    b->type.p, b->orig.p);
            b->zero3(b->fx.p, b->fy.p, b->fz.p, cntb);
            b->neigh_ready = false;
            taken += cntb;
        }
        MDB_CUDA(cudaStreamSynchronize(stream)); // host buffers may be reused by the caller
        if (taken != n) throw Error(fmt("mdb_dd_setAtoms: %lld of %lld atoms lie outside this process's bricks", n - taken, n));
    }
    void setEam(int nrho, double drho, int nr, double dr, double cut, double mass, const double* frho, const double* zr,
        const double* rhor) override
    {
        G.force_field = MDB_FF_EAM;
        for (Brick* b : bricks) b->setEam(nrho, drho, nr, dr, cut, mass, frho, zr, rhor);
    }

    // ------------------------------------------------------------------ thermo (common/thermo.c)
    void global_vel_sums(double out[4])
    {
        for (size_t k = 0; k < bricks.size(); k++) bricks[k]->vel_sums(d_sum.p + 4 * k);
        MDB_CUDA(cudaMemcpyAsync(h_sum + 8, d_sum.p, 4 * bricks.size() * sizeof(double), cudaMemcpyDeviceToHost, stream));
        MDB_CUDA(cudaStreamSynchronize(stream));
        for (int c = 0; c < 4; c++) {
            h_sum[c] = 0;
            for (size_t k = 0; k < bricks.size(); k++) h_sum[c] += h_sum[8 + 4 * k + c];
        }
        sum_over_procs(h_sum, 4);
        for (int c = 0; c < 4; c++) out[c] = h_sum[c];
    }
    void computeThermo(double* T, double* P) override
    {
        for (Brick* b : bricks)
            if (!b->thermo_ready) b->setupThermo();
        double s[4];
        global_vel_sums(s);
        bricks[0]->thermo_from_sum(s[3], T, P);
    }
    void adjustThermo() // thermo.c:82-122 over all bricks
    {
        double s[4];
        global_vel_sums(s);
        const real vxtot = (real)s[0] / (real)gNatoms, vytot = (real)s[1] / (real)gNatoms, vztot = (real)s[2] / (real)gNatoms;
        for (Brick* b : bricks)
            MDB_LAUNCH(launches, k_vel_shift<real>, grid_for(b->Nlocal, 256), 256, 0, stream, b->Nlocal, b->vx.p, b->vy.p,
                b->vz.p, vxtot, vytot, vztot);
        global_vel_sums(s);
        real t = (real)s[3];
        t *= bricks[0]->t_scale;
        const real factor = (real)sqrt((double)(bricks[0]->temp / t));
        for (Brick* b : bricks)
            MDB_LAUNCH(launches, k_vel_scale<real>, grid_for(b->Nlocal, 256), 256, 0, stream, b->Nlocal, b->vx.p, b->vy.p,
                b->vz.p, factor);
    }

    // ------------------------------------------------------------------ ghosts: setupPbc + updatePbc across bricks
    void setupGhosts()
    {
        for (Brick* b : bricks) b->brick_border_A();
        MDB_CUDA(cudaStreamSynchronize(stream));
        for (Brick* b : bricks) b->brick_sendlist_B(false);
        gather_offsets();
        std::vector<Seg> segs;
        for (Brick* b : bricks) {
            b->brick_finish_table(false, h_off_all + b->brick_id * 32);
            b->brick_ghost_alloc(incoming(b->brick_id, segs));
        }
        // ghost type + tag (once per rebuild), then the per-step schedules
        std::vector<Op> int_ops;
        build_ops(int_ops, 2, sizeof(int), false, [](Brick* s) { return (const void*)s->dd_isend.p; },
            [](Brick* r, int k) { return (void*)((k == 0 ? r->type.p : r->orig.p) + r->Nlocal); });
        for (Brick* b : bricks) b->brick_pack_ints();
        run_ops(int_ops);
        build_ops(pos_ops, 3, sizeof(real), false, [](Brick* s) { return (const void*)s->dd_sendbuf.p; },
            [](Brick* r, int k) { return (void*)((k == 0 ? r->x.p : (k == 1 ? r->y.p : r->z.p)) + r->Nlocal); });
        if (push_ready && push_wanted) {
            dd_schedule(topo, proc, cnt.data(), pos_plan);
            prepare_push();
        }
        if (G.force_field == MDB_FF_EAM)
            build_ops(fp_ops, 1, sizeof(real), false, [](Brick* s) { return (const void*)s->dd_sendbuf.p; },
                [](Brick* r, int) { return (void*)(r->fp.p + r->Nlocal); });
    }
    void forward() // updatePbc (pbc.c:42-55): fresh positions of the border atoms to their images
    {
        if (timing) MDB_CUDA(cudaEventRecord(ev[2], stream));
        if (push_ready && push_wanted) {
            forward_push();
        } else {
            for (Brick* b : bricks) b->brick_pack_pos();
            run_ops(pos_ops);
        }
        if (timing) {
            float ms = 0;
            MDB_CUDA(cudaEventRecord(ev[3], stream));
            MDB_CUDA(cudaEventSynchronize(ev[3]));
            MDB_CUDA(cudaEventElapsedTime(&ms, ev[2], ev[3]));
            comm_ms += ms;
        }
    }
    // ------------------------------------------------------------------ migration: updateAtomsPbc across bricks
    void migrate()
    {
        for (Brick* b : bricks) b->brick_migrate_A();
        MDB_CUDA(cudaStreamSynchronize(stream));
        for (Brick* b : bricks) b->brick_sendlist_B(true);
        gather_offsets();
        std::vector<Seg> segs;
        std::vector<int> nin(bricks.size());
        for (size_t k = 0; k < bricks.size(); k++) {
            Brick* b = bricks[k];
            b->brick_finish_table(true, h_off_all + b->brick_id * 32);
            nin[k] = incoming(b->brick_id, segs);
            b->brick_migrate_pack(nin[k]);
        }
        std::vector<Op> ops;
        build_ops(ops, 6, sizeof(real), true, [](Brick* s) { return (const void*)s->dd_sendbuf.p; },
            [](Brick* r, int k) {
                real* a[6] = { r->x.p, r->y.p, r->z.p, r->vx.p, r->vy.p, r->vz.p };
                return (void*)(a[k] + r->nstay);
            });
        run_ops(ops);
        build_ops(ops, 2, sizeof(int), true, [](Brick* s) { return (const void*)s->dd_isend.p; },
            [](Brick* r, int k) { return (void*)((k == 0 ? r->type.p : r->orig.p) + r->nstay); });
        run_ops(ops);
        for (size_t k = 0; k < bricks.size(); k++) {
            bricks[k]->Nlocal = bricks[k]->nstay + nin[k];
            bricks[k]->Natoms = bricks[k]->Nlocal;
            bricks[k]->Nghost = 0;
        }
    }

    // ------------------------------------------------------------------ driver flow (verletlist/main.c)
    void setup(bool adjust) override // main.c:58-72
    {
        for (Brick* b : bricks) {
            b->setupNeighbor();
            b->thermo_ready = false;
            b->derive_dtforce();
            b->setupThermo();
        }
        if (adjust) adjustThermo();
        for (Brick* b : bricks) b->sort_atoms();
        setupGhosts();
        forward();
        for (Brick* b : bricks) b->buildNeighbor();
    }
    void reneighbour() override // main.c:76-95
    {
        NvtxRange nvtx_range_("reneighbour");
        check_push_error();
        for (Brick* b : bricks) b->xy_valid = false; // migration and sorting move atoms between slots
        migrate();
        for (Brick* b : bricks) b->sort_atoms();
        setupGhosts();
        forward();
        for (Brick* b : bricks) b->buildNeighbor();
    }
    void force()
    {
        if (G.force_field == MDB_FF_EAM) { // force_eam.c: density -> fp of the images -> force
            for (Brick* b : bricks) b->eam_density();
            for (Brick* b : bricks) b->brick_pack_fp();
            run_ops(fp_ops);
            for (Brick* b : bricks) { b->eam_force(); b->force_launches++; }
        } else {
            for (Brick* b : bricks) b->launch_force(FORCE_DISPATCH);
        }
    }
    void run(int nsteps, double* thermo_out, int max_records, int* nrecords, double* timers) override // main.c:244-288
    {
        for (Brick* b : bricks) {
            if (!b->thermo_ready) b->setupThermo();
            b->xy_valid = false; // positions may have been set from outside since the last run
        }
        const int nstat = G.nstat > 0 ? G.nstat : nsteps + 1;
        const int every = G.reneigh_every > 0 ? G.reneigh_every : nsteps + 1;
        const size_t maxrec = nsteps / nstat + 3;
        if (4 * maxrec * bricks.size() > 2048) throw Error("run: too many thermo records");
        for (Brick* b : bricks) b->d_thermo.ensure(4 * maxrec, false, stream);
        std::vector<int> rec_step;
        auto record = [&](int step) {
            for (Brick* b : bricks) b->vel_sums(b->d_thermo.p + 4 * rec_step.size());
            rec_step.push_back(step);
        };
        double f0 = 0, n0 = 0;
        for (Brick* b : bricks) { f0 += b->force_ms; n0 += b->neigh_ms; }
        record(0);
        force();
        MDB_CUDA(cudaEventRecord(ev[0], stream)); // TOTAL starts after the first force, main.c:252
        bool initial_done = false;
        for (int n = 0; n < nsteps; n++) {
            const bool reneigh = (n + 1) % every == 0;
            if (!initial_done)
                for (Brick* b : bricks) b->initialIntegrate();
            const bool rec = !((n + 1) % nstat) && (n + 1) < nsteps;
            // force(n) + finalIntegrate(n) + initialIntegrate(n+1) in one launch per brick (x, y, z updated in place, gathers
            // from the double-buffered copies), unless something reads the state in between
            bool fuse = !rec && n + 1 < nsteps && G.force_field == MDB_FF_LJ;
            for (Brick* b : bricks) fuse = fuse && b->can_fuse_force_inplace();
            if (fuse) {
                if (reneigh) reneighbour();
                else forward();
                for (Brick* b : bricks) b->forceFinalInitialIntegrateInPlace();
                initial_done = true;
                continue;
            }
            if (reneigh) {
                reneighbour();
                force();
            } else {
                forward();
                force();
            }
            if (rec || n + 1 == nsteps || !bricks[0]->fuse_integrate) {
                for (Brick* b : bricks) b->finalIntegrate();
                initial_done = false;
                if (rec) record(n + 1);
            } else {
                for (Brick* b : bricks) b->finalInitialIntegrate();
                initial_done = true;
            }
        }
        MDB_CUDA(cudaEventRecord(ev[1], stream));
        record(nsteps);
        check_push_error();
        const size_t nr4 = 4 * rec_step.size();
        for (size_t k = 0; k < bricks.size(); k++)
            MDB_CUDA(cudaMemcpyAsync(h_sum + 2048 + k * nr4, bricks[k]->d_thermo.p, nr4 * sizeof(double), cudaMemcpyDeviceToHost, stream));
        MDB_CUDA(cudaStreamSynchronize(stream));
        for (size_t q = 0; q < nr4; q++) {
            h_sum[q] = 0;
            for (size_t k = 0; k < bricks.size(); k++) h_sum[q] += h_sum[2048 + k * nr4 + q];
        }
        sum_over_procs(h_sum, (int)nr4);
        float ms = 0;
        MDB_CUDA(cudaEventElapsedTime(&ms, ev[0], ev[1]));
        int nr = 0;
        for (size_t r = 0; r < rec_step.size(); r++)
            if (thermo_out && nr < max_records) {
                thermo_out[3 * nr] = rec_step[r];
                bricks[0]->thermo_from_sum(h_sum[4 * r + 3], &thermo_out[3 * nr + 1], &thermo_out[3 * nr + 2]);
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
    void getCounts(long long* v) override // global atoms, this process: local atoms, ghosts, largest maxneighs, bricks
    {
        v[0] = gNatoms; v[1] = v[2] = v[3] = 0;
        for (Brick* b : bricks) {
            v[1] += b->Nlocal;
            v[2] += b->Nghost;
            v[3] = std::max<long long>(v[3], b->maxneighs);
        }
        v[4] = (long long)bricks.size();
    }
    void getAtoms(int which, int* tags, void* ax, void* ay, void* az) override // slot order of the bricks, global frame
    {
        size_t off = 0;
        for (Brick* b : bricks) {
            const size_t n = b->Nlocal;
            const real *p, *q, *r;
            if (which == 'x') { p = b->x.p; q = b->y.p; r = b->z.p; }
            else if (which == 'v') { p = b->vx.p; q = b->vy.p; r = b->vz.p; }
            else if (which == 'f') { p = b->fx.p; q = b->fy.p; r = b->fz.p; }
            else throw Error("getAtoms: which must be 'x', 'v' or 'f'");
            if (which == 'x') {
                b->tx.ensure(n, false, stream); b->ty.ensure(n, false, stream); b->tz.ensure(n, false, stream);
                MDB_LAUNCH(launches, k_dd_to_global<real>, grid_for(n, 256), 256, 0, stream, (int)n,
                    (real)(b->bcoord[0] * b->xprd), (real)(b->bcoord[1] * b->yprd), (real)(b->bcoord[2] * b->zprd), p, q, r,
                    b->tx.p, b->ty.p, b->tz.p);
                p = b->tx.p; q = b->ty.p; r = b->tz.p;
            }
            MDB_CUDA(cudaMemcpyAsync((real*)ax + off, p, n * sizeof(real), cudaMemcpyDeviceToHost, stream));
            MDB_CUDA(cudaMemcpyAsync((real*)ay + off, q, n * sizeof(real), cudaMemcpyDeviceToHost, stream));
            MDB_CUDA(cudaMemcpyAsync((real*)az + off, r, n * sizeof(real), cudaMemcpyDeviceToHost, stream));
            if (tags) MDB_CUDA(cudaMemcpyAsync(tags + off, b->orig.p, n * sizeof(int), cudaMemcpyDeviceToHost, stream));
            off += n;
        }
        MDB_CUDA(cudaStreamSynchronize(stream));
    }
    void getNeighborTags(int* tags, int* nn, int* rows, int stride) override
    {
        size_t off = 0;
        for (Brick* b : bricks) {
            const size_t n = b->Nlocal;
            if (b->nstride == 0) throw Error("getNeighborTags: no neighbor list");
            b->rows.ensure(n * stride, false, stream);
            MDB_LAUNCH(launches, k_dd_rows_as_tags, grid_for(n, 128), 128, 0, stream, (int)n, stride, b->LL, b->numneigh.p,
                b->neighbors.p, b->orig.p, b->rows.p);
            MDB_CUDA(cudaMemcpyAsync(rows + off * stride, b->rows.p, n * stride * sizeof(int), cudaMemcpyDeviceToHost, stream));
            MDB_CUDA(cudaMemcpyAsync(nn + off, b->numneigh.p, n * sizeof(int), cudaMemcpyDeviceToHost, stream));
            MDB_CUDA(cudaMemcpyAsync(tags + off, b->orig.p, n * sizeof(int), cudaMemcpyDeviceToHost, stream));
            off += n;
        }
        MDB_CUDA(cudaStreamSynchronize(stream));
    }
    void saveState() override
    {
        for (Brick* b : bricks) b->brick_save();
    }
    void restoreState() override
    {
        for (Brick* b : bricks) b->brick_restore();
    }
    void setOption(const char* name, double v) override
    {
        if (!strcmp(name, "halo_push")) { push_wanted = v != 0; return; }
        for (Brick* b : bricks) b->setOption(name, v);
    }
    void setTiming(bool on) override
    {
        timing = on;
        for (Brick* b : bricks) b->timing = on;
    }
    void stats(double* force_ms, long long* force_launches, double* neigh_ms, long long* neigh_launches, long long* nl,
        double* cms, bool reset) override
    {
        double f = 0, n = 0;
        long long fl = 0, nbl = 0, l = launches;
        for (Brick* b : bricks) {
            f += b->force_ms; n += b->neigh_ms; fl += b->force_launches; nbl += b->neigh_launches; l += b->launches;
            if (reset) { b->force_ms = b->neigh_ms = 0; b->force_launches = b->neigh_launches = b->launches = 0; }
        }
        if (force_ms) *force_ms = f;
        if (force_launches) *force_launches = fl;
        if (neigh_ms) *neigh_ms = n;
        if (neigh_launches) *neigh_launches = nbl;
        if (nl) *nl = l;
        if (cms) *cms = comm_ms;
        if (reset) { launches = 0; comm_ms = 0; }
    }
};

DDBase* make_dd(const mdb_params& g, const int grid[3], int nprocs, int proc, const void* nccl_id, int device)
{
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) throw Error("mdb_dd_create: no CUDA device (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) throw Error(fmt("mdb_dd_create: device %d out of range (%d devices)", device, ndev));
    if (g.precision == MDB_DP) return new DomainGroup<double>(g, grid, nprocs, proc, nccl_id, device);
    if (g.precision == MDB_SP) return new DomainGroup<float>(g, grid, nprocs, proc, nccl_id, device);
    throw Error("mdb_dd_create: precision must be MDB_SP or MDB_DP");
}
