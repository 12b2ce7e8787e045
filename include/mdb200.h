/*
 * mdb200.h -- C ABI of libmdb200: a B200 (sm_100a) implementation of MD-Bench's short-range force
 * hot path (binning -> neighbor build -> LJ/EAM force -> velocity-Verlet -> PBC ghosts).
 *
 * This is the drop-in boundary.  MD-Bench's "plugin API" is a set of global C function pointers
 * (reference src/verletlist/force.h:16-17, neighbor.h:55-56, integrate.h:12-14, pbc.h:15-17) plus
 * a few plain functions called from main.c (setupPbc pbc.h:22, setupNeighbor neighbor.h:59,
 * initDevice device.h:20, computeThermo/adjustThermo thermo.h).  Every entry point below names
 * the reference interface it replaces.  The reference passes its Atom/Neighbor/Parameter structs,
 * whose layout depends on compile-time -DPRECISION / -DAOS; here the same choices are RUN-TIME
 * fields of mdb_params and all state lives behind an opaque handle, so one library serves the
 * SP/DP and AOS/SOA builds of the driver.  INTEGRATION.md shows the shim that binds these entry
 * points to the reference's function pointers.
 *
 * Conventions
 *   - plain C, plain pointers and sizes; no C++ or torch types.
 *   - host buffers are `precision`-typed (float for MDB_SP, double for MDB_DP) and laid out as the
 *     reference's atom_x()/atom_y()/atom_z() macros expect (verletlist/atom.h:51-73):
 *       MDB_AOS: x points to 3*n reals {x0,y0,z0,x1,...}; y,z are ignored (may be NULL);
 *       MDB_SOA: x,y,z point to n reals each.
 *   - functions returning int return 0 on success, non-zero on failure; mdb_last_error() gives the
 *     message.  (The reference prints and exit()s, common/device.c:15-21; the driver in
 *     md-bench_b200/driver does the same on a non-zero return.)
 *   - there is NO CPU fallback: without a CUDA device mdb_create() fails.
 *   - one mdb_ctx = one spatial domain on one GPU; not re-entrant per ctx (like the reference's
 *     file-static state), distinct ctx are independent.
 */
#ifndef MDB200_H
#define MDB200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDB200_ABI_VERSION 1

enum { MDB_SP = 1, MDB_DP = 2 };     /* = reference -DPRECISION (config.mk:103-107) */
enum { MDB_AOS = 0, MDB_SOA = 1 };   /* = reference DATA_LAYOUT / -DAOS (config.mk:100-102) */
enum { MDB_FF_LJ = 0, MDB_FF_EAM = 1 }; /* = enum forcetype, verletlist/force.h:19 */

/* Mirror of the reference's Parameter (common/parameter.h:27-61); same field names and meaning.
 * Reals are passed as double and narrowed to `precision` exactly where the reference narrows. */
typedef struct mdb_params {
    int precision;     /* MDB_SP | MDB_DP */
    int layout;        /* MDB_AOS | MDB_SOA: layout of HOST buffers passed to/from this library */
    int force_field;   /* MDB_FF_LJ | MDB_FF_EAM */
    double epsilon, sigma, temp, rho, mass;
    int ntypes, ntimes, nstat, reneigh_every, half_neigh;
    double dt, skin, cutforce;
    int nx, ny, nz;
    int pbc_x, pbc_y, pbc_z;
    /* box from an input file (readers set xlo..zhi, atom.c:199-562); from_input = 0 -> generated
     * FCC lattice box nx*lattice (main.c:42-45) */
    int from_input;
    double xlo, xhi, ylo, yhi, zlo, zhi;
    /* EAM overrides (initEam, common/eam_utils.c:27-35) are applied by mdb_setEam() */
} mdb_params;

typedef struct mdb_ctx mdb_ctx;

/* ---- lifecycle ------------------------------------------------------------------------------ */
int         mdb_abi_version(void);
const char* mdb_last_error(void);
/* defaults of initParameter, common/parameter.c:16-51 (precision DP, layout AOS) */
void        mdb_default_params(mdb_params* p);
/* replaces initAtom/initPbc/initStats/initNeighbor + initDevice (main.c:47-51,68; device_spec.c:11).
 * device = CUDA ordinal.  Derives lattice, box, cutneigh = cutforce + skin (main.c:233), sigma6,
 * dtforce (parameter.c:115-120).  Returns NULL on failure. */
mdb_ctx*    mdb_create(const mdb_params* p, int device);
void        mdb_destroy(mdb_ctx* c);
/* tuning switches (A/B measurements, debugging); defaults are the fastest measured variants (profiles/):
 *  "sort_atoms" (default 0, like the reference's SORT_ATOMS build option; 1 for the bricks of a decomposed box): re-sort
 *      the local atoms by neighbor bin at every rebuild (neighbor.c:360-426); the permutation is tracked, every accessor
 *      below still speaks the reference's atom numbering.
 *  "sort_block" order of the bins for "sort_atoms": 0 = the reference's x-fastest bin order, B > 0 = blocks of B x B x B bins
 *      (a thread block's 128 atoms then fill a compact box, like the generator's 8x8x8 sub-boxes).
 *  "eam_variant" EAM passes: 2 (default) packed (x, y) / (z, fp) gathers and (value, slope) tables, 1 = packed spline rows
 *      (what a brick of a decomposed box runs), 0 = first kernels (scalar table gathers, IEEE sqrt / division).
 *  "fuse_integrate" (default 1) finalIntegrate(n) + initialIntegrate(n+1) in one pass inside mdb_run.
 *  "fuse_force" (default 1) inside mdb_run, LJ full lists of a single domain: computeForce(n) + finalIntegrate(n) +
 *      initialIntegrate(n+1) in ONE kernel (the integrate halves run in the force kernel's epilogue on the force still
 *      in registers; positions are double-buffered).  Bit-identical to the separate operators; steps that record thermo
 *      and the last step of a run keep the separate kernels, so f/x/v read back after mdb_run are unchanged.
 *  "xy_gather" (default 1) the fused kernel fetches a neighbor's x and y with one 2-element vector gather from a packed
 *      (x, y) copy of the positions (kept current by its own epilogue and by updatePbc) instead of two scalar gathers:
 *      fewer L1 wavefronts per pair.  Bit-identical.
 *  "ghost_epilogue" (default -1 = on for domains of up to 2^19 atoms, 0 off, 1 on) single domain, fused step: the kernel's
 *      epilogue also writes the periodic images of the atom (updatePbc, pbc.c:42-55), so a step between two rebuilds is
 *      ONE launch; the next mdb_updatePbc finds them current.  Bit-identical (the same single fma per coordinate).
 *      BASELINE config 1 (32^3): 2.69 -> 2.93 G atom updates/s; at 128^3 it costs 1.4 % of the force kernel.
 *  "lazy_ops" (default 0; 3.91 -> 4.34 G atom updates/s for MDBench-VL-B200 --operators at 128^3) for drivers that keep
 *      the reference's operator-by-operator loop: mdb_computeForce and mdb_finalIntegrate only record that they are due;
 *      if the next call is mdb_initialIntegrate the three run as the one fused kernel of mdb_run, any other entry point
 *      first launches them separately, in order.  Same results; mdb_computeForce then returns 0 s. */
int         mdb_setOption(mdb_ctx* c, const char* name, double value);
/* run all work of this ctx on the given cudaStream_t (passed as void*); NULL = ctx-owned stream */
int         mdb_setStream(mdb_ctx* c, void* cuda_stream);
int         mdb_sync(mdb_ctx* c);

/* ---- atoms ---------------------------------------------------------------------------------- */
/* createAtom (verletlist/atom.c:67-187): FCC lattice + Park-Miller velocities, generated on the
 * device with the reference's emission order (so atom indices match). Returns Natoms or <0. */
long long   mdb_createAtom(mdb_ctx* c);
/* what the file readers (atom.c:199-562) hand over: n atoms from HOST buffers */
int         mdb_setAtoms(mdb_ctx* c, long long n, const void* x, const void* y, const void* z,
                         const void* vx, const void* vy, const void* vz, const int* type);
/* same, buffers already resident in device memory (same precision/layout rules) */
int         mdb_setAtomsDevice(mdb_ctx* c, long long n, const void* x, const void* y, const void* z,
                               const void* vx, const void* vy, const void* vz, const int* type);
/* copy local atoms back to HOST buffers; which = 'x' | 'v' | 'f'; with_ghosts!=0 appends the
 * ghost atoms (x only), i.e. Nlocal+Nghost entries as in the reference's arrays */
int         mdb_getAtoms(mdb_ctx* c, int which, int with_ghosts, void* x, void* y, void* z);
/* Atom.type of the local atoms (Nlocal ints).  ntypes > 1 (the reference's EXPLICIT_TYPES builds, force_lj.c:61-67):
 * createAtom draws type = rand() % ntypes per atom from the host's rand() sequence in emission order (atom.c:159), the
 * readers hand types over through mdb_setAtoms; the pair parameters the reference looks up by type pair are the same for
 * every pair (atom.c:84-89 fills all ntypes^2 entries alike), so the kernels keep the scalars and results do not depend
 * on the types.  Single verletlist domains only (mdb_cp_create / mdb_dd_create reject ntypes != 1). */
int         mdb_getTypes(mdb_ctx* c, int* types);
/* Natoms, Nlocal, Nghost, Nmax, maxneighs (Atom / Neighbor scalar fields) */
int         mdb_getCounts(mdb_ctx* c, long long* Natoms, long long* Nlocal, long long* Nghost,
                          long long* Nmax, int* maxneighs);
/* keep / restore a device-side copy of the current x,v (bench: restart a run without H2D) */
int         mdb_saveState(mdb_ctx* c);
int         mdb_restoreState(mdb_ctx* c);

/* ---- thermo (common/thermo.c) ----------------------------------------------------------------- */
int         mdb_setupThermo(mdb_ctx* c);                         /* setupThermo, thermo.c:30-53 */
int         mdb_adjustThermo(mdb_ctx* c);                        /* adjustThermo, thermo.c:82-122 */
/* computeThermo, thermo.c:55-80: kinetic temperature and pressure of the local atoms */
int         mdb_computeThermo(mdb_ctx* c, double* T, double* P);

/* ---- operators: one entry point per reference function pointer ------------------------------ */
int         mdb_setupNeighbor(mdb_ctx* c);                       /* neighbor.c:64-184 */
int         mdb_setupPbc(mdb_ctx* c);                            /* pbc.c:98-227 (on the device) */
int         mdb_updatePbc(mdb_ctx* c, int reneigh);              /* UpdatePbcFunction updatePbc, pbc.c:42-55 */
int         mdb_updateAtomsPbc(mdb_ctx* c, int reneigh);         /* updateAtomsPbc, pbc.c:59-84 */
int         mdb_buildNeighbor(mdb_ctx* c);                       /* BuildNeighborFunction, neighbor.c:186-264 (+binatoms 329-358) */
/* ComputeForceFunction (force.h:16): return the seconds spent, like the reference; <0 on error */
double      mdb_computeForce(mdb_ctx* c);                        /* dispatch of initForce, force.c:13-34 */
double      mdb_computeForceLJFullNeigh(mdb_ctx* c);             /* force_lj.c:14-105 */
double      mdb_computeForceLJHalfNeigh(mdb_ctx* c);             /* force_lj.c:107-198 */
double      mdb_computeForceEam(mdb_ctx* c);                     /* force_eam.c:19-231 */
int         mdb_initialIntegrate(mdb_ctx* c, int reneigh);       /* IntegrationFunction, integrate.c:21-31 */
int         mdb_finalIntegrate(mdb_ctx* c, int reneigh);         /* integrate.c:33-40 */

/* ---- driver-level flow ------------------------------------------------------------------------ */
/* main.c:58-72 after the atoms exist: setupNeighbor, setupThermo, [adjustThermo if adjust!=0],
 * setupPbc, updatePbc, buildNeighbor */
int         mdb_setup(mdb_ctx* c, int adjust);
int         mdb_reneighbour(mdb_ctx* c);                         /* main.c:76-95 */
/* main.c:244-288: thermo(0), first force, nsteps iterations, thermo(-1).  The whole loop is
 * enqueued on the ctx stream without host round-trips except one flag read per rebuild.
 * thermo_out (may be NULL) receives up to max_records (step, T, P) triples; *nrecords their count.
 * timers (may be NULL) = {TOTAL, FORCE, NEIGH} seconds as the reference reports them
 * (TOTAL excludes the first force call, main.c:250-252); FORCE/NEIGH are only split out when
 * mdb_setTiming(c, 1) was called (adds event records + syncs per phase). */
int         mdb_run(mdb_ctx* c, int nsteps, double* thermo_out, int max_records, int* nrecords,
                    double* timers);
int         mdb_setTiming(mdb_ctx* c, int on);
/* accumulated CUDA-event time (ms) and launch count of the force kernel since the last reset
 * (only collected while timing is on); used by bench.py for the roofline */
int         mdb_getKernelStats(mdb_ctx* c, double* force_ms, long long* force_launches,
                               double* neigh_ms, long long* neigh_launches,
                               long long* total_launches);
int         mdb_resetKernelStats(mdb_ctx* c);

/* ---- EAM (common/eam_utils.c, verletlist/force_eam.c) ---------------------------------------- */
/* hand over funcfl tables as read from the potential file (0-based frho[nrho], zr[nr], rhor[nr],
 * doubles): applies initEam's overrides (mass, cutforce = cut, cutneigh = cut + 1, temp = 600,
 * dt = 0.001, rho = 0.07041125; eam_utils.c:27-35) and builds the spline tables (file2array,
 * array2spline, interpolate; eam_utils.c:95-284).  Call before mdb_createAtom/mdb_setup. */
int         mdb_setEam(mdb_ctx* c, int nrho, double drho, int nr, double dr, double cut, double mass,
                       const double* frho, const double* zr, const double* rhor);
/* or hand over finished spline tables (precision-typed) */
int         mdb_setEamSplines(mdb_ctx* c, int nr, int nrho, int nr_tot, int nrho_tot, double rdr,
                              double rdrho, const void* rhor_spline, const void* frho_spline,
                              const void* z2r_spline);
int         mdb_getEamSplines(mdb_ctx* c, int* nr, int* nrho, int* nr_tot, int* nrho_tot,
                              double* rdr, double* rdrho, void* rhor_spline, void* frho_spline,
                              void* z2r_spline);

/* ---- parity accessors (what the reference keeps in host-visible arrays) ------------------------ */
/* Neighbor.numneigh[Nlocal] and Neighbor.neighbors as row-major rows of `row_stride` ints
 * (neighbor.h:18-28); indices are the reference's (ghost g has index Nlocal+g). */
int         mdb_getNeighbors(mdb_ctx* c, int* numneigh, int* neighbors, int row_stride);
/* Atom.border_map and the file-static PBCx/y/z of pbc.c:17-18, Nghost entries each */
int         mdb_getGhostMap(mdb_ctx* c, int* border_map, int* PBCx, int* PBCy, int* PBCz);
/* bin geometry of setupNeighbor (neighbor.c:24-38): ints {nbinx,nbiny,nbinz,mbinx,mbiny,mbinz,
 * mbinxlo,mbinylo,mbinzlo,mbins,nstencil,max bin count}, reals {bininvx,y,z,binsizex,y,z,
 * cutneighsq,cutneigh,xprd,yprd,zprd,lattice,dtforce,cutforce} */
int         mdb_getNeighborParams(mdb_ctx* c, int ints[12], double reals[14]);
int         mdb_getStencil(mdb_ctx* c, int* stencil);            /* nstencil ints */
int         mdb_getBinCounts(mdb_ctx* c, int* bincount);          /* mbins ints, after buildNeighbor */
int         mdb_getEamFp(mdb_ctx* c, void* fp, int with_ghosts);   /* Eam.fp after computeForceEam */

/* ---- workload counters for the roofline (verletlist/stats.h:13-18 equivalents) ----------------- */
/* listed pairs (sum of numneigh) of the current list, and pairs inside the force cutoff for the
 * current positions; computed on demand by a counting kernel */
int         mdb_countPairs(mdb_ctx* c, long long* listed, long long* in_cutoff);

/* ---- kernel micro-benchmark (reference src/verletlist/main-stub.c) ----------------------------------- */
/* replaces createNeighbors (main-stub.c:62-106): a synthetic neighbor list for the atoms handed over with
 * mdb_setAtoms -- pattern 0 "seq" (i+1, i+2, ...), 1 "fix" (0 .. nneighs-1 for every atom), 2 "rand" --
 * nneighs entries replicated nreps times; mdb_computeForceLJFullNeigh / HalfNeigh then time the kernel alone */
enum { MDB_STUB_SEQ = 0, MDB_STUB_FIX = 1, MDB_STUB_RAND = 2 };
int         mdb_stubNeighbors(mdb_ctx* c, int pattern, int nneighs, int nreps, unsigned seed);

/* ---- multi-GPU: spatial decomposition ----------------------------------------------- */
/* The reference runs one domain ("replicas only").  Here a box may be cut into gx*gy*gz bricks, each
 * with a ghost shell of width cutneigh; what is generalised is the setupPbc / updatePbc /
 * updateAtomsPbc trio (verletlist/pbc.c:98-227, 42-55, 59-84): a periodic image now comes from the
 * neighbor brick in that direction (which is the brick itself along an axis with one brick, i.e. the
 * reference's scheme).  Bricks are dealt to processes in consecutive blocks (one process per GPU;
 * #bricks must be a multiple of #processes).  Bricks of one process exchange by device copies,
 * bricks of different processes over NVLink: per step by peer stores into the neighbor GPU's arrays (CUDA IPC + flag
 * handshake, one brick per process), per rebuild by NCCL send/recv (NCCL is dlopen()ed on first use).
 * Atoms carry GLOBAL tags = the index createAtom (atom.c:67-187) gives them in the undecomposed box,
 * so results are comparable with a single-domain run atom by atom. */
typedef struct mdb_dd mdb_dd;
int         mdb_dd_uniqueIdBytes(void);                 /* sizeof(ncclUniqueId) */
int         mdb_dd_getUniqueId(void* id);               /* process 0 calls this, the launcher broadcasts it */
/* host-only: the send (send!=0) or receive slots of `brick`: directions of the ADDGHOST ladder
 * (pbc.c:107-224) ordered by (peer brick, direction), their peer bricks and owning processes;
 * returns the number of slots or <0.  No CUDA call: used to check that all ranks derive one plan. */
int         mdb_dd_plan(int gx, int gy, int gz, int perx, int pery, int perz, int nprocs, int brick,
                        int send, int dir[26], int peer[26], int owner[26]);
/* host-only: the transfer schedule of one exchange as process `proc` executes it, given
 * cnt[brick*26 + direction] = entries each brick sends per direction.  ops receives 7 ints per
 * transfer {kind 0 copy/1 send/2 recv, src brick, dst brick, first entry in src's send list, first
 * entry in dst's receive area, entries, peer process}; returns the number of transfers. */
int         mdb_dd_schedule(int gx, int gy, int gz, int perx, int pery, int perz, int nprocs, int proc,
                            const int* cnt, int max_ops, int* ops);
/* p = parameters of the WHOLE box (nx,ny,nz multiples of gx,gy,gz).  nccl_id may be NULL iff nprocs == 1 */
mdb_dd*     mdb_dd_create(const mdb_params* p, int gx, int gy, int gz, int nprocs, int proc,
                          const void* nccl_id, int device);
/* the same brick grid for the CLUSTERPAIR scheme (cluster_n = 4 or 8, full lists, generated lattices): ghost j-CLUSTERS come
 * from the neighbor bricks as whole tiles (setupPbc / updatePbcCPU, clusterpair/pbc.c:183-323, 45-114), atoms migrate at every
 * rebuild (pbc.c:117-144); NCCL send/recv between processes, device copies inside one.  Every mdb_dd_* call below applies
 * (mdb_dd_setEam and mdb_dd_getNeighborTags answer with an error; options are mdb_cp_setOption's);
 * mdb_dd_getAtoms returns the atom arrays as of the last updateSingleAtoms; mdb_dd_getCounts v[2] counts ghost clusters. */
mdb_dd*     mdb_dd_create_cp(const mdb_params* p, int cluster_n, int gx, int gy, int gz, int nprocs, int proc,
                             const void* nccl_id, int device);
void        mdb_dd_destroy(mdb_dd* d);
int         mdb_dd_setStream(mdb_dd* d, void* cuda_stream);
int         mdb_dd_sync(mdb_dd* d);
long long   mdb_dd_createAtom(mdb_dd* d);               /* createAtom per brick; returns global Natoms */
/* what the file readers (atom.c:199-562) hand over, decomposed: n atoms lying in THIS process's bricks,
 * HOST SoA buffers of `precision` reals in the global frame, global tags; vx..vz may be NULL */
int         mdb_dd_setAtoms(mdb_dd* d, long long n, const int* tags, const void* x, const void* y,
                            const void* z, const void* vx, const void* vy, const void* vz);
int         mdb_dd_setEam(mdb_dd* d, int nrho, double drho, int nr, double dr, double cut, double mass,
                          const double* frho, const double* zr, const double* rhor);
int         mdb_dd_setup(mdb_dd* d, int adjust);        /* main.c:58-72 over all bricks */
int         mdb_dd_reneighbour(mdb_dd* d);              /* main.c:76-95: migrate, ghosts, lists */
int         mdb_dd_run(mdb_dd* d, int nsteps, double* thermo_out, int max_records, int* nrecords,
                       double* timers);                 /* main.c:244-288, thermo summed over all bricks */
int         mdb_dd_computeThermo(mdb_dd* d, double* T, double* P);
/* v = {global Natoms, local atoms of this process, ghosts of this process, largest maxneighs, bricks here} */
int         mdb_dd_getCounts(mdb_dd* d, long long v[5]);
/* local atoms of this process (SoA host buffers of `precision` reals; positions in the global frame)
 * and their global tags (may be NULL) */
int         mdb_dd_getAtoms(mdb_dd* d, int which, int* tags, void* x, void* y, void* z);
/* neighbor lists of the local atoms with every entry translated to its global tag (images map to the
 * tag of their source atom): tags[n], numneigh[n], rows[n*stride] */
int         mdb_dd_getNeighborTags(mdb_dd* d, int* tags, int* numneigh, int* rows, int stride);
int         mdb_dd_saveState(mdb_dd* d);
int         mdb_dd_restoreState(mdb_dd* d);
/* every mdb_setOption name (applied to all bricks of this process), plus "halo_push" (default 1: per-step ghost
 * positions by peer stores into the IPC-mapped arrays of the neighbor GPUs; 0: NCCL send/recv).  Same value on every process. */
int         mdb_dd_setOption(mdb_dd* d, const char* name, double value);
int         mdb_dd_setTiming(mdb_dd* d, int on);
int         mdb_dd_getKernelStats(mdb_dd* d, double* force_ms, long long* force_launches, double* neigh_ms,
                                  long long* neigh_launches, long long* total_launches, double* comm_ms);
int         mdb_dd_resetKernelStats(mdb_dd* d);

/* ---- clusterpair scheme (reference src/clusterpair/, OPT_SCHEME=clusterpair) --------------------- */
/* GROMACS-style M x N cluster pairs with M = 4 i-atoms per cluster (force.h:48) and cluster_n = 4 or 8
 * j-atoms per cluster (the reference fixes N at compile time through VECTOR_WIDTH, force.h:50-58; here
 * it is chosen at create time).  The driver of this scheme (clusterpair/main.c:40-93, 225-300) calls
 * plain functions buildClusters / defineJClusters / binClusters / updateSingleAtoms (neighbor.h:42-50)
 * next to the function pointers computeForce (force.h), buildNeighbor (neighbor.h:40),
 * initialIntegrate / finalIntegrate (integrate.h:14), updatePbc / updateAtomsPbc (pbc.h); one entry
 * point per function below.  Cluster data use the reference's layout (force.h:62-91): tile t of
 * cluster_n atoms stored as [x0..x(N-1) | y0.. | z0..]; tile t is j-cluster t and, for N = 8, also
 * the i-clusters 2t (lanes 0-3) and 2t+1 (lanes 4-7); ghost j-clusters follow the local ones and the
 * last tile (dummy_cj) is all +infinity.  p->layout applies to atom positions only, velocities are
 * always SoA (clusterpair/atom.h:66-92).  pbc_x/y/z are ignored like in the reference (pbc.c:183-323).  With
 * p->from_input the box lengths are xhi - xlo ... (what the readers hand to setupNeighbor, neighbor.c:78-82) and the
 * box is treated as [0, length) like the reference does; the atoms then come from mdb_cp_setAtoms. */
typedef struct mdb_cp mdb_cp;
mdb_cp*     mdb_cp_create(const mdb_params* p, int cluster_n, int device);
void        mdb_cp_destroy(mdb_cp* c);
int         mdb_cp_setStream(mdb_cp* c, void* cuda_stream);
int         mdb_cp_sync(mdb_cp* c);
/* "prune_every" (default 1000, common/parameter.c:40): pruneNeighbor period inside mdb_cp_run; "force_variant" 0 = auto
 * (full lists: lane per i atom, packed FP32 in SP; half lists: warp per i-cluster), 1 = lane per i atom scalar,
 * 2 = lane per i atom packed FP32 (SP full), 3 = warp per i-cluster / lane per j atom;
 * "sp_kernel" (SP full lists, packed FP32): 2 (default) = two lanes per i-cluster, two i atoms per lane, reciprocal =
 * MUFU.RCP (1 ulp; the reference's SP kernel uses the 14-bit _mm512_rcp14_ps), 1 = the same with a Newton step on it
 * (bit-identical to 0), 0 = lane per i atom;
 * "ghost_epilogue" (default -1 = on for domains of up to 2^19 atoms, 0 off, 1 on): the fused step's epilogue also writes the
 * atom's lanes of the ghost tiles (updatePbcCPU, pbc.c:45-114), so a step between two rebuilds is ONE launch; bit-identical;
 * "fuse_force" (default 1): inside mdb_cp_run with full lists, computeForce(n) + finalIntegrate(n) + initialIntegrate(n+1)
 * run as ONE kernel (integrate halves in the force kernel's epilogue, second cluster position array); bit-identical. */
int         mdb_cp_setOption(mdb_cp* c, const char* name, double value);
long long   mdb_cp_createAtom(mdb_cp* c);                         /* clusterpair/atom.c:49-180 */
int         mdb_cp_setAtoms(mdb_cp* c, long long n, const void* x, const void* y, const void* z,
                            const void* vx, const void* vy, const void* vz);
/* atom arrays as they are (refreshed by updateSingleAtoms only, like the reference's); which = 'x'|'v';
 * tag (may be NULL) = index each atom had when it was handed over (the reference does not track it) */
int         mdb_cp_getAtoms(mdb_cp* c, int which, void* x, void* y, void* z, int* tag);
/* v = {Natoms, Nlocal, Nghost (atoms), Nclusters_local, Nclusters_ghost, dummy_cj, maxneighs,
 *      local j-clusters (= local tiles)} */
int         mdb_cp_getCounts(mdb_cp* c, long long v[8]);
int         mdb_cp_setupThermo(mdb_cp* c);
int         mdb_cp_adjustThermo(mdb_cp* c);
int         mdb_cp_computeThermo(mdb_cp* c, double* T, double* P); /* reads the ATOM arrays, thermo.c:55-80 */
int         mdb_cp_setupNeighbor(mdb_cp* c);                      /* neighbor.c:70-172 */
int         mdb_cp_buildClusters(mdb_cp* c);                      /* neighbor.c:663-753 (+binAtoms, sortAtomsByZCoord) */
int         mdb_cp_defineJClusters(mdb_cp* c);                    /* neighbor.c:755-873 */
int         mdb_cp_setupPbc(mdb_cp* c);                           /* pbc.c:183-323 (incl. updatePbc(first)) */
int         mdb_cp_binClusters(mdb_cp* c);                        /* neighbor.c:875-1021 */
int         mdb_cp_buildNeighbor(mdb_cp* c);                      /* buildNeighborCPU, neighbor.c:262-481 */
int         mdb_cp_pruneNeighbor(mdb_cp* c);                      /* pruneNeighborCPU, neighbor.c:483-531 */
int         mdb_cp_updateSingleAtoms(mdb_cp* c);                  /* neighbor.c:1023-1049 */
int         mdb_cp_updateAtomsPbc(mdb_cp* c);                     /* updateAtomsPbcCPU, pbc.c:117-144 */
int         mdb_cp_updatePbc(mdb_cp* c, int first);               /* updatePbcCPU, pbc.c:45-114 */
double      mdb_cp_computeForce(mdb_cp* c);                       /* computeForceLJ 4xN / 2xNN / Ref, force_lj.c:47-164 */
int         mdb_cp_initialIntegrate(mdb_cp* c);                   /* integrate.c:23-44 */
int         mdb_cp_finalIntegrate(mdb_cp* c);                     /* integrate.c:46-63 */
int         mdb_cp_setup(mdb_cp* c, int adjust);                  /* clusterpair/main.c:40-76 after the atoms exist */
int         mdb_cp_reneighbour(mdb_cp* c);                        /* clusterpair/main.c:78-93 */
/* clusterpair/main.c:225-300; thermo records are taken from the atom arrays exactly when the reference
 * takes them (so an intermediate record shows the velocities of the last rebuild) */
int         mdb_cp_run(mdb_cp* c, int nsteps, double* thermo_out, int max_records, int* nrecords,
                       double* timers);
int         mdb_cp_saveState(mdb_cp* c);                          /* atom arrays x, v kept on the device */
int         mdb_cp_restoreState(mdb_cp* c);
int         mdb_cp_setTiming(mdb_cp* c, int on);
int         mdb_cp_getKernelStats(mdb_cp* c, double* force_ms, long long* force_launches, double* neigh_ms,
                                  long long* neigh_launches, long long* total_launches);
int         mdb_cp_resetKernelStats(mdb_cp* c);
/* cluster pairs listed, atom pairs inside the force cutoff (clusterpair/stats.h equivalents) */
int         mdb_cp_countPairs(mdb_cp* c, long long* cluster_pairs, long long* atom_pairs_in_cutoff);
/* parity accessors: which = 'i' (Nclusters_local entries) | 'j' (local + ghost j-clusters):
 * Cluster.natoms and the bounding box {minx,maxx,miny,maxy,minz,maxz} (atom.h:19-24) */
int         mdb_cp_getClusters(mdb_cp* c, int which, int* natoms, void* bbox);
/* cl_x / cl_v / cl_f ('x' | 'v' | 'f') as tiles x 3 x cluster_n reals; 'x': local + ghost tiles, else local */
int         mdb_cp_getClusterData(mdb_cp* c, int which, void* out);
int         mdb_cp_getClusterTags(mdb_cp* c, int* tags);          /* (local + ghost tiles) x cluster_n, -1 = padding */
int         mdb_cp_getClusterBins(mdb_cp* c, int* icluster_bin);  /* Atom.icluster_bin */
/* Neighbor.numneigh / numneigh_masked / neighbors (row-major rows of row_stride ints), neighbor.h:31-40 */
int         mdb_cp_getLists(mdb_cp* c, int* numneigh, int* numneigh_masked, int* neighbors, int row_stride);
int         mdb_cp_getGhostMap(mdb_cp* c, int* border_map, int* PBCx, int* PBCy, int* PBCz);
/* ints {nbinx,nbiny,mbinx,mbiny,mbins,mbinxlo,mbinylo,nstencil}, reals {binsizex,binsizey,bininvx,
 * bininvy,cutneighsq,cutneigh,xprd,yprd,zprd,rbb_sq}; stencil (may be NULL) nstencil ints */
int         mdb_cp_getNeighborParams(mdb_cp* c, int ints[8], double reals[10], int* stencil);

/* kernel micro-benchmark of the clusterpair scheme (reference src/clusterpair/main-stub.c:227-272, 61-122): niclusters
 * synthetic i-clusters of iclusters_natoms atoms at x = y = z = index * 1e-5 (use cutforce 1e6), j-clusters defined from
 * them, and a list of nneighs j-clusters per i-cluster in pattern MDB_STUB_SEQ / FIX / RAND, replicated nreps times;
 * masked != 0: every entry goes through the masked loop.  mdb_cp_computeForce then times the kernel alone. */
int         mdb_cp_stub(mdb_cp* c, int niclusters, int iclusters_natoms, int pattern, int nneighs, int nreps,
                        int masked, unsigned seed);

/* ---- measurement ------------------------------------------------------------------------------ */
/* FMA issue-rate micro-benchmark on `device`: dense FP32 (MDB_SP) or FP64 (MDB_DP) vector peak in
 * TFLOP/s (FMA = 2 flop).  The roofline denominator for the force kernels. */
int         mdb_measureFmaPeak(int precision, int device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* MDB200_H */
