/* main_cp.c -- MD-Bench driver, CLUSTERPAIR scheme (reference OPT_SCHEME=clusterpair) over libmdb200.
 * Same command line, parameter-file keys and report as reference src/clusterpair/main.c:95-320; setup() and
 * reneighbour() follow main.c:40-93, the time loop 225-300 (including pruneNeighbor every prune_every steps and the
 * thermo records taken from the atom arrays as they are at that moment).  Extras replacing build options of the
 * reference: --precision sp|dp (DATA_TYPE; default sp like BASELINE config 2), --cluster-n 4|8 (the reference derives
 * CLUSTER_N from VECTOR_WIDTH, force.h:50-58), --device <n>, --operators (run the loop operator by operator through
 * the reference's driver functions instead of the device-resident mdb_cp_run()). */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "mdbench.h"

enum { TOTAL = 0, NEIGH, FORCE, NUMTIMER };

static mdb_cp* ctx;

static double getTimeStamp(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + (double)ts.tv_nsec * 1.e-9;
}
#define CK(call, where)                                                                          \
    do {                                                                                         \
        if ((call) != 0) mdb_die(where);                                                         \
    } while (0)

/* the plain functions / function pointers clusterpair/main.c calls (neighbor.h:40-50, pbc.h, integrate.h:14, force.h) */
static void buildClusters(void) { CK(mdb_cp_buildClusters(ctx), "buildClusters"); }
static void defineJClusters(void) { CK(mdb_cp_defineJClusters(ctx), "defineJClusters"); }
static void setupPbcCP(void) { CK(mdb_cp_setupPbc(ctx), "setupPbc"); }
static void binClusters(void) { CK(mdb_cp_binClusters(ctx), "binClusters"); }
static void buildNeighborCP(void) { CK(mdb_cp_buildNeighbor(ctx), "buildNeighbor"); }
static void pruneNeighbor(void) { CK(mdb_cp_pruneNeighbor(ctx), "pruneNeighbor"); }
static void updateSingleAtoms(void) { CK(mdb_cp_updateSingleAtoms(ctx), "updateSingleAtoms"); }
static void updateAtomsPbcCP(void) { CK(mdb_cp_updateAtomsPbc(ctx), "updateAtomsPbc"); }
static void updatePbcCP(int first) { CK(mdb_cp_updatePbc(ctx, first), "updatePbc"); }
static void initialIntegrateCP(void) { CK(mdb_cp_initialIntegrate(ctx), "initialIntegrate"); }
static void finalIntegrateCP(void) { CK(mdb_cp_finalIntegrate(ctx), "finalIntegrate"); }
static double computeForceCP(void)
{
    double t = mdb_cp_computeForce(ctx);
    if (t < 0) mdb_die("computeForce");
    return t;
}
static void computeThermoCP(int iflag)
{
    double T, P;
    CK(mdb_cp_computeThermo(ctx, &T, &P), "computeThermo");
    fprintf(stdout, "%i\t%e\t%e\n", iflag < 0 ? -iflag - 1 : iflag, T, P);
}

static double setup(Parameter* param, int cluster_n) /* clusterpair/main.c:40-76 */
{
    double timeStart = getTimeStamp();
    Atom atom;
    initAtom(&atom);
    param->lattice = pow((4.0 / param->rho), (1.0 / 3.0));
    param->xprd = param->nx * param->lattice;
    param->yprd = param->ny * param->lattice;
    param->zprd = param->nz * param->lattice;
    if (param->input_file != NULL) { /* main.c:57-66: readAtom sets the box, positions are used as they are */
        param->layout = MDB_AOS;
        readAtom(&atom, param);
    }
    mdb_params p;
    mdb_default_params(&p);
    p.precision = param->precision; p.layout = MDB_AOS; p.force_field = MDB_FF_LJ;
    p.epsilon = param->epsilon; p.sigma = param->sigma; p.temp = param->temp; p.rho = param->rho; p.mass = param->mass;
    p.ntypes = param->ntypes; p.ntimes = param->ntimes; p.nstat = param->nstat; p.reneigh_every = param->reneigh_every;
    p.half_neigh = param->half_neigh; p.dt = param->dt; p.skin = param->skin; p.cutforce = param->cutforce;
    p.nx = param->nx; p.ny = param->ny; p.nz = param->nz;
    p.pbc_x = param->pbc_x; p.pbc_y = param->pbc_y; p.pbc_z = param->pbc_z;
    p.from_input = param->input_file != NULL;
    p.xlo = param->xlo; p.xhi = param->xhi; p.ylo = param->ylo; p.yhi = param->yhi; p.zlo = param->zlo; p.zhi = param->zhi;
    ctx = mdb_cp_create(&p, cluster_n, param->device);
    if (!ctx) mdb_die("initDevice");
    CK(mdb_cp_setOption(ctx, "prune_every", (double)param->prune_every), "initNeighbor");
    if (param->input_file == NULL) {
        if (mdb_cp_createAtom(ctx) < 0) mdb_die("createAtom");
    } else { /* the driver's reader stages positions AoS {x,y,z}* and velocities AoS too: hand velocities over as SoA */
        const size_t n = atom.Nlocal, es = param->precision == MDB_SP ? sizeof(float) : sizeof(double);
        char* v = (char*)malloc(3 * n * es);
        for (size_t i = 0; i < n; i++)
            for (int k = 0; k < 3; k++) memcpy(v + (k * n + i) * es, (char*)atom.vx + (3 * i + k) * es, es);
        if (mdb_cp_setAtoms(ctx, (long long)n, atom.x, NULL, NULL, v, v + n * es, v + 2 * n * es) != 0) mdb_die("readAtom");
        free(v);
    }
    CK(mdb_cp_setupNeighbor(ctx), "setupNeighbor");
    CK(mdb_cp_setupThermo(ctx), "setupThermo");
    if (param->input_file == NULL) CK(mdb_cp_adjustThermo(ctx), "adjustThermo");
    buildClusters();
    defineJClusters();
    setupPbcCP();
    binClusters();
    buildNeighborCP();
    return getTimeStamp() - timeStart;
}

static double reneighbour(void) /* clusterpair/main.c:78-93 */
{
    double timeStart = getTimeStamp();
    updateSingleAtoms();
    updateAtomsPbcCP();
    buildClusters();
    defineJClusters();
    setupPbcCP();
    binClusters();
    buildNeighborCP();
    CK(mdb_cp_sync(ctx), "reneighbour");
    return getTimeStamp() - timeStart;
}

int main(int argc, char** argv)
{
    double timer[NUMTIMER];
    Parameter param;
    int operators = 0, cluster_n = 4;

    initParameter(&param);
    param.precision = MDB_SP;
    for (int i = 0; i < argc; i++) {
        if ((strcmp(argv[i], "-p") == 0) || (strcmp(argv[i], "--param") == 0)) { readParameter(&param, argv[++i]); continue; }
        if ((strcmp(argv[i], "-f") == 0)) {
            if ((param.force_field = str2ff(argv[++i])) < 0) {
                fprintf(stderr, "Invalid force field!\n");
                exit(-1);
            }
            continue;
        }
        if ((strcmp(argv[i], "-i") == 0)) { param.input_file = strdup(argv[++i]); continue; }
        if ((strcmp(argv[i], "-e") == 0)) { param.eam_file = strdup(argv[++i]); continue; }
        if ((strcmp(argv[i], "-n") == 0) || (strcmp(argv[i], "--nsteps") == 0)) { param.ntimes = atoi(argv[++i]); continue; }
        if ((strcmp(argv[i], "-nx") == 0)) { param.nx = atoi(argv[++i]); continue; }
        if ((strcmp(argv[i], "-ny") == 0)) { param.ny = atoi(argv[++i]); continue; }
        if ((strcmp(argv[i], "-nz") == 0)) { param.nz = atoi(argv[++i]); continue; }
        if ((strcmp(argv[i], "-half") == 0)) { param.half_neigh = atoi(argv[++i]); continue; }
        if ((strcmp(argv[i], "-m") == 0) || (strcmp(argv[i], "--mass") == 0)) { param.mass = atof(argv[++i]); continue; }
        if ((strcmp(argv[i], "-r") == 0) || (strcmp(argv[i], "--radius") == 0)) { param.cutforce = atof(argv[++i]); continue; }
        if ((strcmp(argv[i], "-s") == 0) || (strcmp(argv[i], "--skin") == 0)) { param.skin = atof(argv[++i]); continue; }
        if ((strcmp(argv[i], "--freq") == 0)) { param.proc_freq = atof(argv[++i]); continue; }
        if ((strcmp(argv[i], "--vtk") == 0)) { param.vtk_file = strdup(argv[++i]); continue; }
        if ((strcmp(argv[i], "--xtc") == 0)) { param.xtc_file = strdup(argv[++i]); continue; }
        if ((strcmp(argv[i], "--precision") == 0)) { param.precision = strcmp(argv[++i], "sp") == 0 ? MDB_SP : MDB_DP; continue; }
        if ((strcmp(argv[i], "--cluster-n") == 0)) { cluster_n = atoi(argv[++i]); continue; }
        if ((strcmp(argv[i], "--device") == 0)) { param.device = atoi(argv[++i]); continue; }
        if ((strcmp(argv[i], "--operators") == 0)) { operators = 1; continue; }
        if ((strcmp(argv[i], "-h") == 0) || (strcmp(argv[i], "--help") == 0)) {
            printf("MD Bench: A minimalistic re-implementation of miniMD (B200 build, clusterpair scheme)\n");
            printf(HLINE);
            printf("-p / --param <string>:      file to read parameters from (can be specified more than once)\n");
            printf("-f <string>:                force field (lj), default lj\n");
            printf("-i <string>:                input file with atom positions (dump)\n");
            printf("-n / --nsteps <int>:        set number of timesteps for simulation\n");
            printf("-nx/-ny/-nz <int>:          set linear dimension of systembox in x/y/z direction\n");
            printf("-half <int>:                use half (1) or full (0) neighbor lists\n");
            printf("-m / --mass <real>:         set mass of atoms\n");
            printf("-r / --radius <real>:       set cutoff radius\n");
            printf("-s / --skin <real>:         set skin (verlet buffer)\n");
            printf("--freq <real>:              processor frequency (GHz)\n");
            printf("--vtk / --xtc <string>:     trajectory output (not supported in this build)\n");
            printf("--precision sp|dp           floating-point precision (reference: DATA_TYPE), default sp\n");
            printf("--cluster-n 4|8             j-cluster size N of the 4 x N scheme (reference: from VECTOR_WIDTH)\n");
            printf("--device <int>              CUDA device ordinal\n");
            printf("--operators                 run the loop through the driver functions, operator by operator\n");
            printf(HLINE);
            exit(EXIT_SUCCESS);
        }
    }
    if (param.force_field != FF_LJ) {
        fprintf(stderr, "Error: the clusterpair scheme has only the LJ kernels\n");
        exit(EXIT_FAILURE);
    }
    param.cutneigh = param.cutforce + param.skin; /* clusterpair/main.c:216 */
    setup(&param, cluster_n);
    printParameter(&param);
    printf("\tKernel: CUDA sm_100a %dx%d, %s neighbor lists\n", 4, cluster_n, param.half_neigh ? "half" : "full");
    printf(HLINE);
    printf("step\ttemp\t\tpressure\n");

    if (operators) {
        computeThermoCP(0);
        timer[FORCE] = computeForceCP();
        timer[NEIGH] = 0.0;
        timer[TOTAL] = getTimeStamp();
        for (int n = 0; n < param.ntimes; n++) {
            initialIntegrateCP();
            if ((n + 1) % param.reneigh_every) {
                if (!((n + 1) % param.prune_every)) pruneNeighbor();
                updatePbcCP(0);
            } else {
                timer[NEIGH] += reneighbour();
            }
            timer[FORCE] += computeForceCP();
            finalIntegrateCP();
            if (!((n + 1) % param.nstat) && (n + 1) < param.ntimes) computeThermoCP(n + 1);
        }
        CK(mdb_cp_sync(ctx), "run");
        timer[TOTAL] = getTimeStamp() - timer[TOTAL];
        updateSingleAtoms();
        computeThermoCP(-1 - param.ntimes);
    } else {
        int maxrec = param.ntimes / (param.nstat > 0 ? param.nstat : 1) + 4, nrec = 0;
        double* rec = (double*)malloc(3 * maxrec * sizeof(double));
        double tm[3];
        mdb_cp_setTiming(ctx, getenv("MDB_PHASE_TIMERS") != NULL);
        CK(mdb_cp_run(ctx, param.ntimes, rec, maxrec, &nrec, tm), "run");
        for (int r = 0; r < nrec; r++) fprintf(stdout, "%i\t%e\t%e\n", (int)rec[3 * r], rec[3 * r + 1], rec[3 * r + 2]);
        timer[TOTAL] = tm[0]; timer[FORCE] = tm[1]; timer[NEIGH] = tm[2];
        free(rec);
    }
    long long v[8];
    CK(mdb_cp_getCounts(ctx, v), "getCounts");
    printf(HLINE);
    printf("System: %d atoms %d ghost atoms, Steps: %d\n", (int)v[0], (int)v[2], param.ntimes);
    printf("TOTAL %.2fs FORCE %.2fs NEIGH %.2fs REST %.2fs\n", timer[TOTAL], timer[FORCE], timer[NEIGH],
        timer[TOTAL] - timer[FORCE] - timer[NEIGH]);
    printf(HLINE);
    printf("Performance: %.2f million atom updates per second\n", 1e-6 * (double)v[0] * param.ntimes / timer[TOTAL]);
    mdb_cp_destroy(ctx);
    return EXIT_SUCCESS;
}
