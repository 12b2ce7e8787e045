/* main.c -- MD-Bench driver (verletlist scheme) over libmdb200.
 * Same command line, parameter-file keys and report as reference src/verletlist/main.c:129-344;
 * setup() and the time loop follow main.c:36-95 and 244-288.  Two extras replace build options of
 * the reference: --precision sp|dp (DATA_TYPE), --layout aos|soa (DATA_LAYOUT), --sort (SORT_ATOMS),
 * --device <n>, and --operators runs the time loop operator by operator through the function
 * pointers (like the reference) instead of the device-resident mdb_run(). */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "mdbench.h"

enum { TOTAL = 0, NEIGH, FORCE, NUMTIMER };

static double getTimeStamp(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + (double)ts.tv_nsec * 1.e-9;
}

static Funcfl funcfl;

static double setup(Parameter* param, Atom* atom, Neighbor* neighbor, Stats* stats)
{
    double timeStart = getTimeStamp();
    initAtom(atom);
    memset(stats, 0, sizeof *stats);
    memset(neighbor, 0, sizeof *neighbor);
    neighbor->maxneighs  = 100;
    neighbor->half_neigh = param->half_neigh;
    if (param->force_field == FF_EAM) { /* initEam, eam_utils.c:22-40: overrides come from the potential file */
        readEamFile(&funcfl, param->eam_file);
        param->mass = funcfl.mass; param->cutforce = funcfl.cut; param->cutneigh = param->cutforce + 1.0;
        param->temp = 600.0; param->dt = 0.001; param->rho = 0.07041125;
        param->dtforce = 0.5 * param->dt / param->mass;
    }
    param->lattice = pow((4.0 / param->rho), (1.0 / 3.0));
    param->xprd = param->nx * param->lattice;
    param->yprd = param->ny * param->lattice;
    param->zprd = param->nz * param->lattice;
    if (param->input_file != NULL) readAtom(atom, param); /* sets the box from the file */
    initDevice(atom, param);
    if (param->force_field == FF_EAM &&
        mdb_setEam(atom->d_atom, funcfl.nrho, funcfl.drho, funcfl.nr, funcfl.dr, funcfl.cut, funcfl.mass,
            funcfl.frho, funcfl.zr, funcfl.rhor) != 0)
        mdb_die("initEam");
    if (param->input_file == NULL) {
        createAtom(atom, param);
    } else if (mdb_setAtoms(atom->d_atom, atom->Nlocal, atom->x, atom->y, atom->z, atom->vx, atom->vy, atom->vz, NULL) != 0) {
        mdb_die("readAtom");
    }
    setupNeighbor(param, atom);
    setupThermo(param, atom);
    if (param->input_file == NULL) adjustThermo(param, atom);
    if (param->sort_atoms && mdb_setOption(atom->d_atom, "sort_atoms", 1.0) != 0) mdb_die("sortAtom");
    setupPbc(atom, param);
    updatePbc(atom, param, true);
    buildNeighbor(atom, neighbor);
    initForce(param);
    return getTimeStamp() - timeStart;
}

static double reneighbour(int n, Parameter* param, Atom* atom, Neighbor* neighbor)
{
    double timeStart = getTimeStamp();
    updateAtomsPbc(atom, param, true);
    setupPbc(atom, param);
    updatePbc(atom, param, true);
    buildNeighbor(atom, neighbor);
    return getTimeStamp() - timeStart;
}

int main(int argc, char** argv)
{
    double timer[NUMTIMER];
    Atom atom;
    Neighbor neighbor;
    Stats stats;
    Parameter param;
    int operators = 0;

    initParameter(&param);
    for (int i = 0; i < argc; i++) {
        if ((strcmp(argv[i], "-p") == 0) || strcmp(argv[i], "--params") == 0) { readParameter(&param, argv[++i]); continue; }
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
        if ((strcmp(argv[i], "-r") == 0) || (strcmp(argv[i], "--radius") == 0)) { param.cutforce = atof(argv[++i]); continue; }
        if ((strcmp(argv[i], "-s") == 0) || (strcmp(argv[i], "--skin") == 0)) { param.skin = atof(argv[++i]); continue; }
        if ((strcmp(argv[i], "--freq") == 0)) { param.proc_freq = atof(argv[++i]); continue; }
        if ((strcmp(argv[i], "--vtk") == 0)) { param.vtk_file = strdup(argv[++i]); continue; }
        if ((strcmp(argv[i], "-w") == 0)) { param.write_atom_file = strdup(argv[++i]); continue; }
        if ((strcmp(argv[i], "--precision") == 0)) { param.precision = strcmp(argv[++i], "sp") == 0 ? MDB_SP : MDB_DP; continue; }
        if ((strcmp(argv[i], "--layout") == 0)) { param.layout = strcmp(argv[++i], "soa") == 0 ? MDB_SOA : MDB_AOS; continue; }
        if ((strcmp(argv[i], "--device") == 0)) { param.device = atoi(argv[++i]); continue; }
        if ((strcmp(argv[i], "--sort") == 0)) { param.sort_atoms = 1; continue; }
        if ((strcmp(argv[i], "--operators") == 0)) { operators = 1; continue; }
        if ((strcmp(argv[i], "-h") == 0) || (strcmp(argv[i], "--help") == 0)) {
            printf("MD Bench: A performance-oriented prototyping harness for MD algorithms (B200 build)\n");
            printf(HLINE);
            printf("-p / --params <string>:     file to read parameters from (can be specified more than once)\n");
            printf("-f <string>:                force field (lj or eam), default lj\n");
            printf("-i <string>:                input file with atom positions (dump)\n");
            printf("-e <string>:                input file for EAM\n");
            printf("-n / --nsteps <int>:        set number of timesteps for simulation\n");
            printf("-nx/-ny/-nz <int>:          set linear dimension of systembox in x/y/z direction\n");
            printf("-half <int>:                use half (1) or full (0) neighbor lists\n");
            printf("-r / --radius <real>:       set cutoff radius\n");
            printf("-s / --skin <real>:         set skin (verlet buffer)\n");
            printf("-w <file>:                  write input atoms to file\n");
            printf("--freq <real>:              processor frequency (GHz)\n");
            printf("--vtk <string>:             VTK file for visualization (not supported in this build)\n");
            printf("--precision sp|dp           floating-point precision (reference: DATA_TYPE), default dp\n");
            printf("--layout aos|soa            host data layout (reference: DATA_LAYOUT), default aos\n");
            printf("--sort                      re-sort atoms at every rebuild (reference: SORT_ATOMS)\n");
            printf("--device <int>              CUDA device ordinal\n");
            printf("--operators                 run the loop through the operator function pointers\n");
            printf(HLINE);
            exit(EXIT_SUCCESS);
        }
    }

    param.cutneigh = param.cutforce + param.skin;
    setup(&param, &atom, &neighbor, &stats);
    printParameter(&param);
    printf(HLINE);
    printf("step\ttemp\t\tpressure\n");

    if (param.write_atom_file != NULL) {
        const size_t es = param.precision == MDB_SP ? sizeof(float) : sizeof(double);
        const size_t n  = atom.Nlocal;
        if (param.layout == MDB_AOS) {
            atom.x = realloc(atom.x, 3 * n * es); atom.vx = realloc(atom.vx, 3 * n * es);
        } else {
            void** a[] = { &atom.x, &atom.y, &atom.z, &atom.vx, &atom.vy, &atom.vz };
            for (int k = 0; k < 6; k++) *a[k] = realloc(*a[k], n * es);
        }
        if (mdb_getAtoms(atom.d_atom, 'x', 0, atom.x, atom.y, atom.z) || mdb_getAtoms(atom.d_atom, 'v', 0, atom.vx, atom.vy, atom.vz))
            mdb_die("writeAtom");
        writeAtom(&atom, &param);
    }

    if (operators) {
        /* MDB_LAZY_OPS=1: keep this loop, let the library fuse computeForce + finalIntegrate + initialIntegrate (mdb200.h) */
        const bool lazy = getenv("MDB_LAZY_OPS") != NULL, phase_timers = getenv("MDB_PHASE_TIMERS") != NULL;
        if (lazy && mdb_setOption(atom.d_atom, "lazy_ops", 1.0) != 0) mdb_die("lazy_ops");
        /* with lazy_ops a deferred computeForce returns 0 s: the force time then comes from the library's CUDA-event
         * timers (MDB_PHASE_TIMERS=1, one event pair and sync per launch), else it stays inside REST and the report says so */
        if (lazy && phase_timers) { mdb_setTiming(atom.d_atom, 1); mdb_resetKernelStats(atom.d_atom); }
        computeThermo(0, &param, &atom);
        timer[FORCE] = computeForce(&param, &atom, &neighbor, &stats);
        timer[NEIGH] = 0.0;
        timer[TOTAL] = getTimeStamp();
        for (int n = 0; n < param.ntimes; n++) {
            bool reneigh = (n + 1) % param.reneigh_every == 0;
            initialIntegrate(reneigh, &param, &atom);
            if (reneigh) timer[NEIGH] += reneighbour(n, &param, &atom, &neighbor);
            else updatePbc(&atom, &param, false);
            timer[FORCE] += computeForce(&param, &atom, &neighbor, &stats);
            finalIntegrate(reneigh, &param, &atom);
            if (!((n + 1) % param.nstat) && (n + 1) < param.ntimes) computeThermo(n + 1, &param, &atom);
        }
        mdb_sync(atom.d_atom);
        timer[TOTAL] = getTimeStamp() - timer[TOTAL];
        if (lazy) {
            double fms = 0.0, nms = 0.0;
            long long fl = 0, nl = 0, tl = 0;
            if (phase_timers && mdb_getKernelStats(atom.d_atom, &fms, &fl, &nms, &nl, &tl) == 0) timer[FORCE] = fms * 1e-3;
            else printf("lazy_ops: the force kernels run deferred, their time is part of REST (MDB_PHASE_TIMERS=1 times them)\n");
        }
        computeThermo(-1, &param, &atom);
    } else {
        /* the whole loop stays on the device; thermo records come back at the end */
        int maxrec = param.ntimes / (param.nstat > 0 ? param.nstat : 1) + 4, nrec = 0;
        double* rec = (double*)malloc(3 * maxrec * sizeof(double));
        double tm[3];
        mdb_setTiming(atom.d_atom, getenv("MDB_PHASE_TIMERS") != NULL);
        if (mdb_run(atom.d_atom, param.ntimes, rec, maxrec, &nrec, tm) != 0) mdb_die("run");
        for (int r = 0; r < nrec; r++) fprintf(stdout, "%i\t%e\t%e\n", (int)rec[3 * r], rec[3 * r + 1], rec[3 * r + 2]);
        timer[TOTAL] = tm[0]; timer[FORCE] = tm[1]; timer[NEIGH] = tm[2];
        free(rec);
    }
    long long ng;
    mdb_getCounts(atom.d_atom, NULL, NULL, &ng, NULL, NULL);
    atom.Nghost = (int)ng;

    printf(HLINE);
    printf("System: %d atoms %d ghost atoms, Steps: %d\n", atom.Natoms, atom.Nghost, param.ntimes);
    printf("TOTAL %.2fs FORCE %.2fs NEIGH %.2fs REST %.2fs\n", timer[TOTAL], timer[FORCE], timer[NEIGH],
        timer[TOTAL] - timer[FORCE] - timer[NEIGH]);
    printf(HLINE);
    printf("Performance: %.2f million atom updates per second\n", 1e-6 * (double)atom.Natoms * param.ntimes / timer[TOTAL]);
    mdb_destroy(atom.d_atom);
    return EXIT_SUCCESS;
}
