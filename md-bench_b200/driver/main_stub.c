/* main_stub.c -- the reference's force-kernel micro-benchmark (src/verletlist/main-stub.c) over libmdb200: same options
 * (-f -p seq|fix|rand -n -na -nn -nr --freq --csv), same synthetic atoms (x = y = z = i * 1e-5, cutoff 1e6: every listed
 * pair is inside) and neighbor-list patterns, same report lines.  The list is generated on the device
 * (mdb_stubNeighbors); each timestep is one mdb_computeForceLJ{Full,Half}Neigh call timed by CUDA events.
 * Extras: --precision sp|dp, --device n.  -f eam is not available here (LJ kernels only, like the driver's default). */
#include <stdlib.h>
#include <string.h>

#include "mdbench.h"

int main(int argc, char** argv)
{
    Parameter param;
    const char* pattern_str = "seq";
    int pattern = MDB_STUB_SEQ, natoms = 256, nneighs = 76, nreps = 1, csv = 0;
    initParameter(&param);
    param.ntimes = 200;
    param.cutforce = 1000000.0; /* main-stub.c:46-47 */
    param.skin = 0.0;
    param.proc_freq = 2.4;
    for (int i = 1; i < argc; i++) {
        if (strcmp(argv[i], "-f") == 0) {
            if ((param.force_field = str2ff(argv[++i])) < 0) { fprintf(stderr, "Invalid force field!\n"); exit(-1); }
            continue;
        }
        if (strcmp(argv[i], "-p") == 0) {
            pattern_str = argv[++i];
            if (strncmp(pattern_str, "seq", 3) == 0) pattern = MDB_STUB_SEQ;
            else if (strncmp(pattern_str, "fix", 3) == 0) pattern = MDB_STUB_FIX;
            else if (strncmp(pattern_str, "rand", 3) == 0) pattern = MDB_STUB_RAND;
            else { fprintf(stderr, "Invalid pattern!\n"); exit(-1); }
            continue;
        }
        if (strcmp(argv[i], "-n") == 0 || strcmp(argv[i], "--nsteps") == 0) { param.ntimes = atoi(argv[++i]); continue; }
        if (strcmp(argv[i], "-na") == 0) { natoms = atoi(argv[++i]); continue; }
        if (strcmp(argv[i], "-nn") == 0) { nneighs = atoi(argv[++i]); continue; }
        if (strcmp(argv[i], "-nr") == 0) { nreps = atoi(argv[++i]); continue; }
        if (strcmp(argv[i], "-half") == 0) { param.half_neigh = atoi(argv[++i]); continue; }
        if (strcmp(argv[i], "--freq") == 0) { param.proc_freq = atof(argv[++i]); continue; }
        if (strcmp(argv[i], "--csv") == 0) { csv = 1; continue; }
        if (strcmp(argv[i], "--precision") == 0) { param.precision = strcmp(argv[++i], "sp") == 0 ? MDB_SP : MDB_DP; continue; }
        if (strcmp(argv[i], "--device") == 0) { param.device = atoi(argv[++i]); continue; }
        if (strcmp(argv[i], "-h") == 0 || strcmp(argv[i], "--help") == 0) {
            printf("MD Bench: A minimalistic re-implementation of miniMD (B200 build, kernel micro-benchmark)\n");
            printf(HLINE);
            printf("-f <string>:          force field (lj), default lj\n");
            printf("-p <string>:          pattern for data accesses (seq, fix or rand)\n");
            printf("-n / --nsteps <int>:  number of timesteps for simulation\n");
            printf("-na <int>:            number of atoms (default 256)\n");
            printf("-nn <int>:            number of neighbors per atom (default 76)\n");
            printf("-nr <int>:            number of times neighbor lists should be replicated (default 1)\n");
            printf("--freq <real>:        set clock frequency (GHz) and display average cycles per atom and neighbors\n");
            printf("--csv:                set output as CSV style\n");
            printf("--precision sp|dp, --device <int>, -half 0|1\n");
            printf(HLINE);
            exit(EXIT_SUCCESS);
        }
    }
    if (param.force_field != FF_LJ) { fprintf(stderr, "Error: the kernel micro-benchmark of this build has the LJ kernels only\n"); exit(-1); }

    mdb_params p;
    mdb_default_params(&p);
    p.precision = param.precision; p.layout = MDB_SOA; p.epsilon = param.epsilon; p.sigma = 1.0; /* sigma6 = 1 */
    p.cutforce = param.cutforce; p.skin = 0.0; p.half_neigh = param.half_neigh; p.nx = p.ny = p.nz = 1;
    mdb_ctx* ctx = mdb_create(&p, param.device);
    if (!ctx) mdb_die("initDevice");
    const size_t es = param.precision == MDB_SP ? sizeof(float) : sizeof(double);
    void* xs = malloc((size_t)natoms * es);
    for (int i = 0; i < natoms; i++) { /* main-stub.c:245-247 */
        if (param.precision == MDB_SP) ((float*)xs)[i] = (float)i * 0.00001f;
        else ((double*)xs)[i] = (double)i * 0.00001;
    }
    if (mdb_setAtoms(ctx, natoms, xs, xs, xs, NULL, NULL, NULL, NULL) != 0) mdb_die("createAtoms");
    free(xs);

    const double estim_atom_volume      = (double)natoms * 3 * es;
    const double estim_neighbors_volume = (double)natoms * (nneighs + 2) * sizeof(int);
    const double estim_volume           = (double)natoms * 6 * es + estim_neighbors_volume;
    if (!csv) {
        printf("Pattern: %s\n", pattern_str);
        printf("Number of timesteps: %d\n", param.ntimes);
        printf("Number of atoms: %d\n", natoms);
        printf("Number of neighbors per atom: %d\n", nneighs);
        printf("Number of times to replicate neighbor lists: %d\n", nreps);
        printf("Estimated total data volume (kB): %.4f\n", estim_volume / 1000.0);
        printf("Estimated atom data volume (kB): %.4f\n", estim_atom_volume / 1000.0);
        printf("Estimated neighborlist data volume (kB): %.4f\n", estim_neighbors_volume / 1000.0);
    }
    if (mdb_stubNeighbors(ctx, pattern, nneighs, nreps, 12345u) != 0) mdb_die("createNeighbors");

    double T_accum = 0.0;
    for (int i = 0; i < param.ntimes; i++) {
        const double t = param.half_neigh ? mdb_computeForceLJHalfNeigh(ctx) : mdb_computeForceLJFullNeigh(ctx);
        if (t < 0) mdb_die("computeForce");
        T_accum += t;
    }
    const double freq_hz               = param.proc_freq * 1.e9;
    const double atoms_updates_per_sec = (double)natoms / T_accum * (double)param.ntimes;
    const double cycles_per_atom       = T_accum / (double)natoms / (double)param.ntimes * freq_hz;
    const double cycles_per_neigh      = cycles_per_atom / (double)(nneighs * nreps);
    if (!csv) {
        printf("Total time: %.4f, Mega atom updates/s: %.4f\n", T_accum, atoms_updates_per_sec / 1.e6);
        if (param.proc_freq > 0.0) printf("Cycles per atom: %.4f, Cycles per neighbor: %.4f\n", cycles_per_atom, cycles_per_neigh);
    } else {
        printf("steps,pattern,natoms,nneighs,nreps,total vol.(kB),atoms vol.(kB),neigh vol.(kB),time(s),atom upds/s(M)");
        if (param.proc_freq > 0.0) printf(",cy/atom,cy/neigh");
        printf("\n");
        printf("%d,%s,%d,%d,%d,%.4f,%.4f,%.4f,%.4f,%.4f", param.ntimes, pattern_str, natoms, nneighs, nreps, estim_volume / 1.e3,
            estim_atom_volume / 1.e3, estim_neighbors_volume / 1.e3, T_accum, atoms_updates_per_sec / 1.e6);
        if (param.proc_freq > 0.0) printf(",%.4f,%.4f", cycles_per_atom, cycles_per_neigh);
        printf("\n");
    }
    if (!csv) { /* displayStatistics of the reference's default (COMPUTE_STATS) build, verletlist/stats.c:22-68 and main-stub.c:324.  The
                 * SIMD unit here is a warp over 32 ATOMS: one "SIMD iteration" is one trip of a warp's pair loop (every row of the
                 * synthetic list has nneighs * nreps entries), against VECTOR_WIDTH neighbors of one atom in the reference. */
        const double calls       = (double)param.ntimes;
        const double per_atom    = (double)nneighs * (double)nreps;
        const double force_neighs = (double)natoms * per_atom * calls;
        const double force_iters  = (double)((natoms + 31) / 32) * per_atom * calls;
        const double useful_volume = 1e-9 * ((double)natoms * calls * (double)(es * 6 + sizeof(int)) + force_neighs * (double)(es * 3 + sizeof(int)));
        printf("Statistics:\n");
        printf("\tVector width: %d, Processor frequency: %.4f GHz\n", 32, param.proc_freq);
        printf("\tAverage neighbors per atom: %.4f\n", per_atom);
        printf("\tAverage SIMD iterations per atom: %.4f\n", force_iters / ((double)natoms * calls));
        printf("\tTotal number of computed pair interactions: %.0f\n", force_neighs);
        printf("\tTotal number of SIMD iterations: %.0f\n", force_iters);
        printf("\tUseful read data volume for force computation: %.2fGB\n", useful_volume);
        if (param.proc_freq > 0.0) printf("\tCycles/SIMD iteration: %.4f\n", T_accum * freq_hz / force_iters);
    }
    mdb_destroy(ctx);
    return EXIT_SUCCESS;
}
