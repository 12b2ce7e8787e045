/* main_stub_cp.c -- the reference's clusterpair force-kernel micro-benchmark (src/clusterpair/main-stub.c) over libmdb200:
 * same options (-f -p seq|fix|rand -m <masked> -n -ni -na -nn -nr --freq --csv), same synthetic clusters (atoms at
 * x = y = z = index * 1e-5, cutoff 1e6) and list patterns, same report lines.  Clusters and lists are generated on the
 * device (mdb_cp_stub); each timestep is one mdb_cp_computeForce call timed by CUDA events.
 * Extras: --precision sp|dp (default sp), --cluster-n 4|8, -half 0|1, --device n. */
#include <stdlib.h>
#include <string.h>

#include "mdbench.h"

int main(int argc, char** argv)
{
    Parameter param;
    const char* pattern_str = "seq";
    int pattern = MDB_STUB_SEQ, niclusters = 256, iclusters_natoms = 4, nneighs = 9, nreps = 1, masked = 0, csv = 0, cluster_n = 4;
    initParameter(&param);
    param.precision = MDB_SP;
    param.ntimes    = 200;
    param.cutforce  = 1000000.0; /* main-stub.c:46-47 */
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
        if (strcmp(argv[i], "-m") == 0) { masked = atoi(argv[++i]); continue; }
        if (strcmp(argv[i], "-n") == 0 || strcmp(argv[i], "--nsteps") == 0) { param.ntimes = atoi(argv[++i]); continue; }
        if (strcmp(argv[i], "-ni") == 0) { niclusters = atoi(argv[++i]); continue; }
        if (strcmp(argv[i], "-na") == 0) { iclusters_natoms = atoi(argv[++i]); continue; }
        if (strcmp(argv[i], "-nn") == 0) { nneighs = atoi(argv[++i]); continue; }
        if (strcmp(argv[i], "-nr") == 0) { nreps = atoi(argv[++i]); continue; }
        if (strcmp(argv[i], "-half") == 0) { param.half_neigh = atoi(argv[++i]); continue; }
        if (strcmp(argv[i], "--freq") == 0) { param.proc_freq = atof(argv[++i]); continue; }
        if (strcmp(argv[i], "--csv") == 0) { csv = 1; continue; }
        if (strcmp(argv[i], "--precision") == 0) { param.precision = strcmp(argv[++i], "sp") == 0 ? MDB_SP : MDB_DP; continue; }
        if (strcmp(argv[i], "--cluster-n") == 0) { cluster_n = atoi(argv[++i]); continue; }
        if (strcmp(argv[i], "--device") == 0) { param.device = atoi(argv[++i]); continue; }
        if (strcmp(argv[i], "-h") == 0 || strcmp(argv[i], "--help") == 0) {
            printf("MD Bench: A minimalistic re-implementation of miniMD (B200 build, clusterpair kernel micro-benchmark)\n");
            printf(HLINE);
            printf("-f <string>:          force field (lj), default lj\n");
            printf("-p <string>:          pattern for data accesses (seq, fix or rand)\n");
            printf("-m <int>:             use the masked loop for every list entry (default 0)\n");
            printf("-n / --nsteps <int>:  number of timesteps for simulation\n");
            printf("-ni <int>:            number of i-clusters (default 256)\n");
            printf("-na <int>:            number of atoms per i-cluster (default 4)\n");
            printf("-nn <int>:            number of j-cluster neighbors per i-cluster (default 9)\n");
            printf("-nr <int>:            number of times neighbor lists should be replicated (default 1)\n");
            printf("--freq <real>:        set clock frequency (GHz) and display average cycles per atom and neighbors\n");
            printf("--csv:                set output as CSV style\n");
            printf("--precision sp|dp, --cluster-n 4|8, -half 0|1, --device <int>\n");
            printf(HLINE);
            exit(EXIT_SUCCESS);
        }
    }
    if (param.force_field != FF_LJ) { fprintf(stderr, "Error: the clusterpair scheme has only the LJ kernels\n"); exit(-1); }

    mdb_params p;
    mdb_default_params(&p);
    p.precision = param.precision; p.epsilon = param.epsilon; p.sigma = 1.0; p.cutforce = param.cutforce; p.skin = 0.0;
    p.half_neigh = param.half_neigh; p.nx = p.ny = p.nz = 1;
    mdb_cp* ctx = mdb_cp_create(&p, cluster_n, param.device);
    if (!ctx) mdb_die("initDevice");
    if (mdb_cp_stub(ctx, niclusters, iclusters_natoms, pattern, nneighs, nreps, masked, 12345u) != 0) mdb_die("createNeighbors");
    long long v[8];
    if (mdb_cp_getCounts(ctx, v) != 0) mdb_die("getCounts");
    const long long nlocal = v[1];
    const size_t es = param.precision == MDB_SP ? sizeof(float) : sizeof(double);
    const double estim_atom_volume      = (double)nlocal * 3 * es;
    const double estim_neighbors_volume = (double)nlocal * (nneighs + 2) * sizeof(int);
    const double estim_volume           = (double)nlocal * 6 * es + estim_neighbors_volume;
    if (!csv) {
        printf("Kernel: CUDA sm_100a, MxN: %dx%d\n", 4, cluster_n);
        printf("Floating-point precision: %s\n", param.precision == MDB_SP ? "single" : "double");
        printf("Pattern: %s\n", pattern_str);
        printf("Number of timesteps: %d\n", param.ntimes);
        printf("Number of i-clusters: %d\n", niclusters);
        printf("Number of atoms per i-cluster: %d\n", iclusters_natoms);
        printf("Number of j-cluster neighbors per i-cluster: %d\n", nneighs);
        printf("Number of times to replicate neighbor lists: %d\n", nreps);
        printf("Estimated total data volume (kB): %.4f\n", estim_volume / 1000.0);
        printf("Estimated atom data volume (kB): %.4f\n", estim_atom_volume / 1000.0);
        printf("Estimated neighborlist data volume (kB): %.4f\n", estim_neighbors_volume / 1000.0);
    }
    double T_accum = 0.0;
    for (int i = 0; i < param.ntimes; i++) {
        const double t = mdb_cp_computeForce(ctx);
        if (t < 0) mdb_die("computeForce");
        T_accum += t;
    }
    const double freq_hz               = param.proc_freq * 1.e9;
    const double atoms_updates_per_sec = (double)nlocal / T_accum * (double)param.ntimes;
    const double cycles_per_atom       = T_accum / (double)nlocal / (double)param.ntimes * freq_hz;
    const double cycles_per_neigh      = cycles_per_atom / (double)nneighs;
    if (!csv) {
        printf("Total time: %.4f, Mega atom updates/s: %.4f\n", T_accum, atoms_updates_per_sec / 1.e6);
        if (param.proc_freq > 0.0) printf("Cycles per atom: %.4f, Cycles per neighbor: %.4f\n", cycles_per_atom, cycles_per_neigh);
    } else {
        printf("steps,pattern,niclusters,iclusters_natoms,nneighs,nreps,total vol.(kB),atoms vol.(kB),neigh vol.(kB),time(s),atom upds/s(M)");
        if (param.proc_freq > 0.0) printf(",cy/atom,cy/neigh");
        printf("\n");
        printf("%d,%s,%d,%d,%d,%d,%.4f,%.4f,%.4f,%.4f,%.4f", param.ntimes, pattern_str, niclusters, iclusters_natoms, nneighs, nreps,
            estim_volume / 1.e3, estim_atom_volume / 1.e3, estim_neighbors_volume / 1.e3, T_accum, atoms_updates_per_sec / 1.e6);
        if (param.proc_freq > 0.0) printf(",%.4f,%.4f", cycles_per_atom, cycles_per_neigh);
        printf("\n");
    }
    mdb_cp_destroy(ctx);
    return EXIT_SUCCESS;
}
