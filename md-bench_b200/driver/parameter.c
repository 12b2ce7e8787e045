/* parameter.c -- run parameters: defaults, "key value" parameter files, report block.
 * Behaviour follows reference src/common/parameter.c:16-51 (defaults), 53-122 (file format: one
 * "key value" pair per line, '#' starts a comment, keys are matched by prefix), 124-187 (report). */
#include <stdlib.h>
#include <string.h>

#include "mdbench.h"

int str2ff(const char* s)
{
    if (strncmp(s, "lj", 2) == 0) return FF_LJ;
    if (strncmp(s, "eam", 3) == 0) return FF_EAM;
    return -1;
}
const char* ff2str(int ff) { return ff == FF_LJ ? "lj" : (ff == FF_EAM ? "eam" : "invalid"); }

void initParameter(Parameter* p)
{
    memset(p, 0, sizeof *p);
    p->force_field   = FF_LJ;
    p->epsilon       = 1.0;
    p->sigma         = 1.0;
    p->sigma6        = 1.0;
    p->rho           = 0.8442;
    p->ntypes        = 1;
    p->ntimes        = 200;
    p->dt            = 0.005;
    p->nx = p->ny = p->nz = 32;
    p->pbc_x = p->pbc_y = p->pbc_z = 1;
    p->cutforce      = 2.5;
    p->skin          = 0.3;
    p->cutneigh      = p->cutforce + p->skin;
    p->temp          = 1.44;
    p->nstat         = 100;
    p->mass          = 1.0;
    p->dtforce       = 0.5 * p->dt;
    p->reneigh_every = 20;
    p->resort_every  = 400;
    p->prune_every   = 1000;
    p->x_out_every   = 20;
    p->v_out_every   = 5;
    p->half_neigh    = 0;
    p->proc_freq     = 2.4;
    p->precision     = MDB_DP; /* config.mk: DATA_TYPE ?= DP */
    p->layout        = MDB_AOS; /* config.mk: DATA_LAYOUT ?= AOS */
}

void readParameter(Parameter* p, const char* filename)
{
    FILE* fp = fopen(filename, "r");
    char line[MAXLINE];
    if (!fp) {
        fprintf(stderr, "Could not open parameter file: %s\n", filename);
        exit(-1);
    }
    while (fgets(line, MAXLINE, fp)) {
        char* hash = strchr(line, '#');
        if (hash) *hash = '\0';
        char* tok = strtok(line, " \t\r\n");
        char* val = strtok(NULL, " \t\r\n");
        if (!tok || !val) continue;
        /* prefix match on the key, first hit in this order is NOT exclusive: every matching key is
         * assigned, exactly like the reference's chain of independent ifs (parameter.c:75-111) */
#define KEY(name) (strncmp(tok, #name, sizeof(#name) - 1) == 0)
        if (KEY(force_field)) p->force_field = str2ff(val);
        if (KEY(input_file)) p->input_file = strdup(val);
        if (KEY(eam_file)) p->eam_file = strdup(val);
        if (KEY(vtk_file)) p->vtk_file = strdup(val);
        if (KEY(xtc_file)) p->xtc_file = strdup(val);
        if (KEY(epsilon)) p->epsilon = atof(val);
        if (KEY(sigma)) p->sigma = atof(val);
        if (KEY(rho)) p->rho = atof(val);
        if (KEY(dt)) p->dt = atof(val);
        if (KEY(cutforce)) p->cutforce = atof(val);
        if (KEY(skin)) p->skin = atof(val);
        if (KEY(temp)) p->temp = atof(val);
        if (KEY(mass)) p->mass = atof(val);
        if (KEY(proc_freq)) p->proc_freq = atof(val);
        if (KEY(ntypes)) p->ntypes = atoi(val);
        if (KEY(ntimes)) p->ntimes = atoi(val);
        if (KEY(nx)) p->nx = atoi(val);
        if (KEY(ny)) p->ny = atoi(val);
        if (KEY(nz)) p->nz = atoi(val);
        if (KEY(pbc_x)) p->pbc_x = atoi(val);
        if (KEY(pbc_y)) p->pbc_y = atoi(val);
        if (KEY(pbc_z)) p->pbc_z = atoi(val);
        if (KEY(nstat)) p->nstat = atoi(val);
        if (KEY(reneigh_every)) p->reneigh_every = atoi(val);
        if (KEY(resort_every)) p->resort_every = atoi(val);
        if (KEY(prune_every)) p->prune_every = atoi(val);
        if (KEY(x_out_every)) p->x_out_every = atoi(val);
        if (KEY(v_out_every)) p->v_out_every = atoi(val);
        if (KEY(half_neigh)) p->half_neigh = atoi(val);
#undef KEY
    }
    p->dtforce = 0.5 * p->dt;
    double s2  = p->sigma * p->sigma;
    p->sigma6  = s2 * s2 * s2;
    fclose(fp);
}

void printParameter(Parameter* p)
{
    printf("Parameters:\n");
    if (p->input_file) printf("\tInput file: %s\n", p->input_file);
    if (p->vtk_file) printf("\tVTK file: %s\n", p->vtk_file);
    if (p->xtc_file) printf("\tXTC file: %s\n", p->xtc_file);
    if (p->eam_file) printf("\tEAM file: %s\n", p->eam_file);
    printf("\tForce field: %s\n", ff2str(p->force_field));
    printf("\tKernel: %s\n", "CUDA-sm_100a");
    printf("\tData layout: %s\n", p->layout == MDB_AOS ? "AoS" : "SoA");
    printf("\tFloating-point precision: %s\n", p->precision == MDB_SP ? "single" : "double");
    printf("\tUnit cells (nx, ny, nz): %d, %d, %d\n", p->nx, p->ny, p->nz);
    printf("\tDomain box sizes (x, y, z): %e, %e, %e\n", p->xprd, p->yprd, p->zprd);
    printf("\tPeriodic (x, y, z): %d, %d, %d\n", p->pbc_x, p->pbc_y, p->pbc_z);
    printf("\tLattice size: %e\n", p->lattice);
    printf("\tEpsilon: %e\n", p->epsilon);
    printf("\tSigma: %e\n", p->sigma);
    printf("\tTemperature: %e\n", p->temp);
    printf("\tRHO: %e\n", p->rho);
    printf("\tMass: %e\n", p->mass);
    printf("\tNumber of types: %d\n", p->ntypes);
    printf("\tNumber of timesteps: %d\n", p->ntimes);
    printf("\tReport stats every (timesteps): %d\n", p->nstat);
    printf("\tReneighbor every (timesteps): %d\n", p->reneigh_every);
    if (p->sort_atoms) printf("\tResort atoms every (timesteps): %d\n", p->reneigh_every);
    else printf("\tSort atoms: no\n");
    printf("\tPrune every (timesteps): %d\n", p->prune_every);
    printf("\tOutput positions every (timesteps): %d\n", p->x_out_every);
    printf("\tOutput velocities every (timesteps): %d\n", p->v_out_every);
    printf("\tDelta time (dt): %e\n", p->dt);
    printf("\tCutoff radius: %e\n", p->cutforce);
    printf("\tSkin: %e\n", p->skin);
    printf("\tHalf neighbor lists: %d\n", p->half_neigh);
    printf("\tProcessor frequency (GHz): %.4f\n", p->proc_freq);
}
