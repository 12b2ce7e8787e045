/*
 * mdbench.h -- host-side driver structures of the B200 build of MD-Bench (verletlist scheme).
 * Mirrors the reference's driver-facing interface so that main.c reads like the reference's:
 *   Parameter      <- reference src/common/parameter.h:27-61   (same field names and meaning)
 *   Atom, Neighbor <- reference src/verletlist/atom.h:24-39, neighbor.h:18-28 (scalar fields only:
 *                     the arrays live on the device behind the mdb_ctx handle; host copies of
 *                     x/v/f are fetched on demand into Atom.x/vx/fx)
 *   function-pointer operators computeForce / buildNeighbor / initialIntegrate / finalIntegrate /
 *   updatePbc / updateAtomsPbc <- force.h:16-17, neighbor.h:55-56, integrate.h:12-14, pbc.h:15-17
 * PRECISION (-DPRECISION=1|2) and AOS/SOA (-DAOS) of the reference are run-time options here
 * (--precision sp|dp, --layout aos|soa).
 */
#ifndef MDBENCH_DRIVER_H
#define MDBENCH_DRIVER_H
#include <stdbool.h>
#include <stdio.h>

#include "../../include/mdb200.h"

#ifndef MAXLINE
#define MAXLINE 4096
#endif
#define HLINE "------------------------------------------------------------------\n"

enum forcetype { FF_LJ = 0, FF_EAM };

typedef struct {
    int force_field;
    char* param_file;
    char* input_file;
    char* vtk_file;
    char* xtc_file;
    char* write_atom_file;
    double epsilon, sigma, sigma6, temp, rho, mass;
    int ntypes, ntimes, nstat, reneigh_every, resort_every, prune_every, x_out_every, v_out_every, half_neigh;
    double dt, dtforce, skin, cutforce, cutneigh;
    int nx, ny, nz, pbc_x, pbc_y, pbc_z;
    double lattice, xlo, xhi, ylo, yhi, zlo, zhi, xprd, yprd, zprd;
    double proc_freq;
    char* eam_file;
    /* run-time replacements of the reference's build options */
    int precision; /* MDB_SP | MDB_DP  (DATA_TYPE)   */
    int layout;    /* MDB_AOS | MDB_SOA (DATA_LAYOUT) */
    int device;    /* CUDA ordinal */
    int sort_atoms; /* SORT_ATOMS */
} Parameter;

typedef struct {
    int Natoms, Nlocal, Nghost, Nmax;
    /* host staging in `precision` / `layout`: what the readers fill and what -w / thermo read */
    void *x, *y, *z, *vx, *vy, *vz;
    int* type;
    int ntypes;
    mdb_ctx* d_atom; /* device state (reference: DeviceAtom d_atom, atom.h:12-22,38) */
} Atom;

typedef struct {
    int every, ncalls, maxneighs, half_neigh;
} Neighbor;

typedef struct {
    long long total_force_neighs, total_force_iters, atoms_within_cutoff, atoms_outside_cutoff;
} Stats;

typedef struct {
    int nrho, nr;
    double drho, dr, cut, mass;
    double *frho, *rhor, *zr; /* 0-based as read from the funcfl file */
} Funcfl;

/* parameter.c */
void initParameter(Parameter*);
void readParameter(Parameter*, const char*);
void printParameter(Parameter*);
int str2ff(const char*);
const char* ff2str(int);
/* atom_io.c */
void initAtom(Atom*);
int readAtom(Atom*, Parameter*);
void writeAtom(Atom*, Parameter*);
void readEamFile(Funcfl*, const char*);
/* operators.c: the reference's operator API bound to libmdb200 */
typedef double (*ComputeForceFunction)(Parameter*, Atom*, Neighbor*, Stats*);
typedef void (*BuildNeighborFunction)(Atom*, Neighbor*);
typedef void (*IntegrationFunction)(bool, Parameter*, Atom*);
typedef void (*UpdatePbcFunction)(Atom*, Parameter*, bool);
extern ComputeForceFunction computeForce;
extern BuildNeighborFunction buildNeighbor;
extern IntegrationFunction initialIntegrate, finalIntegrate;
extern UpdatePbcFunction updatePbc, updateAtomsPbc;
void initForce(Parameter*);
void initDevice(Atom*, Parameter*);
void setupNeighbor(Parameter*, Atom*);
void setupPbc(Atom*, Parameter*);
void setupThermo(Parameter*, Atom*);
void adjustThermo(Parameter*, Atom*);
void computeThermo(int iflag, Parameter*, Atom*);
void createAtom(Atom*, Parameter*);
void mdb_die(const char* where); /* print mdb_last_error() and exit(-1), like cuda_assert (device.c:15-21) */
#endif
