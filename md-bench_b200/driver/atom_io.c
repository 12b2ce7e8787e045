/* atom_io.c -- input readers (.pdb .gro .dmp .in), funcfl EAM reader, -w writer.
 * File formats and field handling follow reference src/verletlist/atom.c:199-562 (readers),
 * 564-588 (writer) and src/common/eam_utils.c:42-93 (readEamFile).  Values are parsed as double and
 * narrowed to the run's precision when staged (the reference narrows on assignment to MD_FLOAT). */
#include <stdlib.h>
#include <string.h>

#include "mdbench.h"

static double *hx, *hy, *hz, *hvx, *hvy, *hvz; /* parsed values, double */
static int hcap;

static void grow(Atom* atom, int need)
{
    if (need <= hcap) return;
    int ncap = hcap ? hcap : 20000; /* DELTA, atom.c:17 */
    while (ncap < need) ncap += 20000;
    double** a[] = { &hx, &hy, &hz, &hvx, &hvy, &hvz };
    for (int k = 0; k < 6; k++) {
        *a[k] = (double*)realloc(*a[k], ncap * sizeof(double));
        memset(*a[k] + hcap, 0, (ncap - hcap) * sizeof(double));
    }
    atom->type = (int*)realloc(atom->type, ncap * sizeof(int));
    memset(atom->type + hcap, 0, (ncap - hcap) * sizeof(int));
    hcap       = ncap;
    atom->Nmax = ncap;
}

void initAtom(Atom* atom) { memset(atom, 0, sizeof *atom); }

static void rdline(char* line, FILE* fp)
{
    if (!fgets(line, MAXLINE, fp)) line[0] = '\0';
}
static int type_str2int(const char* t)
{
    if (t && strncmp(t, "Ar", 2) == 0) return 0;
    fprintf(stderr, "Invalid atom type: %s\n", t ? t : "(null)");
    exit(-1);
}
static void set_box(Parameter* p, double xhi, double yhi, double zhi)
{
    p->xlo = p->ylo = p->zlo = 0.0;
    p->xhi = xhi; p->yhi = yhi; p->zhi = zhi;
    p->xprd = p->xhi - p->xlo; p->yprd = p->yhi - p->ylo; p->zprd = p->zhi - p->zlo;
}
#define TOKD() atof(strtok(NULL, " "))

static int read_pdb(Atom* atom, Parameter* p, FILE* fp)
{
    char line[MAXLINE];
    int n = 0;
    while (!feof(fp)) {
        rdline(line, fp);
        char* item = strtok(line, " ");
        if (!item) continue;
        if (strncmp(item, "CRYST1", 6) == 0) {
            double a = TOKD(), b = TOKD(), c = TOKD();
            set_box(p, a, b, c);
        } else if (strncmp(item, "ATOM", 4) == 0) {
            int id = atoi(strtok(NULL, " ")) - 1;
            grow(atom, id + 2);
            atom->type[id] = type_str2int(strtok(NULL, " "));
            strtok(NULL, " ");       /* label */
            strtok(NULL, " ");       /* comp id */
            hx[id] = TOKD(); hy[id] = TOKD(); hz[id] = TOKD();
            hvx[id] = hvy[id] = hvz[id] = 0.0;
            if (atom->type[id] + 1 > atom->ntypes) atom->ntypes = atom->type[id] + 1;
            atom->Natoms++; atom->Nlocal++; n++;
        } else if (strncmp(item, "HEADER", 6) == 0 || strncmp(item, "REMARK", 6) == 0 ||
                   strncmp(item, "MODEL", 5) == 0 || strncmp(item, "TER", 3) == 0 ||
                   strncmp(item, "ENDMDL", 6) == 0 || item[0] == '\n') {
        } else {
            fprintf(stderr, "Invalid item: %s\n", item);
            exit(-1);
        }
    }
    return n;
}

static int read_gro(Atom* atom, Parameter* p, FILE* fp)
{
    char line[MAXLINE], desc[MAXLINE];
    rdline(desc, fp);
    desc[strcspn(desc, "\n")] = '\0';
    rdline(line, fp);
    int want = atoi(strtok(line, " ")), n = 0;
    fprintf(stdout, "System: %s with %d atoms\n", desc, want);
    while (!feof(fp) && n < want) {
        rdline(line, fp);
        strtok(line, " "); /* residue label */
        int type = type_str2int(strtok(NULL, " "));
        strtok(NULL, " "); /* atom number: the running index is used (atom.c:333-334) */
        int id = n;
        grow(atom, id + 2);
        atom->type[id] = type;
        hx[id] = TOKD(); hy[id] = TOKD(); hz[id] = TOKD();
        hvx[id] = TOKD(); hvy[id] = TOKD(); hvz[id] = TOKD();
        if (type + 1 > atom->ntypes) atom->ntypes = type + 1;
        atom->Natoms++; atom->Nlocal++; n++;
    }
    if (!feof(fp)) {
        rdline(line, fp);
        double a = atof(strtok(line, " ")), b = TOKD(), c = TOKD();
        set_box(p, a, b, c);
    }
    if (n != want) {
        fprintf(stderr, "Input error: Number of atoms read do not match (%d/%d).\n", n, want);
        exit(-1);
    }
    return n;
}

static int read_dmp(Atom* atom, Parameter* p, FILE* fp)
{
    char line[MAXLINE];
    int natoms = 0, n = 0, ts = -1;
    while (!feof(fp) && ts < 1 && !n) {
        rdline(line, fp);
        if (strncmp(line, "ITEM: ", 6) != 0) {
            fprintf(stderr, "Invalid input from file, expected item reference but got:\n%s\n", line);
            exit(-1);
        }
        char* item = &line[6];
        if (strncmp(item, "TIMESTEP", 8) == 0) {
            rdline(line, fp);
            ts = atoi(line);
        } else if (strncmp(item, "NUMBER OF ATOMS", 15) == 0) {
            rdline(line, fp);
            natoms       = atoi(line);
            atom->Natoms = atom->Nlocal = natoms;
            grow(atom, natoms + 1);
        } else if (strncmp(item, "BOX BOUNDS pp pp pp", 19) == 0) {
            rdline(line, fp); p->xlo = atof(strtok(line, " ")); p->xhi = TOKD(); p->xprd = p->xhi - p->xlo;
            rdline(line, fp); p->ylo = atof(strtok(line, " ")); p->yhi = TOKD(); p->yprd = p->yhi - p->ylo;
            rdline(line, fp); p->zlo = atof(strtok(line, " ")); p->zhi = TOKD(); p->zprd = p->zhi - p->zlo;
        } else if (strncmp(item, "ATOMS id type x y z vx vy vz", 28) == 0) {
            for (int i = 0; i < natoms; i++) {
                rdline(line, fp);
                int id         = atoi(strtok(line, " ")) - 1;
                atom->type[id] = atoi(strtok(NULL, " "));
                hx[id] = TOKD(); hy[id] = TOKD(); hz[id] = TOKD();
                hvx[id] = TOKD(); hvy[id] = TOKD(); hvz[id] = TOKD();
                if (atom->type[id] > atom->ntypes) atom->ntypes = atom->type[id];
                n++;
            }
        } else {
            fprintf(stderr, "Invalid item: %s\n", item);
            exit(-1);
        }
    }
    if (ts < 0 || !natoms || !n) {
        fprintf(stderr, "Input error: atom data was not read!\n");
        exit(-1);
    }
    return natoms;
}

static int read_in(Atom* atom, Parameter* p, FILE* fp)
{
    char line[MAXLINE];
    rdline(line, fp);
    int natoms = atoi(strtok(line, " "));
    p->xlo = TOKD(); p->xhi = TOKD(); p->ylo = TOKD(); p->yhi = TOKD(); p->zlo = TOKD(); p->zhi = TOKD();
    p->xprd = p->xhi - p->xlo; p->yprd = p->yhi - p->ylo; p->zprd = p->zhi - p->zlo;
    atom->Natoms = atom->Nlocal = natoms;
    atom->ntypes = 1;
    grow(atom, natoms + 1);
    for (int i = 0; i < natoms; i++) {
        rdline(line, fp);
        char* m = strtok(line, " ");
        if (m && strncmp(m, "inf", 3) != 0) p->mass = atof(m);
        hx[i] = TOKD(); hy[i] = TOKD(); hz[i] = TOKD();
        hvx[i] = TOKD(); hvy[i] = TOKD(); hvz[i] = TOKD();
        atom->type[i] = 0;
    }
    if (!natoms) {
        fprintf(stderr, "Input error: atom data was not read!\n");
        exit(-1);
    }
    return natoms;
}

/* stage the parsed doubles in the run's precision and layout (what mdb_setAtoms expects) */
static void stage(Atom* atom, Parameter* p)
{
    const int n = atom->Nlocal;
    const size_t es = p->precision == MDB_SP ? sizeof(float) : sizeof(double);
    double* src[6] = { hx, hy, hz, hvx, hvy, hvz };
    void** dst[6]  = { &atom->x, &atom->y, &atom->z, &atom->vx, &atom->vy, &atom->vz };
    if (p->layout == MDB_AOS) {
        atom->x  = malloc(3 * n * es);
        atom->vx = malloc(3 * n * es);
        for (int i = 0; i < n; i++)
            for (int c = 0; c < 3; c++) {
                if (p->precision == MDB_SP) {
                    ((float*)atom->x)[3 * i + c]  = (float)src[c][i];
                    ((float*)atom->vx)[3 * i + c] = (float)src[3 + c][i];
                } else {
                    ((double*)atom->x)[3 * i + c]  = src[c][i];
                    ((double*)atom->vx)[3 * i + c] = src[3 + c][i];
                }
            }
    } else {
        for (int k = 0; k < 6; k++) {
            *dst[k] = malloc(n * es);
            for (int i = 0; i < n; i++) {
                if (p->precision == MDB_SP) ((float*)*dst[k])[i] = (float)src[k][i];
                else ((double*)*dst[k])[i] = src[k][i];
            }
        }
    }
}

int readAtom(Atom* atom, Parameter* p)
{
    const int len = (int)strlen(p->input_file);
    FILE* fp      = fopen(p->input_file, "r");
    if (!fp) {
        fprintf(stderr, "Could not open input file: %s\n", p->input_file);
        exit(-1);
    }
    int n;
    if (len >= 4 && strncmp(&p->input_file[len - 4], ".pdb", 4) == 0) n = read_pdb(atom, p, fp);
    else if (len >= 4 && strncmp(&p->input_file[len - 4], ".gro", 4) == 0) n = read_gro(atom, p, fp);
    else if (len >= 4 && strncmp(&p->input_file[len - 4], ".dmp", 4) == 0) n = read_dmp(atom, p, fp);
    else if (len >= 3 && strncmp(&p->input_file[len - 3], ".in", 3) == 0) n = read_in(atom, p, fp);
    else {
        fprintf(stderr, "Invalid input file extension: %s\nValid choices are: pdb, gro, dmp, in\n", p->input_file);
        exit(-1);
    }
    fclose(fp);
    if (!n) {
        fprintf(stderr, "Input error: No atoms read!\n");
        exit(-1);
    }
    stage(atom, p);
    fprintf(stdout, "Read %d atoms from %s\n", n, p->input_file);
    return n;
}

/* -w <file>: "type,1.0,x,y,z,vx,vy,vz,0" per atom (atom.c:564-588); expects Atom.x/vx filled */
void writeAtom(Atom* atom, Parameter* p)
{
    FILE* fp = fopen(p->write_atom_file, "w");
    if (!fp) {
        fprintf(stderr, "Could not open %s\n", p->write_atom_file);
        exit(-1);
    }
    for (int i = 0; i < atom->Nlocal; i++) {
        double v[6];
        for (int c = 0; c < 3; c++) {
            if (p->layout == MDB_AOS) {
                v[c]     = p->precision == MDB_SP ? ((float*)atom->x)[3 * i + c] : ((double*)atom->x)[3 * i + c];
                v[3 + c] = p->precision == MDB_SP ? ((float*)atom->vx)[3 * i + c] : ((double*)atom->vx)[3 * i + c];
            } else {
                void* px[3] = { atom->x, atom->y, atom->z };
                void* pv[3] = { atom->vx, atom->vy, atom->vz };
                v[c]        = p->precision == MDB_SP ? ((float*)px[c])[i] : ((double*)px[c])[i];
                v[3 + c]    = p->precision == MDB_SP ? ((float*)pv[c])[i] : ((double*)pv[c])[i];
            }
        }
        fprintf(fp, "%d,%f,%f,%f,%f,%f,%f,%f,0\n", atom->type ? atom->type[i] : 0, 1.0, v[0], v[1], v[2], v[3], v[4], v[5]);
    }
    fclose(fp);
    fprintf(stdout, "Wrote input data to %s, grid size: %f, %f, %f\n", p->write_atom_file, p->xprd, p->yprd, p->zprd);
}

/* funcfl potential file: 2 header lines (comment; Z mass ...), then "nrho drho nr dr cut", then
 * nrho values of F(rho), nr of Z(r), nr of rho(r), free format (eam_utils.c:42-93) */
static void grab(FILE* fp, int n, double* list)
{
    char line[MAXLINE];
    int i = 0;
    while (i < n) {
        if (!fgets(line, MAXLINE, fp)) {
            fprintf(stderr, "EAM potential file is truncated\n");
            exit(-1);
        }
        for (char* t = strtok(line, " \t\n\r\f"); t && i < n; t = strtok(NULL, " \t\n\r\f")) list[i++] = atof(t);
    }
}
void readEamFile(Funcfl* f, const char* filename)
{
    FILE* fp = fopen(filename, "r");
    char line[MAXLINE];
    if (!fp) {
        printf("Can't open EAM Potential file: %s\n", filename);
        exit(0);
    }
    int tmp;
    rdline(line, fp);
    rdline(line, fp);
    sscanf(line, "%d %lg", &tmp, &f->mass);
    rdline(line, fp);
    sscanf(line, "%d %lg %d %lg %lg", &f->nrho, &f->drho, &f->nr, &f->dr, &f->cut);
    f->frho = (double*)malloc(f->nrho * sizeof(double));
    f->zr   = (double*)malloc(f->nr * sizeof(double));
    f->rhor = (double*)malloc(f->nr * sizeof(double));
    grab(fp, f->nrho, f->frho);
    grab(fp, f->nr, f->zr);
    grab(fp, f->nr, f->rhor);
    fclose(fp);
}
