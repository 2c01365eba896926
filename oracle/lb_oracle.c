/* lb_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see lb_oracle.h).
 *
 * Dense-grid CPU restatement of hybird's per-timestep LB update.  Where the reference walks
 * std::vector index lists (fluidNodes / interfaceNodes / particleNodes / activeNodes) this file
 * scans the dense type map in ascending index order, which is the order those lists have after
 * LB::cleanLists (LB.cpp:1000-1060) sorted them; the few places where the reference's list
 * *order* changes the result (updateInterface's mutant lists, the flood fill in findNewSolid,
 * the sequential sums of massSurplus / extraMass / element forces) keep explicit lists so that
 * every floating-point operation happens in the same order as in the reference run at
 * OMP_NUM_THREADS=1.  Floating-point expressions are written with the reference's association
 * (x86-64 SSE2, no FMA contraction: build with -ffp-contract=off).
 *
 * Curved walls (type 9, DRUM/AVALANCHE/NET geometries, LB.cpp:1278-1319) need their link fractions
 * (lbo_set_curves); LB::enforceMassConservation (DRUM only, LB.cpp:1806-1822) runs once a mass
 * target is set (lbo_set_mass_target).
 */
#include "lb_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define Q 19

/* D3Q19 velocity set, lattice.h:36-60 */
static const double VX[Q] = { 0, 1, -1, 0, 0, 0, 0, 1, -1, -1, 1, 0, 0, 0, 0, 1, -1, 1, -1 };
static const double VY[Q] = { 0, 0, 0, 1, -1, 0, 0, 1, -1, 1, -1, 1, -1, -1, 1, 0, 0, 0, 0 };
static const double VZ[Q] = { 0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, -1, 1 };
static const int CX[Q] = { 0, 1, -1, 0, 0, 0, 0, 1, -1, -1, 1, 0, 0, 0, 0, 1, -1, 1, -1 };
static const int CY[Q] = { 0, 0, 0, 1, -1, 0, 0, 1, -1, 1, -1, 1, -1, -1, 1, 0, 0, 0, 0 };
static const int CZ[Q] = { 0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, -1, 1 };
/* lattice.h:85 */
static const int OPP[Q] = { 0, 2, 1, 4, 3, 6, 5, 8, 7, 10, 9, 12, 11, 14, 13, 16, 15, 18, 17 };
/* lattice.h:88-91 */
static const int SLIP1CHECK[Q] = { 0, 0, 0, 0, 0, 0, 0, 1, 2, 3, 4, 3, 4, 5, 6, 1, 2, 6, 5 };
static const int SLIP1[Q] = { 0, 0, 0, 0, 0, 0, 0, 9, 10, 8, 7, 13, 14, 12, 11, 18, 17, 15, 16 };
static const int SLIP2CHECK[Q] = { 0, 0, 0, 0, 0, 0, 0, 3, 4, 2, 1, 5, 6, 4, 3, 5, 6, 1, 2 };
static const int SLIP2[Q] = { 0, 0, 0, 0, 0, 0, 0, 10, 9, 7, 8, 14, 13, 11, 12, 17, 18, 16, 15 };
/* lattice.h:98-101 */
static const double W[Q] = { 12.0 / 36.0, 2.0 / 36.0, 2.0 / 36.0, 2.0 / 36.0, 2.0 / 36.0, 2.0 / 36.0, 2.0 / 36.0,
                             1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0,
                             1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0 };
static const double LBM_DT = 1.0;   /* lattice.h:26 */
static const double MIN_TAU = 0.501; /* lattice.h:29 */
static const double MAX_TAU = 1.8;   /* lattice.h:30 */

struct LboState {
    int X, Y, Z;
    uint32_t N;
    LboParams p;
    double lbF[3]; /* zeroed for good by the first collision when !forceField (LB.cpp:1074-1076) */
    double initDensity;
    double uLength, uSpeed, uAngVel, uForce, uTorque, uVolume; /* node.cpp:476-488 */
    uint8_t* type; /* t | p<<4 | node<<5 */
    uint32_t* solidIndex;
    double *f, *fs, *n, *u, *hydroForce, *mass, *newMass, *visc, *shearRate;
    uint8_t* mark;
    uint32_t* curveRow;  /* per cell: row of curveDelta, UINT32_MAX for cells without a `curve` object (LB.cpp:363-367) */
    double* curveDelta;  /* [nCurves][19]: curve::delta (node.h:133-148) */
    uint32_t nCurves;
    int enforceMass;     /* problemName == DRUM (LB.cpp:239-244) */
    double totalMass;    /* LB::totalMass (LB.cpp:204-218) */
    uint32_t *listA, *listB, *listC, *listD; /* scratch index lists, capacity N each */
};

static int g_threads = 1;
void lbo_set_threads(int n) { g_threads = n < 1 ? 1 : n; }

static inline int T(const LboState* s, uint32_t i) { return s->type[i] & LBO_TYPE_MASK; }
static inline int isActiveT(int t) { return t == LBO_FLUID || t == LBO_INTERFACE; } /* node.cpp:299-308 */
static inline void setT(LboState* s, uint32_t i, int t) { s->type[i] = (uint8_t)((s->type[i] & ~LBO_TYPE_MASK) | t); }

/* neighbors[i].d[j] as built by LB::initializeLatticeBoundaries (LB.cpp:377-472): shell cells
 * point to themselves; interior cells get i+ne[j] with a per-axis periodic wrap by domain[k];
 * d[0] of interior cells is never assigned and stays 0. */
static uint32_t nbr(const LboState* s, uint32_t i, int j) {
    const int X = s->X, Y = s->Y, Z = s->Z;
    int x = (int)(i % (uint32_t)X), y = (int)((i / (uint32_t)X) % (uint32_t)Y), z = (int)(i / ((uint32_t)X * (uint32_t)Y));
    if (x == 0 || x == X - 1 || y == 0 || y == Y - 1 || z == 0 || z == Z - 1) return i;
    if (j == 0) return 0;
    x += CX[j]; y += CY[j]; z += CZ[j];
    if (x == X - 1 && s->p.boundary[1] == LBO_PERIODIC) x = 1;
    else if (x == 0 && s->p.boundary[0] == LBO_PERIODIC) x = X - 2;
    if (y == Y - 1 && s->p.boundary[3] == LBO_PERIODIC) y = 1;
    else if (y == 0 && s->p.boundary[2] == LBO_PERIODIC) y = Y - 2;
    if (z == Z - 1 && s->p.boundary[5] == LBO_PERIODIC) z = 1;
    else if (z == 0 && s->p.boundary[4] == LBO_PERIODIC) z = Z - 2;
    return (uint32_t)x + (uint32_t)X * ((uint32_t)y + (uint32_t)Y * (uint32_t)z);
}
uint32_t lbo_neighbor(const LboState* s, uint32_t i, int j) { return nbr(s, i, j); }

/* node::computeEquilibrium / node::setEquilibrium (node.cpp:40-61, 90-101) */
static void equilibrium(double n, const double u[3], double feq[Q]) {
    const double C1 = 3.0 * LBM_DT * LBM_DT, C2 = 4.5 * LBM_DT * LBM_DT * LBM_DT * LBM_DT, C3 = 1.5 * LBM_DT * LBM_DT;
    const double usq = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
    for (int j = 0; j < Q; ++j) {
        const double vu = u[0] * VX[j] + u[1] * VY[j] + u[2] * VZ[j];
        feq[j] = W[j] * n * (1.0 + C1 * vu + C2 * vu * vu - C3 * usq);
    }
}

/* node::initialize (node.cpp:26-38) */
static void nodeInitialize(LboState* s, uint32_t i, double density, const double vel[3], double massFunction,
                           double viscosity, const double F[3]) {
    double feq[Q];
    s->shearRate[i] = 0.0;
    s->n[i] = density;
    s->mass[i] = massFunction;
    for (int k = 0; k < 3; ++k) s->u[3 * i + k] = F[k] * LBM_DT / 2.0 / density + vel[k];
    s->visc[i] = viscosity;
    equilibrium(density, &s->u[3 * i], feq);
    for (int j = 0; j < Q; ++j) s->f[(size_t)Q * i + j] = s->fs[(size_t)Q * i + j] = feq[j];
}

static void nodeDelete(LboState* s, uint32_t i) {
    s->type[i] &= (uint8_t)~LBO_NODE_BIT;
    memset(&s->f[(size_t)Q * i], 0, Q * sizeof(double));
    memset(&s->fs[(size_t)Q * i], 0, Q * sizeof(double));
    s->n[i] = s->mass[i] = s->newMass[i] = s->visc[i] = s->shearRate[i] = 0.0;
    for (int k = 0; k < 3; ++k) s->u[3 * i + k] = s->hydroForce[3 * i + k] = 0.0;
}

static void nodeCreate(LboState* s, uint32_t i) {
    /* node::node() (node.h:36-46) */
    nodeDelete(s, i);
    s->type[i] |= LBO_NODE_BIT;
    s->visc[i] = 1.0;
}

LboState* lbo_create(const LboParams* p, const uint8_t* type_flags, const uint32_t* solidIndex, const double* f,
                     const double* n, const double* u, const double* mass, const double* visc) {
    LboState* s = (LboState*)calloc(1, sizeof(LboState));
    s->p = *p;
    s->X = p->size[0]; s->Y = p->size[1]; s->Z = p->size[2];
    s->N = (uint32_t)s->X * (uint32_t)s->Y * (uint32_t)s->Z;
    const uint32_t N = s->N;
    for (int k = 0; k < 3; ++k) s->lbF[k] = p->lbF[k];
    s->initDensity = 1.0; /* LB.cpp:147 */
    /* measureUnits::setComposite (node.cpp:476-488) */
    const double L = p->unitLength, Tm = p->unitTime, D = p->unitDensity;
    s->uLength = L;
    s->uVolume = L * L * L;
    s->uSpeed = L / Tm;
    s->uAngVel = 1.0 / Tm;
    s->uForce = D * L * L * L * L / Tm / Tm;
    s->uTorque = D * L * L * L * L * L / Tm / Tm;
    s->type = (uint8_t*)malloc(N);
    s->solidIndex = (uint32_t*)malloc(sizeof(uint32_t) * N);
    s->f = (double*)calloc((size_t)Q * N, sizeof(double));
    s->fs = (double*)calloc((size_t)Q * N, sizeof(double));
    s->n = (double*)calloc(N, sizeof(double));
    s->u = (double*)calloc((size_t)3 * N, sizeof(double));
    s->hydroForce = (double*)calloc((size_t)3 * N, sizeof(double));
    s->mass = (double*)calloc(N, sizeof(double));
    s->newMass = (double*)calloc(N, sizeof(double));
    s->visc = (double*)calloc(N, sizeof(double));
    s->shearRate = (double*)calloc(N, sizeof(double));
    s->mark = (uint8_t*)calloc(N, 1);
    s->listA = (uint32_t*)malloc(sizeof(uint32_t) * N);
    s->listB = (uint32_t*)malloc(sizeof(uint32_t) * N);
    s->listC = (uint32_t*)malloc(sizeof(uint32_t) * N);
    s->listD = (uint32_t*)malloc(sizeof(uint32_t) * N);
    memcpy(s->type, type_flags, N);
    memcpy(s->solidIndex, solidIndex, sizeof(uint32_t) * N);
    for (uint32_t i = 0; i < N; ++i) {
        if (!(s->type[i] & LBO_NODE_BIT)) continue;
        s->n[i] = n[i];
        s->mass[i] = mass[i];
        s->visc[i] = visc[i];
        for (int k = 0; k < 3; ++k) s->u[3 * i + k] = u[3 * i + k];
        if (f) {
            for (int j = 0; j < Q; ++j) s->f[(size_t)Q * i + j] = s->fs[(size_t)Q * i + j] = f[(size_t)Q * i + j];
        } else {
            double feq[Q];
            equilibrium(s->n[i], &s->u[3 * i], feq);
            for (int j = 0; j < Q; ++j) s->f[(size_t)Q * i + j] = s->fs[(size_t)Q * i + j] = feq[j];
        }
    }
    return s;
}

void lbo_destroy(LboState* s) {
    if (!s) return;
    free(s->type); free(s->solidIndex); free(s->f); free(s->fs); free(s->n); free(s->u); free(s->hydroForce);
    free(s->mass); free(s->newMass); free(s->visc); free(s->shearRate); free(s->mark);
    free(s->listA); free(s->listB); free(s->listC); free(s->listD);
    free(s->curveRow); free(s->curveDelta);
    free(s);
}

/* LB::curves as LB::initializeCurved left it (LB.cpp:589-603): delta[j] per curved-wall cell */
void lbo_set_curves(LboState* s, uint32_t nCurves, const uint32_t* cells, const double* delta) {
    free(s->curveRow); free(s->curveDelta);
    s->curveRow = (uint32_t*)malloc(sizeof(uint32_t) * s->N);
    for (uint32_t i = 0; i < s->N; ++i) s->curveRow[i] = UINT32_MAX;
    s->curveDelta = (double*)malloc(sizeof(double) * Q * (nCurves ? nCurves : 1));
    memcpy(s->curveDelta, delta, sizeof(double) * Q * nCurves);
    for (uint32_t k = 0; k < nCurves; ++k) s->curveRow[cells[k]] = k;
    s->nCurves = nCurves;
}

void lbo_set_mass_target(LboState* s, double totalMass) {
    s->enforceMass = 1;
    s->totalMass = totalMass;
}

uint32_t lbo_nodes(const LboState* s) { return s->N; }
const uint8_t* lbo_type_flags(const LboState* s) { return s->type; }
const uint32_t* lbo_solid_index(const LboState* s) { return s->solidIndex; }
const double* lbo_f(const LboState* s) { return s->f; }
const double* lbo_fs(const LboState* s) { return s->fs; }
const double* lbo_n(const LboState* s) { return s->n; }
const double* lbo_u(const LboState* s) { return s->u; }
const double* lbo_hydro_force(const LboState* s) { return s->hydroForce; }
const double* lbo_mass(const LboState* s) { return s->mass; }
const double* lbo_visc(const LboState* s) { return s->visc; }
const double* lbo_shear_rate(const LboState* s) { return s->shearRate; }
uint32_t lbo_count_type(const LboState* s, int t) {
    uint32_t c = 0;
    for (uint32_t i = 0; i < s->N; ++i) c += (T(s, i) == t);
    return c;
}

/* ------------------------------------------------------------------------------------------ */
/* free surface: LB::updateMass (LB.cpp:1492-1590)                                             */
/* ------------------------------------------------------------------------------------------ */
static void updateMass(LboState* s) {
    const uint32_t N = s->N;
#pragma omp parallel for schedule(static) num_threads(g_threads)
    for (uint32_t i = 0; i < N; ++i) {
        if (T(s, i) != LBO_INTERFACE) continue;
        const double* f = &s->f[(size_t)Q * i];
        const double* fs = &s->fs[(size_t)Q * i];
        double deltaMass = 0.0;
        for (int j = 1; j < Q; ++j) {
            const uint32_t link = nbr(s, i, j);
            const int tl = T(s, link);
            double averageMass = 0.0;
            if (tl == LBO_INTERFACE) averageMass = 0.5 * (s->mass[link] + s->mass[i]);
            else if (tl == LBO_FLUID) averageMass = 1.0;
            else if (tl == LBO_GAS) averageMass = 0.0;
            else if (tl == LBO_DYN_WALL) averageMass = 1.0 * s->mass[i];
            else if (tl == LBO_CURVED) averageMass = 1.0 * s->mass[i];
            else if (tl == LBO_SLIP_DYN) {
                if (j > 6) {
                    const int a1 = isActiveT(T(s, nbr(s, i, SLIP1CHECK[j])));
                    const int a2 = isActiveT(T(s, nbr(s, i, SLIP2CHECK[j])));
                    if ((a1 && !a2) || (!a1 && a2)) averageMass += 1.0 * (1.0 - s->p.slipCoefficient) * s->mass[i];
                    else averageMass += 1.0 * s->mass[i];
                } else averageMass += 1.0 * s->mass[i];
            }
            /* node::massStream (node.cpp:293-295) */
            deltaMass += LBM_DT * averageMass * (f[OPP[j]] - fs[j]);
        }
        s->newMass[i] = s->mass[i];
        s->newMass[i] += deltaMass;
    }
#pragma omp parallel for schedule(static) num_threads(g_threads)
    for (uint32_t i = 0; i < N; ++i) {
        const int t = T(s, i);
        if (t == LBO_FLUID) s->mass[i] = s->n[i];
        else if (t == LBO_INTERFACE) s->mass[i] = s->newMass[i];
    }
}

/* LB::redistributeMass (LB.cpp:1796-1804) */
static void redistributeMass(LboState* s, double massSurplus, const uint32_t* interfaceList, uint32_t nInterface) {
    const double addMass = massSurplus / (double)nInterface;
    for (uint32_t k = 0; k < nInterface; ++k) s->mass[interfaceList[k]] += addMass;
}

static uint32_t listInterface(const LboState* s, uint32_t* out) {
    uint32_t c = 0;
    for (uint32_t i = 0; i < s->N; ++i)
        if (T(s, i) == LBO_INTERFACE) out[c++] = i;
    return c;
}

/* LB::updateInterface (LB.cpp:1592-1618) with its five passes */
static void updateInterface(LboState* s) {
    uint32_t* iface = s->listA;   /* interfaceNodes, ascending (as cleanLists left it) */
    uint32_t* filled = s->listB;
    uint32_t* emptied = s->listC;
    uint32_t* newIf = s->listD;
    uint32_t nIf = listInterface(s, iface), nFilled = 0, nEmptied = 0, nNew = 0;
    double massSurplus = 0.0;

    /* findInterfaceMutants (LB.cpp:1620-1650): backwards over the list => descending lists */
    {
        uint32_t keep = 0;
        for (int64_t ind = (int64_t)nIf - 1; ind >= 0; --ind) {
            const uint32_t i = iface[ind];
            if (s->mass[i] > s->n[i]) { filled[nFilled++] = i; setT(s, i, LBO_FLUID); iface[ind] = UINT32_MAX; }
            else if (s->mass[i] < 0.0) { emptied[nEmptied++] = i; setT(s, i, LBO_GAS); iface[ind] = UINT32_MAX; }
        }
        for (uint32_t k = 0; k < nIf; ++k) if (iface[k] != UINT32_MAX) iface[keep++] = iface[k];
        nIf = keep;
    }

    /* smoothenInterface (LB.cpp:1652-1701) */
    for (uint32_t it = 0; it < nFilled; ++it) {
        const uint32_t index = filled[it];
        for (int j = 1; j < Q; ++j) {
            const uint32_t link = nbr(s, index, j);
            if (T(s, link) == LBO_GAS) {
                setT(s, link, LBO_INTERFACE);
                newIf[nNew++] = link;
                s->mark[link] = 1;
                double donorU[3], donorF[3];
                for (int k = 0; k < 3; ++k) {
                    donorU[k] = s->u[3 * index + k];
                    donorF[k] = s->hydroForce[3 * index + k] + s->lbF[k];
                }
                const double donorVisc = s->visc[index];
                nodeCreate(s, link);
                nodeInitialize(s, link, s->initDensity, donorU, 0.01 * s->initDensity, donorVisc, donorF);
                massSurplus -= 0.01 * s->initDensity;
            }
        }
    }
    for (uint32_t it = 0; it < nEmptied; ++it) {
        const uint32_t index = emptied[it];
        for (int j = 0; j < Q; ++j) {
            const uint32_t link = nbr(s, index, j);
            if (T(s, link) == LBO_FLUID) {
                setT(s, link, LBO_INTERFACE);
                newIf[nNew++] = link;
                s->mark[link] = 1;
                s->mass[link] = 0.99 * s->n[link];
                massSurplus += 0.01 * s->n[link];
            }
        }
    }

    /* updateMutants (LB.cpp:1703-1742) */
    for (uint32_t it = 0; it < nEmptied; ++it) {
        const uint32_t i = emptied[it];
        if (!s->mark[i]) { massSurplus += s->mass[i]; nodeDelete(s, i); }
    }
    for (uint32_t it = 0; it < nFilled; ++it) {
        const uint32_t i = filled[it];
        if (!s->mark[i]) { massSurplus += s->mass[i] - s->n[i]; s->mass[i] = s->n[i]; }
    }
    for (uint32_t it = 0; it < nNew; ++it) { iface[nIf++] = newIf[it]; s->mark[newIf[it]] = 0; }

    /* removeIsolated (LB.cpp:1744-1794) */
    for (int64_t ind = (int64_t)nIf - 1; ind >= 0; --ind) {
        const uint32_t i = iface[ind];
        int surroundedFluid = 1;
        for (int j = 1; j < Q; ++j)
            if (T(s, nbr(s, i, j)) == LBO_GAS) { surroundedFluid = 0; break; }
        if (surroundedFluid) {
            massSurplus += s->mass[i] - s->n[i];
            s->mass[i] = s->n[i];
            setT(s, i, LBO_FLUID);
            iface[ind] = UINT32_MAX;
        }
    }
    for (int64_t ind = (int64_t)nIf - 1; ind >= 0; --ind) {
        const uint32_t i = iface[ind];
        if (i == UINT32_MAX) continue;
        int surroundedGas = 1;
        for (int j = 1; j < Q; ++j)
            if (T(s, nbr(s, i, j)) == LBO_FLUID) { surroundedGas = 0; break; }
        if (surroundedGas) {
            massSurplus += s->mass[i];
            setT(s, i, LBO_GAS);
            nodeDelete(s, i);
            iface[ind] = UINT32_MAX;
        }
    }
    {
        uint32_t keep = 0;
        for (uint32_t k = 0; k < nIf; ++k) if (iface[k] != UINT32_MAX) iface[keep++] = iface[k];
        nIf = keep;
    }
    redistributeMass(s, massSurplus, iface, nIf);
}

/* LB::enforceMassConservation (LB.cpp:1806-1822) */
static void enforceMassConservation(LboState* s) {
    double thisMass = 0.0;
    for (uint32_t i = 0; i < s->N; ++i)
        if (isActiveT(T(s, i)) && !(s->type[i] & LBO_P_BIT)) thisMass += s->mass[i];
    const double massDeficit = (thisMass - s->totalMass);
    const uint32_t nIf = listInterface(s, s->listA);
    redistributeMass(s, -0.01 * massDeficit, s->listA, nIf);
}

void lbo_free_surface_step(LboState* s) {
    updateMass(s);
    updateInterface(s);
    if (s->enforceMass) enforceMassConservation(s);
}

/* ------------------------------------------------------------------------------------------ */
/* particle coupling flags: LB.cpp:475-495, 1826-1849, 1921-2033                               */
/* ------------------------------------------------------------------------------------------ */
/* tVect::insideSphere (vector.cpp:153-158) with center = x0/unit.Length, radius = r/unit.Length */
static int insideParticle(const LboState* s, uint32_t i, const LboParticle* pt) {
    const double pos[3] = { (double)(i % (uint32_t)s->X), (double)((i / (uint32_t)s->X) % (uint32_t)s->Y),
                            (double)(i / ((uint32_t)s->X * (uint32_t)s->Y)) };
    const double radius = pt->r / s->uLength;
    double d[3];
    for (int k = 0; k < 3; ++k) d[k] = pos[k] - pt->x0[k] / s->uLength;
    return (d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) < radius * radius;
}

void lbo_coupling_step(LboState* s, int newNeighborList, const LboParticle* parts, uint32_t nParts,
                       const LboElement* elmts, uint32_t nElmts, const uint32_t* components) {
    const uint32_t N = s->N;
    (void)nElmts;
    if (newNeighborList) {
        /* LB::updateIndices -> initializeParticleBoundaries (LB.cpp:475-495): highest index wins */
#pragma omp parallel for schedule(static) num_threads(g_threads)
        for (uint32_t i = 0; i < N; ++i) {
            s->type[i] &= (uint8_t)~LBO_P_BIT;
            if (!isActiveT(T(s, i))) continue;
            for (uint32_t n = 0; n < nParts; ++n)
                if (insideParticle(s, i, &parts[n])) { s->type[i] |= LBO_P_BIT; s->solidIndex[i] = parts[n].particleIndex; }
        }
    }
    /* particleNodes, ascending */
    uint32_t* plist = s->listA;
    uint32_t nP = 0;
    for (uint32_t i = 0; i < N; ++i) if (s->type[i] & LBO_P_BIT) plist[nP++] = i;
    /* findNewActive (LB.cpp:1921-1967) */
    {
        uint32_t keep = 0;
        for (uint32_t ip = 0; ip < nP; ++ip) {
            const uint32_t index = plist[ip];
            const LboElement* e = &elmts[parts[s->solidIndex[index]].clusterIndex];
            int newActive = 1;
            for (uint32_t c = e->compBegin; c < e->compEnd; ++c)
                if (insideParticle(s, index, &parts[components[c]])) { newActive = 0; break; }
            if (newActive) s->type[index] &= (uint8_t)~LBO_P_BIT;
            else plist[keep++] = index;
        }
        nP = keep;
    }
    /* findNewSolid (LB.cpp:1969-2033): the list grows while it is walked (flood fill) */
    for (uint32_t it = 0; it < nP; ++it) {
        const uint32_t index = plist[it];
        const LboElement* e = &elmts[parts[s->solidIndex[index]].clusterIndex];
        for (int k = 1; k < 7; ++k) {
            const uint32_t link = nbr(s, index, k);
            if (s->type[link] & LBO_P_BIT) continue;
            for (uint32_t c = e->compBegin; c < e->compEnd; ++c) {
                if (insideParticle(s, link, &parts[components[c]])) {
                    s->solidIndex[link] = components[c];
                    s->type[link] |= LBO_P_BIT;
                    plist[nP++] = link;
                    break;
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* LB step: reconstruction, computeHydroForces, collision, streaming                           */
/* ------------------------------------------------------------------------------------------ */
/* node::reconstruct (node.cpp:63-83) */
static void reconstruction(LboState* s) {
#pragma omp parallel for schedule(static) num_threads(g_threads)
    for (uint32_t i = 0; i < s->N; ++i) {
        if (!isActiveT(T(s, i))) continue;
        const double* f = &s->f[(size_t)Q * i];
        double n = 0.0, ux = 0.0, uy = 0.0, uz = 0.0;
        for (int j = 0; j < Q; ++j) n += f[j];
        for (int j = 0; j < Q; ++j) { ux += f[j] * VX[j]; uy += f[j] * VY[j]; uz += f[j] * VZ[j]; }
        s->n[i] = n;
        s->u[3 * i] = ux / n; s->u[3 * i + 1] = uy / n; s->u[3 * i + 2] = uz / n;
    }
}

/* LB::computeHydroForces (LB.cpp:1851-1919); element sums in ascending cell order */
static void computeHydroForces(LboState* s, const LboParticle* parts, const LboElement* elmts, uint32_t nElmts,
                               double* FHydro, double* MHydro, double* fluidVolume) {
    double* F = (double*)calloc((size_t)7 * (nElmts ? nElmts : 1), sizeof(double));
    for (uint32_t i = 0; i < s->N; ++i) {
        if (!isActiveT(T(s, i))) continue;
        double* h = &s->hydroForce[3 * i];
        h[0] = h[1] = h[2] = 0.0;
        if (!(s->type[i] & LBO_P_BIT)) continue;
        const LboParticle* pt = &parts[s->solidIndex[i]];
        const uint32_t ci = pt->clusterIndex;
        const double pos[3] = { (double)(i % (uint32_t)s->X), (double)((i / (uint32_t)s->X) % (uint32_t)s->Y),
                                (double)(i / ((uint32_t)s->X * (uint32_t)s->Y)) };
        double radius[3], wxr[3], localVel[3], diffVel[3];
        for (int k = 0; k < 3; ++k) radius[k] = pos[k] - pt->x0[k] / s->uLength + pt->radiusVec[k] / s->uLength;
        const double* w = elmts[ci].wGlobal;
        wxr[0] = w[1] * radius[2] - w[2] * radius[1];
        wxr[1] = w[2] * radius[0] - w[0] * radius[2];
        wxr[2] = w[0] * radius[1] - w[1] * radius[0];
        for (int k = 0; k < 3; ++k) localVel[k] = elmts[ci].x1[k] / s->uSpeed + wxr[k] / s->uAngVel;
        const double liquidFraction = s->mass[i] / s->n[i]; /* node.cpp:12-14 */
        for (int k = 0; k < 3; ++k) diffVel[k] = (s->u[3 * i + k] - localVel[k]) * liquidFraction;
        for (int k = 0; k < 3; ++k) h[k] += diffVel[k] * -1.0;
        double* Fe = &F[(size_t)7 * ci];
        Fe[6] += s->mass[i];
        for (int k = 0; k < 3; ++k) Fe[k] += diffVel[k] * 1.0;
        Fe[3] += (radius[1] * diffVel[2] - radius[2] * diffVel[1]) * 1.0;
        Fe[4] += (radius[2] * diffVel[0] - radius[0] * diffVel[2]) * 1.0;
        Fe[5] += (radius[0] * diffVel[1] - radius[1] * diffVel[0]) * 1.0;
    }
    for (uint32_t e = 0; e < nElmts; ++e) {
        for (int k = 0; k < 3; ++k) {
            if (FHydro) FHydro[3 * e + k] = F[(size_t)7 * e + k] * s->uForce;
            if (MHydro) MHydro[3 * e + k] = F[(size_t)7 * e + 3 + k] * s->uTorque;
        }
        if (fluidVolume) fluidVolume[e] = F[(size_t)7 * e + 6] * s->uVolume;
    }
    free(F);
}

/* LB::collision (LB.cpp:1072-1112) and the node kernels it calls (node.cpp:85-184) */
static void collision(LboState* s) {
    if (!s->p.forceField) s->lbF[0] = s->lbF[1] = s->lbF[2] = 0.0;
    const double minVisc = (MIN_TAU - 0.5) / 3 / LBM_DT, maxVisc = (MAX_TAU - 0.5) / 3 / LBM_DT;
    const int shear = s->p.nonNewtonian || s->p.turbulence;
#pragma omp parallel for schedule(static) num_threads(g_threads)
    for (uint32_t i = 0; i < s->N; ++i) {
        if (!isActiveT(T(s, i))) continue;
        double* f = &s->f[(size_t)Q * i];
        double* u = &s->u[3 * i];
        const double n = s->n[i];
        double feq[Q], totalForce[3];
        /* shiftVelocity */
        for (int k = 0; k < 3; ++k) totalForce[k] = s->lbF[k] + s->hydroForce[3 * i + k];
        for (int k = 0; k < 3; ++k) u[k] += totalForce[k] * (0.5 * LBM_DT) / n;
        equilibrium(n, u, feq);
        if (shear) {
            /* computeShearRate; vv[j] = v[j] (x) v[j] (lattice.h:64-82), tMat::magnitude (vector.cpp:453-457) */
            const double tau = 0.5 + 3.0 * s->visc[i] * LBM_DT;
            double g[3][3] = { { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 } };
            for (int j = 0; j < Q; ++j) {
                const double d = f[j] - feq[j];
                const double v[3] = { VX[j], VY[j], VZ[j] };
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < 3; ++b) g[a][b] += (v[a] * v[b]) * d;
            }
            const double sc = 1.5 * LBM_DT / (tau * n);
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) g[a][b] *= sc;
            const double shearRate = sqrt(0.5 * (g[0][0] * g[0][0] + g[1][1] * g[1][1] + g[2][2] * g[2][2] +
                                                 2.0 * (g[0][1] * g[1][0] + g[2][0] * g[0][2] + g[1][2] * g[2][1])));
            s->shearRate[i] = shearRate;
            double nuTurb = 0.0, nuApp;
            if (s->p.turbulence) nuTurb = s->p.turbConst * shearRate;
            if (s->p.nonNewtonian) nuApp = s->p.plasticVisc + s->p.yieldStress / (n * 2.0 * shearRate);
            else nuApp = s->visc[i];
            const double x = nuApp + nuTurb;
            const double lo = (x < maxVisc) ? x : maxVisc; /* std::min(maxVisc, x) */
            s->visc[i] = (minVisc < lo) ? lo : minVisc;    /* std::max(minVisc, lo) */
        }
        /* solveCollision */
        const double omega = 1.0 / (0.5 + 3.0 * s->visc[i] * LBM_DT);
        for (int j = 0; j < Q; ++j) f[j] += omega * (feq[j] - f[j]);
        /* addForce */
        const double F1 = 3.0 * LBM_DT * LBM_DT, F2 = 9.0 * LBM_DT * LBM_DT * LBM_DT * LBM_DT;
        const double omegaf = 1.0 - 1.0 / (1.0 + 6.0 * s->visc[i] * LBM_DT);
        for (int j = 0; j < Q; ++j) {
            const double vu = u[0] * VX[j] + u[1] * VY[j] + u[2] * VZ[j];
            const double vmu[3] = { VX[j] - u[0], VY[j] - u[1], VZ[j] - u[2] };
            const double c = F2 * vu;
            const double fp[3] = { VX[j] * c + vmu[0] * F1, VY[j] * c + vmu[1] * F1, VZ[j] * c + vmu[2] * F1 };
            f[j] += LBM_DT * omegaf * W[j] * (fp[0] * totalForce[0] + fp[1] * totalForce[1] + fp[2] * totalForce[2]);
        }
    }
}

/* LB::streaming (LB.cpp:1144-1488) */
static void streaming(LboState* s, double* wallFHydro, uint32_t nWalls) {
    const double C2x2 = 9.0 * LBM_DT * LBM_DT * LBM_DT * LBM_DT, C3x2 = 3.0 * LBM_DT * LBM_DT;
    const double S1 = s->p.slipCoefficient, S2 = (1.0 - s->p.slipCoefficient);
    const double BBCoeff = 2.0 * 3.0 * LBM_DT * LBM_DT;
    double staticPres[Q];
    for (int j = 0; j < Q; ++j) staticPres[j] = s->initDensity * W[j];
    double extraMass = 0.0;
    double* wallF = (double*)calloc((size_t)3 * (nWalls ? nWalls : 1), sizeof(double));
    const uint32_t N = s->N;
    /* node::store */
#pragma omp parallel for schedule(static) num_threads(g_threads)
    for (uint32_t i = 0; i < N; ++i)
        if (isActiveT(T(s, i))) memcpy(&s->fs[(size_t)Q * i], &s->f[(size_t)Q * i], Q * sizeof(double));
    /* sequential: extraMass and wall forces are order-dependent sums */
    for (uint32_t it = 0; it < N; ++it) {
        if (!isActiveT(T(s, it))) continue;
        double* f = &s->f[(size_t)Q * it];
        const double* fs = &s->fs[(size_t)Q * it];
        const double* u = &s->u[3 * it];
        for (int j = 1; j < Q; ++j) {
            const uint32_t link = nbr(s, it, j);
            const int tl = T(s, link);
            if (isActiveT(tl)) {
                f[OPP[j]] = s->fs[(size_t)Q * link + OPP[j]];
            } else if (tl == LBO_GAS) {
                const double usq = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
                const double vuj = u[0] * VX[j] + u[1] * VY[j] + u[2] * VZ[j];
                f[OPP[j]] = -fs[j] + W[j] * s->initDensity * (2.0 + C2x2 * (vuj * vuj) - C3x2 * usq);
            } else if (tl == LBO_CURVED) {
                /* Mei-Luo-Shyy (LB.cpp:1278-1319), curve::getChi / computeCoefficients (node.cpp:458-472) */
                if (!s->curveRow || s->curveRow[link] == UINT32_MAX) {
                    fprintf(stderr, "lb_oracle: curved wall cell %u without curve data\n", link);
                    exit(3);
                }
                const double C1 = 3.0 * LBM_DT * LBM_DT, C2 = 4.5 * LBM_DT * LBM_DT * LBM_DT * LBM_DT, C3 = 1.5 * LBM_DT * LBM_DT;
                const double* vel = &s->u[3 * link];
                const double BBi = BBCoeff * s->n[it] * W[j] * (vel[0] * VX[j] + vel[1] * VY[j] + vel[2] * VZ[j]);
                const double delta = s->curveDelta[(size_t)Q * s->curveRow[link] + OPP[j]];
                const double tau = 0.5 + 3.0 * s->visc[it];
                const double chi = (delta >= 0.5) ? (2.0 * delta - 1.0) / tau : (2.0 * delta - 1.0) / (tau - 2.0);
                double ubf[3] = { 0.0, 0.0, 0.0 };
                if (delta >= 0.5) {
                    const double m1 = (delta - 1) / delta, m2 = 1 / delta;
                    for (int k = 0; k < 3; ++k) ubf[k] = m1 * u[k] + m2 * vel[k];
                } else {
                    const uint32_t linkBack = nbr(s, it, OPP[j]);
                    const double* src = (T(s, linkBack) == LBO_FLUID && (s->type[linkBack] & LBO_NODE_BIT)) ? &s->u[3 * linkBack] : vel;
                    for (int k = 0; k < 3; ++k) ubf[k] = src[k];
                }
                const double usq = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
                const double vu = u[0] * VX[j] + u[1] * VY[j] + u[2] * VZ[j];
                const double fStar = s->n[it] * W[j] * (1 + C1 * (VX[j] * ubf[0] + VY[j] * ubf[1] + VZ[j] * ubf[2]) + C2 * vu * vu + C3 * usq);
                f[OPP[j]] = (1.0 - chi) * fs[j] + chi * fStar - BBi;
                extraMass += s->mass[it] * (chi * fs[j] - chi * fStar + BBi);
            } else if (tl == LBO_DYN_WALL) {
                const double* vel = &s->u[3 * link];
                const double BBi = BBCoeff * s->n[it] * W[j] * (vel[0] * VX[j] + vel[1] * VY[j] + vel[2] * VZ[j]);
                /* node::bounceBackForce (node.cpp:283-291) */
                const double sc = (2.0 * (fs[j] - staticPres[j]) - BBi) * LBM_DT;
                const uint32_t w = s->solidIndex[link];
                if (w < nWalls) { wallF[3 * w] += VX[j] * sc; wallF[3 * w + 1] += VY[j] * sc; wallF[3 * w + 2] += VZ[j] * sc; }
                f[OPP[j]] = fs[j] - BBi;
                extraMass += BBi * s->mass[it];
            } else if (tl == LBO_STAT_WALL) {
                f[OPP[j]] = fs[j];
            } else if (tl == LBO_SLIP_STAT) {
                if (j > 6) {
                    const uint32_t c1 = nbr(s, it, SLIP1CHECK[j]), c2 = nbr(s, it, SLIP2CHECK[j]);
                    const int a1 = isActiveT(T(s, c1)), a2 = isActiveT(T(s, c2));
                    if (a1 && !a2) f[OPP[j]] = S1 * s->fs[(size_t)Q * c1 + SLIP1[j]] + S2 * fs[j];
                    else if (!a1 && a2) f[OPP[j]] = S1 * s->fs[(size_t)Q * c2 + SLIP2[j]] + S2 * fs[j];
                    else f[OPP[j]] = fs[j];
                } else f[OPP[j]] = fs[j];
            } else if (tl == LBO_SLIP_DYN) {
                const double* vel = &s->u[3 * link];
                const double BBi = BBCoeff * s->n[it] * W[j] * (vel[0] * VX[j] + vel[1] * VY[j] + vel[2] * VZ[j]);
                if (j > 6) {
                    const uint32_t c1 = nbr(s, it, SLIP1CHECK[j]), c2 = nbr(s, it, SLIP2CHECK[j]);
                    const int a1 = isActiveT(T(s, c1)), a2 = isActiveT(T(s, c2));
                    if (a1 && !a2) {
                        f[OPP[j]] = S1 * s->fs[(size_t)Q * c1 + SLIP1[j]] + S2 * (fs[j] - BBi);
                        extraMass += S2 * s->mass[it] * BBi;
                    } else if (!a1 && a2) {
                        f[OPP[j]] = S1 * s->fs[(size_t)Q * c2 + SLIP2[j]] + S2 * (fs[j] - BBi);
                        extraMass += S2 * s->mass[it] * BBi;
                    } else {
                        f[OPP[j]] = fs[j] - BBi;
                        extraMass += s->mass[it] * BBi;
                    }
                } else {
                    f[OPP[j]] = fs[j] - BBi;
                    extraMass += s->mass[it] * BBi;
                }
            } else {
                fprintf(stderr, "lb_oracle: %u type %d TYPE ERROR\n", link, tl); /* LB.cpp:1458-1461 */
                exit(3);
            }
        }
    }
    /* redistributeMass(extraMass) over the interface cells (LB.cpp:1477) */
    uint32_t nIf = listInterface(s, s->listA);
    redistributeMass(s, extraMass, s->listA, nIf);
    if (wallFHydro)
        for (uint32_t w = 0; w < 3 * nWalls; ++w) wallFHydro[w] = wallF[w] * s->uForce;
    free(wallF);
}

void lbo_step(LboState* s, const LboParticle* parts, uint32_t nParts, const LboElement* elmts, uint32_t nElmts,
              double* FHydro, double* MHydro, double* fluidVolume, double* wallFHydro, uint32_t nWalls) {
    (void)nParts;
    reconstruction(s);
    computeHydroForces(s, parts, elmts, nElmts, FHydro, MHydro, fluidVolume);
    collision(s);
    streaming(s, wallFHydro, nWalls);
}
