/* lb_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, dense grid) of the per-timestep lattice-Boltzmann update of
 * gnomeCreative/hybird (LB.cpp / node.cpp / lattice.h).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg may use it; the CUDA product never links or calls it.
 *
 * Parity status: PINNED.  The reference ships no golden vectors, so this
 * restatement is pinned against the unmodified reference compiled by oracle/Makefile
 * (oracle/_ref/ref_harness): tests/test_oracle_vs_reference.py runs both on the cases of
 * oracle/cases.py and demands bit-identical populations, macroscopic fields, masses, type maps
 * and element forces (reference at OMP_NUM_THREADS=1), and tests/golden/ holds reference outputs
 * for boxes without /root/reference.
 */
#ifndef LB_ORACLE_H
#define LB_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cell type codes: node.h:71-84 */
enum { LBO_FLUID = 0, LBO_GAS = 2, LBO_INTERFACE = 3, LBO_PERIODIC = 4, LBO_SLIP_STAT = 5, LBO_SLIP_DYN = 6,
       LBO_STAT_WALL = 7, LBO_DYN_WALL = 8, LBO_CURVED = 9 };
#define LBO_TYPE_MASK 0x0F
#define LBO_P_BIT 0x10    /* nodeType::p "inside particle" */
#define LBO_NODE_BIT 0x20 /* a `node` object exists (IO.cpp:747 tests nodes[i]!=0) */

typedef struct {
    int32_t size[3];     /* lbSize incl. the boundary shell (LB.cpp:110-112) */
    int32_t boundary[6]; /* boundary0..5 (LB.cpp:175-181) */
    double lbF[3];       /* lattice units (LB.cpp:173) */
    double initDynVisc, plasticVisc, yieldStress, turbConst, slipCoefficient; /* lattice units */
    int32_t freeSurface, forceField, nonNewtonian, turbulence;               /* hybird.cpp:184-190 */
    double unitLength, unitTime, unitDensity;                                  /* LB.cpp:93-98 */
} LboParams;

typedef struct {
    double x0[3], r, radiusVec[3]; /* physical units, as DEM holds them (elmt.h:17-36) */
    uint32_t clusterIndex, particleIndex;
} LboParticle;

typedef struct {
    double x1[3], wGlobal[3]; /* physical units (elmt.h:70-90) */
    uint32_t compBegin, compEnd; /* range into the flattened components array */
} LboElement;

typedef struct LboState LboState;

/* type_flags: t | p<<4 | node<<5 per cell, reference index order i = x + X*(y + Y*z).
 * f may be NULL: populations are then set to the equilibrium of (n,u) like node::initialize. */
LboState* lbo_create(const LboParams* p, const uint8_t* type_flags, const uint32_t* solidIndex, const double* f,
                     const double* n, const double* u, const double* mass, const double* visc);
void lbo_destroy(LboState* s);
/* curved-wall cells (type 9): delta[19] of each (curve::delta, node.h:133-148), as LB::initializeCurved computed them */
void lbo_set_curves(LboState* s, uint32_t nCurves, const uint32_t* cells, const double* delta);
/* problemName == DRUM: LB::enforceMassConservation (LB.cpp:1806-1822) after every free-surface update */
void lbo_set_mass_target(LboState* s, double totalMass);

/* LB::latticeBoltzmannFreeSurfaceStep (LB.cpp:235-245) */
void lbo_free_surface_step(LboState* s);
/* LB::latticeBoltzmannCouplingStep (LB.cpp:247-280) */
void lbo_coupling_step(LboState* s, int newNeighborList, const LboParticle* parts, uint32_t nParts,
                       const LboElement* elmts, uint32_t nElmts, const uint32_t* components);
/* LB::latticeBolzmannStep (LB.cpp:221-233). Outputs in physical units; any may be NULL. */
void lbo_step(LboState* s, const LboParticle* parts, uint32_t nParts, const LboElement* elmts, uint32_t nElmts,
              double* FHydro, double* MHydro, double* fluidVolume, double* wallFHydro, uint32_t nWalls);

/* state access (arrays owned by the state; N = X*Y*Z; f/fs are [N][19], u/hydroForce [N][3]) */
uint32_t lbo_nodes(const LboState* s);
const uint8_t* lbo_type_flags(const LboState* s);
const uint32_t* lbo_solid_index(const LboState* s);
const double* lbo_f(const LboState* s);
const double* lbo_fs(const LboState* s);
const double* lbo_n(const LboState* s);
const double* lbo_u(const LboState* s);
const double* lbo_hydro_force(const LboState* s);
const double* lbo_mass(const LboState* s);
const double* lbo_visc(const LboState* s);
const double* lbo_shear_rate(const LboState* s);
/* neighbour table entry d[j] of cell i as the reference builds it (LB.cpp:377-472) */
uint32_t lbo_neighbor(const LboState* s, uint32_t i, int j);
uint32_t lbo_count_type(const LboState* s, int t);
void lbo_set_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
