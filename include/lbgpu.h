/* lbgpu.h -- C ABI of the B200 lattice-Boltzmann engine (liblbgpu.so).
 *
 * The reference (gnomeCreative/hybird) has no FFI: its boundary for the LB hot path is the C++
 * class `LB` (LB.h:32-215), called from goCycle (hybird.cpp:35-66) and read by IO (IO.cpp).
 * Each entry point below replaces one of those calls; the reference-side binding is the
 * `LB` shim in hybird_b200/shim/LB_gpu.cpp.
 *
 *   lbGpuInit            <- LB::latticeBolzmannInit        (LB.h:158, LB.cpp:190-219) state upload
 *   lbGpuStep            <- LB::latticeBoltzmannFreeSurfaceStep (LB.h:161, LB.cpp:235-245)
 *                           LB::latticeBoltzmannCouplingStep    (LB.h:160, LB.cpp:247-280)
 *                           LB::latticeBolzmannStep             (LB.h:159, LB.cpp:221-233)
 *                           in the order goCycle issues them (hybird.cpp:49-57)
 *   lbGpuParticleForces  <- elmts[].FHydro/MHydro/fluidVolume written by LB::computeHydroForces
 *                           (LB.cpp:1851-1919) and walls[].FHydro written by LB::streaming
 *                           (LB.cpp:1321-1341,1485-1487)
 *   lbGpuFetchFields     <- IO's direct reads of lb.types[i], lb.nodes[i]->{n,u,visc,mass,
 *                           shearRate} (IO.cpp:698-831, 835-895)
 *
 * Conventions: plain C, POD only, no exceptions cross the boundary.  Every function returns 0 on
 * success or a negative LBGPU_E* code; lbGpuLastError() gives the message.  All host arrays are
 * in the reference's cell order i = x + X*(y + Y*z) (LB.cpp:2317-2337) including the 1-cell
 * boundary shell.  The caller owns host buffers; the library owns device memory.  Calls must come
 * from one host thread per handle.  There is no CPU fallback: without a CUDA device every entry
 * point fails with LBGPU_ENODEVICE.
 */
#ifndef LBGPU_H
#define LBGPU_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LBGPU_ABI_VERSION 1

enum {
    LBGPU_OK = 0,
    LBGPU_EINVAL = -1,      /* bad argument */
    LBGPU_ENODEVICE = -2,   /* no CUDA device / driver */
    LBGPU_ECUDA = -3,       /* CUDA runtime error */
    LBGPU_EUNSUPPORTED = -4,/* configuration outside what the engine handles (slab axis, list capacities) */
    LBGPU_ETYPE = -5,       /* "TYPE ERROR" of LB::streaming (LB.cpp:1458-1461): link to an illegal cell type */
    LBGPU_ECOMM = -6        /* multi-GPU halo transport error */
};

/* cell-type byte: nodeType::t (node.h:71-84) in the low nibble plus flag bits */
#define LBGPU_TYPE_MASK 0x0F
#define LBGPU_P_BIT 0x10    /* nodeType::p, "inside particle" (node.h:86) */
#define LBGPU_NODE_BIT 0x20 /* lb.nodes[i] != 0 (IO.cpp:747) */

typedef struct {
    int32_t size[3];      /* lbSize[0..2] incl. shell (LB.cpp:110-112) */
    int32_t boundary[6];  /* boundary0..5 (LB.cpp:175-181): 4 periodic, 5/6 slip stat/dyn, 7/8 no-slip stat/dyn */
    double lbF[3];        /* body force, lattice units (LB.cpp:173) */
    double initDynVisc;   /* lattice units (LB.cpp:140) */
    double plasticVisc;   /* lattice units (LB.cpp:141) */
    double yieldStress;   /* lattice units (LB.cpp:142) */
    double turbConst;     /* LB.cpp:185 */
    double slipCoefficient; /* LB.cpp:183 */
    int32_t freeSurface, forceField, nonNewtonian, turbulence; /* hybird.cpp:184-190 */
    double unitLength, unitTime, unitDensity; /* LB.cpp:93-98; composites as measureUnits::setComposite */
    int32_t nWalls;       /* dem.walls.size(): length of the wall-force output */
    int32_t device;       /* CUDA device ordinal; -1 = current device */
    /* slab decomposition: the interior planes 1..size[2]-2 are cut along z (slabAxis = 2, the slowest
     * index, so halo planes are contiguous) into nSlabs slabs (lbGpuSlabRange). This handle owns slabs
     * [slabIndex, slabIndex + nLocalSlabs); its host arrays cover the planes of those slabs plus one
     * plane below and above.  nSlabs <= 1: the whole lattice in one slab. */
    int32_t slabAxis, nSlabs, slabIndex, nLocalSlabs;
    int32_t reserved[4];
} LbGpuParams;

typedef struct {
    double x0[3], r, radiusVec[3]; /* particle::x0, r, radiusVec in physical units (elmt.h:17-36) */
    uint32_t clusterIndex, particleIndex;
} LbGpuParticle;                  /* 64 bytes */

typedef struct {
    double x1[3], wGlobal[3];     /* elmt::x1, wGlobal in physical units (elmt.h:70-90) */
    uint32_t compBegin, compEnd;  /* elmt::components as a range of the flattened `components` array */
} LbGpuElement;                   /* 56 bytes */

typedef struct LbGpuHandle LbGpuHandle;

const char* lbGpuLastError(void);
int lbGpuAbiVersion(void);
int lbGpuDeviceCount(void);
/* global planes [zBegin, zEnd) owned by slab `slab` of `nSlabs` for a lattice of sizeZ planes */
int lbGpuSlabRange(int32_t sizeZ, int32_t nSlabs, int32_t slab, int32_t* zBegin, int32_t* zEnd);

/* Multi-GPU: one process per GPU, each owning nLocalSlabs consecutive slabs (rank r: slabs [r*nLocalSlabs,
 * (r+1)*nLocalSlabs)).  The reference is a single process (no communication backend), so these
 * entry points replace nothing upstream; they carry the face halo of SURVEY.md 8e over NCCL send/recv and the
 * three small sums (mass surplus, interface count, per-element forces) over ncclAllReduce.  Rank 0 obtains the
 * id, the host runtime (torch.distributed in hybird_b200/slabs.py, MPI in a C++ driver) broadcasts it, every rank
 * calls lbGpuCommInit BEFORE lbGpuInit.  NCCL is loaded at run time (libnccl.so.2; override with LBGPU_NCCL_LIB). */
int lbGpuCommUniqueId(uint8_t id[128]);
int lbGpuCommInit(const uint8_t id[128], int32_t rank, int32_t world, int32_t device);
int lbGpuCommInfo(int32_t* rank, int32_t* world, int32_t* ncclVersion);
int lbGpuCommFinalize(void);
/* *on = 1 when this handle's per-step halo goes through peer memory: every rank maps the arrays of its neighbour ranks
 * (cudaIpc over NVLink / NVSwitch) and one put kernel per step stores the face planes straight into the neighbours'
 * ghost planes and raises a flag there; 0: NCCL send/recv carries it (no peer access between the GPUs, or
 * LBGPU_PEER_HALO=0) -- lbGpuLastError() then says why. */
int lbGpuPeerHalo(LbGpuHandle* h, int32_t* on);

/* Upload the state LB::latticeBolzmannInit produced.
 *   type_flags  N bytes: t | p<<4 | node<<5
 *   solidIndex  N
 *   f           19*N doubles, cell-major ([i][j] like node::f) or NULL => equilibrium of (n,u) as node::initialize
 *   n,u,mass,visc  N, 3N (cell-major), N, N; read where the node bit is set (wall cells carry the wall velocity in u)
 */
int lbGpuInit(const LbGpuParams* params, const uint8_t* type_flags, const uint32_t* solidIndex, const double* f,
              const double* n, const double* u, const double* mass, const double* visc, LbGpuHandle** out);

/* The same state built ON THE DEVICE for a box problem -- a lattice bounded by its six planes (problemName NONE /
 * demChute): LB::latticeBolzmannInit's cell types, DEM wall indices, particle flags, gas region with the interface
 * closure, hydrostatic density, initial velocity, masses and wall nodes (LB.cpp:190-219, 324-996; DEM.cpp:435-640)
 * computed per cell, instead of by the reference's serial O(N x n_geom) host loops and a 35 GB host mirror at 67 M
 * cells.  Bit-identical to uploading the host-built state.  Works per slab: with a communicator every rank
 * initialises its own planes (the hydrostatic reference height is max-reduced over the ranks).
 *   initVelocity   lattice units (LB.cpp:143)
 *   wallVelocity   6 x 3, physical units, velocity of the DEM wall of boundary k (used where boundary k is 6 or 8); may be NULL
 *   regions        applied in order to fluid cells: a cell inside (gasInside) / outside (!gasInside) becomes gas
 *                  kind 0: box a0<=x<=a1, a2<=y<=a3, a4<=z<=a5;  1: sphere centre (a0,a1,a2) radius a3 (strictly inside);
 *                  2: half space z > a0 (demChute: a0 = 0.025/unit.Length + 0.5, LB.cpp:612-627)       [lattice units]
 *   parts          particles in physical units as in lbGpuStep (LB::initializeParticleBoundaries) */
typedef struct { int32_t kind, gasInside; double a[6]; } LbGpuRegion;
int lbGpuInitBox(const LbGpuParams* params, const double initVelocity[3], const double* wallVelocity, const LbGpuRegion* regions,
                 uint32_t nRegions, const LbGpuParticle* parts, uint32_t nParts, LbGpuHandle** out);

/* Curved walls (type 9; problem geometries with a cylinder: DRUM / AVALANCHE / NET).  LB::curves (LB.h:63-64) as
 * LB::initializeCurved (LB.cpp:589-603) left it: cells[k] is the cell index of the k-th `curve` object in the host
 * arrays' order, delta[19*k + j] its curve::delta[j] (node.h:133-148; m1, m2 and chi follow from delta, node.cpp:458-472).
 * Must be called after lbGpuInit and before the first step of a lattice that holds type-9 cells; a slab handle passes
 * the cells of its own planes (indices relative to its host arrays), the others are ignored. */
int lbGpuSetCurves(LbGpuHandle* h, uint32_t nCurves, const uint32_t* cells, const double* delta /* 19*nCurves */);

/* problemName == DRUM: LB::enforceMassConservation (LB.cpp:1806-1822) runs after every free-surface update with
 * LB::totalMass = totalMass (lattice units, LB.cpp:204-218). */
int lbGpuSetMassTarget(LbGpuHandle* h, double totalMass);

/* One goCycle worth of LB work, in the reference's order: free-surface step (if doFreeSurface),
 * coupling step (if doCoupling; rescanParticles = dem.newNeighborList), LB step.
 * parts/elmts/components may be NULL when nParts == 0. Asynchronous on the handle's stream. */
int lbGpuStep(LbGpuHandle* h, int doFreeSurface, int doCoupling, int rescanParticles, const LbGpuParticle* parts,
              uint32_t nParts, const LbGpuElement* elmts, uint32_t nElmts, const uint32_t* components,
              uint32_t nComponents);

/* The coupling step alone (goCycle's branch for demTime <= demInitialRepeat, hybird.cpp:60-64: particle flags
 * follow the particles while the fluid is not stepped yet). */
int lbGpuCouple(LbGpuHandle* h, int rescanParticles, const LbGpuParticle* parts, uint32_t nParts, const LbGpuElement* elmts,
                uint32_t nElmts, const uint32_t* components, uint32_t nComponents);

/* `count` cycles back to back without a host round trip. The particle state of the last lbGpuStep/lbGpuCouple upload
 * (none: pure-fluid / free-surface runs) stays resident and fixed: the coupling step (particle flags, without rescan)
 * and the hydrodynamic force reduction run every cycle as in goCycle (hybird.cpp:49-57). */
int lbGpuRun(LbGpuHandle* h, int doFreeSurface, uint32_t count);

/* DEM sub-steps on the device: DEM::discreteElementStep (DEM.cpp:331-376) for single-sphere elements
 * bounded by plane walls -- neighbour table + wall table on the displacement trigger (DEM.cpp:1314-1324, 1377-1513),
 * 5th-order Gear predictor / corrector (elmt.cpp:139-254), LINEAR / HERTZIAN particle-particle and particle-wall contacts
 * (DEM.cpp:1668-1717, 1801-1982, 2138-2224), Newton's equations with the hydrodynamic force and torque of the last LB step
 * (DEM.cpp:1150-1181).  The particle / element lists of the coupling step and of LB::computeHydroForces are refreshed on
 * the device, so a coupled cycle has no host round trip.  Elements are single spheres or clusters of 2-4 spheres; with
 * periodic DEM boundaries (single spheres) the ghost particles of DEM::createGhosts are rebuilt with the tables, the coupling
 * step rescans after a rebuild as dem.newNeighborList asks (DEM.cpp:1414), and the host reads the particle count back once
 * per DEM step.  Not covered: cylinders, objects -- those keep the host DEM and lbGpuStep.  All values in physical units, as the
 * reference's DEM holds them.  With a communicator every rank advances the same (replicated) elements.
 *   contactModel  0 LINEAR, 1 HERTZIAN (material::contactModel, DEM.cpp:150-158)
 *   deltat, multiStep, nebrRange, maxDisp: DEM::deltat / multiStep / nebrRange / maxDisp after DEM::discreteElementInit */
typedef struct {
    int32_t contactModel, multiStep;
    double knConst, ksConst, dampCoeff, viscTang, linearStiff, frictionCoefPart, frictionCoefWall, numVisc;
    double demF[3], deltat, nebrRange, maxDisp;
    /* periodic DEM boundaries (DEM::initializePbcs, DEM.cpp:937-988): pbc::p and pbc::v of each; single spheres only */
    int32_t nPbc, pad;
    double pbcP[3][3], pbcV[3][3];
} LbGpuDemParams;
typedef struct { double x0[3], x1[3], w0[3], radius, m, I[3]; int32_t size, pad; } LbGpuDemElement;
/* elmt::x0, x1, w0, radius, m, I, size (elmt.h:60-110).  size 1: a sphere; 2-4: the reference's clusters (DEM::compositeProperties,
 * DEM.cpp:404-433) with their orientation quaternion (q0 = identity, q1 = 0 as in the particle file of a fresh run); I is the
 * inertia elmt::initialize computed (principal axes, transport terms included); 0 is read as 1. */
typedef struct { double n[3], p[3], vel[3], omega[3], rotCenter[3]; int32_t moving, pad; } LbGpuDemWall; /* wall.h */
int lbGpuDemInit(LbGpuHandle* h, const LbGpuDemParams* params, const LbGpuDemElement* elmts, uint32_t nElmts,
                 const LbGpuDemWall* walls, uint32_t nWalls);
/* One DEM::discreteElementStep (multiStep sub-steps).  hydro = NULL: FHydro / MHydro of the last LB step as they lie on the
 * device; otherwise 7 doubles per element {FHydro, MHydro, -} from the host (tests). */
int lbGpuDemStep(LbGpuHandle* h, const double* hydro);
/* `count` goCycles (hybird.cpp:35-66) back to back on the device: DEM step, free-surface step (if doFreeSurface), coupling
 * step, LB step. */
int lbGpuRunDem(LbGpuHandle* h, int doFreeSurface, uint32_t count);
/* elmt::x0, x1, w0 (3*nElmts each; any may be NULL); info[0] = DEM::maxDisp, info[1] = neighbour-table rebuilds so far,
 * info[2] = longest partner list.  Synchronises. */
int lbGpuDemState(LbGpuHandle* h, double* x0, double* x1, double* w0, double info[3]);
/* elmt::FParticle, FWall, MParticle, MWall of the last sub-step (3*nElmts each; any may be NULL): what IO::exportForces and
 * the particle files print (IO.cpp:930-938).  Synchronises. */
int lbGpuDemContacts(LbGpuHandle* h, double* FParticle, double* FWall, double* MParticle, double* MWall);
/* the particles (spheres) of the elements as particle::updateCorrected left them: *nParticles = their number; x0, radiusVec
 * (3 per particle) and clusterIndex may be NULL.  Synchronises. */
int lbGpuDemParticles(LbGpuHandle* h, uint32_t* nParticles, double* x0, double* radiusVec, uint32_t* clusterIndex);

/* Results of the last step in physical units; any pointer may be NULL. Synchronises. */
int lbGpuParticleForces(LbGpuHandle* h, double* FHydro /*3*nElmts*/, double* MHydro /*3*nElmts*/,
                        double* fluidVolume /*nElmts*/, double* wallFHydro /*3*nWalls*/);

/* Host mirrors for IO; any pointer may be NULL. f is cell-major post-collision populations
 * (node::fs after the last step). Synchronises. */
int lbGpuFetchFields(LbGpuHandle* h, uint8_t* type_flags, uint32_t* solidIndex, double* n, double* u, double* mass,
                     double* visc, double* shearRate, double* hydroForce, double* f);

/* Output path.
 * lbGpuFluidSummary: what IO's screen export derives from the whole lattice, as one device reduction instead of a
 * full-field fetch: out[0] = max |u|^2 over the active cells (IO::exportMaxSpeedFluid, IO.cpp:835-851; lattice units,
 * the caller takes the root), out[1] = sum of mass over the active cells outside particles (IO::totFluidMass,
 * IO.cpp:987-999), out[2] = active cells, out[3] = active cells with visc > 0.95 maxVisc (IO::totPlastic,
 * IO.cpp:969-985).  Over all ranks of the communicator.
 * lbGpuWriteVti: the file of IO::exportParaviewFluidOld (IO.cpp:698-831) -- same extents, spacing, array names
 * (type, v, pressure, dynVisc, AAAmass, solidIndex), types, order and values in physical units -- with the data as raw
 * appended binary instead of formatted text.  withSolidIndex = IO::demSolve.  One process. */
int lbGpuFluidSummary(LbGpuHandle* h, double out[4]);
int lbGpuWriteVti(LbGpuHandle* h, const char* path, int withSolidIndex);

/* Checkpoint / restart of the fluid state (the reference has none: SURVEY.md 5; DEM's recycle file covers the particles,
 * IO.cpp:486-535).  The blob holds the dynamic device state (populations, types, macroscopic fields, masses, resident
 * particle lists, step counter); it is loaded into a handle created by lbGpuInit with the SAME parameters and wall
 * geometry (+ lbGpuSetCurves / lbGpuSetMassTarget where used) -- the initial field values of that handle do not matter.
 * A run continued from a loaded state is bit-identical to the uninterrupted run.  Call between cycles only. */
int lbGpuStateBytes(LbGpuHandle* h, uint64_t* bytes);
int lbGpuSaveState(LbGpuHandle* h, void* buffer, uint64_t bytes);
int lbGpuLoadState(LbGpuHandle* h, const void* buffer, uint64_t bytes);

/* counters: [0]=fluid cells, [1]=interface cells, [2]=cells with p flag, [3]=LB steps done */
int lbGpuCounts(LbGpuHandle* h, uint64_t counts[4]);
/* the same over the planes this handle owns only (lbGpuCounts sums over the ranks of the communicator) */
int lbGpuCountsLocal(LbGpuHandle* h, uint64_t counts[4]);
int lbGpuSynchronize(LbGpuHandle* h);
/* device time of the kernels launched by the last lbGpuStep/lbGpuRun call, milliseconds (CUDA events) */
int lbGpuLastStepMs(LbGpuHandle* h, float* ms);
/* device time of the fused stream-collide kernel alone: sum over the (at most 512 most recent)
 * launches of the last lbGpuStep/lbGpuRun call, CUDA events on the engine's stream. Synchronises. */
int lbGpuLastKernelMs(LbGpuHandle* h, float* msSum, uint32_t* launches);
/* Where the device time of a cycle goes: with the trace on, CUDA events are recorded at the phase boundaries of every cycle
 * (and lbGpuRun launches its cycles one by one); lbGpuPhaseMs averages over the last (at most 64) cycles:
 * ms[0] DEM sub-steps, [1] list build of the free-surface update, [2] free-surface update, [3] coupling step,
 * [4] step kernels + step halo, [5] wall slots / moving-wall sums, [6] element forces + type sync.  Synchronises. */
int lbGpuPhaseTrace(LbGpuHandle* h, int on);
int lbGpuPhaseMs(LbGpuHandle* h, float ms[7], uint32_t* cycles);
/* lbGpuRun replays two consecutive cycles of a free-surface lattice without particles as one CUDA graph (single process;
 * LBGPU_GRAPH=0 turns it off): info[0] = captures, info[1] = replays (of two cycles each) so far */
int lbGpuGraphInfo(LbGpuHandle* h, uint64_t info[2]);
/* number of kernels this handle launched so far (kernels inside a replayed graph included) */
int lbGpuLaunchCount(LbGpuHandle* h, uint64_t* launches);
/* device self-test of the engine's shared-reciprocal fp64 division against IEEE division on `count`
 * pseudo-random operand pairs: result[0] = mismatches (must be 0), [1] = quotients compared, [2] = operand
 * pairs outside the fast-path domain (handled by the ordinary division in the kernels) */
int lbGpuSelfTest(uint64_t count, uint64_t seed, uint64_t result[3]);
int lbGpuFinalize(LbGpuHandle* h);

#ifdef __cplusplus
}
#endif
#endif
