// lb_dem.cuh -- the DEM sub-steps of a coupled cycle on the device: elements of one to four spheres
// (single spheres and the reference's clusters) between plane walls.
//
// What DEM::discreteElementStep (DEM.cpp:331-376) does between two LB steps, restated per element:
//   k_dem_trigger          DEM::evalMaxDisp + the neighbour-table trigger (DEM.cpp:1314-1324, 340-346)
//   k_dem_neighbours       when triggered: DEM::evalNeighborTable (DEM.cpp:1377-1494) as one partner list per PARTICLE -- every
//                          particle of another element whose centre is closer than nebrRange NOW, in ascending index order
//                          (k_grid_* + k_dem_neighbours_grid: the same lists through a uniform grid, for large beds)
//   k_dem_predict          DEM::evalNearWallTable when triggered (DEM.cpp:1496-1513: the FIRST wall within nebrRange of a
//                          particle, at the corrected position), elmt::predict (elmt.cpp:139-177: positions, orientation
//                          quaternions, spins), particle::updatePredicted (elmt.cpp:261-275)
//   k_dem_forces_correct   particle-particle and wall-particle contacts (DEM.cpp:1668-1717, 1801-1982) with the LINEAR /
//                          HERTZIAN laws (DEM.cpp:2138-2224) and the lever arms of non-spherical elements, Newton's equations
//                          in the body frame (DEM.cpp:1150-1181), elmt::correct (elmt.cpp:179-254), particle::updateCorrected
//   k_dem_export           the particle / element lists LB::latticeBoltzmannCouplingStep and LB::computeHydroForces read
// The hydrodynamic force and torque come straight from the LB step's per-element reduction (physical units), the
// positions and velocities go straight into the coupling step: a coupled cycle has no host round trip for the particles.
// The contact laws are memoryless (no tangential spring), and both threads of a pair evaluate the contact with the lower
// particle index as "I", so the pair's force is identical on both sides and every element sums the contacts of its particles
// with their partners in ascending index order: deterministic, no atomics.  Broad phase: the reference searches its table with
// linked cells (DEM.cpp:1326-1375) whose width is at least nebrRange wherever the domain is wider than six radii, so its
// table IS the set of pairs within nebrRange at rebuild time; here a rebuild compares all pairs through shared-memory
// tiles (small beds) or bins the particles into a uniform grid (k_grid_*, large beds) and the sub-steps in between only walk
// the partner lists.  Periodic boundaries (ghost particles) for single spheres (k_dem_pbc).  No cylinders, objects.
// Compiled with -fmad=false like the LB kernels: the reference's operation order is kept.
#pragma once
#include <stdint.h>

namespace lbdem {

struct Params {
    int contactModel, multiStep;
    double knConst, ksConst, dampCoeff, viscTang, linearStiff, frictionCoefPart, frictionCoefWall, numVisc;
    double demF[3], deltat, nebrRange;
    double c[5], coeff1[6], coeff2[6];  // DEM::predictor / DEM::corrector constants (DEM.cpp:1067-1112), computed on the host
    double proto[5][4][3];              // DEM::compositeProperties (DEM.cpp:404-433): sphere i of an element of `size`, unit = radius
    // periodic DEM boundaries (DEM::initializePbcs, DEM.cpp:937-988; pbc::setPlanes, utils.cpp:240-245): plane 1 through p with
    // normal v / |v|, plane 2 through p + v with the opposite normal
    int nPbc, padPbc;
    double pbcP[3][3], pbcV[3][3], pbcN[3][3];
};
struct Wall { double n[3], p[3], vel[3], omega[3], rotCenter[3]; int moving, pad; };
struct Elmt {
    double x[6][3], xp[6][3], w[6][3], wp[6][3];
    double q[6][4], qp[6][4];  // orientation (elmt::q0..q5): the identity and zeros for a single sphere, never touched then
    double radius, m, I[3];
    double fc[4][3];  // FParticle, FWall, MParticle, MWall of the last sub-step (IO::exportForces reads the first two)
    int size, pBegin;  // its particles: [pBegin, pBegin + size)
};
// one sphere of an element: corrected (c) and predicted (p) centre, lever arm (radiusVec) and velocity
struct Part {
    double xc[3], rvc[3], xp[3], rvp[3], vp[3];
    double shift[3];  // a ghost particle (periodic boundaries): its origin's centre + shift; zero for a standard particle
    int cluster, proto, nearWall, pad;
};
struct V3 { double x, y, z; };
struct Q4 { double q0, q1, q2, q3; };
__device__ __forceinline__ V3 v3(const double* a) { return { a[0], a[1], a[2] }; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
__device__ __forceinline__ V3 operator-(V3 a) { return { -a.x, -a.y, -a.z }; }
__device__ __forceinline__ V3 operator*(V3 a, double s) { return { a.x * s, a.y * s, a.z * s }; }
__device__ __forceinline__ V3 operator*(double s, V3 a) { return { s * a.x, s * a.y, s * a.z }; }
__device__ __forceinline__ V3 operator/(V3 a, double s) { return { a.x / s, a.y / s, a.z / s }; }
__device__ __forceinline__ double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ double norm2(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
__device__ __forceinline__ void put(double* d, V3 a) { d[0] = a.x; d[1] = a.y; d[2] = a.z; }
// quaternions as the reference writes them (vector.cpp:269-310, 463-504)
__device__ __forceinline__ Q4 q4(const double* a) { return { a[0], a[1], a[2], a[3] }; }
__device__ __forceinline__ void putq(double* d, Q4 a) { d[0] = a.q0; d[1] = a.q1; d[2] = a.q2; d[3] = a.q3; }
__device__ __forceinline__ Q4 operator+(Q4 a, Q4 b) { return { a.q0 + b.q0, a.q1 + b.q1, a.q2 + b.q2, a.q3 + b.q3 }; }
__device__ __forceinline__ Q4 operator-(Q4 a, Q4 b) { return { a.q0 - b.q0, a.q1 - b.q1, a.q2 - b.q2, a.q3 - b.q3 }; }
__device__ __forceinline__ Q4 operator*(Q4 a, double s) { return { a.q0 * s, a.q1 * s, a.q2 * s, a.q3 * s }; }
__device__ __forceinline__ Q4 operator*(double s, Q4 a) { return { s * a.q0, s * a.q1, s * a.q2, s * a.q3 }; }
__device__ __forceinline__ Q4 qadj(Q4 a) { return { a.q0, -a.q1, -a.q2, -a.q3 }; }
__device__ __forceinline__ Q4 qmul(Q4 q, Q4 r) {  // q.multiply(r)
    return { r.q0 * q.q0 - r.q1 * q.q1 - r.q2 * q.q2 - r.q3 * q.q3, r.q0 * q.q1 + r.q1 * q.q0 - r.q2 * q.q3 + r.q3 * q.q2,
             r.q0 * q.q2 + r.q1 * q.q3 + r.q2 * q.q0 - r.q3 * q.q1, r.q0 * q.q3 - r.q1 * q.q2 + r.q2 * q.q1 + r.q3 * q.q0 };
}
__device__ __forceinline__ Q4 qnormalize(Q4 a) {
    const double n = sqrt(a.q0 * a.q0 + a.q1 * a.q1 + a.q2 * a.q2 + a.q3 * a.q3);
    return { a.q0 / n, a.q1 / n, a.q2 / n, a.q3 / n };
}
__device__ __forceinline__ V3 project(V3 v, Q4 q) {  // quat2vec(q (0, v) q*)
    const Q4 r = qmul(qmul(q, Q4{ 0.0, v.x, v.y, v.z }), qadj(q));
    return { r.q1, r.q2, r.q3 };
}

// scal[0] = maxDisp, flag[0] = rebuild the tables in this sub-step, flag[2] = rebuilds so far
__global__ void __launch_bounds__(1024) k_dem_trigger(const Elmt* __restrict__ e, uint32_t n, double deltat, double nebrRange,
                                                      double* __restrict__ scal, uint32_t* __restrict__ flag) {
    __shared__ double sm[32];
    double mx = 0.0;
    for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) mx = fmax(mx, norm2(v3(e[k].x[1])));
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 32; ++k) mx = fmax(mx, sm[k]);
        double md = scal[0] + sqrt(mx) * deltat;
        const bool rebuild = md > 0.25 * nebrRange;
        if (rebuild) md = 0.0;
        scal[0] = md;
        flag[0] = rebuild ? 1u : 0u;
        if (rebuild) flag[2] += 1u;
    }
}

// partner lists of the PARTICLES: nbr[a * MAX_NBR + q], q < nNbr[a], ascending particle indices of other elements.
// status[0] = largest list length seen (> MAX_NBR: error)
constexpr int MAX_NBR = 48;
// nPdev (may be null): the number of particles when it lives on the device (standard particles + the ghosts of this rebuild)
__global__ void __launch_bounds__(128) k_dem_neighbours(const Part* __restrict__ pt, uint32_t nP, double nebrRange, const uint32_t* __restrict__ flag,
                                                        uint32_t* __restrict__ nbr, uint32_t* __restrict__ nNbr, uint32_t* __restrict__ status,
                                                        const uint32_t* __restrict__ nPdev) {
    if (!*flag) return;
    if (nPdev) nP = *nPdev;
    if (blockIdx.x * blockDim.x >= nP) return;
    __shared__ double sx[128], sy[128], sz[128];
    __shared__ int sc[128];
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = a < nP;
    const V3 xa = live ? v3(pt[a].xc) : V3{ 0, 0, 0 };
    const int ca = live ? pt[a].cluster : -1;
    const double r2 = nebrRange * nebrRange;
    uint32_t cnt = 0;
    for (uint32_t base = 0; base < nP; base += 128) {
        const uint32_t j = base + threadIdx.x;
        __syncthreads();
        if (j < nP) { sx[threadIdx.x] = pt[j].xc[0]; sy[threadIdx.x] = pt[j].xc[1]; sz[threadIdx.x] = pt[j].xc[2]; sc[threadIdx.x] = pt[j].cluster; }
        __syncthreads();
        const uint32_t m = nP - base < 128u ? nP - base : 128u;
        if (live)
            for (uint32_t q = 0; q < m; ++q) {
                const V3 d = { sx[q] - xa.x, sy[q] - xa.y, sz[q] - xa.z };
                if (sc[q] != ca && norm2(d) < r2) {
                    if (cnt < (uint32_t)MAX_NBR) nbr[(size_t)a * MAX_NBR + cnt] = base + q;
                    ++cnt;
                }
            }
    }
    if (live) {
        nNbr[a] = cnt < (uint32_t)MAX_NBR ? cnt : (uint32_t)MAX_NBR;
        if (cnt > (uint32_t)MAX_NBR) atomicMax(status, cnt);
    }
}

// ---------------------------------------------------------------------------------------------
// The same partner lists through a uniform grid (the reference's linked cells, DEM.cpp:1326-1375, as a counting sort): cells
// at least nebrRange wide over the bounding box of the particles, particles binned by cell (count - scan - fill), every
// particle looks at the 27 cells around its own and sorts what it finds, so the lists are the ascending lists of the all-pairs
// pass -- identical results, O(n) instead of O(n^2): 20 000 spheres 1.7 ms -> 0.2 ms, and beds of 1e6 become possible.
// All launches are gated on the rebuild flag.
// ---------------------------------------------------------------------------------------------
struct Grid { double org[3], inv[3]; uint32_t dim[3], nCells; };
constexpr uint32_t GRID_MAX_CELLS = 1u << 21;

__global__ void __launch_bounds__(1024) k_grid_bounds(const Part* __restrict__ pt, uint32_t nP, double nebrRange, const uint32_t* __restrict__ flag,
                                                      Grid* __restrict__ g, uint32_t* __restrict__ cellCount, const uint32_t* __restrict__ nPdev) {
    if (!*flag) return;
    if (nPdev) nP = *nPdev;
    __shared__ double smin[3][32], smax[3][32];
    double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
    for (uint32_t k = threadIdx.x; k < nP; k += blockDim.x)
        for (int c = 0; c < 3; ++c) { lo[c] = fmin(lo[c], pt[k].xc[c]); hi[c] = fmax(hi[c], pt[k].xc[c]); }
    for (int c = 0; c < 3; ++c) {
        for (int o = 16; o > 0; o >>= 1) { lo[c] = fmin(lo[c], __shfl_down_sync(0xffffffffu, lo[c], o)); hi[c] = fmax(hi[c], __shfl_down_sync(0xffffffffu, hi[c], o)); }
        if ((threadIdx.x & 31) == 0) { smin[c][threadIdx.x >> 5] = lo[c]; smax[c][threadIdx.x >> 5] = hi[c]; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double width = nebrRange;
        for (;;) {  // cells of nebrRange, wider if that would be too many of them
            unsigned long long cells = 1;
            for (int c = 0; c < 3; ++c) {
                double a = smin[c][0], b = smax[c][0];
                for (int k = 1; k < 32; ++k) { a = fmin(a, smin[c][k]); b = fmax(b, smax[c][k]); }
                g->org[c] = a; g->inv[c] = 1.0 / width;
                g->dim[c] = (uint32_t)((b - a) / width) + 1u;
                cells *= g->dim[c];
            }
            if (cells <= GRID_MAX_CELLS) { g->nCells = (uint32_t)cells; break; }
            width *= 2.0;
        }
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k <= g->nCells; k += blockDim.x) cellCount[k] = 0u;
}
__device__ __forceinline__ uint32_t grid_cell(const Grid& g, const double* x, int* cx, int* cy, int* cz) {
    *cx = min((int)g.dim[0] - 1, max(0, (int)((x[0] - g.org[0]) * g.inv[0])));
    *cy = min((int)g.dim[1] - 1, max(0, (int)((x[1] - g.org[1]) * g.inv[1])));
    *cz = min((int)g.dim[2] - 1, max(0, (int)((x[2] - g.org[2]) * g.inv[2])));
    return (uint32_t)*cx + g.dim[0] * ((uint32_t)*cy + g.dim[1] * (uint32_t)*cz);
}
__global__ void __launch_bounds__(128) k_grid_count(const Part* __restrict__ pt, uint32_t nP, const uint32_t* __restrict__ flag, const Grid* __restrict__ g,
                                                    uint32_t* __restrict__ cellCount, uint32_t* __restrict__ cellOf, const uint32_t* __restrict__ nPdev) {
    if (!*flag) return;
    if (nPdev) nP = *nPdev;
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nP) return;
    int cx, cy, cz;
    const uint32_t c = grid_cell(*g, pt[k].xc, &cx, &cy, &cz);
    cellOf[k] = c;
    atomicAdd(&cellCount[c], 1u);
}
// exclusive scan of the cell counts in place (one block; cellCount[nCells] = n afterwards), fill counters zeroed
__global__ void __launch_bounds__(1024) k_grid_scan(const uint32_t* __restrict__ flag, const Grid* __restrict__ g, uint32_t* __restrict__ cellCount,
                                                    uint32_t* __restrict__ cellFill) {
    if (!*flag) return;
    __shared__ uint32_t sa[1024];
    const uint32_t nC = g->nCells, per = (nC + 1023u) / 1024u;
    const uint32_t b0 = threadIdx.x * per, b1 = min(nC, b0 + per);
    uint32_t s = 0;
    for (uint32_t c = b0; c < b1; ++c) s += cellCount[c];
    sa[threadIdx.x] = s;
    __syncthreads();
    for (uint32_t o = 1; o < 1024; o <<= 1) {
        uint32_t v = 0;
        if (threadIdx.x >= o) v = sa[threadIdx.x - o];
        __syncthreads();
        sa[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = sa[threadIdx.x] - s;
    for (uint32_t c = b0; c < b1; ++c) { const uint32_t v = cellCount[c]; cellCount[c] = run; cellFill[c] = 0u; run += v; }
    if (threadIdx.x == 1023) cellCount[nC] = sa[1023];
}
__global__ void __launch_bounds__(128) k_grid_fill(uint32_t nP, const uint32_t* __restrict__ flag, const uint32_t* __restrict__ cellStart,
                                                   uint32_t* __restrict__ cellFill, const uint32_t* __restrict__ cellOf, uint32_t* __restrict__ sorted,
                                                   const uint32_t* __restrict__ nPdev) {
    if (!*flag) return;
    if (nPdev) nP = *nPdev;
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nP) return;
    const uint32_t c = cellOf[k];
    sorted[cellStart[c] + atomicAdd(&cellFill[c], 1u)] = k;
}
__global__ void __launch_bounds__(128) k_dem_neighbours_grid(const Part* __restrict__ pt, uint32_t nP, double nebrRange, const uint32_t* __restrict__ flag,
                                                             const Grid* __restrict__ gp, const uint32_t* __restrict__ cellStart,
                                                             const uint32_t* __restrict__ sorted, uint32_t* __restrict__ nbr, uint32_t* __restrict__ nNbr,
                                                             uint32_t* __restrict__ status, const uint32_t* __restrict__ nPdev) {
    if (!*flag) return;
    if (nPdev) nP = *nPdev;
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= nP) return;
    const Grid g = *gp;
    const V3 xa = v3(pt[a].xc);
    const int ca = pt[a].cluster;
    const double r2 = nebrRange * nebrRange;
    int cx, cy, cz;
    grid_cell(g, pt[a].xc, &cx, &cy, &cz);
    uint32_t found[MAX_NBR];
    uint32_t cnt = 0;
    // cells may be wider than nebrRange but never narrower: the 27 cells around cover the range
    for (int dz = -1; dz <= 1; ++dz) {
        const int z = cz + dz;
        if (z < 0 || z >= (int)g.dim[2]) continue;
        for (int dy = -1; dy <= 1; ++dy) {
            const int y = cy + dy;
            if (y < 0 || y >= (int)g.dim[1]) continue;
            const int x0 = max(0, cx - 1), x1 = min((int)g.dim[0] - 1, cx + 1);
            // the cells of a row are consecutive: one range of the sorted array
            const uint32_t c0 = (uint32_t)x0 + g.dim[0] * ((uint32_t)y + g.dim[1] * (uint32_t)z), c1 = c0 + (uint32_t)(x1 - x0);
            for (uint32_t q = cellStart[c0]; q < cellStart[c1 + 1]; ++q) {
                const uint32_t j = sorted[q];
                if (pt[j].cluster == ca) continue;
                const V3 d = v3(pt[j].xc) - xa;
                if (norm2(d) < r2) {
                    if (cnt < (uint32_t)MAX_NBR) found[cnt] = j;
                    ++cnt;
                }
            }
        }
    }
    const uint32_t m = cnt < (uint32_t)MAX_NBR ? cnt : (uint32_t)MAX_NBR;
    for (uint32_t u = 1; u < m; ++u) {  // ascending, like the all-pairs pass (the order within a cell is whatever the atomics gave)
        const uint32_t v = found[u];
        uint32_t b = u;
        while (b > 0 && found[b - 1] > v) { found[b] = found[b - 1]; --b; }
        found[b] = v;
    }
    for (uint32_t u = 0; u < m; ++u) nbr[(size_t)a * MAX_NBR + u] = found[u];
    nNbr[a] = m;
    if (cnt > (uint32_t)MAX_NBR) atomicMax(status, cnt);
}

// ---------------------------------------------------------------------------------------------
// Periodic DEM boundaries (single spheres).  When the tables are rebuilt: DEM::pbcShift (DEM.cpp:1554-1584) brings back the
// elements that left the domain, DEM::createGhosts (DEM.cpp:1586-1660) gives every particle within nebrRange of a periodic
// plane a ghost on the other side and every particle with two ghosts one more in the corner.  Ghosts are further entries of
// the particle array (index nStd + g, g in the reference's creation order: per boundary, per element; then the corner
// ghosts in the order of their first parent), so the tables, the contacts (which act on the standard particle of a pair only,
// DEM.cpp:1849-1890) and the lists of the LB side see them as the reference's do.  One block.
// out: cnt[0] = particles + ghosts, cnt[1] |= 1 (a rebuild happened in this cycle: dem.newNeighborList, DEM.cpp:1414);
// comps[7 k ..] = element k's particle and its ghosts in creation order, nComp[k] their number.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t* __restrict__ a, uint32_t len, uint32_t* __restrict__ sm) {
    // exclusive scan of a[0 .. len) in place by the whole block (1024 threads); returns the total
    const uint32_t per = (len + blockDim.x - 1) / blockDim.x, b0 = threadIdx.x * per, b1 = min(len, b0 + per);
    uint32_t s = 0;
    for (uint32_t k = b0; k < b1; ++k) s += a[k];
    sm[threadIdx.x] = s;
    __syncthreads();
    for (uint32_t o = 1; o < blockDim.x; o <<= 1) {
        uint32_t v = 0;
        if (threadIdx.x >= o) v = sm[threadIdx.x - o];
        __syncthreads();
        sm[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = sm[threadIdx.x] - s;
    for (uint32_t k = b0; k < b1; ++k) { const uint32_t v = a[k]; a[k] = run; run += v; }
    const uint32_t total = sm[blockDim.x - 1];
    __syncthreads();
    return total;
}
__global__ void __launch_bounds__(1024) k_dem_pbc(Elmt* __restrict__ e, uint32_t n, const __grid_constant__ Params p, Part* __restrict__ pt,
                                                  const uint32_t* __restrict__ flag, int8_t* __restrict__ gflag, uint32_t* __restrict__ basePos,
                                                  uint32_t* __restrict__ cornPos, uint32_t* __restrict__ comps, uint32_t* __restrict__ nComp,
                                                  uint32_t* __restrict__ cnt) {
    if (!*flag) return;
    __shared__ uint32_t sm[1024];
    const int nb = p.nPbc;
    // pbcShift + the ghost test of every (boundary, element)
    for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) {
        V3 x0 = v3(e[k].x[0]);
        for (int b = 0; b < nb; ++b) {
            const V3 nrm = v3(p.pbcN[b]), v = v3(p.pbcV[b]);
            const double left = dot(nrm, x0 - v3(p.pbcP[b])), right = dot(-1.0 * nrm, x0 - (v3(p.pbcP[b]) + v));
            if (left < 0.0) x0 = x0 + v;
            if (right < 0.0) x0 = x0 + (-1.0 * v);
        }
        put(e[k].x[0], x0);
        put(pt[k].xc, x0);
        for (int b = 0; b < nb; ++b) {
            const V3 nrm = v3(p.pbcN[b]), v = v3(p.pbcV[b]);
            const double left = dot(nrm, x0 - v3(p.pbcP[b])), right = dot(-1.0 * nrm, x0 - (v3(p.pbcP[b]) + v));
            const int8_t f = left < p.nebrRange ? 1 : (right < p.nebrRange ? -1 : 0);
            gflag[(size_t)b * n + k] = f;
            basePos[(size_t)b * n + k] = f ? 1u : 0u;
        }
    }
    __syncthreads();
    // corner ghosts: one per pair (g1 < g2) of base ghosts of the same particle, listed under g1
    for (uint32_t q = threadIdx.x; q < (uint32_t)nb * n; q += blockDim.x) {
        const uint32_t b1 = q / n, k = q % n;
        uint32_t c = 0;
        if (gflag[q]) for (int b2 = (int)b1 + 1; b2 < nb; ++b2) c += gflag[(size_t)b2 * n + k] ? 1u : 0u;
        cornPos[q] = c;
    }
    __syncthreads();
    const uint32_t nBase = block_excl_scan(basePos, (uint32_t)nb * n, sm);
    const uint32_t nCorn = block_excl_scan(cornPos, (uint32_t)nb * n, sm);
    // the ghosts themselves, and every element's component list: its particle, its base ghosts by boundary, its corner ghosts
    for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) {
        const V3 x0 = v3(e[k].x[0]);
        uint32_t nc = 0;
        comps[(size_t)7 * k + nc++] = k;
        for (int b = 0; b < nb; ++b) {
            const int8_t f = gflag[(size_t)b * n + k];
            if (!f) continue;
            const uint32_t g = n + basePos[(size_t)b * n + k];
            const V3 sh = f > 0 ? v3(p.pbcV[b]) : -1.0 * v3(p.pbcV[b]);
            Part a = pt[k];
            put(a.shift, sh); put(a.xc, x0 + sh); a.nearWall = -1;
            pt[g] = a;
            comps[(size_t)7 * k + nc++] = g;
        }
        for (int b1 = 0; b1 < nb; ++b1) {
            const int8_t f1 = gflag[(size_t)b1 * n + k];
            if (!f1) continue;
            uint32_t r = 0;
            for (int b2 = b1 + 1; b2 < nb; ++b2) {
                const int8_t f2 = gflag[(size_t)b2 * n + k];
                if (!f2) continue;
                const uint32_t g = n + nBase + cornPos[(size_t)b1 * n + k] + r++;
                const V3 s1 = f1 > 0 ? v3(p.pbcV[b1]) : -1.0 * v3(p.pbcV[b1]), s2 = f2 > 0 ? v3(p.pbcV[b2]) : -1.0 * v3(p.pbcV[b2]);
                const V3 sh = s1 + s2;
                Part a = pt[k];
                put(a.shift, sh); put(a.xc, x0 + sh); a.nearWall = -1;
                pt[g] = a;
                comps[(size_t)7 * k + nc++] = g;
            }
        }
        nComp[k] = nc;
    }
    if (threadIdx.x == 0) { cnt[0] = n + nBase + nCorn; cnt[1] = 1u; }
}
// particle::ghostUpdate (elmt.cpp:292-299): a ghost follows its origin.  PRED: the predicted state (for the contacts), else
// the corrected one (for the lists of the LB side)
template <bool PRED>
__global__ void __launch_bounds__(128) k_dem_ghost_update(Part* __restrict__ pt, uint32_t nStd, const uint32_t* __restrict__ cnt) {
    const uint32_t a = nStd + blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= cnt[0]) return;
    const Part& o = pt[pt[a].cluster];  // single spheres: particle index = element index
    const V3 sh = v3(pt[a].shift);
    if (PRED) { put(pt[a].xp, v3(o.xp) + sh); put(pt[a].vp, v3(o.vp)); }
    else put(pt[a].xc, v3(o.xc) + sh);
}
// particle::updateCorrected (elmt.cpp:277-290): the spheres of element el at its corrected state
__device__ __forceinline__ void particles_corrected(const Params& p, const Elmt& el, Part* __restrict__ pt) {
    const V3 x0 = v3(el.x[0]);
    for (int i = 0; i < el.size; ++i) {
        Part& a = pt[el.pBegin + i];
        V3 xa = x0, rv = { 0.0, 0.0, 0.0 };
        if (el.size > 1) {
            xa = x0 + el.radius * project(v3(p.proto[el.size][i]), q4(el.q[0]));
            rv = xa - x0;
        }
        put(a.xc, xa); put(a.rvc, rv);
    }
}
__global__ void __launch_bounds__(128) k_dem_init_particles(const Elmt* __restrict__ e, uint32_t n, const __grid_constant__ Params p, Part* __restrict__ pt) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) particles_corrected(p, e[k], pt);
}

__global__ void __launch_bounds__(128) k_dem_predict(Elmt* __restrict__ e, uint32_t n, const __grid_constant__ Params p, Part* __restrict__ pt,
                                                     const Wall* __restrict__ walls, uint32_t nWalls, const uint32_t* __restrict__ flag) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    Elmt& el = e[k];
    if (*flag) {
        for (int i = 0; i < el.size; ++i) {
            Part& a = pt[el.pBegin + i];
            int nw = -1;
            const V3 xa = v3(a.xc);
            for (uint32_t w = 0; w < nWalls; ++w)
                if (dot(v3(walls[w].n), xa - v3(walls[w].p)) < p.nebrRange) { nw = (int)w; break; }
            a.nearWall = nw;
        }
    }
    const double* c = p.c;
    V3 x[6], w[6];
    for (int q = 0; q < 6; ++q) { x[q] = v3(el.x[q]); w[q] = v3(el.w[q]); }
    const V3 xp0 = x[0] + x[1] * c[0] + x[2] * c[1] + x[3] * c[2] + x[4] * c[3] + x[5] * c[4];
    const V3 xp1 = x[1] + x[2] * c[0] + x[3] * c[1] + x[4] * c[2] + x[5] * c[3];
    put(el.xp[0], xp0);
    put(el.xp[1], xp1);
    put(el.xp[2], x[2] + x[3] * c[0] + x[4] * c[1] + x[5] * c[2]);
    put(el.xp[3], x[3] + x[4] * c[0] + x[5] * c[1]);
    put(el.xp[4], x[4] + x[5] * c[0]);
    put(el.xp[5], x[5]);
    const V3 wp0 = w[0] + w[1] * c[0] + w[2] * c[1] + w[3] * c[2] + w[4] * c[3] + w[5] * c[4];
    put(el.wp[0], wp0);
    put(el.wp[1], w[1] + w[2] * c[0] + w[3] * c[1] + w[4] * c[2] + w[5] * c[3]);
    put(el.wp[2], w[2] + w[3] * c[0] + w[4] * c[1] + w[5] * c[2]);
    put(el.wp[3], w[3] + w[4] * c[0] + w[5] * c[1]);
    put(el.wp[4], w[4] + w[5] * c[0]);
    put(el.wp[5], w[5]);
    Q4 qp0 = { 1.0, 0.0, 0.0, 0.0 };
    if (el.size > 1) {
        Q4 q[6];
        for (int s = 0; s < 6; ++s) q[s] = q4(el.q[s]);
        qp0 = qnormalize(q[0] + q[1] * c[0] + q[2] * c[1] + q[3] * c[2] + q[4] * c[3] + q[5] * c[4]);
        putq(el.qp[0], qp0);
        putq(el.qp[1], q[1] + q[2] * c[0] + q[3] * c[1] + q[4] * c[2] + q[5] * c[3]);
        putq(el.qp[2], q[2] + q[3] * c[0] + q[4] * c[1] + q[5] * c[2]);
        putq(el.qp[3], q[3] + q[4] * c[0] + q[5] * c[1]);
        putq(el.qp[4], q[4] + q[5] * c[0]);
        putq(el.qp[5], q[5]);
    }
    // particle::updatePredicted (elmt.cpp:261-275); elmt::wpGlobal = wp0 (wSolver)
    for (int i = 0; i < el.size; ++i) {
        Part& a = pt[el.pBegin + i];
        V3 xa = xp0, rv = { 0.0, 0.0, 0.0 }, va = xp1;
        if (el.size > 1) {
            xa = xp0 + el.radius * project(v3(p.proto[el.size][i]), qp0);
            rv = xa - xp0;
            va = xp1 + cross(wp0, rv);
        }
        put(a.xp, xa); put(a.rvp, rv); put(a.vp, va);
    }
}

__device__ __forceinline__ double normal_contact(const Params& p, double overlap, double vreln, double effRad, double effMass) {
    if (p.contactModel == 1) {
        const double kn = p.knConst * sqrt(effRad) * sqrt(overlap);
        const double gamman = 2.0 * p.dampCoeff * sqrt(kn * effMass);
        return fmax(kn * overlap + (-gamman * vreln), 0.0);
    }
    const double gamman = 2.0 * p.dampCoeff * sqrt(p.linearStiff * effMass);
    return fmax(p.linearStiff * overlap + (-gamman * vreln), 0.0);
}
__device__ __forceinline__ double tangential_contact(const Params& p, double vrelt, double fn, double effRad, double effMass, double friction) {
    const double ks = p.contactModel == 1 ? p.ksConst * sqrt(effRad) * pow(fabs(fn), 1.0 / 3.0) : p.linearStiff;
    const double fsMax = friction * fn;
    const double gammas = 2.0 * p.viscTang * sqrt(effMass * ks);
    return fmin(gammas * vrelt, fsMax);
}

// hydro: per element {FHydro(3), MHydro(3), fluidVolume} in physical units, as the LB step's reduction left them
__global__ void __launch_bounds__(128) k_dem_forces_correct(Elmt* __restrict__ e, uint32_t n, const __grid_constant__ Params p, Part* __restrict__ pt,
                                                            const Wall* __restrict__ walls, const double* __restrict__ hydro,
                                                            const uint32_t* __restrict__ nbr, const uint32_t* __restrict__ nNbr) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    Elmt& el = e[k];
    const V3 xk = v3(el.xp[0]), vk = v3(el.xp[1]), wk = v3(el.wp[0]);
    const double rk = el.radius, mk = el.m;
    const bool cluster = el.size > 1;
    V3 FP = { 0, 0, 0 }, FW = { 0, 0, 0 }, MP = { 0, 0, 0 }, MW = { 0, 0, 0 };
    for (int i = 0; i < el.size; ++i) {
        const uint32_t a = (uint32_t)el.pBegin + (uint32_t)i;
        const Part pa = pt[a];
        const uint32_t nn = nNbr[a];
        for (uint32_t q = 0; q < nn; ++q) {
            const uint32_t b = nbr[(size_t)a * MAX_NBR + q];
            const Part& pb = pt[b];
            const Elmt& o = e[pb.cluster];
            const bool iAmI = a < b;  // the pair's force is evaluated in one role assignment (lower particle index = I) on both sides
            const V3 xI = iAmI ? v3(pa.xp) : v3(pb.xp), xJ = iAmI ? v3(pb.xp) : v3(pa.xp);
            const V3 d = xJ - xI;  // partJ->x0 - partI->x0
            const double rI = iAmI ? rk : o.radius, rJ = iAmI ? o.radius : rk;
            const double sig = rI + rJ;
            if (!(norm2(d) < sig * sig)) continue;
            const double mI = iAmI ? mk : o.m, mJ = iAmI ? o.m : mk;
            const V3 vI = iAmI ? v3(pa.vp) : v3(pb.vp), vJ = iAmI ? v3(pb.vp) : v3(pa.vp);
            const V3 wI = iAmI ? wk : v3(o.wp[0]), wJ = iAmI ? v3(o.wp[0]) : wk;
            const double dist = sqrt(norm2(d));
            const double overlap = rI + rJ - dist;
            const V3 relVel = vJ - vI;
            const V3 en = d / dist;
            const double vn = dot(relVel, en);
            const V3 normalRelVel = en * vn;
            const double effMass = mI * mJ / (mI + mJ);
            const double effRad = rI * rJ / (rI + rJ);
            const double fn = normal_contact(p, overlap, vn, effRad, effMass);
            const V3 nf = en * fn;
            const V3 vecRadI = rI * en, vecRadJ = -rJ * en;
            // lever arm of the contact point about the element's centre: vecRad + particle::radiusVec (DEM.cpp:1841-1846)
            const V3 cdMine = iAmI ? vecRadI + v3(pa.rvp) : vecRadJ + v3(pa.rvp);
            if (iAmI) { FP = FP - nf; if (cluster) MP = MP - cross(cdMine, nf); }
            else { FP = FP + nf; if (cluster) MP = MP + cross(cdMine, nf); }
            const V3 relC = relVel - cross(wI, vecRadI) + cross(wJ, vecRadJ);
            const V3 tang = relC - normalRelVel;
            const double nt = sqrt(norm2(tang));
            if (nt != 0.0) {
                const double ft = tangential_contact(p, nt, fn, effRad, effMass, p.frictionCoefPart);
                const V3 et = tang / nt;
                const V3 tf = ft * et;
                if (iAmI) { MP = MP + cross(cdMine, tf); FP = FP + tf; }
                else { MP = MP - cross(cdMine, tf); FP = FP - tf; }
            }
        }
        if (pa.nearWall >= 0) {
            const Wall wl = walls[pa.nearWall];
            const V3 en = v3(wl.n);
            const V3 xa = v3(pa.xp), va = v3(pa.vp);
            const double dist = dot(en, xa - v3(wl.p));
            const double overlap = rk - dist;
            if (overlap > 0.0) {
                V3 cpv = { 0.0, 0.0, 0.0 };
                if (wl.moving) {
                    const V3 dc = xa - v3(wl.rotCenter);
                    cpv = v3(wl.vel) + cross(v3(wl.omega), dc - dot(dc, en) * en);
                }
                const V3 relVel = va - cpv;
                const double vn = dot(relVel, en);
                const V3 normalRelVel = en * vn;
                const double fn = normal_contact(p, 2.0 * overlap, vn, rk, mk);
                const V3 nf = en * fn;
                const V3 vecRadJ = -rk * en;
                V3 cdJ = vecRadJ;
                if (cluster) cdJ = cdJ + (xa - xk);  // DEM.cpp:1937-1941
                FW = FW + nf;
                if (cluster) MW = MW + cross(cdJ, nf);
                const V3 relC = relVel + cross(wk, vecRadJ);
                const V3 tang = relC - normalRelVel;
                const double nt = sqrt(norm2(tang));
                if (nt != 0.0) {
                    const double ft = tangential_contact(p, nt, fn, rk, mk, p.frictionCoefWall);
                    const V3 et = tang / sqrt(norm2(tang));
                    const V3 tf = ft * et;
                    MW = MW - cross(cdJ, tf);
                    FW = FW - tf;
                }
            }
        }
    }
    // Newton's equations (DEM.cpp:1150-1181)
    const V3 FH = hydro ? v3(hydro + (size_t)7 * k) : V3{ 0, 0, 0 }, MH = hydro ? v3(hydro + (size_t)7 * k + 3) : V3{ 0, 0, 0 };
    const V3 FVisc = -6.0 * M_PI * p.numVisc * rk * vk;
    const V3 MVisc = -8.0 * M_PI * p.numVisc * rk * rk * rk * wk;
    const V3 x2 = (FVisc + FH + FP + FW) / mk + v3(p.demF);
    const V3 mom = MVisc + MH + MP + MW;
    const double* I = el.I;
    V3 w1;
    Q4 q2 = { 0.0, 0.0, 0.0, 0.0 };
    if (!cluster) {
        // a sphere's body frame stays the global one (its quaternion is the identity: the projections are exact identities)
        w1 = { (mom.x + (I[1] - I[2]) * wk.y * wk.z) / I[0], (mom.y + (I[2] - I[0]) * wk.z * wk.x) / I[1],
               (mom.z + (I[0] - I[1]) * wk.x * wk.y) / I[2] };
    } else {
        const Q4 qp0 = q4(el.qp[0]), qp1 = q4(el.qp[1]);
        const V3 momBf = project(mom, qadj(qp0));
        const V3 wl = project(wk, qadj(qp0));  // elmt::wpLocal (elmt.cpp:164-167)
        const V3 waBf = { (momBf.x + (I[1] - I[2]) * wl.y * wl.z) / I[0], (momBf.y + (I[2] - I[0]) * wl.z * wl.x) / I[1],
                          (momBf.z + (I[0] - I[1]) * wl.x * wl.y) / I[2] };
        w1 = project(waBf, qp0);
        const Q4 waQuat = { -2.0 * (qp1.q0 * qp1.q0 + qp1.q1 * qp1.q1 + qp1.q2 * qp1.q2 + qp1.q3 * qp1.q3), waBf.x, waBf.y, waBf.z };
        q2 = 0.5 * qmul(qp0, waQuat);
    }
    // elmt::correct
    const double* c2 = p.coeff2; const double* c1 = p.coeff1;
    V3 xp[6], wp[6];
    for (int q = 0; q < 6; ++q) { xp[q] = v3(el.xp[q]); wp[q] = v3(el.wp[q]); }
    const V3 x2c = x2 - xp[2];
    V3 x[6];
    x[0] = xp[0] + x2c * c2[0]; x[1] = xp[1] + x2c * c2[1]; x[2] = x2;
    x[3] = xp[3] + x2c * c2[3]; x[4] = xp[4] + x2c * c2[4]; x[5] = xp[5] + x2c * c2[5];
    const V3 w1c = w1 - wp[1];
    V3 w[6];
    w[0] = wp[0] + w1c * c1[0]; w[1] = w1;
    w[2] = wp[2] + w1c * c1[2]; w[3] = wp[3] + w1c * c1[3]; w[4] = wp[4] + w1c * c1[4]; w[5] = wp[5] + w1c * c1[5];
    // every element reads the others' PREDICTED state (xp, wp, the particles' xp / rvp / vp) above and writes only its own
    // corrected state (x, w, q, the particles' xc / rvc); xp := x ... (the tail of elmt::correct) happens at the top of the next
    // predict, which overwrites them anyway
    for (int q = 0; q < 6; ++q) { put(el.x[q], x[q]); put(el.w[q], w[q]); }
    if (cluster) {
        Q4 qp[6];
        for (int s = 0; s < 6; ++s) qp[s] = q4(el.qp[s]);
        const Q4 q2c = q2 - qp[2];
        putq(el.q[0], qnormalize(qp[0] + q2c * c2[0]));
        putq(el.q[1], qp[1] + q2c * c2[1]);
        putq(el.q[2], q2);
        putq(el.q[3], qp[3] + q2c * c2[3]);
        putq(el.q[4], qp[4] + q2c * c2[4]);
        putq(el.q[5], qp[5] + q2c * c2[5]);
    }
    put(el.fc[0], FP); put(el.fc[1], FW); put(el.fc[2], MP); put(el.fc[3], MW);
    particles_corrected(p, el, pt);
}

// the lists the LB side reads (layout of LbGpuParticle / LbGpuElement, physical units): particle::updateCorrected, elmt::x1,
// elmt::wGlobal = w0 (wSolver, elmt.cpp:232-235), elmt::components = the element's particles
struct OutParticle { double x0[3], r, radiusVec[3]; uint32_t clusterIndex, particleIndex; };
struct OutElement { double x1[3], wGlobal[3]; uint32_t compBegin, compEnd; };
__global__ void __launch_bounds__(128) k_dem_export(const Elmt* __restrict__ e, uint32_t n, const Part* __restrict__ pt, OutParticle* __restrict__ parts,
                                                    OutElement* __restrict__ elmts, uint32_t* __restrict__ comps) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    OutElement oe;
    for (int c = 0; c < 3; ++c) { oe.x1[c] = e[k].x[1][c]; oe.wGlobal[c] = e[k].w[0][c]; }
    oe.compBegin = (uint32_t)e[k].pBegin; oe.compEnd = (uint32_t)(e[k].pBegin + e[k].size);
    elmts[k] = oe;
    for (int i = 0; i < e[k].size; ++i) {
        const uint32_t a = (uint32_t)e[k].pBegin + (uint32_t)i;
        OutParticle op;
        for (int c = 0; c < 3; ++c) { op.x0[c] = pt[a].xc[c]; op.radiusVec[c] = pt[a].rvc[c]; }
        op.r = e[k].radius; op.clusterIndex = k; op.particleIndex = a;
        parts[a] = op; comps[a] = a;
    }
}

// the same with periodic boundaries: the ghosts follow the standard particles in the list, an element's components are the
// 7-slot rows k_dem_pbc filled
__global__ void __launch_bounds__(128) k_dem_export_pbc(const Elmt* __restrict__ e, uint32_t n, const Part* __restrict__ pt, const uint32_t* __restrict__ cnt,
                                                        const uint32_t* __restrict__ nComp, OutParticle* __restrict__ parts, OutElement* __restrict__ elmts) {
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a < cnt[0]) {
        const uint32_t k = (uint32_t)pt[a].cluster;
        OutParticle op;
        for (int c = 0; c < 3; ++c) { op.x0[c] = pt[a].xc[c]; op.radiusVec[c] = 0.0; }
        op.r = e[k].radius; op.clusterIndex = k; op.particleIndex = a;
        parts[a] = op;
    }
    if (a < n) {
        OutElement oe;
        for (int c = 0; c < 3; ++c) { oe.x1[c] = e[a].x[1][c]; oe.wGlobal[c] = e[a].w[0][c]; }
        oe.compBegin = 7u * a; oe.compEnd = 7u * a + nComp[a];
        elmts[a] = oe;
    }
}

}  // namespace lbdem
