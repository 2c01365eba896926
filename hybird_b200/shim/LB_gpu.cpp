// LB_gpu.cpp -- drop-in replacement of the reference's LB time-step methods on top of liblbgpu.so.
//
// The reference (gnomeCreative/hybird) is built UNMODIFIED: hybird.cpp, DEM.cpp, IO.cpp, elmt.cpp, utils.cpp,
// vector.cpp, node.cpp and LB.cpp are compiled where they lie.  In LB.o the five methods that form the boundary of
// the hot path are renamed with objcopy (see Makefile) to plain symbols LB_ref_*; this translation unit defines the
// methods under their original names, so every call site of the reference (hybird.cpp:47-59,193,310) binds here:
//
//   LB::latticeBoltzmannGet            reference parse, plus the export cadence IO will use (screenExpTime, fluidExpTime)
//   LB::latticeBolzmannInit            reference host init (LB.cpp:190-219), then lbGpuInit uploads the state
//   LB::latticeBoltzmannFreeSurfaceStep  records the request                       (LB.cpp:235-245)
//   LB::latticeBoltzmannCouplingStep   packs particles/elements, resets the flag   (LB.cpp:247-280)
//   LB::latticeBolzmannStep            lbGpuStep + lbGpuParticleForces -> elmts[].FHydro/MHydro/fluidVolume,
//                                      walls[].FHydro
//
// The output path.  IO reads the fluid through lb.types / lb.nodes; the four members that walk
// the whole lattice are made weak in IO.o (objcopy --weaken-symbol, see Makefile) and defined here on top of the device:
//
//   IO::exportMaxSpeedFluid, IO::exportTotalMass, IO::exportPlasticity   lbGpuFluidSummary (one device reduction; same
//                                      text on cout / export.dat / maxFluidSpeed.dat / plasticity.dat)   IO.cpp:835-895
//   IO::exportParaviewFluidOld         lbGpuWriteVti: the same .vti (names, types, order, values) with raw appended data
//                                      instead of formatted text                                          IO.cpp:698-831
//
// so an export costs a reduction (screen) or one selective field fetch per array (fluid file) instead of a fetch of
// every field plus a rebuild of the host node lists.  Only the problem-specific screen exports that read the fluid
// themselves (DRUM: exportDrum, SHEARCELL: exportShearCell) still get refreshed host mirrors.
//
// Built twice (Makefile): hybird_gpu -- the product: no reference LB step is linked (unreferenced reference functions are
// dropped by --gc-sections); hybird_gpu_verify -- compiled with LBGPU_SHIM_VERIFY: LBGPU_VERIFY=1 additionally runs the
// reference's own step on the host state every cycle and compares the fields (integration check; the forces handed
// to DEM are always the device's).
// Built with -fno-access-control like the oracle harness: LB's members are "public: //private" (LB.h:36) but
// nodeType's are private.
#include "../../include/lbgpu.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "LB.h"
#include "IO.h"
#include "DEM.h"

extern "C" {
void LB_ref_latticeBoltzmannGet(LB*, GetPot&, GetPot&);
void LB_ref_latticeBolzmannInit(LB*, cylinderList&, wallList&, particleList&, objectList&);
void DEM_ref_discreteElementStep(DEM*, IO&);
#ifdef LBGPU_SHIM_VERIFY
void LB_ref_latticeBolzmannStep(LB*, elmtList&, particleList&, wallList&);
void LB_ref_latticeBoltzmannCouplingStep(LB*, bool&, elmtList&, particleList&);
void LB_ref_latticeBoltzmannFreeSurfaceStep(LB*);
#endif
}

namespace {

struct GpuState {
    LbGpuHandle* h = nullptr;
    bool fsRequested = false, couplePending = false, rescan = false, verify = false;
    std::vector<LbGpuParticle> parts;
    std::vector<LbGpuElement> elmts;
    std::vector<uint32_t> comps;
    double screenExpTime = 0.0, fluidExpTime = 0.0;
    unsigned int lastScreenExp = 0, lastFluidExp = 0;
    unsigned long long steps = 0, fetches = 0, summaries = 0, vtis = 0;
    double worst = 0.0;
    // LBGPU_DEM=1: DEM::discreteElementStep runs on the device (lbGpuDem*), for single-sphere elements between plane walls
    bool demChecked = false, demOnDevice = false, demPending = false;
    DEM* dem = nullptr;
    unsigned long long demSteps = 0;
    // fetch buffers
    std::vector<uint8_t> tf;
    std::vector<uint32_t> solid;
    std::vector<double> n, u, mass, visc, shear;
};

std::map<LB*, GpuState>& states() {
    static std::map<LB*, GpuState> m;
    return m;
}

void die(const char* what) {
    // the reference's convention for fatal conditions on this path: message on cout, exit(1) (macros.h:8)
    cout << "lbgpu shim: ERROR in " << what << ": " << lbGpuLastError() << endl;
    exit(1);
}

void pack(GpuState& st, elmtList& elmts, particleList& particles) {
    st.parts.resize(particles.size());
    for (size_t k = 0; k < particles.size(); ++k) {
        const particle& p = particles[k];
        LbGpuParticle& o = st.parts[k];
        o.x0[0] = p.x0.x; o.x0[1] = p.x0.y; o.x0[2] = p.x0.z;
        o.r = p.r;
        o.radiusVec[0] = p.radiusVec.x; o.radiusVec[1] = p.radiusVec.y; o.radiusVec[2] = p.radiusVec.z;
        o.clusterIndex = p.clusterIndex; o.particleIndex = p.particleIndex;
    }
    st.elmts.resize(elmts.size());
    st.comps.clear();
    for (size_t e = 0; e < elmts.size(); ++e) {
        const elmt& el = elmts[e];
        LbGpuElement& o = st.elmts[e];
        o.x1[0] = el.x1.x; o.x1[1] = el.x1.y; o.x1[2] = el.x1.z;
        o.wGlobal[0] = el.wGlobal.x; o.wGlobal[1] = el.wGlobal.y; o.wGlobal[2] = el.wGlobal.z;
        o.compBegin = (uint32_t)st.comps.size();
        for (size_t c = 0; c < el.components.size(); ++c) st.comps.push_back((uint32_t)el.components[c]);
        o.compEnd = (uint32_t)st.comps.size();
    }
}

// device state -> the host mirrors IO reads (IO.cpp:698-831, 835-895): types, node existence, n, u, visc, mass,
// shearRate, and the node lists (activeNodes is what exportMaxSpeedFluid / exportTotalMass walk)
void refresh_mirrors(LB* lb, GpuState& st) {
    const size_t N = lb->totNodes;
    st.tf.resize(N); st.solid.resize(N); st.n.resize(N); st.u.resize(3 * N); st.mass.resize(N); st.visc.resize(N); st.shear.resize(N);
    if (lbGpuFetchFields(st.h, st.tf.data(), st.solid.data(), st.n.data(), st.u.data(), st.mass.data(), st.visc.data(),
                         st.shear.data(), nullptr, nullptr))
        die("lbGpuFetchFields");
    ++st.fetches;
    lb->fluidNodes.clear(); lb->interfaceNodes.clear(); lb->particleNodes.clear(); lb->activeNodes.clear();
    for (size_t i = 0; i < N; ++i) {
        const uint8_t b = st.tf[i];
        unsigned int t = b & LBGPU_TYPE_MASK;
        nodeType& ty = lb->types[i];
        ty.setType(t);
        if (b & LBGPU_P_BIT) ty.setInsideParticle(); else ty.setOutsideParticle();
        unsigned int si = st.solid[i];
        ty.setSolidIndex(si);
        const bool has = (b & LBGPU_NODE_BIT) != 0;
        if (has && lb->nodes[i] == 0) lb->nodes[i] = new node;
        if (!has && lb->nodes[i] != 0) { delete lb->nodes[i]; lb->nodes[i] = 0; }
        if (has) {
            node* nd = lb->nodes[i];
            nd->n = st.n[i];
            nd->u = tVect(st.u[3 * i], st.u[3 * i + 1], st.u[3 * i + 2]);
            nd->mass = st.mass[i];
            nd->visc = st.visc[i];
            nd->shearRate = st.shear[i];
        }
        if (t == 0) { lb->fluidNodes.push_back((unsigned int)i); lb->activeNodes.push_back((unsigned int)i); }
        else if (t == 3) { lb->interfaceNodes.push_back((unsigned int)i); lb->activeNodes.push_back((unsigned int)i); }
        if ((b & LBGPU_P_BIT) && (t == 0 || t == 3)) lb->particleNodes.push_back((unsigned int)i);
    }
}

// Will IO::outputStep of the NEXT cycle read the fluid state?  It tests uint(realTime/expTime)+1 > last with
// realTime = dem.demTime (IO.cpp:134,148-149,254-255), which by then equals lb.time * unit.Time up to the rounding of
// DEM's accumulated sub-steps; both sides of that rounding are tested, so an export is never missed (an extra
// refresh costs time only).
bool export_due(LB* lb, GpuState& st) {
    const double t = (double)lb->time * lb->unit.Time;
    const double lo = t * (1.0 - 1e-7), hi = t * (1.0 + 1e-7);
    bool due = false;
    if (st.screenExpTime > 0) {
        const unsigned int cLo = (unsigned int)(lo / st.screenExpTime) + 1, cHi = (unsigned int)(hi / st.screenExpTime) + 1;
        if (cHi > st.lastScreenExp) { due = true; st.lastScreenExp = cLo > st.lastScreenExp ? cLo : st.lastScreenExp; }
    }
    if (st.fluidExpTime > 0) {
        const unsigned int cLo = (unsigned int)(lo / st.fluidExpTime) + 1, cHi = (unsigned int)(hi / st.fluidExpTime) + 1;
        if (cHi > st.lastFluidExp) { due = true; st.lastFluidExp = cLo > st.lastFluidExp ? cLo : st.lastFluidExp; }
    }
    return due;
}

#ifdef LBGPU_SHIM_VERIFY
double rel_diff(double a, double b) {
    const double s = fmax(fabs(a), fabs(b));
    return s > 0 ? fabs(a - b) / s : 0.0;
}

// LBGPU_VERIFY: host state (just advanced by the reference's own step) against the device state
void verify_against_host(LB* lb, GpuState& st, const std::vector<tVect>& refF, elmtList& elmts) {
    const size_t N = lb->totNodes;
    st.tf.resize(N); st.solid.resize(N); st.n.resize(N); st.u.resize(3 * N); st.mass.resize(N); st.visc.resize(N); st.shear.resize(N);
    std::vector<double> f(19 * N);
    if (lbGpuFetchFields(st.h, st.tf.data(), st.solid.data(), st.n.data(), st.u.data(), st.mass.data(), st.visc.data(),
                         st.shear.data(), nullptr, f.data()))
        die("lbGpuFetchFields");
    size_t typeDiff = 0;
    double worst = 0.0;
    for (size_t i = 0; i < N; ++i) {
        const unsigned int t = st.tf[i] & LBGPU_TYPE_MASK;
        if (t != lb->types[i].getType() || ((st.tf[i] & LBGPU_P_BIT) != 0) != lb->types[i].isInsideParticle()) { ++typeDiff; continue; }
        if (!(t == 0 || t == 3)) continue;
        const node* nd = lb->nodes[i];
        worst = fmax(worst, rel_diff(st.n[i], nd->n));
        worst = fmax(worst, rel_diff(st.mass[i], nd->mass));
        worst = fmax(worst, rel_diff(st.visc[i], nd->visc));
        for (int j = 0; j < 19; ++j) worst = fmax(worst, rel_diff(f[19 * i + j], nd->fs[j]));  // post-collision populations
    }
    double fWorst = 0.0, fScale = 0.0;
    for (size_t e = 0; e < elmts.size(); ++e) fScale = fmax(fScale, refF[e].norm());
    for (size_t e = 0; e < elmts.size(); ++e) fWorst = fmax(fWorst, (refF[e] - elmts[e].FHydro).norm() / (fScale > 0 ? fScale : 1.0));
    st.worst = fmax(st.worst, worst);
    cout << "lbgpu verify: step " << st.steps << " type mismatches " << typeDiff << ", max rel field diff " << worst
         << ", max rel FHydro diff " << fWorst << endl;
    if (typeDiff != 0 || worst > 1e-9 || fWorst > 1e-9) { cout << "lbgpu verify: FAILED" << endl; exit(2); }
}
#endif

// the screen exports of these problems read lb.nodes themselves (IO::exportDrum, IO::exportShearCell)
bool needs_host_mirrors() { return problemName == DRUM || problemName == SHEARCELL; }

GpuState& state_of(const LB& lb) { return states()[const_cast<LB*>(&lb)]; }

}  // namespace

void LB::latticeBoltzmannGet(GetPot& lbmCfgFile, GetPot& command_line) {
    LB_ref_latticeBoltzmannGet(this, lbmCfgFile, command_line);
    GpuState& st = states()[this];
    // the cadence IO::outputStep will follow (parsed there with the same macro, hybird.cpp:149-176)
    st.screenExpTime = lbmCfgFile("screenExpTime", 0.0);
    if (command_line.search("-screenExpTime")) st.screenExpTime = command_line.next(st.screenExpTime);
    st.fluidExpTime = lbmCfgFile("fluidExpTime", 0.0);
    if (command_line.search("-fluidExpTime")) st.fluidExpTime = command_line.next(st.fluidExpTime);
#ifdef LBGPU_SHIM_VERIFY
    const char* v = getenv("LBGPU_VERIFY");
    st.verify = v && v[0] == '1';
#else
    if (const char* v = getenv("LBGPU_VERIFY"))
        if (v[0] == '1') { cout << "lbgpu shim: LBGPU_VERIFY needs the hybird_gpu_verify binary (this one links no reference LB step)" << endl; exit(1); }
#endif
}

void LB::latticeBolzmannInit(cylinderList& cylinders, wallList& walls, particleList& particles, objectList& objects) {
    LB_ref_latticeBolzmannInit(this, cylinders, walls, particles, objects);
    GpuState& st = states()[this];
    const size_t N = totNodes;
    LbGpuParams prm;
    memset(&prm, 0, sizeof prm);
    for (int k = 0; k < 3; ++k) prm.size[k] = (int32_t)lbSize[k];
    for (int k = 0; k < 6; ++k) prm.boundary[k] = (int32_t)boundary[k];  // capacity 6 although size() is 3 (LB.cpp:175-181)
    prm.lbF[0] = lbF.x; prm.lbF[1] = lbF.y; prm.lbF[2] = lbF.z;
    prm.initDynVisc = initDynVisc; prm.plasticVisc = plasticVisc; prm.yieldStress = yieldStress;
    prm.turbConst = turbConst; prm.slipCoefficient = slipCoefficient;
    prm.freeSurface = freeSurface; prm.forceField = forceField; prm.nonNewtonian = nonNewtonian; prm.turbulence = turbulenceOn;
    prm.unitLength = unit.Length; prm.unitTime = unit.Time; prm.unitDensity = unit.Density;
    prm.nWalls = (int32_t)walls.size();
    prm.device = -1;
    prm.slabAxis = 2; prm.nSlabs = 1; prm.slabIndex = 0; prm.nLocalSlabs = 1;
    std::vector<uint8_t> tf(N);
    std::vector<uint32_t> solid(N);
    std::vector<double> f(19 * N, 0.0), n(N, 0.0), u(3 * N, 0.0), mass(N, 0.0), visc(N, 0.0);
    for (size_t i = 0; i < N; ++i) {
        const nodeType& ty = types[i];
        tf[i] = (uint8_t)(ty.getType() | (ty.isInsideParticle() ? LBGPU_P_BIT : 0) | (nodes[i] != 0 ? LBGPU_NODE_BIT : 0));
        solid[i] = ty.getSolidIndex();
        if (const node* nd = nodes[i]) {
            n[i] = nd->n; mass[i] = nd->mass; visc[i] = nd->visc;
            u[3 * i] = nd->u.x; u[3 * i + 1] = nd->u.y; u[3 * i + 2] = nd->u.z;
            for (int j = 0; j < 19; ++j) f[19 * i + j] = nd->f[j];
        }
    }
    if (lbGpuInit(&prm, tf.data(), solid.data(), f.data(), n.data(), u.data(), mass.data(), visc.data(), &st.h)) die("lbGpuInit");
    // curved walls (LB::curves, LB.cpp:363-367, 581-603) and the DRUM mass target (LB.cpp:205-209, 239-244)
    {
        std::vector<uint32_t> cells;
        std::vector<double> delta;
        for (size_t i = 0; i < N; ++i) {
            if (curves[i] == 0) continue;
            cells.push_back((uint32_t)i);
            delta.push_back(0.0);  // delta[0] is neither initialised nor read by the reference
            for (int j = 1; j < 19; ++j) delta.push_back(curves[i]->delta[j]);
        }
        if (!cells.empty() && lbGpuSetCurves(st.h, (uint32_t)cells.size(), cells.data(), delta.data())) die("lbGpuSetCurves");
        if (problemName == DRUM && freeSurface && lbGpuSetMassTarget(st.h, totalMass)) die("lbGpuSetMassTarget");
    }
    cout << "lbgpu shim: " << N << " cells uploaded to the GPU" << (st.verify ? " (verify mode: the reference steps alongside)" : "") << endl;
}

void LB::latticeBoltzmannFreeSurfaceStep() {
    GpuState& st = states()[this];
    st.fsRequested = true;
#ifdef LBGPU_SHIM_VERIFY
    if (st.verify) LB_ref_latticeBoltzmannFreeSurfaceStep(this);
#endif
}

void LB::latticeBoltzmannCouplingStep(bool& newNeighborList, elmtList& elmts, particleList& particles) {
    GpuState& st = states()[this];
#ifdef LBGPU_SHIM_VERIFY
    if (st.verify) { bool flag = newNeighborList; LB_ref_latticeBoltzmannCouplingStep(this, flag, elmts, particles); }
#endif
    if (st.couplePending) {
        // the previous cycle coupled without stepping the fluid (demTime <= demInitialRepeat, hybird.cpp:60-64)
        if (lbGpuCouple(st.h, st.rescan, st.parts.data(), (uint32_t)st.parts.size(), st.elmts.data(), (uint32_t)st.elmts.size(),
                        st.comps.data(), (uint32_t)st.comps.size()))
            die("lbGpuCouple");
    }
    if (!st.demPending) pack(st, elmts, particles);  // with the DEM on the device the resident lists are the current ones
    st.rescan = newNeighborList;
    st.couplePending = true;
    newNeighborList = false;  // LB.cpp:258
}

// ---------------------------------------------------------------------------------------------
// DEM::discreteElementStep (DEM.cpp:331-376; renamed to DEM_ref_discreteElementStep in DEM.o, see Makefile).  With
// LBGPU_DEM=1 and a DEM the device covers -- single-sphere elements, plane walls, no periodic DEM boundaries, cylinders or
// objects, the fluid stepped from the first cycle on -- the sub-steps run on the device inside the same lbGpuRunDem call
// that runs the LB side of this cycle; the host DEM object only keeps its clock and receives the elements' state back
// for IO.  Anything else falls through to the reference's own step.
// ---------------------------------------------------------------------------------------------
void DEM::discreteElementStep(IO& io) {
    GpuState* st = states().empty() ? nullptr : &states().begin()->second;
    if (st && st->h && !st->demChecked) {
        st->demChecked = true;
        const char* e = getenv("LBGPU_DEM");
        bool ok = e && e[0] == '1' && io.lbmSolve && demInitialRepeat == 0.0 && io.saveCount == 0 && !st->verify && !elmts.empty() &&
                  pbcs.empty() && cylinders.empty() && objects.empty() && ghosts.empty();
        for (size_t k = 0; ok && k < elmts.size(); ++k) ok = elmts[k].size == 1;
        if (ok) {
            LbGpuDemParams P;
            memset(&P, 0, sizeof P);
            P.contactModel = sphereMat.contactModel == HERTZIAN ? 1 : 0;
            P.multiStep = (int32_t)multiStep;
            P.knConst = sphereMat.knConst; P.ksConst = sphereMat.ksConst; P.dampCoeff = sphereMat.dampCoeff; P.viscTang = sphereMat.viscTang;
            P.linearStiff = sphereMat.linearStiff; P.frictionCoefPart = sphereMat.frictionCoefPart; P.frictionCoefWall = sphereMat.frictionCoefWall;
            P.numVisc = numVisc; P.demF[0] = demF.x; P.demF[1] = demF.y; P.demF[2] = demF.z;
            P.deltat = deltat; P.nebrRange = nebrRange; P.maxDisp = maxDisp;
            std::vector<LbGpuDemElement> E(elmts.size());
            for (size_t k = 0; k < elmts.size(); ++k) {
                const elmt& el = elmts[k];
                LbGpuDemElement& o = E[k];
                o.x0[0] = el.x0.x; o.x0[1] = el.x0.y; o.x0[2] = el.x0.z;
                o.x1[0] = el.x1.x; o.x1[1] = el.x1.y; o.x1[2] = el.x1.z;
                o.w0[0] = el.w0.x; o.w0[1] = el.w0.y; o.w0[2] = el.w0.z;
                o.radius = el.radius; o.m = el.m; o.I[0] = el.I.x; o.I[1] = el.I.y; o.I[2] = el.I.z; o.size = 1; o.pad = 0;
            }
            std::vector<LbGpuDemWall> W(walls.size());
            for (size_t k = 0; k < walls.size(); ++k) {
                const wall& w = walls[k];
                LbGpuDemWall& o = W[k];
                memset(&o, 0, sizeof o);
                o.n[0] = w.n.x; o.n[1] = w.n.y; o.n[2] = w.n.z; o.p[0] = w.p.x; o.p[1] = w.p.y; o.p[2] = w.p.z;
                o.vel[0] = w.vel.x; o.vel[1] = w.vel.y; o.vel[2] = w.vel.z; o.omega[0] = w.omega.x; o.omega[1] = w.omega.y; o.omega[2] = w.omega.z;
                o.rotCenter[0] = w.rotCenter.x; o.rotCenter[1] = w.rotCenter.y; o.rotCenter[2] = w.rotCenter.z;
                o.moving = w.moving ? 1 : 0;
            }
            if (lbGpuDemInit(st->h, &P, E.data(), (uint32_t)E.size(), W.empty() ? nullptr : W.data(), (uint32_t)W.size())) die("lbGpuDemInit");
            st->demOnDevice = true;
            st->dem = this;
            cout << "lbgpu shim: DEM sub-steps on the GPU (" << elmts.size() << " spheres, " << walls.size() << " walls, " << multiStep << " per LB step)" << endl;
        } else if (e && e[0] == '1') {
            cout << "lbgpu shim: LBGPU_DEM=1, but this DEM is outside what the device covers (clusters, periodic boundaries, cylinders, "
                    "objects, demInitialRepeat, saveCount): the host DEM runs" << endl;
        }
    }
    if (!st || !st->demOnDevice) { DEM_ref_discreteElementStep(this, io); return; }
    for (unsigned int it = 0; it < multiStep; ++it) { demTimeStep++; demTime += deltat; }  // the clock, accumulated as the reference does
    st->demPending = true;  // the sub-steps themselves run inside this cycle's lbGpuRunDem (LB::latticeBolzmannStep)
}

void LB::latticeBolzmannStep(elmtList& elmts, particleList& particles, wallList& walls) {
    GpuState& st = states()[this];
#ifdef LBGPU_SHIM_VERIFY
    std::vector<tVect> refF;
    if (st.verify) {
        elmtList refElmts(elmts);  // elmt has const members: copy-construct, never assign
        wallList refWalls(walls);
        LB_ref_latticeBolzmannStep(this, refElmts, particles, refWalls);
        for (size_t e = 0; e < refElmts.size(); ++e) refF.push_back(refElmts[e].FHydro);
    }
#endif
    if (st.demPending) {
        // goCycle's order on the device: DEM step, free-surface step, coupling step, LB step
        if (lbGpuRunDem(st.h, st.fsRequested ? 1 : 0, 1)) die("lbGpuRunDem");
        ++st.demSteps;
        const size_t nE = elmts.size();
        std::vector<double> x0(3 * nE), x1(3 * nE), w0(3 * nE), FP(3 * nE), FW(3 * nE), MP(3 * nE), MW(3 * nE);
        double info[3];
        if (lbGpuDemState(st.h, x0.data(), x1.data(), w0.data(), info)) die("lbGpuDemState");
        if (lbGpuDemContacts(st.h, FP.data(), FW.data(), MP.data(), MW.data())) die("lbGpuDemContacts");
        // what IO and the driver read from the DEM object between steps: the corrected state of every element, its particle
        for (size_t e = 0; e < nE; ++e) {
            elmt& el = elmts[e];
            el.x0 = el.xp0 = tVect(x0[3 * e], x0[3 * e + 1], x0[3 * e + 2]);
            el.x1 = el.xp1 = tVect(x1[3 * e], x1[3 * e + 1], x1[3 * e + 2]);
            el.w0 = el.wp0 = el.wGlobal = el.wpGlobal = el.wLocal = el.wpLocal = tVect(w0[3 * e], w0[3 * e + 1], w0[3 * e + 2]);
            el.FParticle = tVect(FP[3 * e], FP[3 * e + 1], FP[3 * e + 2]); el.FWall = tVect(FW[3 * e], FW[3 * e + 1], FW[3 * e + 2]);
            el.MParticle = tVect(MP[3 * e], MP[3 * e + 1], MP[3 * e + 2]); el.MWall = tVect(MW[3 * e], MW[3 * e + 1], MW[3 * e + 2]);
        }
        for (size_t k = 0; k < particles.size(); ++k) {
            const elmt& el = elmts[particles[k].clusterIndex];
            particles[k].x0 = el.x0; particles[k].x1 = el.x1;  // particle::updateCorrected for a one-sphere element
        }
        st.dem->maxDisp = info[0];
    } else {
    if (!st.couplePending) pack(st, elmts, particles);  // computeHydroForces reads the lists it is given (LB.cpp:1851)
    if (lbGpuStep(st.h, st.fsRequested ? 1 : 0, st.couplePending ? 1 : 0, st.rescan ? 1 : 0, st.parts.data(), (uint32_t)st.parts.size(),
                  st.elmts.data(), (uint32_t)st.elmts.size(), st.comps.data(), (uint32_t)st.comps.size()))
        die("lbGpuStep");
    }
    st.fsRequested = false; st.couplePending = false; st.rescan = false; st.demPending = false;
    ++st.steps;
    const size_t nE = elmts.size(), nW = walls.size();
    std::vector<double> F(3 * nE + 1), M(3 * nE + 1), V(nE + 1), W(3 * nW + 1);
    if (lbGpuParticleForces(st.h, F.data(), M.data(), V.data(), W.data())) die("lbGpuParticleForces");
    for (size_t e = 0; e < nE; ++e) {
        elmts[e].FHydro = tVect(F[3 * e], F[3 * e + 1], F[3 * e + 2]);
        elmts[e].MHydro = tVect(M[3 * e], M[3 * e + 1], M[3 * e + 2]);
        elmts[e].fluidVolume = V[e];
    }
    for (size_t w = 0; w < nW; ++w) walls[w].FHydro = tVect(W[3 * w], W[3 * w + 1], W[3 * w + 2]);
#ifdef LBGPU_SHIM_VERIFY
    if (st.verify) { verify_against_host(this, st, refF, elmts); return; }
#endif
    if (needs_host_mirrors() && export_due(this, st)) refresh_mirrors(this, st);
}

// ---------------------------------------------------------------------------------------------
// IO's whole-lattice walks, served by the device (weak in IO.o, see the header comment)
// ---------------------------------------------------------------------------------------------
namespace {
void summary_of(const LB& lb, double out[4]) {
    GpuState& st = state_of(lb);
    if (!st.h) { out[0] = out[1] = out[2] = out[3] = 0.0; return; }
    if (lbGpuFluidSummary(st.h, out)) die("lbGpuFluidSummary");
    ++st.summaries;
}
}  // namespace

void IO::exportMaxSpeedFluid(const LB& lb) {  // IO.cpp:835-851
    double s[4];
    summary_of(lb, s);
    const double maxFluidSpeed = sqrt(s[0]);
    cout << "MaxFSpeed= " << maxFluidSpeed * lb.unit.Speed << "(" << int(lbmDt * lbmDt * maxFluidSpeed * 100.0 / 0.01) << "%)\t";
    exportFile << "MaxFSpeed= " << maxFluidSpeed * lb.unit.Speed << "(" << int(lbmDt * lbmDt * maxFluidSpeed * 100.0 / 0.01) << "%)\t";
    maxSpeedFile.open(maxSpeedFileName.c_str(), ios::app);
    maxSpeedFile << realTime << " " << maxFluidSpeed * lb.unit.Speed << "\n";
    maxSpeedFile.close();
}

void IO::exportTotalMass(const LB& lb) {  // IO.cpp:878-883, 987-999
    double s[4] = { 0.0, 0.0, 0.0, 0.0 };
    if (lbmSolve) summary_of(lb, s);
    const double massTot = s[1];
    cout << "Volume=" << massTot * lb.unit.Volume << "; Mass = " << massTot * lb.unit.Mass << "\t";
    exportFile << "Volume=" << massTot * lb.unit.Volume << "; Mass = " << massTot * lb.unit.Mass << "\t";
}

void IO::exportPlasticity(const LB& lb) {  // IO.cpp:885-895, 969-985
    double s[4];
    summary_of(lb, s);
    const double percPlastic = 100.0 * double((unsigned int)s[3]) / double((unsigned int)s[2]);
    cout << "Plastic =" << int(percPlastic) << "%;\t";
    exportFile << "Plastic =" << int(percPlastic) << "%;\t";
    plasticityFile.open(plasticityFileName.c_str(), ios::app);
    plasticityFile << realTime << " " << percPlastic << "\n";
    plasticityFile.close();
}

void IO::exportParaviewFluidOld(const LB& lb, const string& fluidFile) {  // IO.cpp:698-831
    GpuState& st = state_of(lb);
    if (lbGpuWriteVti(st.h, fluidFile.c_str(), demSolve ? 1 : 0)) die("lbGpuWriteVti");
    ++st.vtis;
}
