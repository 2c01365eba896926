// ref_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Drives the UNMODIFIED reference implementation (gnomeCreative/hybird, sources compiled where
// they lie under /root/reference by oracle/Makefile; outputs only under oracle/_ref/) so that it
// can serve as (i) the ground truth the C restatement in oracle/lb_oracle.c is pinned against,
// (ii) the generator of the golden vectors under tests/golden/ and (iii) the "reference" CPU
// baseline timed by bench.py.
//
// This translation unit is compiled with -fno-access-control so it can call the private phases
// of the reference's LB class one by one.  It contains no code taken from
// the reference: it only *calls* it.  What it replays:
//   * main()'s initialisation order ............ hybird.cpp:299-316
//   * parseConfigFile()'s switches ............. hybird.cpp:132-223
//   * LB::latticeBolzmannInit's call order ..... LB.cpp:190-219 (split so gas can be injected
//                                                between the geometry passes and the interface
//                                                closure loops of LB.cpp:787-814)
//   * the LB part of goCycle() ................. hybird.cpp:47-59
//
// Usage: ref_harness -c case.cfg [-key value ...] --out PREFIX --steps K [options]
//   --dump a,b,c        full state dumps after these LB steps (0 = after init)
//   --types-every       append the 1-byte type|p map after every step to PREFIX_types.bin
//   --types-until K     ... only after init and after the steps <= K (and after every --dump step)
//   --lite              state dumps without the pre-collision f (full-size lattices: 238 instead of 390 B per cell)
//   --fluid-box x0 x1 y0 y1 z0 z1    fluid cells outside the (inclusive) box become gas
//   --gas-box   x0 x1 y0 y1 z0 z1    fluid cells inside the (inclusive) box become gas
//   --gas-sphere cx cy cz r          fluid cells with |pos-c|<r become gas
//   --fluid-sphere cx cy cz r        fluid cells with |pos-c|>=r become gas
//   --wall-vel i vx vy vz            set dem.walls[i].vel (physical units) before LB init
//   --motion none|kin|dem            particle motion between LB steps (default none)
//   --rescan-every K                 raise dem.newNeighborList every K steps (kin/none motion)
//   --time                           print per-phase wall-clock JSON (CPU baseline); no dumps
//   --warmup W                       untimed steps before --time measurement
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <string>
#include <vector>
#include <chrono>
#include <set>
#include <sstream>

#include "getpot.h"
#include "IO.h"
#include "DEM.h"
#include "LB.h"

ProblemName problemName = NONE;

namespace {

struct Box { int lo[3], hi[3]; bool inside; };
struct Sphere { double c[3], r; bool inside; };
struct WallVel { int idx; double v[3]; };

struct Args {
    std::string cfg, out = "ref";
    unsigned steps = 1, warmup = 0, rescanEvery = 0, typesUntil = 0xffffffffu;
    bool lite = false;
    std::set<unsigned> dumps;
    bool typesEvery = false, timeMode = false, dumpNeighbors = false;
    std::string motion = "none";
    std::vector<Box> boxes;
    std::vector<Sphere> spheres;
    std::vector<WallVel> wallVels;
};

double now_s() {
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

Args parseArgs(int argc, char** argv) {
    Args a;
    for (int i = 1; i < argc; ++i) {
        std::string s = argv[i];
        auto need = [&](int k) { if (i + k >= argc) { fprintf(stderr, "missing value for %s\n", s.c_str()); exit(2); } };
        if (s == "-c") { need(1); a.cfg = argv[++i]; }
        else if (s == "--out") { need(1); a.out = argv[++i]; }
        else if (s == "--steps") { need(1); a.steps = atoi(argv[++i]); }
        else if (s == "--warmup") { need(1); a.warmup = atoi(argv[++i]); }
        else if (s == "--rescan-every") { need(1); a.rescanEvery = atoi(argv[++i]); }
        else if (s == "--dump") {
            need(1);
            std::stringstream ss(argv[++i]);
            std::string tok;
            while (std::getline(ss, tok, ',')) a.dumps.insert((unsigned)atoi(tok.c_str()));
        }
        else if (s == "--types-every") a.typesEvery = true;
        else if (s == "--types-until") { need(1); a.typesEvery = true; a.typesUntil = atoi(argv[++i]); }
        else if (s == "--lite") a.lite = true;
        else if (s == "--dump-neighbors") a.dumpNeighbors = true;
        else if (s == "--time") a.timeMode = true;
        else if (s == "--motion") { need(1); a.motion = argv[++i]; }
        else if (s == "--fluid-box" || s == "--gas-box") {
            need(6);
            Box b; b.inside = (s == "--gas-box");
            for (int k = 0; k < 3; ++k) { b.lo[k] = atoi(argv[++i]); b.hi[k] = atoi(argv[++i]); }
            a.boxes.push_back(b);
        }
        else if (s == "--fluid-sphere" || s == "--gas-sphere") {
            need(4);
            Sphere sp; sp.inside = (s == "--gas-sphere");
            for (int k = 0; k < 3; ++k) sp.c[k] = atof(argv[++i]);
            sp.r = atof(argv[++i]);
            a.spheres.push_back(sp);
        }
        else if (s == "--wall-vel") {
            need(4);
            WallVel w; w.idx = atoi(argv[++i]);
            for (int k = 0; k < 3; ++k) w.v[k] = atof(argv[++i]);
            a.wallVels.push_back(w);
        }
    }
    if (a.cfg.empty()) { fprintf(stderr, "need -c cfg\n"); exit(2); }
    return a;
}

// The switches main() reads before handing the GetPot objects to LB and DEM (hybird.cpp:132-223).
void readSwitches(IO& io, DEM& dem, LB& lb, GetPot& cfgFile, GetPot& command_line) {
    PARSE_CLASS_MEMBER(cfgFile, io.problemNameString, "problemName", "none");
    const std::string& p = io.problemNameString;
    if (p == "DRUM") problemName = DRUM; else if (p == "SHEARCELL") problemName = SHEARCELL;
    else if (p == "NONE") problemName = NONE; else if (p == "AVALANCHE") problemName = AVALANCHE;
    else if (p == "SPLASH") problemName = SPLASH; else if (p == "BOX") problemName = BOX;
    else if (p == "NET") problemName = NET; else if (p == "DIFF") problemName = DIFF;
    else if (p == "BARRIER") problemName = BARRIER; else if (p == "demChute") problemName = demChute;
    PARSE_CLASS_MEMBER(cfgFile, dem.demInitialRepeat, "demInitialRepeat", 0.0);
    PARSE_CLASS_MEMBER(cfgFile, io.demSolve, "demSolve", 0);
    PARSE_CLASS_MEMBER(cfgFile, io.lbmSolve, "lbSolve", 0);
    PARSE_CLASS_MEMBER(cfgFile, lb.freeSurface, "freeSurfaceSolve", 0);
    PARSE_CLASS_MEMBER(cfgFile, lb.forceField, "forceFieldSolve", 0);
    PARSE_CLASS_MEMBER(cfgFile, lb.nonNewtonian, "nonNewtonianSolve", 0);
    PARSE_CLASS_MEMBER(cfgFile, lb.turbulenceOn, "turbulenceSolve", 0);
    lb.latticeBoltzmannGet(cfgFile, command_line);
    dem.discreteElementGet(cfgFile, command_line);
    switch (problemName) {
        case DRUM:
            PARSE_CLASS_MEMBER(cfgFile, dem.drumSpeed, "drumSpeed", 0.0);
            PARSE_CLASS_MEMBER(cfgFile, lb.fluidMass, "fluidMass", 0.0);
            break;
        case NET:
        case BARRIER:
            PARSE_CLASS_MEMBER(cfgFile, lb.avalanchePosit, "avalanchePosit", 0.0);
            break;
        default: break;
    }
    io.saveCount = 0;
    io.maximumTimeSteps = 0;
}

void injectGas(LB& lb, const Args& a) {
    for (unsigned it = 0; it < lb.totNodes; ++it) {
        if (!lb.types[it].isFluid()) continue;
        const int c[3] = { (int)lb.getX(it), (int)lb.getY(it), (int)lb.getZ(it) };
        bool gas = false;
        for (const Box& b : a.boxes) {
            bool in = true;
            for (int k = 0; k < 3; ++k) in = in && c[k] >= b.lo[k] && c[k] <= b.hi[k];
            if (b.inside ? in : !in) gas = true;
        }
        for (const Sphere& s : a.spheres) {
            double d2 = 0.0;
            for (int k = 0; k < 3; ++k) d2 += (c[k] - s.c[k]) * (c[k] - s.c[k]);
            const bool in = d2 < s.r * s.r;
            if (s.inside ? in : !in) gas = true;
        }
        if (gas) lb.types[it].setGas();
    }
}

// LB::latticeBolzmannInit (LB.cpp:190-219) with initializeTypes (LB.cpp:324-340) opened up so
// the gas region can be injected before the closure loops inside initializeInterface.
void initLB(LB& lb, DEM& dem, const Args& a) {
    lb.initializeNodes();
    lb.initializeLatticeBoundaries();
    lb.initializeParticleBoundaries(dem.particles);
    lb.initializeWallBoundaries(dem.walls);
    lb.initializeCylinderBoundaries(dem.cylinders);
    lb.initializeObjectBoundaries(dem.objects);
    lb.initializeCurved(dem.cylinders);
    injectGas(lb, a);
    lb.initializeInterface(dem.particles.size());
    lb.initializeLists();
    lb.initializeVariables();
    lb.initializeWalls(dem.walls, dem.cylinders, dem.objects);
    lb.cleanLists();
    lb.totalMass = 0.0;
    if (problemName == DRUM) {  // LB.cpp:205-209
        lb.totalMass = lb.fluidMass / lb.unit.Mass;
    } else {
        for (unsigned it = 0; it < lb.nodes.size(); ++it)
            if (lb.types[it].isActive() && !lb.types[it].isInsideParticle()) lb.totalMass += lb.nodes[it]->mass;
    }
}

template <class T> void wr(FILE* f, const T* p, size_t n) {
    if (fwrite(p, sizeof(T), n, f) != n) { perror("fwrite"); exit(3); }
}

void dumpState(const LB& lb, const DEM& dem, const std::string& path, unsigned step, bool withNeighbors, bool lite = false) {
    FILE* fp = fopen(path.c_str(), "wb");
    if (!fp) { perror(path.c_str()); exit(3); }
    const unsigned N = lb.totNodes;
    const char magic[8] = { 'H', 'B', 'D', 'U', 'M', 'P', '2', 0 };
    wr(fp, magic, 8);
    uint32_t hdr[8] = { lb.lbSize[0], lb.lbSize[1], lb.lbSize[2], step, (uint32_t)dem.elmts.size(),
                        (uint32_t)dem.walls.size(), (uint32_t)withNeighbors | (lite ? 2u : 0u), (uint32_t)dem.particles.size() };
    wr(fp, hdr, 8);
    std::vector<uint8_t> t(N), fl(N);
    std::vector<uint32_t> si(N);
    for (unsigned i = 0; i < N; ++i) {
        t[i] = (uint8_t)lb.types[i].getType();
        fl[i] = (uint8_t)((lb.types[i].isInsideParticle() ? 1 : 0) | (lb.nodes[i] != 0 ? 2 : 0));
        si[i] = lb.types[i].getSolidIndex();
    }
    wr(fp, t.data(), N); wr(fp, fl.data(), N); wr(fp, si.data(), N);
    std::vector<double> buf((size_t)19 * N);
    auto field = [&](int width, void (*get)(const node&, double*)) {
        std::fill(buf.begin(), buf.begin() + (size_t)width * N, 0.0);
        for (unsigned i = 0; i < N; ++i) if (lb.nodes[i]) get(*lb.nodes[i], &buf[(size_t)width * i]);
        wr(fp, buf.data(), (size_t)width * N);
    };
    if (!lite) field(19, [](const node& nd, double* o) { for (int j = 0; j < 19; ++j) o[j] = nd.f[j]; });
    field(19, [](const node& nd, double* o) { for (int j = 0; j < 19; ++j) o[j] = nd.fs[j]; });
    field(1, [](const node& nd, double* o) { o[0] = nd.n; });
    field(3, [](const node& nd, double* o) { o[0] = nd.u.x; o[1] = nd.u.y; o[2] = nd.u.z; });
    field(3, [](const node& nd, double* o) { o[0] = nd.hydroForce.x; o[1] = nd.hydroForce.y; o[2] = nd.hydroForce.z; });
    field(1, [](const node& nd, double* o) { o[0] = nd.mass; });
    field(1, [](const node& nd, double* o) { o[0] = nd.visc; });
    field(1, [](const node& nd, double* o) { o[0] = nd.shearRate; });
    if (withNeighbors) {
        std::vector<uint32_t> d((size_t)19 * N);
        for (unsigned i = 0; i < N; ++i) for (int j = 0; j < 19; ++j) d[(size_t)19 * i + j] = lb.neighbors[i].d[j];
        wr(fp, d.data(), d.size());
    }
    // trailer: curved-wall cells (LB::curves, node.h:133-148): count, cell indices, delta[19] each (delta[0] is never
    // initialised by the reference nor read: written as 0), then totalMass (LB::enforceMassConservation's target)
    {
        const char cm[8] = { 'C', 'U', 'R', 'V', 'E', 'S', '1', 0 };
        wr(fp, cm, 8);
        std::vector<uint32_t> idx;
        for (unsigned i = 0; i < N; ++i) if (lb.curves[i] != 0) idx.push_back(i);
        uint32_t nc = (uint32_t)idx.size();
        wr(fp, &nc, 1);
        wr(fp, idx.data(), idx.size());
        std::vector<double> dl((size_t)19 * idx.size(), 0.0);
        for (size_t k = 0; k < idx.size(); ++k) for (int j = 1; j < 19; ++j) dl[19 * k + j] = lb.curves[idx[k]]->delta[j];
        wr(fp, dl.data(), dl.size());
        double tm = lb.totalMass;
        wr(fp, &tm, 1);
    }
    fclose(fp);
}

// What the LB calls of one step consume from DEM ("in"), recorded so a test
// can replay the identical inputs into another implementation.
void traceParticles(FILE* fp, const DEM& dem, bool newNeighborList) {
    uint32_t hdr[3] = { (uint32_t)dem.particles.size(), (uint32_t)dem.elmts.size(), (uint32_t)newNeighborList };
    wr(fp, hdr, 3);
    for (const particle& p : dem.particles) {
        double d[7] = { p.x0.x, p.x0.y, p.x0.z, p.r, p.radiusVec.x, p.radiusVec.y, p.radiusVec.z };
        uint32_t u[2] = { p.clusterIndex, p.particleIndex };
        wr(fp, d, 7); wr(fp, u, 2);
    }
    for (const elmt& e : dem.elmts) {
        double d[6] = { e.x1.x, e.x1.y, e.x1.z, e.wGlobal.x, e.wGlobal.y, e.wGlobal.z };
        wr(fp, d, 6);
        uint32_t nc = (uint32_t)e.components.size();
        wr(fp, &nc, 1);
        for (uint32_t k = 0; k < nc; ++k) { uint32_t c = (uint32_t)e.components[k]; wr(fp, &c, 1); }
    }
}

void moveKinematic(DEM& dem, double dt) {
    for (elmt& e : dem.elmts) {
        e.x0 = e.x0 + e.x1 * dt;
        e.xp0 = e.x0;
    }
    for (particle& p : dem.particles) {
        const elmt& e = dem.elmts[p.clusterIndex];
        p.x0 = e.x0 + p.radiusVec;
        p.x1 = e.x1 + e.wGlobal.cross(p.radiusVec);
    }
}

}  // namespace

int main(int argc, char** argv) {
    Args a = parseArgs(argc, argv);
    IO io;
    DEM dem;
    LB lb;
    GetPot command_line(argc, argv);
    GetPot cfgFile(a.cfg);
    io.lbmCfgName = a.cfg;
    readSwitches(io, dem, lb, cfgFile, command_line);
    io.currentTimeStep = 0;

    lb.latticeDefinition();
    dem.discreteElementInit(lb.boundary, lb.lbSize, lb.unit, lb.lbF);
    for (const WallVel& w : a.wallVels) {
        if (w.idx < 0 || w.idx >= (int)dem.walls.size()) { fprintf(stderr, "bad wall index\n"); return 2; }
        dem.walls[w.idx].vel = tVect(w.v[0], w.v[1], w.v[2]);
    }
    initLB(lb, dem, a);
    lb.time = 0;
    dem.demTime = 0.0;
    dem.demTimeStep = 0;

    const unsigned N = lb.totNodes;
    FILE* typesFp = 0; FILE* elmtFp = 0; FILE* partFp = 0; FILE* logFp = 0;
    if (!a.timeMode) {
        if (a.typesEvery) typesFp = fopen((a.out + "_types.bin").c_str(), "wb");
        elmtFp = fopen((a.out + "_forces.bin").c_str(), "wb");
        partFp = fopen((a.out + "_parts.bin").c_str(), "wb");
        logFp = fopen((a.out + "_log.txt").c_str(), "w");
        fprintf(logFp, "# size %u %u %u nodes %u elmts %zu parts %zu walls %zu unitLength %.17g unitTime %.17g unitDensity %.17g\n",
                lb.lbSize[0], lb.lbSize[1], lb.lbSize[2], N, dem.elmts.size(), dem.particles.size(), dem.walls.size(),
                lb.unit.Length, lb.unit.Time, lb.unit.Density);
        fprintf(logFp, "# lbF %.17g %.17g %.17g initVisc %.17g plasticVisc %.17g yieldStress %.17g turbConst %.17g slip %.17g\n",
                lb.lbF.x, lb.lbF.y, lb.lbF.z, lb.initDynVisc, lb.plasticVisc, lb.yieldStress, lb.turbConst, lb.slipCoefficient);
        fprintf(logFp, "# initVelocity %.17g %.17g %.17g boundary %u %u %u %u %u %u flags fs %d ff %d nn %d turb %d\n",
                lb.initVelocity.x, lb.initVelocity.y, lb.initVelocity.z, lb.boundary[0], lb.boundary[1], lb.boundary[2],
                lb.boundary[3], lb.boundary[4], lb.boundary[5], (int)lb.freeSurface, (int)lb.forceField,
                (int)lb.nonNewtonian, (int)lb.turbulenceOn);
        fprintf(logFp, "# totalMass %.17g enforceMass %d\n", lb.totalMass, (int)(problemName == DRUM));
        // what a device-side DEM needs to restate DEM::discreteElementStep for single-sphere elements (tests only)
        if (FILE* df = fopen((a.out + "_dem.txt").c_str(), "w")) {
            fprintf(df, "contactModel %d knConst %.17g ksConst %.17g dampCoeff %.17g viscTang %.17g linearStiff %.17g frictionCoefPart %.17g "
                        "frictionCoefWall %.17g numVisc %.17g demF %.17g %.17g %.17g deltat %.17g multiStep %u nebrRange %.17g maxDisp %.17g\n",
                    (int)(dem.sphereMat.contactModel == HERTZIAN ? 1 : 0), dem.sphereMat.knConst, dem.sphereMat.ksConst, dem.sphereMat.dampCoeff,
                    dem.sphereMat.viscTang, dem.sphereMat.linearStiff, dem.sphereMat.frictionCoefPart, dem.sphereMat.frictionCoefWall, dem.numVisc,
                    dem.demF.x, dem.demF.y, dem.demF.z, dem.deltat, dem.multiStep, dem.nebrRange, dem.maxDisp);
            for (const elmt& e : dem.elmts)
                fprintf(df, "elmt %u size %u radius %.17g m %.17g I %.17g %.17g %.17g x0 %.17g %.17g %.17g x1 %.17g %.17g %.17g w0 %.17g %.17g %.17g\n",
                        e.index, e.size, e.radius, e.m, e.I.x, e.I.y, e.I.z, e.x0.x, e.x0.y, e.x0.z, e.x1.x, e.x1.y, e.x1.z, e.w0.x, e.w0.y, e.w0.z);
            for (const wall& w : dem.walls)
                fprintf(df, "wall %u n %.17g %.17g %.17g p %.17g %.17g %.17g vel %.17g %.17g %.17g omega %.17g %.17g %.17g rotCenter %.17g %.17g %.17g moving %d\n",
                        w.index, w.n.x, w.n.y, w.n.z, w.p.x, w.p.y, w.p.z, w.vel.x, w.vel.y, w.vel.z, w.omega.x, w.omega.y, w.omega.z,
                        w.rotCenter.x, w.rotCenter.y, w.rotCenter.z, (int)w.moving);
            for (const pbc& b : dem.pbcs)
                fprintf(df, "pbc %d p %.17g %.17g %.17g v %.17g %.17g %.17g\n", b.index, b.p.x, b.p.y, b.p.z, b.v.x, b.v.y, b.v.z);
            fprintf(df, "pbcs %zu cylinders %zu objects %zu ghosts %zu\n", dem.pbcs.size(), dem.cylinders.size(), dem.objects.size(), dem.ghosts.size());
            fclose(df);
        }
        if (a.dumps.count(0)) dumpState(lb, dem, a.out + "_state000000.bin", 0, a.dumpNeighbors, a.lite);
    }
    auto writeTypes = [&]() {
        std::vector<uint8_t> t(N);
        for (unsigned i = 0; i < N; ++i) t[i] = (uint8_t)(lb.types[i].getType() | (lb.types[i].isInsideParticle() ? 16 : 0));
        wr(typesFp, t.data(), N);
    };
    if (typesFp) writeTypes();

    double tPhase[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    double tTotal = 0.0;
    unsigned long long activeSum = 0;
    const unsigned total = a.steps + (a.timeMode ? a.warmup : 0);
    for (unsigned s = 1; s <= total; ++s) {
        ++io.currentTimeStep;
        ++lb.time;
        // particle motion (the DEM side of goCycle, hybird.cpp:43-45)
        if (a.motion == "dem") dem.discreteElementStep(io);
        else {
            if (a.motion == "kin") moveKinematic(dem, lb.unit.Time);
            dem.demTime += lb.unit.Time;
            if (a.rescanEvery && s % a.rescanEvery == 0) dem.newNeighborList = true;
        }
        if (partFp) traceParticles(partFp, dem, dem.newNeighborList);
        const bool timed = a.timeMode && s > a.warmup;
        // the LB side of goCycle (hybird.cpp:47-59), phases opened up for timing
        double t0 = now_s();
        if (lb.freeSurface) {
            lb.updateMass();
            double t1 = now_s(); if (timed) tPhase[0] += t1 - t0;
            lb.updateInterface();
            if (problemName == DRUM) lb.enforceMassConservation();
            double t2 = now_s(); if (timed) tPhase[1] += t2 - t1;
        }
        double t3 = now_s();
        if (io.demSolve) lb.latticeBoltzmannCouplingStep(dem.newNeighborList, dem.elmts, dem.particles);
        double t4 = now_s(); if (timed) tPhase[2] += t4 - t3;
        lb.cleanLists();
        double t5 = now_s(); if (timed) tPhase[3] += t5 - t4;
        lb.reconstruction();
        double t6 = now_s(); if (timed) tPhase[4] += t6 - t5;
        lb.computeHydroForces(dem.elmts, dem.particles);
        double t7 = now_s(); if (timed) tPhase[5] += t7 - t6;
        lb.collision();
        double t8 = now_s(); if (timed) tPhase[6] += t8 - t7;
        lb.streaming(dem.elmts, dem.particles, dem.walls);
        double t9 = now_s(); if (timed) tPhase[7] += t9 - t8;
        if (timed) { tTotal += t9 - t0; activeSum += lb.activeNodes.size(); }

        if (!a.timeMode) {
            for (const elmt& e : dem.elmts) {
                double d[7] = { e.FHydro.x, e.FHydro.y, e.FHydro.z, e.MHydro.x, e.MHydro.y, e.MHydro.z, e.fluidVolume };
                wr(elmtFp, d, 7);
            }
            for (const wall& w : dem.walls) { double d[3] = { w.FHydro.x, w.FHydro.y, w.FHydro.z }; wr(elmtFp, d, 3); }
            if (typesFp && (s <= a.typesUntil || a.dumps.count(s))) writeTypes();
            double mass = 0.0;
            for (unsigned i = 0; i < N; ++i)
                if (lb.types[i].isActive() && !lb.types[i].isInsideParticle()) mass += lb.nodes[i]->mass;
            fprintf(logFp, "step %u fluid %zu interface %zu particleNodes %zu mass %.17g\n", s, lb.fluidNodes.size(),
                    lb.interfaceNodes.size(), lb.particleNodes.size(), mass);
            if (a.dumps.count(s)) {
                char name[64];
                snprintf(name, sizeof name, "_state%06u.bin", s);
                dumpState(lb, dem, a.out + name, s, false, a.lite);
            }
        }
    }
    if (a.timeMode) {
        int threads = 1;
#ifdef _OPENMP
        threads = omp_get_max_threads();
#endif
        const double ms = 1e3 * tTotal / a.steps;
        printf("{\"impl\": \"reference\", \"threads\": %d, \"steps\": %u, \"warmup\": %u, \"nodes\": %u, "
               "\"active_mean\": %.1f, \"ms_per_step\": %.6f, \"mlups_active\": %.6f, \"mlups_all\": %.6f, "
               "\"phase_ms\": {\"updateMass\": %.4f, \"updateInterface\": %.4f, \"coupling\": %.4f, \"cleanLists\": %.4f, "
               "\"reconstruction\": %.4f, \"computeHydroForces\": %.4f, \"collision\": %.4f, \"streaming\": %.4f}}\n",
               threads, a.steps, a.warmup, N, (double)activeSum / a.steps, ms,
               (double)activeSum / a.steps / (ms * 1e3), (double)N / (ms * 1e3),
               1e3 * tPhase[0] / a.steps, 1e3 * tPhase[1] / a.steps, 1e3 * tPhase[2] / a.steps, 1e3 * tPhase[3] / a.steps,
               1e3 * tPhase[4] / a.steps, 1e3 * tPhase[5] / a.steps, 1e3 * tPhase[6] / a.steps, 1e3 * tPhase[7] / a.steps);
    }
    if (typesFp) fclose(typesFp);
    if (elmtFp) fclose(elmtFp);
    if (partFp) fclose(partFp);
    if (logFp) fclose(logFp);
    return 0;
}
