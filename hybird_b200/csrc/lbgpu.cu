// lbgpu.cu -- host side of liblbgpu.so: the C ABI of include/lbgpu.h on top of lb_kernels.cuh.
//
// One handle = one CUDA device = one slab of the lattice.  All work of a handle is issued on its
// own stream; lbGpuStep is asynchronous except for the particle flood fill, which needs a 4-byte
// read-back per generation.  There is deliberately no CPU path: without a device every entry
// point returns LBGPU_ENODEVICE.
#include "../../include/lbgpu.h"

#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "lb_kernels.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) return fail(LBGPU_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

lb::FastDiv make_div(uint32_t d) {
    lb::FastDiv f;
    f.d = d;
    if (d <= 1) { f.mul = 0; f.shr = 0; return f; }
    uint32_t s = 0;
    while ((1ull << s) < d) ++s;  // s = ceil(log2 d) >= 1
    const unsigned long long num = 1ull << (31 + s);
    f.mul = (uint32_t)((num + d - 1) / d);
    f.shr = s - 1;
    return f;
}

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t alloc(size_t count) {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        return cudaMalloc((void**)&p, count * sizeof(T));
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    ~DevBuf() { release(); }
};

}  // namespace

struct LbGpuHandle {
    LbGpuParams prm;
    int device = 0;
    uint32_t N = 0;
    size_t stride = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t evA = nullptr, evB = nullptr;
    lb::Dev dev;  // template for kernel parameters (pointers filled per launch)
    DevBuf<double> fA, fB, n, ux, uy, uz, mass, newMass, visc, shearRate, hfx, hfy, hfz;
    DevBuf<uint8_t> type0, type1, mark;
    DevBuf<uint32_t> solidIndex;
    DevBuf<double> partial, sums, scal, elemOut, wallOut;
    DevBuf<unsigned long long> counters;  // [0] nInterface [1..3] k_count scratch
    DevBuf<uint32_t> status;              // [0] type error, [1] flood-fill counter
    DevBuf<lb::RawParticle> rawParts;
    DevBuf<lb::RawElement> rawElmts;
    DevBuf<lb::Particle> parts;
    DevBuf<lb::Element> elmts;
    DevBuf<uint32_t> comps;
    void* pinned = nullptr;
    size_t pinnedBytes = 0;
    uint32_t* pinnedStatus = nullptr;
    uint32_t nParts = 0, nElmts = 0, nComps = 0;
    int cur = 0;      // population buffer holding the latest post-collision state (0 = A)
    int curType = 0;  // type buffer holding the current types
    bool fs = false, shear = false, force = false, macroAlways = false, dynWall = false;
    bool macroValid = true, lastStepFirst = false, lastStepCoupled = false, typesFlipped = false;
    uint64_t steps = 0, launches = 0;
    uint32_t blocks = 0;
    double uLength = 1, uSpeed = 1, uAngVel = 1, uForce = 1, uTorque = 1, uVolume = 1;
    float lastMs = 0.f;
    // CUDA-event pairs around the fused step kernel of the last lbGpuStep/lbGpuRun call (ring of KEV)
    static constexpr uint32_t KEV = 512;
    std::vector<cudaEvent_t> kev0, kev1;
    uint32_t kevCount = 0;

    double* fbuf(int k) { return k == 0 ? fA.p : fB.p; }
    uint8_t* tbuf(int k) { return k == 0 ? type0.p : type1.p; }
};

namespace {

using namespace lb;

typedef void (*StepKernel)(const Dev);

// FS / DYNWALL are compile-time in k_step; only the combinations that can occur are instantiated.
template <bool FORCE, bool SHEAR, bool MACRO, bool COUPLE>
StepKernel pick_step(bool fsOn, bool dyn) {
    if constexpr (!MACRO) {
        return k_step<FORCE, SHEAR, false, COUPLE, false, false>;
    } else {
        if (fsOn) return dyn ? k_step<FORCE, SHEAR, true, COUPLE, true, true> : k_step<FORCE, SHEAR, true, COUPLE, true, false>;
        return dyn ? k_step<FORCE, SHEAR, true, COUPLE, false, true> : k_step<FORCE, SHEAR, true, COUPLE, false, false>;
    }
}

StepKernel select_step(bool force, bool shear, bool macro, bool couple, bool fsOn, bool dyn) {
    if (!macro) {
        if (couple) return shear ? pick_step<true, true, false, true>(false, false) : pick_step<true, false, false, true>(false, false);
        if (force) return shear ? pick_step<true, true, false, false>(false, false) : pick_step<true, false, false, false>(false, false);
        return shear ? pick_step<false, true, false, false>(false, false) : pick_step<false, false, false, false>(false, false);
    }
    // the full variants always carry the force path (exact when the force is zero)
    if (couple) return shear ? pick_step<true, true, true, true>(fsOn, dyn) : pick_step<true, false, true, true>(fsOn, dyn);
    return shear ? pick_step<true, true, true, false>(fsOn, dyn) : pick_step<true, false, true, false>(fsOn, dyn);
}

Dev dev_for(LbGpuHandle* h, bool fsStep) {
    Dev d = h->dev;
    d.fsrc = h->fbuf(h->cur);
    d.fdst = h->fbuf(h->cur ^ 1);
    if (fsStep) {
        d.typeOld = h->tbuf(h->curType);
        d.type = h->tbuf(h->curType ^ 1);
    } else {
        d.typeOld = h->tbuf(h->curType);
        d.type = h->tbuf(h->curType);
    }
    d.parts = h->parts.p; d.elmts = h->elmts.p; d.comps = h->comps.p;
    d.nParts = h->nParts; d.nElmts = h->nElmts;
    return d;
}

int ensure_pinned(LbGpuHandle* h, size_t bytes) {
    if (bytes <= h->pinnedBytes) return 0;
    if (h->pinned) cudaFreeHost(h->pinned);
    h->pinned = nullptr; h->pinnedBytes = 0;
    size_t want = bytes + bytes / 2 + 4096;
    CU(cudaMallocHost(&h->pinned, want));
    h->pinnedBytes = want;
    return 0;
}

int upload_particles(LbGpuHandle* h, const LbGpuParticle* parts, uint32_t nParts, const LbGpuElement* elmts,
                     uint32_t nElmts, const uint32_t* comps, uint32_t nComps) {
    static_assert(sizeof(LbGpuParticle) == sizeof(RawParticle) && sizeof(LbGpuElement) == sizeof(RawElement), "ABI layout");
    if (nParts > h->rawParts.n) { CU(h->rawParts.alloc(nParts + nParts / 2 + 16)); CU(h->parts.alloc(h->rawParts.n)); }
    if (nElmts > h->rawElmts.n) {
        CU(h->rawElmts.alloc(nElmts + nElmts / 2 + 16));
        CU(h->elmts.alloc(h->rawElmts.n));
        CU(h->elemOut.alloc(h->rawElmts.n * 7));
    }
    if (nComps > h->comps.n) CU(h->comps.alloc(nComps + nComps / 2 + 16));
    const size_t bP = sizeof(RawParticle) * nParts, bE = sizeof(RawElement) * nElmts, bC = sizeof(uint32_t) * nComps;
    if (int rc = ensure_pinned(h, bP + bE + bC)) return rc;
    // the previous step's async copies out of the staging buffer must have completed
    CU(cudaStreamSynchronize(h->stream));
    char* st = (char*)h->pinned;
    if (bP) memcpy(st, parts, bP);
    if (bE) memcpy(st + bP, elmts, bE);
    if (bC) memcpy(st + bP + bE, comps, bC);
    if (bP) CU(cudaMemcpyAsync(h->rawParts.p, st, bP, cudaMemcpyHostToDevice, h->stream));
    if (bE) CU(cudaMemcpyAsync(h->rawElmts.p, st + bP, bE, cudaMemcpyHostToDevice, h->stream));
    if (bC) CU(cudaMemcpyAsync(h->comps.p, st + bP + bE, bC, cudaMemcpyHostToDevice, h->stream));
    h->nParts = nParts; h->nElmts = nElmts; h->nComps = nComps;
    const uint32_t m = nParts > nElmts ? nParts : nElmts;
    if (m) {
        k_prepare_particles<<<(m + 127) / 128, 128, 0, h->stream>>>(h->rawParts.p, nParts, h->rawElmts.p, nElmts, h->uLength,
                                                                   h->uSpeed, h->parts.p, h->elmts.p);
        ++h->launches;
    }
    return 0;
}

int free_surface_step(LbGpuHandle* h) {
    Dev d = dev_for(h, true);
    const uint32_t B = h->blocks;
    cudaStream_t s = h->stream;
    d.pull = h->steps > 0;
    k_fs_mass<<<B, BLOCK, 0, s>>>(d);
    k_fs_mutate<<<B, BLOCK, 0, s>>>(d, h->mark.p);
    k_fs_smooth<<<B, BLOCK, 0, s>>>(d, h->mark.p, h->partial.p);
    k_fs_isolated<0><<<B, BLOCK, 0, s>>>(d, h->partial.p + B, h->counters.p);
    CU(cudaMemsetAsync(h->counters.p, 0, sizeof(unsigned long long), s));
    k_fs_isolated<1><<<B, BLOCK, 0, s>>>(d, h->partial.p + 2 * (size_t)B, h->counters.p);
    k_reduce_partials<<<1, 1024, 0, s>>>(h->partial.p, B, 3, h->sums.p, 0);
    k_fs_finalize<<<1, 1, 0, s>>>(h->sums.p, h->counters.p, h->scal.p);
    k_redistribute<<<B, BLOCK, 0, s>>>(d, h->scal.p + 1);
    h->launches += 8;
    h->curType ^= 1;
    h->typesFlipped = true;
    CU(cudaGetLastError());
    return 0;
}

int coupling_step(LbGpuHandle* h, bool rescan) {
    Dev d = dev_for(h, false);
    const uint32_t B = h->blocks;
    cudaStream_t s = h->stream;
    if (h->nParts == 0) {
        if (rescan) { k_clear_p<<<B, BLOCK, 0, s>>>(d); ++h->launches; }
        return 0;
    }
    const uint32_t wb = (h->nParts * 32 + BLOCK - 1) / BLOCK;
    if (rescan) {
        k_clear_p<<<B, BLOCK, 0, s>>>(d);
        k_rescan<0><<<wb, BLOCK, 0, s>>>(d);
        k_rescan<1><<<wb, BLOCK, 0, s>>>(d);
        h->launches += 3;
    }
    k_find_new_active<<<B, BLOCK, 0, s>>>(d);
    ++h->launches;
    for (int gen = 0; gen < 4096; ++gen) {
        CU(cudaMemsetAsync(h->status.p + 1, 0, sizeof(uint32_t), s));
        k_find_new_solid<<<B, BLOCK, 0, s>>>(d, h->status.p + 1);
        k_commit_pending<<<B, BLOCK, 0, s>>>(d);
        h->launches += 2;
        CU(cudaMemcpyAsync(h->pinnedStatus, h->status.p + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        if (*h->pinnedStatus == 0) break;
    }
    CU(cudaGetLastError());
    return 0;
}

int lb_step(LbGpuHandle* h) {
    const bool first = (h->steps == 0);
    const bool couple = h->nParts > 0;
    const bool macro = h->macroAlways;
    const bool fsOn = h->fs;
    Dev d = dev_for(h, false);
    // the streaming being evaluated happened under the type map of before this cycle's free-surface step
    if (h->typesFlipped) d.typeOld = h->tbuf(h->curType ^ 1);
    h->typesFlipped = false;
    d.pull = !first;
    const uint32_t B = h->blocks;
    cudaStream_t s = h->stream;
    const int nSums = 1 + 3 * h->prm.nWalls;
    if (h->dynWall) CU(cudaMemsetAsync(h->partial.p, 0, sizeof(double) * (size_t)B * nSums, s));
    StepKernel k = select_step(h->force, h->shear, macro, couple, fsOn, h->dynWall);
    const uint32_t ke = h->kevCount % LbGpuHandle::KEV;
    CU(cudaEventRecord(h->kev0[ke], s));
    k<<<B, BLOCK, 0, s>>>(d);
    CU(cudaEventRecord(h->kev1[ke], s));
    ++h->kevCount;
    ++h->launches;
    if (h->dynWall) {
        k_reduce_partials<<<1, 1024, 0, s>>>(h->partial.p, B, nSums, h->sums.p + 4, 0);
        ++h->launches;
        if (fsOn) {
            // LB::redistributeMass(extraMass) at the end of LB::streaming (LB.cpp:1477)
            k_extra_mass_finalize<<<1, 1, 0, s>>>(h->sums.p + 4, h->counters.p, h->scal.p + 2);
            Dev dn = d;
            k_redistribute<<<B, BLOCK, 0, s>>>(dn, h->scal.p + 2);
            h->launches += 2;
        }
    }
    if (h->nElmts > 0 && couple) {
        const uint32_t eb = (h->nElmts * 32 + BLOCK - 1) / BLOCK;
        k_element_forces<<<eb, BLOCK, 0, s>>>(d, h->uForce, h->uTorque, h->uVolume, h->elemOut.p);
        ++h->launches;
    }
    CU(cudaGetLastError());
    h->cur ^= 1;
    h->macroValid = macro;
    h->lastStepFirst = first;
    h->lastStepCoupled = couple;
    ++h->steps;
    return 0;
}

int check_status(LbGpuHandle* h) {
    CU(cudaMemcpyAsync(h->pinnedStatus, h->status.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (*h->pinnedStatus) {
        const unsigned t = *h->pinnedStatus - 1;
        CU(cudaMemsetAsync(h->status.p, 0, sizeof(uint32_t), h->stream));
        return fail(LBGPU_ETYPE, "TYPE ERROR: an active cell links to a cell of type %u (LB.cpp:1458-1461)", t);
    }
    return 0;
}

}  // namespace

extern "C" {

const char* lbGpuLastError(void) { return g_err.c_str(); }
int lbGpuAbiVersion(void) { return LBGPU_ABI_VERSION; }

int lbGpuDeviceCount(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int lbGpuInit(const LbGpuParams* prm, const uint8_t* type_flags, const uint32_t* solidIndex, const double* f,
              const double* n, const double* u, const double* mass, const double* visc, LbGpuHandle** out) {
    if (!prm || !type_flags || !solidIndex || !n || !u || !mass || !visc || !out) return fail(LBGPU_EINVAL, "lbGpuInit: null argument");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(LBGPU_ENODEVICE, "lbGpuInit: no CUDA device available (this engine has no CPU fallback)");
    }
    for (int k = 0; k < 3; ++k)
        if (prm->size[k] < 3) return fail(LBGPU_EINVAL, "lbGpuInit: lbSize[%d]=%d < 3", k, prm->size[k]);
    const unsigned long long N64 = (unsigned long long)prm->size[0] * prm->size[1] * prm->size[2];
    if (N64 >= (1ull << 31)) return fail(LBGPU_EINVAL, "lbGpuInit: %llu cells exceed the 2^31 index range", N64);
    for (int a = 0; a < 3; ++a) {
        const bool lo = prm->boundary[2 * a] == T_PERIODIC, hi = prm->boundary[2 * a + 1] == T_PERIODIC;
        if (lo != hi) return fail(LBGPU_EINVAL, "lbGpuInit: boundary%d/%d must both be periodic (4) or neither", 2 * a, 2 * a + 1);
    }
    if (prm->nSlabs > 1) return fail(LBGPU_EUNSUPPORTED, "lbGpuInit: slab decomposition is driven by lbGpuInitSlab");
    if (prm->nWalls < 0 || prm->nWalls > 64) return fail(LBGPU_EINVAL, "lbGpuInit: nWalls=%d out of range", prm->nWalls);
    const uint32_t N = (uint32_t)N64;
    bool anyDyn = false, anyGas = false, anyIface = false;
    for (uint32_t i = 0; i < N; ++i) {
        const int t = type_flags[i] & LBGPU_TYPE_MASK;
        if (t == T_CURVED) return fail(LBGPU_EUNSUPPORTED, "lbGpuInit: curved walls (type 9, LB.cpp:1278-1319) are not implemented");
        if (t == 1 || t > 9) return fail(LBGPU_EINVAL, "lbGpuInit: cell %u has undefined type %d", i, t);
        anyDyn |= (t == T_DYN_WALL || t == T_SLIP_DYN);
        anyGas |= (t == T_GAS);
        anyIface |= (t == T_INTERFACE);
    }

    LbGpuHandle* h = new (std::nothrow) LbGpuHandle();
    if (!h) return fail(LBGPU_EINVAL, "out of host memory");
    h->prm = *prm;
    h->N = N;
    int rc = 0;
    auto body = [&]() -> int {
        if (prm->device >= 0) {
            if (prm->device >= ndev) return fail(LBGPU_EINVAL, "lbGpuInit: device %d of %d", prm->device, ndev);
            CU(cudaSetDevice(prm->device));
        }
        CU(cudaGetDevice(&h->device));
        CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        CU(cudaEventCreate(&h->evA));
        CU(cudaEventCreate(&h->evB));
        h->kev0.resize(LbGpuHandle::KEV); h->kev1.resize(LbGpuHandle::KEV);
        for (uint32_t k = 0; k < LbGpuHandle::KEV; ++k) { CU(cudaEventCreate(&h->kev0[k])); CU(cudaEventCreate(&h->kev1[k])); }
        h->stride = ((size_t)N + 31) / 32 * 32;
        h->blocks = (N + BLOCK - 1) / BLOCK;
        h->fs = prm->freeSurface != 0;
        h->shear = prm->nonNewtonian || prm->turbulence;
        h->force = prm->forceField && (prm->lbF[0] != 0.0 || prm->lbF[1] != 0.0 || prm->lbF[2] != 0.0);
        h->dynWall = anyDyn;
        h->macroAlways = h->fs || anyDyn || anyGas || anyIface;
        // measureUnits::setComposite (node.cpp:476-488)
        const double L = prm->unitLength, Tm = prm->unitTime, D = prm->unitDensity;
        h->uLength = L; h->uVolume = L * L * L; h->uSpeed = L / Tm; h->uAngVel = 1.0 / Tm;
        h->uForce = D * L * L * L * L / Tm / Tm; h->uTorque = D * L * L * L * L * L / Tm / Tm;

        CU(h->fA.alloc(h->stride * Q)); CU(h->fB.alloc(h->stride * Q));
        CU(h->n.alloc(N)); CU(h->ux.alloc(N)); CU(h->uy.alloc(N)); CU(h->uz.alloc(N));
        CU(h->mass.alloc(N)); CU(h->visc.alloc(N)); CU(h->shearRate.alloc(N));
        CU(h->hfx.alloc(N)); CU(h->hfy.alloc(N)); CU(h->hfz.alloc(N));
        CU(h->type0.alloc(N)); CU(h->solidIndex.alloc(N));
        if (h->fs) { CU(h->type1.alloc(N)); CU(h->mark.alloc(N)); CU(h->newMass.alloc(N)); }
        const size_t nPartial = (size_t)h->blocks * (size_t)(3 > 1 + 3 * prm->nWalls ? 3 : 1 + 3 * prm->nWalls);
        CU(h->partial.alloc(nPartial));
        CU(h->sums.alloc(8 + 3 * 64)); CU(h->scal.alloc(8)); CU(h->wallOut.alloc(3 * 64));
        CU(h->counters.alloc(8)); CU(h->status.alloc(4));
        CU(cudaMallocHost((void**)&h->pinnedStatus, 64));
        cudaStream_t s = h->stream;
        CU(cudaMemsetAsync(h->shearRate.p, 0, sizeof(double) * N, s));
        CU(cudaMemsetAsync(h->hfx.p, 0, sizeof(double) * N, s));
        CU(cudaMemsetAsync(h->hfy.p, 0, sizeof(double) * N, s));
        CU(cudaMemsetAsync(h->hfz.p, 0, sizeof(double) * N, s));
        CU(cudaMemsetAsync(h->sums.p, 0, sizeof(double) * h->sums.n, s));
        CU(cudaMemsetAsync(h->scal.p, 0, sizeof(double) * 8, s));
        CU(cudaMemsetAsync(h->counters.p, 0, sizeof(unsigned long long) * 8, s));
        CU(cudaMemsetAsync(h->status.p, 0, sizeof(uint32_t) * 4, s));
        CU(cudaMemsetAsync(h->partial.p, 0, sizeof(double) * nPartial, s));

        Dev& d = h->dev;
        memset(&d, 0, sizeof d);
        d.X = prm->size[0]; d.Y = prm->size[1]; d.Z = prm->size[2]; d.N = N; d.stride = h->stride;
        d.divX = make_div((uint32_t)d.X); d.divXY = make_div((uint32_t)d.X * (uint32_t)d.Y);
        for (int k = 0; k < 6; ++k) d.per[k] = prm->boundary[k] == T_PERIODIC;
        d.solidIndex = h->solidIndex.p;
        d.n = h->n.p; d.ux = h->ux.p; d.uy = h->uy.p; d.uz = h->uz.p;
        d.mass = h->mass.p; d.newMass = h->newMass.p; d.visc = h->visc.p; d.shearRate = h->shearRate.p;
        d.hfx = h->hfx.p; d.hfy = h->hfy.p; d.hfz = h->hfz.p;
        for (int k = 0; k < 3; ++k) { d.lbF[k] = prm->forceField ? prm->lbF[k] : 0.0; d.lbFInit[k] = prm->lbF[k]; }
        d.initVisc = prm->initDynVisc; d.plasticVisc = prm->plasticVisc; d.yieldStress = prm->yieldStress;
        d.turbConst = prm->turbConst;
        d.S1 = prm->slipCoefficient; d.S2 = 1.0 - prm->slipCoefficient;
        d.uAngVel = h->uAngVel;
        d.nonNewtonian = prm->nonNewtonian; d.turbulence = prm->turbulence;
        d.nWalls = prm->nWalls;
        d.partial = h->partial.p; d.status = h->status.p;

        // staged upload: host (pageable) -> device scratch -> SoA
        CU(cudaMemcpyAsync(h->type0.p, type_flags, N, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(h->solidIndex.p, solidIndex, sizeof(uint32_t) * N, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(h->n.p, n, sizeof(double) * N, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(h->mass.p, mass, sizeof(double) * N, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(h->visc.p, visc, sizeof(double) * N, cudaMemcpyHostToDevice, s));
        {
            DevBuf<double> tmp;
            CU(tmp.alloc((size_t)3 * N));
            CU(cudaMemcpyAsync(tmp.p, u, sizeof(double) * 3 * N, cudaMemcpyHostToDevice, s));
            k_split3<<<h->blocks, BLOCK, 0, s>>>(N, tmp.p, h->ux.p, h->uy.p, h->uz.p);
            CU(cudaStreamSynchronize(s));
        }
        Dev dd = dev_for(h, false);
        if (f) {
            // upload the populations in slices to bound the staging memory
            DevBuf<double> tmp;
            CU(tmp.alloc((size_t)Q * N));
            CU(cudaMemcpyAsync(tmp.p, f, sizeof(double) * Q * N, cudaMemcpyHostToDevice, s));
            k_upload_f<<<h->blocks, BLOCK, 0, s>>>(dd, tmp.p, h->fA.p, h->fB.p);
            CU(cudaStreamSynchronize(s));
        } else {
            k_upload_f<<<h->blocks, BLOCK, 0, s>>>(dd, nullptr, h->fA.p, h->fB.p);
        }
        h->launches += 2;
        // initial interface count (LB::redistributeMass divides by interfaceNodes.size())
        k_count<<<h->blocks, BLOCK, 0, s>>>(dd, h->counters.p + 1);
        CU(cudaMemcpyAsync(h->counters.p, h->counters.p + 2, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, s));
        ++h->launches;
        CU(cudaStreamSynchronize(s));
        CU(cudaGetLastError());
        return 0;
    };
    rc = body();
    if (rc) { lbGpuFinalize(h); return rc; }
    *out = h;
    return LBGPU_OK;
}

int lbGpuStep(LbGpuHandle* h, int doFreeSurface, int doCoupling, int rescanParticles, const LbGpuParticle* parts,
              uint32_t nParts, const LbGpuElement* elmts, uint32_t nElmts, const uint32_t* components, uint32_t nComponents) {
    if (!h) return fail(LBGPU_EINVAL, "lbGpuStep: null handle");
    if (nParts && (!parts || !elmts || !components)) return fail(LBGPU_EINVAL, "lbGpuStep: particle arrays missing");
    CU(cudaSetDevice(h->device));
    CU(cudaEventRecord(h->evA, h->stream));
    h->kevCount = 0;
    int rc;
    if (doFreeSurface && h->fs) { if ((rc = free_surface_step(h))) return rc; }
    if (doCoupling) {
        if ((rc = upload_particles(h, parts, nParts, elmts, nElmts, components, nComponents))) return rc;
        if ((rc = coupling_step(h, rescanParticles != 0))) return rc;
    }
    if ((rc = lb_step(h))) return rc;
    CU(cudaEventRecord(h->evB, h->stream));
    return LBGPU_OK;
}

int lbGpuRun(LbGpuHandle* h, int doFreeSurface, uint32_t count) {
    if (!h) return fail(LBGPU_EINVAL, "lbGpuRun: null handle");
    CU(cudaSetDevice(h->device));
    CU(cudaEventRecord(h->evA, h->stream));
    h->kevCount = 0;
    for (uint32_t k = 0; k < count; ++k) {
        int rc;
        if (doFreeSurface && h->fs) { if ((rc = free_surface_step(h))) return rc; }
        if ((rc = lb_step(h))) return rc;
    }
    CU(cudaEventRecord(h->evB, h->stream));
    return LBGPU_OK;
}

int lbGpuSynchronize(LbGpuHandle* h) {
    if (!h) return fail(LBGPU_EINVAL, "null handle");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    return check_status(h);
}

int lbGpuLastStepMs(LbGpuHandle* h, float* ms) {
    if (!h || !ms) return fail(LBGPU_EINVAL, "null argument");
    CU(cudaSetDevice(h->device));
    CU(cudaEventSynchronize(h->evB));
    CU(cudaEventElapsedTime(ms, h->evA, h->evB));
    return LBGPU_OK;
}

int lbGpuLastKernelMs(LbGpuHandle* h, float* msSum, uint32_t* launches) {
    if (!h || !msSum || !launches) return fail(LBGPU_EINVAL, "null argument");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    const uint32_t n = h->kevCount < LbGpuHandle::KEV ? h->kevCount : LbGpuHandle::KEV;
    float sum = 0.f;
    for (uint32_t k = 0; k < n; ++k) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, h->kev0[k], h->kev1[k]));
        sum += ms;
    }
    *msSum = sum; *launches = n;
    return LBGPU_OK;
}

int lbGpuLaunchCount(LbGpuHandle* h, uint64_t* launches) {
    if (!h || !launches) return fail(LBGPU_EINVAL, "null argument");
    *launches = h->launches;
    return LBGPU_OK;
}

int lbGpuParticleForces(LbGpuHandle* h, double* FHydro, double* MHydro, double* fluidVolume, double* wallFHydro) {
    if (!h) return fail(LBGPU_EINVAL, "null handle");
    CU(cudaSetDevice(h->device));
    if (int rc = check_status(h)) return rc;
    const uint32_t nE = h->nElmts;
    if (nE && (FHydro || MHydro || fluidVolume)) {
        std::vector<double> tmp((size_t)7 * nE, 0.0);
        if (h->lastStepCoupled) {
            CU(cudaMemcpyAsync(tmp.data(), h->elemOut.p, sizeof(double) * 7 * nE, cudaMemcpyDeviceToHost, h->stream));
            CU(cudaStreamSynchronize(h->stream));
        }
        for (uint32_t e = 0; e < nE; ++e) {
            for (int k = 0; k < 3; ++k) {
                if (FHydro) FHydro[3 * e + k] = tmp[(size_t)7 * e + k];
                if (MHydro) MHydro[3 * e + k] = tmp[(size_t)7 * e + 3 + k];
            }
            if (fluidVolume) fluidVolume[e] = tmp[(size_t)7 * e + 6];
        }
    }
    if (wallFHydro && h->prm.nWalls > 0) {
        const int nW = h->prm.nWalls;
        std::vector<double> tmp((size_t)3 * nW, 0.0);
        if (h->dynWall && h->steps > 0) {
            CU(cudaMemcpyAsync(tmp.data(), h->sums.p + 5, sizeof(double) * 3 * nW, cudaMemcpyDeviceToHost, h->stream));
            CU(cudaStreamSynchronize(h->stream));
        }
        for (int k = 0; k < 3 * nW; ++k) wallFHydro[k] = tmp[k] * h->uForce;  // LB.cpp:1485-1487
    }
    return LBGPU_OK;
}

int lbGpuFetchFields(LbGpuHandle* h, uint8_t* type_flags, uint32_t* solidIndex, double* n, double* u, double* mass,
                     double* visc, double* shearRate, double* hydroForce, double* f) {
    if (!h) return fail(LBGPU_EINVAL, "null handle");
    CU(cudaSetDevice(h->device));
    if (int rc = check_status(h)) return rc;
    cudaStream_t s = h->stream;
    const uint32_t N = h->N, B = h->blocks;
    Dev d = dev_for(h, false);
    if (!h->macroValid && (n || u)) {
        // n and the shifted u of the last step, recomputed from the previous population buffer
        Dev dm = d;
        dm.fsrc = h->fbuf(h->cur ^ 1);
        const bool force = h->force || h->lastStepCoupled, couple = h->lastStepCoupled;
        dm.pull = !h->lastStepFirst;
        if (couple) k_macro<true, true><<<B, BLOCK, 0, s>>>(dm);
        else if (force) k_macro<true, false><<<B, BLOCK, 0, s>>>(dm);
        else k_macro<false, false><<<B, BLOCK, 0, s>>>(dm);
        ++h->launches;
        h->macroValid = true;
    }
    if (type_flags) {
        DevBuf<uint8_t> tmp;
        CU(tmp.alloc(N));
        k_fetch_types<<<B, BLOCK, 0, s>>>(d, tmp.p);
        CU(cudaMemcpyAsync(type_flags, tmp.p, N, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
    }
    if (solidIndex) CU(cudaMemcpyAsync(solidIndex, h->solidIndex.p, sizeof(uint32_t) * N, cudaMemcpyDeviceToHost, s));
    DevBuf<double> tmp;
    if (n || u || mass || visc || shearRate || hydroForce) CU(tmp.alloc((size_t)3 * N));
    auto scalar = [&](const double* src, double* dst, int activeOnly) -> int {
        k_fetch_scalar<<<B, BLOCK, 0, s>>>(d, src, tmp.p, activeOnly);
        CU(cudaMemcpyAsync(dst, tmp.p, sizeof(double) * N, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        return 0;
    };
    int rc;
    if (n && (rc = scalar(h->n.p, n, 0))) return rc;
    if (mass && (rc = scalar(h->mass.p, mass, 0))) return rc;
    if (visc && (rc = scalar(h->visc.p, visc, 0))) return rc;
    if (shearRate && (rc = scalar(h->shearRate.p, shearRate, 1))) return rc;
    if (u) {
        k_fetch_vec<<<B, BLOCK, 0, s>>>(d, h->ux.p, h->uy.p, h->uz.p, tmp.p, 0);
        CU(cudaMemcpyAsync(u, tmp.p, sizeof(double) * 3 * N, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
    }
    if (hydroForce) {
        k_fetch_vec<<<B, BLOCK, 0, s>>>(d, h->hfx.p, h->hfy.p, h->hfz.p, tmp.p, 1);
        CU(cudaMemcpyAsync(hydroForce, tmp.p, sizeof(double) * 3 * N, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
    }
    if (f) {
        DevBuf<double> tf;
        CU(tf.alloc((size_t)Q * N));
        k_download_f<<<B, BLOCK, 0, s>>>(d, h->fbuf(h->cur), tf.p);
        CU(cudaMemcpyAsync(f, tf.p, sizeof(double) * Q * N, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
    }
    CU(cudaStreamSynchronize(s));
    CU(cudaGetLastError());
    return LBGPU_OK;
}

int lbGpuCounts(LbGpuHandle* h, uint64_t counts[4]) {
    if (!h || !counts) return fail(LBGPU_EINVAL, "null argument");
    CU(cudaSetDevice(h->device));
    Dev d = dev_for(h, false);
    CU(cudaMemsetAsync(h->counters.p + 1, 0, sizeof(unsigned long long) * 3, h->stream));
    k_count<<<h->blocks, BLOCK, 0, h->stream>>>(d, h->counters.p + 1);
    ++h->launches;
    unsigned long long tmp[3];
    CU(cudaMemcpyAsync(tmp, h->counters.p + 1, sizeof tmp, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    counts[0] = tmp[0]; counts[1] = tmp[1]; counts[2] = tmp[2]; counts[3] = h->steps;
    return LBGPU_OK;
}

int lbGpuFinalize(LbGpuHandle* h) {
    if (!h) return LBGPU_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->evA) cudaEventDestroy(h->evA);
    if (h->evB) cudaEventDestroy(h->evB);
    for (cudaEvent_t e : h->kev0) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : h->kev1) if (e) cudaEventDestroy(e);
    if (h->pinned) cudaFreeHost(h->pinned);
    if (h->pinnedStatus) cudaFreeHost(h->pinnedStatus);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return LBGPU_OK;
}

}  // extern "C"
