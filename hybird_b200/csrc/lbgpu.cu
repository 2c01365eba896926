// lbgpu.cu -- host side of liblbgpu.so: the C ABI of include/lbgpu.h on top of lb_kernels.cuh.
//
// A handle owns one or more SLABS of the lattice (cut along z, the slowest index, so halo planes
// are contiguous).  Each slab is a self-contained device lattice with one ghost plane on every cut
// side; periodic boundaries use the same ghost mechanism inside a slab (k_fill_ghosts).  The ghost
// planes between slabs of one handle are refreshed by device-to-device plane copies on the
// handle's stream; slabs of other processes (one process per GPU) are reached through NCCL
// send/recv (lbgpu_comm.cuh).
// All work of a handle is issued on its own stream; lbGpuStep is asynchronous except for the
// particle flood fill, which needs a 4-byte read-back per generation.  There is deliberately no
// CPU path: without a device every entry point returns LBGPU_ENODEVICE.
#include "../../include/lbgpu.h"

#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <memory>
#include <map>
#include <mutex>
#include <new>
#include <chrono>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "lb_kernels.cuh"
#include "lb_dem.cuh"
#include "lbgpu_comm.h"

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

#define NC(call)                                                                                         \
    do {                                                                                                 \
        int r_ = (call);                                                                                 \
        if (r_ != lbcomm::ncclSuccess) return fail(LBGPU_ECOMM, "%s failed: %s (%s:%d)", #call, lbcomm::api().GetErrorString(r_), __FILE__, __LINE__); \
    } while (0)

// LBGPU_TRACE=1: wall-clock marks of the bulk-transfer entry points on stderr (lbGpuInit, lbGpuFetchFields)
struct Trace {
    bool on;
    std::chrono::steady_clock::time_point t0, last;
    const char* what;
    explicit Trace(const char* w) : on(getenv("LBGPU_TRACE") != nullptr), what(w) { t0 = last = std::chrono::steady_clock::now(); }
    void mark(const char* label) {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[lbgpu trace] %s: %-28s %8.2f ms (total %8.2f)\n", what, label,
                std::chrono::duration<double, std::milli>(now - last).count(), std::chrono::duration<double, std::milli>(now - t0).count());
        last = now;
    }
};

// memcpy between pageable host memory and a pinned stage on several threads: one core moves 5-8 GB/s (less when the
// pages of a freshly allocated destination are touched for the first time), PCIe 5 x16 about 50 GB/s
int copy_threads() {
    static int n = 0;
    if (n == 0) {
        if (const char* e = getenv("LBGPU_COPY_THREADS")) n = atoi(e);
        if (n <= 0) { const unsigned hc = std::thread::hardware_concurrency(); n = hc >= 16 ? 8 : (hc >= 8 ? 4 : (hc >= 4 ? 2 : 1)); }
        if (n > 16) n = 16;
    }
    return n;
}
// a small pool of copy threads, created on first use and kept for the life of the process (spawning threads per 16 MB
// chunk cost as much as the copy itself)
struct CopyPool {
    std::vector<std::thread> workers;
    std::mutex m;
    std::condition_variable cvWork, cvDone;
    uint64_t generation = 0;
    int pending = 0;
    char* dst = nullptr; const char* src = nullptr; size_t bytes = 0, per = 0;
    bool stop = false;
    void start(int n) {
        for (int k = 0; k < n; ++k)
            workers.emplace_back([this, k] {
                uint64_t seen = 0;
                for (;;) {
                    std::unique_lock<std::mutex> lk(m);
                    cvWork.wait(lk, [&] { return stop || generation != seen; });
                    if (stop) return;
                    seen = generation;
                    char* d = dst; const char* s = src; const size_t n = bytes, p = per;
                    lk.unlock();
                    const size_t off = (size_t)(k + 1) * p;
                    if (off < n) memcpy(d + off, s + off, n - off < p ? n - off : p);
                    lk.lock();
                    if (--pending == 0) cvDone.notify_one();
                }
            });
    }
    void copy(void* d, const void* s, size_t n) {
        const int T = (int)workers.size() + 1;
        std::unique_lock<std::mutex> lk(m);
        dst = (char*)d; src = (const char*)s; bytes = n;
        per = ((n + T - 1) / T + 4095) / 4096 * 4096;
        pending = (int)workers.size();
        ++generation;
        lk.unlock();
        cvWork.notify_all();
        memcpy(d, s, per < n ? per : n);
        lk.lock();
        cvDone.wait(lk, [&] { return pending == 0; });
    }
    ~CopyPool() {
        { std::lock_guard<std::mutex> lk(m); stop = true; }
        cvWork.notify_all();
        for (auto& t : workers) t.join();
    }
};
void par_memcpy(void* dst, const void* src, size_t bytes) {
    const int T = copy_threads();
    if (T <= 1 || bytes < (4u << 20)) { memcpy(dst, src, bytes); return; }
    static CopyPool* pool = nullptr;  // never destroyed: worker threads must not be joined from a static destructor at exit
    static std::mutex poolMutex;
    std::lock_guard<std::mutex> lk(poolMutex);
    if (!pool) { pool = new CopyPool(); pool->start(T - 1); }
    pool->copy(dst, src, bytes);
}

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

// Device memory of closed engines is kept for the next one (per device, exact sizes of 1 MB and more, LBGPU_CACHE=0 turns it
// off): a host that builds one lattice after another of the same shape -- a parameter sweep, a restart, the bench's end-to-end
// job -- then issues no cudaMalloc at all.  A fresh cudaMalloc of a few GB normally takes ~1 ms, but was measured at 100-280 ms
// per call right after another process had released its memory.  Nothing is assumed about the contents of an allocation.
struct DevCache {
    std::mutex m;
    std::multimap<std::pair<int, size_t>, void*> blocks;
    size_t bytes = 0;
    bool on = true;
    static constexpr size_t MIN_BYTES = 1u << 20, LIMIT = 120ull << 30;
    DevCache() { if (const char* e = getenv("LBGPU_CACHE")) on = atoi(e) != 0; }
    void* take(int dev, size_t b) {
        std::lock_guard<std::mutex> g(m);
        auto it = blocks.find({ dev, b });
        if (it == blocks.end()) return nullptr;
        void* p = it->second;
        blocks.erase(it);
        bytes -= b;
        return p;
    }
    bool give(int dev, size_t b, void* p) {
        if (!on || b < MIN_BYTES) return false;
        std::lock_guard<std::mutex> g(m);
        if (bytes + b > LIMIT) return false;
        blocks.insert({ { dev, b }, p });
        bytes += b;
        return true;
    }
    void trim(int dev) {  // out of memory: hand everything cached on this device back to the driver
        std::lock_guard<std::mutex> g(m);
        for (auto it = blocks.begin(); it != blocks.end();) {
            if (it->first.first == dev) { cudaFree(it->second); bytes -= it->first.second; it = blocks.erase(it); } else ++it;
        }
    }
};
inline DevCache& dev_cache() { static DevCache* c = new DevCache(); return *c; }  // never destroyed (the driver may be gone at exit)

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    bool view = false;  // part of somebody else's allocation (one cudaMalloc costs ~1 ms whatever its size)
    int dev = -1;
    cudaError_t alloc(size_t count) {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        const size_t b = count * sizeof(T);
        cudaGetDevice(&dev);
        if (b >= DevCache::MIN_BYTES) { p = (T*)dev_cache().take(dev, b); if (p) return cudaSuccess; }
        cudaError_t e = cudaMalloc((void**)&p, b);
        if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); dev_cache().trim(dev); e = cudaMalloc((void**)&p, b); }
        if (e != cudaSuccess) { p = nullptr; n = 0; }
        return e;
    }
    void set_view(T* ptr, size_t count) { release(); p = ptr; n = count; view = true; }
    void release() {
        if (p && !view) {
            // like cudaFree, which waits for the device: work that still uses the block may be in flight on some stream
            if (dev_cache().on && n * sizeof(T) >= DevCache::MIN_BYTES) cudaDeviceSynchronize();
            if (!dev_cache().give(dev, n * sizeof(T), (void*)p)) cudaFree(p);
        }
        p = nullptr; n = 0; view = false;
    }
    ~DevBuf() { release(); }
};

using namespace lb;

// One device-resident part of the lattice: global planes [zBegin, zEnd) plus a plane below and above.
struct Slab {
    int index = 0;          // slab number in the global decomposition
    int slot = 0;           // position among the handle's slabs
    int zBegin = 1, zEnd = 1;
    uint32_t N = 0, XY = 0;
    size_t stride = 0, pad = 0;
    uint32_t blocks = 0;                 // blocks covering all N cells
    uint32_t ownBegin = 0, ownEnd = 0;   // cells of the planes the per-cell kernels cover (everything but remote ghost planes)
    Dev dev;                             // template for kernel parameters (pointers filled per launch)
    DevBuf<double> fA, fB, macroPool, n, ux, uy, uz, mass, newMass, visc, shearRate, hfx, hfy, hfz;  // n .. hfz (but newMass) are views into macroPool
    DevBuf<uint8_t> type0, type1, mark;
    DevBuf<uint32_t> solidIndex, bulk;
    DevBuf<uint32_t> curveRow;  // curved walls: row of the handle's curveDelta per cell
    // particle coupling: the flagged cells as a list (k_plist_*), {count clamped, -, count} and the per-block scan scratch
    DevBuf<uint32_t> pList, pCounts, pBlockCount;
    uint32_t pCap = 0, pGroups = 0, pScanBlocks = 0;
    DevBuf<uint32_t> staticList, staticCount;  // PART 2 of the split step kernel: owned non-bulk cells that can be active
    uint32_t nStatic = 0;
    DevBuf<uint32_t> wallCells, wallMasks;     // wall-push list (k_wall_push): bulk cells with links to static no-slip walls
    uint32_t nWallPush = 0;
    // free surface: the interface-cell list and the list of tiles the step kernel visits (lb_kernels.cuh, k_list_*);
    // listCounts = {interface cells, tiles, interface cells before clamping to the capacity}
    DevBuf<uint8_t> tileFlags;
    DevBuf<uint32_t> cellList, tileList, listCounts, listBlockCount, listTileOffset, candList, candBlockCount;
    uint32_t candCap = 0;
    uint32_t listGrid = 0, listBlocks = 0, cellCap = 0;  // blocks of a list-driven launch; blocks of the list passes
    DevBuf<uint32_t> rangeCells, listRanges;  // face / interior launches of a slab between processes: cell ranges, their list positions
    bool hasRanges = false;
    DevBuf<uint32_t> gDst, gSrc, gPop;   // local ghost list (periodic mirrors)
    uint32_t nGhost = 0;
    DevBuf<double> partial, sums, scal, elemOut, elemPartial, summary;
    DevBuf<uint32_t> elemDone;
    DevBuf<unsigned long long> counters;  // [0] nInterface [1..3] k_count scratch
    DevBuf<uint32_t> status;              // [0] type error, [1] flood-fill counter
    // host copies of what lbGpuInit received for the ghost cells (they are dead cells of the reference)
    std::vector<uint32_t> ghostIdx, ghostSrc, ghostSolid;
    std::vector<uint8_t> ghostType;
    std::vector<double> ghostN, ghostU;  // density / velocity of the dead cells that carry a wall node (cell 0 at most)
    // z-periodic lattice cut into several slabs: the reference's shell planes z=0 / z=Z-1 are replaced by remote ghost
    // planes on the device; what lbGpuInit received for them is kept for lbGpuFetchFields
    std::vector<uint8_t> shellTypeLo, shellTypeHi;
    std::vector<uint32_t> shellSolidLo, shellSolidHi;
    bool remoteLo = false, remoteHi = false;  // the z-/z+ ghost plane belongs to another slab
    double* fbuf(int k) { return (k == 0 ? fA.p : fB.p) + pad; }
    uint8_t* tbuf(int k) { return k == 0 ? type0.p : type1.p; }
};

// ---------------------------------------------------------------------------------------------
// Peer halo: the per-step halo between ranks without NCCL.  Every rank exports the arrays of its two edge slabs
// (cudaIpcGetMemHandle), maps the ones of its neighbour ranks (NVLink / NVSwitch peers) and, once the face planes of a
// step are updated, a put kernel on the high-priority comm stream stores them straight into the neighbours' ghost
// planes, then raises a sequence flag in the neighbour's memory; the neighbour's stream waits on its own flag word
// before the next kernel that reads the ghosts.  One launch per step instead of a grouped send/recv per plane, no
// proxy thread, no rendezvous: the sender knows the destination is free because it has itself received the
// neighbour's flag of the previous step, which the neighbour raises after its last read of that ghost plane.
// ---------------------------------------------------------------------------------------------
enum PeerBuf { PB_FA, PB_FB, PB_N, PB_UX, PB_UY, PB_UZ, PB_VISC, PB_HFX, PB_HFY, PB_HFZ, PB_TYPE, PB_MARK, PB_SOLID, PB_MASS, PB_FLAGS, PB_COUNT };
struct PeerExport {  // what a rank tells the others about one of its edge slabs (side 0: first slab, 1: last slab)
    cudaIpcMemHandle_t handle[PB_COUNT];
    unsigned long long offset[PB_COUNT];
    uint32_t valid[PB_COUNT];
    unsigned long long stride, pad;
    uint32_t Z, XY;
};
constexpr int PUT_MAX = 64;
struct PutPlan {  // one launch of k_put_halo: plane copies into peer memory + the flags to raise afterwards
    int n;
    const char* src[PUT_MAX];
    char* dst[PUT_MAX];
    uint32_t bytes[PUT_MAX];
    uint32_t* flagDst[2];
    uint32_t* counter;
    uint32_t chunk;
    // in-cycle exchanges (k_xput): the receiver must have reached the exchange before its ghost planes are overwritten
    uint32_t* readyDst[2];          // where we tell each neighbour "my stream is here" ...
    const uint32_t* readyLocal[2];  // ... and where they tell us
};
struct PeerHalo {
    bool on = false;
    std::string note;                    // why it is off
    char* nb[2][PB_COUNT] = {};          // [0] neighbour below (its LAST slab), [1] neighbour above (its FIRST slab)
    PeerExport nbInfo[2];
    std::vector<std::pair<cudaIpcMemHandle_t, void*>> opened;
    // flag words: [0] step halo landed from the rank below, [1] from the rank above; in-cycle exchanges: [2]/[3] the rank
    // below / above has reached exchange number x ("ready"), [4]/[5] its planes of exchange x have landed ("done");
    // [8], [9] block counters of the two put kernels
    DevBuf<uint32_t> flags;
    uint32_t seq = 0, xseq = 0;
    PutPlan plan[2];
    uint32_t planWhat[2] = { 0, 0 };
    std::vector<std::pair<uint64_t, PutPlan>> xplans;  // in-cycle exchanges, by (fields, population buffer)
    bool inCycle = true;                 // in-cycle exchanges through peer memory too (LBGPU_PEER_CYCLE=0: NCCL)
};

__global__ void __launch_bounds__(128) k_put_halo(const __grid_constant__ PutPlan pl, uint32_t seq) {
    const int e = blockIdx.y;
    const size_t begin = (size_t)blockIdx.x * pl.chunk;
    const uint32_t bytes = pl.bytes[e];
    if (begin < bytes) {
        const size_t len = bytes - begin < pl.chunk ? bytes - begin : pl.chunk;
        const char* s = pl.src[e] + begin;
        char* d = pl.dst[e] + begin;
        if ((((uintptr_t)s | (uintptr_t)d | len) & 15) == 0) {
            const uint4* s4 = (const uint4*)s; uint4* d4 = (uint4*)d;
            for (size_t k = threadIdx.x; k < len / 16; k += blockDim.x) d4[k] = s4[k];
        } else if ((((uintptr_t)s | (uintptr_t)d | len) & 7) == 0) {
            const unsigned long long* s8 = (const unsigned long long*)s; unsigned long long* d8 = (unsigned long long*)d;
            for (size_t k = threadIdx.x; k < len / 8; k += blockDim.x) d8[k] = s8[k];
        } else {
            for (size_t k = threadIdx.x; k < len; k += blockDim.x) d[k] = s[k];
        }
    }
    // every block's stores are ordered before its count; the last block to count raises the flags
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t total = gridDim.x * gridDim.y;
        if (atomicAdd(pl.counter, 1u) == total - 1) {
            *pl.counter = 0;
            __threadfence_system();
            for (int k = 0; k < 2; ++k)
                if (pl.flagDst[k]) *(volatile uint32_t*)pl.flagDst[k] = seq;
        }
    }
}

// the stream stalls until the neighbours' flags reach `seq` (their planes of this step have landed in our ghosts)
__global__ void k_wait_halo(volatile uint32_t* flags, uint32_t seq, int needDown, int needUp, uint32_t* status) {
    const int k = threadIdx.x;
    if (k > 1 || !(k == 0 ? needDown : needUp)) return;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while ((int)(flags[k] - seq) < 0) {
        __nanosleep(200);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 20000000000ull) { atomicExch(&status[3], 1u + (uint32_t)k); return; }  // 20 s: the neighbour is gone
    }
    __threadfence_system();
}

// In-cycle exchange through peer memory (type / mark / mass / ... planes between the passes of the free-surface update
// and of the particle-flag flood fill; the same planes ncclSend/ncclRecv would carry, in one launch).  Unlike the step
// halo these arrays are not double-buffered, so the receiver's earlier kernels may still be reading the ghost planes:
// block 0 first tells both neighbours that this rank's stream has reached exchange `seq` (every earlier kernel of the
// stream has completed), every block then waits until the neighbours have said the same, copies its chunk, and the
// last block raises the neighbours' "done" words; k_wait_halo on the done words completes the exchange.
__global__ void __launch_bounds__(128) k_xput(const __grid_constant__ PutPlan pl, uint32_t seq, uint32_t* status) {
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < 2 && pl.readyDst[threadIdx.x]) {
        *(volatile uint32_t*)pl.readyDst[threadIdx.x] = seq;
        __threadfence_system();
    }
    if (threadIdx.x < 2 && pl.readyLocal[threadIdx.x]) {
        const volatile uint32_t* f = pl.readyLocal[threadIdx.x];
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while ((int)(*f - seq) < 0) {
            __nanosleep(100);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 20000000000ull) { atomicExch(&status[3], 3u); break; }
        }
    }
    __syncthreads();
    const int e = blockIdx.y;
    const size_t begin = (size_t)blockIdx.x * pl.chunk;
    const uint32_t bytes = pl.bytes[e];
    if (begin < bytes) {
        const size_t len = bytes - begin < pl.chunk ? bytes - begin : pl.chunk;
        const char* s = pl.src[e] + begin;
        char* d = pl.dst[e] + begin;
        if ((((uintptr_t)s | (uintptr_t)d | len) & 15) == 0) {
            const uint4* s4 = (const uint4*)s; uint4* d4 = (uint4*)d;
            for (size_t k = threadIdx.x; k < len / 16; k += blockDim.x) d4[k] = s4[k];
        } else if ((((uintptr_t)s | (uintptr_t)d | len) & 3) == 0) {
            const uint32_t* s1 = (const uint32_t*)s; uint32_t* d1 = (uint32_t*)d;
            for (size_t k = threadIdx.x; k < len / 4; k += blockDim.x) d1[k] = s1[k];
        } else {
            for (size_t k = threadIdx.x; k < len; k += blockDim.x) d[k] = s[k];
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t total = gridDim.x * gridDim.y;
        if (atomicAdd(pl.counter, 1u) == total - 1) {
            *pl.counter = 0;
            __threadfence_system();
            for (int k = 0; k < 2; ++k)
                if (pl.flagDst[k]) *(volatile uint32_t*)pl.flagDst[k] = seq;
        }
    }
}

}  // namespace

struct LbGpuHandle {
    LbGpuParams prm;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t evA = nullptr, evB = nullptr;
    cudaStream_t commStream = nullptr;  // halo transport to other processes, overlapped with the interior update
    cudaEvent_t evFaces = nullptr, evHalo = nullptr;
    PeerHalo peer;                      // neighbour ranks' arrays mapped into this process (see "peer halo")
    std::vector<std::unique_ptr<Slab>> slabs;
    int nSlabsGlobal = 1, firstSlab = 0;
    DevBuf<lb::RawParticle> rawParts;
    DevBuf<lb::RawElement> rawElmts;
    DevBuf<lb::Particle> parts;
    DevBuf<lb::Element> elmts;
    DevBuf<uint32_t> comps;
    void* pinned = nullptr;
    size_t pinnedBytes = 0;
    // bulk transfers between the caller's pageable arrays and the device go through two pinned stages (the CPU copy of
    // one chunk overlaps the DMA of the other; a plain cudaMemcpy from pageable memory serialises the two)
    char* stage[2] = { nullptr, nullptr };
    cudaEvent_t stageEv[2] = { nullptr, nullptr };
    static constexpr size_t STAGE_MAX = 16u << 20;
    static size_t stage_bytes() {  // LBGPU_STAGE_MB: 1..16 (pinning costs ~1 ms per MB: two 16 MB stages were 30 ms of the first upload)
        static size_t v = 0;
        if (!v) { int mb = 4; if (const char* e = getenv("LBGPU_STAGE_MB")) mb = atoi(e); if (mb < 1) mb = 1; if (mb > 16) mb = 16; v = (size_t)mb << 20; }
        return v;
    }
    uint32_t* pinnedStatus = nullptr;
    uint32_t nParts = 0, nElmts = 0, nComps = 0;
    int cur = 0;      // population buffer holding the latest post-collision state (0 = A)
    bool fs = false, shear = false, force = false, macroAlways = false, dynWall = false, slip = false;
    bool macroValid = true, lastStepFirst = false, lastStepCoupled = false, typesFlipped = false, listsFresh = false;
    uint32_t* pinnedCounts = nullptr;  // per slab: {interface cells, visited tiles, unclamped interface cells, -} of the last list build
    uint64_t steps = 0, launches = 0;
    // curved walls (type 9) and LB::enforceMassConservation (problemName DRUM)
    // dense step launches: L2 prefetch distance in blocks (Dev::prefetch).  Default: one generation of resident blocks
    // (5 per SM); LBGPU_PREFETCH overrides (0 = off).  A/B on the 256^3 channel: 0 -> 0.872-0.900 of the copy peak,
    // 370 / 740 -> 0.907-0.951, 1480 -> 0.81, 2960+ -> 0.68 (the prefetched rows no longer survive in L2).
    uint32_t prefetchBlocks = 0xffffffffu, prefetchTiles = 0;  // (tile lists: one request per population for four consecutive full tiles)
    bool wallPushAllowed = true;  // LBGPU_WALL_PUSH=0 keeps the list-driven launch for every wall-adjacent cell (A/B)
    bool ghostCopy = false;  // with wallPush, single process: the periodic mirrors are written by k_fill_ghosts after the step
                             // instead of by the step kernel, so the cells next to periodic faces take the bulk path too
    bool wallPush = false;  // static no-slip links are served by slots pre-stored in the wall cells (k_wall_push)
    int fsGridPerSM = 0;  // > 0: persistent launch of the free-surface step kernel, this many blocks per SM (LBGPU_FS_GRID)
    bool hasCurved = false, curvesSet = false, enforceMass = false;
    double totalMass = 0.0;
    DevBuf<double> curveDelta;
    double uLength = 1, uSpeed = 1, uAngVel = 1, uForce = 1, uTorque = 1, uVolume = 1;
    // CUDA-event pairs around the fused step kernel of the last lbGpuStep/lbGpuRun call (ring of KEV)
    static constexpr uint32_t KEV = 512;
    std::vector<cudaEvent_t> kev0, kev1;
    uint32_t kevCount = 0;
    int numSMs = 148;
    // CUDA graph of two consecutive cycles of a free-surface lattice without particles (lbGpuRun): ~20 launches per cycle
    // replayed as one graph launch.  Two cycles, because the population buffers alternate.  LBGPU_GRAPH=0 turns it off.
    struct CycleGraph {
        cudaGraphExec_t exec = nullptr;
        std::vector<uint32_t> tiles;   // visited tiles per slab the step kernels' grids were sized for
        uint64_t launches = 0;         // kernels per replay
        int cur = -1, fs = -1, kind = -1;
        uint32_t nParts = 0;
        uint64_t replays = 0, captures = 0;
    } graph;
    bool graphAllowed = true, capturing = false;
    bool forcesPending = false, lazyForces = false;  // several processes: the element sums of the last step are still per rank / lbGpuRun defers them
    int floodGens = 3;       // single process: flood-fill generations issued per coupling step without a host round trip (0: ask the host)
    uint32_t eagerCycles = 0;  // cycles launched one by one since the particle lists / settings last changed (a graph needs settled buffers)
    // phase trace (lbGpuPhaseTrace): CUDA events at the phase boundaries of a cycle, a ring of the last PH_RING cycles
    static constexpr int PH_MARKS = 8, PH_RING = 64;
    bool phaseOn = false;
    std::vector<cudaEvent_t> phaseEv;  // PH_RING x PH_MARKS
    uint64_t phaseCycles = 0;
    // DEM sub-steps on the device (lb_dem.cuh): the elements' Gear state, partner lists, wall table
    struct Dem {
        bool on = false;
        lbdem::Params prm;
        uint32_t n = 0, nP = 0, nWalls = 0;  // elements, their particles (spheres), walls
        DevBuf<lbdem::Elmt> e;
        DevBuf<lbdem::Part> pt;
        // periodic DEM boundaries (single spheres): ghost particles behind the standard ones in pt, rebuilt with the tables
        bool pbc = false;
        uint32_t cap = 0;  // particle slots: nP, or 7 per sphere with periodic boundaries (3 ghosts + 3 corner ghosts at most)
        DevBuf<int8_t> gflag;
        DevBuf<uint32_t> basePos, cornPos, nComp, cnt;  // cnt[0] = particles + ghosts, cnt[1] = a rebuild happened since the last coupling step
        bool rescan = false;
        DevBuf<lbdem::Wall> walls;
        DevBuf<uint32_t> nbr, nNbr, flag;  // flag[0] = rebuild in this sub-step, [1] = longest partner list, [2] = rebuilds so far
        // uniform grid for the table rebuild of large beds (k_grid_*): LBGPU_DEM_GRID=0/1 overrides the choice by size
        bool grid = false;
        DevBuf<lbdem::Grid> gridDesc;
        DevBuf<uint32_t> cellCount, cellFill, cellOf, sorted;
        DevBuf<double> scal, hydro;         // scal[0] = DEM::maxDisp; hydro: forces handed in from the host (lbGpuDemStep)
    } dem;
};

namespace {

typedef void (*StepKernel)(const Dev);

// phase boundaries of a cycle: 0 start, 1 after the DEM sub-steps, 2 after the list build, 3 after the free-surface update,
// 4 after the coupling step, 5 after the step kernels + halo, 6 after wall slots / moving-wall sums, 7 after the element forces
// and the type sync (end of the cycle)
inline void phase_mark(LbGpuHandle* h, int m) {
    if (!h->phaseOn || h->capturing) return;
    cudaEventRecord(h->phaseEv[(size_t)(h->phaseCycles % LbGpuHandle::PH_RING) * LbGpuHandle::PH_MARKS + m], h->stream);
    if (m == LbGpuHandle::PH_MARKS - 1) ++h->phaseCycles;
}

// FS / DYNWALL / PART are compile-time in k_step; only the combinations that can occur are instantiated.
// part 0: the whole step in one launch (lean variants); parts 1 + 2: bulk cells, then the rest (lb_kernels.cuh).
template <bool FORCE, bool SHEAR, bool MACRO, bool COUPLE>
StepKernel pick_step(bool fsOn, bool dyn, int part) {
    if constexpr (!MACRO) {
        if (part == 1) return k_step<FORCE, SHEAR, false, COUPLE, false, false, 1>;
        if (part == 2) return k_step<FORCE, SHEAR, false, COUPLE, false, false, 2>;
        return k_step<FORCE, SHEAR, false, COUPLE, false, false, 0>;
    } else {
        // moving-wall sums only arise next to walls, i.e. never on cells with the static bulk bit
        if (part == 1) return fsOn ? k_step<FORCE, SHEAR, true, COUPLE, true, false, 1> : k_step<FORCE, SHEAR, true, COUPLE, false, false, 1>;
        if (part == 3) return k_step<FORCE, SHEAR, true, COUPLE, true, false, 3>;
        if (fsOn) return dyn ? k_step<FORCE, SHEAR, true, COUPLE, true, true, 2> : k_step<FORCE, SHEAR, true, COUPLE, true, false, 2>;
        return dyn ? k_step<FORCE, SHEAR, true, COUPLE, false, true, 2> : k_step<FORCE, SHEAR, true, COUPLE, false, false, 2>;
    }
}

// the variants with a heavy generic path are split into a bulk launch and a launch for the other cells
bool step_is_split(bool shear, bool macro, bool couple, bool fsOn, bool dyn) { return macro || shear || couple || fsOn || dyn; }

StepKernel select_step(bool force, bool shear, bool macro, bool couple, bool fsOn, bool dyn, int part) {
    if (!macro) {
        if (couple) return shear ? pick_step<true, true, false, true>(false, false, part) : pick_step<true, false, false, true>(false, false, part);
        if (force) return shear ? pick_step<true, true, false, false>(false, false, part) : pick_step<true, false, false, false>(false, false, part);
        return shear ? pick_step<false, true, false, false>(false, false, part) : pick_step<false, false, false, false>(false, false, part);
    }
    // the full variants always carry the force path (exact when the force is zero)
    if (couple) return shear ? pick_step<true, true, true, true>(fsOn, dyn, part) : pick_step<true, false, true, true>(fsOn, dyn, part);
    return shear ? pick_step<true, true, true, false>(fsOn, dyn, part) : pick_step<true, false, true, false>(fsOn, dyn, part);
}

// source populations of a launch: buffer `buf`, pulled through the links (pull) or taken in place
void set_src(Dev& d, Slab* s, int buf, bool pull) {
    d.fsrc = s->fbuf(buf);
    d.pull = pull ? 1 : 0;
    for (int k = 0; k < Q; ++k) {
        d.fsrcK[k] = d.fsrc + (size_t)k * s->stride;
        d.fsrcP[k] = d.fsrcK[k] - (pull ? d.off[k] : 0);
    }
}

Dev dev_for(LbGpuHandle* h, Slab* s) {
    Dev d = s->dev;
    set_src(d, s, h->cur, h->steps > 0);
    d.fdst = s->fbuf(h->cur ^ 1);
    for (int k = 0; k < Q; ++k) d.fdstK[k] = d.fdst + (size_t)k * s->stride;
    d.type = s->tbuf(0);     // current types (updated in place by the free-surface step)
    d.typeOld = s->tbuf(0);  // the step kernel of a cycle with a free-surface step looks its links up in tbuf(1) instead
    d.list = s->cellList.p; d.nList = s->listCounts.p;  // list-driven kernels: the interface cells unless told otherwise
    d.cand = s->candList.p; d.nCand = s->listCounts.p + 3;
    d.lazyMass = 0;
    if (h->ghostCopy) d.push = 0;
    d.curveRow = s->curveRow.p; d.curveDelta = h->curveDelta.p; d.shearState = h->shear ? 1 : 0;
    d.prefetch = h->prefetchBlocks; d.prefetchTiles = h->prefetchTiles;
    d.parts = h->parts.p; d.elmts = h->elmts.p; d.comps = h->comps.p;
    d.nParts = h->nParts; d.nElmts = h->nElmts;
    d.cellBegin = s->ownBegin; d.cellEnd = s->ownEnd;
    return d;
}
Dev dev_all(LbGpuHandle* h, Slab* s) {  // same, covering every cell including remote ghost planes
    Dev d = dev_for(h, s);
    d.cellBegin = 0; d.cellEnd = s->N;
    return d;
}
uint32_t own_blocks(const Slab* s) { return (s->ownEnd - s->ownBegin + BLOCK - 1) / BLOCK; }

// The pinned stages are pooled over the life of the process: pinning 2 x 16 MB costs several milliseconds, which a host that
// creates one engine after another (a parameter sweep, the bench's end-to-end job) would pay every time.
struct StagePool {
    std::mutex m;
    std::vector<char*> free_;
    char* get() {
        { std::lock_guard<std::mutex> g(m); if (!free_.empty()) { char* p = free_.back(); free_.pop_back(); return p; } }
        char* p = nullptr;
        if (cudaHostAlloc((void**)&p, LbGpuHandle::stage_bytes(), getenv("LBGPU_STAGE_PLAIN") ? cudaHostAllocDefault : cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        return p;
    }
    void put(char* p) { std::lock_guard<std::mutex> g(m); free_.push_back(p); }
};
StagePool& stage_pool() { static StagePool* p = new StagePool(); return *p; }  // never destroyed (the driver may be gone at exit)

int ensure_stages(LbGpuHandle* h) {
    for (int k = 0; k < 2; ++k) {
        if (!h->stage[k]) { h->stage[k] = stage_pool().get(); if (!h->stage[k]) return fail(LBGPU_ECUDA, "pinned staging buffer: out of memory"); }
        if (!h->stageEv[k]) CU(cudaEventCreateWithFlags(&h->stageEv[k], cudaEventDisableTiming));
    }
    return 0;
}
// pageable host -> device, chunked through the two pinned stages; asynchronous on the handle's stream for the caller
// (the source may be reused on return)
int h2d_staged(LbGpuHandle* h, void* dst, const void* src, size_t bytes) {
    if (bytes < LbGpuHandle::stage_bytes()) {  // small lattices: not worth pinning two stages
        CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
        return 0;
    }
    if (int rc = ensure_stages(h)) return rc;
    size_t off = 0;
    for (int k = 0; off < bytes; ++k, off += LbGpuHandle::stage_bytes()) {
        const size_t len = bytes - off < LbGpuHandle::stage_bytes() ? bytes - off : LbGpuHandle::stage_bytes();
        const int b = k & 1;
        if (k >= 2) CU(cudaEventSynchronize(h->stageEv[b]));
        par_memcpy(h->stage[b], (const char*)src + off, len);
        CU(cudaMemcpyAsync((char*)dst + off, h->stage[b], len, cudaMemcpyHostToDevice, h->stream));
        CU(cudaEventRecord(h->stageEv[b], h->stream));
    }
    CU(cudaEventSynchronize(h->stageEv[0]));
    CU(cudaEventSynchronize(h->stageEv[1]));
    return 0;
}
// device -> pageable host, the same way; complete on return
int d2h_staged(LbGpuHandle* h, void* dst, const void* src, size_t bytes) {
    if (bytes < LbGpuHandle::stage_bytes()) {
        CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        return 0;
    }
    if (int rc = ensure_stages(h)) return rc;
    size_t off = 0, done = 0;
    int k = 0;
    for (; off < bytes; ++k, off += LbGpuHandle::stage_bytes()) {
        const size_t len = bytes - off < LbGpuHandle::stage_bytes() ? bytes - off : LbGpuHandle::stage_bytes();
        const int b = k & 1;
        if (k >= 2) {  // the chunk that used this stage two rounds ago has landed: hand it to the caller
            CU(cudaEventSynchronize(h->stageEv[b]));
            par_memcpy((char*)dst + done, h->stage[b], LbGpuHandle::stage_bytes());
            done += LbGpuHandle::stage_bytes();
        }
        CU(cudaMemcpyAsync(h->stage[b], (const char*)src + off, len, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaEventRecord(h->stageEv[b], h->stream));
    }
    for (int r = (k >= 2 ? k - 2 : 0); r < k; ++r) {
        const int b = r & 1;
        const size_t len = bytes - done < LbGpuHandle::stage_bytes() ? bytes - done : LbGpuHandle::stage_bytes();
        CU(cudaEventSynchronize(h->stageEv[b]));
        par_memcpy((char*)dst + done, h->stage[b], len);
        done += len;
    }
    return 0;
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

// ---------------------------------------------------------------------------------------------
// Ghost refresh: local periodic mirrors inside every slab, then the planes between slabs.
// `what` is a mask of lb::G_* fields; G_POPS refers to the destination population buffer.
// ---------------------------------------------------------------------------------------------
const int POPS_UP[5] = { 5, 11, 13, 15, 18 };    // CZ > 0: pulled by the slab above out of its lower ghost plane
const int POPS_DOWN[5] = { 6, 12, 14, 16, 17 };  // CZ < 0

template <class T>
int copy_plane(T* dst, uint32_t dstPlane, const T* src, uint32_t srcPlane, uint32_t XY, cudaStream_t st) {
    CU(cudaMemcpyAsync(dst + (size_t)dstPlane * XY, src + (size_t)srcPlane * XY, sizeof(T) * XY, cudaMemcpyDeviceToDevice, st));
    return 0;
}

// plane `sp` of slab a -> ghost plane `dp` of slab b (same process)
int copy_face(LbGpuHandle* h, Slab* a, uint32_t sp, Slab* b, uint32_t dp, uint32_t what, bool up) {
    cudaStream_t st = h->stream;
    const uint32_t XY = a->XY;
    int rc;
    if (what & G_POPS) {
        double* src = a->fbuf(h->cur ^ 1); double* dst = b->fbuf(h->cur ^ 1);
        if (h->slip) {
            for (int k = 0; k < Q; ++k) if ((rc = copy_plane(dst + (size_t)k * b->stride, dp, src + (size_t)k * a->stride, sp, XY, st))) return rc;
        } else {
            const int* ks = up ? POPS_UP : POPS_DOWN;
            for (int q = 0; q < 5; ++q)
                if ((rc = copy_plane(dst + (size_t)ks[q] * b->stride, dp, src + (size_t)ks[q] * a->stride, sp, XY, st))) return rc;
        }
    }
    if (what & G_POPS_SRC) {
        double* src = a->fbuf(h->cur); double* dst = b->fbuf(h->cur);
        for (int k = 0; k < Q; ++k) if ((rc = copy_plane(dst + (size_t)k * b->stride, dp, src + (size_t)k * a->stride, sp, XY, st))) return rc;
    }
    if (what & G_TYPE) if ((rc = copy_plane(b->tbuf(0), dp, a->tbuf(0), sp, XY, st))) return rc;
    if (what & G_SOLID) if ((rc = copy_plane(b->solidIndex.p, dp, a->solidIndex.p, sp, XY, st))) return rc;
    if (what & G_MASS) if ((rc = copy_plane(b->mass.p, dp, a->mass.p, sp, XY, st))) return rc;
    if (what & G_MACRO) {
        if ((rc = copy_plane(b->n.p, dp, a->n.p, sp, XY, st))) return rc;
        if ((rc = copy_plane(b->ux.p, dp, a->ux.p, sp, XY, st))) return rc;
        if ((rc = copy_plane(b->uy.p, dp, a->uy.p, sp, XY, st))) return rc;
        if ((rc = copy_plane(b->uz.p, dp, a->uz.p, sp, XY, st))) return rc;
    }
    if (what & G_VISC) if ((rc = copy_plane(b->visc.p, dp, a->visc.p, sp, XY, st))) return rc;
    if (what & G_HF) {
        if ((rc = copy_plane(b->hfx.p, dp, a->hfx.p, sp, XY, st))) return rc;
        if ((rc = copy_plane(b->hfy.p, dp, a->hfy.p, sp, XY, st))) return rc;
        if ((rc = copy_plane(b->hfz.p, dp, a->hfz.p, sp, XY, st))) return rc;
    }
    if ((what & G_MARK) && a->mark.p) if ((rc = copy_plane(b->mark.p, dp, a->mark.p, sp, XY, st))) return rc;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Slabs of other processes: NCCL send/recv of the face planes (every field plane is contiguous, so the planes are
// sent straight out of / received straight into the SoA arrays, no packing), one group per exchange.
// ---------------------------------------------------------------------------------------------
struct Xfer { void* ptr; size_t bytes; };

// field planes exchanged with the slab above (up) or below: what this slab sends and where it receives
void face_planes(LbGpuHandle* h, Slab* s, uint32_t what, bool up, std::vector<Xfer>& snd, std::vector<Xfer>& rcv) {
    const size_t XY = s->XY;
    const size_t sp = up ? (size_t)s->dev.Z - 2 : 1, dp = up ? (size_t)s->dev.Z - 1 : 0;
    auto add2 = [&](char* sendBase, char* recvBase, size_t elem) {
        snd.push_back({ sendBase + sp * XY * elem, XY * elem });
        rcv.push_back({ recvBase + dp * XY * elem, XY * elem });
    };
    auto add = [&](void* base, size_t elem) { add2((char*)base, (char*)base, elem); };
    if (what & G_POPS) {
        double* f = s->fbuf(h->cur ^ 1);
        if (h->slip) { for (int k = 0; k < Q; ++k) add(f + (size_t)k * s->stride, 8); }
        else {
            // the populations moving towards the neighbour leave, the ones moving away from it arrive
            const int* out = up ? POPS_UP : POPS_DOWN; const int* in = up ? POPS_DOWN : POPS_UP;
            for (int q = 0; q < 5; ++q) add2((char*)(f + (size_t)out[q] * s->stride), (char*)(f + (size_t)in[q] * s->stride), 8);
        }
    }
    if (what & G_POPS_SRC) { double* f = s->fbuf(h->cur); for (int k = 0; k < Q; ++k) add(f + (size_t)k * s->stride, 8); }
    if (what & G_TYPE) add(s->tbuf(0), 1);
    if (what & G_SOLID) add(s->solidIndex.p, 4);
    if (what & G_MASS) add(s->mass.p, 8);
    if (what & G_MACRO) { add(s->n.p, 8); add(s->ux.p, 8); add(s->uy.p, 8); add(s->uz.p, 8); }
    if (what & G_VISC) add(s->visc.p, 8);
    if (what & G_HF) { add(s->hfx.p, 8); add(s->hfy.p, 8); add(s->hfz.p, 8); }
    if ((what & G_MARK) && s->mark.p) add(s->mark.p, 1);
}

// ranks holding the slab below the handle's first / above its last slab (-1: a true lattice boundary)
void neighbour_ranks(LbGpuHandle* h, int* down, int* up) {
    const lbcomm::Comm& c = lbcomm::comm();
    const bool ring = h->prm.boundary[4] == T_PERIODIC;
    *down = c.rank > 0 ? c.rank - 1 : (ring ? c.world - 1 : -1);
    *up = c.rank + 1 < c.world ? c.rank + 1 : (ring ? 0 : -1);
}

int exchange_remote(LbGpuHandle* h, uint32_t what, cudaStream_t st) {
    lbcomm::Api& A = lbcomm::api();
    const lbcomm::Comm& c = lbcomm::comm();
    int down, up;
    neighbour_ranks(h, &down, &up);
    std::vector<Xfer> sUp, rUp, sDn, rDn;
    if (up >= 0) face_planes(h, h->slabs.back().get(), what, true, sUp, rUp);
    if (down >= 0) face_planes(h, h->slabs.front().get(), what, false, sDn, rDn);
    // posting order send-up, send-down, recv-down, recv-up pairs every send with its receive even when both
    // neighbours are the same rank (two ranks on a periodic axis)
    NC(A.GroupStart());
    for (const Xfer& x : sUp) NC(A.Send(x.ptr, x.bytes, lbcomm::ncclUint8, up, c.comm, st));
    for (const Xfer& x : sDn) NC(A.Send(x.ptr, x.bytes, lbcomm::ncclUint8, down, c.comm, st));
    for (const Xfer& x : rDn) NC(A.Recv(x.ptr, x.bytes, lbcomm::ncclUint8, down, c.comm, st));
    for (const Xfer& x : rUp) NC(A.Recv(x.ptr, x.bytes, lbcomm::ncclUint8, up, c.comm, st));
    NC(A.GroupEnd());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Peer halo set-up (collective over the communicator): export the edge slabs' arrays, gather every rank's records,
// map the neighbours'.  Any failure on any rank (no peer access, IPC refused) switches every rank back to NCCL
// send/recv for the step halo.  LBGPU_PEER_HALO=0 does the same (A/B).
// ---------------------------------------------------------------------------------------------
void* peer_buf_ptr(Slab* s, int b, LbGpuHandle* h) {
    switch (b) {
        case PB_FA: return s->fA.p; case PB_FB: return s->fB.p; case PB_N: return s->n.p;
        case PB_UX: return s->ux.p; case PB_UY: return s->uy.p; case PB_UZ: return s->uz.p;
        case PB_VISC: return s->visc.p; case PB_HFX: return s->hfx.p; case PB_HFY: return s->hfy.p; case PB_HFZ: return s->hfz.p;
        case PB_TYPE: return s->type0.p; case PB_MARK: return s->mark.p; case PB_SOLID: return s->solidIndex.p; case PB_MASS: return s->mass.p;
        case PB_FLAGS: return h->peer.flags.p;
    }
    return nullptr;
}

int peer_setup(LbGpuHandle* h) {
    PeerHalo& P = h->peer;
    P.on = false;
    if (!lbcomm::active()) return 0;
    const lbcomm::Comm& c = lbcomm::comm();
    lbcomm::Api& A = lbcomm::api();
    cudaStream_t st = h->stream;
    int down, up;
    neighbour_ranks(h, &down, &up);
    CU(P.flags.alloc(64));
    CU(cudaMemsetAsync(P.flags.p, 0, 64 * sizeof(uint32_t), st));
    std::string why;
    if (const char* e = getenv("LBGPU_PEER_HALO")) { if (atoi(e) == 0) why = "LBGPU_PEER_HALO=0"; }
    if (!A.AllGather) why = "ncclAllGather missing";
    // export
    PeerExport mine[2];
    memset(mine, 0, sizeof mine);
    for (int side = 0; side < 2 && why.empty(); ++side) {
        Slab* s = side == 0 ? h->slabs.front().get() : h->slabs.back().get();
        mine[side].stride = s->stride; mine[side].pad = s->pad; mine[side].Z = (uint32_t)s->dev.Z; mine[side].XY = s->XY;
        for (int b = 0; b < PB_COUNT; ++b) {
            void* ptr = peer_buf_ptr(s, b, h);
            if (!ptr) continue;
            void* base = nullptr; size_t size = 0;
            if (lbcomm::address_range(ptr, &base, &size) != 0) { why = "cuMemGetAddressRange unavailable"; break; }
            if (cudaIpcGetMemHandle(&mine[side].handle[b], base) != cudaSuccess) { cudaGetLastError(); why = "cudaIpcGetMemHandle refused"; break; }
            mine[side].offset[b] = (unsigned long long)((char*)ptr - (char*)base);
            mine[side].valid[b] = 1;
        }
    }
    // gather the records of all ranks (a rank that could not export sends invalid records)
    if (!why.empty()) memset(mine, 0, sizeof mine);
    DevBuf<char> dsend, dall;
    CU(dsend.alloc(sizeof mine)); CU(dall.alloc(sizeof mine * (size_t)c.world));
    CU(cudaMemcpyAsync(dsend.p, mine, sizeof mine, cudaMemcpyHostToDevice, st));
    std::vector<PeerExport> all((size_t)2 * c.world);
    if (A.AllGather) {
        NC(A.AllGather(dsend.p, dall.p, sizeof mine, lbcomm::ncclUint8, c.comm, st));
        CU(cudaMemcpyAsync(all.data(), dall.p, sizeof mine * (size_t)c.world, cudaMemcpyDeviceToHost, st));
    }
    CU(cudaStreamSynchronize(st));
    // map the neighbours' arrays
    auto open = [&](const cudaIpcMemHandle_t& hd, void** out) -> bool {
        for (auto& o : P.opened) if (memcmp(&o.first, &hd, sizeof hd) == 0) { *out = o.second; return true; }
        void* base = nullptr;
        if (cudaIpcOpenMemHandle(&base, hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return false; }
        P.opened.push_back({ hd, base });
        *out = base;
        return true;
    };
    for (int dir = 0; dir < 2 && why.empty(); ++dir) {
        const int nb = dir == 0 ? down : up;
        if (nb < 0) continue;
        const PeerExport& r = all[(size_t)2 * nb + (dir == 0 ? 1 : 0)];  // below us: its last slab; above us: its first slab
        P.nbInfo[dir] = r;
        Slab* s = dir == 0 ? h->slabs.front().get() : h->slabs.back().get();
        if (!r.valid[PB_FA] || !r.valid[PB_FLAGS]) { why = "rank " + std::to_string(nb) + " exported nothing"; break; }
        if (r.XY != s->XY) { why = "neighbour slab has another cross-section"; break; }
        for (int b = 0; b < PB_COUNT; ++b) {
            P.nb[dir][b] = nullptr;
            if (!r.valid[b]) continue;
            void* base = nullptr;
            if (!open(r.handle[b], &base)) { why = "cudaIpcOpenMemHandle refused (no peer access to rank " + std::to_string(nb) + "?)"; break; }
            P.nb[dir][b] = (char*)base + r.offset[b];
        }
    }
    // all or nothing across the ranks
    DevBuf<uint32_t> bad;
    CU(bad.alloc(1));
    const uint32_t mineBad = why.empty() ? 0u : 1u;
    CU(cudaMemcpyAsync(bad.p, &mineBad, sizeof mineBad, cudaMemcpyHostToDevice, st));
    NC(A.AllReduce(bad.p, bad.p, 1, lbcomm::ncclUint32, lbcomm::ncclSum, c.comm, st));
    uint32_t anyBad = 0;
    CU(cudaMemcpyAsync(&anyBad, bad.p, sizeof anyBad, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (anyBad) {
        P.note = why.empty() ? "another rank could not map its neighbours" : why;
        if (getenv("LBGPU_VERBOSE")) fprintf(stderr, "[lbgpu rank %d] peer halo off: %s (NCCL send/recv carries the step halo)\n", c.rank, P.note.c_str());
        return 0;
    }
    P.on = true;
    P.seq = 0; P.xseq = 0;
    P.planWhat[0] = P.planWhat[1] = 0;
    P.xplans.clear();
    if (const char* e = getenv("LBGPU_PEER_CYCLE")) P.inCycle = atoi(e) != 0;
    if (getenv("LBGPU_VERBOSE")) fprintf(stderr, "[lbgpu rank %d] peer halo on (neighbours %d / %d mapped)\n", c.rank, down, up);
    return 0;
}

void peer_teardown(LbGpuHandle* h) {
    for (auto& o : h->peer.opened) cudaIpcCloseMemHandle(o.second);
    h->peer.opened.clear();
    h->peer.on = false;
}

// the plane copies of one step halo: fields `what`, destination population buffer `buf` (0 = A)
void build_put_plan(LbGpuHandle* h, uint32_t what, int buf, PutPlan& pl) {
    PeerHalo& P = h->peer;
    memset(&pl, 0, sizeof pl);
    int down, up;
    neighbour_ranks(h, &down, &up);
    auto add = [&](const void* src, char* dst, size_t bytes) {
        if (pl.n >= PUT_MAX) return;
        pl.src[pl.n] = (const char*)src; pl.dst[pl.n] = dst; pl.bytes[pl.n] = (uint32_t)bytes; ++pl.n;
    };
    for (int dir = 0; dir < 2; ++dir) {
        const int nb = dir == 0 ? down : up;
        if (nb < 0) continue;
        const bool isUp = dir == 1;
        Slab* s = isUp ? h->slabs.back().get() : h->slabs.front().get();
        const PeerExport& r = P.nbInfo[dir];
        const size_t XY = s->XY;
        const size_t sp = isUp ? (size_t)s->dev.Z - 2 : 1;      // my face plane
        const size_t dp = isUp ? 0 : (size_t)r.Z - 1;           // the neighbour's ghost plane
        if (what & G_POPS) {
            const double* f = s->fbuf(buf);
            double* nf = (double*)P.nb[dir][buf == 0 ? PB_FA : PB_FB] + r.pad;
            auto pop = [&](int k) { add(f + (size_t)k * s->stride + sp * XY, (char*)(nf + (size_t)k * r.stride + dp * XY), XY * 8); };
            if (h->slip) { for (int k = 0; k < Q; ++k) pop(k); }
            else { const int* ks = isUp ? POPS_UP : POPS_DOWN; for (int q = 0; q < 5; ++q) pop(ks[q]); }
        }
        auto field = [&](const double* mineP, int b) {
            if (mineP && P.nb[dir][b]) add(mineP + sp * XY, P.nb[dir][b] + dp * XY * 8, XY * 8);
        };
        if (what & G_MACRO) { field(s->n.p, PB_N); field(s->ux.p, PB_UX); field(s->uy.p, PB_UY); field(s->uz.p, PB_UZ); }
        if (what & G_VISC) field(s->visc.p, PB_VISC);
        if (what & G_HF) { field(s->hfx.p, PB_HFX); field(s->hfy.p, PB_HFY); field(s->hfz.p, PB_HFZ); }
        // the flag word the neighbour waits on: we are ITS neighbour above when it is below us, and vice versa
        pl.flagDst[dir] = (uint32_t*)P.nb[dir][PB_FLAGS] + (isUp ? 0 : 1);
    }
    pl.counter = P.flags.p + 8;
    pl.chunk = 32768;
}

// step halo through peer memory, on `st` (after the face planes are final)
int peer_put(LbGpuHandle* h, uint32_t what, cudaStream_t st) {
    PeerHalo& P = h->peer;
    const int buf = h->cur ^ 1;
    if (P.planWhat[buf] != what) { build_put_plan(h, what, buf, P.plan[buf]); P.planWhat[buf] = what; }
    const PutPlan& pl = P.plan[buf];
    uint32_t maxBytes = 0;
    for (int k = 0; k < pl.n; ++k) maxBytes = pl.bytes[k] > maxBytes ? pl.bytes[k] : maxBytes;
    if (pl.n == 0) return 0;
    dim3 grid((maxBytes + pl.chunk - 1) / pl.chunk, (unsigned)pl.n);
    k_put_halo<<<grid, 128, 0, st>>>(pl, P.seq);
    ++h->launches;
    CU(cudaGetLastError());
    return 0;
}

// in-cycle exchange of the face planes of `what` through peer memory, on the handle's stream (see k_xput)
int peer_exchange(LbGpuHandle* h, uint32_t what) {
    PeerHalo& P = h->peer;
    cudaStream_t st = h->stream;
    int down, up;
    neighbour_ranks(h, &down, &up);
    const uint64_t key = ((uint64_t)what << 1) | (uint64_t)(h->cur & 1);
    PutPlan* pl = nullptr;
    for (auto& kv : P.xplans) if (kv.first == key) pl = &kv.second;
    if (!pl) {
        P.xplans.emplace_back(key, PutPlan());
        pl = &P.xplans.back().second;
        memset(pl, 0, sizeof *pl);
        auto add = [&](const void* src, char* dst, size_t bytes) {
            if (pl->n >= PUT_MAX) return;
            pl->src[pl->n] = (const char*)src; pl->dst[pl->n] = dst; pl->bytes[pl->n] = (uint32_t)bytes; ++pl->n;
        };
        for (int dir = 0; dir < 2; ++dir) {
            const int nb = dir == 0 ? down : up;
            if (nb < 0) continue;
            const bool isUp = dir == 1;
            Slab* s = isUp ? h->slabs.back().get() : h->slabs.front().get();
            const PeerExport& r = P.nbInfo[dir];
            const size_t XY = s->XY;
            const size_t sp = isUp ? (size_t)s->dev.Z - 2 : 1, dp = isUp ? 0 : (size_t)r.Z - 1;
            auto pops = [&](int buf, bool all) {
                const double* f = s->fbuf(buf);
                double* nf = (double*)P.nb[dir][buf == 0 ? PB_FA : PB_FB] + r.pad;
                auto pop = [&](int k) { add(f + (size_t)k * s->stride + sp * XY, (char*)(nf + (size_t)k * r.stride + dp * XY), XY * 8); };
                if (all) { for (int k = 0; k < Q; ++k) pop(k); }
                else { const int* ks = isUp ? POPS_UP : POPS_DOWN; for (int q = 0; q < 5; ++q) pop(ks[q]); }
            };
            if (what & G_POPS) pops(h->cur ^ 1, h->slip);
            if (what & G_POPS_SRC) pops(h->cur, true);
            auto field = [&](const void* mineP, int b, size_t elem) {
                if (mineP && P.nb[dir][b]) add((const char*)mineP + sp * XY * elem, P.nb[dir][b] + dp * XY * elem, XY * elem);
            };
            if (what & G_TYPE) field(s->tbuf(0), PB_TYPE, 1);
            if (what & G_SOLID) field(s->solidIndex.p, PB_SOLID, 4);
            if (what & G_MASS) field(s->mass.p, PB_MASS, 8);
            if (what & G_MACRO) { field(s->n.p, PB_N, 8); field(s->ux.p, PB_UX, 8); field(s->uy.p, PB_UY, 8); field(s->uz.p, PB_UZ, 8); }
            if (what & G_VISC) field(s->visc.p, PB_VISC, 8);
            if (what & G_HF) { field(s->hfx.p, PB_HFX, 8); field(s->hfy.p, PB_HFY, 8); field(s->hfz.p, PB_HFZ, 8); }
            if ((what & G_MARK) && s->mark.p) field(s->mark.p, PB_MARK, 1);
            uint32_t* nbFlags = (uint32_t*)P.nb[dir][PB_FLAGS];
            // we are the neighbour ABOVE for the rank below us (its slots [3], [5]) and the neighbour BELOW for the rank above ([2], [4])
            pl->readyDst[dir] = nbFlags + (isUp ? 2 : 3);
            pl->flagDst[dir] = nbFlags + (isUp ? 4 : 5);
            pl->readyLocal[dir] = P.flags.p + (isUp ? 3 : 2);
        }
        pl->counter = P.flags.p + 9;
        pl->chunk = 32768;
    }
    ++P.xseq;
    if (pl->n > 0) {
        uint32_t maxBytes = 0;
        for (int k = 0; k < pl->n; ++k) maxBytes = pl->bytes[k] > maxBytes ? pl->bytes[k] : maxBytes;
        dim3 grid((maxBytes + pl->chunk - 1) / pl->chunk, (unsigned)pl->n);
        k_xput<<<grid, 128, 0, st>>>(*pl, P.xseq, h->slabs[0]->status.p);
        k_wait_halo<<<1, 32, 0, st>>>(P.flags.p + 4, P.xseq, down >= 0, up >= 0, h->slabs[0]->status.p);
        h->launches += 2;
    }
    CU(cudaGetLastError());
    return 0;
}

// in-place sum over the ranks of a small device array (no-op in a single-process run)
int allreduce_sum(LbGpuHandle* h, void* buf, size_t count, int dtype);
// the per-element sums of the last LB step over all ranks (collective: every rank calls it at the same point)
int forces_reduced(LbGpuHandle* h) {
    if (!h->forcesPending) return 0;
    h->forcesPending = false;
    return allreduce_sum(h, h->slabs[0]->elemOut.p, (size_t)7 * h->nElmts, lbcomm::ncclFloat64);
}

int allreduce_sum(LbGpuHandle* h, void* buf, size_t count, int dtype) {
    if (!lbcomm::active() || count == 0) return 0;
    NC(lbcomm::api().AllReduce(buf, buf, count, dtype, lbcomm::ncclSum, lbcomm::comm().comm, h->stream));
    return 0;
}

// popsPushed: the local mirrors of everything the step kernel stores (populations, n, u, visc, hydroForce) were
// already written by that kernel (ghost push)
// remote: also move the planes shared with other processes (false when the caller overlaps that transport itself)
int exchange(LbGpuHandle* h, uint32_t what, bool popsPushed = false, bool remote = true) {
    cudaStream_t st = h->stream;
    const uint32_t local = popsPushed ? (what & ~(uint32_t)(G_POPS | G_MACRO | G_VISC | G_HF)) : what;
    for (auto& sp : h->slabs) {
        Slab* s = sp.get();
        if (!s->nGhost || !local) continue;
        Dev d = dev_for(h, s);
        k_fill_ghosts<<<(s->nGhost + BLOCK - 1) / BLOCK, BLOCK, 0, st>>>(d, s->gDst.p, s->gSrc.p, s->gPop.p, 0, s->nGhost, local, s->mark.p);
        ++h->launches;
    }
    const int G = (int)h->slabs.size();
    const bool ring = h->prm.boundary[4] == T_PERIODIC;
    const bool multiProc = lbcomm::active();
    if (G > 1) {
        for (int k = 0; k < G; ++k) {
            Slab* a = h->slabs[k].get();
            const int ku = (k + 1 < G) ? k + 1 : ((ring && !multiProc) ? 0 : -1);
            if (ku < 0) continue;
            Slab* b = h->slabs[ku].get();
            int rc;
            // a's top owned plane -> b's lower ghost plane; b's bottom owned plane -> a's upper ghost plane
            if ((rc = copy_face(h, a, (uint32_t)a->dev.Z - 2, b, 0, what, true))) return rc;
            if ((rc = copy_face(h, b, 1, a, (uint32_t)a->dev.Z - 1, what, false))) return rc;
        }
    }
    if (multiProc && remote) {
        if (h->peer.on && h->peer.inCycle) { if (int rc = peer_exchange(h, what)) return rc; }
        else if (int rc = exchange_remote(h, what, st)) return rc;
    }
    CU(cudaGetLastError());
    return 0;
}

// sums over the slabs of small per-slab device arrays (fixed slab order), result in slab 0's array
__global__ void k_add_arrays(double* __restrict__ dst, const double* __restrict__ src, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] += src[i];
}
__global__ void k_add_counters(unsigned long long* __restrict__ dst, const unsigned long long* __restrict__ src, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] += src[i];
}
__global__ void k_add_u32(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] += src[i];
}

int particle_capacity(LbGpuHandle* h, uint32_t nParts, uint32_t nElmts, uint32_t nComps) {
    if (nParts > h->rawParts.n) { CU(h->rawParts.alloc(nParts + nParts / 2 + 16)); CU(h->parts.alloc(h->rawParts.n)); }
    if (nElmts > h->rawElmts.n) {
        CU(h->rawElmts.alloc(nElmts + nElmts / 2 + 16));
        CU(h->elmts.alloc(h->rawElmts.n));
        for (auto& s : h->slabs) CU(s->elemOut.alloc(h->rawElmts.n * 7));
    }
    if (nComps > h->comps.n) CU(h->comps.alloc(nComps + nComps / 2 + 16));
    return 0;
}

int upload_particles(LbGpuHandle* h, const LbGpuParticle* parts, uint32_t nParts, const LbGpuElement* elmts,
                     uint32_t nElmts, const uint32_t* comps, uint32_t nComps) {
    static_assert(sizeof(LbGpuParticle) == sizeof(RawParticle) && sizeof(LbGpuElement) == sizeof(RawElement), "ABI layout");
    if (h->graph.exec) { cudaGraphExecDestroy(h->graph.exec); h->graph.exec = nullptr; }  // the lists may move
    if (int rc = particle_capacity(h, nParts, nElmts, nComps)) return rc;
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
    if (h->nParts != nParts || h->nElmts != nElmts || h->nComps != nComps) h->eagerCycles = 0;  // buffers may be allocated by the next cycles
    h->nParts = nParts; h->nElmts = nElmts; h->nComps = nComps;
    const uint32_t m = nParts > nElmts ? nParts : nElmts;
    if (m) {
        k_prepare_particles<<<(m + 127) / 128, 128, 0, h->stream>>>(h->rawParts.p, nParts, h->rawElmts.p, nElmts, h->uLength,
                                                                   h->uSpeed, h->parts.p, h->elmts.p);
        ++h->launches;
    }
    return 0;
}

// interface-cell list and visited-tile list of every slab from the current types: count - scan - write for the cells
// (k_list_*), then for the tiles (k_tile_*)
// The interface-cell and candidate lists grow AHEAD of need: the counts of the last list build that has reached the host
// (they trail by a step at most, and an interface moves one cell per step) are compared with half the capacity, so a
// list is never found too small by the kernels that consume it.  Returns true when a list was re-allocated.
bool lists_need_growth(const LbGpuHandle* h) {
    for (size_t q = 0; q < h->slabs.size(); ++q) {
        const Slab* s = h->slabs[q].get();
        if ((h->pinnedCounts[8 * q + 2] > s->cellCap / 2 && s->cellCap < s->N + 16u) || (h->pinnedCounts[8 * q + 4] > s->candCap / 2 && s->candCap < s->N + 16u)) return true;
    }
    return false;
}

int grow_lists(LbGpuHandle* h, bool* grown) {
    *grown = false;
    if (h->capturing) return 0;  // lbGpuRun looks at the counts before it captures or replays a graph
    for (size_t q = 0; q < h->slabs.size(); ++q) {
        Slab* s = h->slabs[q].get();
        const uint32_t cells = h->pinnedCounts[8 * q + 2], cand = h->pinnedCounts[8 * q + 4];
        if (cells <= s->cellCap / 2 && cand <= s->candCap / 2) continue;
        CU(cudaStreamSynchronize(h->stream));
        auto target = [&](uint32_t seen, uint32_t cap) {
            unsigned long long t = 2ull * (seen > cap ? seen : cap) + 1024ull;
            return (uint32_t)(t < (unsigned long long)s->N + 16ull ? t : (unsigned long long)s->N + 16ull);
        };
        if (cells > s->cellCap / 2 && s->cellCap < s->N + 16u) { s->cellCap = target(cells, s->cellCap); CU(s->cellList.alloc(s->cellCap)); *grown = true; }
        if (cand > s->candCap / 2 && s->candCap < s->N + 16u) { s->candCap = target(cand, s->candCap); CU(s->candList.alloc(s->candCap)); *grown = true; }
    }
    return 0;
}

// Slabs of other processes: the face planes of an edge slab are updated by launches of their own, ahead of the interior, so
// that they can travel while the interior is updated (lb_step).  Which faces this handle's slab q splits off:
void face_plan(LbGpuHandle* h, size_t q, bool* loFace, bool* hiFace) {
    *loFace = *hiFace = false;
    if (!lbcomm::active() || h->dynWall) return;
    int down, up;
    neighbour_ranks(h, &down, &up);
    // (an edge slab with fewer than three owned planes has no interior to hide the transport behind)
    if ((down >= 0 && h->slabs.front()->dev.Z < 5) || (up >= 0 && h->slabs.back()->dev.Z < 5)) return;
    *loFace = q == 0 && down >= 0;
    *hiFace = q + 1 == h->slabs.size() && up >= 0;
}

// coarse blocks of a list build (lb_kernels.cuh, SCAN_MAX_BLOCKS): `chunks` 128-thread chunks on `grid` blocks of `per` chunks
struct Coarse { uint32_t grid, per; };
Coarse coarse_blocks(uint32_t chunks) {
    const uint32_t per = chunks > SCAN_MAX_BLOCKS ? (chunks + SCAN_MAX_BLOCKS - 1) / SCAN_MAX_BLOCKS : 1u;
    return { chunks ? (chunks + per - 1) / per : 1u, per };
}

int build_lists(LbGpuHandle* h) {
    cudaStream_t st = h->stream;
    { bool grown; if (int rc = grow_lists(h, &grown)) return rc; }
    for (size_t q = 0; q < h->slabs.size(); ++q) {
        Slab* s = h->slabs[q].get();
        const uint32_t nT = s->blocks * TILES_PER_BLOCK, tb = (nT + BLOCK - 1) / BLOCK;  // tiles of 32 cells
        const Coarse cl = coarse_blocks(s->listBlocks), ct = coarse_blocks(tb);
        k_list_count<<<cl.grid, BLOCK, 0, st>>>(s->tbuf(0), nT, s->tileFlags.p, s->listBlockCount.p, cl.per);
        k_list_write<<<cl.grid, BLOCK, 0, st>>>(dev_all(h, s), nT, s->tileFlags.p, s->listBlockCount.p, s->cellList.p, s->cellCap, cl.per, s->listCounts.p);
        {   // candidates of the update: owners decided on the (identical) pre-update copy of the types
            Dev d = dev_for(h, s);
            d.typeOld = s->tbuf(1);
            const uint32_t groups = (s->N + 15u) / 16u, gb = (groups + BLOCK - 1) / BLOCK;
            const Coarse cc = coarse_blocks(gb);
            k_cand_mark<<<s->listGrid, BLOCK, 0, st>>>(d, s->mark.p, s->tileFlags.p);
            k_plist_count<<<cc.grid, BLOCK, 0, st>>>(s->mark.p, groups, s->candBlockCount.p, nullptr, MARK_CAND, cc.per);
            k_plist_write<<<cc.grid, BLOCK, 0, st>>>(s->mark.p, groups, s->candBlockCount.p, s->candList.p, s->candCap, nullptr, MARK_CAND, cc.per, s->listCounts.p, 3);
            h->launches += 3;
        }
        k_tile_count<<<ct.grid, BLOCK, 0, st>>>(s->tileFlags.p, nT, s->listTileOffset.p, ct.per);
        k_tile_write<<<ct.grid, BLOCK, 0, st>>>(s->tileFlags.p, nT, s->listTileOffset.p, s->tileList.p, ct.per, s->listCounts.p);
        h->launches += 4;
        bool loFace, hiFace;
        face_plan(h, q, &loFace, &hiFace);
        s->hasRanges = loFace || hiFace;
        if (s->hasRanges) {  // where the face planes and the interior lie in the two ascending lists
            if (!s->rangeCells.p) {
                CU(s->rangeCells.alloc(6)); CU(s->listRanges.alloc(12));
                const uint32_t b = s->ownBegin + (loFace ? s->XY : 0u), e = s->ownEnd - (hiFace ? s->XY : 0u);
                const uint32_t rc6[6] = { s->ownBegin, b, b, e, e, s->ownEnd };
                CU(cudaMemcpyAsync(s->rangeCells.p, rc6, sizeof rc6, cudaMemcpyHostToDevice, st));
                CU(cudaStreamSynchronize(st));  // rc6 is on the stack; once per slab
            }
            k_list_ranges<<<1, 32, 0, st>>>(s->tileList.p, s->candList.p, s->listCounts.p, s->rangeCells.p, 3, s->listRanges.p);
            ++h->launches;
        }
        // the host only needs the tile count to size the step kernel's grid; a stale value is fine (grid-stride loop)
        CU(cudaMemcpyAsync(h->pinnedCounts + 8 * q, s->listCounts.p, 5 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    }
    h->listsFresh = true;
    CU(cudaGetLastError());
    return 0;
}

int free_surface_step(LbGpuHandle* h) {
    cudaStream_t st = h->stream;
    int rc;
    if ((rc = build_lists(h))) return rc;
    phase_mark(h, 2);
    // the kernels of the update see the types of before it in typeOld: tbuf(1)
    auto fsdev = [&](Slab* s) {
        Dev d = dev_for(h, s);
        d.typeOld = s->tbuf(1);
        return d;
    };
    for (auto& sp : h->slabs) {
        Slab* s = sp.get();
        const Dev d = fsdev(s);
        k_fs_mass<<<s->listGrid, BLOCK, 0, st>>>(d);
        k_fs_mutate<<<s->listGrid, BLOCK, 0, st>>>(d, s->mark.p);
        h->launches += 2;
    }
    if ((rc = exchange(h, G_MARK | G_TYPE))) return rc;
    for (auto& sp : h->slabs) {
        Slab* s = sp.get();
        k_fs_smooth<<<s->listGrid, BLOCK, 0, st>>>(fsdev(s), s->mark.p, s->partial.p);
        ++h->launches;
    }
    if ((rc = exchange(h, G_TYPE))) return rc;
    for (auto& sp : h->slabs) {
        Slab* s = sp.get();
        k_fs_isolated<0><<<s->listGrid, BLOCK, 0, st>>>(fsdev(s), s->partial.p + s->listGrid, s->counters.p);
        ++h->launches;
    }
    if ((rc = exchange(h, G_TYPE))) return rc;
    for (auto& sp : h->slabs) {
        Slab* s = sp.get();
        CU(cudaMemsetAsync(s->counters.p, 0, sizeof(unsigned long long), st));
        k_fs_isolated<1><<<s->listGrid, BLOCK, 0, st>>>(fsdev(s), s->partial.p + 2 * (size_t)s->listGrid, s->counters.p);
        // three partial arrays (one entry per persistent block), summed in fixed order
        k_reduce_partials<<<1, 1024, 0, st>>>(s->partial.p, s->listGrid, 3, s->sums.p, 0);
        h->launches += 2;
    }
    Slab* s0 = h->slabs[0].get();
    for (size_t k = 1; k < h->slabs.size(); ++k) {
        k_add_arrays<<<1, 32, 0, st>>>(s0->sums.p, h->slabs[k]->sums.p, 3);
        k_add_counters<<<1, 32, 0, st>>>(s0->counters.p, h->slabs[k]->counters.p, 1);
        h->launches += 2;
    }
    if (lbcomm::active()) NC(lbcomm::api().GroupStart());  // one launch for both sums
    if ((rc = allreduce_sum(h, s0->sums.p, 3, lbcomm::ncclFloat64))) return rc;
    if ((rc = allreduce_sum(h, s0->counters.p, 1, lbcomm::ncclUint64))) return rc;
    if (lbcomm::active()) NC(lbcomm::api().GroupEnd());
    k_fs_finalize<<<1, 1, 0, st>>>(s0->sums.p, s0->counters.p, s0->scal.p);
    ++h->launches;
    for (auto& sp : h->slabs) {
        Slab* s = sp.get();
        if (s != s0) CU(cudaMemcpyAsync(s->counters.p, s0->counters.p, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
        k_redistribute<true><<<s->listGrid, BLOCK, 0, st>>>(fsdev(s), s0->scal.p + 1);
        ++h->launches;
    }
    if (h->enforceMass) {
        // LB::enforceMassConservation (LB.cpp:1806-1822; DRUM): total mass, then redistributeMass(-0.01*deficit)
        for (auto& sp : h->slabs) {
            Slab* s = sp.get();
            k_mass_total<<<own_blocks(s), BLOCK, 0, st>>>(fsdev(s), s->partial.p);
            k_reduce_partials<<<1, 1024, 0, st>>>(s->partial.p, own_blocks(s), 1, s->sums.p + 3, 0);
            h->launches += 2;
            if (s != s0) { k_add_arrays<<<1, 32, 0, st>>>(s0->sums.p + 3, s->sums.p + 3, 1); ++h->launches; }
        }
        if ((rc = allreduce_sum(h, s0->sums.p + 3, 1, lbcomm::ncclFloat64))) return rc;
        k_mass_deficit_finalize<<<1, 1, 0, st>>>(s0->sums.p + 3, h->totalMass, s0->counters.p, s0->scal.p + 3);
        ++h->launches;
        for (auto& sp : h->slabs) {
            k_redistribute<true><<<sp->listGrid, BLOCK, 0, st>>>(fsdev(sp.get()), s0->scal.p + 3);
            ++h->launches;
        }
    }
    // new interface cells carry n, u, visc taken from their donors: refresh everything a neighbour may read
    if ((rc = exchange(h, G_TYPE | G_MASS | G_MACRO | G_VISC | G_HF))) return rc;
    phase_mark(h, 3);
    h->typesFlipped = true;  // until the step kernel has run: tbuf(1) holds the types of before this update
    h->listsFresh = false;   // the interface list describes the interface of before this update
    CU(cudaGetLastError());
    return 0;
}

// after the step kernel of a cycle with a free-surface step: tbuf(1) := tbuf(0) wherever types can have changed
// (band tiles, periodic ghost cells, remote ghost planes), marks cleared there
int sync_old_types(LbGpuHandle* h) {
    cudaStream_t st = h->stream;
    for (auto& sp : h->slabs) {
        Slab* s = sp.get();
        k_fs_sync<<<s->listGrid, BLOCK, 0, st>>>(dev_all(h, s), s->tbuf(1), s->mark.p);
        ++h->launches;
        if (s->nGhost) {
            k_fs_sync_ghosts<<<(s->nGhost + BLOCK - 1) / BLOCK, BLOCK, 0, st>>>(dev_all(h, s), s->gDst.p, s->nGhost, s->tbuf(1), s->mark.p);
            ++h->launches;
        }
        const size_t XY = s->XY, top = (size_t)(s->dev.Z - 1) * XY;
        if (s->remoteLo) {
            CU(cudaMemcpyAsync(s->tbuf(1), s->tbuf(0), XY, cudaMemcpyDeviceToDevice, st));
            CU(cudaMemsetAsync(s->mark.p, 0, XY, st));
        }
        if (s->remoteHi) {
            CU(cudaMemcpyAsync(s->tbuf(1) + top, s->tbuf(0) + top, XY, cudaMemcpyDeviceToDevice, st));
            CU(cudaMemsetAsync(s->mark.p + top, 0, XY, st));
        }
    }
    CU(cudaGetLastError());
    return 0;
}

int coupling_step(LbGpuHandle* h, bool rescan) {
    cudaStream_t st = h->stream;
    int rc;
    if (h->nParts == 0) {
        if (rescan) {
            for (auto& sp : h->slabs) { k_clear_p<<<sp->blocks, BLOCK, 0, st>>>(dev_all(h, sp.get())); ++h->launches; }
        }
        return 0;
    }
    const uint32_t wb = (h->nParts * 32 + BLOCK - 1) / BLOCK;
    for (auto& sp : h->slabs) {  // list buffers, on first use
        Slab* s = sp.get();
        if (s->pCap) continue;
        s->pGroups = (s->N + 15u) / 16u;
        s->pScanBlocks = (s->pGroups + BLOCK - 1) / BLOCK;
        s->pCap = s->N + 16;  // every cell may be flagged (a packed bed)
        CU(s->pList.alloc(s->pCap)); CU(s->pCounts.alloc(4)); CU(s->pBlockCount.alloc(s->pScanBlocks));
    }
    if (rescan) {
        for (auto& sp : h->slabs) {
            k_clear_p<<<sp->blocks, BLOCK, 0, st>>>(dev_all(h, sp.get()));
            k_rescan<0><<<wb, BLOCK, 0, st>>>(dev_for(h, sp.get()));
            k_rescan<1><<<wb, BLOCK, 0, st>>>(dev_for(h, sp.get()));
            h->launches += 3;
        }
        if ((rc = exchange(h, G_TYPE | G_SOLID))) return rc;  // the list below holds the flagged ghost cells too
    }
    // flagged cells (ghosts included) of every slab as a list: count - scan - write
    auto build_plist = [&](int gateSlot) -> int {
        for (auto& sp : h->slabs) {
            Slab* s = sp.get();
            const uint32_t* gate = gateSlot >= 0 ? s->status.p + gateSlot : nullptr;
            const Coarse cp = coarse_blocks(s->pScanBlocks);
            k_plist_count<<<cp.grid, BLOCK, 0, st>>>(s->tbuf(0), s->pGroups, s->pBlockCount.p, gate, P_BIT, cp.per);
            k_plist_write<<<cp.grid, BLOCK, 0, st>>>(s->tbuf(0), s->pGroups, s->pBlockCount.p, s->pList.p, s->pCap, gate, P_BIT, cp.per, s->pCounts.p, 0);
            h->launches += 2;
        }
        return 0;
    };
    auto list_dev = [&](Slab* s) {
        Dev d = dev_for(h, s);
        d.list = s->pList.p; d.nList = s->pCounts.p;
        return d;
    };
    const uint32_t lg = (uint32_t)h->numSMs * 16u;
    if ((rc = build_plist(-1))) return rc;
    for (auto& sp : h->slabs) {
        k_find_new_active<<<lg, BLOCK, 0, st>>>(list_dev(sp.get()));
        ++h->launches;
    }
    if ((rc = exchange(h, G_TYPE | G_SOLID | G_HF))) return rc;
    Slab* s0 = h->slabs[0].get();
    // Flood fill, one generation per round.  The count of cells a generation flagged lives in status[1 + (gen & 1)] of
    // every slab (the global sum); the next generation's launches are gated on it, so the first SPEC generations are
    // issued without waiting for the host -- only then is the count read back, once per further generation.
    // (particles move a fraction of a cell per step: the cells they newly cover touch flagged cells, so generation 0 flags
    // them and generation 1 only confirms that nothing is left)
    // One process: the host is not asked at all.  floodGens generations are issued (each gated on the one before, so the
    // idle ones cost a few empty launches); should the last of them still flag cells the fill is unfinished, which
    // check_status reports as an error (LBGPU_FLOOD_GENS raises the number, 0 restores the round trip).  A particle would
    // have to advance more than floodGens - 1 cells in one step for that -- far beyond what the LB step itself tolerates.
    const bool deferred = !lbcomm::active() && h->floodGens >= 2;
    const int SPEC = deferred ? h->floodGens : 2, MAX_GEN = deferred ? h->floodGens : 4096;
    bool converged = false;
    for (int gen = 0; gen < MAX_GEN; ++gen) {
        const int cur = 1 + (gen & 1), prev = 1 + ((gen + 1) & 1);
        const int gate = gen > 0 ? prev : -1;
        // generation 0 reuses the list (flags were only cleared since); later ones see the cells flagged meanwhile
        if (gen > 0) { if ((rc = build_plist(gate))) return rc; }
        for (auto& sp : h->slabs) {
            Slab* s = sp.get();
            CU(cudaMemsetAsync(s->status.p + cur, 0, sizeof(uint32_t), st));
            k_find_new_solid<<<lg, BLOCK, 0, st>>>(list_dev(s), s->status.p + cur, gate >= 0 ? s->status.p + gate : nullptr);
            k_commit_pending<<<s->pScanBlocks, BLOCK, 0, st>>>(s->tbuf(0), s->pGroups, s->status.p + cur);
            h->launches += 2;
        }
        for (size_t k = 1; k < h->slabs.size(); ++k) { k_add_u32<<<1, 32, 0, st>>>(s0->status.p + cur, h->slabs[k]->status.p + cur, 1); ++h->launches; }
        if ((rc = allreduce_sum(h, s0->status.p + cur, 1, lbcomm::ncclUint32))) return rc;
        for (size_t k = 1; k < h->slabs.size(); ++k)
            CU(cudaMemcpyAsync(h->slabs[k]->status.p + cur, s0->status.p + cur, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
        if (deferred) {
            if (gen + 1 == SPEC) {
                k_flood_leftover<<<1, 1, 0, st>>>(s0->status.p + cur, s0->status.p + 4);
                ++h->launches;
                converged = true;
                break;
            }
        } else if (gen + 1 >= SPEC) {
            CU(cudaMemcpyAsync(h->pinnedStatus, s0->status.p + cur, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            for (auto& sp : h->slabs)
                CU(cudaMemcpyAsync(h->pinnedStatus + 16 + sp->slot, sp->pCounts.p + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            for (auto& sp : h->slabs)
                if (h->pinnedStatus[16 + sp->slot] > sp->pCap)
                    return fail(LBGPU_EUNSUPPORTED, "particle coupling: %u flagged cells exceed the list capacity %u", h->pinnedStatus[16 + sp->slot], sp->pCap);
            if (*h->pinnedStatus == 0) { converged = true; break; }
        }
        if ((rc = exchange(h, G_TYPE | G_SOLID))) return rc;
    }
    if (!converged) return fail(LBGPU_EUNSUPPORTED, "particle coupling: the flood fill of LB::findNewSolid did not finish in %d generations", MAX_GEN);
    CU(cudaGetLastError());
    return 0;
}

// Bulk bitmap, static list (PART 2 of the split step kernel) and wall-push list of one slab: count - scan - write.
// Built at init and once more when a pure-fluid lattice switches to the split kernel (first particles).
int build_static(LbGpuHandle* h, Slab* s) {
    cudaStream_t st = h->stream;
    const int wallOk = h->wallPush ? 1 : 0;
    // without a free surface cell activity never changes; with one the bitmap holds the static part
    if (!s->bulk.p) CU(s->bulk.alloc((size_t)s->blocks * (BLOCK / 32) + 1));
    CU(cudaMemsetAsync(s->bulk.p, 0, sizeof(uint32_t) * s->bulk.n, st));
    if (h->fs) k_build_bulk<true><<<s->blocks, BLOCK, 0, st>>>(dev_all(h, s), s->bulk.p, wallOk);
    else k_build_bulk<false><<<s->blocks, BLOCK, 0, st>>>(dev_all(h, s), s->bulk.p, wallOk);
    ++h->launches;
    s->dev.bulk = s->bulk.p;
    DevBuf<uint32_t> bc;
    CU(bc.alloc(s->blocks));
    if (!s->staticCount.p) CU(s->staticCount.alloc(8));  // [1] static cells, [3] wall-push cells (+ [4]: k_list_offsets mirrors slot 3 unclamped)
    CU(cudaMemsetAsync(s->staticCount.p, 0, 8 * sizeof(uint32_t), st));
    if (h->fs) k_static_count<true><<<s->blocks, BLOCK, 0, st>>>(dev_all(h, s), s->bulk.p, bc.p);
    else k_static_count<false><<<s->blocks, BLOCK, 0, st>>>(dev_all(h, s), s->bulk.p, bc.p);
    k_list_offsets<<<1, 1024, 0, st>>>(bc.p, s->blocks, s->staticCount.p, 1, s->N);
    CU(cudaMemcpyAsync(h->pinnedStatus + 8, s->staticCount.p + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    s->nStatic = h->pinnedStatus[8];
    CU(s->staticList.alloc(s->nStatic + 1));
    if (h->fs) k_static_write<true><<<s->blocks, BLOCK, 0, st>>>(dev_all(h, s), s->bulk.p, bc.p, s->staticList.p);
    else k_static_write<false><<<s->blocks, BLOCK, 0, st>>>(dev_all(h, s), s->bulk.p, bc.p, s->staticList.p);
    h->launches += 3;
    s->nWallPush = 0;
    if (h->wallPush) {
        k_wall_count<<<s->blocks, BLOCK, 0, st>>>(dev_all(h, s), s->bulk.p, bc.p);
        k_list_offsets<<<1, 1024, 0, st>>>(bc.p, s->blocks, s->staticCount.p, 3, s->N);
        CU(cudaMemcpyAsync(h->pinnedStatus + 8, s->staticCount.p + 3, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        s->nWallPush = h->pinnedStatus[8];
        CU(s->wallCells.alloc(s->nWallPush + 1)); CU(s->wallMasks.alloc(s->nWallPush + 1));
        k_wall_write<<<s->blocks, BLOCK, 0, st>>>(dev_all(h, s), s->bulk.p, bc.p, s->wallCells.p, s->wallMasks.p);
        h->launches += 3;
    }
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    return 0;
}

// bounce-back slots of the static no-slip walls from the populations in buffer `buf` (k_wall_push)
int wall_push(LbGpuHandle* h, int buf) {
    for (auto& sp : h->slabs) {
        Slab* s = sp.get();
        if (!s->nWallPush) continue;
        k_wall_push<<<(s->nWallPush + BLOCK - 1) / BLOCK, BLOCK, 0, h->stream>>>(dev_for(h, s), s->wallCells.p, s->wallMasks.p, s->nWallPush, s->fbuf(buf));
        ++h->launches;
    }
    return 0;
}

int lb_step(LbGpuHandle* h) {
    const bool first = (h->steps == 0);
    const bool couple = h->nParts > 0;
    const bool macro = h->macroAlways;
    const bool fsOn = h->fs;
    cudaStream_t st = h->stream;
    int rc;
    if (h->hasCurved && !h->curvesSet)
        return fail(LBGPU_EINVAL, "the lattice has curved-wall cells (type 9): call lbGpuSetCurves before the first step");
    const int nSums = 1 + 3 * h->prm.nWalls;
    const bool split = step_is_split(h->shear, macro, couple, fsOn, h->dynWall);
    if (split && !h->wallPush && h->wallPushAllowed) {
        // a pure-fluid lattice got its first particles: from now on the split kernel runs, with the cells next to plain
        // walls on the bulk path; their slots are filled from the current post-collision populations
        h->wallPush = true;
        h->ghostCopy = !lbcomm::active();
        for (auto& sp : h->slabs) { if ((rc = build_static(h, sp.get()))) return rc; }
        if (h->steps > 0) wall_push(h, h->cur);
    }
    StepKernel k = select_step(h->force, h->shear, macro, couple, fsOn, h->dynWall, split ? 1 : 0);
    StepKernel k2 = split ? select_step(h->force, h->shear, macro, couple, fsOn, h->dynWall, 2) : nullptr;
    StepKernel k3 = (split && fsOn) ? select_step(h->force, h->shear, macro, couple, fsOn, h->dynWall, 3) : nullptr;
    if (fsOn && !h->typesFlipped && !h->listsFresh) { if ((rc = build_lists(h))) return rc; }  // a cycle without its own update
    const uint32_t ke = h->kevCount % LbGpuHandle::KEV;
    if (h->dynWall)
        for (auto& sp : h->slabs) CU(cudaMemsetAsync(sp->partial.p, 0, sizeof(double) * (size_t)sp->blocks * nSums, st));
    if (!h->capturing) CU(cudaEventRecord(h->kev0[ke], st));
    const uint32_t what = G_POPS | (fsOn ? (G_MACRO | G_VISC | G_HF) : 0u) | (h->hasCurved ? (G_MACRO | G_VISC) : 0u);
    // Slabs of other processes: the two face planes are updated first and travel (NCCL, comm stream) while the
    // interior is updated.  Moving walls keep the plain order (their per-block partial sums are indexed by block).
    int down = -1, up = -1;
    if (lbcomm::active()) neighbour_ranks(h, &down, &up);
    bool overlap = false;
    {
        bool lo, hi;
        for (size_t q = 0; q < h->slabs.size(); ++q) { face_plan(h, q, &lo, &hi); overlap = overlap || lo || hi; }
    }
    // part: 0 lower face plane, 1 interior (or the whole slab), 2 upper face plane
    auto launch = [&](Slab* s, uint32_t begin, uint32_t end, int part) {
        if (end <= begin) return;
        Dev d = dev_for(h, s);
        // the streaming being evaluated happened under the type map of before this cycle's free-surface step
        if (h->typesFlipped) { d.typeOld = s->tbuf(1); d.lazyMass = 1; }
        d.pStride = s->blocks; d.pBase = 0;
        d.cellBegin = begin; d.cellEnd = end;
        if (fsOn) {
            // one tile per block, sized with the tile count of the last list build that has reached the host (any
            // value is correct: the kernel strides over the list)
            d.list = s->tileList.p; d.nList = s->listCounts.p + 1;
            d.range = s->hasRanges ? s->listRanges.p + 4 * part : nullptr;
            uint32_t g = (h->pinnedCounts[8 * s->slot + 1] + TILES_PER_BLOCK - 1) / TILES_PER_BLOCK;
            g = g + g / 16 + 8;
            if (g > s->blocks) g = s->blocks;
            if (s->hasRanges && part != 1) { const uint32_t gf = (s->XY / TILE + 2 + TILES_PER_BLOCK - 1) / TILES_PER_BLOCK + 1; if (g > gf) g = gf; }  // one plane
            if (h->fsGridPerSM > 0 && g > (uint32_t)(h->fsGridPerSM * h->numSMs)) g = (uint32_t)(h->fsGridPerSM * h->numSMs);
            k<<<g, BLOCK, 0, st>>>(d);
        } else {
            k<<<(end - begin + BLOCK - 1) / BLOCK, BLOCK, 0, st>>>(d);
        }
        ++h->launches;
        if (k2 && s->nStatic) {  // the cells next to walls, shells and periodic faces
            d.list = s->staticList.p; d.nList = s->staticCount.p + 1;
            d.range = nullptr;
            k2<<<(s->nStatic + BLOCK - 1) / BLOCK, BLOCK, 0, st>>>(d);
            ++h->launches;
        }
        if (k3) {  // interface cells old and new
            d.list = s->cellList.p; d.nList = s->listCounts.p;
            d.range = s->hasRanges ? s->listRanges.p + 4 * part + 2 : nullptr;
            k3<<<s->listGrid, BLOCK, 0, st>>>(d);
            ++h->launches;
        }
    };
    std::vector<std::pair<uint32_t, uint32_t>> rest(h->slabs.size());
    for (size_t q = 0; q < h->slabs.size(); ++q) {
        Slab* s = h->slabs[q].get();
        uint32_t b = s->ownBegin, e = s->ownEnd;
        bool lo, hi;
        face_plan(h, q, &lo, &hi);
        if (lo) { launch(s, b, b + s->XY, 0); b += s->XY; }
        if (hi) { launch(s, e - s->XY, e, 2); e -= s->XY; }
        rest[q] = { b, e };
    }
    if (overlap) {
        CU(cudaEventRecord(h->evFaces, st));
        CU(cudaStreamWaitEvent(h->commStream, h->evFaces, 0));
        if (h->peer.on) { ++h->peer.seq; if ((rc = peer_put(h, what, h->commStream))) return rc; }
        else if ((rc = exchange_remote(h, what, h->commStream))) return rc;
        CU(cudaEventRecord(h->evHalo, h->commStream));
    }
    for (size_t q = 0; q < h->slabs.size(); ++q) launch(h->slabs[q].get(), rest[q].first, rest[q].second, 1);
    if (!h->capturing) { CU(cudaEventRecord(h->kev1[ke], st)); ++h->kevCount; }
    // ghostCopy: the mirrors of everything the step kernel stored are copied now (the kernel did not push them)
    const uint32_t whatLocal = what | (h->ghostCopy ? ((macro ? G_MACRO : 0u) | (h->shear ? G_VISC : 0u) | (couple ? G_HF : 0u)) : 0u);
    if ((rc = exchange(h, whatLocal, !h->ghostCopy, !overlap))) return rc;
    if (overlap) {
        CU(cudaStreamWaitEvent(st, h->evHalo, 0));
        if (h->peer.on) {  // ... and the neighbours' planes of this step have landed in our ghost planes
            k_wait_halo<<<1, 32, 0, st>>>(h->peer.flags.p, h->peer.seq, down >= 0, up >= 0, h->slabs[0]->status.p);
            ++h->launches;
        }
    }
    phase_mark(h, 5);
    Slab* s0 = h->slabs[0].get();
    if (h->hasCurved) {
        // streaming through the curved links + their extraMass, after every cell's n, u of this step are in place
        // (ghosts included) and after the population exchange of the step
        for (auto& sp : h->slabs) {
            Slab* s = sp.get();
            if (!s->nStatic) continue;
            Dev d = dev_for(h, s);
            d.pStride = s->blocks; d.pBase = 0;
            d.list = s->staticList.p; d.nList = s->staticCount.p + 1;
            k_curved_stream<<<(s->nStatic + BLOCK - 1) / BLOCK, BLOCK, 0, st>>>(d);
            ++h->launches;
        }
    }
    if (h->wallPush) wall_push(h, h->cur ^ 1);  // bounce-back slots of the plain walls for the next pull
    if (h->dynWall) {
        for (auto& sp : h->slabs) {
            Slab* s = sp.get();
            for (int a = 0; a < nSums; ++a)
                k_reduce_partials<<<1, 1024, 0, st>>>(s->partial.p + (size_t)a * s->blocks, own_blocks(s), 1, s->sums.p + 4 + a, 0);
            h->launches += nSums;
            if (s != s0) { k_add_arrays<<<(nSums + 31) / 32, 32, 0, st>>>(s0->sums.p + 4, s->sums.p + 4, nSums); ++h->launches; }
        }
        if ((rc = allreduce_sum(h, s0->sums.p + 4, (size_t)nSums, lbcomm::ncclFloat64))) return rc;
        if (fsOn) {
            // LB::redistributeMass(extraMass) at the end of LB::streaming (LB.cpp:1477)
            k_extra_mass_finalize<<<1, 1, 0, st>>>(s0->sums.p + 4, s0->counters.p, s0->scal.p + 2);
            ++h->launches;
            if (!h->typesFlipped && !h->listsFresh) { if ((rc = build_lists(h))) return rc; }
            for (auto& sp : h->slabs) {
                if (h->typesFlipped) { Dev d = dev_for(h, sp.get()); d.typeOld = sp->tbuf(1); k_redistribute<true><<<sp->listGrid, BLOCK, 0, st>>>(d, s0->scal.p + 2); }
                else k_redistribute<false><<<sp->listGrid, BLOCK, 0, st>>>(dev_for(h, sp.get()), s0->scal.p + 2);
                ++h->launches;
            }
            if ((rc = exchange(h, G_MASS))) return rc;
        }
    }
    phase_mark(h, 6);
    if (h->nElmts > 0 && couple) {
        // few elements: several blocks share one (the launch should fill the device: ~2 blocks per SM)
        uint32_t split = 1;
        while (split < 64 && h->nElmts * split * 2 <= (uint32_t)h->numSMs * 2) split *= 2;
        const uint32_t eb = h->nElmts * split;
        for (auto& sp : h->slabs) {
            Slab* s = sp.get();
            if (split > 1 && s->elemPartial.n < (size_t)h->nElmts * split * 7) {
                CU(s->elemPartial.alloc((size_t)h->rawElmts.n * 64 * 7));
                CU(s->elemDone.alloc(h->rawElmts.n));
                CU(cudaMemsetAsync(s->elemDone.p, 0, sizeof(uint32_t) * s->elemDone.n, st));
            }
            k_element_forces<<<eb, BLOCK, 0, st>>>(dev_for(h, s), h->uForce, h->uTorque, h->uVolume, s->elemOut.p, split,
                                                   s->elemPartial.p, s->elemDone.p);
            ++h->launches;
            if (s != s0) { k_add_arrays<<<(7 * h->nElmts + 127) / 128, 128, 0, st>>>(s0->elemOut.p, s->elemOut.p, 7 * h->nElmts); ++h->launches; }
        }
        // every rank ends up with the forces of all elements (LB::computeHydroForces sums over the whole lattice) -- when
        // somebody asks for them: the DEM step, lbGpuParticleForces, a checkpoint (forces_reduced).  lbGpuRun's cycles keep
        // the particles fixed and read no forces, so its ranks skip 7 x nElmts doubles of all-reduce per cycle.
        h->forcesPending = lbcomm::active();
        if (!h->lazyForces) { if ((rc = forces_reduced(h))) return rc; }
    }
    if (h->typesFlipped) { if ((rc = sync_old_types(h))) return rc; }
    h->typesFlipped = false;
    phase_mark(h, 7);
    CU(cudaGetLastError());
    if (!h->capturing) ++h->eagerCycles;
    h->cur ^= 1;
    h->macroValid = macro;
    h->lastStepFirst = first;
    h->lastStepCoupled = couple;
    ++h->steps;
    return 0;
}

int check_status(LbGpuHandle* h) {
    if (h->dem.on) {
        CU(cudaMemcpyAsync(h->pinnedStatus, h->dem.flag.p + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        if (*h->pinnedStatus > (uint32_t)lbdem::MAX_NBR)
            return fail(LBGPU_EUNSUPPORTED, "DEM: an element has %u partners within nebrRange, the partner lists hold %d", *h->pinnedStatus, lbdem::MAX_NBR);
    }
    if (h->peer.on) {
        CU(cudaMemcpyAsync(h->pinnedStatus, h->slabs[0]->status.p + 3, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        if (*h->pinnedStatus) return fail(LBGPU_ECOMM, "peer halo: the planes of the rank %s never arrived (20 s)", *h->pinnedStatus == 1 ? "below" : "above");
    }
    if (h->nParts > 0 && !lbcomm::active() && h->floodGens >= 2) {
        CU(cudaMemcpyAsync(h->pinnedStatus, h->slabs[0]->status.p + 4, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        if (*h->pinnedStatus) {
            CU(cudaMemsetAsync(h->slabs[0]->status.p + 4, 0, sizeof(uint32_t), h->stream));
            return fail(LBGPU_EUNSUPPORTED, "particle coupling: the flood fill of LB::findNewSolid was still flagging cells (%u) after %d generations: "
                        "a particle moved several cells in one step (LBGPU_FLOOD_GENS raises the number of generations, 0 asks the host each generation)",
                        *h->pinnedStatus, h->floodGens);
        }
    }
    for (auto& sp : h->slabs) {
        if (h->fs && h->pinnedCounts[8 * sp->slot + 2] > sp->cellCap)
            return fail(LBGPU_EUNSUPPORTED, "free surface: %u interface cells exceed the list capacity %u", h->pinnedCounts[8 * sp->slot + 2], sp->cellCap);
        if (h->fs && h->pinnedCounts[8 * sp->slot + 4] > sp->candCap)
            return fail(LBGPU_EUNSUPPORTED, "free surface: %u candidate cells exceed the list capacity %u", h->pinnedCounts[8 * sp->slot + 4], sp->candCap);
        CU(cudaMemcpyAsync(h->pinnedStatus, sp->status.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        if (*h->pinnedStatus) {
            const unsigned t = *h->pinnedStatus - 1;
            CU(cudaMemsetAsync(sp->status.p, 0, sizeof(uint32_t), h->stream));
            return fail(LBGPU_ETYPE, "TYPE ERROR: an active cell links to a cell of type %u (LB.cpp:1458-1461)", t);
        }
    }
    return 0;
}

// Build one slab: allocate, upload planes [zBegin-1, zEnd+1) of the host arrays (which start at plane hostZ0), ghost lists.
int build_slab(LbGpuHandle* h, Slab* s, int hostZ0, const uint8_t* type_flags, const uint32_t* solidIndex, const double* f,
               const double* n, const double* u, const double* mass, const double* visc) {
    const LbGpuParams* prm = &h->prm;
    cudaStream_t st = h->stream;
    Trace tr("build_slab");
    const int X = prm->size[0], Y = prm->size[1], gZ = prm->size[2];
    const int Zl = s->zEnd - s->zBegin + 2;
    s->XY = (uint32_t)X * (uint32_t)Y;
    s->N = s->XY * (uint32_t)Zl;
    const uint32_t N = s->N;
    s->stride = ((size_t)N + 31) / 32 * 32;
    // the speculative pulls of the step kernel reach one plane + one row + 1 cell (+ the tail of the last block) beyond
    // either end of a population plane
    s->pad = ((size_t)s->XY + X + 2 + 511) / 32 * 32;
    s->blocks = (N + BLOCK - 1) / BLOCK;
    const bool perZ = prm->boundary[4] == T_PERIODIC;
    s->remoteLo = s->zBegin > 1 || (perZ && h->nSlabsGlobal > 1);
    s->remoteHi = s->zEnd < gZ - 1 || (perZ && h->nSlabsGlobal > 1);
    s->ownBegin = s->remoteLo ? s->XY : 0;
    s->ownEnd = s->remoteHi ? N - s->XY : N;
    const size_t hostOff = (size_t)(s->zBegin - 1 - hostZ0) * s->XY;

    CU(s->fA.alloc(s->stride * Q + 2 * s->pad)); CU(s->fB.alloc(s->stride * Q + 2 * s->pad));
    tr.mark("population buffers allocated");
    // every cell's 19 slots of both buffers are written by k_upload_f; only the pads in front of / behind the planes and
    // the (< 32 doubles) tail of each plane need zeros for the speculative pulls (a first touch of 5 GB costs ~40 ms)
    for (double* fb : { s->fA.p, s->fB.p }) {
        CU(cudaMemsetAsync(fb, 0, sizeof(double) * s->pad, st));
        CU(cudaMemsetAsync(fb + s->pad + s->stride * Q, 0, sizeof(double) * s->pad, st));
        if (s->stride > N)
            CU(cudaMemset2DAsync(fb + s->pad + N, sizeof(double) * s->stride, 0, sizeof(double) * (s->stride - N), Q, st));
    }
    tr.mark("their memsets issued");
    {   // ten per-cell arrays out of one allocation (each 256-byte aligned)
        const size_t NA = ((size_t)N + 31) / 32 * 32;
        CU(s->macroPool.alloc(10 * NA));
        DevBuf<double>* arr[10] = { &s->n, &s->ux, &s->uy, &s->uz, &s->mass, &s->visc, &s->shearRate, &s->hfx, &s->hfy, &s->hfz };
        for (int k = 0; k < 10; ++k) arr[k]->set_view(s->macroPool.p + (size_t)k * NA, N);
    }
    tr.mark("macroscopic arrays allocated");
    const size_t NT = ((size_t)N + LIST_CELLS - 1) / LIST_CELLS * LIST_CELLS + BLOCK;  // the last block of a pass reads whole
    CU(s->type0.alloc(NT)); CU(s->solidIndex.alloc(N));
    CU(cudaMemsetAsync(s->type0.p, T_STAT_WALL, NT, st));
    if (h->fs) {
        CU(s->type1.alloc(NT)); CU(s->mark.alloc(NT)); CU(s->newMass.alloc(N)); CU(cudaMemsetAsync(s->mark.p, 0, NT, st));
        s->listBlocks = (s->blocks + LIST_TILES - 1) / LIST_TILES;
        // first capacities: the interface is a sheet (1 cell in 8 at most); the lists grow ahead of need (grow_lists)
        s->cellCap = N / 8 + 1024;
        if (const char* e = getenv("LBGPU_LIST_CAP")) s->cellCap = (uint32_t)atoi(e) > 16u ? (uint32_t)atoi(e) : 16u;  // tests: exercise the growth
        s->candCap = 4 * s->cellCap;
        CU(s->candList.alloc(s->candCap));
        CU(s->candBlockCount.alloc((((size_t)N + 15) / 16 + BLOCK - 1) / BLOCK + 1));
        CU(s->tileFlags.alloc((size_t)s->listBlocks * LIST_TILES * TILES_PER_BLOCK + 16)); CU(s->tileList.alloc((size_t)s->blocks * TILES_PER_BLOCK)); CU(s->cellList.alloc(s->cellCap));
        CU(s->listCounts.alloc(8)); CU(s->listBlockCount.alloc(s->listBlocks)); CU(s->listTileOffset.alloc(((size_t)s->blocks * TILES_PER_BLOCK + BLOCK - 1) / BLOCK + 1));
        CU(cudaMemsetAsync(s->listCounts.p, 0, 8 * sizeof(uint32_t), st));
        CU(cudaMemsetAsync(s->tileFlags.p, 0, s->tileFlags.n, st));
        s->listGrid = (uint32_t)h->numSMs * 16u;
    }
    const int nSums = 1 + 3 * prm->nWalls;
    // per-block partial sums: one slot per block of the widest launch that writes them (the list-driven free-surface
    // kernels run h->numSMs*16 blocks whatever the lattice size)
    const size_t widest = (size_t)s->blocks > (size_t)h->numSMs * 16 ? (size_t)s->blocks : (size_t)h->numSMs * 16;
    const size_t nPartial = widest * (size_t)(3 > nSums ? 3 : nSums);
    CU(s->partial.alloc(nPartial));
    CU(s->sums.alloc(8 + 3 * 64)); CU(s->scal.alloc(8));
    CU(s->counters.alloc(8)); CU(s->status.alloc(8));  // status: [0] type error, [1],[2] flood-fill counts, [3] peer time-out, [4] unfinished flood fill
    CU(cudaMemsetAsync(s->shearRate.p, 0, sizeof(double) * N, st));
    CU(cudaMemsetAsync(s->hfx.p, 0, sizeof(double) * N, st));
    CU(cudaMemsetAsync(s->hfy.p, 0, sizeof(double) * N, st));
    CU(cudaMemsetAsync(s->hfz.p, 0, sizeof(double) * N, st));
    CU(cudaMemsetAsync(s->sums.p, 0, sizeof(double) * s->sums.n, st));
    CU(cudaMemsetAsync(s->scal.p, 0, sizeof(double) * 8, st));
    CU(cudaMemsetAsync(s->counters.p, 0, sizeof(unsigned long long) * 8, st));
    CU(cudaMemsetAsync(s->status.p, 0, sizeof(uint32_t) * 8, st));
    CU(cudaMemsetAsync(s->partial.p, 0, sizeof(double) * nPartial, st));

    tr.mark("allocations + memsets issued");
    Dev& d = s->dev;
    memset(&d, 0, sizeof d);
    d.X = X; d.Y = Y; d.Z = Zl; d.N = N; d.stride = s->stride;
    d.divX = make_div((uint32_t)X); d.divXY = make_div(s->XY);
    for (int k = 0; k < 4; ++k) d.ghost[k] = prm->boundary[k] == T_PERIODIC;
    d.ghost[4] = s->remoteLo || perZ; d.ghost[5] = s->remoteHi || perZ;
    d.perZ = perZ; d.zOff = s->zBegin - 1; d.gZ = gZ;
    // periodic wraps served inside this lattice: the step kernel pushes into the mirroring ghost cells itself
    d.push = (d.ghost[0] ? 1 : 0) | (d.ghost[2] ? 2 : 0) | ((perZ && !s->remoteLo) ? 4 : 0);
    for (int j = 0; j < Q; ++j) d.off[j] = CXh(j) + X * (CYh(j) + Y * CZh(j));
    d.solidIndex = s->solidIndex.p;
    d.n = s->n.p; d.ux = s->ux.p; d.uy = s->uy.p; d.uz = s->uz.p;
    d.mass = s->mass.p; d.newMass = s->newMass.p; d.visc = s->visc.p; d.shearRate = s->shearRate.p;
    d.hfx = s->hfx.p; d.hfy = s->hfy.p; d.hfz = s->hfz.p;
    for (int k = 0; k < 3; ++k) { d.lbF[k] = prm->forceField ? prm->lbF[k] : 0.0; d.lbFInit[k] = prm->lbF[k]; }
    d.initVisc = prm->initDynVisc; d.plasticVisc = prm->plasticVisc; d.yieldStress = prm->yieldStress;
    d.turbConst = prm->turbConst;
    d.S1 = prm->slipCoefficient; d.S2 = 1.0 - prm->slipCoefficient;
    d.uAngVel = h->uAngVel;
    d.omega0 = 1.0 / (0.5 + 3.0 * prm->initDynVisc);           // node::solveCollision (node.cpp:158-165)
    d.omegaf0 = 1.0 - 1.0 / (1.0 + 6.0 * prm->initDynVisc);    // node::addForce (node.cpp:167-184)
    d.nonNewtonian = prm->nonNewtonian; d.turbulence = prm->turbulence;
    d.nWalls = prm->nWalls;
    d.partial = s->partial.p; d.status = s->status.p; d.pStride = s->blocks; d.pBase = 0;
    d.bulk = nullptr;

    // local ghost list: cells of the periodic x/y shells (and of the z shells when z is periodic inside this slab)
    // mirror the cell at the wrapped position (LB.cpp:438-472: per-axis wrap, diagonals wrap twice)
    {
        const bool gx = d.ghost[0], gy = d.ghost[2], gzLocal = perZ && !s->remoteLo;
        std::vector<uint32_t> gd, gs, gp;
        std::vector<uint32_t> pmTable(512, 0xffffffffu);
        { const size_t guess = 2 * ((size_t)X * Y + (size_t)Y * Zl + (size_t)X * Zl) + 64; gd.reserve(guess); gs.reserve(guess); gp.reserve(guess); }
        int cx[Q], cy[Q], cz[Q];  // (cvec() rebuilds its table on every run-time call)
        for (int j = 0; j < Q; ++j) { cx[j] = CXh(j); cy[j] = CYh(j); cz[j] = CZh(j); }
        auto idx = [&](int x, int y, int z) { return (uint32_t)x + (uint32_t)X * ((uint32_t)y + (uint32_t)Y * (uint32_t)z); };
        for (int z = 0; z < Zl; ++z) {
            const bool zs = (z == 0 || z == Zl - 1);
            // remote ghost planes arrive complete (including their x/y ghosts) from the neighbour slab
            if ((z == 0 && s->remoteLo) || (z == Zl - 1 && s->remoteHi)) continue;
            for (int y = 0; y < Y; ++y) {
                const bool ys = (y == 0 || y == Y - 1);
                // rows inside the lattice only touch the two x shell cells (the lattice has 16 M cells, its shell 0.4 M)
                const bool wholeRow = (ys && gy) || (zs && gzLocal);
                if (!wholeRow && !gx) continue;
                for (int x = 0; x < X; x += (wholeRow ? 1 : X - 1)) {
                    const bool xs = (x == 0 || x == X - 1);
                    const bool isG = (xs && gx) || (ys && gy) || (zs && gzLocal);
                    if (!isG) continue;
                    int sx = x, sy = y, sz = z;
                    if (gx) { if (x == 0) sx = X - 2; else if (x == X - 1) sx = 1; }
                    if (gy) { if (y == 0) sy = Y - 2; else if (y == Y - 1) sy = 1; }
                    if (gzLocal) { if (z == 0) sz = Zl - 2; else if (z == Zl - 1) sz = 1; }
                    // population j is pulled out of this ghost by the cell at ghost + c_j, if that is an interior cell (across a
                    // slab cut the pulling cell is the neighbour slab's: its copy of this plane is taken from here).  Whether
                    // coordinate v + c lies inside depends on v's class only: the 19-bit masks of the 125 class triples are
                    // tabulated on first use (the loop over the 18 links per shell cell cost 6 ms on 256^3)
                    auto cls = [](int v, int lo, int hi) {  // bit (c + 1): lo <= v + c <= hi
                        return (uint32_t)((v - 1 >= lo && v - 1 <= hi) ? 1 : 0) | (uint32_t)((v >= lo && v <= hi) ? 2 : 0) | (uint32_t)((v + 1 >= lo && v + 1 <= hi) ? 4 : 0);
                    };
                    const uint32_t key = cls(x, 1, X - 2) | (cls(y, 1, Y - 2) << 3) | (cls(z, s->remoteLo ? 0 : 1, s->remoteHi ? Zl - 1 : Zl - 2) << 6);
                    if (pmTable[key] == 0xffffffffu) {
                        uint32_t m = 0;
                        for (int j = 1; j < Q; ++j)
                            if (((key >> (cx[j] + 1)) & 1u) && ((key >> (3 + cy[j] + 1)) & 1u) && ((key >> (6 + cz[j] + 1)) & 1u)) m |= 1u << j;
                        pmTable[key] = m;
                    }
                    uint32_t pm = pmTable[key];
                    if (h->slip) pm = (1u << Q) - 1u;
                    gd.push_back(idx(x, y, z)); gs.push_back(idx(sx, sy, sz)); gp.push_back(pm);
                }
            }
        }
        s->nGhost = (uint32_t)gd.size();
        if (s->nGhost) {
            CU(s->gDst.alloc(s->nGhost)); CU(s->gSrc.alloc(s->nGhost)); CU(s->gPop.alloc(s->nGhost));
            CU(cudaMemcpyAsync(s->gDst.p, gd.data(), 4 * (size_t)s->nGhost, cudaMemcpyHostToDevice, st));
            CU(cudaMemcpyAsync(s->gSrc.p, gs.data(), 4 * (size_t)s->nGhost, cudaMemcpyHostToDevice, st));
            CU(cudaMemcpyAsync(s->gPop.p, gp.data(), 4 * (size_t)s->nGhost, cudaMemcpyHostToDevice, st));
            CU(cudaStreamSynchronize(st));
            tr.mark("ghost lists: built + uploaded");
            s->ghostIdx = gd; s->ghostSrc = gs;
            s->ghostType.resize(s->nGhost); s->ghostSolid.resize(s->nGhost);
            s->ghostN.assign(s->nGhost, 0.0); s->ghostU.assign((size_t)3 * s->nGhost, 0.0);
            if (type_flags)
                for (uint32_t k = 0; k < s->nGhost; ++k) {
                    s->ghostType[k] = type_flags[hostOff + gd[k]]; s->ghostSolid[k] = solidIndex[hostOff + gd[k]];
                    if (s->ghostType[k] & NODE_BIT) {
                        s->ghostN[k] = n[hostOff + gd[k]];
                        for (int c = 0; c < 3; ++c) s->ghostU[3 * k + c] = u[3 * (hostOff + gd[k]) + c];
                    }
                }
        }
    }

    tr.mark("ghost lists");
    if (!type_flags) { CU(cudaGetLastError()); return 0; }  // device-side initialisation fills the arrays (device_init)
    if (perZ && s->remoteLo && s->zBegin == 1) {
        s->shellTypeLo.assign(type_flags + hostOff, type_flags + hostOff + s->XY);
        s->shellSolidLo.assign(solidIndex + hostOff, solidIndex + hostOff + s->XY);
    }
    if (perZ && s->remoteHi && s->zEnd == gZ - 1) {
        const size_t o = hostOff + (size_t)(Zl - 1) * s->XY;
        s->shellTypeHi.assign(type_flags + o, type_flags + o + s->XY);
        s->shellSolidHi.assign(solidIndex + o, solidIndex + o + s->XY);
    }
    // staged upload: host (pageable) -> device scratch -> SoA
    { int rc;
      if ((rc = h2d_staged(h, s->type0.p, type_flags + hostOff, N))) return rc;
      tr.mark("type up");
      if ((rc = h2d_staged(h, s->solidIndex.p, solidIndex + hostOff, sizeof(uint32_t) * N))) return rc;
      tr.mark("solidIndex up");
      if ((rc = h2d_staged(h, s->n.p, n + hostOff, sizeof(double) * N))) return rc;
      tr.mark("n up");
      if ((rc = h2d_staged(h, s->mass.p, mass + hostOff, sizeof(double) * N))) return rc;
      tr.mark("mass up");
      if ((rc = h2d_staged(h, s->visc.p, visc + hostOff, sizeof(double) * N))) return rc; }
    tr.mark("visc up");
    {
        // the cell-major velocities land in population buffer B, which nothing has written yet (k_upload_f below fills every
        // cell's slots of both buffers), and are split into the three arrays from there: no temporary allocation
        double* scratch = s->fB.p + s->pad;
        if (int rc = h2d_staged(h, scratch, u + 3 * hostOff, sizeof(double) * 3 * N)) return rc;
        k_split3<<<s->blocks, BLOCK, 0, st>>>(N, scratch, s->ux.p, s->uy.p, s->uz.p);
        if (s->stride > N)  // the tails of the planes the scratch covered are zero again for the speculative pulls
            CU(cudaMemset2DAsync(s->fB.p + s->pad + N, sizeof(double) * s->stride, 0, sizeof(double) * (s->stride - N), Q, st));
    }
    tr.mark("u up + split");
    Dev dd = dev_all(h, s);
    if (f) {
        DevBuf<double> tmp;
        CU(tmp.alloc((size_t)Q * N));
        if (int rc = h2d_staged(h, tmp.p, f + (size_t)Q * hostOff, sizeof(double) * Q * N)) return rc;
        k_upload_f<<<s->blocks, BLOCK, 0, st>>>(dd, tmp.p, s->fbuf(0), s->fbuf(1));
        CU(cudaStreamSynchronize(st));
    } else {
        k_upload_f<<<s->blocks, BLOCK, 0, st>>>(dd, nullptr, s->fbuf(0), s->fbuf(1));
    }
    h->launches += 2;
    CU(cudaGetLastError());
    tr.mark("populations");
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

int lbGpuSlabRange(int32_t sizeZ, int32_t nSlabs, int32_t slab, int32_t* zBegin, int32_t* zEnd) {
    if (sizeZ < 3 || nSlabs < 1 || slab < 0 || slab >= nSlabs || nSlabs > sizeZ - 2 || !zBegin || !zEnd)
        return fail(LBGPU_EINVAL, "lbGpuSlabRange: bad arguments");
    const long long inner = sizeZ - 2;
    *zBegin = 1 + (int32_t)((long long)slab * inner / nSlabs);
    *zEnd = 1 + (int32_t)((long long)(slab + 1) * inner / nSlabs);
    return LBGPU_OK;
}

}  // extern "C"

namespace {

struct BoxSetup {
    InitBox box;
    std::vector<InitRegion> regions;
    const LbGpuParticle* parts = nullptr;
    uint32_t nParts = 0;
};

__global__ void k_set_cell0_node(Dev p, InitBox b) {
    // the reference's wall-node loop starts at j = 0, where neighbors[i].d[0] of interior cells is 0 (never assigned,
    // LB.cpp:377-387, 955): cell 0 gets a node whenever it is a wall
    const uint8_t tb = p.type[0];
    const int t = tb & TYPE_MASK;
    if (!is_wall_type(t) || (tb & NODE_BIT)) return;
    p.type[0] = tb | NODE_BIT;
    p.n[0] = 1.0;
    if (t == T_DYN_WALL || t == T_SLIP_DYN)
        for (int k = 0; k < 6; ++k)
            if (b.wallOfBoundary[k] == (int)p.solidIndex[0]) { p.ux[0] = b.wallVel[k][0]; p.uy[0] = b.wallVel[k][1]; p.uz[0] = b.wallVel[k][2]; }
}

// LB::latticeBolzmannInit for a box problem, on the device (kernels: lb_kernels.cuh, k_init_*)
int device_init(LbGpuHandle* h, const BoxSetup& bs) {
    cudaStream_t st = h->stream;
    const InitBox& b = bs.box;
    int rc;
    DevBuf<InitRegion> dreg;
    DevBuf<uint32_t> flag;
    DevBuf<int> maxp;
    CU(dreg.alloc(bs.regions.size() + 1)); CU(flag.alloc(1)); CU(maxp.alloc(3));
    if (!bs.regions.empty()) CU(cudaMemcpyAsync(dreg.p, bs.regions.data(), sizeof(InitRegion) * bs.regions.size(), cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(flag.p, 0, sizeof(uint32_t), st));
    CU(cudaMemsetAsync(maxp.p, 0xff, 3 * sizeof(int), st));  // -1
    for (auto& sp : h->slabs) { k_init_types<<<sp->blocks, BLOCK, 0, st>>>(dev_all(h, sp.get()), b); ++h->launches; }
    // what the reference holds in the cells the device treats as periodic ghosts (dead cells; lbGpuFetchFields reports them)
    for (auto& sp : h->slabs) {
        Slab* s = sp.get();
        if (!s->nGhost) continue;
        std::vector<uint8_t> t(s->N);
        std::vector<uint32_t> si(s->N);
        CU(cudaMemcpyAsync(t.data(), s->type0.p, s->N, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(si.data(), s->solidIndex.p, sizeof(uint32_t) * s->N, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        for (uint32_t k = 0; k < s->nGhost; ++k) { s->ghostType[k] = t[s->ghostIdx[k]]; s->ghostSolid[k] = si[s->ghostIdx[k]]; }
        // cell 0 of the lattice (see k_set_cell0_node) when it is such a dead cell
        if (s->zBegin == 1)
            for (uint32_t k = 0; k < s->nGhost; ++k)
                if (s->ghostIdx[k] == 0 && is_wall_type(s->ghostType[k] & TYPE_MASK)) {
                    s->ghostType[k] |= NODE_BIT;
                    s->ghostN[k] = 1.0;  // node::initialize(initDensity, wall velocity, ...) (LB.cpp:946-996)
                    const int t0 = s->ghostType[k] & TYPE_MASK;
                    if (t0 == T_DYN_WALL || t0 == T_SLIP_DYN)
                        for (int kb = 0; kb < 6; ++kb)
                            if (b.wallOfBoundary[kb] == (int)s->ghostSolid[k])
                                for (int c = 0; c < 3; ++c) s->ghostU[3 * k + c] = b.wallVel[kb][c];
                }
    }
    if ((rc = exchange(h, G_TYPE | G_SOLID))) return rc;
    if (bs.nParts) {
        // LB::initializeParticleBoundaries (LB.cpp:475-495)
        if ((rc = upload_particles(h, bs.parts, bs.nParts, nullptr, 0, nullptr, 0))) return rc;
        const uint32_t wb = (bs.nParts * 32 + BLOCK - 1) / BLOCK;
        for (auto& sp : h->slabs) {
            k_rescan<0><<<wb, BLOCK, 0, st>>>(dev_for(h, sp.get()));
            k_rescan<1><<<wb, BLOCK, 0, st>>>(dev_for(h, sp.get()));
            h->launches += 2;
        }
        if ((rc = exchange(h, G_TYPE | G_SOLID))) return rc;
        CU(cudaStreamSynchronize(st));
        h->nParts = 0; h->nElmts = 0; h->nComps = 0;  // the particles become resident with the first coupling step
    }
    if (!bs.regions.empty()) {
        for (auto& sp : h->slabs) { k_init_gas<<<sp->blocks, BLOCK, 0, st>>>(dev_all(h, sp.get()), dreg.p, (int)bs.regions.size(), flag.p); ++h->launches; }
        if ((rc = allreduce_sum(h, flag.p, 1, lbcomm::ncclUint32))) return rc;  // every rank takes the same branch
        CU(cudaMemcpyAsync(h->pinnedStatus, flag.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (*h->pinnedStatus) {
            if ((rc = exchange(h, G_TYPE))) return rc;
            for (auto& sp : h->slabs) { k_init_closure<0><<<own_blocks(sp.get()), BLOCK, 0, st>>>(dev_for(h, sp.get())); ++h->launches; }
            if ((rc = exchange(h, G_TYPE))) return rc;
            for (auto& sp : h->slabs) { k_init_closure<1><<<own_blocks(sp.get()), BLOCK, 0, st>>>(dev_for(h, sp.get())); ++h->launches; }
            if ((rc = exchange(h, G_TYPE))) return rc;
        }
    }
    for (auto& sp : h->slabs) { k_init_maxp<<<own_blocks(sp.get()), BLOCK, 0, st>>>(dev_for(h, sp.get()), maxp.p); ++h->launches; }
    // the reference height of the hydrostatic density is the highest active cell of the WHOLE lattice
    if (lbcomm::active()) NC(lbcomm::api().AllReduce(maxp.p, maxp.p, 3, lbcomm::ncclInt32, lbcomm::ncclMax, lbcomm::comm().comm, st));
    int mp[3];
    CU(cudaMemcpyAsync(mp, maxp.p, sizeof mp, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const bool any = mp[0] >= 0;
    for (auto& sp : h->slabs) {
        k_init_fields<<<sp->blocks, BLOCK, 0, st>>>(dev_all(h, sp.get()), b, any ? (double)mp[0] : 0.0, any ? (double)mp[1] : 0.0, any ? (double)mp[2] : 0.0);
        ++h->launches;
    }
    if ((rc = exchange(h, G_TYPE))) return rc;
    for (auto& sp : h->slabs) {
        Slab* s = sp.get();
        k_init_wall_nodes<<<own_blocks(s), BLOCK, 0, st>>>(dev_for(h, s), b);
        ++h->launches;
        if (s->zBegin == 1 && !s->dev.ghost[0] && !s->dev.ghost[2] && !s->remoteLo) { k_set_cell0_node<<<1, 1, 0, st>>>(dev_all(h, s), b); ++h->launches; }
    }
    for (auto& sp : h->slabs) {
        Slab* s = sp.get();
        k_upload_f<<<s->blocks, BLOCK, 0, st>>>(dev_all(h, s), nullptr, s->fbuf(0), s->fbuf(1));
        ++h->launches;
    }
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    return 0;
}

int init_impl(const LbGpuParams* prm, const uint8_t* type_flags, const uint32_t* solidIndex, const double* f,
              const double* n, const double* u, const double* mass, const double* visc, const BoxSetup* box, LbGpuHandle** out) {
    if (!prm || !out || (!box && (!type_flags || !solidIndex || !n || !u || !mass || !visc))) return fail(LBGPU_EINVAL, "lbGpuInit: null argument");
    *out = nullptr;
    Trace tr("lbGpuInit");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(LBGPU_ENODEVICE, "lbGpuInit: no CUDA device available (this engine has no CPU fallback)");
    }
    for (int k = 0; k < 3; ++k)
        if (prm->size[k] < 3) return fail(LBGPU_EINVAL, "lbGpuInit: lbSize[%d]=%d < 3", k, prm->size[k]);
    for (int a = 0; a < 3; ++a) {
        const bool lo = prm->boundary[2 * a] == T_PERIODIC, hi = prm->boundary[2 * a + 1] == T_PERIODIC;
        if (lo != hi) return fail(LBGPU_EINVAL, "lbGpuInit: boundary%d/%d must both be periodic (4) or neither", 2 * a, 2 * a + 1);
    }
    if (prm->nWalls < 0 || prm->nWalls > 64) return fail(LBGPU_EINVAL, "lbGpuInit: nWalls=%d out of range", prm->nWalls);
    const int G = prm->nSlabs > 1 ? prm->nSlabs : 1;
    const int first = G > 1 ? prm->slabIndex : 0;
    const int nLocal = G > 1 ? (prm->nLocalSlabs > 0 ? prm->nLocalSlabs : 1) : 1;
    if (G > 1 && prm->slabAxis != 2) return fail(LBGPU_EUNSUPPORTED, "lbGpuInit: slabs are cut along z (slabAxis=2)");
    if (first < 0 || first + nLocal > G || G > prm->size[2] - 2) return fail(LBGPU_EINVAL, "lbGpuInit: slab %d+%d of %d", first, nLocal, G);
    if (first != 0 || nLocal != G) {
        const lbcomm::Comm& c = lbcomm::comm();
        if (!lbcomm::active()) return fail(LBGPU_ECOMM, "lbGpuInit: slabs %d..%d of %d live in other processes: call lbGpuCommInit first", first, first + nLocal - 1, G);
        if (c.world * nLocal != G || first != c.rank * nLocal)
            return fail(LBGPU_EINVAL, "lbGpuInit: rank %d of %d must own slabs [%d, %d) of %d", c.rank, c.world, c.rank * nLocal, (c.rank + 1) * nLocal, c.world * nLocal);
    } else if (lbcomm::active() && G > 1) {
        return fail(LBGPU_EINVAL, "lbGpuInit: a communicator of %d ranks is active but this handle claims every slab", lbcomm::comm().world);
    }
    int32_t zLo, zHi, tmpz;
    lbGpuSlabRange(prm->size[2], G, first, &zLo, &tmpz);
    lbGpuSlabRange(prm->size[2], G, first + nLocal - 1, &tmpz, &zHi);
    const unsigned long long XY = (unsigned long long)prm->size[0] * prm->size[1];
    const unsigned long long Nh = XY * (unsigned long long)(zHi - zLo + 2);  // cells in the host arrays
    if (Nh >= (1ull << 31)) return fail(LBGPU_EINVAL, "lbGpuInit: %llu cells exceed the 2^31 index range", Nh);
    bool anyDyn = false, anyGas = false, anyIface = false, anySlip = false, anyCurved = false;
    if (box) {
        for (int k = 0; k < 6; ++k) {
            anyDyn |= (prm->boundary[k] == T_DYN_WALL || prm->boundary[k] == T_SLIP_DYN);
            anySlip |= (prm->boundary[k] == T_SLIP_STAT || prm->boundary[k] == T_SLIP_DYN);
        }
        anyGas = anyIface = !box->regions.empty();
    }
    if (type_flags) {
        // which cell types occur: byte histogram on a few threads (a branch per cell costs 30 ms on 16 M cells)
        const int T = copy_threads();
        std::vector<std::vector<unsigned long long>> hist((size_t)T, std::vector<unsigned long long>(256, 0ull));
        std::vector<std::thread> th;
        const unsigned long long per = (Nh + T - 1) / T;
        auto work = [&](int k) {
            unsigned long long* hk = hist[(size_t)k].data();
            const unsigned long long b = (unsigned long long)k * per, e = b + per < Nh ? b + per : Nh;
            for (unsigned long long i = b; i < e; ++i) ++hk[type_flags[i]];
        };
        for (int k = 1; k < T; ++k) th.emplace_back(work, k);
        work(0);
        for (auto& t : th) t.join();
        bool present[16] = {};
        for (int k = 0; k < T; ++k) for (int v = 0; v < 256; ++v) if (hist[(size_t)k][(size_t)v]) present[v & LBGPU_TYPE_MASK] = true;
        for (int t = 0; t < 16; ++t) {
            if (!present[t]) continue;
            if (t == 1 || t > 9) {
                unsigned long long i = 0;
                while (i < Nh && (type_flags[i] & LBGPU_TYPE_MASK) != t) ++i;
                return fail(LBGPU_EINVAL, "lbGpuInit: cell %llu has undefined type %d", i, t);
            }
        }
        anyCurved = present[T_CURVED];
        anyDyn = present[T_DYN_WALL] || present[T_SLIP_DYN];
        anySlip = present[T_SLIP_STAT] || present[T_SLIP_DYN];
        anyGas = present[T_GAS];
        anyIface = present[T_INTERFACE];
    }

    tr.mark("type scan");
    LbGpuHandle* h = new (std::nothrow) LbGpuHandle();
    if (!h) return fail(LBGPU_EINVAL, "out of host memory");
    h->prm = *prm;
    h->nSlabsGlobal = G; h->firstSlab = first;
    int rc = 0;
    auto body = [&]() -> int {
        if (prm->device >= 0) {
            if (prm->device >= ndev) return fail(LBGPU_EINVAL, "lbGpuInit: device %d of %d", prm->device, ndev);
            CU(cudaSetDevice(prm->device));
        }
        CU(cudaGetDevice(&h->device));
        CU(cudaDeviceGetAttribute(&h->numSMs, cudaDevAttrMultiProcessorCount, h->device));
        CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        CU(cudaEventCreate(&h->evA));
        CU(cudaEventCreate(&h->evB));
        {   // the halo transport must not queue behind the interior launch's blocks: highest priority
            int prLo = 0, prHi = 0;
            CU(cudaDeviceGetStreamPriorityRange(&prLo, &prHi));
            CU(cudaStreamCreateWithPriority(&h->commStream, cudaStreamNonBlocking, prHi));
        }
        CU(cudaEventCreateWithFlags(&h->evFaces, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&h->evHalo, cudaEventDisableTiming));
        h->kev0.resize(LbGpuHandle::KEV); h->kev1.resize(LbGpuHandle::KEV);
        for (uint32_t k = 0; k < LbGpuHandle::KEV; ++k) { CU(cudaEventCreate(&h->kev0[k])); CU(cudaEventCreate(&h->kev1[k])); }
        h->fs = prm->freeSurface != 0;
        if (const char* e = getenv("LBGPU_FS_GRID")) h->fsGridPerSM = atoi(e);
        if (const char* e = getenv("LBGPU_WALL_PUSH")) h->wallPushAllowed = atoi(e) != 0;
        if (const char* e = getenv("LBGPU_GRAPH")) h->graphAllowed = atoi(e) != 0;
        if (const char* e = getenv("LBGPU_FLOOD_GENS")) h->floodGens = atoi(e);
        if (const char* e = getenv("LBGPU_PREFETCH")) h->prefetchBlocks = (uint32_t)atoi(e);
        if (h->prefetchBlocks == 0xffffffffu) h->prefetchBlocks = (uint32_t)h->numSMs * 5u;
        h->prefetchTiles = (uint32_t)h->numSMs * 5u;  // A/B, distance 0 / 370 / 740 / 1480 blocks: cfg4 0.504 / 0.493 / 0.492 / 0.498, cfg5 5.33 / 5.22 / 5.18 / 5.26 ms per step
        if (const char* e = getenv("LBGPU_PREFETCH_TILES")) h->prefetchTiles = (uint32_t)atoi(e);

        h->shear = prm->nonNewtonian || prm->turbulence;
        h->force = prm->forceField && (prm->lbF[0] != 0.0 || prm->lbF[1] != 0.0 || prm->lbF[2] != 0.0);
        // curved links use the moving-wall machinery: stored n, u and the extraMass sum (LB.cpp:1278-1319)
        h->hasCurved = anyCurved;
        h->dynWall = anyDyn || anyCurved;
        h->slip = anySlip;
        h->macroAlways = h->fs || anyDyn || anyCurved || anyGas || anyIface;
        // lattices that run the split step kernel from the start serve plain-wall links through pre-stored slots
        h->wallPush = h->wallPushAllowed && step_is_split(h->shear, h->macroAlways, false, h->fs, h->dynWall);
        // (between processes the face planes travel before the step ends: there the step kernel keeps pushing)
        h->ghostCopy = h->wallPush && !lbcomm::active();
        // measureUnits::setComposite (node.cpp:476-488)
        const double L = prm->unitLength, Tm = prm->unitTime, D = prm->unitDensity;
        h->uLength = L; h->uVolume = L * L * L; h->uSpeed = L / Tm; h->uAngVel = 1.0 / Tm;
        h->uForce = D * L * L * L * L / Tm / Tm; h->uTorque = D * L * L * L * L * L / Tm / Tm;
        CU(cudaMallocHost((void**)&h->pinnedStatus, 4096));
        CU(cudaMallocHost((void**)&h->pinnedCounts, sizeof(uint32_t) * 8 * (size_t)nLocal));
        memset(h->pinnedCounts, 0, sizeof(uint32_t) * 8 * (size_t)nLocal);
        tr.mark("streams, events, pinned");
        for (int k = 0; k < nLocal; ++k) {
            h->slabs.emplace_back(new Slab());
            Slab* s = h->slabs.back().get();
            s->index = first + k; s->slot = k;
            int32_t zb, ze;
            lbGpuSlabRange(prm->size[2], G, first + k, &zb, &ze);
            s->zBegin = zb; s->zEnd = ze;
            if (int r = build_slab(h, s, zLo - 1, type_flags, solidIndex, f, n, u, mass, visc)) return r;
        }
        tr.mark("slabs built");
        if (box) { if (int r = device_init(h, *box)) return r; }
        // ghosts of every field, both population buffers
        if (int r = exchange(h, G_POPS | G_POPS_SRC | G_TYPE | G_SOLID | G_MASS | G_MACRO | G_VISC | G_HF)) return r;
        cudaStream_t st = h->stream;
        for (auto& sp : h->slabs) {
            Slab* s = sp.get();
            if (h->fs) CU(cudaMemcpyAsync(s->type1.p, s->type0.p, s->type0.n, cudaMemcpyDeviceToDevice, st));
            if (int r = build_static(h, s)) return r;
            // initial interface count (LB::redistributeMass divides by interfaceNodes.size())
            k_count<<<own_blocks(s), BLOCK, 0, st>>>(dev_for(h, s), s->counters.p + 1);
            ++h->launches;
        }
        if (h->fs) {  // valid until the first free-surface step rebuilds them
            if (int r = build_lists(h)) return r;
            // the initial interface decides the first capacities (a lattice that is mostly interface band is legal)
            for (int pass = 0; pass < 4; ++pass) {
                CU(cudaStreamSynchronize(st));
                bool grown = false;
                if (int r = grow_lists(h, &grown)) return r;
                if (!grown) break;
                for (auto& sp : h->slabs) CU(cudaMemsetAsync(sp->mark.p, 0, sp->mark.n, st));  // marks of the overflowed build
                if (int r = build_lists(h)) return r;
            }
        }
        Slab* s0 = h->slabs[0].get();
        for (size_t k = 1; k < h->slabs.size(); ++k) { k_add_counters<<<1, 32, 0, st>>>(s0->counters.p + 1, h->slabs[k]->counters.p + 1, 3); ++h->launches; }
        if (int r = allreduce_sum(h, s0->counters.p + 1, 3, lbcomm::ncclUint64)) return r;
        for (auto& sp : h->slabs) CU(cudaMemcpyAsync(sp->counters.p, s0->counters.p + 2, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
        CU(cudaStreamSynchronize(st));
        CU(cudaGetLastError());
        tr.mark("ghosts, static lists, counts");
        if (int r = peer_setup(h)) return r;
        tr.mark("peer set-up");
        return 0;
    };
    rc = body();
    if (rc) { lbGpuFinalize(h); return rc; }
    *out = h;
    return LBGPU_OK;
}

}  // namespace

extern "C" {

int lbGpuInit(const LbGpuParams* prm, const uint8_t* type_flags, const uint32_t* solidIndex, const double* f,
              const double* n, const double* u, const double* mass, const double* visc, LbGpuHandle** out) {
    if (!type_flags) return fail(LBGPU_EINVAL, "lbGpuInit: null argument");
    return init_impl(prm, type_flags, solidIndex, f, n, u, mass, visc, nullptr, out);
}

int lbGpuInitBox(const LbGpuParams* prm, const double initVelocity[3], const double* wallVelocity, const LbGpuRegion* regions,
                 uint32_t nRegions, const LbGpuParticle* parts, uint32_t nParts, LbGpuHandle** out) {
    if (!prm || !out || (nRegions && !regions) || (nParts && !parts)) return fail(LBGPU_EINVAL, "lbGpuInitBox: null argument");
    if (prm->boundary[4] == T_PERIODIC && prm->nSlabs > 1) return fail(LBGPU_EUNSUPPORTED, "lbGpuInitBox: z-periodic lattice cut into slabs");
    BoxSetup bs;
    memset(&bs.box, 0, sizeof bs.box);
    int nWalls = 0;
    const double speed = prm->unitLength / prm->unitTime;  // measureUnits::setComposite (node.cpp:476-488)
    for (int k = 0; k < 6; ++k) {
        const int bt = prm->boundary[k];
        if (bt < T_PERIODIC || bt > T_DYN_WALL) return fail(LBGPU_EINVAL, "lbGpuInitBox: boundary%d = %d (4..8 expected)", k, bt);
        bs.box.boundary[k] = bt;
        bs.box.wallOfBoundary[k] = (bt >= T_SLIP_STAT) ? nWalls++ : -1;  // DEM::initializeWalls (DEM.cpp:435-640)
        const bool moving = (bt == T_SLIP_DYN || bt == T_DYN_WALL);
        for (int c = 0; c < 3; ++c) bs.box.wallVel[k][c] = (moving && wallVelocity) ? wallVelocity[3 * k + c] / speed : 0.0;
    }
    if (nWalls != prm->nWalls) return fail(LBGPU_EINVAL, "lbGpuInitBox: %d boundaries are walls but nWalls = %d", nWalls, prm->nWalls);
    bs.box.nRegions = (int)nRegions;
    for (int c = 0; c < 3; ++c) { bs.box.lbF[c] = prm->lbF[c]; bs.box.initVelocity[c] = initVelocity ? initVelocity[c] : 0.0; }
    bs.box.initVisc = prm->initDynVisc;
    for (uint32_t r = 0; r < nRegions; ++r) {
        InitRegion g;
        g.kind = regions[r].kind; g.gasInside = regions[r].gasInside;
        if (g.kind < 0 || g.kind > 2) return fail(LBGPU_EINVAL, "lbGpuInitBox: region %u has kind %d", r, g.kind);
        for (int c = 0; c < 6; ++c) g.a[c] = regions[r].a[c];
        bs.regions.push_back(g);
    }
    bs.parts = parts; bs.nParts = nParts;
    return init_impl(prm, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, &bs, out);
}

int lbGpuSetCurves(LbGpuHandle* h, uint32_t nCurves, const uint32_t* cells, const double* delta) {
    if (!h) return fail(LBGPU_EINVAL, "lbGpuSetCurves: null handle");
    if (nCurves && (!cells || !delta)) return fail(LBGPU_EINVAL, "lbGpuSetCurves: null array");
    if (h->prm.boundary[4] == T_PERIODIC && h->nSlabsGlobal > 1)
        return fail(LBGPU_EUNSUPPORTED, "lbGpuSetCurves: curved walls on a z-periodic lattice cut into slabs");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    CU(h->curveDelta.alloc((size_t)nCurves * Q + 1));
    if (nCurves) CU(cudaMemcpy(h->curveDelta.p, delta, sizeof(double) * Q * (size_t)nCurves, cudaMemcpyHostToDevice));
    // host arrays start at the plane below the handle's first slab
    const long long hostZ0 = (long long)h->slabs.front()->zBegin - 1;
    for (auto& sp : h->slabs) {
        Slab* s = sp.get();
        std::vector<uint32_t> row(s->N, 0xffffffffu);
        const long long shift = ((long long)s->zBegin - 1 - hostZ0) * (long long)s->XY;
        for (uint32_t k = 0; k < nCurves; ++k) {
            const long long loc = (long long)cells[k] - shift;
            if (loc >= 0 && loc < (long long)s->N) row[(size_t)loc] = k;
        }
        // a periodic ghost cell stands for the cell it mirrors (the reference's own curve object of a shell cell is
        // never reached: LB.cpp:438-472 wraps the links away from it)
        for (size_t k = 0; k < s->ghostIdx.size(); ++k) row[s->ghostIdx[k]] = row[s->ghostSrc[k]];
        CU(s->curveRow.alloc(s->N));
        CU(cudaMemcpy(s->curveRow.p, row.data(), sizeof(uint32_t) * s->N, cudaMemcpyHostToDevice));
    }
    CU(cudaDeviceSynchronize());  // the copies ran on the legacy stream; the handle's stream does not order against it
    h->curvesSet = true;
    return LBGPU_OK;
}

int lbGpuSetMassTarget(LbGpuHandle* h, double totalMass) {
    if (!h) return fail(LBGPU_EINVAL, "lbGpuSetMassTarget: null handle");
    if (!h->fs) return fail(LBGPU_EINVAL, "lbGpuSetMassTarget: the lattice has no free surface");
    h->enforceMass = true;
    h->totalMass = totalMass;
    if (h->graph.exec) { cudaGraphExecDestroy(h->graph.exec); h->graph.exec = nullptr; }  // the captured cycle had no mass target
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
    phase_mark(h, 0); phase_mark(h, 1);
    if (doFreeSurface && h->fs) { if ((rc = free_surface_step(h))) return rc; } else { phase_mark(h, 2); phase_mark(h, 3); }
    // LB::computeHydroForces (LB.cpp:1851-1919) applies the direct forcing on every cell flagged inside a particle whether
    // or not goCycle ran the coupling step (demSolve = 0 keeps the flags of the initialisation): the particle lists
    // become resident whenever they are given, only the flag update is tied to doCoupling
    if (doCoupling || nParts > 0) {
        if ((rc = upload_particles(h, parts, nParts, elmts, nElmts, components, nComponents))) return rc;
    }
    if (doCoupling) { if ((rc = coupling_step(h, rescanParticles != 0))) return rc; }
    phase_mark(h, 4);
    if ((rc = lb_step(h))) return rc;
    CU(cudaEventRecord(h->evB, h->stream));
    return LBGPU_OK;
}

int lbGpuCouple(LbGpuHandle* h, int rescanParticles, const LbGpuParticle* parts, uint32_t nParts, const LbGpuElement* elmts,
                uint32_t nElmts, const uint32_t* components, uint32_t nComponents) {
    if (!h) return fail(LBGPU_EINVAL, "lbGpuCouple: null handle");
    if (nParts && (!parts || !elmts || !components)) return fail(LBGPU_EINVAL, "lbGpuCouple: particle arrays missing");
    CU(cudaSetDevice(h->device));
    int rc;
    if ((rc = upload_particles(h, parts, nParts, elmts, nElmts, components, nComponents))) return rc;
    return coupling_step(h, rescanParticles != 0);
}

}  // extern "C"
namespace {
// `count` cycles.  A free-surface or coupled cycle is 10-25 small launches around the step kernel: on small lattices their
// issue cost bounds the cycle.  In one process the cycle is a fixed sequence of stream operations (the flood fill of the
// coupling step gates its generations on the device, see coupling_step; so are the DEM sub-steps), so two consecutive cycles
// -- the population buffers alternate -- are captured once and replayed.  The last cycles of a call run eagerly: they carry
// the CUDA events lbGpuLastKernelMs reads, and the list counts the host sizes grids with.  kind: which cycle (0 lbGpuRun, 1
// lbGpuRunDem) the cached graph holds.
template <class Cycle>
int run_cycles(LbGpuHandle* h, uint32_t count, bool fsCycle, int kind, Cycle&& cycle) {
    uint32_t k = 0;
    const bool coupled = h->nParts > 0;
    const bool graphable = h->graphAllowed && !h->phaseOn && (fsCycle || coupled) && (!h->fs || fsCycle) &&
                           (!coupled || h->floodGens >= 2) && !lbcomm::active() && !h->dynWall && !(kind == 1 && h->dem.pbc) && count >= 8;
    if (graphable) {
        constexpr uint32_t TAIL = 2;
        for (; k < count && (h->steps < 2 || h->eagerCycles < 2); ++k) { if (int rc = cycle()) return rc; }
        while (count - k >= 2 + TAIL) {
            if (lists_need_growth(h)) {  // an eager cycle re-allocates the lists; the graph holds the old pointers
                if (h->graph.exec) { cudaGraphExecDestroy(h->graph.exec); h->graph.exec = nullptr; }
                if (int rc = cycle()) return rc;
                ++k;
                continue;
            }
            bool valid = h->graph.exec != nullptr && h->graph.cur == h->cur && h->graph.fs == (int)fsCycle && h->graph.nParts == h->nParts && h->graph.kind == kind;
            for (size_t q = 0; valid && q < h->slabs.size(); ++q) {
                const uint32_t now = h->pinnedCounts[8 * q + 1], then = h->graph.tiles[q];
                valid = now <= then + then / 32 && now + now / 4 + 64 >= then;  // the grids carry 1/16 of headroom
            }
            if (!valid) {
                if (h->graph.exec) { cudaGraphExecDestroy(h->graph.exec); h->graph.exec = nullptr; }
                const uint64_t steps0 = h->steps, launches0 = h->launches;
                h->graph.tiles.resize(h->slabs.size());
                for (size_t q = 0; q < h->slabs.size(); ++q) h->graph.tiles[q] = h->pinnedCounts[8 * q + 1];
                cudaGraph_t g = nullptr;
                CU(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
                h->capturing = true;
                int rc = cycle();
                if (!rc) rc = cycle();
                h->capturing = false;
                const cudaError_t ce = cudaStreamEndCapture(h->stream, &g);
                h->steps = steps0;  // nothing ran yet
                h->graph.launches = h->launches - launches0;
                h->launches = launches0;
                if (rc) { if (g) cudaGraphDestroy(g); return rc; }
                if (ce != cudaSuccess) return fail(LBGPU_ECUDA, "graph capture of the cycle: %s", cudaGetErrorString(ce));
                const cudaError_t ie = cudaGraphInstantiate(&h->graph.exec, g, 0);
                cudaGraphDestroy(g);
                if (ie != cudaSuccess) { h->graph.exec = nullptr; return fail(LBGPU_ECUDA, "graph instantiation: %s", cudaGetErrorString(ie)); }
                h->graph.cur = h->cur; h->graph.fs = (int)fsCycle; h->graph.nParts = h->nParts; h->graph.kind = kind;
                ++h->graph.captures;
            }
            CU(cudaGraphLaunch(h->graph.exec, h->stream));
            h->steps += 2; h->launches += h->graph.launches; ++h->graph.replays;
            k += 2;
        }
    }
    for (; k < count; ++k) { if (int rc = cycle()) return rc; }
    return 0;
}
}  // namespace
extern "C" {

int lbGpuRun(LbGpuHandle* h, int doFreeSurface, uint32_t count) {
    if (!h) return fail(LBGPU_EINVAL, "lbGpuRun: null handle");
    CU(cudaSetDevice(h->device));
    CU(cudaEventRecord(h->evA, h->stream));
    h->kevCount = 0;
    auto cycle = [&]() -> int {
        int rc;
        phase_mark(h, 0); phase_mark(h, 1);
        if (doFreeSurface && h->fs) { if ((rc = free_surface_step(h))) return rc; } else { phase_mark(h, 2); phase_mark(h, 3); }
        if (h->nParts > 0) { if ((rc = coupling_step(h, false))) return rc; }  // goCycle's order, particles of the last upload
        phase_mark(h, 4);
        return lb_step(h);
    };
    struct Lazy { LbGpuHandle* h; explicit Lazy(LbGpuHandle* hh) : h(hh) { h->lazyForces = true; } ~Lazy() { h->lazyForces = false; } } lazy(h);
    if (int rc = run_cycles(h, count, h->fs && doFreeSurface, 0, cycle)) return rc;
    if (int rc = forces_reduced(h)) return rc;  // of the last cycle
    CU(cudaEventRecord(h->evB, h->stream));
    return LBGPU_OK;
}

}  // extern "C"
namespace {
// DEM::discreteElementStep (DEM.cpp:331-376) on the device, then the lists the LB side reads.  hydro: 7 doubles per
// element on the device, or nullptr = no hydrodynamic force yet (before the first LB step)
int dem_step(LbGpuHandle* h, const double* hydro) {
    static_assert(sizeof(lbdem::OutParticle) == sizeof(RawParticle) && sizeof(lbdem::OutElement) == sizeof(RawElement), "list layout");
    auto& D = h->dem;
    cudaStream_t st = h->stream;
    const uint32_t n = D.n, nb = (n + 127) / 128, nP = D.nP, pb = (D.cap + 127) / 128;
    const uint32_t* nPdev = D.pbc ? D.cnt.p : nullptr;  // with periodic boundaries the particle count lives on the device
    for (int sub = 0; sub < D.prm.multiStep; ++sub) {
        lbdem::k_dem_trigger<<<1, 1024, 0, st>>>(D.e.p, n, D.prm.deltat, D.prm.nebrRange, D.scal.p, D.flag.p);
        if (D.pbc) {
            lbdem::k_dem_pbc<<<1, 1024, 0, st>>>(D.e.p, n, D.prm, D.pt.p, D.flag.p, D.gflag.p, D.basePos.p, D.cornPos.p, h->comps.p, D.nComp.p, D.cnt.p);
            ++h->launches;
        }
        if (D.grid) {
            lbdem::k_grid_bounds<<<1, 1024, 0, st>>>(D.pt.p, nP, D.prm.nebrRange, D.flag.p, D.gridDesc.p, D.cellCount.p, nPdev);
            lbdem::k_grid_count<<<pb, 128, 0, st>>>(D.pt.p, nP, D.flag.p, D.gridDesc.p, D.cellCount.p, D.cellOf.p, nPdev);
            lbdem::k_grid_scan<<<1, 1024, 0, st>>>(D.flag.p, D.gridDesc.p, D.cellCount.p, D.cellFill.p);
            lbdem::k_grid_fill<<<pb, 128, 0, st>>>(nP, D.flag.p, D.cellCount.p, D.cellFill.p, D.cellOf.p, D.sorted.p, nPdev);
            lbdem::k_dem_neighbours_grid<<<pb, 128, 0, st>>>(D.pt.p, nP, D.prm.nebrRange, D.flag.p, D.gridDesc.p, D.cellCount.p, D.sorted.p, D.nbr.p, D.nNbr.p,
                                                             D.flag.p + 1, nPdev);
            h->launches += 4;
        } else {
            lbdem::k_dem_neighbours<<<pb, 128, 0, st>>>(D.pt.p, nP, D.prm.nebrRange, D.flag.p, D.nbr.p, D.nNbr.p, D.flag.p + 1, nPdev);
        }
        lbdem::k_dem_predict<<<nb, 128, 0, st>>>(D.e.p, n, D.prm, D.pt.p, D.walls.p, D.nWalls, D.flag.p);
        if (D.pbc) { lbdem::k_dem_ghost_update<true><<<(D.cap - n + 127) / 128, 128, 0, st>>>(D.pt.p, n, D.cnt.p); ++h->launches; }
        lbdem::k_dem_forces_correct<<<nb, 128, 0, st>>>(D.e.p, n, D.prm, D.pt.p, D.walls.p, hydro, D.nbr.p, D.nNbr.p);
        h->launches += 4;
    }
    if (D.pbc) {
        lbdem::k_dem_ghost_update<false><<<(D.cap - n + 127) / 128, 128, 0, st>>>(D.pt.p, n, D.cnt.p);
        lbdem::k_dem_export_pbc<<<pb, 128, 0, st>>>(D.e.p, n, D.pt.p, D.cnt.p, D.nComp.p, (lbdem::OutParticle*)h->rawParts.p, (lbdem::OutElement*)h->rawElmts.p);
        // the LB side needs the particle count (and whether to rescan) on the host: one small read-back per DEM step
        CU(cudaMemcpyAsync(h->pinnedStatus + 512, D.cnt.p, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        h->nParts = h->pinnedStatus[512];
        if (h->pinnedStatus[513]) { D.rescan = true; CU(cudaMemsetAsync(D.cnt.p + 1, 0, sizeof(uint32_t), st)); }
        if (h->nParts > D.cap) return fail(LBGPU_EUNSUPPORTED, "DEM: %u particles and ghosts exceed the %u slots", h->nParts, D.cap);
        k_prepare_particles<<<(std::max(h->nParts, n) + 127) / 128, 128, 0, st>>>(h->rawParts.p, h->nParts, h->rawElmts.p, n, h->uLength, h->uSpeed, h->parts.p, h->elmts.p);
        h->launches += 3;
    } else {
    lbdem::k_dem_export<<<nb, 128, 0, st>>>(D.e.p, n, D.pt.p, (lbdem::OutParticle*)h->rawParts.p, (lbdem::OutElement*)h->rawElmts.p, h->comps.p);
    k_prepare_particles<<<(std::max(nP, n) + 127) / 128, 128, 0, st>>>(h->rawParts.p, nP, h->rawElmts.p, n, h->uLength, h->uSpeed, h->parts.p, h->elmts.p);
    h->launches += 2;
    }
    CU(cudaGetLastError());
    return 0;
}
}  // namespace
extern "C" {

int lbGpuDemInit(LbGpuHandle* h, const LbGpuDemParams* prm, const LbGpuDemElement* elmts, uint32_t nElmts, const LbGpuDemWall* walls,
                 uint32_t nWalls) {
    if (!h || !prm || !elmts || nElmts == 0) return fail(LBGPU_EINVAL, "lbGpuDemInit: null argument or no elements");
    if (nWalls && !walls) return fail(LBGPU_EINVAL, "lbGpuDemInit: wall array missing");
    if (prm->multiStep < 1 || !(prm->deltat > 0.0) || !(prm->nebrRange > 0.0)) return fail(LBGPU_EINVAL, "lbGpuDemInit: multiStep, deltat and nebrRange must be positive");
    static_assert(sizeof(LbGpuDemWall) == sizeof(lbdem::Wall), "ABI layout");
    CU(cudaSetDevice(h->device));
    auto& D = h->dem;
    lbdem::Params& P = D.prm;
    P.contactModel = prm->contactModel; P.multiStep = prm->multiStep;
    P.knConst = prm->knConst; P.ksConst = prm->ksConst; P.dampCoeff = prm->dampCoeff; P.viscTang = prm->viscTang;
    P.linearStiff = prm->linearStiff; P.frictionCoefPart = prm->frictionCoefPart; P.frictionCoefWall = prm->frictionCoefWall;
    P.numVisc = prm->numVisc; P.deltat = prm->deltat; P.nebrRange = prm->nebrRange;
    for (int k = 0; k < 3; ++k) P.demF[k] = prm->demF[k];
    // DEM::predictor / DEM::corrector constants (DEM.cpp:1067-1112), in the reference's expressions
    const double dt = prm->deltat;
    const double c[5] = { dt, dt * dt / 2.0, dt * dt * dt / 6.0, dt * dt * dt * dt / 24.0, dt * dt * dt * dt * dt / 120.0 };
    const double g1[6] = { 95.0 / 288.0, 1.0, 25.0 / 24.0, 35.0 / 72.0, 5.0 / 48.0, 1.0 / 120.0 };
    const double g2[6] = { 3.0 / 16.0, 251.0 / 360.0, 1.0, 11.0 / 18.0, 1.0 / 6.0, 1.0 / 60.0 };
    for (int k = 0; k < 5; ++k) P.c[k] = c[k];
    P.coeff1[0] = g1[0] * c[0]; P.coeff2[0] = g2[0] * c[1];
    for (int k = 1; k < 6; ++k) { P.coeff1[k] = g1[k] * c[0] / c[k - 1]; P.coeff2[k] = g2[k] * c[1] / c[k - 1]; }
    // DEM::compositeProperties (DEM.cpp:404-433), unit = radius.  The triangle's "-1/2" is an integer division in the reference:
    // its second and third sphere sit at y = 0 (kept: the reference's inertia and contacts are those of that shape)
    memset(P.proto, 0, sizeof P.proto);
    {
        const double s2 = sqrt(2.0), s3 = sqrt(3.0), s6 = sqrt(6.0);
        const double p2[2][3] = { { 0.5, 0.0, 0.0 }, { -0.5, 0.0, 0.0 } };
        const double p3[3][3] = { { 0.0, 1.0, 0.0 }, { -s3 / 2, (double)(-1 / 2), 0.0 }, { s3 / 2, (double)(-1 / 2), 0.0 } };
        const double p4[4][3] = { { 0.0, 0.0, 1.0 }, { 0.0, 2.0 * s2 / 3.0, -1.0 / 3.0 }, { 2.0 * s6 / 6.0, -2.0 * s2 / 6.0, -1.0 / 3.0 },
                                  { -2.0 * s6 / 6.0, -2.0 * s2 / 6.0, -1.0 / 3.0 } };
        memcpy(P.proto[2], p2, sizeof p2); memcpy(P.proto[3], p3, sizeof p3); memcpy(P.proto[4], p4, sizeof p4);
    }
    P.nPbc = prm->nPbc; P.padPbc = 0;
    if (P.nPbc < 0 || P.nPbc > 3) return fail(LBGPU_EINVAL, "lbGpuDemInit: %d periodic boundaries (0-3)", P.nPbc);
    for (int b = 0; b < P.nPbc; ++b) {
        double nv = 0.0;
        for (int q = 0; q < 3; ++q) { P.pbcP[b][q] = prm->pbcP[b][q]; P.pbcV[b][q] = prm->pbcV[b][q]; nv += prm->pbcV[b][q] * prm->pbcV[b][q]; }
        nv = sqrt(nv);
        if (!(nv > 0.0)) return fail(LBGPU_EINVAL, "lbGpuDemInit: periodic boundary %d has no translation vector", b);
        for (int q = 0; q < 3; ++q) P.pbcN[b][q] = prm->pbcV[b][q] / nv;  // pbc::setPlanes
    }
    D.pbc = P.nPbc > 0;
    D.rescan = false;
    std::vector<lbdem::Elmt> E(nElmts);
    memset(E.data(), 0, sizeof(lbdem::Elmt) * nElmts);
    std::vector<lbdem::Part> PT;
    for (uint32_t k = 0; k < nElmts; ++k) {
        for (int q = 0; q < 3; ++q) {
            E[k].x[0][q] = E[k].xp[0][q] = elmts[k].x0[q];
            E[k].x[1][q] = E[k].xp[1][q] = elmts[k].x1[q];
            E[k].w[0][q] = elmts[k].w0[q];
            E[k].I[q] = elmts[k].I[q];
        }
        E[k].q[0][0] = E[k].qp[0][0] = 1.0;  // elmt::q0 = (1, 0, 0, 0), q1 = 0 (the particle file of a fresh run)
        E[k].radius = elmts[k].radius; E[k].m = elmts[k].m;
        E[k].size = elmts[k].size > 0 ? elmts[k].size : 1;
        E[k].pBegin = (int)PT.size();
        if (E[k].size > 4) return fail(LBGPU_EINVAL, "lbGpuDemInit: element %u has %d spheres (DEM::compositeProperties knows 1-4)", k, E[k].size);
        if (E[k].size > 1 && D.pbc) return fail(LBGPU_EUNSUPPORTED, "lbGpuDemInit: periodic DEM boundaries are covered for single spheres only (element %u has %d)", k, E[k].size);
        if (!(E[k].radius > 0.0) || !(E[k].m > 0.0)) return fail(LBGPU_EINVAL, "lbGpuDemInit: element %u has no radius or mass", k);
        for (int i = 0; i < E[k].size; ++i) {
            lbdem::Part a;
            memset(&a, 0, sizeof a);
            a.cluster = (int)k; a.proto = i; a.nearWall = -1;
            PT.push_back(a);
        }
    }
    const uint32_t nP = (uint32_t)PT.size();
    D.cap = D.pbc ? 7u * nP : nP;
    PT.resize(D.cap, PT.empty() ? lbdem::Part() : PT[0]);
    CU(D.e.alloc(nElmts)); CU(D.pt.alloc(D.cap)); CU(D.walls.alloc(nWalls ? nWalls : 1)); CU(D.nbr.alloc((size_t)D.cap * lbdem::MAX_NBR)); CU(D.nNbr.alloc(D.cap));
    if (D.pbc) {
        CU(D.gflag.alloc((size_t)3 * nElmts)); CU(D.basePos.alloc((size_t)3 * nElmts)); CU(D.cornPos.alloc((size_t)3 * nElmts)); CU(D.nComp.alloc(nElmts)); CU(D.cnt.alloc(4));
    }
    CU(D.flag.alloc(4)); CU(D.scal.alloc(2)); CU(D.hydro.alloc((size_t)7 * nElmts));
    D.grid = D.cap >= 4096;  // below that the all-pairs pass (one launch) is as fast as the five launches of the grid
    if (const char* e = getenv("LBGPU_DEM_GRID")) D.grid = atoi(e) != 0;
    if (D.grid) {
        CU(D.gridDesc.alloc(1)); CU(D.cellCount.alloc(lbdem::GRID_MAX_CELLS + 2)); CU(D.cellFill.alloc(lbdem::GRID_MAX_CELLS + 2));
        CU(D.cellOf.alloc(D.cap)); CU(D.sorted.alloc(D.cap));
    }
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaMemcpy(D.e.p, E.data(), sizeof(lbdem::Elmt) * nElmts, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(D.pt.p, PT.data(), sizeof(lbdem::Part) * D.cap, cudaMemcpyHostToDevice));
    if (nWalls) CU(cudaMemcpy(D.walls.p, walls, sizeof(lbdem::Wall) * nWalls, cudaMemcpyHostToDevice));
    CU(cudaMemset(D.nNbr.p, 0, sizeof(uint32_t) * D.cap));
    CU(cudaMemset(D.flag.p, 0, sizeof(uint32_t) * 4));
    const double sc[2] = { prm->maxDisp, 0.0 };
    CU(cudaMemcpy(D.scal.p, sc, sizeof sc, cudaMemcpyHostToDevice));
    CU(cudaDeviceSynchronize());  // the copies ran on the legacy stream (see lbGpuSetCurves)
    D.n = nElmts; D.nP = nP; D.nWalls = nWalls; D.on = true;
    // the resident lists of the coupling step: the particles of every element (+ the ghost slots: 7 components per sphere)
    if (int rc = particle_capacity(h, D.cap, nElmts, D.cap)) return rc;
    h->nParts = nP; h->nElmts = nElmts; h->nComps = D.cap;
    if (D.pbc) {  // until the first rebuild (the first sub-step): no ghosts
        std::vector<uint32_t> comps0((size_t)D.cap, 0u), one(nElmts, 1u);
        for (uint32_t k = 0; k < nElmts; ++k) comps0[(size_t)7 * k] = k;
        const uint32_t c0[4] = { nP, 0u, 0u, 0u };
        CU(cudaMemcpy(h->comps.p, comps0.data(), sizeof(uint32_t) * D.cap, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(D.nComp.p, one.data(), sizeof(uint32_t) * nElmts, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(D.cnt.p, c0, sizeof c0, cudaMemcpyHostToDevice));
        CU(cudaDeviceSynchronize());
    }
    const uint32_t nb = (nElmts + 127) / 128;
    lbdem::k_dem_init_particles<<<nb, 128, 0, h->stream>>>(D.e.p, nElmts, D.prm, D.pt.p);
    if (D.pbc) lbdem::k_dem_export_pbc<<<(D.cap + 127) / 128, 128, 0, h->stream>>>(D.e.p, nElmts, D.pt.p, D.cnt.p, D.nComp.p, (lbdem::OutParticle*)h->rawParts.p, (lbdem::OutElement*)h->rawElmts.p);
    else lbdem::k_dem_export<<<nb, 128, 0, h->stream>>>(D.e.p, nElmts, D.pt.p, (lbdem::OutParticle*)h->rawParts.p, (lbdem::OutElement*)h->rawElmts.p, h->comps.p);
    k_prepare_particles<<<(std::max(nP, nElmts) + 127) / 128, 128, 0, h->stream>>>(h->rawParts.p, nP, h->rawElmts.p, nElmts, h->uLength, h->uSpeed, h->parts.p, h->elmts.p);
    h->launches += 3;
    CU(cudaGetLastError());
    return LBGPU_OK;
}

int lbGpuDemStep(LbGpuHandle* h, const double* hydro) {
    if (!h || !h->dem.on) return fail(LBGPU_EINVAL, "lbGpuDemStep: no device-side DEM on this handle (lbGpuDemInit)");
    CU(cudaSetDevice(h->device));
    const double* src = h->lastStepCoupled ? h->slabs[0]->elemOut.p : nullptr;
    if (hydro) {
        CU(cudaStreamSynchronize(h->stream));
        CU(cudaMemcpy(h->dem.hydro.p, hydro, sizeof(double) * 7 * h->dem.n, cudaMemcpyHostToDevice));
        CU(cudaDeviceSynchronize());
        src = h->dem.hydro.p;
    }
    return dem_step(h, src);
}

int lbGpuRunDem(LbGpuHandle* h, int doFreeSurface, uint32_t count) {
    if (!h || !h->dem.on) return fail(LBGPU_EINVAL, "lbGpuRunDem: no device-side DEM on this handle (lbGpuDemInit)");
    CU(cudaSetDevice(h->device));
    CU(cudaEventRecord(h->evA, h->stream));
    h->kevCount = 0;
    auto cycle = [&]() -> int {
        int rc;
        phase_mark(h, 0);
        // (after the first LB step the element sums lie on the device; before it the reference's FHydro is zero too)
        if ((rc = dem_step(h, (h->lastStepCoupled || h->capturing) ? h->slabs[0]->elemOut.p : nullptr))) return rc;
        phase_mark(h, 1);
        if (doFreeSurface && h->fs) { if ((rc = free_surface_step(h))) return rc; } else { phase_mark(h, 2); phase_mark(h, 3); }
        // dem.newNeighborList: raised by a rebuild of the tables with periodic DEM boundaries (DEM.cpp:1414), reset by the coupling step
        const bool rescan = h->dem.rescan;
        h->dem.rescan = false;
        if ((rc = coupling_step(h, rescan))) return rc;
        phase_mark(h, 4);
        return lb_step(h);
    };
    if (int rc = run_cycles(h, count, h->fs && doFreeSurface, 1, cycle)) return rc;
    CU(cudaEventRecord(h->evB, h->stream));
    return LBGPU_OK;
}

int lbGpuDemState(LbGpuHandle* h, double* x0, double* x1, double* w0, double info[3]) {
    if (!h || !h->dem.on) return fail(LBGPU_EINVAL, "lbGpuDemState: no device-side DEM on this handle (lbGpuDemInit)");
    CU(cudaSetDevice(h->device));
    if (int rc = check_status(h)) return rc;
    auto& D = h->dem;
    std::vector<lbdem::Elmt> E(D.n);
    uint32_t f[3]; double sc[2];
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaMemcpy(E.data(), D.e.p, sizeof(lbdem::Elmt) * D.n, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(f, D.flag.p, sizeof f, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(sc, D.scal.p, sizeof sc, cudaMemcpyDeviceToHost));
    for (uint32_t k = 0; k < D.n; ++k)
        for (int q = 0; q < 3; ++q) {
            if (x0) x0[3 * k + q] = E[k].x[0][q];
            if (x1) x1[3 * k + q] = E[k].x[1][q];
            if (w0) w0[3 * k + q] = E[k].w[0][q];
        }
    if (info) { info[0] = sc[0]; info[1] = (double)f[2]; info[2] = (double)f[1]; }
    return LBGPU_OK;
}

int lbGpuDemParticles(LbGpuHandle* h, uint32_t* nParticles, double* x0, double* radiusVec, uint32_t* clusterIndex) {
    if (!h || !h->dem.on) return fail(LBGPU_EINVAL, "lbGpuDemParticles: no device-side DEM on this handle (lbGpuDemInit)");
    CU(cudaSetDevice(h->device));
    auto& D = h->dem;
    const uint32_t nNow = D.pbc ? h->nParts : D.nP;  // with periodic boundaries: the ghosts of the last rebuild included
    if (nParticles) *nParticles = nNow;
    if (!x0 && !radiusVec && !clusterIndex) return LBGPU_OK;
    std::vector<lbdem::Part> PT(nNow);
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaMemcpy(PT.data(), D.pt.p, sizeof(lbdem::Part) * nNow, cudaMemcpyDeviceToHost));
    for (uint32_t a = 0; a < nNow; ++a) {
        for (int q = 0; q < 3; ++q) { if (x0) x0[3 * a + q] = PT[a].xc[q]; if (radiusVec) radiusVec[3 * a + q] = PT[a].rvc[q]; }
        if (clusterIndex) clusterIndex[a] = (uint32_t)PT[a].cluster;
    }
    return LBGPU_OK;
}

int lbGpuDemContacts(LbGpuHandle* h, double* FParticle, double* FWall, double* MParticle, double* MWall) {
    if (!h || !h->dem.on) return fail(LBGPU_EINVAL, "lbGpuDemContacts: no device-side DEM on this handle (lbGpuDemInit)");
    CU(cudaSetDevice(h->device));
    auto& D = h->dem;
    std::vector<lbdem::Elmt> E(D.n);
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaMemcpy(E.data(), D.e.p, sizeof(lbdem::Elmt) * D.n, cudaMemcpyDeviceToHost));
    double* out[4] = { FParticle, FWall, MParticle, MWall };
    for (uint32_t k = 0; k < D.n; ++k)
        for (int a = 0; a < 4; ++a)
            if (out[a]) for (int q = 0; q < 3; ++q) out[a][3 * k + q] = E[k].fc[a][q];
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

int lbGpuPhaseTrace(LbGpuHandle* h, int on) {
    if (!h) return fail(LBGPU_EINVAL, "null handle");
    CU(cudaSetDevice(h->device));
    if (on && h->phaseEv.empty()) {
        h->phaseEv.resize((size_t)LbGpuHandle::PH_RING * LbGpuHandle::PH_MARKS);
        for (auto& e : h->phaseEv) CU(cudaEventCreate(&e));
    }
    h->phaseOn = on != 0;
    h->phaseCycles = 0;
    return LBGPU_OK;
}

int lbGpuPhaseMs(LbGpuHandle* h, float ms[7], uint32_t* cycles) {
    if (!h || !ms || !cycles) return fail(LBGPU_EINVAL, "null argument");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    const uint32_t n = (uint32_t)(h->phaseCycles < (uint64_t)LbGpuHandle::PH_RING ? h->phaseCycles : (uint64_t)LbGpuHandle::PH_RING);
    for (int p = 0; p < 7; ++p) ms[p] = 0.f;
    for (uint32_t c = 0; c < n; ++c)
        for (int p = 0; p < 7; ++p) {
            float t = 0.f;
            CU(cudaEventElapsedTime(&t, h->phaseEv[(size_t)c * LbGpuHandle::PH_MARKS + p], h->phaseEv[(size_t)c * LbGpuHandle::PH_MARKS + p + 1]));
            ms[p] += t;
        }
    for (int p = 0; p < 7; ++p) ms[p] = n ? ms[p] / (float)n : 0.f;
    *cycles = n;
    return LBGPU_OK;
}

int lbGpuGraphInfo(LbGpuHandle* h, uint64_t info[2]) {
    if (!h || !info) return fail(LBGPU_EINVAL, "null argument");
    info[0] = h->graph.captures; info[1] = h->graph.replays;
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
    Slab* s0 = h->slabs[0].get();
    const uint32_t nE = h->nElmts;
    if (nE && (FHydro || MHydro || fluidVolume)) {
        std::vector<double> tmp((size_t)7 * nE, 0.0);
        if (h->lastStepCoupled) {
            CU(cudaMemcpyAsync(tmp.data(), s0->elemOut.p, sizeof(double) * 7 * nE, cudaMemcpyDeviceToHost, h->stream));
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
            CU(cudaMemcpyAsync(tmp.data(), s0->sums.p + 5, sizeof(double) * 3 * nW, cudaMemcpyDeviceToHost, h->stream));
            CU(cudaStreamSynchronize(h->stream));
        }
        for (int k = 0; k < 3 * nW; ++k) wallFHydro[k] = tmp[k] * h->uForce;  // LB.cpp:1485-1487
    }
    return LBGPU_OK;
}

}  // extern "C"
namespace {
// n and the shifted u of the last step, recomputed from the previous population buffer when the step kernel did not
// store them (lattices whose streaming never reads them)
int ensure_macro(LbGpuHandle* h) {
    if (h->macroValid) return 0;
    cudaStream_t st = h->stream;
    for (auto& sp : h->slabs) {
        Dev dm = dev_for(h, sp.get());
        set_src(dm, sp.get(), h->cur ^ 1, !h->lastStepFirst);
        const bool force = h->force || h->lastStepCoupled, couple = h->lastStepCoupled;
        const uint32_t B = own_blocks(sp.get());
        if (couple) k_macro<true, true><<<B, BLOCK, 0, st>>>(dm);
        else if (force) k_macro<true, false><<<B, BLOCK, 0, st>>>(dm);
        else k_macro<false, false><<<B, BLOCK, 0, st>>>(dm);
        ++h->launches;
    }
    h->macroValid = true;
    CU(cudaGetLastError());
    return 0;
}
}  // namespace
extern "C" {

int lbGpuFluidSummary(LbGpuHandle* h, double out[4]) {
    if (!h || !out) return fail(LBGPU_EINVAL, "lbGpuFluidSummary: null argument");
    CU(cudaSetDevice(h->device));
    if (int rc = check_status(h)) return rc;
    if (int rc = ensure_macro(h)) return rc;
    cudaStream_t st = h->stream;
    double tot[4] = { 0.0, 0.0, 0.0, 0.0 };
    for (auto& sp : h->slabs) {
        Slab* s = sp.get();
        const uint32_t B = own_blocks(s);
        if (s->summary.n < (size_t)4 * B + 4) CU(s->summary.alloc((size_t)4 * B + 4));
        k_summary<<<B, BLOCK, 0, st>>>(dev_for(h, s), s->summary.p + 4, B);
        k_summary_final<<<1, 1024, 0, st>>>(s->summary.p + 4, B, B, s->summary.p);
        h->launches += 2;
        double r[4];
        CU(cudaMemcpyAsync(r, s->summary.p, sizeof r, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        tot[0] = r[0] > tot[0] ? r[0] : tot[0];
        for (int k = 1; k < 4; ++k) tot[k] += r[k];
    }
    if (lbcomm::active()) {
        // maximum and sums over the ranks
        Slab* s0 = h->slabs[0].get();
        CU(cudaMemcpyAsync(s0->summary.p, tot, sizeof tot, cudaMemcpyHostToDevice, st));
        NC(lbcomm::api().AllReduce(s0->summary.p, s0->summary.p, 1, lbcomm::ncclFloat64, lbcomm::ncclMax, lbcomm::comm().comm, st));
        NC(lbcomm::api().AllReduce(s0->summary.p + 1, s0->summary.p + 1, 3, lbcomm::ncclFloat64, lbcomm::ncclSum, lbcomm::comm().comm, st));
        CU(cudaMemcpyAsync(tot, s0->summary.p, sizeof tot, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    for (int k = 0; k < 4; ++k) out[k] = tot[k];
    return LBGPU_OK;
}

// IO::exportParaviewFluidOld (IO.cpp:698-831): the same ImageData file -- extents, origin, spacing, array names, types,
// order and values -- with the data as raw appended binary instead of 17 M formatted numbers per array.  One field at
// a time passes through lbGpuFetchFields (which reproduces what the reference holds in dead shell cells), so the
// host never holds more than the largest array.
int lbGpuWriteVti(LbGpuHandle* h, const char* path, int withSolidIndex) {
    if (!h || !path) return fail(LBGPU_EINVAL, "lbGpuWriteVti: null argument");
    if (lbcomm::active()) return fail(LBGPU_EUNSUPPORTED, "lbGpuWriteVti: one process only (a rank holds its own planes)");
    const LbGpuParams& prm = h->prm;
    const size_t N = (size_t)prm.size[0] * prm.size[1] * prm.size[2];
    if (N * 24 >= (1ull << 32)) return fail(LBGPU_EUNSUPPORTED, "lbGpuWriteVti: arrays of more than 4 GB need 64-bit block headers");
    const double L = prm.unitLength, T = prm.unitTime, D = prm.unitDensity;
    const double uSpeed = L / T, uPressure = D * L * L / T / T, uDynVisc = D * L * L / T, uDensity = D;  // node.cpp:476-488
    struct Arr { const char* name; const char* type; int comps; size_t bytes; };
    std::vector<Arr> arrs;
    arrs.push_back({ "type", "Int8", 1, N });
    arrs.push_back({ "v", "Float64", 3, N * 24 });
    arrs.push_back({ "pressure", "Float64", 1, N * 8 });
    if (prm.nonNewtonian) arrs.push_back({ "dynVisc", "Float64", 1, N * 8 });
    if (prm.freeSurface) arrs.push_back({ "AAAmass", "Float64", 1, N * 8 });
    if (withSolidIndex) arrs.push_back({ "solidIndex", "Int16", 1, N * 2 });
    FILE* fp = fopen(path, "wb");
    if (!fp) return fail(LBGPU_EINVAL, "lbGpuWriteVti: cannot open %s", path);
    int rc = LBGPU_OK;
    auto body = [&]() -> int {
        fprintf(fp, "<VTKFile type=\"ImageData\" version=\"0.1\" byte_order=\"LittleEndian\">\n");
        fprintf(fp, " <ImageData WholeExtent=\"0 %d 0 %d 0 %d\" Origin=\"0.0 0.0 0.0\" Spacing=\"%g %g %g\">\n", prm.size[0] - 1,
                prm.size[1] - 1, prm.size[2] - 1, L, L, L);
        fprintf(fp, "  <Piece Extent=\"0 %d 0 %d 0 %d\">\n   <PointData>\n", prm.size[0] - 1, prm.size[1] - 1, prm.size[2] - 1);
        size_t off = 0;
        for (const Arr& a : arrs) {
            if (a.comps > 1) fprintf(fp, "    <DataArray type=\"%s\" Name=\"%s\" NumberOfComponents=\"%d\" format=\"appended\" offset=\"%zu\"/>\n", a.type, a.name, a.comps, off);
            else fprintf(fp, "    <DataArray type=\"%s\" Name=\"%s\" format=\"appended\" offset=\"%zu\"/>\n", a.type, a.name, off);
            off += 4 + a.bytes;
        }
        fprintf(fp, "   </PointData>\n   <CellData>\n   </CellData>\n  </Piece>\n </ImageData>\n <AppendedData encoding=\"raw\">\n_");
        std::vector<uint8_t> tf(N);
        std::vector<double> buf;
        auto block = [&](const void* data, size_t bytes) -> int {
            const uint32_t nb = (uint32_t)bytes;
            if (fwrite(&nb, 4, 1, fp) != 1 || fwrite(data, 1, bytes, fp) != bytes) return fail(LBGPU_EINVAL, "lbGpuWriteVti: write to %s failed", path);
            return 0;
        };
        int r;
        // type: nodeType::getType, 1 inside particles (IO.cpp:726-733)
        if ((r = lbGpuFetchFields(h, tf.data(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr))) return r;
        {
            std::vector<int8_t> t8(N);
            for (size_t i = 0; i < N; ++i) t8[i] = (tf[i] & LBGPU_P_BIT) ? (int8_t)1 : (int8_t)(tf[i] & LBGPU_TYPE_MASK);
            if ((r = block(t8.data(), N))) return r;
        }
        // v = u * unit.Speed where a node exists (IO.cpp:746-754): lbGpuFetchFields returns zeros elsewhere
        buf.resize(3 * N);
        if ((r = lbGpuFetchFields(h, nullptr, nullptr, nullptr, buf.data(), nullptr, nullptr, nullptr, nullptr, nullptr))) return r;
        for (size_t k = 0; k < 3 * N; ++k) buf[k] *= uSpeed;
        if ((r = block(buf.data(), N * 24))) return r;
        // pressure = 0.3333333 (n - 1) unit.Pressure where a node exists and n != 0 (IO.cpp:768-781)
        buf.resize(N);
        if ((r = lbGpuFetchFields(h, nullptr, nullptr, buf.data(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr))) return r;
        for (size_t i = 0; i < N; ++i)
            if ((tf[i] & LBGPU_NODE_BIT) && buf[i] != 0.0) buf[i] = 0.3333333 * (buf[i] - 1.0) * uPressure;
        if ((r = block(buf.data(), N * 8))) return r;
        if (prm.nonNewtonian) {
            if ((r = lbGpuFetchFields(h, nullptr, nullptr, nullptr, nullptr, nullptr, buf.data(), nullptr, nullptr, nullptr))) return r;
            for (size_t i = 0; i < N; ++i) buf[i] *= uDynVisc;
            if ((r = block(buf.data(), N * 8))) return r;
        }
        if (prm.freeSurface) {
            if ((r = lbGpuFetchFields(h, nullptr, nullptr, nullptr, nullptr, buf.data(), nullptr, nullptr, nullptr, nullptr))) return r;
            for (size_t i = 0; i < N; ++i) buf[i] *= uDensity;
            if ((r = block(buf.data(), N * 8))) return r;
        }
        if (withSolidIndex) {
            std::vector<uint32_t> si(N);
            if ((r = lbGpuFetchFields(h, nullptr, si.data(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr))) return r;
            std::vector<int16_t> s16(N);
            for (size_t i = 0; i < N; ++i) s16[i] = (int16_t)si[i];
            if ((r = block(s16.data(), N * 2))) return r;
        }
        fprintf(fp, "\n </AppendedData>\n</VTKFile>\n");
        return 0;
    };
    rc = body();
    if (fclose(fp) != 0 && rc == 0) rc = fail(LBGPU_EINVAL, "lbGpuWriteVti: closing %s failed", path);
    return rc;
}

int lbGpuFetchFields(LbGpuHandle* h, uint8_t* type_flags, uint32_t* solidIndex, double* n, double* u, double* mass,
                     double* visc, double* shearRate, double* hydroForce, double* f) {
    if (!h) return fail(LBGPU_EINVAL, "null handle");
    CU(cudaSetDevice(h->device));
    if (int rc = check_status(h)) return rc;
    Trace tr("lbGpuFetchFields");
    cudaStream_t st = h->stream;
    const int zLoHost = h->slabs[0]->zBegin - 1;
    if (n || u) { if (int rc = ensure_macro(h)) return rc; }
    for (auto& sp : h->slabs) {
        Slab* s = sp.get();
        const uint32_t B = s->blocks;
        // every slab writes the planes it owns (plus the true shell planes at the ends of the lattice)
        const uint32_t p0 = s->remoteLo ? 1u : 0u, p1 = (uint32_t)s->dev.Z - (s->remoteHi ? 1u : 0u);
        const size_t cOff = (size_t)p0 * s->XY, cCnt = (size_t)(p1 - p0) * s->XY;
        const size_t hBase = (size_t)(s->zBegin - 1 - zLoHost) * s->XY;  // host index of the slab's local cell 0
        const size_t hOff = hBase + cOff;
        Dev d = dev_all(h, s);
        // periodic shell planes that exist only as remote ghosts on the device: dead cells of the reference
        const size_t hTop = hBase + (size_t)(s->dev.Z - 1) * s->XY;
        auto shell_zero = [&](double* dst, int comps) {
            if (!s->shellTypeLo.empty()) memset(dst + comps * hBase, 0, sizeof(double) * comps * s->XY);
            if (!s->shellTypeHi.empty()) memset(dst + comps * hTop, 0, sizeof(double) * comps * s->XY);
        };
        if (type_flags) {
            DevBuf<uint8_t> tmp;
            CU(tmp.alloc(s->N));
            k_fetch_types<<<B, BLOCK, 0, st>>>(d, tmp.p);
            if (int rc2 = d2h_staged(h, type_flags + hOff, tmp.p + cOff, cCnt)) return rc2;
            CU(cudaStreamSynchronize(st));
            for (uint32_t k = 0; k < s->nGhost; ++k) type_flags[hBase + s->ghostIdx[k]] = s->ghostType[k];
            if (!s->shellTypeLo.empty()) memcpy(type_flags + hBase, s->shellTypeLo.data(), s->XY);
            if (!s->shellTypeHi.empty()) memcpy(type_flags + hTop, s->shellTypeHi.data(), s->XY);
            tr.mark("types");
        }
        if (solidIndex) {
            if (int rc2 = d2h_staged(h, solidIndex + hOff, s->solidIndex.p + cOff, sizeof(uint32_t) * cCnt)) return rc2;
            CU(cudaStreamSynchronize(st));
            for (uint32_t k = 0; k < s->nGhost; ++k) solidIndex[hBase + s->ghostIdx[k]] = s->ghostSolid[k];
            if (!s->shellSolidLo.empty()) memcpy(solidIndex + hBase, s->shellSolidLo.data(), sizeof(uint32_t) * s->XY);
            if (!s->shellSolidHi.empty()) memcpy(solidIndex + hTop, s->shellSolidHi.data(), sizeof(uint32_t) * s->XY);
        }
        DevBuf<double> tmp;
        if (n || u || mass || visc || shearRate || hydroForce) CU(tmp.alloc((size_t)3 * s->N));
        auto scalar = [&](const double* src, double* dst, int activeOnly) -> int {
            k_fetch_scalar<<<B, BLOCK, 0, st>>>(d, src, tmp.p, activeOnly);
            if (int rc2 = d2h_staged(h, dst + hOff, tmp.p + cOff, sizeof(double) * cCnt)) return rc2;
            CU(cudaStreamSynchronize(st));
            shell_zero(dst, 1);
            return 0;
        };
        int rc;
        if (n && (rc = scalar(s->n.p, n, 0))) return rc;
        if (n) {
            // a dead cell of the reference (periodic shell) that carries a node is a wall node: density initDensity = 1
            // (LB.cpp:946-996; only cell 0 can be one, through d[0] = 0 of the interior cells)
            for (uint32_t k = 0; k < s->nGhost; ++k)
                if ((s->ghostType[k] & NODE_BIT) && is_wall_type(s->ghostType[k] & TYPE_MASK)) n[hBase + s->ghostIdx[k]] = s->ghostN[k];
        }
        tr.mark("n");
        if (mass && (rc = scalar(s->mass.p, mass, 0))) return rc;
        tr.mark("mass");
        if (visc && (rc = scalar(s->visc.p, visc, 0))) return rc;
        if (shearRate && (rc = scalar(s->shearRate.p, shearRate, 1))) return rc;
        if (u) {
            k_fetch_vec<<<B, BLOCK, 0, st>>>(d, s->ux.p, s->uy.p, s->uz.p, tmp.p, 0);
            if (int rc2 = d2h_staged(h, u + 3 * hOff, tmp.p + 3 * cOff, sizeof(double) * 3 * cCnt)) return rc2;
            CU(cudaStreamSynchronize(st));
            shell_zero(u, 3);
            for (uint32_t k = 0; k < s->nGhost; ++k)
                if ((s->ghostType[k] & NODE_BIT) && is_wall_type(s->ghostType[k] & TYPE_MASK))
                    for (int c = 0; c < 3; ++c) u[3 * (hBase + s->ghostIdx[k]) + c] = s->ghostU[3 * k + c];
            tr.mark("u");
        }
        if (hydroForce) {
            k_fetch_vec<<<B, BLOCK, 0, st>>>(d, s->hfx.p, s->hfy.p, s->hfz.p, tmp.p, 1);
            if (int rc2 = d2h_staged(h, hydroForce + 3 * hOff, tmp.p + 3 * cOff, sizeof(double) * 3 * cCnt)) return rc2;
            CU(cudaStreamSynchronize(st));
            shell_zero(hydroForce, 3);
        }
        if (f) {
            DevBuf<double> tf;
            CU(tf.alloc((size_t)Q * s->N));
            k_download_f<<<B, BLOCK, 0, st>>>(d, s->fbuf(h->cur), tf.p);
            if (int rc2 = d2h_staged(h, f + (size_t)Q * hOff, tf.p + (size_t)Q * cOff, sizeof(double) * Q * cCnt)) return rc2;
            CU(cudaStreamSynchronize(st));
            shell_zero(f, Q);
        }
    }
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    return LBGPU_OK;
}

}  // extern "C"
namespace { int counts_impl(LbGpuHandle* h, uint64_t counts[4], bool global); }
extern "C" {
int lbGpuCounts(LbGpuHandle* h, uint64_t counts[4]) { return counts_impl(h, counts, true); }
int lbGpuCountsLocal(LbGpuHandle* h, uint64_t counts[4]) { return counts_impl(h, counts, false); }
}  // extern "C"
namespace {
int counts_impl(LbGpuHandle* h, uint64_t counts[4], bool global) {
    if (!h || !counts) return fail(LBGPU_EINVAL, "null argument");
    CU(cudaSetDevice(h->device));
    unsigned long long tot[3] = { 0, 0, 0 };
    for (auto& sp : h->slabs) {
        Slab* s = sp.get();
        CU(cudaMemsetAsync(s->counters.p + 1, 0, sizeof(unsigned long long) * 3, h->stream));
        k_count<<<own_blocks(s), BLOCK, 0, h->stream>>>(dev_for(h, s), s->counters.p + 1);
        ++h->launches;
        unsigned long long tmp[3];
        CU(cudaMemcpyAsync(tmp, s->counters.p + 1, sizeof tmp, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        for (int k = 0; k < 3; ++k) tot[k] += tmp[k];
    }
    if (global && lbcomm::active()) {
        Slab* s0 = h->slabs[0].get();
        CU(cudaMemcpyAsync(s0->counters.p + 4, tot, sizeof tot, cudaMemcpyHostToDevice, h->stream));
        if (int rc = allreduce_sum(h, s0->counters.p + 4, 3, lbcomm::ncclUint64)) return rc;
        CU(cudaMemcpyAsync(tot, s0->counters.p + 4, sizeof tot, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    counts[0] = tot[0]; counts[1] = tot[1]; counts[2] = tot[2]; counts[3] = h->steps;
    return LBGPU_OK;
}
}  // namespace
extern "C" {

int lbGpuCommUniqueId(uint8_t id[128]) {
    if (!id) return fail(LBGPU_EINVAL, "null argument");
    const std::string e = lbcomm::load();
    if (!e.empty()) return fail(LBGPU_ECOMM, "lbGpuCommUniqueId: %s", e.c_str());
    lbcomm::ncclUniqueId u;
    NC(lbcomm::api().GetUniqueId(&u));
    memcpy(id, u.internal, 128);
    return LBGPU_OK;
}

int lbGpuCommInit(const uint8_t id[128], int32_t rank, int32_t world, int32_t device) {
    if (!id || world < 1 || rank < 0 || rank >= world) return fail(LBGPU_EINVAL, "lbGpuCommInit: bad arguments");
    lbcomm::Comm& c = lbcomm::comm();
    if (c.comm) return fail(LBGPU_EINVAL, "lbGpuCommInit: a communicator is already active");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(LBGPU_ENODEVICE, "lbGpuCommInit: no CUDA device"); }
    if (device >= ndev) return fail(LBGPU_EINVAL, "lbGpuCommInit: device %d of %d", device, ndev);
    if (device >= 0) CU(cudaSetDevice(device));
    CU(cudaGetDevice(&c.device));
    c.rank = rank; c.world = world;
    if (world == 1) return LBGPU_OK;  // nothing to talk to
    const std::string e = lbcomm::load();
    if (!e.empty()) return fail(LBGPU_ECOMM, "lbGpuCommInit: %s", e.c_str());
    lbcomm::ncclUniqueId u;
    memcpy(u.internal, id, 128);
    NC(lbcomm::api().CommInitRank(&c.comm, world, u, rank));
    {   // NCCL connects its rings / trees on the first collective (seconds at 8 ranks): pay that here, once per
        // process, rather than inside the first lattice's initialisation
        double* w = nullptr;
        CU(cudaMalloc((void**)&w, sizeof(double)));
        CU(cudaMemset(w, 0, sizeof(double)));
        NC(lbcomm::api().AllReduce(w, w, 1, lbcomm::ncclFloat64, lbcomm::ncclSum, c.comm, (cudaStream_t)0));
        CU(cudaDeviceSynchronize());
        // ... and so is the first mapping of a neighbour's memory (peer access between the two devices is enabled then):
        // every rank exports a dummy block, the handles are all-gathered, the ring neighbours' blocks are opened and closed
        if (lbcomm::api().AllGather) {
            cudaIpcMemHandle_t mine;
            char* dh = nullptr;
            if (cudaIpcGetMemHandle(&mine, w) == cudaSuccess && cudaMalloc((void**)&dh, sizeof mine * (size_t)(world + 1)) == cudaSuccess) {
                std::vector<cudaIpcMemHandle_t> all((size_t)world);
                CU(cudaMemcpy(dh, &mine, sizeof mine, cudaMemcpyHostToDevice));
                NC(lbcomm::api().AllGather(dh, dh + sizeof mine, sizeof mine, lbcomm::ncclUint8, c.comm, (cudaStream_t)0));
                CU(cudaMemcpy(all.data(), dh + sizeof mine, sizeof mine * (size_t)world, cudaMemcpyDeviceToHost));
                const int nbs[2] = { (rank + world - 1) % world, (rank + 1) % world };
                for (int k = 0; k < (world > 2 ? 2 : 1); ++k) {
                    void* base = nullptr;
                    if (nbs[k] != rank && cudaIpcOpenMemHandle(&base, all[(size_t)nbs[k]], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess) cudaIpcCloseMemHandle(base);
                }
                cudaGetLastError();
                // nobody frees its block while a neighbour may still be opening it
                NC(lbcomm::api().AllReduce(w, w, 1, lbcomm::ncclFloat64, lbcomm::ncclSum, c.comm, (cudaStream_t)0));
                CU(cudaDeviceSynchronize());
                cudaFree(dh);
            } else cudaGetLastError();
        }
        CU(cudaFree(w));
    }
    return LBGPU_OK;
}

int lbGpuCommInfo(int32_t* rank, int32_t* world, int32_t* ncclVersion) {
    const lbcomm::Comm& c = lbcomm::comm();
    if (rank) *rank = c.rank;
    if (world) *world = c.comm ? c.world : 1;
    if (ncclVersion) { int v = 0; if (lbcomm::api().GetVersion) lbcomm::api().GetVersion(&v); *ncclVersion = v; }
    return LBGPU_OK;
}

int lbGpuPeerHalo(LbGpuHandle* h, int32_t* on) {
    if (!h || !on) return fail(LBGPU_EINVAL, "null argument");
    *on = h->peer.on ? 1 : 0;
    if (!h->peer.on && !h->peer.note.empty()) g_err = h->peer.note;  // why not (lbGpuLastError)
    return LBGPU_OK;
}

int lbGpuCommFinalize(void) {
    lbcomm::Comm& c = lbcomm::comm();
    if (c.comm) { lbcomm::api().CommDestroy(c.comm); c.comm = nullptr; }
    c.rank = 0; c.world = 1; c.device = -1;
    return LBGPU_OK;
}

int lbGpuSelfTest(uint64_t count, uint64_t seed, uint64_t result[3]) {
    if (!result) return fail(LBGPU_EINVAL, "null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(LBGPU_ENODEVICE, "lbGpuSelfTest: no CUDA device"); }
    DevBuf<unsigned long long> out;
    CU(out.alloc(3));
    CU(cudaMemset(out.p, 0, 3 * sizeof(unsigned long long)));
    k_selftest_div<<<148 * 8, 256>>>(count, seed, out.p);
    CU(cudaGetLastError());
    unsigned long long r[3];
    CU(cudaMemcpy(r, out.p, sizeof r, cudaMemcpyDeviceToHost));
    result[0] = r[0]; result[1] = r[1]; result[2] = r[2];
    return LBGPU_OK;
}

// ---------------------------------------------------------------------------------------------
// Checkpoint / restart (the reference has none for the fluid, IO.cpp:486-535 covers the particles).
// The blob is the dynamic device state of the handle in a fixed order behind a small header; everything derived from
// the static geometry (bulk bitmap, static list, ghost lists, curves) is rebuilt by lbGpuInit / lbGpuSetCurves on the
// handle the state is loaded into, the per-step lists are rebuilt by the next step.
// ---------------------------------------------------------------------------------------------
}  // extern "C"
namespace {
struct StateHeader {
    char magic[8];
    uint32_t version, nSlabs;
    int32_t size[3], fs;
    uint64_t cellsTotal, steps;
    uint32_t cur, macroValid, lastStepFirst, lastStepCoupled, nParts, nElmts, nComps, pad;
};

// visits every piece of dynamic state in a fixed order: fn(device pointer, bytes)
template <class F>
int visit_state(LbGpuHandle* h, F&& fn) {
    int rc;
    for (auto& sp : h->slabs) {
        Slab* s = sp.get();
        auto buf = [&](auto& b) -> int { return b.n ? fn((void*)b.p, b.n * sizeof(*b.p)) : 0; };
        if ((rc = buf(s->fA)) || (rc = buf(s->fB)) || (rc = buf(s->n)) || (rc = buf(s->ux)) || (rc = buf(s->uy)) || (rc = buf(s->uz)) ||
            (rc = buf(s->mass)) || (rc = buf(s->newMass)) || (rc = buf(s->visc)) || (rc = buf(s->shearRate)) || (rc = buf(s->hfx)) ||
            (rc = buf(s->hfy)) || (rc = buf(s->hfz)) || (rc = buf(s->type0)) || (rc = buf(s->type1)) || (rc = buf(s->mark)) ||
            (rc = buf(s->solidIndex)) || (rc = buf(s->counters)) || (rc = buf(s->scal)) || (rc = buf(s->sums)) || (rc = buf(s->status)))
            return rc;
        if (h->nElmts && (rc = fn((void*)s->elemOut.p, sizeof(double) * 7 * h->nElmts))) return rc;
    }
    if (h->nParts && (rc = fn((void*)h->parts.p, sizeof(lb::Particle) * h->nParts))) return rc;
    if (h->nElmts && (rc = fn((void*)h->elmts.p, sizeof(lb::Element) * h->nElmts))) return rc;
    if (h->nComps && (rc = fn((void*)h->comps.p, sizeof(uint32_t) * h->nComps))) return rc;
    if (h->dem.on) {  // the elements' Gear state and tables (the handle the blob is loaded into went through lbGpuDemInit)
        auto& D = h->dem;
        if ((rc = fn((void*)D.e.p, sizeof(lbdem::Elmt) * D.n)) || (rc = fn((void*)D.pt.p, sizeof(lbdem::Part) * D.cap)) ||
            (rc = fn((void*)D.nbr.p, sizeof(uint32_t) * D.nbr.n)) || (rc = fn((void*)D.nNbr.p, sizeof(uint32_t) * D.cap)) || (rc = fn((void*)D.flag.p, sizeof(uint32_t) * 4)) ||
            (rc = fn((void*)D.scal.p, sizeof(double) * 2)))
            return rc;
        if (D.pbc && ((rc = fn((void*)D.cnt.p, sizeof(uint32_t) * 4)) || (rc = fn((void*)D.nComp.p, sizeof(uint32_t) * D.n)))) return rc;  // ghosts of the last rebuild
    }
    return 0;
}
}  // namespace
extern "C" {

int lbGpuStateBytes(LbGpuHandle* h, uint64_t* bytes) {
    if (!h || !bytes) return fail(LBGPU_EINVAL, "lbGpuStateBytes: null argument");
    uint64_t total = sizeof(StateHeader);
    visit_state(h, [&](void*, size_t b) { total += b; return 0; });
    *bytes = total;
    return LBGPU_OK;
}

int lbGpuSaveState(LbGpuHandle* h, void* buffer, uint64_t bytes) {
    if (!h || !buffer) return fail(LBGPU_EINVAL, "lbGpuSaveState: null argument");
    uint64_t need = 0;
    lbGpuStateBytes(h, &need);
    if (bytes < need) return fail(LBGPU_EINVAL, "lbGpuSaveState: buffer of %llu bytes, %llu needed", (unsigned long long)bytes, (unsigned long long)need);
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    if (int rc = check_status(h)) return rc;
    StateHeader hd;
    memset(&hd, 0, sizeof hd);
    memcpy(hd.magic, "LBGPUST", 8);
    hd.version = 1; hd.nSlabs = (uint32_t)h->slabs.size();
    for (int k = 0; k < 3; ++k) hd.size[k] = h->prm.size[k];
    hd.fs = h->fs ? 1 : 0;
    for (auto& sp : h->slabs) hd.cellsTotal += sp->N;
    hd.steps = h->steps; hd.cur = (uint32_t)h->cur; hd.macroValid = h->macroValid; hd.lastStepFirst = h->lastStepFirst;
    hd.lastStepCoupled = h->lastStepCoupled; hd.nParts = h->nParts; hd.nElmts = h->nElmts; hd.nComps = h->nComps;
    hd.pad = h->dem.on ? h->dem.n : 0;  // elements of the device-side DEM
    char* out = (char*)buffer;
    memcpy(out, &hd, sizeof hd);
    size_t off = sizeof hd;
    return visit_state(h, [&](void* p, size_t b) -> int {
        CU(cudaMemcpy(out + off, p, b, cudaMemcpyDeviceToHost));
        off += b;
        return 0;
    });
}

int lbGpuLoadState(LbGpuHandle* h, const void* buffer, uint64_t bytes) {
    if (!h || !buffer || bytes < sizeof(StateHeader)) return fail(LBGPU_EINVAL, "lbGpuLoadState: bad argument");
    StateHeader hd;
    memcpy(&hd, buffer, sizeof hd);
    uint64_t cells = 0;
    for (auto& sp : h->slabs) cells += sp->N;
    if (memcmp(hd.magic, "LBGPUST", 8) != 0 || hd.version != 1) return fail(LBGPU_EINVAL, "lbGpuLoadState: not a state blob of this library");
    if (hd.nSlabs != h->slabs.size() || hd.cellsTotal != cells || hd.size[0] != h->prm.size[0] || hd.size[1] != h->prm.size[1] ||
        hd.size[2] != h->prm.size[2] || (hd.fs != 0) != h->fs)
        return fail(LBGPU_EINVAL, "lbGpuLoadState: the state was saved from a different lattice (size / slabs / free surface)");
    if (hd.pad != (h->dem.on ? h->dem.n : 0u))
        return fail(LBGPU_EINVAL, "lbGpuLoadState: the state holds %u device-side DEM elements, this handle %u (lbGpuDemInit first)", hd.pad, h->dem.on ? h->dem.n : 0u);
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    // room for the particle lists of the saved state (as upload_particles provides it)
    if (hd.nParts > h->rawParts.n) { CU(h->rawParts.alloc(hd.nParts + 16)); CU(h->parts.alloc(h->rawParts.n)); }
    if (hd.nElmts > h->rawElmts.n) {
        CU(h->rawElmts.alloc(hd.nElmts + 16));
        CU(h->elmts.alloc(h->rawElmts.n));
        for (auto& s : h->slabs) CU(s->elemOut.alloc(h->rawElmts.n * 7));
    }
    if (hd.nComps > h->comps.n) CU(h->comps.alloc(hd.nComps + 16));
    h->nParts = hd.nParts; h->nElmts = hd.nElmts; h->nComps = hd.nComps;
    uint64_t need = 0;
    lbGpuStateBytes(h, &need);
    if (bytes < need) return fail(LBGPU_EINVAL, "lbGpuLoadState: blob of %llu bytes, %llu expected", (unsigned long long)bytes, (unsigned long long)need);
    const char* in = (const char*)buffer;
    size_t off = sizeof hd;
    if (int rc = visit_state(h, [&](void* p, size_t b) -> int {
            CU(cudaMemcpy(p, in + off, b, cudaMemcpyHostToDevice));
            off += b;
            return 0;
        }))
        return rc;
    if (h->graph.exec) { cudaGraphExecDestroy(h->graph.exec); h->graph.exec = nullptr; }
    h->steps = hd.steps; h->cur = (int)hd.cur; h->macroValid = hd.macroValid != 0; h->lastStepFirst = hd.lastStepFirst != 0;
    h->lastStepCoupled = hd.lastStepCoupled != 0;
    h->typesFlipped = false; h->listsFresh = false;
    CU(cudaDeviceSynchronize());  // (see lbGpuSetCurves)
    if (h->wallPush && h->steps > 0) wall_push(h, h->cur);
    return LBGPU_OK;
}

int lbGpuFinalize(LbGpuHandle* h) {
    if (!h) return LBGPU_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->commStream) cudaStreamSynchronize(h->commStream);
    if (h->graph.exec) cudaGraphExecDestroy(h->graph.exec);
    for (cudaEvent_t e : h->phaseEv) if (e) cudaEventDestroy(e);
    peer_teardown(h);
    h->slabs.clear();
    if (h->evA) cudaEventDestroy(h->evA);
    if (h->evB) cudaEventDestroy(h->evB);
    if (h->evFaces) cudaEventDestroy(h->evFaces);
    if (h->evHalo) cudaEventDestroy(h->evHalo);
    if (h->commStream) { cudaStreamSynchronize(h->commStream); cudaStreamDestroy(h->commStream); }
    for (cudaEvent_t e : h->kev0) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : h->kev1) if (e) cudaEventDestroy(e);
    if (h->pinned) cudaFreeHost(h->pinned);
    for (int k = 0; k < 2; ++k) { if (h->stage[k]) stage_pool().put(h->stage[k]); if (h->stageEv[k]) cudaEventDestroy(h->stageEv[k]); }
    if (h->pinnedStatus) cudaFreeHost(h->pinnedStatus);
    if (h->pinnedCounts) cudaFreeHost(h->pinnedCounts);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return LBGPU_OK;
}

}  // extern "C"
