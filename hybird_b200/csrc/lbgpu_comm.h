// lbgpu_comm.h -- halo transport between slabs that live in different processes (one process per GPU).
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy PyTorch already loaded when the host is a
// torch.distributed job, else the system library), so liblbgpu.so has no link-time dependency on it and a
// single-GPU user never needs it.  Only point-to-point ncclSend/ncclRecv inside one group per exchange (NVLink 5 /
// NVSwitch peer copies on a B200 box) and small fp64/u64/u32 sum all-reduces are used; there is no data-path
// collective (the path shards, the only exchange is the face halo).
//
// The per-step population halo itself does not go through NCCL when the neighbour GPUs are peers (NVLink / NVSwitch):
// each rank maps its neighbours' arrays (cudaIpc) and a put kernel stores the face planes straight into the
// neighbour's ghost planes and raises a flag there (lbgpu.cu, "peer halo").  NCCL then carries only the handles at
// start-up, the small sums and the in-cycle byte planes of the free-surface / particle-flag updates.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>

#include <cstdlib>
#include <string>

namespace lbcomm {

// the subset of nccl.h this engine uses (ABI-stable since NCCL 2.7)
typedef struct ncclComm* ncclComm_t;
struct ncclUniqueId { char internal[128]; };
enum { ncclSuccess = 0 };
enum { ncclUint8 = 1, ncclInt32 = 2, ncclUint32 = 3, ncclUint64 = 5, ncclFloat64 = 8 };
enum { ncclSum = 0, ncclMax = 2 };

struct Api {
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int*) = nullptr;
    std::string path;
};

inline Api& api() {
    static Api a;
    return a;
}

// returns an empty string on success, else what went wrong
inline std::string load() {
    Api& a = api();
    if (a.lib) return "";
    const char* cands[] = { getenv("LBGPU_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
    std::string tried;
    for (const char* c : cands) {
        if (!c || !*c) continue;
        void* h = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
        if (h) { a.lib = h; a.path = c; break; }
        tried += std::string(c) + ": " + (dlerror() ? dlerror() : "?") + "; ";
    }
    if (!a.lib) return "cannot load NCCL (" + tried + ")";
    bool ok = true;
    auto sym = [&](const char* n) { void* p = dlsym(a.lib, n); if (!p) ok = false; return p; };
    a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
    a.Send = (decltype(a.Send))sym("ncclSend");
    a.Recv = (decltype(a.Recv))sym("ncclRecv");
    a.AllReduce = (decltype(a.AllReduce))sym("ncclAllReduce");
    a.AllGather = (decltype(a.AllGather))sym("ncclAllGather");
    a.GroupStart = (decltype(a.GroupStart))sym("ncclGroupStart");
    a.GroupEnd = (decltype(a.GroupEnd))sym("ncclGroupEnd");
    a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
    a.GetVersion = (decltype(a.GetVersion))sym("ncclGetVersion");
    if (!ok) { dlclose(a.lib); a.lib = nullptr; return "NCCL library " + a.path + " lacks a required symbol"; }
    return "";
}

// process-wide communicator: rank r owns the slabs [r*nLocal, (r+1)*nLocal)
struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = -1;
};
inline Comm& comm() {
    static Comm c;
    return c;
}
inline bool active() { return comm().comm != nullptr && comm().world > 1; }

// cuMemGetAddressRange of the driver API (bound at run time like NCCL): cudaIpcGetMemHandle describes the whole
// allocation a pointer lies in (small cudaMalloc blocks are sub-allocated), so the exporter sends the offset along
inline int address_range(const void* p, void** base, size_t* size) {
    typedef int (*Fn)(unsigned long long*, size_t*, unsigned long long);
    static Fn fn = nullptr;
    if (!fn) {
        void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return -1;
        fn = (Fn)dlsym(h, "cuMemGetAddressRange_v2");
        if (!fn) return -1;
    }
    unsigned long long b = 0;
    size_t sz = 0;
    const int rc = fn(&b, &sz, (unsigned long long)(uintptr_t)p);
    if (rc != 0) return rc;
    *base = (void*)(uintptr_t)b; *size = sz;
    return 0;
}

}  // namespace lbcomm
