// membench.cu -- what does the D3Q19 pull access pattern cost on this GPU, without the collision?
// 19 shifted read streams + 19 write streams over a 256^3 lattice (the cfg2 shape), against a plain copy.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
constexpr int Q = 19;
struct Args { const double* src[Q]; double* dst[Q]; int off[Q]; unsigned n0, n1; };

template <int CPT>
__global__ void __launch_bounds__(128) k_pull(const __grid_constant__ Args a) {
    const unsigned base = a.n0 + (blockIdx.x * 128u + threadIdx.x) * CPT;
    if (base + CPT > a.n1) return;
    double f[Q][CPT];
#pragma unroll
    for (int k = 0; k < Q; ++k)
#pragma unroll
        for (int c = 0; c < CPT; ++c) f[k][c] = a.src[k][(ptrdiff_t)(base + c) - a.off[k]];
#pragma unroll
    for (int k = 0; k < Q; ++k)
#pragma unroll
        for (int c = 0; c < CPT; ++c) a.dst[k][base + c] = f[k][c] + 1.0;
}
// same, adding `work` dependent-free fp64 adds per population to emulate the collision's instruction load
template <int WORK>
__global__ void __launch_bounds__(128, 4) k_pull_work(const __grid_constant__ Args a) {
    const unsigned i = a.n0 + blockIdx.x * 128u + threadIdx.x;
    if (i >= a.n1) return;
    double f[Q];
#pragma unroll
    for (int k = 0; k < Q; ++k) f[k] = a.src[k][(ptrdiff_t)i - a.off[k]];
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < Q; ++k) s += f[k];
#pragma unroll
    for (int w = 0; w < WORK; ++w)
#pragma unroll
        for (int k = 0; k < Q; ++k) f[k] = __dadd_rn(__dmul_rn(f[k], 0.999), s);
#pragma unroll
    for (int k = 0; k < Q; ++k) a.dst[k][i] = f[k];
}
__global__ void k_copy(const double2* __restrict__ s, double2* __restrict__ d, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t st = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += st) d[i] = s[i];
}
template <class F> float timeit(F f, int reps) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(a); for (int i = 0; i < reps; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); CK(cudaGetLastError()); CK(cudaDeviceSynchronize()); return ms / reps;
}
int main() {
    const int X = 256, Y = 256, Z = 256; const size_t N = (size_t)X * Y * Z; const size_t pad = (size_t)X * Y + X + 64;
    const size_t stride = N;
    double *A, *B;
    CK(cudaMalloc(&A, (stride * Q + 2 * pad) * 8)); CK(cudaMalloc(&B, (stride * Q + 2 * pad) * 8));
    CK(cudaMemset(A, 0, (stride * Q + 2 * pad) * 8)); CK(cudaMemset(B, 0, (stride * Q + 2 * pad) * 8));
    const int cx[Q] = { 0, 1, -1, 0, 0, 0, 0, 1, -1, -1, 1, 0, 0, 0, 0, 1, -1, 1, -1 };
    const int cy[Q] = { 0, 0, 0, 1, -1, 0, 0, 1, -1, 1, -1, 1, -1, -1, 1, 0, 0, 0, 0 };
    const int cz[Q] = { 0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, -1, 1 };
    Args a; 
    for (int k = 0; k < Q; ++k) { a.src[k] = A + pad + k * stride; a.dst[k] = B + pad + k * stride; a.off[k] = cx[k] + X * (cy[k] + Y * cz[k]); }
    a.n0 = 0; a.n1 = (unsigned)N;
    const double bytes = (double)N * 304;
    float ms;
    ms = timeit([&] { k_copy<<<148 * 16, 512>>>((const double2*)A, (double2*)B, stride * Q / 2); }, 20);
    printf("copy (grid-stride double2, 19 planes)  %.4f ms  %.1f GB/s\n", ms, (double)stride * Q * 16 / ms / 1e6);
    ms = timeit([&] { cudaMemcpyAsync(B, A, stride * Q * 8, cudaMemcpyDeviceToDevice); }, 20);
    printf("cudaMemcpy D2D                         %.4f ms  %.1f GB/s\n", ms, (double)stride * Q * 16 / ms / 1e6);
    ms = timeit([&] { k_pull<1><<<(unsigned)(N / 128), 128>>>(a); }, 20);
    printf("pull 19 streams, 1 cell/thread         %.4f ms  %.1f GB/s  %.0f MLUPS\n", ms, bytes / ms / 1e6, N / ms / 1e3);
    ms = timeit([&] { k_pull<2><<<(unsigned)(N / 256), 128>>>(a); }, 20);
    printf("pull 19 streams, 2 cells/thread        %.4f ms  %.1f GB/s  %.0f MLUPS\n", ms, bytes / ms / 1e6, N / ms / 1e3);
    ms = timeit([&] { k_pull_work<4><<<(unsigned)(N / 128), 128>>>(a); }, 20);
    printf("pull + 171 fp64 instr                  %.4f ms  %.1f GB/s  %.0f MLUPS\n", ms, bytes / ms / 1e6, N / ms / 1e3);
    ms = timeit([&] { k_pull_work<12><<<(unsigned)(N / 128), 128>>>(a); }, 20);
    printf("pull + 475 fp64 instr                  %.4f ms  %.1f GB/s  %.0f MLUPS\n", ms, bytes / ms / 1e6, N / ms / 1e3);
    ms = timeit([&] { k_pull_work<20><<<(unsigned)(N / 128), 128>>>(a); }, 20);
    printf("pull + 780 fp64 instr                  %.4f ms  %.1f GB/s  %.0f MLUPS\n", ms, bytes / ms / 1e6, N / ms / 1e3);
    return 0;
}
