// CUDA-semantics emulator for kernel debugging on a machine without a GPU.  TEST INFRASTRUCTURE ONLY.
//
// The build container has nvcc but no GPU, and a gpurun round trip takes minutes.  Compiling the kernel
// sources of fdn_tip2025_b200/csrc with `g++ -DFDN_EMU -I tests/emu` maps the small subset of CUDA they
// use (threadIdx/blockIdx, __shared__, __syncthreads, warp shuffles, float2/float4, launch syntax through
// the FDN_LAUNCH macros) onto host threads so the very same source can be checked against the oracle on
// tiny shapes before any GPU time is spent.  The product never loads the resulting library: the package
// loader (fdn_tip2025_b200/_lib.py) only opens libfdn_b200.so and raises if it is missing.  Kernels that
// use inline PTX (tcgen05 / TMA / mbarrier) are compiled out under FDN_EMU and are only tested on the GPU.
//
// Execution model: blocks run one after another; FDN_LAUNCH_SEQ runs the threads of a block as a plain
// loop (kernel must not synchronise), FDN_LAUNCH runs them on a persistent pool of host threads with a
// std::barrier behind __syncthreads() and per-warp barriers behind the shuffles.
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct int2 { int x, y; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

typedef void* cudaStream_t;
typedef int cudaError_t;
#define cudaSuccess 0
static inline cudaError_t cudaGetLastError() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256); return 0; }
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
enum { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2 };
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, int) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
enum cudaStreamCaptureStatus { cudaStreamCaptureStatusNone = 0, cudaStreamCaptureStatusActive = 1 };
static inline cudaError_t cudaStreamIsCapturing(cudaStream_t, cudaStreamCaptureStatus* s) { *s = cudaStreamCaptureStatusNone; return 0; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__ static
static inline double atomicAdd(double* p, double v) {          // blocks / threads of the emulator may be real host threads
    unsigned long long* q = reinterpret_cast<unsigned long long*>(p);
    unsigned long long old = __atomic_load_n(q, __ATOMIC_RELAXED), nxt;
    double cur;
    do {
        memcpy(&cur, &old, 8);
        cur += v;
        memcpy(&nxt, &cur, 8);
    } while (!__atomic_compare_exchange_n(q, &old, nxt, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    memcpy(&cur, &old, 8);
    return cur;
}
static inline int atomicMax(int* p, int v) {
    int old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
#define __align__(n) __attribute__((aligned(n)))
#define __ldg(p) (*(p))
#define __fdividef(a, b) ((a) / (b))
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline float __frcp_rn(float x) { return 1.0f / x; }
static inline int __float2int_rz(float x) { return (int)x; }
static inline float __saturatef(float x) { return x < 0.f ? 0.f : (x > 1.f ? 1.f : x); }
using std::max;
using std::min;

namespace emu {
struct ThreadCtx {
    uint3 tid, bid;
    dim3 bdim, gdim;
    int linear_tid;
    bool threaded;
};
extern thread_local ThreadCtx ctx;
unsigned char* dyn_smem();
void sync_block();
float shfl(float v, int src_lane);
void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body, bool threaded);
}  // namespace emu

#define threadIdx (emu::ctx.tid)
#define blockIdx (emu::ctx.bid)
#define blockDim (emu::ctx.bdim)
#define gridDim (emu::ctx.gdim)
#define __syncthreads() emu::sync_block()
#define __syncwarp(...) ((void)0)

static inline float __shfl_sync(unsigned, float v, int src, int = 32) { return emu::shfl(v, src); }
static inline float __shfl_xor_sync(unsigned, float v, int m, int = 32) { return emu::shfl(v, (emu::ctx.linear_tid & 31) ^ m); }
static inline float __shfl_down_sync(unsigned, float v, int d, int = 32) {
    int l = (emu::ctx.linear_tid & 31) + d;
    return emu::shfl(v, l > 31 ? (emu::ctx.linear_tid & 31) : l);
}
