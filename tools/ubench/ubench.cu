// Round-2 micro-benchmarks that decide kernel designs (run on the B200 through gpurun; not product code).
//   1. fp32 issue rate: FFMA vs packed FFMA2 (fma.rn.f32x2), alone and mixed with ALU work
//   2. how fast one CTA per SM can stage [K channel rows][128 pixels] tiles of an NCHW tensor (rows one plane apart) into
//      shared memory: (a) one cp.async.bulk per 512-byte row issued by NW warps (round-1 k_pw_mma loader),
//                     (b) 16-byte cp.async issued by all 256 consumer threads with commit / wait groups
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/ubench tools/ubench/ubench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

// ------------------------------------------------------------------------------------------------ 1. FFMA vs FFMA2
template <int MODE>
__global__ void __launch_bounds__(256) k_fma(float* out, int iters, float a, float b) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
    unsigned u = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
        } else {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                float2 v = __ffma2_rn(make_float2(x[i], x[i + 1]), make_float2(a, a), make_float2(b, b));
                x[i] = v.x; x[i + 1] = v.y;
            }
        }
        if (MODE >= 2) {          // 8 integer ALU ops per 16 FMAs
#pragma unroll
            for (int i = 0; i < 8; ++i) u = (u ^ (u >> 3)) + 0x9e3779b9u * (unsigned)i;
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)u;
}

static void bench_fma() {
    float* out;
    CK(cudaMalloc(&out, 148 * 8 * 256 * 4));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int iters = 20000;
    const char* names[4] = {"FFMA", "FFMA2", "FFMA + ALU", "FFMA2 + ALU"};
    for (int mode = 0; mode < 4; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0));
            if (mode == 0) k_fma<0><<<148 * 8, 256>>>(out, iters, 1.0001f, 0.5f);
            if (mode == 1) k_fma<1><<<148 * 8, 256>>>(out, iters, 1.0001f, 0.5f);
            if (mode == 2) k_fma<2><<<148 * 8, 256>>>(out, iters, 1.0001f, 0.5f);
            if (mode == 3) k_fma<3><<<148 * 8, 256>>>(out, iters, 1.0001f, 0.5f);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
        }
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        double fma = 148.0 * 8 * 256 * 16.0 * iters;
        printf("fma  %-12s %8.3f ms  %7.1f TFLOP/s (2 flop per FMA)\n", names[mode], ms, 2 * fma / ms / 1e9);
    }
    CK(cudaFree(out));
}

// ------------------------------------------------------------------------------------------------ 2. tile staging
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.b32 %0, 1, 0, P1;\n}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}

#define SLOT_FLOATS (32 * 128)

// (a) NW loader warps issue one bulk copy per row; 256 consumer threads read every element once
template <int NW>
__global__ void __launch_bounds__(256 + 32 * NW, 1) k_stage_bulk(const float* __restrict__ x, float* __restrict__ out, int K, long long HW, int ntiles,
                                                                 int ring) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* s_raw = reinterpret_cast<float*>(smem);
    uint64_t* full = reinterpret_cast<uint64_t*>(s_raw + ring * SLOT_FLOATS);
    uint64_t* empty = full + 16;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < ring; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int nkb = K / 32;
    const int tiles_per_img = (int)(HW / 128);
    if (warp >= 8) {
        const int lw = warp - 8;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int b = tile / tiles_per_img, p0 = (tile % tiles_per_img) * 128;
            for (int kb = 0; kb < nkb; ++kb, ++it) {
                const int r = it % ring;
                if (it >= (uint32_t)ring) mbar_wait(&empty[r], ((it / ring) - 1) & 1);
                if (lw == 0 && lane == 0) mbar_expect_tx(&full[r], 32 * 512);
                for (int row = lane * NW + lw; row < 32; row += 32 * NW)
                    bulk_g2s(s_raw + r * SLOT_FLOATS + row * 128, x + ((size_t)b * K + kb * 32 + row) * HW + p0, 512, &full[r]);
            }
        }
    } else {
        const int pix = tid & 127, half = tid >> 7;
        float acc = 0.f;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
            for (int kb = 0; kb < nkb; ++kb, ++it) {
                const int r = it % ring;
                mbar_wait(&full[r], (it / ring) & 1);
                const float* raw = s_raw + r * SLOT_FLOATS + pix;
#pragma unroll
                for (int k = 0; k < 16; ++k) acc += raw[(2 * k + half) * 128];
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[r]);
            }
        out[blockIdx.x * 256 + tid] = acc;
    }
}

// (b) the 256 consumer threads issue 16-byte cp.async themselves: warp w copies rows w, w+8, w+16, w+24 of the K block DEPTH-1 ahead
__global__ void __launch_bounds__(256, 1) k_stage_cpasync(const float* __restrict__ x, float* __restrict__ out, int K, long long HW, int ntiles, int ring) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* s_raw = reinterpret_cast<float*>(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nkb = K / 32;
    const int tiles_per_img = (int)(HW / 128);
    const int my_tiles = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const uint32_t total = (uint32_t)my_tiles * nkb;
    auto issue = [&](uint32_t j) {          // K block number j of this CTA's sequence
        if (j < total) {
            const int tile = blockIdx.x + (j / nkb) * gridDim.x, kb = j % nkb;
            const int b = tile / tiles_per_img, p0 = (tile % tiles_per_img) * 128;
            const float* src = x + ((size_t)b * K + kb * 32 + warp) * HW + p0 + lane * 4;
            float* dst = s_raw + (j % ring) * SLOT_FLOATS + warp * 128 + lane * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q) cp_async16(dst + q * 8 * 128, src + (size_t)q * 8 * HW);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (int j = 0; j < ring - 1; ++j) issue(j);
    const int pix = tid & 127, half = tid >> 7;
    float acc = 0.f;
    for (uint32_t it = 0; it < total; ++it) {
        // groups it+1 .. it+ring-2 may still be in flight
        if (ring == 2) asm volatile("cp.async.wait_group 0;" ::: "memory");
        else if (ring == 3) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else if (ring == 4) asm volatile("cp.async.wait_group 2;" ::: "memory");
        else if (ring == 6) asm volatile("cp.async.wait_group 4;" ::: "memory");
        else if (ring == 8) asm volatile("cp.async.wait_group 6;" ::: "memory");
        else asm volatile("cp.async.wait_group 10;" ::: "memory");      // ring 12
        __syncthreads();                       // every thread's copies of K block `it` have landed; slot (it-1)%ring is free
        issue(it + ring - 1);
        const float* raw = s_raw + (it % ring) * SLOT_FLOATS + pix;
#pragma unroll
        for (int k = 0; k < 16; ++k) acc += raw[(2 * k + half) * 128];
    }
    out[blockIdx.x * 256 + tid] = acc;
}

static void bench_stage() {
    const int B = 4;
    const long long HW = 640LL * 1120;
    const int Kmax = 128;
    float* x; float* out;
    CK(cudaMalloc(&x, (size_t)B * Kmax * HW * 4));
    CK(cudaMemset(x, 0, (size_t)B * Kmax * HW * 4));
    CK(cudaMalloc(&out, 148 * 512 * 4));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int ntiles = (int)(HW / 128) * B;
    for (int K : {32, 96, 128}) {
        const double bytes = (double)B * K * HW * 4;
        for (int variant = 0; variant < 7; ++variant) {
            const int rings[7] = {8, 8, 4, 8, 12, 8, 12};
            const int ring = rings[variant];
            const size_t smem = (size_t)ring * SLOT_FLOATS * 4 + 512;
            float ms = 0;
            for (int rep = 0; rep < 2; ++rep) {
                CK(cudaEventRecord(e0));
                if (variant == 0) {
                    CK(cudaFuncSetAttribute(k_stage_bulk<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    k_stage_bulk<3><<<148, 256 + 96, smem>>>(x, out, K, HW, ntiles, ring);
                } else if (variant == 1) {
                    CK(cudaFuncSetAttribute(k_stage_bulk<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    k_stage_bulk<8><<<148, 256 + 256, smem>>>(x, out, K, HW, ntiles, ring);
                } else {
                    CK(cudaFuncSetAttribute(k_stage_cpasync, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    // variants 5, 6: two CTAs per SM share the work (grid 296)
                    const int grid = variant >= 5 ? 296 : 148;
                    k_stage_cpasync<<<grid, 256, variant >= 5 ? (size_t)(ring / 2) * SLOT_FLOATS * 4 + 512 : smem>>>(x, out, K, HW, ntiles, variant >= 5 ? ring / 2 : ring);
                }
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                CK(cudaGetLastError());
                CK(cudaEventElapsedTime(&ms, e0, e1));
            }
            const char* nm[7] = {"bulk 3 warps ring 8", "bulk 8 warps ring 8", "cp.async ring 4", "cp.async ring 8", "cp.async ring 12",
                                 "cp.async 2 CTA/SM ring 4", "cp.async 2 CTA/SM ring 6"};
            printf("stage K=%3d %-26s %7.3f ms  %7.1f GB/s  (%.1f cycles per 512-byte row per SM at 1.9 GHz)\n", K, nm[variant], ms, bytes / ms / 1e6,
                   ms * 1e-3 * 1.9e9 / ((double)ntiles * K / 148));
        }
    }
}

// ------------------------------------------------------------------------------------------------ 3. legacy mma.sync tf32 rate
__global__ void __launch_bounds__(256) k_hmma(float* out, int iters) {
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    uint32_t a[4] = {0x3f800000u + threadIdx.x, 0x3f800000u, 0x3f000000u, 0x3f800000u}, b[2] = {0x3f800000u, 0x3f000000u + threadIdx.x};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static void bench_hmma() {
    float* out;
    CK(cudaMalloc(&out, 148 * 4 * 256 * 4));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int iters = 20000;
    for (int ctas : {1, 2, 4}) {
        float ms = 0;
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0));
            k_hmma<<<148 * ctas, 256>>>(out, iters);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms, e0, e1));
        }
        const double mma = 148.0 * ctas * 8 * 8.0 * iters;          // warp-level instructions
        printf("hmma m16n8k8 tf32, %d CTAs/SM of 8 warps: %8.3f ms  %7.1f TFLOP/s  (%.2f HMMA per cycle per SM at 1.9 GHz)\n", ctas, ms,
               mma * 2 * 16 * 8 * 8 / ms / 1e9, mma / 148 / (ms * 1e-3 * 1.9e9));
    }
    CK(cudaFree(out));
}

int main(int argc, char** argv) {
    if (argc > 1 && argv[1][0] == 'h') { bench_hmma(); return 0; }
    bench_fma();
    bench_stage();
    bench_hmma();
    return 0;
}
