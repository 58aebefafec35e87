// Asynchronous global -> shared staging for the persistent FFT kernels: TMA bulk copies (cp.async.bulk, one instruction per
// contiguous tile, completion counted in bytes on an mbarrier) and 8-byte cp.async for strided column tiles.
//
// Why: a CTA that loads its tile, computes and stores is in its load phase a third of the time, so only a few tens of KB per SM are
// in flight - the row kernels sat at 3.7 TB/s however few instructions they executed (profiles/r3_fft_ab.txt).  The persistent
// kernels keep two or three tiles per CTA in flight while computing (Little: 6.5 TB/s x ~1.5 us = 66 KB per SM).
//
// The emulation build (tests/emu) has no asynchronous engine: copies are done on the spot by the issuing thread and waits are
// no-ops, which is equivalent because every consumer is separated from the issue by a __syncthreads() in these kernels.
#pragma once
#include "fdn_common.cuh"

namespace fasync {

#ifdef FDN_EMU
struct Bar { int dummy[2]; };
__device__ __forceinline__ void init(Bar*, unsigned) {}
__device__ __forceinline__ void fence_init() {}
__device__ __forceinline__ void wait(Bar*, unsigned) {}
__device__ __forceinline__ void expect_tx(Bar*, unsigned) {}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, Bar*) { memcpy(dst, src, bytes); }
__device__ __forceinline__ void cp8(void* dst, const void* src) { memcpy(dst, src, 8); }
__device__ __forceinline__ void cp_commit() {}
template <int N> __device__ __forceinline__ void cp_wait() {}
#else
typedef unsigned long long Bar;
__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void init(Bar* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// wait for the phase with the given parity (nanosleep back-off between probes; traps after ~1 s)
__device__ __forceinline__ void wait(Bar* bar, unsigned parity) {
    unsigned done;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.b32 %0, 1, 0, P1;\n"
        "}\n" : "=r"(done) : "r"(s32(bar)), "r"(parity) : "memory");
    unsigned spins = 0;
    while (!done) {
        if (++spins > (1u << 24)) __trap();                            // a copy that never lands is a bug: fail the launch instead of hanging the GPU
        asm volatile("nanosleep.u32 32;" ::: "memory");
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
            "selp.b32 %0, 1, 0, P1;\n"
            "}\n" : "=r"(done) : "r"(s32(bar)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void expect_tx(Bar* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
}
// one TMA bulk copy: bytes % 16 == 0, src and dst 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, Bar* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src),
                 "r"(bytes), "r"(s32(bar))
                 : "memory");
}
__device__ __forceinline__ void cp8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
#endif

}  // namespace fasync

// SM count of the current device (persistent grids), cached per device
static inline int fdn_sm_count() {
#ifdef FDN_EMU
    return 4;
#else
    static int n_dev[FDN_MAX_DEVICES] = {0};
    const int dev = fdn_device();
    if (n_dev[dev] == 0) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        n_dev[dev] = n > 0 ? n : 148;
    }
    return n_dev[dev];
#endif
}
