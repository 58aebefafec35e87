// Shared definitions for the libfdn_b200 kernels (sm_100a).
#pragma once

#ifdef FDN_EMU
#include "cuda_emu.h"   // tests/emu: host emulation used only for debugging in the GPU-less container
#define FDN_DYN_SMEM(name) unsigned char* name = emu::dyn_smem()
#define FDN_LAUNCH(kern, grid, block, smem, stream, ...) \
    emu::launch((grid), (block), (smem), [=]() { kern(__VA_ARGS__); }, true)
#define FDN_LAUNCH_SEQ(kern, grid, block, smem, stream, ...) \
    emu::launch((grid), (block), (smem), [=]() { kern(__VA_ARGS__); }, false)
#else
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#define FDN_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define FDN_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
// kernels that neither synchronise nor shuffle (the emulator may run their threads as a plain loop)
#define FDN_LAUNCH_SEQ(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

#include <string>

#define FDN_API extern "C" __attribute__((visibility("default")))

// ---- error plumbing (fdn_last_error_string) -------------------------------------------------------
void fdn_set_error(const std::string& msg);
int fdn_check_launch(const char* what);   // returns 0 or the cudaError_t of the last launch

#define FDN_REQUIRE(cond, msg)                                                   \
    do {                                                                         \
        if (!(cond)) {                                                           \
            fdn_set_error(std::string(__func__) + ": " + (msg) + " [" #cond "]"); \
            return -1;                                                           \
        }                                                                        \
    } while (0)

// Library state that depends on the GPU (twiddle tables, cudaFuncSetAttribute opt-ins, SM counts) is kept per device: one
// process may drive several GPUs from worker threads, each with its own current device (SURVEY.md section 8(b) "Threading / streams").
#define FDN_MAX_DEVICES 64
static inline int fdn_device() {
    int d = 0;
    cudaGetDevice(&d);
    return (d >= 0 && d < FDN_MAX_DEVICES) ? d : 0;
}

static inline bool fdn_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline int fdn_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- small device helpers --------------------------------------------------------------------------
#ifdef FDN_EMU
#define FDN_EXPF(x) expf(x)
__device__ __forceinline__ float fdn_rcp_fast(float x) { return 1.0f / x; }
#else
// e^x = ex2.approx.ftz(x log2 e) and 1/x = rcp.approx.ftz(x) as single MUFU operations: __expf / __fdividef wrap the same
// instructions in denormal-range scaling (a compare and two multiplies each) that the GELU arguments never need - a result below
// 2^-126 is flushed to zero, where it multiplies into an output of that size anyway.
__device__ __forceinline__ float fdn_exp_fast(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
    return r;
}
__device__ __forceinline__ float fdn_rcp_fast(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fdn_ex2_fast(float x) {           // 2^x, one MUFU
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
#define FDN_EXPF(x) fdn_exp_fast(x)
#endif
#ifdef FDN_EMU
__device__ __forceinline__ float fdn_ex2_fast(float x) { return exp2f(x); }
#endif
// GELU works on u = |x| sqrt(log2 e) / sqrt 2 so that e^{-z^2} = 2^{-u^2} needs no pre-scaling: z = u / sqrt(log2 e)
#define FDN_GELU_KU 0.8493218002880191f      /* sqrt(log2 e) / sqrt 2 */
#define FDN_GELU_K39 0.32469629835150216f    /* 0.39 / sqrt(log2 e) */
#define FDN_GELU_KH (-0.5887050112577373f)   /* -1 / (sqrt 2 sqrt(log2 e)):  KH u = -0.5 |x| */
// erf-form GELU (F.gelu default): 0.5 x (1 + erf(x / sqrt 2)), evaluated through erfc(z) = t P(t) e^{-z^2}, t = 1/(1 + 0.39 z), z = |x| / sqrt 2
// (degree-6 fit, |error| < 2e-8 before rounding).  The result is formed as relu(x) - 0.5 |x| erfc(z), which keeps small results of
// negative arguments free of cancellation (there it is just -0.5 |x| erfc(z)).  In fp32 its error equals the erff() formulation's (max 3.8e-7 / rms 6.1e-8 vs 4.5e-7 / 6.8e-8 over [-8, 8])
// at about half the instructions: two MUFU (rcp, ex2) and ten FMA-pipe operations.
__device__ __forceinline__ float fdn_gelu(float x) {
    const float u = fabsf(x) * FDN_GELU_KU;
    const float t = fdn_rcp_fast(fmaf(FDN_GELU_K39, u, 1.0f));
    float q = fmaf(-2.280578155e-01f, t, 8.887884326e-01f);
    q = fmaf(q, t, -6.388445279e-01f);
    q = fmaf(q, t, 6.523336912e-01f);
    q = fmaf(q, t, 9.027867826e-02f);
    q = fmaf(q, t, 2.355015248e-01f);
    const float ec = q * t * fdn_ex2_fast(-u * u);             // erfc(z)
    // 0.5 x (2 - erfc) for x >= 0 and 0.5 x erfc for x < 0 are both  relu(x) - 0.5 |x| erfc(z): no select, one rounding fewer
    return fmaf(u * FDN_GELU_KH, ec, fmaxf(x, 0.f));
}
// ---- packed fp32x2 helpers (Blackwell FMUL2 / FFMA2; scalar forms in the emulation build, bit-identical) ----------------------------
// A scalar operand written as the pair {s, s} is encoded by ptxas as a broadcast register operand (no move).
#ifdef FDN_EMU
__device__ __forceinline__ float2 f2mul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
__device__ __forceinline__ float2 f2mul_s(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
__device__ __forceinline__ float2 f2fma_c(float2 a, float2 b, float c) { return make_float2(fmaf(a.x, b.x, c), fmaf(a.y, b.y, c)); }
__device__ __forceinline__ float2 f2add_s(float2 a, float s) { return make_float2(a.x + s, a.y + s); }
#else
__device__ __forceinline__ float2 f2add_s(float2 a, float s) {
    float2 r;
    asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %4}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(s));
    return r;
}
__device__ __forceinline__ float2 f2mul(float2 a, float2 b) {
    float2 r;
    asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 f2mul_s(float2 a, float s) {
    float2 r;
    asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %4}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(s));
    return r;
}
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) {
    float2 r;
    asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %7}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0, %1}, rd; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}
__device__ __forceinline__ float2 f2fma_c(float2 a, float2 b, float c) {   // a*b + {c, c}
    float2 r;
    asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %6}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0, %1}, rd; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c));
    return r;
}
#endif
// gelu of two values at once (and gelu(g) * l below): the same operation sequence as fdn_gelu (bit-identical), the polynomial and the
// products packed
__device__ __forceinline__ float2 fdn_gelu2(float2 g) {
    const float2 z = make_float2(fabsf(g.x) * FDN_GELU_KU, fabsf(g.y) * FDN_GELU_KU);          // u of fdn_gelu
    const float2 den = f2fma_c(make_float2(FDN_GELU_K39, FDN_GELU_K39), z, 1.0f);
    const float2 t = make_float2(fdn_rcp_fast(den.x), fdn_rcp_fast(den.y));
    float2 q = f2fma_c(make_float2(-2.280578155e-01f, -2.280578155e-01f), t, 8.887884326e-01f);
    q = f2fma_c(q, t, -6.388445279e-01f);
    q = f2fma_c(q, t, 6.523336912e-01f);
    q = f2fma_c(q, t, 9.027867826e-02f);
    q = f2fma_c(q, t, 2.355015248e-01f);
    const float2 nz2 = f2mul(make_float2(-z.x, -z.y), z);
    const float2 e = make_float2(fdn_ex2_fast(nz2.x), fdn_ex2_fast(nz2.y));
    const float2 ec = f2mul(f2mul(q, t), e);
    return f2fma(f2mul_s(z, FDN_GELU_KH), ec, make_float2(fmaxf(g.x, 0.f), fmaxf(g.y, 0.f)));     // relu(g) - 0.5 |g| erfc(z)
}
__device__ __forceinline__ float2 fdn_gelu_gate2(float2 g, float2 l) {
    const float2 z = make_float2(fabsf(g.x) * FDN_GELU_KU, fabsf(g.y) * FDN_GELU_KU);          // u of fdn_gelu
    const float2 den = f2fma_c(make_float2(FDN_GELU_K39, FDN_GELU_K39), z, 1.0f);
    const float2 t = make_float2(fdn_rcp_fast(den.x), fdn_rcp_fast(den.y));
    float2 q = f2fma_c(make_float2(-2.280578155e-01f, -2.280578155e-01f), t, 8.887884326e-01f);
    q = f2fma_c(q, t, -6.388445279e-01f);
    q = f2fma_c(q, t, 6.523336912e-01f);
    q = f2fma_c(q, t, 9.027867826e-02f);
    q = f2fma_c(q, t, 2.355015248e-01f);
    const float2 nz2 = f2mul(make_float2(-z.x, -z.y), z);
    const float2 e = make_float2(fdn_ex2_fast(nz2.x), fdn_ex2_fast(nz2.y));
    const float2 ec = f2mul(f2mul(q, t), e);
    return f2mul(f2fma(f2mul_s(z, FDN_GELU_KH), ec, make_float2(fmaxf(g.x, 0.f), fmaxf(g.y, 0.f))), l);
}
__device__ __forceinline__ float fdn_lrelu(float x) { return x > 0.f ? x : 0.1f * x; }
__device__ __forceinline__ float fdn_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
// replace_denormals (FDN_arch.py:548-553): |v| < 1e-10 -> +1e-10
__device__ __forceinline__ float fdn_rd(float v) { return (v < 1e-10f && v > -1e-10f) ? 1e-10f : v; }
__device__ __forceinline__ float fdn_act(float v, int act) {
    return act == 1 ? fdn_lrelu(v) : (act == 2 ? fmaxf(v, 0.f) : v);
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }  // a*conj(b)
// Complex add / subtract / multiply-by-a-real on Blackwell's packed fp32x2 pipe (FADD2 / FFMA2: one issue slot for both halves of a
// float2 that sits in an aligned register pair, which is how LDS.64 / LDG.64 and other packed results deliver it).  IEEE round-to-nearest
// per half, no FTZ: bit-identical to the scalar forms, which the emulation build keeps.  The FFT butterflies are mostly these.
#ifdef FDN_EMU
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cfma(float2 a, float s, float2 c) { return make_float2(fmaf(a.x, s, c.x), fmaf(a.y, s, c.y)); }   // a*s + c
#else
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
    float2 r;
    asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
    float2 r;
    asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; sub.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 cfma(float2 a, float s, float2 c) {   // a*s + c
    float2 r;
    asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %4}; mov.b64 rc, {%5, %6}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0, %1}, rd; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(s), "f"(c.x), "f"(c.y));
    return r;
}
#endif
