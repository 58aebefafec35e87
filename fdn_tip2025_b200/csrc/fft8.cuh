// 8x8 real 2-D FFT kept entirely in registers: one thread owns one 8x8 patch of one channel.
// Conventions match torch.fft.rfft2 / irfft2 with norm='backward' (FDN_arch.py:460,469,585-589,614-630):
//   S[ky][kx] = sum_{y,x} p[y][x] e^{-2 pi i (ky y + kx x)/8},  kx = 0..4
//   irfft2: C2C inverse along ky, then C2R along kx (imag of the kx=0 and kx=4 columns ignored), scale 1/64.
// The self-conjugate bins (0,0),(4,0),(0,4),(4,4) come out with an imaginary part that is exactly zero, as
// they do in the reference's CPU FFT, so replace_denormals maps them to +1e-10.
#pragma once
#include "fdn_common.cuh"

#define FDN_SQRT1_2 0.70710678118654752440f

// forward real FFT of 8 samples -> bins 0..4 (X0 and X4 have imag exactly 0)
__device__ __forceinline__ void fft8_r2c(const float x[8], float2 X[5]) {
    const float c = FDN_SQRT1_2;
    float a0 = x[0] + x[4], a1 = x[0] - x[4], a2 = x[2] + x[6], a3 = x[2] - x[6];
    float a4 = x[1] + x[5], a5 = x[1] - x[5], a6 = x[3] + x[7], a7 = x[3] - x[7];
    float e0 = a0 + a2, e2 = a0 - a2, o0 = a4 + a6, o2 = a4 - a6;
    X[0] = make_float2(e0 + o0, 0.f);
    X[4] = make_float2(e0 - o0, 0.f);
    X[2] = make_float2(e2, -o2);
    X[1] = make_float2(a1 + c * (a5 - a7), -a3 - c * (a5 + a7));
    X[3] = make_float2(a1 + c * (a7 - a5), a3 - c * (a5 + a7));
}

// in-place complex FFT of 8 points, natural order.  SGN = -1 forward, +1 inverse (unscaled)
template <int SGN>
__device__ __forceinline__ void fft8_c2c(float2 v[8]) {
    const float c = FDN_SQRT1_2;
    float2 a0 = cadd(v[0], v[4]), a1 = csub(v[0], v[4]), a2 = cadd(v[2], v[6]), a3 = csub(v[2], v[6]);
    float2 a4 = cadd(v[1], v[5]), a5 = csub(v[1], v[5]), a6 = cadd(v[3], v[7]), a7 = csub(v[3], v[7]);
    // multiply by SGN*i : (x,y) -> (-SGN*y, SGN*x)
    float2 ia3 = make_float2(-SGN * a3.y, SGN * a3.x);
    float2 ia7 = make_float2(-SGN * a7.y, SGN * a7.x);
    float2 e0 = cadd(a0, a2), e2 = csub(a0, a2), e1 = cadd(a1, ia3), e3 = csub(a1, ia3);
    float2 o0 = cadd(a4, a6), o2 = csub(a4, a6), o1 = cadd(a5, ia7), o3 = csub(a5, ia7);
    // twiddles w^k = e^{SGN i pi k/4}
    float2 t1 = make_float2(c * (o1.x - SGN * o1.y), c * (o1.y + SGN * o1.x));     // o1 * (c + SGN i c)
    float2 t2 = make_float2(-SGN * o2.y, SGN * o2.x);                               // o2 * (SGN i)
    float2 t3 = make_float2(c * (-o3.x - SGN * o3.y), c * (-o3.y + SGN * o3.x));    // o3 * (-c + SGN i c)
    v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
    v[1] = cadd(e1, t1); v[5] = csub(e1, t1);
    v[2] = cadd(e2, t2); v[6] = csub(e2, t2);
    v[3] = cadd(e3, t3); v[7] = csub(e3, t3);
}

// inverse real FFT: bins 0..4 (imag of X0, X4 ignored) -> 8 samples, unscaled
__device__ __forceinline__ void fft8_c2r(const float2 X[5], float x[8]) {
    const float c = FDN_SQRT1_2;
    float A = X[0].x + X[4].x, D = X[0].x - X[4].x, B = 2.f * X[2].x, G = 2.f * X[2].y;
    float sr = 2.f * (X[1].x + X[3].x), si = 2.f * (X[1].y - X[3].y);
    x[0] = A + B + sr;
    x[4] = A + B - sr;
    x[2] = A - B - si;
    x[6] = A - B + si;
    float pr = c * (X[1].x - X[1].y), pi = c * (X[1].x + X[1].y);       // X1 * e^{i pi/4}
    float qr = c * (-X[3].x - X[3].y), qi = c * (X[3].x - X[3].y);      // X3 * e^{3 i pi/4}
    float u = 2.f * (pr + qr), w = 2.f * (pi - qi);
    x[1] = D - G + u;
    x[5] = D - G - u;
    x[3] = D + G - w;
    x[7] = D + G + w;
}

// p[y][x] (row-major 64 floats) -> S[ky][kx]
__device__ __forceinline__ void rfft2_8x8(const float p[64], float2 S[8][5]) {
#pragma unroll
    for (int y = 0; y < 8; ++y) fft8_r2c(p + 8 * y, S[y]);
#pragma unroll
    for (int kx = 0; kx < 5; ++kx) {
        float2 col[8];
#pragma unroll
        for (int y = 0; y < 8; ++y) col[y] = S[y][kx];
        fft8_c2c<-1>(col);
#pragma unroll
        for (int y = 0; y < 8; ++y) S[y][kx] = col[y];
    }
}

// S[ky][kx] -> p[y][x], including the 1/64 scale.  S is destroyed.
__device__ __forceinline__ void irfft2_8x8(float2 S[8][5], float p[64]) {
#pragma unroll
    for (int kx = 0; kx < 5; ++kx) {
        float2 col[8];
#pragma unroll
        for (int y = 0; y < 8; ++y) col[y] = S[y][kx];
        fft8_c2c<1>(col);
#pragma unroll
        for (int y = 0; y < 8; ++y) S[y][kx] = col[y];
    }
#pragma unroll
    for (int y = 0; y < 8; ++y) {
        float r[8];
        fft8_c2r(S[y], r);
#pragma unroll
        for (int x = 0; x < 8; ++x) p[8 * y + x] = r[x] * (1.0f / 64.0f);
    }
}
