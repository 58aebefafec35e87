// 8x8-patch spectral operators of the FDformer blocks, one thread per (channel, patch), FFT in registers.
//   FDFFN spectral branch  FDN_arch.py:458-470   Y = rd(X) * ffta * e^{-i fftp}  (== |rd X| ffta e^{i(angle - fftp)})
//   FDSA bin algebra       FDN_arch.py:585-630   out1 = |v'| e^{i(th_q - th_k)}, out2 = |rd(qk)| e^{i th_v'},
//                                                out3 = |rd(qk)| e^{i(th_q - th_k)},  v' = rd(v * fft)
// The polar forms are evaluated in closed form (no atan2/sincos): e^{i(th_q - th_k)} = (q/|q|) conj(k/|k|) with
// q, k already passed through replace_denormals, which bounds every modulus away from zero.
#include "fdn_common.cuh"
#include "fft8.cuh"

__device__ __forceinline__ void load_patch(const float* __restrict__ base, int W, float p[64]) {
#pragma unroll
    for (int y = 0; y < 8; ++y) {
        const float4* r = reinterpret_cast<const float4*>(base + (size_t)y * W);
        float4 a = r[0], b = r[1];
        p[8 * y + 0] = a.x; p[8 * y + 1] = a.y; p[8 * y + 2] = a.z; p[8 * y + 3] = a.w;
        p[8 * y + 4] = b.x; p[8 * y + 5] = b.y; p[8 * y + 6] = b.z; p[8 * y + 7] = b.w;
    }
}
__device__ __forceinline__ void store_patch(float* __restrict__ base, int W, const float p[64]) {
#pragma unroll
    for (int y = 0; y < 8; ++y) {
        float4* r = reinterpret_cast<float4*>(base + (size_t)y * W);
        r[0] = make_float4(p[8 * y + 0], p[8 * y + 1], p[8 * y + 2], p[8 * y + 3]);
        r[1] = make_float4(p[8 * y + 4], p[8 * y + 5], p[8 * y + 6], p[8 * y + 7]);
    }
}

// ---------------------------------------------------------------------------------------------------
// FDFFN: out = irfft2(rd(rfft2(x)) * wspec[c]) + add
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_fdffn_patch(const float* __restrict__ x, const float* __restrict__ add,
                                                     const float2* __restrict__ wspec, float* __restrict__ out,
                                                     int C, int H, int W, long long nitems) {
    long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= nitems) return;
    const int pw = W >> 3, ph = H >> 3;
    int px = (int)(item % pw);
    long long t = item / pw;
    int py = (int)(t % ph);
    long long plane = t / ph;              // b*C + c
    int c = (int)(plane % C);
    size_t off = (size_t)plane * H * W + (size_t)(py * 8) * W + px * 8;
    float p[64];
    float2 S[8][5];
    load_patch(x + off, W, p);
    rfft2_8x8(p, S);
    const float2* w = wspec + c * 40;
#pragma unroll
    for (int ky = 0; ky < 8; ++ky)
#pragma unroll
        for (int kx = 0; kx < 5; ++kx) {
            float2 z = make_float2(fdn_rd(S[ky][kx].x), fdn_rd(S[ky][kx].y));
            S[ky][kx] = cmul(z, w[ky * 5 + kx]);
        }
    irfft2_8x8(S, p);
    if (add) {
        float q[64];
        load_patch(add + off, W, q);
#pragma unroll
        for (int i = 0; i < 64; ++i) p[i] += q[i];
    }
    store_patch(out + off, W, p);
}

// One bin of the FDSA algebra.  Moduli use the hardware reciprocal square root (2 ulp): they only scale magnitudes,
// the phases come from exact complex products of the clamped inputs.  replace_denormals keeps every modulus >= 1.4e-10.
__device__ __forceinline__ float2 fdsa_bin(float2 q, float2 k, float2 v, int role) {
    float2 qk = cmul(q, k);
    qk.x = fdn_rd(qk.x);
    qk.y = fdn_rd(qk.y);
    const float A = sqrtf(qk.x * qk.x + qk.y * qk.y);                       // |rd(q k)|
    const float2 qc = make_float2(fdn_rd(q.x), fdn_rd(q.y)), kc = make_float2(fdn_rd(k.x), fdn_rd(k.y));
    const float iq = rsqrtf(qc.x * qc.x + qc.y * qc.y), ik = rsqrtf(kc.x * kc.x + kc.y * kc.y);
    const float2 u = cmulc(make_float2(qc.x * iq, qc.y * iq), make_float2(kc.x * ik, kc.y * ik));   // e^{i(th_q - th_k)}
    const float2 vc = make_float2(fdn_rd(v.x), fdn_rd(v.y));                 // v' = rd(v * fft)
    const float sv = vc.x * vc.x + vc.y * vc.y, iv = rsqrtf(sv);
    if (role == 0) { const float m = sv * iv; return make_float2(m * u.x, m * u.y); }      // |v'| e^{i dtheta}
    if (role == 1) { const float s = A * iv; return make_float2(s * vc.x, s * vc.y); }      // |qk| e^{i angle v'}
    return make_float2(A * u.x, A * u.y);                                                   // |qk| e^{i dtheta}
}

// ---------------------------------------------------------------------------------------------------
// FDSA: three lanes (q, k, v roles) cooperate on one (channel, patch); bins are exchanged with shuffles
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_fdsa_patch(const float* __restrict__ hid, const float* __restrict__ wfft,
                                                    float* __restrict__ out, int E, int H, int W, long long nitems) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int t = lane / 3, role = lane - 3 * t;
    const long long item = warp * 10 + t;
    const bool valid = (lane < 30) && (item < nitems);
    const int pw = W >> 3, ph = H >> 3;
    float p[64];
    float2 S[8][5];
    size_t off_out = 0;
    int e = 0;
    if (valid) {
        int px = (int)(item % pw);
        long long r = item / pw;
        int py = (int)(r % ph);
        r /= ph;
        e = (int)(r % E);
        long long b = r / E;
        size_t sp = (size_t)(py * 8) * W + px * 8;
        load_patch(hid + ((size_t)b * 4 * E + role * E + e) * H * W + sp, W, p);
        off_out = ((size_t)b * 3 * E + role * E + e) * H * W + sp;
    } else {
#pragma unroll
        for (int i = 0; i < 64; ++i) p[i] = 0.f;
    }
    rfft2_8x8(p, S);
    const int l0 = 3 * t;
#pragma unroll
    for (int ky = 0; ky < 8; ++ky)
#pragma unroll
        for (int kx = 0; kx < 5; ++kx) {
            float wsel = (valid && role == 2) ? wfft[e * 40 + ky * 5 + kx] : 1.0f;
            float sx = S[ky][kx].x * wsel, sy = S[ky][kx].y * wsel;
            float2 q = make_float2(__shfl_sync(0xffffffffu, sx, l0), __shfl_sync(0xffffffffu, sy, l0));
            float2 k = make_float2(__shfl_sync(0xffffffffu, sx, l0 + 1), __shfl_sync(0xffffffffu, sy, l0 + 1));
            float2 v = make_float2(__shfl_sync(0xffffffffu, sx, l0 + 2), __shfl_sync(0xffffffffu, sy, l0 + 2));
            S[ky][kx] = fdsa_bin(q, k, v, role);
        }
    irfft2_8x8(S, p);
    if (valid) store_patch(out + off_out, W, p);
}

// ---------------------------------------------------------------------------------------------------
// FDSA with the depthwise 3x3 (to_hidden_dw, FDN_arch.py:563,578) fused in front: four lanes per (channel, patch) -
// q, k, v and v_value.  Every lane convolves its 10x10 halo window of the pre-dw hidden tensor in registers; the q/k/v
// lanes then run the spectral algebra exactly as above, the v_value lane stores its convolved patch for the gate.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_row10(const float* __restrict__ plane, int H, int W, int yy, int x0, float r[10]) {
    if (yy >= 0 && yy < H) {
        const float* p = plane + (size_t)yy * W + x0;
        const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
        r[0] = x0 > 0 ? p[-1] : 0.f;
        r[1] = a.x; r[2] = a.y; r[3] = a.z; r[4] = a.w; r[5] = b.x; r[6] = b.y; r[7] = b.z; r[8] = b.w;
        r[9] = x0 + 8 < W ? p[8] : 0.f;
    } else {
#pragma unroll
        for (int i = 0; i < 10; ++i) r[i] = 0.f;
    }
}

// p[8*y + x] = sum_{dy,dx} k[dy][dx] * in[y0 + y + dy - 1][x0 + x + dx - 1], zero outside the image
__device__ __forceinline__ void dw3_patch(const float* __restrict__ plane, const float* __restrict__ w, int H, int W, int y0, int x0,
                                          float p[64]) {
    float k[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) k[i] = w[i];
    float r0[10], r1[10], r2[10];
    load_row10(plane, H, W, y0 - 1, x0, r0);
    load_row10(plane, H, W, y0, x0, r1);
#pragma unroll
    for (int y = 0; y < 8; ++y) {
        load_row10(plane, H, W, y0 + y + 1, x0, r2);
#pragma unroll
        for (int x = 0; x < 8; ++x) {
            float a = k[0] * r0[x];
            a += k[1] * r0[x + 1]; a += k[2] * r0[x + 2];
            a += k[3] * r1[x]; a += k[4] * r1[x + 1]; a += k[5] * r1[x + 2];
            a += k[6] * r2[x]; a += k[7] * r2[x + 1]; a += k[8] * r2[x + 2];
            p[8 * y + x] = a;
        }
#pragma unroll
        for (int i = 0; i < 10; ++i) { r0[i] = r1[i]; r1[i] = r2[i]; }
    }
}

// p[8*y + x] += depthwise 3x3 of `plane` around the patch (rolling three input rows, zero outside the image)
__device__ __forceinline__ void dw3_patch_add(const float* __restrict__ plane, const float* __restrict__ w, int H, int W, int y0, int x0,
                                              float p[64]) {
    float k[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) k[i] = w[i];
    float r0[10], r1[10], r2[10];
    load_row10(plane, H, W, y0 - 1, x0, r0);
    load_row10(plane, H, W, y0, x0, r1);
#pragma unroll
    for (int y = 0; y < 8; ++y) {
        load_row10(plane, H, W, y0 + y + 1, x0, r2);
#pragma unroll
        for (int x = 0; x < 8; ++x) {
            float a = k[0] * r0[x];
            a += k[1] * r0[x + 1]; a += k[2] * r0[x + 2];
            a += k[3] * r1[x]; a += k[4] * r1[x + 1]; a += k[5] * r1[x + 2];
            a += k[6] * r2[x]; a += k[7] * r2[x + 1]; a += k[8] * r2[x + 2];
            p[8 * y + x] += a;
        }
#pragma unroll
        for (int i = 0; i < 10; ++i) { r0[i] = r1[i]; r1[i] = r2[i]; }
    }
}

// FDFFN: out = irfft2(rd(rfft2(h)) * wspec[c]) + dw_b(s1)      - the second depthwise conv of the spatial branch
// (space.2, FDN_arch.py:439-441,457) is evaluated on the patch from s1 with a one-pixel halo, so s2 never exists in HBM.
__global__ void __launch_bounds__(128, 4) k_fdffn_patch_dw(const float* __restrict__ h, const float* __restrict__ s1, const float* __restrict__ wb,
                                                        const float2* __restrict__ wspec, float* __restrict__ out, int C, int H, int W,
                                                        int per_plane) {
    // grid.y = plane (b*C + c): channel, spectral weights and depthwise taps are uniform over the CTA
    const int item = blockIdx.x * blockDim.x + threadIdx.x;     // over (H/8)*(W/8)
    if (item >= per_plane) return;
    const int pw = W >> 3;
    const int px = item % pw, py = item / pw;
    const size_t plane = blockIdx.y;
    const int c = blockIdx.y % C;
    size_t off = plane * H * W + (size_t)(py * 8) * W + px * 8;
    float p[64];
    {
        float2 S[8][5];
        load_patch(h + off, W, p);
        rfft2_8x8(p, S);
        const float2* w = wspec + c * 40;
#pragma unroll
        for (int ky = 0; ky < 8; ++ky)
#pragma unroll
            for (int kx = 0; kx < 5; ++kx) {
                float2 z = make_float2(fdn_rd(S[ky][kx].x), fdn_rd(S[ky][kx].y));
                S[ky][kx] = cmul(z, w[ky * 5 + kx]);
            }
        irfft2_8x8(S, p);
    }
    dw3_patch_add(s1 + (size_t)plane * H * W, wb + c * 9, H, W, py * 8, px * 8, p);
    store_patch(out + off, W, p);
}

// FDFFN middle section per patch, nothing through HBM (FDN_arch.py:457-470):
//     out = irfft2_8x8(rd(rfft2_8x8(h)) * wspec) + dw_b(gelu(dw_a(h)))
// One thread per (channel, patch).  The spectral branch is computed first into p[64].  The spatial branch then runs as a rolling
// pipeline over the twelve input rows of the patch's halo-2 window: three rows of h (12 wide) give one row of s1 = gelu(dw_a(h)) on
// the halo-1 window (10 wide, zero outside the image - it is dw_b's zero padding), and that row is scattered into the (up to) three
// output rows it contributes to, so only one row of s1 is ever live.  Against the two-kernel form (GELU depthwise kernel writing s1,
// k_fdffn_patch_dw reading it back with a halo) this removes a write and a read of the Hd-channel tensor per FDFFN at the price of
// evaluating gelu(dw_a) on 100 instead of 64 positions per patch.
__device__ __forceinline__ void load_row12(const float* __restrict__ plane, int H, int W, int yy, int x0, float r[12]) {
    if (yy >= 0 && yy < H) {
        const float* p = plane + (size_t)yy * W + x0;
        const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
        float2 l = make_float2(0.f, 0.f), rr = l;
        if (x0 > 0) l = *reinterpret_cast<const float2*>(p - 2);              // x0 is a multiple of 8: 8-byte aligned
        if (x0 + 8 < W) rr = *reinterpret_cast<const float2*>(p + 8);
        r[0] = l.x; r[1] = l.y;
        r[2] = a.x; r[3] = a.y; r[4] = a.z; r[5] = a.w; r[6] = b.x; r[7] = b.y; r[8] = b.z; r[9] = b.w;
        r[10] = rr.x; r[11] = rr.y;
    } else {
#pragma unroll
        for (int i = 0; i < 12; ++i) r[i] = 0.f;
    }
}

__global__ void __launch_bounds__(128, 3) k_fdffn_patch_fused(const float* __restrict__ h, const float* __restrict__ wa, const float* __restrict__ wb,
                                                              const float2* __restrict__ wspec, float* __restrict__ out, int C, int H, int W,
                                                              int per_plane) {
    // grid.y = plane (b*C + c): channel, spectral weights and depthwise taps are uniform over the CTA
    const int item = blockIdx.x * blockDim.x + threadIdx.x;     // over (H/8)*(W/8)
    if (item >= per_plane) return;
    const int pw = W >> 3;
    const int px = item % pw, py = item / pw;
    const size_t plane = blockIdx.y;
    const int c = blockIdx.y % C;
    const int x0 = px * 8, y0 = py * 8;
    const float* hp = h + plane * H * W;
    const size_t off = plane * H * W + (size_t)y0 * W + x0;
    float p[64];
    {
        float2 S[8][5];
        load_patch(h + off, W, p);
        rfft2_8x8(p, S);
        const float2* w = wspec + c * 40;
#pragma unroll
        for (int ky = 0; ky < 8; ++ky)
#pragma unroll
            for (int kx = 0; kx < 5; ++kx) {
                float2 z = make_float2(fdn_rd(S[ky][kx].x), fdn_rd(S[ky][kx].y));
                S[ky][kx] = cmul(z, w[ky * 5 + kx]);
            }
        irfft2_8x8(S, p);
    }
    float ka[9], kb[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) { ka[i] = wa[c * 9 + i]; kb[i] = wb[c * 9 + i]; }
    const bool colL = x0 > 0, colR = x0 + 8 < W;                 // s1 columns -1 and 8 lie inside the image
    float h0[12], h1[12], h2[12];
    load_row12(hp, H, W, y0 - 2, x0, h0);
    load_row12(hp, H, W, y0 - 1, x0, h1);
#pragma unroll
    for (int r = -1; r <= 8; ++r) {                              // s1 row r of the patch frame (image row y0 + r)
        load_row12(hp, H, W, y0 + r + 1, x0, h2);
        const int gy = y0 + r;
        if (gy >= 0 && gy < H) {                                 // warp-uniform except where a warp spans two patch rows
            // Two adjacent outputs per packed operation: the rows sit in even-aligned register pairs (that is how the 64- / 128-bit loads
            // deliver them), so for an even output column x the taps dx = 0 and dx = 2 read the aligned pairs (h[x], h[x+1]) and
            // (h[x+2], h[x+3]) - one FFMA2 each with the weight as a broadcast operand - and only the dx = 1 taps, whose pair would
            // straddle two register pairs, stay scalar: 12 instead of 18 issue slots per output pair, no operand moves.
            float2 s2[5];                                        // (s1 column 2i - 1, s1 column 2i)
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const int x = 2 * i;
                float2 a = f2mul_s(make_float2(h0[x], h0[x + 1]), ka[0]);
                a = cfma(make_float2(h0[x + 2], h0[x + 3]), ka[2], a);
                a = cfma(make_float2(h1[x], h1[x + 1]), ka[3], a);
                a = cfma(make_float2(h1[x + 2], h1[x + 3]), ka[5], a);
                a = cfma(make_float2(h2[x], h2[x + 1]), ka[6], a);
                a = cfma(make_float2(h2[x + 2], h2[x + 3]), ka[8], a);
                a.x = fmaf(ka[1], h0[x + 1], a.x); a.y = fmaf(ka[1], h0[x + 2], a.y);
                a.x = fmaf(ka[4], h1[x + 1], a.x); a.y = fmaf(ka[4], h1[x + 2], a.y);
                a.x = fmaf(ka[7], h2[x + 1], a.x); a.y = fmaf(ka[7], h2[x + 2], a.y);
                s2[i] = fdn_gelu2(a);
            }
            float s[10];
#pragma unroll
            for (int i = 0; i < 5; ++i) { s[2 * i] = s2[i].x; s[2 * i + 1] = s2[i].y; }
            if (!colL) s[0] = 0.f;
            if (!colR) s[9] = 0.f;
            // scatter: s1 row r is the (dy = r - y + 1) row of output row y, y = r - 1 .. r + 1; same pairing for the output columns
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                const int y = r + 1 - dy;
                if (y >= 0 && y < 8) {
#pragma unroll
                    for (int x = 0; x < 8; x += 2) {
                        float2 a = make_float2(p[8 * y + x], p[8 * y + x + 1]);
                        a = cfma(make_float2(s[x], s[x + 1]), kb[dy * 3], a);
                        a = cfma(make_float2(s[x + 2], s[x + 3]), kb[dy * 3 + 2], a);
                        a.x = fmaf(kb[dy * 3 + 1], s[x + 1], a.x);
                        a.y = fmaf(kb[dy * 3 + 1], s[x + 2], a.y);
                        p[8 * y + x] = a.x; p[8 * y + x + 1] = a.y;
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 12; ++i) { h0[i] = h1[i]; h1[i] = h2[i]; }
    }
    store_patch(out + off, W, p);
}

#ifndef FDSA_MIN_BLOCKS
#define FDSA_MIN_BLOCKS 4
#endif
#ifdef FDN_EMU
#define FDN_WARP_SYNC() __syncthreads()      // the emulator runs the lanes of a warp as free-running host threads
#else
#define FDN_WARP_SYNC() __syncwarp()
#endif
#define FDSA_IW 248        // floats of exchange space per item: 40 bins x 3 roles x float2 = 240, padded so that the eight items of a
                           // warp land on distinct bank pairs (every 64-bit access of the exchange is the 2-wavefront minimum)

// All three FDSA outputs of one bin (see fdsa_bin): out1 = |v'| e^{i dtheta}, out2 = |qk| e^{i angle v'}, out3 = |qk| e^{i dtheta}.
__device__ __forceinline__ void fdsa_bin3(float2 q, float2 k, float2 v, float2& o1, float2& o2, float2& o3) {
    float2 qk = cmul(q, k);
    qk.x = fdn_rd(qk.x);
    qk.y = fdn_rd(qk.y);
    const float s = qk.x * qk.x + qk.y * qk.y;
    const float A = s * rsqrtf(s);                                           // |rd(q k)|
    const float2 qc = make_float2(fdn_rd(q.x), fdn_rd(q.y)), kc = make_float2(fdn_rd(k.x), fdn_rd(k.y));
    const float iq = rsqrtf(qc.x * qc.x + qc.y * qc.y), ik = rsqrtf(kc.x * kc.x + kc.y * kc.y);
    const float2 u = cmulc(make_float2(qc.x * iq, qc.y * iq), make_float2(kc.x * ik, kc.y * ik));   // e^{i(th_q - th_k)}
    const float2 vc = make_float2(fdn_rd(v.x), fdn_rd(v.y));                 // v' = rd(v * fft)
    const float sv = vc.x * vc.x + vc.y * vc.y, iv = rsqrtf(sv);
    const float m = sv * iv, sc = A * iv;
    o1 = make_float2(m * u.x, m * u.y);
    o2 = make_float2(sc * vc.x, sc * vc.y);
    o3 = make_float2(A * u.x, A * u.y);
}

// Four lanes per (channel, patch): q, k, v and v_value.  Each lane convolves its 10x10 halo window (to_hidden_dw) in registers and
// the q/k/v lanes transform their patch.  The bin algebra needs q, k and v of a bin together, so the three spectra are exchanged
// through shared memory and the 40 bins are split over the four lanes (ten each, the v_value lane included): every lane evaluates
// all three outputs of its bins once - instead of every role lane re-deriving the shared moduli and phases of all 40 bins from
// 240 shuffles - and the role lanes read their output spectrum back for the inverse transform.
__global__ void __launch_bounds__(128, FDSA_MIN_BLOCKS) k_fdsa_patch_dw(const float* __restrict__ hid, const float* __restrict__ wdw,
                                                       const float* __restrict__ wfft, float* __restrict__ out, float* __restrict__ vv,
                                                       int E, int H, int W, int per_plane) {
    __shared__ __align__(16) float s_x[4 * 8 * FDSA_IW];
    const int lane = threadIdx.x & 31;
    const int t = lane >> 2, role = lane & 3;
    // grid.y = (image, channel e): uniform over the CTA; grid.x walks the patches of the plane, 32 per CTA
    const int item = (blockIdx.x * 4 + (threadIdx.x >> 5)) * 8 + t;
    const bool valid = item < per_plane;
    const int pw = W >> 3;
    const int e = blockIdx.y % E;
    const size_t b = blockIdx.y / E;
    float2* sx = reinterpret_cast<float2*>(s_x + ((threadIdx.x >> 5) * 8 + t) * FDSA_IW);
    float p[64];
    size_t off_out = 0;
    if (valid) {
        const int px = item % pw, py = item / pw;
        const int ch = role * E + e;
        dw3_patch(hid + ((size_t)b * 4 * E + ch) * H * W, wdw + ch * 9, H, W, py * 8, px * 8, p);
        const size_t sp = (size_t)(py * 8) * W + px * 8;
        if (role == 3) {
            store_patch(vv + ((size_t)b * E + e) * H * W + sp, W, p);
        } else {
            off_out = ((size_t)b * 3 * E + role * E + e) * H * W + sp;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 64; ++i) p[i] = 0.f;
    }
    {
        float2 S[8][5];
        rfft2_8x8(p, S);
        if (role < 3) {
#pragma unroll
            for (int ky = 0; ky < 8; ++ky)
#pragma unroll
                for (int kx = 0; kx < 5; ++kx) sx[(ky * 5 + kx) * 3 + role] = S[ky][kx];
        }
    }
    FDN_WARP_SYNC();
#pragma unroll
    for (int j = 0; j < 10; ++j) {
        const int bin = 4 * j + role;
        const float wv = valid ? wfft[e * 40 + bin] : 1.0f;
        const float2 q = sx[bin * 3], k = sx[bin * 3 + 1];
        float2 v = sx[bin * 3 + 2];
        v.x *= wv;
        v.y *= wv;
        float2 o1, o2, o3;
        fdsa_bin3(q, k, v, o1, o2, o3);
        sx[bin * 3] = o1;
        sx[bin * 3 + 1] = o2;
        sx[bin * 3 + 2] = o3;
    }
    FDN_WARP_SYNC();
    if (role < 3) {          // warp-divergent from here on: no synchronisation below
        float2 S[8][5];
#pragma unroll
        for (int ky = 0; ky < 8; ++ky)
#pragma unroll
            for (int kx = 0; kx < 5; ++kx) S[ky][kx] = sx[(ky * 5 + kx) * 3 + role];
        irfft2_8x8(S, p);
        if (valid) store_patch(out + off_out, W, p);
    }
}

// ---------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------
// out = irfft2_8x8( rd(rfft2_8x8(x)) * wspec[c] ) + add.   x, add, out: [B][C][H][W]; wspec: [C][8][5] complex.
FDN_API int fdn_fdffn_patch(const float* x, const float* add, const float* wspec, float* out, int B, int C, int H, int W,
                            cudaStream_t st) {
    FDN_REQUIRE(x && wspec && out && B > 0 && C > 0, "bad arguments");
    FDN_REQUIRE(H % 8 == 0 && W % 8 == 0, "H and W must be multiples of the 8x8 patch");
    FDN_REQUIRE(fdn_aligned16(x) && fdn_aligned16(out) && (!add || fdn_aligned16(add)), "pointers must be 16-byte aligned");
    long long n = (long long)B * C * (H / 8) * (W / 8);
    FDN_LAUNCH_SEQ(k_fdffn_patch, dim3(fdn_cdiv(n, 128)), dim3(128), 0, st, x, add, reinterpret_cast<const float2*>(wspec), out, C,
                   H, W, n);
    return fdn_check_launch("k_fdffn_patch");
}

// hid [B][4E][H][W] (q,k,v,v_value groups, after the depthwise 3x3); wfft [E][8][5]; out [B][3E][H][W] = (out1,out2,out3)
// before their LayerNorms.
FDN_API int fdn_fdsa_patch(const float* hid, const float* wfft, float* out, int B, int E, int H, int W, cudaStream_t st) {
    FDN_REQUIRE(hid && wfft && out && B > 0 && E > 0, "bad arguments");
    FDN_REQUIRE(H % 8 == 0 && W % 8 == 0, "H and W must be multiples of the 8x8 patch");
    FDN_REQUIRE(fdn_aligned16(hid) && fdn_aligned16(out), "pointers must be 16-byte aligned");
    long long n = (long long)B * E * (H / 8) * (W / 8);
    long long warps = (n + 9) / 10;
    FDN_LAUNCH(k_fdsa_patch, dim3(fdn_cdiv(warps, 4)), dim3(128), 0, st, hid, wfft, out, E, H, W, n);
    return fdn_check_launch("k_fdsa_patch");
}

// FDSA bin algebra with to_hidden_dw fused: hid [B][4E][H][W] is the *pre*-depthwise hidden tensor, wdw [4E][9] the
// depthwise weights.  out [B][3E][H][W] = (out1,out2,out3) before norm1..3, vv [B][E][H][W] = depthwise-convolved v_value.
FDN_API int fdn_fdsa_patch_dw(const float* hid, const float* wdw, const float* wfft, float* out, float* vv, int B, int E, int H, int W,
                              cudaStream_t st) {
    FDN_REQUIRE(hid && wdw && wfft && out && vv && B > 0 && E > 0, "bad arguments");
    FDN_REQUIRE(H % 8 == 0 && W % 8 == 0, "H and W must be multiples of the 8x8 patch");
    FDN_REQUIRE(fdn_aligned16(hid) && fdn_aligned16(out) && fdn_aligned16(vv), "pointers must be 16-byte aligned");
    FDN_REQUIRE((long long)B * E <= 65535, "too many planes for one launch");
    const int n = (H / 8) * (W / 8);
    FDN_LAUNCH(k_fdsa_patch_dw, dim3(fdn_cdiv(n, 32), B * E), dim3(128), 0, st, hid, wdw, wfft, out, vv, E, H, W, n);
    return fdn_check_launch("k_fdsa_patch_dw");
}

// FDFFN middle section in one kernel: out = dw_b(gelu(dw_a(h))) + irfft2_8x8(rd(rfft2_8x8(h)) * wspec)   (FDN_arch.py:457-470)
// h, out [B][C][H][W]; wa, wb [C][9] depthwise 3x3 weights (space.0, space.2); wspec [C][8][5] complex.
FDN_API int fdn_fdffn_spatial(const float* h, const float* wa, const float* wb, const float* wspec, float* out, int B, int C, int H, int W,
                              cudaStream_t st) {
    FDN_REQUIRE(h && wa && wb && wspec && out && B > 0 && C > 0, "bad arguments");
    FDN_REQUIRE(H % 8 == 0 && W % 8 == 0, "H and W must be multiples of the 8x8 patch");
    FDN_REQUIRE(fdn_aligned16(h) && fdn_aligned16(out), "h and out must be 16-byte aligned");
    FDN_REQUIRE((long long)B * C <= 65535, "too many planes for one launch");
    const int n = (H / 8) * (W / 8);
    FDN_LAUNCH_SEQ(k_fdffn_patch_fused, dim3(fdn_cdiv(n, 128), B * C), dim3(128), 0, st, h, wa, wb, reinterpret_cast<const float2*>(wspec), out, C,
                   H, W, n);
    return fdn_check_launch("k_fdffn_patch_fused");
}

// out = irfft2_8x8(rd(rfft2_8x8(h)) * wspec[c]) + depthwise3x3(s1; wb[c])   (FDFFN spectral branch + space.2, FDN_arch.py:439-441,457-470)
FDN_API int fdn_fdffn_patch_dw(const float* h, const float* s1, const float* wb, const float* wspec, float* out, int B, int C, int H, int W,
                               cudaStream_t st) {
    FDN_REQUIRE(h && s1 && wb && wspec && out && B > 0 && C > 0, "bad arguments");
    FDN_REQUIRE(H % 8 == 0 && W % 8 == 0, "H and W must be multiples of the 8x8 patch");
    FDN_REQUIRE(fdn_aligned16(h) && fdn_aligned16(s1) && fdn_aligned16(out), "pointers must be 16-byte aligned");
    FDN_REQUIRE((long long)B * C <= 65535, "too many planes for one launch");
    const int n = (H / 8) * (W / 8);
    FDN_LAUNCH_SEQ(k_fdffn_patch_dw, dim3(fdn_cdiv(n, 128), B * C), dim3(128), 0, st, h, s1, wb, reinterpret_cast<const float2*>(wspec), out, C,
                   H, W, n);
    return fdn_check_launch("k_fdffn_patch_dw");
}
