// Per-pixel kernels: 1x1 convolutions (channel GEMMs) with fused LayerNorm prologue and bias / activation /
// FiLM / residual / ratio epilogues, channel LayerNorms, resampling and the MAR gamma curve.
//
// Reference ops covered: every nn.Conv2d(k=1) in FDN_arch.py / fdnlol24_arch.py, LayerNorm
// (FDN_arch.py:326-342), torch.cat of nearest-resampled maps feeding fourier_fuse / FAM / Convs
// (FDN_arch.py:230-251), bilinear 0.5x / 2x (FDN_arch.py:719,730), PixelUnshuffle (:199-200) and the gamma
// curve 1-(1-x)^(40 i) (:282-284).
#include "fdn_common.cuh"

// ---------------------------------------------------------------------------------------------------
// 1x1 convolution, fp32 FFMA tile kernel
// ---------------------------------------------------------------------------------------------------
struct PwSrc {
    const float* p;   // [B][C][Hs][Ws]
    int C;
    int shift;        // 0: same size; s>0: source is 2^s smaller (nearest upsample); s<0: source is 2^-s larger (nearest down)
    int Hs, Ws;
};

struct PwParams {
    PwSrc src[3];
    int nsrc;
    int K, N, H, W, B;
    const float* wt;          // [K][N]  (transposed weight)
    const float* bias;        // [N] or null
    const float* ln_w;        // [K] LayerNorm over the K input channels (single unshifted source only), or null
    const float* ln_b;
    int act;                  // 0 none, 1 LeakyReLU(0.1), 2 ReLU
    const float* film_mul;    // [B][N][H][W] or null:  y = y*mul + add
    const float* film_add;
    const float* res;         // [B][N][H][W] or null:  y += res_coef*res
    float res_coef;
    const float* img_scale;   // [B] or null: y *= img_scale[b]   (applied last)
    float* out;               // element (b,n,y,x) at out[b*out_bs + n*out_ps + y*out_rs + x]
    long long out_bs, out_ps;
    int out_rs;
};

#define PW_TP 128
#define PW_KC 32

template <int NPT>
__global__ void __launch_bounds__(256) k_pw_conv(PwParams q) {
    constexpr int TN = 8 * NPT;
    __shared__ __align__(16) float Xs[PW_KC][PW_TP];
    __shared__ __align__(16) float Ws[PW_KC][TN];
    __shared__ float s_mu[PW_TP], s_rs[PW_TP];
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int b = blockIdx.z, n0 = blockIdx.y * TN, p0 = blockIdx.x * PW_TP;
    const int HW = q.H * q.W;
    const bool ln = q.ln_w != nullptr;

    if (ln) {
        if (tid < PW_TP) {
            int p = p0 + tid;
            float mu = 0.f, rs = 0.f;
            if (p < HW) {
                const float* xp = q.src[0].p + (size_t)b * q.K * HW + p;
                float s = 0.f;
                for (int k = 0; k < q.K; ++k) s += xp[(size_t)k * HW];
                mu = s / (float)q.K;
                float v = 0.f;
                for (int k = 0; k < q.K; ++k) {
                    float d = xp[(size_t)k * HW] - mu;
                    v += d * d;
                }
                rs = 1.0f / sqrtf(v / (float)q.K + 1e-5f);
            }
            s_mu[tid] = mu;
            s_rs[tid] = rs;
        }
        __syncthreads();
    }

    float acc[NPT][4];
#pragma unroll
    for (int i = 0; i < NPT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    // Staging: a thread always stages the same pixel (256 threads = 2 channels x 128 pixels), so the pixel's offset inside every source
    // - including the nearest-resampled ones of the torch.cat-fused convolutions - is computed once, not per element (a run-time
    // division and a source search per element made the 84-channel fourier_fuse convolution 3 ms per launch).
    const int spp = tid & (PW_TP - 1), skk0 = tid >> 7;
    const int sp = p0 + spp;
    const bool sp_ok = sp < HW;
    const float* sbase[3] = {nullptr, nullptr, nullptr};
    size_t sstride[3] = {0, 0, 0};
    int send[3] = {0, 0, 0}, sbeg[3] = {0, 0, 0};                  // channel range [sbeg, send) of source s
    {
        const int y = sp_ok ? sp / q.W : 0, x = sp_ok ? sp - y * q.W : 0;
        int cum = 0;
#pragma unroll
        for (int s = 0; s < 3; ++s)
            if (s < q.nsrc) {
                const PwSrc& S = q.src[s];
                size_t off;
                if (S.shift == 0) {
                    sstride[s] = (size_t)HW;
                    off = (size_t)(sp_ok ? sp : 0);
                } else {
                    const int ys = S.shift > 0 ? (y >> S.shift) : (y << -S.shift);
                    const int xs = S.shift > 0 ? (x >> S.shift) : (x << -S.shift);
                    sstride[s] = (size_t)S.Hs * S.Ws;
                    off = (size_t)ys * S.Ws + xs;
                }
                sbase[s] = S.p + (size_t)b * S.C * sstride[s] + off;
                sbeg[s] = cum;
                cum += S.C;
                send[s] = cum;
            }
    }
    for (int k0 = 0; k0 < q.K; k0 += PW_KC) {
#pragma unroll 4
        for (int kk = skk0; kk < PW_KC; kk += 2) {
            const int k = k0 + kk;
            float v = 0.f;
            if (k < q.K && sp_ok) {
                const int s = (k < send[0] || q.nsrc == 1) ? 0 : ((k < send[1] || q.nsrc == 2) ? 1 : 2);
                v = sbase[s][(size_t)(k - sbeg[s]) * sstride[s]];
                if (ln) v = (v - s_mu[spp]) * s_rs[spp] * q.ln_w[k] + q.ln_b[k];
            }
            Xs[kk][spp] = v;
        }
        for (int i = tid; i < PW_KC * TN; i += 256) {
            int kk = i / TN, nn = i - kk * TN;
            int k = k0 + kk, n = n0 + nn;
            Ws[kk][nn] = (k < q.K && n < q.N) ? q.wt[(size_t)k * q.N + n] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < PW_KC; ++kk) {
            float4 xv = *reinterpret_cast<const float4*>(&Xs[kk][4 * tx]);
            float wv[NPT];
#pragma unroll
            for (int i = 0; i < NPT; i += 4) {
                float4 w4 = *reinterpret_cast<const float4*>(&Ws[kk][ty * NPT + i]);
                wv[i] = w4.x; wv[i + 1] = w4.y; wv[i + 2] = w4.z; wv[i + 3] = w4.w;
            }
#pragma unroll
            for (int i = 0; i < NPT; ++i) {
                acc[i][0] += wv[i] * xv.x;
                acc[i][1] += wv[i] * xv.y;
                acc[i][2] += wv[i] * xv.z;
                acc[i][3] += wv[i] * xv.w;
            }
        }
        __syncthreads();
    }

    const float scale = q.img_scale ? q.img_scale[b] : 1.0f;
    const bool flat = (q.out_rs == q.W) && (q.out_ps == (long long)HW) && ((HW & 3) == 0);
    const int pbase = p0 + 4 * tx;
    size_t ooff[4];                                                 // strided output view: row / column of the four pixels, once
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int p = pbase + j < HW ? pbase + j : 0;
        const int y = p / q.W, x = p - y * q.W;
        ooff[j] = (size_t)y * q.out_rs + x;
    }
#pragma unroll
    for (int i = 0; i < NPT; ++i) {
        int n = n0 + ty * NPT + i;
        if (n >= q.N) continue;
        float bias = q.bias ? q.bias[n] : 0.f;
        size_t cidx = ((size_t)b * q.N + n) * HW;      // compact [B][N][H][W] index of film / res
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int p = pbase + j;
            float y = fdn_act(acc[i][j] + bias, q.act);
            if (p < HW) {
                if (q.film_mul) y = y * q.film_mul[cidx + p] + q.film_add[cidx + p];
                if (q.res) y += q.res_coef * q.res[cidx + p];
            }
            v[j] = y * scale;
        }
        if (flat && pbase + 3 < HW) {
            *reinterpret_cast<float4*>(q.out + (size_t)b * q.out_bs + (size_t)n * q.out_ps + pbase) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int p = pbase + j;
                if (p < HW) q.out[(size_t)b * q.out_bs + (size_t)n * q.out_ps + ooff[j]] = v[j];
            }
        }
    }
}

// Few output channels (N <= NO <= 16; the level-1 convolutions of MAR: 84 -> 12, 12 -> 12, 12 -> 3): a thread owns four adjacent pixels and
// ALL outputs.  Inputs come straight from global memory (one 128-bit load per channel, or four cached scalar loads for a
// nearest-resampled source), the weights of a channel are NO / 4 broadcast 128-bit shared-memory reads feeding 4 NO FMAs - no input
// staging, no barrier in the K loop, no output tile padded to 32 channels (k_pw_conv spent 3 ms on the 84 -> 12 fourier_fuse
// convolution).  Same accumulation order over k as k_pw_conv.
#define PWS_MAXK 256
template <int NO>
__global__ void __launch_bounds__(256) k_pw_conv_small(PwParams q) {
    __shared__ __align__(16) float sw[PWS_MAXK * NO];
    const int HW = q.H * q.W, b = blockIdx.z;
    for (int i = threadIdx.x; i < q.K * NO; i += blockDim.x) {
        const int k = i / NO, n = i - k * NO;
        sw[i] = n < q.N ? q.wt[(size_t)k * q.N + n] : 0.f;
    }
    __syncthreads();
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) * 4;             // HW % 4 == 0, W % 4 == 0: the four pixels share a row
    if (p >= HW) return;
    const int y = p / q.W, x = p - y * q.W;
    float acc[4][NO];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int n = 0; n < NO; ++n) acc[j][n] = 0.f;
    int kg = 0;
    for (int s = 0; s < q.nsrc; ++s) {
        const PwSrc& S = q.src[s];
        const float* wrow = sw + kg * NO;
        if (S.shift == 0) {
            const float* xp = S.p + (size_t)b * S.C * HW + p;
#pragma unroll 2
            for (int k = 0; k < S.C; ++k) {
                const float4 xv = *reinterpret_cast<const float4*>(xp + (size_t)k * HW);
                const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                for (int n4 = 0; n4 < NO; n4 += 4) {
                    const float4 w4 = *reinterpret_cast<const float4*>(wrow + k * NO + n4);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        acc[j][n4] += w4.x * xs[j]; acc[j][n4 + 1] += w4.y * xs[j];
                        acc[j][n4 + 2] += w4.z * xs[j]; acc[j][n4 + 3] += w4.w * xs[j];
                    }
                }
            }
        } else {
            const int ys = S.shift > 0 ? (y >> S.shift) : (y << -S.shift);
            int xo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) xo[j] = S.shift > 0 ? ((x + j) >> S.shift) : ((x + j) << -S.shift);
            const size_t plane = (size_t)S.Hs * S.Ws;
            const float* xp = S.p + (size_t)b * S.C * plane + (size_t)ys * S.Ws;
#pragma unroll 2
            for (int k = 0; k < S.C; ++k) {
                const float* r = xp + (size_t)k * plane;
                const float xs[4] = {r[xo[0]], r[xo[1]], r[xo[2]], r[xo[3]]};
#pragma unroll
                for (int n4 = 0; n4 < NO; n4 += 4) {
                    const float4 w4 = *reinterpret_cast<const float4*>(wrow + k * NO + n4);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        acc[j][n4] += w4.x * xs[j]; acc[j][n4 + 1] += w4.y * xs[j];
                        acc[j][n4 + 2] += w4.z * xs[j]; acc[j][n4 + 3] += w4.w * xs[j];
                    }
                }
            }
        }
        kg += S.C;
    }
    const float scale = q.img_scale ? q.img_scale[b] : 1.0f;
    const bool flat = (q.out_rs == q.W) && (q.out_ps == (long long)HW);
    const size_t ooff = (size_t)y * q.out_rs + x;
#pragma unroll
    for (int n = 0; n < NO; ++n) {
        if (n >= q.N) break;
        const float bias = q.bias ? q.bias[n] : 0.f;
        const size_t cidx = ((size_t)b * q.N + n) * HW + p;              // compact [B][N][H][W] index of film / res
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float t = fdn_act(acc[j][n] + bias, q.act);
            if (q.film_mul) t = t * q.film_mul[cidx + j] + q.film_add[cidx + j];
            if (q.res) t += q.res_coef * q.res[cidx + j];
            v[j] = t * scale;
        }
        float* o = q.out + (size_t)b * q.out_bs + (size_t)n * q.out_ps + ooff;
        if (flat && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
            *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = v[j];
        }
    }
}

// Generic 1x1 convolution.  Sources are concatenated along channels; src_shift[i] > 0 means source i is 2^s times
// smaller and is nearest-upsampled, < 0 means it is larger and is nearest-downsampled (x[::2^s, ::2^s]).
// wt is the weight transposed to [K][N].  Epilogue order: +bias, act, *film_mul+film_add, +res_coef*res, *img_scale[b].
FDN_API int fdn_pw_conv(const float* src0, int c0, int shift0, const float* src1, int c1, int shift1, const float* src2, int c2,
                        int shift2, const float* wt, const float* bias, const float* ln_w, const float* ln_b, int act,
                        const float* film_mul, const float* film_add, const float* res, float res_coef,
                        const float* img_scale, float* out, long long out_bs, long long out_ps, int out_rs, int B, int N,
                        int H, int W, cudaStream_t st) {
    FDN_REQUIRE(src0 && wt && out && B > 0 && N > 0 && H > 0 && W > 0 && c0 > 0, "bad arguments");
    FDN_REQUIRE((film_mul == nullptr) == (film_add == nullptr), "film needs both maps");
    PwParams q;
    const float* sp[3] = {src0, src1, src2};
    int sc[3] = {c0, c1, c2}, sh[3] = {shift0, shift1, shift2};
    q.nsrc = 0;
    q.K = 0;
    for (int i = 0; i < 3; ++i) {
        if (!sp[i] || sc[i] <= 0) break;
        PwSrc& S = q.src[q.nsrc++];
        S.p = sp[i];
        S.C = sc[i];
        S.shift = sh[i];
        if (sh[i] >= 0) {
            FDN_REQUIRE(H % (1 << sh[i]) == 0 && W % (1 << sh[i]) == 0, "upsampled source must divide the output size");
            S.Hs = H >> sh[i];
            S.Ws = W >> sh[i];
        } else {
            S.Hs = H << -sh[i];
            S.Ws = W << -sh[i];
        }
        q.K += sc[i];
    }
    for (int i = q.nsrc; i < 3; ++i) q.src[i] = q.src[0];
    if (ln_w) FDN_REQUIRE(ln_b && q.nsrc == 1 && shift0 == 0, "LayerNorm prologue needs one unshifted source");
    q.N = N; q.H = H; q.W = W; q.B = B;
    q.wt = wt; q.bias = bias; q.ln_w = ln_w; q.ln_b = ln_b; q.act = act;
    q.film_mul = film_mul; q.film_add = film_add; q.res = res; q.res_coef = res_coef; q.img_scale = img_scale;
    q.out = out; q.out_bs = out_bs; q.out_ps = out_ps; q.out_rs = out_rs;
    FDN_REQUIRE(fdn_aligned16(out) || !(out_rs == W && out_ps == (long long)H * W), "output must be 16-byte aligned");
    int HW = H * W;
    if (N <= 16 && ln_w == nullptr && q.K <= PWS_MAXK && W % 4 == 0) {
        dim3 grid(fdn_cdiv(HW / 4, 256), 1, B);
        if (N <= 4) { auto k = k_pw_conv_small<4>; FDN_LAUNCH(k, grid, dim3(256), 0, st, q); }
        else if (N <= 8) { auto k = k_pw_conv_small<8>; FDN_LAUNCH(k, grid, dim3(256), 0, st, q); }
        else if (N <= 12) { auto k = k_pw_conv_small<12>; FDN_LAUNCH(k, grid, dim3(256), 0, st, q); }
        else { auto k = k_pw_conv_small<16>; FDN_LAUNCH(k, grid, dim3(256), 0, st, q); }
        return fdn_check_launch("k_pw_conv_small");
    }
    if (N <= 32) {
        auto k = k_pw_conv<4>;
        FDN_LAUNCH(k, dim3(fdn_cdiv(HW, PW_TP), fdn_cdiv(N, 32), B), dim3(256), 0, st, q);
    } else {
        auto k = k_pw_conv<8>;
        FDN_LAUNCH(k, dim3(fdn_cdiv(HW, PW_TP), fdn_cdiv(N, 64), B), dim3(256), 0, st, q);
    }
    return fdn_check_launch("k_pw_conv");
}

// ---------------------------------------------------------------------------------------------------
// grouped channel LayerNorm:  out[b, g*C + c, p] = LN_g(in[b, g*C + c, p]) * mul[b, c, p] + add[b, c, p]
// ---------------------------------------------------------------------------------------------------
struct ChanLnParams {
    const float* in;      // [B][G*C][HW]
    float* out;           // same shape (may alias in)
    const float* gamma;   // [G][C]
    const float* beta;    // [G][C]
    const float* mul;     // element (b,c,p) at mul[b*mul_bs + c*HW + p], or null
    const float* add;     // same addressing with add_bs, or null
    long long mul_bs, add_bs;
    int G, C, HW;
    long long total;      // B*HW
};

__global__ void __launch_bounds__(256) k_chan_ln(ChanLnParams q) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= q.total) return;
    int b = (int)(i / q.HW), p = (int)(i - (long long)b * q.HW);
    for (int g = 0; g < q.G; ++g) {
        const float* xp = q.in + ((size_t)b * q.G + g) * q.C * q.HW + p;
        float* op = q.out + ((size_t)b * q.G + g) * q.C * q.HW + p;
        float s = 0.f;
        for (int c = 0; c < q.C; ++c) s += xp[(size_t)c * q.HW];
        float mu = s / (float)q.C, v = 0.f;
        for (int c = 0; c < q.C; ++c) {
            float d = xp[(size_t)c * q.HW] - mu;
            v += d * d;
        }
        float rs = 1.0f / sqrtf(v / (float)q.C + 1e-5f);
        for (int c = 0; c < q.C; ++c) {
            float y = (xp[(size_t)c * q.HW] - mu) * rs * q.gamma[g * q.C + c] + q.beta[g * q.C + c];
            if (q.mul) y *= q.mul[(size_t)b * q.mul_bs + (size_t)c * q.HW + p];
            if (q.add) y += q.add[(size_t)b * q.add_bs + (size_t)c * q.HW + p];
            op[(size_t)c * q.HW] = y;
        }
    }
}

FDN_API int fdn_chan_ln(const float* in, float* out, const float* gamma, const float* beta, const float* mul, long long mul_bs,
                        const float* add, long long add_bs, int B, int G, int C, int HW, cudaStream_t st) {
    FDN_REQUIRE(in && out && gamma && beta && B > 0 && G > 0 && C > 0 && HW > 0, "bad arguments");
    ChanLnParams q;
    q.in = in; q.out = out; q.gamma = gamma; q.beta = beta; q.mul = mul; q.add = add;
    q.mul_bs = mul_bs; q.add_bs = add_bs; q.G = G; q.C = C; q.HW = HW; q.total = (long long)B * HW;
    FDN_LAUNCH_SEQ(k_chan_ln, dim3(fdn_cdiv(q.total, 256)), dim3(256), 0, st, q);
    return fdn_check_launch("k_chan_ln");
}

// ---------------------------------------------------------------------------------------------------
// resampling / layout kernels (planes = B*C)
// ---------------------------------------------------------------------------------------------------
__global__ void k_avgpool2(const float* __restrict__ in, float* __restrict__ out, int H, int W, long long total) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // over planes*Ho*Wo
    if (i >= total) return;
    int Wo = W >> 1, Ho = H >> 1;
    int x = (int)(i % Wo);
    long long t = i / Wo;
    int y = (int)(t % Ho);
    long long pl = t / Ho;
    const float* p = in + ((size_t)pl * H + 2 * y) * W + 2 * x;
    out[i] = ((p[0] + p[1]) + (p[W] + p[W + 1])) * 0.25f;
}

__global__ void k_up2_bilinear(const float* __restrict__ in, float* __restrict__ out, int H, int W, long long total) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // over planes*2H*2W
    if (i >= total) return;
    int Wo = 2 * W, Ho = 2 * H;
    int x = (int)(i % Wo);
    long long t = i / Wo;
    int y = (int)(t % Ho);
    long long pl = t / Ho;
    // align_corners=False: src = (dst+0.5)/2 - 0.5, clamped at 0
    float sy = fmaxf(0.5f * (y + 0.5f) - 0.5f, 0.f), sx = fmaxf(0.5f * (x + 0.5f) - 0.5f, 0.f);
    int y0 = (int)sy, x0 = (int)sx;
    int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
    float ly = sy - y0, lx = sx - x0;
    const float* p = in + (size_t)pl * H * W;
    float top = (1.f - lx) * p[(size_t)y0 * W + x0] + lx * p[(size_t)y0 * W + x1];
    float bot = (1.f - lx) * p[(size_t)y1 * W + x0] + lx * p[(size_t)y1 * W + x1];
    out[i] = (1.f - ly) * top + ly * bot;
}

// four horizontally adjacent outputs per thread (W even): the same taps and the same arithmetic per output as k_up2_bilinear, with the
// row / plane decomposition, the two source rows and the 128-bit store shared by the four (the scalar kernel ran at 1 TB/s on index
// arithmetic: three run-time divisions and four scattered loads per output)
__global__ void __launch_bounds__(256) k_up2_bilinear4(const float* __restrict__ in, float* __restrict__ out, int H, int W, long long total4) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // over planes*2H*(2W/4)
    if (i >= total4) return;
    const int Wo = 2 * W, Ho = 2 * H, Wq = Wo >> 2;
    const int xq = (int)(i % Wq);
    long long t = i / Wq;
    const int y = (int)(t % Ho);
    const long long pl = t / Ho;
    const float sy = fmaxf(0.5f * (y + 0.5f) - 0.5f, 0.f);
    const int y0 = (int)sy, y1 = min(y0 + 1, H - 1);
    const float ly = sy - y0;
    const float* r0 = in + ((size_t)pl * H + y0) * W;
    const float* r1 = in + ((size_t)pl * H + y1) * W;
    // outputs 4 xq .. 4 xq + 3 read source columns 2 xq - 1 .. 2 xq + 2 (clamped)
    const int c = 2 * xq;
    const int cm = max(c - 1, 0), cp = min(c + 2, W - 1);
    const float a0 = r0[cm], a1 = r0[c], a2 = r0[c + 1], a3 = r0[cp];
    const float b0 = r1[cm], b1 = r1[c], b2 = r1[c + 1], b3 = r1[cp];
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int x = 4 * xq + j;
        const float sx = fmaxf(0.5f * (x + 0.5f) - 0.5f, 0.f);
        const int x0 = (int)sx;
        const float lx = sx - x0;
        // (x0, x1) is (cm, c), (c, c + 1), (c, c + 1), (c + 1, cp) for j = 0..3; at the left border x0 = x1 = 0 = cm = c
        const float t0 = j == 0 ? a0 : (j == 3 ? a2 : a1), t1 = j == 0 ? a1 : (j == 3 ? a3 : a2);
        const float u0 = j == 0 ? b0 : (j == 3 ? b2 : b1), u1 = j == 0 ? b1 : (j == 3 ? b3 : b2);
        const float top = (1.f - lx) * t0 + lx * t1;
        const float bot = (1.f - lx) * u0 + lx * u1;
        o[j] = (1.f - ly) * top + ly * bot;
    }
    *reinterpret_cast<float4*>(out + ((size_t)pl * Ho + y) * Wo + 4 * xq) = make_float4(o[0], o[1], o[2], o[3]);
}

__global__ void k_pixel_unshuffle(const float* __restrict__ in, float* __restrict__ out, int C, int H, int W, int r, long long total) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // over B*C*r*r*(H/r)*(W/r)
    if (i >= total) return;
    int Wo = W / r, Ho = H / r;
    int x = (int)(i % Wo);
    long long t = i / Wo;
    int y = (int)(t % Ho);
    t /= Ho;
    int co = (int)(t % (C * r * r));
    long long b = t / (C * r * r);
    int c = co / (r * r), ij = co - c * r * r, ii = ij / r, jj = ij - ii * r;
    out[i] = in[(((size_t)b * C + c) * H + (size_t)y * r + ii) * W + (size_t)x * r + jj];
}

// out = 1 - (1 - x)^(40 * i)      (torch.pow semantics, x in [0,1])
__global__ void k_gamma(const float* __restrict__ x, const float* __restrict__ illum, float* __restrict__ out, float scale, long long total) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    out[i] = 1.0f - powf(1.0f - x[i], illum[i] * scale);
}

// border of a [planes][H][W] tensor <- value[plane % C]   (fourier_fuse.fpre[1]: 1x1 depthwise conv with padding 1)
__global__ void k_fill_border(float* __restrict__ t, const float* __restrict__ value, int C, int H, int W, long long total) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // over planes*(2W + 2H)
    if (i >= total) return;
    int per = 2 * W + 2 * H;
    long long pl = i / per;
    int e = (int)(i - pl * per);
    int y, x;
    if (e < W) { y = 0; x = e; }
    else if (e < 2 * W) { y = H - 1; x = e - W; }
    else if (e < 2 * W + H) { y = e - 2 * W; x = 0; }
    else { y = e - 2 * W - H; x = W - 1; }
    t[((size_t)pl * H + y) * W + x] = value[pl % C];
}

FDN_API int fdn_avgpool2(const float* in, float* out, int planes, int H, int W, cudaStream_t st) {
    FDN_REQUIRE(in && out && planes > 0 && H % 2 == 0 && W % 2 == 0, "bad arguments");
    long long n = (long long)planes * (H / 2) * (W / 2);
    FDN_LAUNCH_SEQ(k_avgpool2, dim3(fdn_cdiv(n, 256)), dim3(256), 0, st, in, out, H, W, n);
    return fdn_check_launch("k_avgpool2");
}
FDN_API int fdn_up2_bilinear(const float* in, float* out, int planes, int H, int W, cudaStream_t st) {
    FDN_REQUIRE(in && out && planes > 0 && H > 0 && W > 0, "bad arguments");
    long long n = (long long)planes * H * W * 4;
    if (W % 2 == 0 && W >= 2 && fdn_aligned16(out)) {
        FDN_LAUNCH_SEQ(k_up2_bilinear4, dim3(fdn_cdiv(n / 4, 256)), dim3(256), 0, st, in, out, H, W, n / 4);
        return fdn_check_launch("k_up2_bilinear4");
    }
    FDN_LAUNCH_SEQ(k_up2_bilinear, dim3(fdn_cdiv(n, 256)), dim3(256), 0, st, in, out, H, W, n);
    return fdn_check_launch("k_up2_bilinear");
}
FDN_API int fdn_pixel_unshuffle(const float* in, float* out, int B, int C, int H, int W, int r, cudaStream_t st) {
    FDN_REQUIRE(in && out && B > 0 && C > 0 && r > 0 && H % r == 0 && W % r == 0, "bad arguments");
    long long n = (long long)B * C * H * W;
    FDN_LAUNCH_SEQ(k_pixel_unshuffle, dim3(fdn_cdiv(n, 256)), dim3(256), 0, st, in, out, C, H, W, r, n);
    return fdn_check_launch("k_pixel_unshuffle");
}
FDN_API int fdn_gamma_curve(const float* x, const float* illum, float* out, float scale, long long n, cudaStream_t st) {
    FDN_REQUIRE(x && illum && out && n > 0, "bad arguments");
    FDN_LAUNCH_SEQ(k_gamma, dim3(fdn_cdiv(n, 256)), dim3(256), 0, st, x, illum, out, scale, n);
    return fdn_check_launch("k_gamma");
}
FDN_API int fdn_fill_border(float* t, const float* value, int planes, int C, int H, int W, cudaStream_t st) {
    FDN_REQUIRE(t && value && planes > 0 && C > 0 && H > 1 && W > 1, "bad arguments");
    long long n = (long long)planes * (2 * W + 2 * H);
    FDN_LAUNCH_SEQ(k_fill_border, dim3(fdn_cdiv(n, 256)), dim3(256), 0, st, t, value, C, H, W, n);
    return fdn_check_launch("k_fill_border");
}
