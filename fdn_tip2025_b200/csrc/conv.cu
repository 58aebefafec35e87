// Spatial convolutions of the hot path, fp32 FFMA:
//   dense KxK (3x3 s1/s2, 7x7 s2, strided 1x1) with bias / activation / residual / sigmoid-head epilogues
//       FDformer patch_embed, output, Downsample/Upsample convs   FDN_arch.py:704, 804, 720, 731
//       MAR f3_down/f2_down, FAM.merge2, ConvsOut, out, fourier_out FDN_arch.py:192-196, 57, 135, 241-255
//       FCAFFN conv3_{mul,add}(conv1_{mul,add}(img)) folded into one 3->C 3x3 conv   FDN_arch.py:423
//       LPNet convs with BatchNorm folded                           LPNet_arch.py:46-61, 90-97
//   ConvTranspose2d 4x4 stride 2 pad 1 + LeakyReLU                  FDN_arch.py:21-23, 194-195
//   depthwise 3x3 (plain, +GELU, or C->2C with the GELU gate)       FDN_arch.py:426-427, 435-441, 448, 472-473, 563
#include "fdn_common.cuh"

struct ConvParams {
    const float* in;      // [B][Cin][Hin][Win]
    const float* w;       // [Cout][Cin][K][K]
    const float* bias;    // [Cout] or null
    const float* res;     // residual, element (b,co,y,x) at res[((b*Cout+co)*Hr + (y<<res_shift))*Wr + (x<<res_shift)], or null
    float* out;           // [B][Cout][Hout][Wout]
    int B, Cin, Cout, Hin, Win, Hout, Wout;
    int pad;
    int act;              // 0 none, 1 LeakyReLU(0.1), 2 ReLU   (applied to conv+bias)
    int head;             // 0: y = act(conv+bias) + res ; 1: y = sigmoid(conv + bias + res) + 1e-8
    int res_shift, Hr, Wr;
    int cc;               // input channels staged per iteration
};

#define CV_TX 16   // threads along x, 4 outputs each
#define CV_TY 16   // threads along y

// NO = output channels per CTA (16 for the wide FDformer convs: the staged input tile is reused twice as often and a thread has
// 64 independent accumulators; 8 for the narrow MAR / LPNet layers).  The staged input rows are padded to a multiple of four
// floats so that, at stride 1, a thread reads its K+3 wide window with one conflict-free 128-bit load plus the K-1 values after
// it - once per kernel row, reused by all K taps - instead of four 4-way-conflicting scalar loads per tap.
template <int K, int S, int NO>
__global__ void __launch_bounds__(256) k_conv2d(ConvParams q) {
    FDN_DYN_SMEM(smem);
    constexpr int TW = CV_TX * 4, TH = CV_TY;                 // output tile
    constexpr int IW = ((TW - 1) * S + K + 3) & ~3, IH = (TH - 1) * S + K;   // input tile (row stride padded to 16 bytes)
    float* Ws = reinterpret_cast<float*>(smem);               // [cc][K][K][NO]  (first: keeps float4 reads aligned)
    float* Is = Ws + q.cc * K * K * NO;                       // [cc][IH][IW]
    const int tid = threadIdx.x, tx = tid % CV_TX, ty = tid / CV_TX;
    const int tiles_x = (q.Wout + TW - 1) / TW;
    const int ox0 = (blockIdx.x % tiles_x) * TW, oy0 = (blockIdx.x / tiles_x) * TH;
    const int co0 = blockIdx.y * NO, b = blockIdx.z;
    const int ix0 = ox0 * S - q.pad, iy0 = oy0 * S - q.pad;
    float acc[NO][4];
#pragma unroll
    for (int o = 0; o < NO; ++o)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[o][j] = 0.f;

    for (int c0 = 0; c0 < q.Cin; c0 += q.cc) {
        const int nc = min(q.cc, q.Cin - c0);
        for (int i = tid; i < nc * IH * IW; i += 256) {
            int c = i / (IH * IW), r = i - c * (IH * IW);
            int yy = r / IW, xx = r - yy * IW;
            int gy = iy0 + yy, gx = ix0 + xx;
            float v = 0.f;
            if (gy >= 0 && gy < q.Hin && gx >= 0 && gx < q.Win)
                v = q.in[(((size_t)b * q.Cin + c0 + c) * q.Hin + gy) * q.Win + gx];
            Is[i] = v;
        }
        for (int i = tid; i < nc * K * K * NO; i += 256) {
            int o = i % NO, r = i / NO;
            int kk = r % (K * K), c = r / (K * K);
            int co = co0 + o;
            Ws[i] = co < q.Cout ? q.w[((size_t)co * q.Cin + c0 + c) * K * K + kk] : 0.f;
        }
        __syncthreads();
        for (int c = 0; c < nc; ++c) {
            const float* ip = Is + c * IH * IW + (ty * S) * IW + (tx * 4) * S;
            const float* wp = Ws + c * K * K * NO;
#pragma unroll
            for (int ky = 0; ky < K; ++ky) {
                float win[S == 1 ? K + 3 : 1];
                if (S == 1) {
                    const float4 w4 = *reinterpret_cast<const float4*>(ip + ky * IW);
                    win[0] = w4.x; win[1] = w4.y; win[2] = w4.z; win[3] = w4.w;
#pragma unroll
                    for (int j = 4; j < K + 3; ++j) win[S == 1 ? j : 0] = ip[ky * IW + j];
                }
#pragma unroll
                for (int kx = 0; kx < K; ++kx) {
                    float wv[NO];
#pragma unroll
                    for (int o = 0; o < NO; o += 4) {
                        const float4 w4 = *reinterpret_cast<const float4*>(wp + (ky * K + kx) * NO + o);
                        wv[o] = w4.x; wv[o + 1] = w4.y; wv[o + 2] = w4.z; wv[o + 3] = w4.w;
                    }
                    float xv[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) xv[j] = S == 1 ? win[S == 1 ? j + kx : 0] : ip[ky * IW + j * S + kx];
#pragma unroll
                    for (int o = 0; o < NO; ++o)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[o][j] += wv[o] * xv[j];
                }
            }
        }
        __syncthreads();
    }
    const int oy = oy0 + ty;
    if (oy >= q.Hout) return;
#pragma unroll
    for (int o = 0; o < NO; ++o) {
        int co = co0 + o;
        if (co >= q.Cout) continue;
        float bias = q.bias ? q.bias[co] : 0.f;
        const int oxb = ox0 + tx * 4;
        float* orow = q.out + (((size_t)b * q.Cout + co) * q.Hout + oy) * q.Wout;
        if (!q.res && q.head == 0 && (q.Wout & 3) == 0 && oxb + 3 < q.Wout) {      // common case: one 128-bit store
            *reinterpret_cast<float4*>(orow + oxb) = make_float4(fdn_act(acc[o][0] + bias, q.act), fdn_act(acc[o][1] + bias, q.act),
                                                                  fdn_act(acc[o][2] + bias, q.act), fdn_act(acc[o][3] + bias, q.act));
            continue;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int ox = oxb + j;
            if (ox >= q.Wout) continue;
            float v = acc[o][j] + bias;
            float r = 0.f;
            if (q.res) r = q.res[(((size_t)b * q.Cout + co) * q.Hr + ((size_t)oy << q.res_shift)) * q.Wr + ((size_t)ox << q.res_shift)];
            if (q.head == 1) v = fdn_sigmoid(v + r) + 1e-8f;
            else v = fdn_act(v, q.act) + r;
            orow[ox] = v;
        }
    }
}

// any K / stride: one thread per output element (used for the strided 1x1 convs of LPNet)
__global__ void k_conv2d_naive(ConvParams q, int K, int S, long long total) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int ox = (int)(i % q.Wout);
    long long t = i / q.Wout;
    int oy = (int)(t % q.Hout);
    t /= q.Hout;
    int co = (int)(t % q.Cout);
    int b = (int)(t / q.Cout);
    float acc = q.bias ? q.bias[co] : 0.f;
    for (int c = 0; c < q.Cin; ++c)
        for (int ky = 0; ky < K; ++ky) {
            int gy = oy * S - q.pad + ky;
            if (gy < 0 || gy >= q.Hin) continue;
            for (int kx = 0; kx < K; ++kx) {
                int gx = ox * S - q.pad + kx;
                if (gx < 0 || gx >= q.Win) continue;
                acc += q.w[(((size_t)co * q.Cin + c) * K + ky) * K + kx] * q.in[(((size_t)b * q.Cin + c) * q.Hin + gy) * q.Win + gx];
            }
        }
    float r = 0.f;
    if (q.res) r = q.res[(((size_t)b * q.Cout + co) * q.Hr + ((size_t)oy << q.res_shift)) * q.Wr + ((size_t)ox << q.res_shift)];
    float v = q.head == 1 ? fdn_sigmoid(acc + r) + 1e-8f : fdn_act(acc, q.act) + r;
    q.out[i] = v;
}

template <int K, int S, int NO>
static int launch_conv_no(ConvParams& q, cudaStream_t st) {
    constexpr int IW = ((CV_TX * 4 - 1) * S + K + 3) & ~3, IH = (CV_TY - 1) * S + K;
    size_t per_c = (size_t)(IH * IW + K * K * NO) * sizeof(float);
    int cc = (int)min((size_t)q.Cin, max((size_t)1, (size_t)(64 * 1024) / per_c));
    q.cc = cc;
    size_t smem = per_c * cc;
    auto kern = k_conv2d<K, S, NO>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { fdn_set_error(cudaGetErrorString(e)); return (int)e; }
    }
    int tiles = fdn_cdiv(q.Wout, CV_TX * 4) * fdn_cdiv(q.Hout, CV_TY);
    FDN_LAUNCH(kern, dim3(tiles, fdn_cdiv(q.Cout, NO), q.B), dim3(256), smem, st, q);
    return fdn_check_launch("k_conv2d");
}
template <int K, int S>
static int launch_conv(ConvParams& q, cudaStream_t st) {
    if (K == 3 && q.Cout >= 16 && q.Cout % 16 == 0) return launch_conv_no<K, S, 16>(q, st);
    if (K == 3 && S == 1 && q.Cout <= 4) return launch_conv_no<K, S, 4>(q, st);      // RGB heads (32 -> 3, 12 -> 3): no 8-wide output tile to pad
    return launch_conv_no<K, S, 8>(q, st);
}

// Dense KxK convolution, groups=1.  y = act(conv(x)+bias) + res           (head = 0)
//                                   y = sigmoid(conv(x)+bias+res) + 1e-8  (head = 1, MAR illumination heads)
// res (optional) has spatial size (Hout<<res_shift, Wout<<res_shift) and is sampled at (y<<res_shift, x<<res_shift).
FDN_API int fdn_conv2d(const float* in, const float* w, const float* bias, const float* res, int res_shift, float* out, int B,
                       int Cin, int Hin, int Win, int Cout, int K, int stride, int pad, int act, int head, cudaStream_t st) {
    FDN_REQUIRE(in && w && out && B > 0 && Cin > 0 && Cout > 0 && K > 0 && stride > 0 && pad >= 0, "bad arguments");
    ConvParams q;
    q.in = in; q.w = w; q.bias = bias; q.res = res; q.out = out;
    q.B = B; q.Cin = Cin; q.Cout = Cout; q.Hin = Hin; q.Win = Win;
    q.Hout = (Hin + 2 * pad - K) / stride + 1;
    q.Wout = (Win + 2 * pad - K) / stride + 1;
    FDN_REQUIRE(q.Hout > 0 && q.Wout > 0, "empty output");
    q.pad = pad; q.act = act; q.head = head; q.res_shift = res_shift;
    q.Hr = q.Hout << res_shift; q.Wr = q.Wout << res_shift;
    q.cc = 1;
    if (K == 3 && stride == 1) return launch_conv<3, 1>(q, st);
    if (K == 3 && stride == 2) return launch_conv<3, 2>(q, st);
    if (K == 7 && stride == 2) return launch_conv<7, 2>(q, st);
    long long total = (long long)B * Cout * q.Hout * q.Wout;
    FDN_LAUNCH_SEQ(k_conv2d_naive, dim3(fdn_cdiv(total, 256)), dim3(256), 0, st, q, K, stride, total);
    return fdn_check_launch("k_conv2d_naive");
}

// ---------------------------------------------------------------------------------------------------
// FCAFFN FiLM maps (FDN_arch.py:423): mul = conv3_mul(conv1_mul(img)), add = conv3_add(conv1_add(img)), both folded on the
// host into dense 3->C 3x3 kernels.  One launch produces both maps: a thread keeps the 3x3x6 neighbourhood of its four
// pixels in registers and loops over 8 output channels (weights broadcast from shared memory).
// ---------------------------------------------------------------------------------------------------
#define FILM_CG 8
__global__ void __launch_bounds__(256) k_film_maps(const float* __restrict__ img, const float* __restrict__ wmul, const float* __restrict__ wadd,
                                                   float* __restrict__ omul, float* __restrict__ oadd, int C, int H, int W) {
    __shared__ __align__(16) float2 sw[FILM_CG][28];         // 27 taps (+1 pad) per channel: (mul weight, add weight)
    const int c0 = blockIdx.y * FILM_CG, b = blockIdx.z;
    for (int i = threadIdx.x; i < FILM_CG * 27; i += blockDim.x) {
        const int c = i / 27, t = i - c * 27;
        sw[c][t] = (c0 + c < C) ? make_float2(wmul[(c0 + c) * 27 + t], wadd[(c0 + c) * 27 + t]) : make_float2(0.f, 0.f);
    }
    __syncthreads();
    const int W4 = W >> 2;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // over H * W/4
    if (i >= (long long)H * W4) return;
    const int x0 = (int)(i % W4) * 4, y = (int)(i / W4);
    float nb[3][3][6];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int yy = y + dy - 1;
            if (yy >= 0 && yy < H) {
                const float* p = img + (((size_t)b * 3 + j) * H + yy) * W + x0;
                const float4 m = *reinterpret_cast<const float4*>(p);
                nb[j][dy][0] = x0 > 0 ? p[-1] : 0.f;
                nb[j][dy][1] = m.x; nb[j][dy][2] = m.y; nb[j][dy][3] = m.z; nb[j][dy][4] = m.w;
                nb[j][dy][5] = x0 + 4 < W ? p[4] : 0.f;
            } else {
#pragma unroll
                for (int dx = 0; dx < 6; ++dx) nb[j][dy][dx] = 0.f;
            }
        }
    // the two maps of a channel read the same 27 inputs: one packed FFMA2 per tap and pixel, (mul, add) += (w_mul, w_add) * input,
    // and one 64-bit weight read per tap instead of two
    for (int c = 0; c < FILM_CG && c0 + c < C; ++c) {
        float2 o[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const float2 wv = sw[c][(j * 3 + dy) * 3 + dx];
#pragma unroll
                    for (int px = 0; px < 4; ++px) o[px] = cfma(wv, nb[j][dy][px + dx], o[px]);
                }
        const size_t at = (((size_t)b * C + c0 + c) * H + y) * W + x0;
        *reinterpret_cast<float4*>(omul + at) = make_float4(o[0].x, o[1].x, o[2].x, o[3].x);
        *reinterpret_cast<float4*>(oadd + at) = make_float4(o[0].y, o[1].y, o[2].y, o[3].y);
    }
}

// mul / add [B][C][H][W] = 3x3 conv (padding 1, no bias) of img [B][3][H][W] with wmul / wadd [C][3][3][3]
FDN_API int fdn_film_maps(const float* img, const float* wmul, const float* wadd, float* omul, float* oadd, int B, int C, int H, int W,
                          cudaStream_t st) {
    FDN_REQUIRE(img && wmul && wadd && omul && oadd && B > 0 && C > 0 && H > 0, "bad arguments");
    FDN_REQUIRE(W % 4 == 0 && fdn_aligned16(img) && fdn_aligned16(omul) && fdn_aligned16(oadd), "W must be a multiple of 4 and pointers 16-byte aligned");
    dim3 grid(fdn_cdiv((long long)H * (W / 4), 256), fdn_cdiv(C, FILM_CG), B);
    FDN_LAUNCH(k_film_maps, grid, dim3(256), 0, st, img, wmul, wadd, omul, oadd, C, H, W);
    return fdn_check_launch("k_film_maps");
}

// ---------------------------------------------------------------------------------------------------
// ConvTranspose2d(k=4, s=2, p=1) + LeakyReLU(0.1): out[2H][2W];  w is [Cin][Cout][4][4]
// out[oy][ox] += in[iy][ix] * w[ky][kx]  with oy = 2*iy - 1 + ky
// ---------------------------------------------------------------------------------------------------
__global__ void k_convt4s2(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                           float* __restrict__ out, int Cin, int Cout, int H, int W, int act, long long total) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // over B*Cout*2H*2W
    if (i >= total) return;
    int Wo = 2 * W, Ho = 2 * H;
    int ox = (int)(i % Wo);
    long long t = i / Wo;
    int oy = (int)(t % Ho);
    t /= Ho;
    int co = (int)(t % Cout);
    int b = (int)(t / Cout);
    // ky must have the parity of oy+1; two candidates each
    int ky0 = (oy + 1) & 1, kx0 = (ox + 1) & 1;
    float acc = bias ? bias[co] : 0.f;
    for (int c = 0; c < Cin; ++c) {
        const float* ip = in + ((size_t)b * Cin + c) * H * W;
        const float* wp = w + ((size_t)c * Cout + co) * 16;
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            int ky = ky0 + 2 * a, iy = (oy + 1 - ky) >> 1;
            if (iy < 0 || iy >= H) continue;
#pragma unroll
            for (int d = 0; d < 2; ++d) {
                int kx = kx0 + 2 * d, ix = (ox + 1 - kx) >> 1;
                if (ix < 0 || ix >= W) continue;
                acc += ip[(size_t)iy * W + ix] * wp[ky * 4 + kx];
            }
        }
    }
    out[i] = fdn_act(acc, act);
}

// Gather form for the MAR up-samplers (Cin <= 48, Cout a multiple of 12): a thread owns one input position and produces the 2x2
// output block it maps to, for 12 output channels, from the 3x3 input neighbourhood (9 coalesced loads per input channel feed
// 192 FMAs); the weights of the channel group are broadcast from shared memory.  Even output rows take kernel rows 1 (same input
// row) and 3 (row above), odd rows take 2 (same row) and 0 (row below); columns alike.
#define CT_CO 12
#define CT_MAXCIN 48
__global__ void __launch_bounds__(256) k_convt4s2_g(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                                                    float* __restrict__ out, int Cin, int Cout, int H, int W, int act, long long total) {
    __shared__ __align__(16) float sw[CT_MAXCIN * CT_CO * 16];          // [cin][co][ky][kx]
    const int co0 = blockIdx.y * CT_CO;
    for (int i = threadIdx.x; i < Cin * CT_CO * 16; i += 256) {
        const int c = i / (CT_CO * 16), r = i - c * (CT_CO * 16);
        sw[i] = w[((size_t)c * Cout + co0) * 16 + r];
    }
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // over B*H*W
    if (i >= total) return;
    const int ix = (int)(i % W);
    long long t = i / W;
    const int iy = (int)(t % H);
    const long long b = t / H;
    float acc[CT_CO][4];
#pragma unroll
    for (int o = 0; o < CT_CO; ++o) {
        const float bv = bias ? bias[co0 + o] : 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[o][j] = bv;
    }
    const bool up = iy > 0, dn = iy + 1 < H, lf = ix > 0, rt = ix + 1 < W;
    for (int c = 0; c < Cin; ++c) {
        const float* p = in + (((size_t)b * Cin + c) * H + iy) * W + ix;
        float n[3][3];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const bool ok = (dy == 1 || (dy == 0 ? up : dn)) && (dx == 1 || (dx == 0 ? lf : rt));
                n[dy][dx] = ok ? p[(dy - 1) * W + (dx - 1)] : 0.f;
            }
        const float4* wc = reinterpret_cast<const float4*>(sw + c * CT_CO * 16);
#pragma unroll
        for (int o = 0; o < CT_CO; ++o) {
            const float4 k0 = wc[o * 4 + 0], k1 = wc[o * 4 + 1], k2 = wc[o * 4 + 2], k3 = wc[o * 4 + 3];     // kernel rows 0..3
            // output (a, b2) = (row parity, column parity); rows: a=0 -> ky 1 (n[1]) + ky 3 (n[0]); a=1 -> ky 2 (n[1]) + ky 0 (n[2])
            acc[o][0] += n[1][1] * k1.y + n[1][0] * k1.w + n[0][1] * k3.y + n[0][0] * k3.w;      // even row, even col: kx 1 (same), kx 3 (left)
            acc[o][1] += n[1][1] * k1.z + n[1][2] * k1.x + n[0][1] * k3.z + n[0][2] * k3.x;      // even row, odd col: kx 2 (same), kx 0 (right)
            acc[o][2] += n[1][1] * k2.y + n[1][0] * k2.w + n[2][1] * k0.y + n[2][0] * k0.w;      // odd row, even col
            acc[o][3] += n[1][1] * k2.z + n[1][2] * k2.x + n[2][1] * k0.z + n[2][2] * k0.x;      // odd row, odd col
        }
    }
    const int Wo = 2 * W, Ho = 2 * H;
#pragma unroll
    for (int o = 0; o < CT_CO; ++o) {
        float* op = out + (((size_t)b * Cout + co0 + o) * Ho + 2 * iy) * Wo + 2 * ix;
        *reinterpret_cast<float2*>(op) = make_float2(fdn_act(acc[o][0], act), fdn_act(acc[o][1], act));
        *reinterpret_cast<float2*>(op + Wo) = make_float2(fdn_act(acc[o][2], act), fdn_act(acc[o][3], act));
    }
}

FDN_API int fdn_convt4s2(const float* in, const float* w, const float* bias, float* out, int B, int Cin, int Cout, int H, int W,
                         int act, cudaStream_t st) {
    FDN_REQUIRE(in && w && out && B > 0 && Cin > 0 && Cout > 0, "bad arguments");
    if (Cin <= CT_MAXCIN && Cout % CT_CO == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0) {
        long long npos = (long long)B * H * W;
        FDN_LAUNCH(k_convt4s2_g, dim3(fdn_cdiv(npos, 256), Cout / CT_CO), dim3(256), 0, st, in, w, bias, out, Cin, Cout, H, W, act, npos);
        return fdn_check_launch("k_convt4s2_g");
    }
    long long total = (long long)B * Cout * H * W * 4;
    FDN_LAUNCH_SEQ(k_convt4s2, dim3(fdn_cdiv(total, 256)), dim3(256), 0, st, in, w, bias, out, Cin, Cout, H, W, act, total);
    return fdn_check_launch("k_convt4s2");
}

// ---------------------------------------------------------------------------------------------------
// depthwise 3x3, zero padding 1.  Each thread produces 4 consecutive pixels of one output channel.
//   mode 0: out[c] = conv(in[c]; w[c])            mode 1: out[c] = gelu(conv(in[c]; w[c]))
//   mode 2: gate, w is [2C][9]: out[j] = gelu(conv(in[j/2]; w[j])) * conv(in[(C+j)/2]; w[C+j])
// ---------------------------------------------------------------------------------------------------
// Each thread produces a 4 (x) by 4 (y) block of one output channel: six input rows of six values (one aligned float4 plus
// the two neighbours) feed sixteen outputs, i.e. 18 load instructions per 16 outputs.
__device__ __forceinline__ void dw_load6(const float* __restrict__ plane, int H, int W, int y0, int x0, float r[6][6]) {
#pragma unroll
    for (int dy = 0; dy < 6; ++dy) {
        const int yy = y0 + dy - 1;
        if (yy >= 0 && yy < H) {
            const float* p = plane + (size_t)yy * W + x0;
            const float4 m = *reinterpret_cast<const float4*>(p);
            r[dy][0] = x0 > 0 ? p[-1] : 0.f;
            r[dy][1] = m.x; r[dy][2] = m.y; r[dy][3] = m.z; r[dy][4] = m.w;
            r[dy][5] = x0 + 4 < W ? p[4] : 0.f;
        } else {
#pragma unroll
            for (int dx = 0; dx < 6; ++dx) r[dy][dx] = 0.f;
        }
    }
}
__device__ __forceinline__ void dw_apply16(const float r[6][6], const float* __restrict__ w, float o[4][4]) {
    float k[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) k[i] = w[i];
#pragma unroll
    for (int y = 0; y < 4; ++y)
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            float a = 0.f;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) a += k[dy * 3 + dx] * r[y + dy][x + dx];
            o[y][x] = a;
        }
}

// 8 (x) by 4 (y) outputs per thread with three rolling input rows of ten values (two aligned float4 plus the two neighbours):
// 24 load instructions per 32 outputs; used whenever W is a multiple of 8 (every FDformer level)
__device__ __forceinline__ void dw_row10(const float* __restrict__ plane, int H, int W, int yy, int x0, float r[10]) {
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
__device__ __forceinline__ void dw_row8(const float (&r0)[10], const float (&r1)[10], const float (&r2)[10], const float (&k)[9], float (&o)[8]) {
#pragma unroll
    for (int x = 0; x < 8; ++x) {
        float a = k[0] * r0[x];
        a += k[1] * r0[x + 1]; a += k[2] * r0[x + 2];
        a += k[3] * r1[x]; a += k[4] * r1[x + 1]; a += k[5] * r1[x + 2];
        a += k[6] * r2[x]; a += k[7] * r2[x + 1]; a += k[8] * r2[x + 2];
        o[x] = a;
    }
}

// the same row for a PAIR of output channels that read the same input rows: k[j] = (weight of output 0, weight of output 1)
__device__ __forceinline__ void dw_row8_pair(const float (&r0)[10], const float (&r1)[10], const float (&r2)[10], const float2 (&k)[9], float2 (&o)[8]) {
#pragma unroll
    for (int x = 0; x < 8; ++x) {
        float2 a = f2mul_s(k[0], r0[x]);
        a = cfma(k[1], r0[x + 1], a); a = cfma(k[2], r0[x + 2], a);
        a = cfma(k[3], r1[x], a); a = cfma(k[4], r1[x + 1], a); a = cfma(k[5], r1[x + 2], a);
        a = cfma(k[6], r2[x], a); a = cfma(k[7], r2[x + 1], a); a = cfma(k[8], r2[x + 2], a);
        o[x] = a;
    }
}

// (Packed fp32x2 rows pairing along x - FFMA2 / FMUL2, two outputs per issue slot, bit-identical - were measured neutral: the register moves that
// build the odd-aligned operand pairs eat the saved issue slots; 62.6 vs 60.9 ms per 8-image step.  Kept out.)
#define DW_ROW8 dw_row8

// every output row is finished (activation / gate) and stored as soon as its third input row has arrived, so only the rolling
// input rows are live: ~64 registers for the plain and GELU modes, ~100 for the gate (two input channels in flight)
template <int MODE, int MINB, int ROWS>
__global__ void __launch_bounds__(128, MINB) k_dwconv3_w8(const float* __restrict__ in, const float* __restrict__ w, float* __restrict__ out,
                                                    int C, int H, int W, int per_plane) {
    // grid.y = plane (b*C + c): the channel - and with it the 9 or 18 weights - is uniform over the CTA, so the weights live in
    // uniform registers / constant operands instead of 18 vector registers per thread
    const int i = blockIdx.x * blockDim.x + threadIdx.x;                // over ceil(H/ROWS)*(W/8)
    if (i >= per_plane) return;
    const int W8 = W >> 3;
    const int x0 = (i % W8) * 8;
    const int y0 = (i / W8) * ROWS;
    const int c = blockIdx.y % C;
    const long long b = blockIdx.y / C;
    const int ca = MODE == 2 ? (c >> 1) : c, cb = (C + c) >> 1;
    const float* pa = in + ((size_t)b * C + ca) * H * W;
    const float* pb = in + ((size_t)b * C + cb) * H * W;
    float ka[9], kb[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        ka[j] = w[c * 9 + j];
        kb[j] = MODE == 2 ? w[(C + c) * 9 + j] : 0.f;
    }
    float a0[10], a1[10], a2[10], b0[10], b1[10], b2[10];
    dw_row10(pa, H, W, y0 - 1, x0, a0);
    dw_row10(pa, H, W, y0, x0, a1);
    if (MODE == 2) {
        dw_row10(pb, H, W, y0 - 1, x0, b0);
        dw_row10(pb, H, W, y0, x0, b1);
    }
    float* op = out + (((size_t)b * C + c) * H + y0) * W + x0;
#pragma unroll
    for (int y = 0; y < ROWS; ++y) {
        float o[8];
        dw_row10(pa, H, W, y0 + y + 1, x0, a2);
        DW_ROW8(a0, a1, a2, ka, o);
        if (MODE >= 1) {
#pragma unroll
            for (int x = 0; x < 8; x += 2) {                        // two pixels per packed operation
                const float2 t = fdn_gelu2(make_float2(o[x], o[x + 1]));
                o[x] = t.x; o[x + 1] = t.y;
            }
        }
        if (MODE == 2) {
            float o2[8];
            dw_row10(pb, H, W, y0 + y + 1, x0, b2);
            DW_ROW8(b0, b1, b2, kb, o2);
#pragma unroll
            for (int x = 0; x < 8; ++x) o[x] *= o2[x];
#pragma unroll
            for (int j = 0; j < 10; ++j) { b0[j] = b1[j]; b1[j] = b2[j]; }
        }
#pragma unroll
        for (int j = 0; j < 10; ++j) { a0[j] = a1[j]; a1[j] = a2[j]; }
        if (y0 + y < H) {
            *reinterpret_cast<float4*>(op + (size_t)y * W) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4*>(op + (size_t)y * W + 4) = make_float4(o[4], o[5], o[6], o[7]);
        }
    }
}

// Gate mode for a PAIR of output channels (2m, 2m+1): both read input channel m for the GELU branch and input channel (C + 2m) / 2
// (and (C + 2m + 1) / 2, a different one only when C is odd) for the linear branch, so the rolling input rows are loaded once and
// feed four convolutions - 12 load/store instructions per 8-pixel row pair instead of 20 with one output channel per thread (these
// kernels are bound by the LSU issue rate, not by DRAM).  ODD: C is odd, a third set of rows serves output 2m+1.
template <bool ODD, int MINB>
__global__ void __launch_bounds__(128, MINB) k_dwgate_pair(const float* __restrict__ in, const float* __restrict__ w, float* __restrict__ out,
                                                           int C, int H, int W, int per_plane) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;                // over ceil(H/4)*(W/8)
    if (i >= per_plane) return;
    const int W8 = W >> 3;
    const int x0 = (i % W8) * 8;
    const int y0 = (i / W8) * 4;
    const int npair = (C + 1) >> 1;
    const int m = blockIdx.y % npair;
    const long long b = blockIdx.y / npair;
    const int c0 = 2 * m, c1 = 2 * m + 1;
    const bool has1 = c1 < C;
    const int cb0 = (C + c0) >> 1, cb1 = has1 ? (C + c1) >> 1 : cb0;
    const float* pa = in + ((size_t)b * C + m) * H * W;
    const float* pb = in + ((size_t)b * C + cb0) * H * W;
    const float* pd = in + ((size_t)b * C + cb1) * H * W;
    // The two outputs of the pair share their input rows, so every tap is ONE packed FFMA2: (acc_2m, acc_2m+1) += (k_2m, k_2m+1) * in,
    // the input a broadcast scalar operand and the weight / accumulator pairs in aligned register pairs - half the FMA issue slots of
    // this issue-bound kernel with no operand shuffling (pairing along x instead needs odd-aligned pairs and was measured neutral).
    // Same operation order as dw_row8 / fdn_gelu: bit-identical results.
    float2 ka[9], kb[9];                      // (output 2m, output 2m+1) weights of the GELU branch / the linear branch
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        ka[j] = make_float2(w[c0 * 9 + j], has1 ? w[c1 * 9 + j] : 0.f);
        kb[j] = make_float2(w[(C + c0) * 9 + j], has1 ? w[(C + c1) * 9 + j] : 0.f);
    }
    float a0[10], a1[10], a2[10], b0[10], b1[10], b2[10], d0[ODD ? 10 : 1], d1[ODD ? 10 : 1], d2[ODD ? 10 : 1];
    dw_row10(pa, H, W, y0 - 1, x0, a0);
    dw_row10(pa, H, W, y0, x0, a1);
    dw_row10(pb, H, W, y0 - 1, x0, b0);
    dw_row10(pb, H, W, y0, x0, b1);
    if (ODD) {
        dw_row10(pd, H, W, y0 - 1, x0, reinterpret_cast<float(&)[10]>(d0));
        dw_row10(pd, H, W, y0, x0, reinterpret_cast<float(&)[10]>(d1));
    }
    float* op0 = out + (((size_t)b * C + c0) * H + y0) * W + x0;
    float* op1 = out + (((size_t)b * C + c1) * H + y0) * W + x0;
#pragma unroll
    for (int y = 0; y < 4; ++y) {
        float2 g[8], l[8];
        dw_row10(pa, H, W, y0 + y + 1, x0, a2);
        dw_row10(pb, H, W, y0 + y + 1, x0, b2);
        dw_row8_pair(a0, a1, a2, ka, g);
        if (ODD) {
            dw_row10(pd, H, W, y0 + y + 1, x0, reinterpret_cast<float(&)[10]>(d2));
            float kx[9], ky[9], lx[8], ly[8];
#pragma unroll
            for (int j = 0; j < 9; ++j) { kx[j] = kb[j].x; ky[j] = kb[j].y; }
            dw_row8(b0, b1, b2, kx, lx);
            dw_row8(reinterpret_cast<float(&)[10]>(d0), reinterpret_cast<float(&)[10]>(d1), reinterpret_cast<float(&)[10]>(d2), ky, ly);
#pragma unroll
            for (int x = 0; x < 8; ++x) l[x] = make_float2(lx[x], ly[x]);
#pragma unroll
            for (int j = 0; j < 10; ++j) { d0[ODD ? j : 0] = d1[ODD ? j : 0]; d1[ODD ? j : 0] = d2[ODD ? j : 0]; }
        } else {
            dw_row8_pair(b0, b1, b2, kb, l);
        }
#pragma unroll
        for (int x = 0; x < 8; ++x) g[x] = fdn_gelu_gate2(g[x], l[x]);
#pragma unroll
        for (int j = 0; j < 10; ++j) { a0[j] = a1[j]; a1[j] = a2[j]; b0[j] = b1[j]; b1[j] = b2[j]; }
        if (y0 + y < H) {
            *reinterpret_cast<float4*>(op0 + (size_t)y * W) = make_float4(g[0].x, g[1].x, g[2].x, g[3].x);
            *reinterpret_cast<float4*>(op0 + (size_t)y * W + 4) = make_float4(g[4].x, g[5].x, g[6].x, g[7].x);
            if (has1) {
                *reinterpret_cast<float4*>(op1 + (size_t)y * W) = make_float4(g[0].y, g[1].y, g[2].y, g[3].y);
                *reinterpret_cast<float4*>(op1 + (size_t)y * W + 4) = make_float4(g[4].y, g[5].y, g[6].y, g[7].y);
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_dwconv3(const float* __restrict__ in, const float* __restrict__ w, float* __restrict__ out,
                                                 int C, int H, int W, int mode, long long total) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // over B*C*ceil(H/4)*(W/4)
    if (i >= total) return;
    const int W4 = W >> 2, H4 = (H + 3) >> 2;
    const int x0 = (int)(i % W4) * 4;
    long long t = i / W4;
    const int y0 = (int)(t % H4) * 4;
    t /= H4;
    const int c = (int)(t % C);
    const long long b = t / C;
    float r[6][6], o[4][4];
    if (mode != 2) {
        dw_load6(in + ((size_t)b * C + c) * H * W, H, W, y0, x0, r);
        dw_apply16(r, w + c * 9, o);
        if (mode == 1) {
#pragma unroll
            for (int y = 0; y < 4; ++y)
#pragma unroll
                for (int x = 0; x < 4; ++x) o[y][x] = fdn_gelu(o[y][x]);
        }
    } else {
        float o2[4][4];
        const int ca = c >> 1, cb = (C + c) >> 1;
        dw_load6(in + ((size_t)b * C + ca) * H * W, H, W, y0, x0, r);
        dw_apply16(r, w + c * 9, o);
        if (cb != ca) dw_load6(in + ((size_t)b * C + cb) * H * W, H, W, y0, x0, r);
        dw_apply16(r, w + (C + c) * 9, o2);
#pragma unroll
        for (int y = 0; y < 4; ++y)
#pragma unroll
            for (int x = 0; x < 4; ++x) o[y][x] = fdn_gelu(o[y][x]) * o2[y][x];
    }
    float* op = out + (((size_t)b * C + c) * H + y0) * W + x0;
#pragma unroll
    for (int y = 0; y < 4; ++y)
        if (y0 + y < H) *reinterpret_cast<float4*>(op + (size_t)y * W) = make_float4(o[y][0], o[y][1], o[y][2], o[y][3]);
}

FDN_API int fdn_dwconv3(const float* in, const float* w, float* out, int B, int C, int H, int W, int mode, cudaStream_t st) {
    FDN_REQUIRE(in && w && out && B > 0 && C > 0 && H > 0 && W > 0, "bad arguments");
    FDN_REQUIRE(W % 4 == 0 && fdn_aligned16(out), "W must be a multiple of 4 and out 16-byte aligned");
    FDN_REQUIRE(mode >= 0 && mode <= 2, "bad mode");
    FDN_REQUIRE(fdn_aligned16(in), "in must be 16-byte aligned");
    if (W % 8 == 0 && !getenv("FDN_DWCONV_W4")) {
        FDN_REQUIRE((long long)B * C <= 65535, "too many planes for one launch");
        // rows per thread: 4.  FDN_DW_ROWS=8 (ten input rows feed eight output rows, 17 % fewer load instructions) measured 3 % slower
        // on the 8-image step (63.8 vs 62.1 ms): the halved thread count costs more latency hiding than the loads save
        static const int rows_env = getenv("FDN_DW_ROWS") ? atoi(getenv("FDN_DW_ROWS")) : 4;
        const int rows = (rows_env == 8 && H % 8 == 0 && (long long)B * C * (H / 8) * (W / 8) >= 148LL * 2048) ? 8 : 4;
        const int total8 = ((H + rows - 1) / rows) * (W / 8);
        dim3 grid(fdn_cdiv(total8, 128), B * C), block(128);
        static const int gate_occ = getenv("FDN_DW_GATE_OCC") ? atoi(getenv("FDN_DW_GATE_OCC")) : 5;    // 5 CTAs/SM (96 regs): 8.9 vs 10.2 ms at 4
#define FDN_DW_LAUNCH(M, OCC)                                                                                          \
    {                                                                                                                  \
        if (rows == 8) { auto k = k_dwconv3_w8<M, OCC, 8>; FDN_LAUNCH_SEQ(k, grid, block, 0, st, in, w, out, C, H, W, total8); } \
        else { auto k = k_dwconv3_w8<M, OCC, 4>; FDN_LAUNCH_SEQ(k, grid, block, 0, st, in, w, out, C, H, W, total8); }           \
    }
        static const int pair_env = getenv("FDN_DW_GATE_PAIR") ? atoi(getenv("FDN_DW_GATE_PAIR")) : 1;
        if (mode == 2 && pair_env) {          // two output channels per thread: the shared input rows are loaded once
            const int total4 = ((H + 3) / 4) * (W / 8);
            dim3 pgrid(fdn_cdiv(total4, 128), B * ((C + 1) / 2));
            if (C & 1) { auto k = k_dwgate_pair<true, 3>; FDN_LAUNCH_SEQ(k, pgrid, block, 0, st, in, w, out, C, H, W, total4); }
            else { auto k = k_dwgate_pair<false, 4>; FDN_LAUNCH_SEQ(k, pgrid, block, 0, st, in, w, out, C, H, W, total4); }
            return fdn_check_launch("k_dwgate_pair");
        }
        if (mode == 0) FDN_DW_LAUNCH(0, 8)
        else if (mode == 1) FDN_DW_LAUNCH(1, 8)
        else if (gate_occ == 5) FDN_DW_LAUNCH(2, 5)
        else if (gate_occ == 6) FDN_DW_LAUNCH(2, 6)
        else FDN_DW_LAUNCH(2, 4)
#undef FDN_DW_LAUNCH
        return fdn_check_launch("k_dwconv3_w8");
    }
    long long total = (long long)B * C * ((H + 3) / 4) * (W / 4);
    FDN_LAUNCH_SEQ(k_dwconv3, dim3(fdn_cdiv(total, 256)), dim3(256), 0, st, in, w, out, C, H, W, mode, total);
    return fdn_check_launch("k_dwconv3");
}
