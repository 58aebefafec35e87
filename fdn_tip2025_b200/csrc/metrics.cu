// Validation metrics on the device: PSNR and SSIM of restored frames against ground truth, so that
// validation never leaves the GPU.  Semantics follow basicsr/metrics/psnr_ssim.py of the reference:
//   calculate_psnr   :8-70     mse over all (cropped) elements in float64, max_value = 1 if img1.max() <= 1 else 255
//   _ssim            :84-117   per-channel 11x11 Gaussian (sigma 1.5), valid region only ([5:-5, 5:-5]), mean over channels
//   _ssim_3d         :163-200  the default (ssim3d=True): 11x11x11 Gaussian over the (H, W, C) volume, replicate padding
//   _ssim_cly        :202-240  Y channel (ITU-R BT.601 via to_y_channel / bgr2ycbcr), replicate border, constants for range 255
// All arithmetic is double precision: the images are tiny next to the forward pass, and variance terms E[x^2] - mu^2 of smooth
// frames cancel badly in fp32 (the reference's own _ssim_3d runs its conv3d in fp32 and moves by ~1e-6 with the cuDNN algorithm).
// Inputs are [B][C][H][W] fp32 (our tensors are planar; the reference's HWC arrays hold the same volume).
#include "fdn_common.cuh"

#define SSIM_R 5
#define SSIM_T 16                    // output tile edge
#define SSIM_HT (SSIM_T + 2 * SSIM_R)
#define SSIM_MAXF 15                 // 5 fields x 3 channels

struct SsimParams {
    const float* a;
    const float* b;
    double* ws;          // [B][4]: sum of the SSIM map, max(img1), sum of squared differences, spare
    int B, C, H, W, crop, mode;
};

__device__ __forceinline__ double warp_sum(double v) {
#ifndef FDN_EMU
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
#endif
    return v;
}

// to_y_channel (metric_util.py:34-47) of a pixel whose channels are (c0, c1, c2) in storage order, range [0, 255]
__device__ __forceinline__ double y_of(double c0, double c1, double c2) {
    const float f0 = (float)c0 / 255.f, f1 = (float)c1 / 255.f, f2 = (float)c2 / 255.f;       // img.astype(np.float32) / 255.
    const double y = ((double)f0 * 24.966 + (double)f1 * 128.553 + (double)f2 * 65.481 + 16.0) / 255.0;
    return (double)((float)y * 255.f);                                                           // .astype(np.float32) * 255.
}

// ws[b][1] = max(img1[b]) over the cropped region, ws[b][2] = sum (img1 - img2)^2 (Y channel if ych)
__global__ void __launch_bounds__(256) k_metric_reduce(SsimParams q, int ych) {
    const int b = blockIdx.y;
    const int h = q.H - 2 * q.crop, w = q.W - 2 * q.crop;
    const long long n = (long long)h * w * (ych ? 1 : q.C);
    const size_t plane = (size_t)q.H * q.W;
    const float* a = q.a + (size_t)b * q.C * plane;
    const float* bb = q.b + (size_t)b * q.C * plane;
    double s = 0.0;
    float mx = -3.4e38f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % w);
        const long long t = i / w;
        const int y = (int)(t % h), c = (int)(t / h);
        const size_t o = (size_t)(y + q.crop) * q.W + x + q.crop;
        if (ych) {
            const double ya = y_of(a[o], a[plane + o], a[2 * plane + o]), yb = y_of(bb[o], bb[plane + o], bb[2 * plane + o]);
            s += (ya - yb) * (ya - yb);
            mx = fmaxf(mx, fmaxf(a[o], fmaxf(a[plane + o], a[2 * plane + o])));
        } else {
            const double d = (double)a[c * plane + o] - (double)bb[c * plane + o];
            s += d * d;
            mx = fmaxf(mx, a[c * plane + o]);
        }
    }
    __shared__ double ss[256];
    __shared__ float sm[256];
    ss[threadIdx.x] = s;
    sm[threadIdx.x] = mx;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            ss[threadIdx.x] += ss[threadIdx.x + o];
            sm[threadIdx.x] = fmaxf(sm[threadIdx.x], sm[threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        atomicAdd(&q.ws[b * 4 + 2], ss[0]);
        // max through the ordered-integer image of the float (works for either sign)
        int bits = __float_as_int(sm[0]);
        bits = bits >= 0 ? bits : bits ^ 0x7fffffff;
        atomicMax(reinterpret_cast<int*>(&q.ws[b * 4 + 1]), bits);
    }
}

__device__ __forceinline__ float ordered_to_float(int bits) { return __int_as_float(bits >= 0 ? bits : bits ^ 0x7fffffff); }

__global__ void k_psnr_final(const double* ws, double* out, int B, long long n, int ych) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double mse = ws[b * 4 + 2] / (double)n;
    const float mx = ordered_to_float(*reinterpret_cast<const int*>(&ws[b * 4 + 1]));
    // img1.max() is taken after to_y_channel when test_y_channel: Y of a [0,1] image lies in [16/255, ~0.92] + ... <= 1 only for
    // tiny inputs; Y is always >= 16/255*... the reference's rule is applied to the values it sees: Y >= 0.0627*255 > 1
    const double max_value = ych ? 255.0 : (mx <= 1.f ? 1.0 : 255.0);
    out[b] = mse == 0.0 ? INFINITY : 20.0 * log10(max_value / sqrt(mse));
}

__constant__ double c_gauss[11];          // cv2.getGaussianKernel(11, 1.5)

// One CTA = one 16x16 tile of the SSIM map of one image.  mode 0: 3-D Gaussian (channels mixed by the replicate-padded 11-tap
// kernel, i.e. a 3x3 matrix), replicate spatial border, all positions.  mode 1: per channel, valid region.  mode 2: Y channel,
// replicate border.  Fields per channel: a, b, a*a, b*b, a*b.
__global__ void __launch_bounds__(256) k_ssim(SsimParams q) {
    FDN_DYN_SMEM(smem);
    double* F = reinterpret_cast<double*>(smem);                       // [nf][SSIM_HT][SSIM_HT]   fields on the halo tile
    const int nch = q.mode == 2 ? 1 : q.C;
    const int nf = 5 * nch;
    double* G = F + (size_t)nf * SSIM_HT * SSIM_HT;                    // [nf][SSIM_HT][SSIM_T]    after the horizontal pass
    const int b = blockIdx.z;
    const int h = q.H - 2 * q.crop, w = q.W - 2 * q.crop;              // cropped image
    const int valid = q.mode == 1 ? SSIM_R : 0;                         // mode 1 evaluates [5:-5, 5:-5] only
    const int oh = h - 2 * valid, ow = w - 2 * valid;                   // SSIM map size
    const int x0 = blockIdx.x * SSIM_T + valid, y0 = blockIdx.y * SSIM_T + valid;
    const size_t plane = (size_t)q.H * q.W;
    const float* a = q.a + (size_t)b * q.C * plane;
    const float* bb = q.b + (size_t)b * q.C * plane;
    const int tid = threadIdx.x;
    // channel mixing matrix of the 3-D kernel: M[c][c'] = sum_d g[d] [clamp(c + d - 5, 0, C-1) == c']
    double M[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    if (q.mode == 0) {
        for (int c = 0; c < 3; ++c)
            for (int c2 = 0; c2 < 3; ++c2) M[c][c2] = 0.0;
        for (int c = 0; c < q.C; ++c)
            for (int d = 0; d < 11; ++d) M[c][min(max(c + d - SSIM_R, 0), q.C - 1)] += c_gauss[d];
    }
    for (int i = tid; i < SSIM_HT * SSIM_HT; i += 256) {
        const int yy = i / SSIM_HT, xx = i - yy * SSIM_HT;
        const int gy = min(max(y0 + yy - SSIM_R, 0), h - 1), gx = min(max(x0 + xx - SSIM_R, 0), w - 1);     // replicate border
        const size_t o = (size_t)(gy + q.crop) * q.W + gx + q.crop;
        double va[3], vb[3];
        if (q.mode == 2) {
            va[0] = y_of(a[o], a[plane + o], a[2 * plane + o]);
            vb[0] = y_of(bb[o], bb[plane + o], bb[2 * plane + o]);
        } else {
            for (int c = 0; c < nch; ++c) { va[c] = a[c * plane + o]; vb[c] = bb[c * plane + o]; }
        }
        double f[5][3];
        for (int c = 0; c < nch; ++c) { f[0][c] = va[c]; f[1][c] = vb[c]; f[2][c] = va[c] * va[c]; f[3][c] = vb[c] * vb[c]; f[4][c] = va[c] * vb[c]; }
        for (int k = 0; k < 5; ++k)
            for (int c = 0; c < nch; ++c) {
                double v = f[k][c];
                if (q.mode == 0) {
                    v = 0.0;
                    for (int c2 = 0; c2 < nch; ++c2) v += M[c][c2] * f[k][c2];
                }
                F[(size_t)(k * nch + c) * SSIM_HT * SSIM_HT + i] = v;
            }
    }
    __syncthreads();
    for (int i = tid; i < nf * SSIM_HT * SSIM_T; i += 256) {            // horizontal pass
        const int fi = i / (SSIM_HT * SSIM_T), r = i - fi * (SSIM_HT * SSIM_T);
        const int yy = r / SSIM_T, xx = r - yy * SSIM_T;
        const double* src = F + (size_t)fi * SSIM_HT * SSIM_HT + yy * SSIM_HT + xx;
        double s = 0.0;
#pragma unroll
        for (int d = 0; d < 11; ++d) s += c_gauss[d] * src[d];
        G[i] = s;
    }
    __syncthreads();
    double local = 0.0;
    {
        const int yy = tid / SSIM_T, xx = tid - yy * SSIM_T;            // 256 threads = 16 x 16 outputs
        if (y0 + yy - valid < oh && x0 + xx - valid < ow) {
            const float mxv = ordered_to_float(*reinterpret_cast<const int*>(&q.ws[b * 4 + 1]));
            const double maxv = q.mode == 2 ? 255.0 : (mxv <= 1.f ? 1.0 : 255.0);
            const double C1 = (0.01 * maxv) * (0.01 * maxv), C2 = (0.03 * maxv) * (0.03 * maxv);
            for (int c = 0; c < nch; ++c) {
                double v[5];
                for (int k = 0; k < 5; ++k) {
                    const double* src = G + (size_t)(k * nch + c) * SSIM_HT * SSIM_T + yy * SSIM_T + xx;
                    double s = 0.0;
#pragma unroll
                    for (int d = 0; d < 11; ++d) s += c_gauss[d] * src[d * SSIM_T];
                    v[k] = s;
                }
                const double mu1 = v[0], mu2 = v[1];
                const double s1 = v[2] - mu1 * mu1, s2 = v[3] - mu2 * mu2, s12 = v[4] - mu1 * mu2;
                local += ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s1 + s2 + C2));
            }
        }
    }
    __shared__ double red[256];
    red[tid] = local;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) red[tid] += red[tid + o];
        __syncthreads();
    }
    if (tid == 0) atomicAdd(&q.ws[b * 4 + 0], red[0]);
}

__global__ void k_ssim_final(const double* ws, double* out, int B, double count) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) out[b] = ws[b * 4] / count;
}

static int upload_gauss() {
    static bool done[FDN_MAX_DEVICES] = {};
    const int dev = fdn_device();
    if (done[dev]) return 0;
    // cv2.getGaussianKernel(11, 1.5): exp(-(i - 5)^2 / (2 sigma^2)), normalised to sum 1
    double g[11], s = 0.0;
    for (int i = 0; i < 11; ++i) { g[i] = exp(-(double)((i - 5) * (i - 5)) / (2.0 * 1.5 * 1.5)); s += g[i]; }
    for (int i = 0; i < 11; ++i) g[i] /= s;
#ifndef FDN_EMU
    if (cudaMemcpyToSymbol(c_gauss, g, sizeof(g)) != cudaSuccess) return -1;
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
#else
    memcpy(c_gauss, g, sizeof(g));
#endif
    done[dev] = true;
    return 0;
}

static void init_ws(double* ws, int B, cudaStream_t st) {
    // [sum, ordered-int max (0x80000000.. = most negative), sqdiff, spare]: zero bits = 0.0 / ordered 0 -> max starts at +0.0;
    // images are non-negative in practice, and a frame that is negative everywhere still selects max_value = 1 like the reference
    cudaMemsetAsync(ws, 0, sizeof(double) * 4 * B, st);
}

// psnr[b] of img1[b] vs img2[b] (calculate_psnr, psnr_ssim.py:8-70).  ws: caller-owned scratch of 4*B doubles.
FDN_API int fdn_psnr(const float* img1, const float* img2, double* psnr, double* ws, int B, int C, int H, int W, int crop_border,
                     int test_y_channel, cudaStream_t st) {
    FDN_REQUIRE(img1 && img2 && psnr && ws && B > 0 && C > 0 && H > 0 && W > 0, "bad arguments");
    FDN_REQUIRE(crop_border >= 0 && 2 * crop_border < H && 2 * crop_border < W, "bad crop_border");
    FDN_REQUIRE(!test_y_channel || C == 3, "the Y channel needs three colour channels");
    SsimParams q{img1, img2, ws, B, C, H, W, crop_border, 0};
    init_ws(ws, B, st);
    const long long n = (long long)(H - 2 * crop_border) * (W - 2 * crop_border) * (test_y_channel ? 1 : C);
    FDN_LAUNCH(k_metric_reduce, dim3((unsigned)min((long long)296, (n + 255) / 256), B), dim3(256), 0, st, q, test_y_channel);
    FDN_LAUNCH_SEQ(k_psnr_final, dim3(fdn_cdiv(B, 64)), dim3(64), 0, st, ws, psnr, B, n, test_y_channel);
    return fdn_check_launch("fdn_psnr");
}

// ssim[b] (calculate_ssim, psnr_ssim.py:243-329).  mode 0: ssim3d=True (default of the reference), 1: ssim3d=False (_ssim),
// 2: test_y_channel=True (_ssim_cly).  ws: caller-owned scratch of 4*B doubles.
FDN_API int fdn_ssim(const float* img1, const float* img2, double* ssim, double* ws, int B, int C, int H, int W, int crop_border,
                     int mode, cudaStream_t st) {
    FDN_REQUIRE(img1 && img2 && ssim && ws && B > 0 && C > 0 && H > 0 && W > 0, "bad arguments");
    FDN_REQUIRE(mode >= 0 && mode <= 2, "bad mode");
    FDN_REQUIRE(C <= 3 && (mode != 2 || C == 3), "at most three channels (Y channel: exactly three)");
    FDN_REQUIRE(crop_border >= 0, "bad crop_border");
    const int h = H - 2 * crop_border, w = W - 2 * crop_border;
    const int oh = mode == 1 ? h - 2 * SSIM_R : h, ow = mode == 1 ? w - 2 * SSIM_R : w;
    FDN_REQUIRE(oh > 0 && ow > 0, "image too small for the 11x11 window");
    FDN_REQUIRE(upload_gauss() == 0, "constant upload failed");
    SsimParams q{img1, img2, ws, B, C, H, W, crop_border, mode};
    init_ws(ws, B, st);
    const long long n = (long long)h * w * C;
    FDN_LAUNCH(k_metric_reduce, dim3((unsigned)min((long long)296, (n + 255) / 256), B), dim3(256), 0, st, q, 0);      // max(img1) -> C1, C2
    const int nch = mode == 2 ? 1 : C;
    const size_t smem = (size_t)5 * nch * (SSIM_HT * SSIM_HT + SSIM_HT * SSIM_T) * sizeof(double);
#ifndef FDN_EMU
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_ssim, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { fdn_set_error(cudaGetErrorString(e)); return (int)e; }
    }
#endif
    FDN_LAUNCH(k_ssim, dim3(fdn_cdiv(ow, SSIM_T), fdn_cdiv(oh, SSIM_T), B), dim3(256), smem, st, q);
    FDN_LAUNCH_SEQ(k_ssim_final, dim3(fdn_cdiv(B, 64)), dim3(64), 0, st, ws, ssim, B, (double)oh * ow * nch);
    return fdn_check_launch("fdn_ssim");
}
