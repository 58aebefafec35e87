// LPNet (I_predict_net) specific kernels; the convolutions themselves run through fdn_conv2d / fdn_pw_conv with
// BatchNorm folded into weight and bias at pack time.   LPNet_arch.py:70-81 (SEBlock), 114-134 (forward)
#include "fdn_common.cuh"

// AvgPool2d(kernel 3, stride 2, padding 1), count_include_pad=True (divide by 9 always)
__global__ void k_avgpool3s2(const float* __restrict__ in, float* __restrict__ out, int H, int W, int Ho, int Wo, long long total) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int x = (int)(i % Wo);
    long long t = i / Wo;
    int y = (int)(t % Ho);
    long long pl = t / Ho;
    const float* p = in + (size_t)pl * H * W;
    float s = 0.f;
    for (int dy = -1; dy <= 1; ++dy) {
        int yy = 2 * y + dy;
        if (yy < 0 || yy >= H) continue;
        for (int dx = -1; dx <= 1; ++dx) {
            int xx = 2 * x + dx;
            if (xx < 0 || xx >= W) continue;
            s += p[(size_t)yy * W + xx];
        }
    }
    out[i] = s * (1.0f / 9.0f);
}

// per-plane mean: out[plane] = mean(in[plane][:])     one CTA per plane
__global__ void __launch_bounds__(256) k_plane_mean(const float* __restrict__ in, float* __restrict__ out, int HW) {
    __shared__ float red[256];
    const float* p = in + (size_t)blockIdx.x * HW;
    float s = 0.f;
    for (int i = threadIdx.x; i < HW; i += 256) s += p[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = red[0] / (float)HW;
}

// squeeze-excite gate: s = sigmoid(W2 relu(W1 m + b1) + b2), m [B][C], W1 [C/16][C], W2 [C][C/16].  One CTA per image.
__global__ void __launch_bounds__(128) k_se_fc(const float* __restrict__ m, const float* __restrict__ w1, const float* __restrict__ b1,
                                               const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ s, int C, int R) {
    __shared__ float hid[64];
    const float* mb = m + (size_t)blockIdx.x * C;
    if ((int)threadIdx.x < R) {
        float a = b1[threadIdx.x];
        for (int c = 0; c < C; ++c) a += w1[threadIdx.x * C + c] * mb[c];
        hid[threadIdx.x] = fmaxf(a, 0.f);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a = b2[c];
        for (int r = 0; r < R; ++r) a += w2[c * R + r] * hid[r];
        s[(size_t)blockIdx.x * C + c] = fdn_sigmoid(a);
    }
}

// out = relu(x * s[plane] + shortcut)
__global__ void k_se_apply(const float* __restrict__ x, const float* __restrict__ s, const float* __restrict__ sc, float* __restrict__ out,
                           int HW, long long total) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    out[i] = fmaxf(x[i] * s[i / HW] + sc[i], 0.f);
}

// head: y = sigmoid(fc2(fc(m))), m [B][C]; optionally gray/y.   One CTA per image, C <= 256.
__global__ void __launch_bounds__(256) k_lpnet_head(const float* __restrict__ m, const float* __restrict__ w1, const float* __restrict__ b1,
                                                    const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ gray,
                                                    float* __restrict__ out, int C) {
    __shared__ float h[256];
    const float* mb = m + (size_t)blockIdx.x * C;
    if ((int)threadIdx.x < C) {
        float a = b1[threadIdx.x];
        for (int c = 0; c < C; ++c) a += w1[threadIdx.x * C + c] * mb[c];
        h[threadIdx.x] = a;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = b2[0];
        for (int c = 0; c < C; ++c) a += w2[c] * h[c];
        float y = fdn_sigmoid(a);
        out[blockIdx.x] = gray ? gray[blockIdx.x] / y : y;
    }
}

// gray mean per image: mean(0.2989 R + 0.587 G + 0.114 B)   (torchvision Grayscale, LPNet_arch.py:115-117)
__global__ void __launch_bounds__(256) k_gray_mean(const float* __restrict__ x, float* __restrict__ out, int HW) {
    __shared__ float red[256];
    const float* p = x + (size_t)blockIdx.x * 3 * HW;
    float s = 0.f;
    for (int i = threadIdx.x; i < HW; i += 256) s += 0.2989f * p[i] + 0.587f * p[HW + i] + 0.114f * p[2 * HW + i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = red[0] / (float)HW;
}

FDN_API int fdn_avgpool3s2(const float* in, float* out, int planes, int H, int W, cudaStream_t st) {
    FDN_REQUIRE(in && out && planes > 0 && H > 0 && W > 0, "bad arguments");
    int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    long long total = (long long)planes * Ho * Wo;
    FDN_LAUNCH_SEQ(k_avgpool3s2, dim3(fdn_cdiv(total, 256)), dim3(256), 0, st, in, out, H, W, Ho, Wo, total);
    return fdn_check_launch("k_avgpool3s2");
}
FDN_API int fdn_plane_mean(const float* in, float* out, int planes, int HW, cudaStream_t st) {
    FDN_REQUIRE(in && out && planes > 0 && HW > 0, "bad arguments");
    FDN_LAUNCH(k_plane_mean, dim3(planes), dim3(256), 0, st, in, out, HW);
    return fdn_check_launch("k_plane_mean");
}
FDN_API int fdn_se_fc(const float* m, const float* w1, const float* b1, const float* w2, const float* b2, float* s, int B, int C, int R,
                      cudaStream_t st) {
    FDN_REQUIRE(m && w1 && b1 && w2 && b2 && s && B > 0 && C > 0 && R > 0 && R <= 64, "bad arguments");
    FDN_LAUNCH(k_se_fc, dim3(B), dim3(128), 0, st, m, w1, b1, w2, b2, s, C, R);
    return fdn_check_launch("k_se_fc");
}
FDN_API int fdn_se_apply(const float* x, const float* s, const float* shortcut, float* out, int planes, int HW, cudaStream_t st) {
    FDN_REQUIRE(x && s && shortcut && out && planes > 0 && HW > 0, "bad arguments");
    long long total = (long long)planes * HW;
    FDN_LAUNCH_SEQ(k_se_apply, dim3(fdn_cdiv(total, 256)), dim3(256), 0, st, x, s, shortcut, out, HW, total);
    return fdn_check_launch("k_se_apply");
}
FDN_API int fdn_lpnet_head(const float* m, const float* w1, const float* b1, const float* w2, const float* b2, const float* gray,
                           float* out, int B, int C, cudaStream_t st) {
    FDN_REQUIRE(m && w1 && b1 && w2 && b2 && out && B > 0 && C > 0 && C <= 256, "bad arguments");
    FDN_LAUNCH(k_lpnet_head, dim3(B), dim3(256), 0, st, m, w1, b1, w2, b2, gray, out, C);
    return fdn_check_launch("k_lpnet_head");
}
FDN_API int fdn_gray_mean(const float* x, float* out, int B, int HW, cudaStream_t st) {
    FDN_REQUIRE(x && out && B > 0 && HW > 0, "bad arguments");
    FDN_LAUNCH(k_gray_mean, dim3(B), dim3(256), 0, st, x, out, HW);
    return fdn_check_launch("k_gray_mean");
}
