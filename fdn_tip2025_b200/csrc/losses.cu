// Training-side spectral losses, forward only, on top of the global-FFT kernels:
//   FFTLoss  basicsr/models/losses/losses.py:83-115   loss_weight * mean |rfft2(pred) - rfft2(target)| over real and imaginary parts
//   MARLoss  basicsr/models/losses/losses.py:764-774  mse(x, y_d) + 10 * vgg(x, y_d) + 0.01 * mse(|rfft2(x)|, |rfft2(y_d)|),
//            y_d = nn.Upsample(scale_factor=1/8, 'bilinear', align_corners=False)(y)  (the VGG term stays with the caller)
// This file holds the small kernels around the transforms: difference, reductions in float64, the 1/8 bilinear resample.
#include "fdn_common.cuh"

__global__ void __launch_bounds__(256) k_diff(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] - b[i];
}

// mode 0: sum |a[i]|      mode 1: sum (a[i] - b[i])^2      -> atomicAdd into *out (float64)
__global__ void __launch_bounds__(256) k_reduce(const float* __restrict__ a, const float* __restrict__ b, double* __restrict__ out, long long n, int mode) {
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (mode == 0) {
            s += (double)fabsf(a[i]);
        } else {
            const double d = (double)a[i] - (double)b[i];
            s += d * d;
        }
    }
    __shared__ double red[256];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(out, red[0]);
}

// bilinear scale 1/8, align_corners=False: output (i, j) samples the input at (8 i + 3.5, 8 j + 3.5), i.e. the mean of the four
// centre pixels of its 8x8 block
__global__ void __launch_bounds__(256) k_down8(const float* __restrict__ in, float* __restrict__ out, int H, int W, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // over planes * (H/8) * (W/8)
    if (i >= total) return;
    const int wo = W >> 3, ho = H >> 3;
    const int x = (int)(i % wo);
    const long long t = i / wo;
    const int y = (int)(t % ho);
    const long long plane = t / ho;
    const float* p = in + (size_t)plane * H * W + (size_t)(8 * y + 3) * W + 8 * x + 3;
    // PyTorch interpolates along x inside each row, then along y: w0 * (w0 * a + w1 * b) + w1 * (w0 * c + w1 * d) with w0 = w1 = 0.5
    out[i] = 0.5f * (0.5f * p[0] + 0.5f * p[1]) + 0.5f * (0.5f * p[W] + 0.5f * p[W + 1]);
}

// out[i] = a[i] - b[i]
FDN_API int fdn_diff(const float* a, const float* b, float* out, long long n, cudaStream_t st) {
    FDN_REQUIRE(a && b && out && n > 0, "bad arguments");
    FDN_LAUNCH_SEQ(k_diff, dim3(fdn_cdiv(n, 256)), dim3(256), 0, st, a, b, out, n);
    return fdn_check_launch("k_diff");
}

// *out (float64, zeroed here) = sum |a[i]| (mode 0) or sum (a[i] - b[i])^2 (mode 1)
FDN_API int fdn_reduce_f64(const float* a, const float* b, double* out, long long n, int mode, cudaStream_t st) {
    FDN_REQUIRE(a && out && n > 0 && (mode == 0 || (mode == 1 && b)), "bad arguments");
    cudaMemsetAsync(out, 0, sizeof(double), st);
    const long long blocks = (n + 255) / 256;
    FDN_LAUNCH(k_reduce, dim3((unsigned)(blocks < 1184 ? blocks : 1184)), dim3(256), 0, st, a, b, out, n, mode);
    return fdn_check_launch("k_reduce");
}

// nn.Upsample(scale_factor=1/8, mode='bilinear', align_corners=False): in [planes][H][W] -> out [planes][H/8][W/8]
FDN_API int fdn_down8_bilinear(const float* in, float* out, int planes, int H, int W, cudaStream_t st) {
    FDN_REQUIRE(in && out && planes > 0 && H >= 8 && W >= 8, "bad arguments");
    FDN_REQUIRE(H % 8 == 0 && W % 8 == 0, "H and W must be multiples of 8");
    const long long total = (long long)planes * (H / 8) * (W / 8);
    FDN_LAUNCH_SEQ(k_down8, dim3(fdn_cdiv(total, 256)), dim3(256), 0, st, in, out, H, W, total);
    return fdn_check_launch("k_down8");
}
