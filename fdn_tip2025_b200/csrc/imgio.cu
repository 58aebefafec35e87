// Image pre/post-processing of the inference scripts, on the device:
//   pre : cv2.imread bytes (uint8, HWC, BGR) -> /255 in fp32 -> RGB, CHW (img2tensor, basicsr/utils/img_util.py:9-33) ->
//         F.pad(..., (0, w_n, 0, h_n), mode='reflect') to the padded size (inference_fdn_lolblur.py:47-62)
//   post: crop [:h, :w] -> clamp to [0, 1] -> *255 -> round half to even -> uint8, HWC, BGR
//         (inference_fdn_lolblur.py:72-73, tensor2img img_util.py:36-98)
// Moving uint8 instead of fp32 across PCIe cuts the host<->device bytes of a frame by four.
#include "fdn_common.cuh"

// one thread per padded pixel; the three channel planes are written with coalesced stores
__global__ void __launch_bounds__(256) k_pre_u8hwc(const unsigned char* __restrict__ img, float* __restrict__ out, int h, int w, int Hp,
                                                  int Wp, long long total) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // over B*Hp*Wp
    if (i >= total) return;
    const int x = (int)(i % Wp);
    long long t = i / Wp;
    const int y = (int)(t % Hp);
    const long long b = t / Hp;
    const int sy = y < h ? y : 2 * (h - 1) - y;                          // reflect (no edge repeat), pad < size
    const int sx = x < w ? x : 2 * (w - 1) - x;
    const unsigned char* p = img + (((size_t)b * h + sy) * w + sx) * 3;  // B, G, R
    const size_t plane = (size_t)Hp * Wp;
    float* o = out + (size_t)b * 3 * plane + (size_t)y * Wp + x;
    o[0] = (float)p[2] / 255.0f;
    o[plane] = (float)p[1] / 255.0f;
    o[2 * plane] = (float)p[0] / 255.0f;
}

__global__ void __launch_bounds__(256) k_post_u8hwc(const float* __restrict__ x, unsigned char* __restrict__ img, int h, int w, int Hp,
                                                   int Wp, long long total) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // over B*h*w
    if (i >= total) return;
    const int xx = (int)(i % w);
    long long t = i / w;
    const int y = (int)(t % h);
    const long long b = t / h;
    const size_t plane = (size_t)Hp * Wp;
    const float* p = x + (size_t)b * 3 * plane + (size_t)y * Wp + xx;
    unsigned char* o = img + (size_t)i * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float v = p[(size_t)(2 - c) * plane];                            // output channel c of BGR = input channel 2-c of RGB
        v = fminf(fmaxf(v, 0.f), 1.f);
        o[c] = (unsigned char)rintf(v * 255.0f);                         // numpy .round(): half to even
    }
}

// img [B][h][w][3] uint8 BGR -> out [B][3][Hp][Wp] fp32 RGB in [0,1], reflect-padded on the right / bottom (Hp >= h, Wp >= w,
// Hp - h < h, Wp - w < w as torch's reflect padding requires)
FDN_API int fdn_pre_u8hwc_to_f32chw(const unsigned char* img, float* out, int B, int h, int w, int Hp, int Wp, cudaStream_t st) {
    FDN_REQUIRE(img && out && B > 0 && h > 0 && w > 0, "bad arguments");
    FDN_REQUIRE(Hp >= h && Wp >= w && Hp - h < h && Wp - w < w, "reflect padding must be smaller than the image");
    long long total = (long long)B * Hp * Wp;
    FDN_LAUNCH_SEQ(k_pre_u8hwc, dim3(fdn_cdiv(total, 256)), dim3(256), 0, st, img, out, h, w, Hp, Wp, total);
    return fdn_check_launch("k_pre_u8hwc");
}

// x [B][3][Hp][Wp] fp32 RGB -> img [B][h][w][3] uint8 BGR: crop, clamp, *255, round half to even
FDN_API int fdn_post_f32chw_to_u8hwc(const float* x, unsigned char* img, int B, int h, int w, int Hp, int Wp, cudaStream_t st) {
    FDN_REQUIRE(x && img && B > 0 && h > 0 && w > 0 && Hp >= h && Wp >= w, "bad arguments");
    long long total = (long long)B * h * w;
    FDN_LAUNCH_SEQ(k_post_u8hwc, dim3(fdn_cdiv(total, 256)), dim3(256), 0, st, x, img, h, w, Hp, Wp, total);
    return fdn_check_launch("k_post_u8hwc");
}
