/* libfdn_b200 - C ABI of the B200 (sm_100a) kernels behind the FDN / FDformer / MAR / LPNet inference forward.
 *
 * This is the drop-in boundary of the hot path.  The reference has no FFI of its own: the
 * operators below are what its nn.Module.forward methods call into ATen for.  Each entry point names the reference
 * lines (under basicsr/models/archs/) whose arithmetic it replaces.  INTEGRATION.md shows the Python (ctypes) stub a
 * reference maintainer would add.
 *
 * Conventions
 *  - all tensors are fp32, contiguous NCHW device memory owned by the caller; spectra are interleaved (re,im) pairs laid
 *    out [plane][H][W/2+1]; pointers that are read or written with 128-bit accesses must be 16-byte aligned;
 *  - every call is asynchronous on the caller's stream `st`; the library never synchronises, never allocates per call
 *    (only fdn_fft_prepare / the first use of an FFT length allocates that length's twiddle table);
 *  - return 0 = launched; negative = argument error detected on the host before any launch; positive = cudaError_t.
 *    fdn_last_error_string() describes the last failure on the calling thread.
 */
#ifndef FDN_B200_H
#define FDN_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

const char* fdn_last_error_string(void);
int fdn_abi_version(void);
int fdn_is_device_build(void);

/* ---- whole-plane FFTs: torch.fft.rfft2 / irfft2(norm='backward')
 *      FDN_arch.py:90,98 (FreBlock) 139,147 (fourier_fuse) 411,418 (FCAFFN) 882-914 (FDN prologue) -------------------- */

/* Build and cache the twiddle tables for lengths H and W (call before CUDA-graph capture). */
int fdn_fft_prepare(int H, int W);

/* Row pass R2C: x [planes][H][W] -> spec [planes][H][W/2+1] complex. */
int fdn_fft_rows_r2c(const float* x, float* spec, int planes, int H, int W, cudaStream_t st);

/* Row pass C2R with epilogue: y = img_scale[b] * (irfft_rows(spec) * inv_norm + res_coef * res).
 * res and img_scale may be NULL; b = plane / planes_per_image.  (FreBlock "+ x", ProcessBlock "+ xori", "* ratio":
 * FDN_arch.py:100,118,213-219) */
int fdn_fft_rows_c2r(const float* spec, float* y, int planes, int H, int W, float inv_norm, const float* res, float res_coef,
                     const float* img_scale, int planes_per_image, cudaStream_t st);

/* Column pass over ncols columns of length H.  Element (plane,y,x) of in/out lives at base + plane*ps + y*rs + x
 * (complex elements for spectra, floats for the real maps of modes 3/4).
 *   mode 0 forward; 1 inverse (unscaled); 2 forward, FCAFFN modulation  Y = rd(X) * A * exp(-iP)  with
 *   A = sum_j w_xa[c][j] amp[b][j], P = sum_j w_xp[c][j] pha[b][j]  (FDN_arch.py:411-418), inverse - one kernel;
 *   3 forward then angle(replace_denormals(X)) (FDN_arch.py:882-892); 4 forward then |X| (FDN_arch.py:901-914).
 * W_real (width of the real signal) locates the Nyquist column so the four self-conjugate bins are made exactly real. */
int fdn_fft_cols(const float* in, long long in_ps, int in_rs, float* out, long long out_ps, int out_rs, int planes, int H,
                 int ncols, int W_real, int mode, int C, const float* amp, const float* pha, const float* w_xa,
                 const float* w_xp, cudaStream_t st);

/* MAR per-bin channel MLPs in place on spec [B][NC][nbins] complex (FDN_arch.py:91-97, 140-146):
 * mag' = W2m lrelu(W1m |X| + b1m) + b2m, pha' likewise on angle(X), X' = mag' e^{i pha'}.  NC in {12,24,48}.
 * w packs W1m b1m W2m b2m W1p b1p W2p b2p, matrices row-major [out][in]. */
int fdn_spec_mlp(float* spec, long long plane_stride, long long nbins, int B, int NC, const float* w, cudaStream_t st);

/* ---- 8x8-patch spectral operators ------------------------------------------------------------------------------------ */

/* FDFFN spectral branch (FDN_arch.py:458-470): out = irfft2_8x8(rd(rfft2_8x8(x)) * wspec[c]) + add, wspec [C][8][5] complex
 * = ffta * exp(-i fftp).  add may be NULL. */
int fdn_fdffn_patch(const float* x, const float* add, const float* wspec, float* out, int B, int C, int H, int W, cudaStream_t st);

/* FDFFN spectral branch + the second depthwise conv of the spatial branch (space.2, FDN_arch.py:439-441,457-470):
 * out = irfft2_8x8(rd(rfft2_8x8(h)) * wspec[c]) + depthwise3x3(s1; wb[c]); s1 = gelu(space.0(h)). */
int fdn_fdffn_patch_dw(const float* h, const float* s1, const float* wb, const float* wspec, float* out, int B, int C, int H, int W,
                       cudaStream_t st);

/* FDFFN middle section fused (FDN_arch.py:457-470): out = dw_b(gelu(dw_a(h))) + irfft2_8x8(rd(rfft2_8x8(h)) * wspec);
 * wa, wb [C][9] = space.0 / space.2 depthwise weights.  One thread per (channel, 8x8 patch); the intermediate gelu(dw_a(h)) is
 * produced row by row in registers (zero outside the image, as dw_b's padding sees it) and never written to memory. */
int fdn_fdffn_spatial(const float* h, const float* wa, const float* wb, const float* wspec, float* out, int B, int C, int H, int W,
                      cudaStream_t st);

/* FDSA bin algebra (FDN_arch.py:585-632): hid [B][4E][H][W] = (q,k,v,v_value) after to_hidden_dw; wfft [E][8][5];
 * out [B][3E][H][W] = (out1,out2,out3) before norm1..3. */
int fdn_fdsa_patch(const float* hid, const float* wfft, float* out, int B, int E, int H, int W, cudaStream_t st);

/* Same with to_hidden_dw (depthwise 3x3, FDN_arch.py:563,578) fused in front: hid is the pre-depthwise hidden tensor, wdw [4E][9];
 * vv [B][E][H][W] receives the convolved v_value group for the gate. */
int fdn_fdsa_patch_dw(const float* hid, const float* wdw, const float* wfft, float* out, float* vv, int B, int E, int H, int W,
                      cudaStream_t st);

/* ---- per-pixel operators ---------------------------------------------------------------------------------------------- */

/* 1x1 convolution over the channel concatenation of up to three sources (nn.Conv2d(k=1), torch.cat, F.interpolate nearest:
 * FDN_arch.py:55-60,125,168-169,186-191,230-251,388-389,451-452,562,566,685-692).  shift>0: source is 2^shift smaller
 * (nearest upsample), shift<0: larger (x[::2^-shift]).  wt is the weight transposed to [K][N].  Optional LayerNorm over the
 * K input channels first (FDN_arch.py:326-342).  Epilogue order: +bias, act (1 LeakyReLU 0.1, 2 ReLU), *film_mul+film_add
 * (FDN_arch.py:423), +res_coef*res, *img_scale[b].  Output element (b,n,y,x) at out[b*out_bs + n*out_ps + y*out_rs + x]. */
int fdn_pw_conv(const float* src0, int c0, int shift0, const float* src1, int c1, int shift1, const float* src2, int c2,
                int shift2, const float* wt, const float* bias, const float* ln_w, const float* ln_b, int act,
                const float* film_mul, const float* film_add, const float* res, float res_coef, const float* img_scale,
                float* out, long long out_bs, long long out_ps, int out_rs, int B, int N, int H, int W, cudaStream_t st);

/* Tensor-core (tcgen05, kind::tf32, accumulator in TMEM) 1x1 convolution for the FDformer blocks: M = 128 pixels per CTA,
 * N <= 256 output channels per chunk.  bpack is the weight packed by fdn_tip2025_b200/packing.py into the K-major
 * SWIZZLE_128B shared-memory image (tf32 hi panel + tf32 lo panel per 32-channel block).  prologue (applied per pixel while
 * the A operand is built): 0 none, 1 LayerNorm over the K inputs (FDN_arch.py:671,673), 2 FDSA gate = three LayerNorm groups
 * (statistics = fdn_group_stats output, or NULL to have the kernel compute them - allowed while K/3 <= 40) times v_value=aux (FDN_arch.py:633-639), 3 FCAFFN mix LN(src)*aux + aux (FDN_arch.py:420).  Epilogue: +bias,
 * *film_mul+film_add, +res_coef*res.  passes: 3 = 3xTF32 split (fp32-level accuracy), 1 = single TF32. */
int fdn_has_tcgen05(void);
/* 1 if fdn_pw_mma supports a layer with K inputs (prologue 2: K = 3E), chunks of Nc output columns and this prologue
 * (stats_in_kernel: prologue 2 with stats == NULL); 0 = tile does not fit in shared memory / K too wide: use fdn_pw_conv. */
int fdn_pw_mma_supported(int K, int Nc, int prologue, int stats_in_kernel);
/* Development aid: 8 uint64 device counters receiving the per-role barrier wait cycles of following fdn_pw_mma launches (NULL = off).
 * Counting is compiled in only with -DFDN_MMA_PROFILE=1 (FDN_MMA_PROFILE=1 python -m fdn_tip2025_b200.build); otherwise a no-op. */
int fdn_pw_mma_set_debug(void* counters);
int fdn_pw_mma(const float* src0, int c0, const float* src1, int c1, const float* bpack, int N, int Nc, int nchunks, int prologue,
               const float* ln_w, const float* ln_b, const float* aux, long long aux_bs, const float* stats, const float* bias,
               const float* film_mul, const float* film_add, const float* res, float res_coef, float* out, int B, int HW, int passes,
               cudaStream_t st);
/* Per-pixel LayerNorm statistics of G channel groups of x [B][G*C][HW]: stats [B][G][2][HW] = (mean, 1/sqrt(var + 1e-5)).
 * Feeds prologue 2 of fdn_pw_mma (FDSA norm1..3, FDN_arch.py:633-635). */
int fdn_group_stats(const float* x, float* stats, int B, int G, int C, int HW, cudaStream_t st);

/* Grouped channel LayerNorm (FDN_arch.py:326-342, 420, 633-638):
 * out[b][g*C+c][p] = LN_g(in[b][g*C+c][p]) * mul[b*mul_bs + c*HW + p] + add[b*add_bs + c*HW + p]; mul/add may be NULL. */
int fdn_chan_ln(const float* in, float* out, const float* gamma, const float* beta, const float* mul, long long mul_bs,
                const float* add, long long add_bs, int B, int G, int C, int HW, cudaStream_t st);

/* nn.Upsample(0.5, bilinear) == 2x2 mean (FDN_arch.py:265,719,866) and nn.Upsample(2, bilinear, align_corners=False) (:730) */
int fdn_avgpool2(const float* in, float* out, int planes, int H, int W, cudaStream_t st);
int fdn_up2_bilinear(const float* in, float* out, int planes, int H, int W, cudaStream_t st);
/* nn.PixelUnshuffle(r) (FDN_arch.py:199-200) */
int fdn_pixel_unshuffle(const float* in, float* out, int B, int C, int H, int W, int r, cudaStream_t st);
/* MAR gamma curve out = 1 - (1-x)^(scale*illum) (FDN_arch.py:282-284) */
int fdn_gamma_curve(const float* x, const float* illum, float* out, float scale, long long n, cudaStream_t st);
/* 1-pixel border of t [planes][H][W] <- value[plane % C] (fourier_fuse.fpre[1], padding=1: FDN_arch.py:126) */
int fdn_fill_border(float* t, const float* value, int planes, int C, int H, int W, cudaStream_t st);

/* ---- spatial convolutions --------------------------------------------------------------------------------------------- */

/* Dense KxK conv, groups=1 (FDN_arch.py:26,57,135,192-196,704,720,731,804; LPNet_arch.py:46-61,90).
 * head 0: y = act(conv+bias) + res;  head 1: y = sigmoid(conv+bias+res) + 1e-8 (FDN_arch.py:241,248,255).
 * res is sampled at (y<<res_shift, x<<res_shift) (nearest-downsampled image for the heads). */
int fdn_conv2d(const float* in, const float* w, const float* bias, const float* res, int res_shift, float* out, int B, int Cin,
               int Hin, int Win, int Cout, int K, int stride, int pad, int act, int head, cudaStream_t st);
/* Tensor-core 3x3 convolution, stride 1, padding 1, for the Downsample / Upsample bodies (FDN_arch.py:715-734): implicit GEMM on
 * mma.sync tf32 in 3xTF32 (fp32-level accuracy).  Cin % 8 == 0; Cout a multiple of CN = fdn_conv3x3_mma_cn(Cout) (32 or 24;
 * 0 = unsupported).  wpack: host-packed weights [Cout/CN][Cin/8][hi,lo][tap][8][40] (packing.pack_conv3x3).
 * out = conv(in) + bias + res (bias, res optional). */
int fdn_conv3x3_mma_cn(int Cout);
int fdn_conv3x3_mma(const float* in, const float* wpack, const float* bias, const float* res, float* out, int B, int Cin, int H, int W,
                    int Cout, cudaStream_t st);
/* FCAFFN FiLM maps (FDN_arch.py:423) in one launch: omul / oadd [B][C][H][W] = 3x3 conv (padding 1) of img [B][3][H][W] with the
 * folded kernels wmul / wadd [C][3][3][3] = conv3_x.weight * conv1_x.weight. */
int fdn_film_maps(const float* img, const float* wmul, const float* wadd, float* omul, float* oadd, int B, int C, int H, int W,
                  cudaStream_t st);
/* ConvTranspose2d(k=4,s=2,p=1) + activation (FDN_arch.py:21-23,194-195); w [Cin][Cout][4][4] */
int fdn_convt4s2(const float* in, const float* w, const float* bias, float* out, int B, int Cin, int Cout, int H, int W, int act,
                 cudaStream_t st);
/* Depthwise 3x3, padding 1.  mode 0 plain, 1 +GELU (FDN_arch.py:435-441), 2 C->2C with gelu(x1)*x2 gate, w [2C][9]
 * (FDN_arch.py:401-402,426-427,448,472-473). */
int fdn_dwconv3(const float* in, const float* w, float* out, int B, int C, int H, int W, int mode, cudaStream_t st);

/* ---- LPNet (LPNet_arch.py:70-81, 114-134) ------------------------------------------------------------------------------ */
int fdn_avgpool3s2(const float* in, float* out, int planes, int H, int W, cudaStream_t st);
int fdn_plane_mean(const float* in, float* out, int planes, int HW, cudaStream_t st);
int fdn_se_fc(const float* m, const float* w1, const float* b1, const float* w2, const float* b2, float* s, int B, int C, int R,
              cudaStream_t st);
int fdn_se_apply(const float* x, const float* s, const float* shortcut, float* out, int planes, int HW, cudaStream_t st);
int fdn_lpnet_head(const float* m, const float* w1, const float* b1, const float* w2, const float* b2, const float* gray,
                   float* out, int B, int C, cudaStream_t st);
int fdn_gray_mean(const float* x, float* out, int B, int HW, cudaStream_t st);

/* ---- image pre/post-processing of the inference scripts (the steps either side of the forward) -----------------------------
 * fdn_pre_u8hwc_to_f32chw: cv2.imread bytes [B][h][w][3] uint8 BGR -> [B][3][Hp][Wp] fp32 RGB in [0,1] (astype(float32)/255,
 *   img2tensor(bgr2rgb=True) basicsr/utils/img_util.py:9-33) with F.pad(..., (0, Wp-w, 0, Hp-h), 'reflect')
 *   (inference_fdn_lolblur.py:47-62, inference_fdn_lolv1.py:44-57).  Requires Hp-h < h and Wp-w < w like torch.
 * fdn_post_f32chw_to_u8hwc: [B][3][Hp][Wp] fp32 RGB -> [B][h][w][3] uint8 BGR: crop [:h,:w], clamp [0,1], *255, round half to
 *   even (tensor2img(rgb2bgr=True) img_util.py:36-98; inference_fdn_lolblur.py:72-73). */
int fdn_pre_u8hwc_to_f32chw(const unsigned char* img, float* out, int B, int h, int w, int Hp, int Wp, cudaStream_t st);
int fdn_post_f32chw_to_u8hwc(const float* x, unsigned char* img, int B, int h, int w, int Hp, int Wp, cudaStream_t st);

/* ---- validation metrics on the device (basicsr/metrics/psnr_ssim.py) ---------------------------------
 * img1 / img2 [B][C][H][W] fp32 in [0,1] or [0,255] (the reference decides by img1.max() <= 1, psnr_ssim.py:58,317); results are
 * float64 like numpy's.  ws: caller-owned scratch of 4*B doubles.  crop_border as in the reference.
 * fdn_psnr: calculate_psnr (psnr_ssim.py:8-70); test_y_channel != 0 compares the BT.601 Y channel (metric_util.py:34-47).
 * fdn_ssim: calculate_ssim (psnr_ssim.py:243-329); mode 0 = ssim3d=True, the reference's default: 11x11x11 Gaussian over the
 *   (H, W, C) volume with replicate padding (:143-200); mode 1 = ssim3d=False: per-channel 11x11 Gaussian, valid region (:84-117);
 *   mode 2 = test_y_channel=True (:202-240). */
int fdn_psnr(const float* img1, const float* img2, double* psnr, double* ws, int B, int C, int H, int W, int crop_border,
             int test_y_channel, cudaStream_t st);
int fdn_ssim(const float* img1, const float* img2, double* ssim, double* ws, int B, int C, int H, int W, int crop_border, int mode,
             cudaStream_t st);

/* ---- training-side spectral losses, forward only (basicsr/models/losses/losses.py:83-115, 764-774) ---
 * Small kernels around the global FFT: element-wise difference, float64 reductions (mode 0: sum |a|, mode 1: sum (a-b)^2; *out is
 * zeroed by the call), and nn.Upsample(scale_factor=1/8, 'bilinear', align_corners=False) (losses.py:768). */
int fdn_diff(const float* a, const float* b, float* out, long long n, cudaStream_t st);
int fdn_reduce_f64(const float* a, const float* b, double* out, long long n, int mode, cudaStream_t st);
int fdn_down8_bilinear(const float* in, float* out, int planes, int H, int W, cudaStream_t st);

#ifdef __cplusplus
}
#endif
#endif /* FDN_B200_H */
