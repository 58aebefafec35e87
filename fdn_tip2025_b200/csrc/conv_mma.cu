// Dense 3x3 convolutions of the FDformer resamplers (Downsample / Upsample bodies, FDN_arch.py:715-734: 32->64, 64->128, 128->64,
// 64->32 channels at 1120x640; 24/48/96 for the LOL-v1 network) as implicit GEMMs on the tensor cores.
//
//   out[b, co, y, x] = bias[co] + res[b, co, y, x] + sum_{c, dy, dx} w[co, c, dy, dx] * in[b, c, y + dy - 1, x + dx - 1]     (zero padding)
//
// GEMM view: M = pixels, N = output channels, K = 9 * Cin, evaluated tap by tap: for a fixed tap the A operand is the input tile
// shifted by (dy, dx), so the im2col matrix never exists - a CTA stages an (8 + 2) x (32 + 2) pixel tile of eight input channels in
// shared memory once and reads it nine times.  The contraction runs on mma.sync.m16n8k8 (tf32 operands, fp32 accumulate) in 3xTF32:
// activations and weights are split into tf32 hi + lo and a_lo*b_hi + a_hi*b_lo + a_hi*b_hi is accumulated, which keeps fp32-level
// accuracy (the same split the tcgen05 1x1 kernel uses).  Why the warp-level MMA and not tcgen05 here: the A operand of a tap is a
// strided window of the staged tile, which a register-fragment MMA reads directly, while a UMMA descriptor needs a dense K-major
// panel per tap (nine re-packs of the tile); measured rate of this path on B200: 277 TFLOP/s tf32 (tools/ubench), i.e. ~90 TFLOP/s
// in 3xTF32 against the 27 TFLOP/s the FFMA kernel reaches on these layers.
//
// CTA = 8 warps = output tile of 8 rows x 32 columns x NT*8 output channels; warp w owns row w (two 16-pixel m-tiles).
// Shared memory per 8-channel K chunk: the input tile as hi and lo planes (plane stride 360 = 8 mod 32 floats, so the four channel
// columns of an A fragment fall on disjoint bank octets: conflict free), and the packed weights [hi,lo][tap][c][COP] (COP = 8 mod 32).
// (Interleaved (hi, lo) pairs fetched with 64-bit loads were measured slower: 13.0 vs 10.9 ms per 8-image step - the same number of
// shared-memory wavefronts and more registers.)
// Accuracy: like the tcgen05 unit, the warp-level MMA truncates its fp32 accumulator on every instruction, a bias that grows
// linearly with K (7.8e-6 relative at K = 1152 in one accumulator).  The accumulators are therefore flushed into IEEE fp32 totals
// every two K chunks (54 MMAs per element), which brings every layer back to ~1e-6; the second register set is why a CTA covers 32
// (or 24) output channels.
#include "fdn_common.cuh"

#ifndef FDN_EMU

#define CM_TW 32
#define CM_TH 8
#define CM_RS 36                      // tile row stride (34 used)
#define CM_PS (10 * CM_RS)            // plane stride: 360 = 8 (mod 32)
#define CM_FLUSH 2                    // K chunks per accumulator flush

struct ConvMmaParams {
    const float* in;       // [B][Cin][H][W]
    const float* wpack;    // [Cout/CN][Cin/8][2][9][8][COP]
    const float* bias;     // [Cout] or null
    const float* res;      // [B][Cout][H][W] or null
    float* out;            // [B][Cout][H][W]
    int Cin, Cout, H, W;
};

__device__ __forceinline__ void hmma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int NT>
__global__ void __launch_bounds__(256, 2) k_conv3x3_mma(ConvMmaParams q) {
    constexpr int COP = ((8 * NT + 31) / 32) * 32 + 8;
    constexpr int WCH = 2 * 9 * 8 * COP;                     // floats of one packed weight chunk
    FDN_DYN_SMEM(smem);
    float* s_hi = reinterpret_cast<float*>(smem);             // [8][10][CM_RS]
    float* s_lo = s_hi + 8 * CM_PS;
    float* s_w = s_lo + 8 * CM_PS;                            // [2][9][8][COP]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int tiles_x = (q.W + CM_TW - 1) / CM_TW;
    const int x0 = (blockIdx.x % tiles_x) * CM_TW, y0 = (blockIdx.x / tiles_x) * CM_TH;
    const int zc = blockIdx.y, b = blockIdx.z;
    const int nchunks = q.Cin >> 3;
    const size_t plane = (size_t)q.H * q.W;
    const float* inb = q.in + (size_t)b * q.Cin * plane;
    const float4* wsrc = reinterpret_cast<const float4*>(q.wpack + (size_t)zc * nchunks * WCH);

    float acc[2][NT][4], total[2][NT][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int j = 0; j < 4; ++j) { acc[mt][nt][j] = 0.f; total[mt][nt][j] = 0.f; }

    // The input tile of a chunk is fetched into registers one chunk ahead (the global-memory latency hides behind the MMAs of the
    // current chunk); element i of this thread is tile position tid + 256 * i of the 8 x 10 x 34 tile.
    constexpr int NPRE = (8 * 10 * 34 + 255) / 256;
    float pre[NPRE];
    auto fetch = [&](int chn) {
#pragma unroll
        for (int k = 0; k < NPRE; ++k) {
            const int i = tid + 256 * k;
            const int c = i / 340, r = i - c * 340;
            const int yy = r / 34, xx = r - yy * 34;
            const int gy = y0 + yy - 1, gx = x0 + xx - 1;
            float v = 0.f;
            if (i < 8 * 10 * 34 && gy >= 0 && gy < q.H && gx >= 0 && gx < q.W) v = inb[(size_t)(chn * 8 + c) * plane + (size_t)gy * q.W + gx];
            pre[k] = v;
        }
    };
    fetch(0);
    for (int ch = 0; ch < nchunks; ++ch) {
        __syncthreads();                                      // the previous chunk's fragments have been read
        // ---- stage the input tile of channels 8*ch .. 8*ch+7 with its one-pixel halo (zero outside the image), split into hi / lo
#pragma unroll
        for (int k = 0; k < NPRE; ++k) {
            const int i = tid + 256 * k;
            if (i < 8 * 10 * 34) {
                const int c = i / 340, r = i - c * 340;
                const int yy = r / 34, xx = r - yy * 34;
                const float v = pre[k];
                const float hi = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);
                s_hi[c * CM_PS + yy * CM_RS + xx] = hi;
                s_lo[c * CM_PS + yy * CM_RS + xx] = v - hi;       // the tensor core reads its tf32 bits
            }
        }
        {
            const float4* src = wsrc + (size_t)ch * (WCH / 4);
            float4* dst = reinterpret_cast<float4*>(s_w);
            for (int i = tid; i < WCH / 4; i += 256) dst[i] = src[i];
        }
        __syncthreads();
        if (ch + 1 < nchunks) fetch(ch + 1);
        // ---- nine taps: A = tile shifted by (dy, dx), B = the tap's [8 channels][NT*8 outputs] weights
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3, dx = tap - dy * 3;
            uint32_t bh[NT][2], bl[NT][2];
            const float* wh = s_w + tap * 8 * COP + t * COP + g;
            const float* wl = wh + 9 * 8 * COP;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                bh[nt][0] = __float_as_uint(wh[nt * 8]);
                bh[nt][1] = __float_as_uint(wh[4 * COP + nt * 8]);
                bl[nt][0] = __float_as_uint(wl[nt * 8]);
                bl[nt][1] = __float_as_uint(wl[4 * COP + nt * 8]);
            }
            const int abase = t * CM_PS + (warp + dy) * CM_RS + g + dx;
            uint32_t ah[2][4], al[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const float* ph = s_hi + abase + mt * 16;
                const float* pl = s_lo + abase + mt * 16;
                ah[mt][0] = __float_as_uint(ph[0]);             ah[mt][1] = __float_as_uint(ph[8]);
                ah[mt][2] = __float_as_uint(ph[4 * CM_PS]);     ah[mt][3] = __float_as_uint(ph[4 * CM_PS + 8]);
                al[mt][0] = __float_as_uint(pl[0]);             al[mt][1] = __float_as_uint(pl[8]);
                al[mt][2] = __float_as_uint(pl[4 * CM_PS]);     al[mt][3] = __float_as_uint(pl[4 * CM_PS + 8]);
            }
            // the three terms as three sweeps over the 2 x NT accumulators: consecutive MMAs never share an accumulator (small terms first)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) hmma_tf32(acc[mt][nt], al[mt], bh[nt][0], bh[nt][1]);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) hmma_tf32(acc[mt][nt], ah[mt], bl[nt][0], bl[nt][1]);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) hmma_tf32(acc[mt][nt], ah[mt], bh[nt][0], bh[nt][1]);
        }
        if ((ch % CM_FLUSH) == CM_FLUSH - 1 || ch == nchunks - 1) {       // IEEE fp32 totals (see the header)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int j = 0; j < 4; ++j) { total[mt][nt][j] += acc[mt][nt][j]; acc[mt][nt][j] = 0.f; }
        }
    }
    // ---- epilogue: C fragment (row = pixel g / g+8, columns 2t, 2t+1 = output channels) -> NCHW
    const int oy = y0 + warp;
    if (oy >= q.H) return;
    const int CN = 8 * NT;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int co = zc * CN + nt * 8 + 2 * t + (j & 1);
                const int ox = x0 + mt * 16 + g + ((j & 2) ? 8 : 0);
                if (ox < q.W && co < q.Cout) {
                    const size_t o = ((size_t)b * q.Cout + co) * plane + (size_t)oy * q.W + ox;
                    float v = total[mt][nt][j];
                    if (q.bias) v += q.bias[co];
                    if (q.res) v += q.res[o];
                    q.out[o] = v;
                }
            }
}

template <int NT>
static int launch_conv_mma(const ConvMmaParams& q, int B, cudaStream_t st) {
    constexpr int COP = ((8 * NT + 31) / 32) * 32 + 8;
    const size_t smem = (size_t)(2 * 8 * CM_PS + 2 * 9 * 8 * COP) * sizeof(float);
    auto kern = k_conv3x3_mma<NT>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { fdn_set_error(cudaGetErrorString(e)); return (int)e; }
    }
    dim3 grid(fdn_cdiv(q.W, CM_TW) * fdn_cdiv(q.H, CM_TH), q.Cout / (8 * NT), B);
    kern<<<grid, dim3(256), smem, st>>>(q);
    return fdn_check_launch("k_conv3x3_mma");
}
#endif  // !FDN_EMU

// Output channels per CTA for a layer with Cout outputs: 32 or 24, whichever divides Cout (0: unsupported).
FDN_API int fdn_conv3x3_mma_cn(int Cout) {
    if (Cout % 32 == 0) return 32;
    if (Cout % 24 == 0) return 24;
    return 0;
}

// Tensor-core 3x3 convolution, stride 1, padding 1 (3xTF32: fp32-level accuracy).  Cin a multiple of 8, Cout a multiple of CN =
// fdn_conv3x3_mma_cn(Cout).  wpack = host-packed weights (fdn_tip2025_b200/packing.py::pack_conv3x3):
// [Cout/CN][Cin/8][hi,lo][tap][8 channels][COP = 40].  out = conv + bias + res.
FDN_API int fdn_conv3x3_mma(const float* in, const float* wpack, const float* bias, const float* res, float* out, int B, int Cin,
                            int H, int W, int Cout, cudaStream_t st) {
#ifdef FDN_EMU
    fdn_set_error("fdn_conv3x3_mma: tensor-core kernels are not available in the host emulation build");
    return -1;
#else
    FDN_REQUIRE(in && wpack && out && B > 0 && H > 0 && W > 0, "bad arguments");
    FDN_REQUIRE(Cin > 0 && Cin % 8 == 0, "Cin must be a multiple of 8");
    const int CN = fdn_conv3x3_mma_cn(Cout);
    FDN_REQUIRE(CN > 0, "Cout must be a multiple of 24 or 32");
    FDN_REQUIRE(fdn_aligned16(wpack), "wpack must be 16-byte aligned");
    FDN_REQUIRE(B <= 65535 && Cout / CN <= 65535, "grid too large");
    ConvMmaParams q{in, wpack, bias, res, out, Cin, Cout, H, W};
    return CN == 32 ? launch_conv_mma<4>(q, B, st) : launch_conv_mma<3>(q, B, st);
#endif
}
