// Whole-plane 2-D real FFTs (torch.fft.rfft2 / irfft2, norm='backward') for the global spectral stages:
// FCAFFN (FDN_arch.py:410-420), MAR FreBlock / fourier_fuse (FDN_arch.py:88-100, 136-148) and the FDN
// prologue (FDN_arch.py:882-914).
//
// Structure (data stays fp32; spectra are float2 [plane][H][W/2+1]):
//   rows R2C  : real rows -> half spectrum rows                       (k_rows_r2c)
//   columns   : forward FFT along H on a tile of columns held in shared memory, the per-bin spectral
//               operator applied in place (FCAFFN modulation / angle / abs), and - for FCAFFN - the
//               inverse column FFT in the same kernel, so the modulated spectrum never goes to HBM
//               between the two column passes                          (k_cols)
//   rows C2R  : Hermitian rows -> real rows, 1/(H*W) scale, residual epilogue (k_rows_c2r)
// Every 1-D transform is a mixed-radix Stockham autosort FFT in shared memory with register butterflies
// for radix 2/3/4/5/7/8 and a direct O(p^2) butterfly for any other prime factor (needed by the
// (H+2)x(W+2) transforms of fourier_fuse: 642 = 2*3*107, 562 = 2*281, ...).  Twiddles come from a
// per-length table computed in double precision on the host.
#include "fdn_common.cuh"
#include "fft8.cuh"

#include <map>
#include <mutex>
#include <vector>

#define FDN_FFT_MAX_PASSES 16

struct FftPlanDev {
    int N;
    int npass;
    int radix[FDN_FFT_MAX_PASSES];
    const float2* tw;   // tw[k] = exp(-2 pi i k / N), k in [0, N)
};

// ---------------------------------------------------------------------------------------------------
// host: plan cache
// ---------------------------------------------------------------------------------------------------
// Twiddle tables live in the memory of the device that was current when they were created, so the cache key is (device, N).
static std::mutex g_plan_mutex;
static std::map<std::pair<int, int>, FftPlanDev> g_plans;

// st: the stream the first kernel using the table will run on.  A missing table cannot be created while that stream is being
// captured into a CUDA graph (cudaMalloc / a synchronous upload are illegal there): call fdn_fft_prepare(H, W) before capturing.
static int fdn_fft_get_plan(int N, FftPlanDev* out, cudaStream_t st = nullptr) {
    std::lock_guard<std::mutex> lock(g_plan_mutex);
    const std::pair<int, int> key(fdn_device(), N);
    auto it = g_plans.find(key);
    if (it != g_plans.end()) {
        *out = it->second;
        return 0;
    }
    if (st != nullptr) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone) {
            fdn_set_error("FFT twiddle table for this length is missing while the stream is capturing: call fdn_fft_prepare(H, W) first");
            return -3;
        }
    }
    FftPlanDev p;
    p.N = N;
    p.npass = 0;
    int n = N;
    // Prime factors without an unrolled butterfly (23, 107, 281, ... of the fourier_fuse sizes) take the O(p^2) pass; they go FIRST:
    // the first Stockham pass has no input twiddles (Ns = 1), which is what makes stockham_pass_prime_first cheap.
    int small[FDN_FFT_MAX_PASSES], nsmall = 0;
    const int pref[] = {8, 4, 2, 3, 5, 7, 11, 13, 17, 19};
    for (int r : pref)
        while (n % r == 0 && n > 1) {
            if (nsmall >= FDN_FFT_MAX_PASSES) return -1;
            small[nsmall++] = r;
            n /= r;
        }
    for (int f = 23; n > 1; f += 2) {
        if (f * f > n) f = n;
        while (n % f == 0) {
            if (p.npass >= FDN_FFT_MAX_PASSES) return -1;
            p.radix[p.npass++] = f;
            n /= f;
        }
    }
    if (p.npass + nsmall > FDN_FFT_MAX_PASSES) return -1;
    for (int i = 0; i < nsmall; ++i) p.radix[p.npass++] = small[i];
    std::vector<float2> tw(N);
    for (int k = 0; k < N; ++k) {
        // exact values on the axes so purely real paths stay exactly real
        if ((4LL * k) % N == 0) {
            int q = (int)((4LL * k) / N);
            const float c[4] = {1.f, 0.f, -1.f, 0.f}, s[4] = {0.f, -1.f, 0.f, 1.f};
            tw[k] = make_float2(c[q], s[q]);
        } else {
            double a = -2.0 * M_PI * (double)k / (double)N;
            tw[k] = make_float2((float)cos(a), (float)sin(a));
        }
    }
    float2* dev = nullptr;
    if (cudaMalloc((void**)&dev, sizeof(float2) * N) != cudaSuccess) return -2;
    if (cudaMemcpy(dev, tw.data(), sizeof(float2) * N, cudaMemcpyHostToDevice) != cudaSuccess) return -2;
    // the copy from pageable memory may return before the DMA has landed, and the caller's (non-blocking) stream is not ordered
    // against the legacy default stream: wait for the device once, at table creation
    if (cudaDeviceSynchronize() != cudaSuccess) return -2;
    p.tw = dev;
    g_plans[key] = p;
    *out = p;
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// device: Stockham passes in shared memory
// ---------------------------------------------------------------------------------------------------
template <int SGN>
__device__ __forceinline__ float2 tw_mul(float2 v, float2 w) {   // v * (SGN<0 ? w : conj(w))
    return SGN < 0 ? cmul(v, w) : cmulc(v, w);
}

template <int R> struct Roots;   // cos/sin of 2 pi k / R
template <> struct Roots<3> {
    __device__ static __forceinline__ float c(int k) { return k == 0 ? 1.f : -0.5f; }
    __device__ static __forceinline__ float s(int k) { return k == 0 ? 0.f : (k == 1 ? 0.86602540378443864676f : -0.86602540378443864676f); }
};
template <> struct Roots<5> {
    __device__ static __forceinline__ float c(int k) {
        return k == 0 ? 1.f : ((k == 1 || k == 4) ? 0.30901699437494742410f : -0.80901699437494742410f);
    }
    __device__ static __forceinline__ float s(int k) {
        return k == 0 ? 0.f : (k == 1 ? 0.95105651629515357212f : (k == 2 ? 0.58778525229247312917f
                     : (k == 3 ? -0.58778525229247312917f : -0.95105651629515357212f)));
    }
};
template <> struct Roots<7> {
    __device__ static __forceinline__ float c(int k) {
        return k == 0 ? 1.f : ((k == 1 || k == 6) ? 0.62348980185873353053f
                     : ((k == 2 || k == 5) ? -0.22252093395631440429f : -0.90096886790241912624f));
    }
    __device__ static __forceinline__ float s(int k) {
        const float s1 = 0.78183148246802980871f, s2 = 0.97492791218182360702f, s3 = 0.43388373911755812048f;
        return k == 0 ? 0.f : (k == 1 ? s1 : (k == 2 ? s2 : (k == 3 ? s3 : (k == 4 ? -s3 : (k == 5 ? -s2 : -s1)))));
    }
};

#include "fft_roots.cuh"      // Roots<11>, <13>, <17>, <19>

template <int R, int SGN>
__device__ __forceinline__ void butterfly(float2 v[R]) {
    if constexpr (R == 2) {
        float2 a = v[0];
        v[0] = cadd(a, v[1]);
        v[1] = csub(a, v[1]);
    } else if constexpr (R == 4) {
        float2 a0 = cadd(v[0], v[2]), a1 = csub(v[0], v[2]), a2 = cadd(v[1], v[3]), a3 = csub(v[1], v[3]);
        float2 ia3 = make_float2(-SGN * a3.y, SGN * a3.x);
        v[0] = cadd(a0, a2);
        v[2] = csub(a0, a2);
        v[1] = cadd(a1, ia3);
        v[3] = csub(a1, ia3);
    } else if constexpr (R == 8) {
        fft8_c2c<SGN>(v);
    }
}
// Odd-length DFT (R in {3,5,7}) through the conjugate-pair symmetry of the roots: with a_n = v[n] + v[R-n], b_n = v[n] - v[R-n],
//   X[k] = v[0] + sum_n a_n cos(2 pi n k / R) + i SGN sum_n b_n sin(2 pi n k / R),   X[R-k] = the same with -i,
// i.e. 2 (R-1) real FMAs per output PAIR instead of 4 (R-1) per output for the plain sum (radix 7: 9 operations per output, was 24 -
// the radix-7 / radix-10 passes were the most expensive part of the row transforms).
template <int R, int SGN>
__device__ __forceinline__ void butterfly_direct(float2 v[R]) {
    constexpr int Hf = (R - 1) / 2;
    float2 a[Hf], ib[Hf];                     // ib_n = i b_n = (-b_n.y, b_n.x): formed by the subtraction itself, so everything after it is
    float2 sum = v[0];                        // a complex * real FMA or a complex add (packed FFMA2 / FADD2)
#pragma unroll
    for (int n = 1; n <= Hf; ++n) {
        a[n - 1] = cadd(v[n], v[R - n]);
        ib[n - 1] = make_float2(v[R - n].y - v[n].y, v[n].x - v[R - n].x);
        sum = cadd(sum, a[n - 1]);
    }
    const float2 v0 = v[0];
    v[0] = sum;
#pragma unroll
    for (int k = 1; k <= Hf; ++k) {
        float2 A = v0, B = make_float2(0.f, 0.f);
#pragma unroll
        for (int n = 1; n <= Hf; ++n) {
            const int m = (n * k) % R;
            A = cfma(a[n - 1], Roots<R>::c(m), A);
            B = cfma(ib[n - 1], Roots<R>::s(m), B);
        }
        // X[k] = A + SGN i B_plain = A + SGN B,  X[R-k] = A - SGN B
        v[k] = SGN > 0 ? cadd(A, B) : csub(A, B);
        v[R - k] = SGN > 0 ? csub(A, B) : cadd(A, B);
    }
}

// exact n / d for 0 <= n < 2^22 with rcp = 1.0f / d (one multiply instead of an integer division)
__device__ __forceinline__ int fast_div(int n, float rcp) { return __float2int_rz(((float)n + 0.5f) * rcp); }

// One radix-R Stockham pass over nseq sequences.  Element n of sequence s lives at s*ss + n*es.
// COLS = false: sequences are contiguous rows (es == 1): warps stride over sequences, lanes over butterflies.
// COLS = true : sequences are the nseq = 2^lg interleaved columns of a tile (ss == 1, es == nseq).
// No integer divisions: j / Ns uses an exact float reciprocal.
template <int R, int SGN, bool COLS>
__device__ __forceinline__ void stockham_pass(const float2* __restrict__ in, float2* __restrict__ out, const FftPlanDev& P,
                                              int Ns, int nseq, int ss, int es, int lg) {
    const int N = P.N, NR = N / R, M = N / (Ns * R);
    const float rcpNs = 1.0f / (float)Ns;
    auto one = [&](int seq, int j) {
        const int jh = fast_div(j, rcpNs), k = j - jh * Ns;
        const float2* src = in + seq * ss;
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = src[(j + r * NR) * es];
        if (Ns > 1) {
#pragma unroll
            for (int r = 1; r < R; ++r) v[r] = tw_mul<SGN>(v[r], P.tw[r * k * M]);
        }
        if constexpr (R % 2 == 1) butterfly_direct<R, SGN>(v); else butterfly<R, SGN>(v);
        float2* dst = out + seq * ss;
        if (!COLS && Ns == 1 && (R % 2 == 0)) {
            // first pass of a row transform: the R outputs of a butterfly are contiguous -> 128-bit conflict-free stores
            float4* d4 = reinterpret_cast<float4*>(dst + j * R);
#pragma unroll
            for (int r = 0; r < R; r += 2) d4[r >> 1] = make_float4(v[r].x, v[r].y, v[r + 1].x, v[r + 1].y);
        } else {
            const int j0 = jh * Ns * R + k;
#pragma unroll
            for (int r = 0; r < R; ++r) dst[(j0 + r * Ns) * es] = v[r];
        }
    };
    if (COLS) {
        const int total = NR << lg;
        for (int idx = threadIdx.x; idx < total; idx += blockDim.x) one(idx & (nseq - 1), idx >> lg);
    } else {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
        for (int seq = warp; seq < nseq; seq += nwarps)
            for (int j = lane; j < NR; j += 32) one(seq, j);
    }
}

// Any other odd prime radix (23, 61, 107, 281, ...): one thread per output PAIR (r, R - r) of a butterfly.  With the butterfly's twiddled
// inputs y[q] = x[q] w^{q k M}, a_q = y[q] + y[R-q], b_q = y[q] - y[R-q]:
//   Y[r] = y[0] + sum_{q=1..h} a_q cos(2 pi q r / R) + i SGN sum_q b_q sin(2 pi q r / R),   Y[R-r] = the same with -i        (h = (R-1)/2)
// - half the terms of the plain sum and two outputs from them.  The R roots come from the length-N table at stride N / R.
template <int SGN, bool COLS>
__device__ __forceinline__ void stockham_pass_generic(const float2* __restrict__ in, float2* __restrict__ out, const FftPlanDev& P,
                                                      int R, int Ns, int nseq, int ss, int es, int lg) {
    const int N = P.N, NR = N / R, M = N / (Ns * R), NRs = N / R;      // NRs: stride of the R-th roots in the table
    const int h = (R - 1) >> 1, npair = h + 1;                          // pair 0 = output 0 alone
    const float rcpNs = 1.0f / (float)Ns, rcpP = 1.0f / (float)npair;
    auto one = [&](int seq, int o) {                                    // o = j * npair + r,  j = butterfly (jhi * Ns + k)
        const int j = fast_div(o, rcpP), r = o - j * npair;
        const int jhi = fast_div(j, rcpNs), k = j - jhi * Ns;
        const float2* src = in + seq * ss + j * es;
        const int tstep = k * M;                                        // input twiddle exponent step: y[q] = x[q] w_N^{q k M}
        const int rstep = r * NRs;                                      // root exponent step: w_R^{q r} = w_N^{q r N / R}
        const float2 y0 = src[0];
        float2 A = y0, B = make_float2(0.f, 0.f);
        int tq = 0, tRq = 0, rq = 0;                                    // exponents of y[q], y[R-q] twiddles and of the root, mod N
        // y[R-q] twiddle exponent (R-q) k M = R k M - q k M = (N / Ns) k - q k M  (mod N)
        const int tR0 = (int)(((long long)(N / Ns) * k) % N);
        for (int q = 1; q <= h; ++q) {
            tq += tstep; if (tq >= N) tq -= N;
            rq += rstep; if (rq >= N) rq -= N;
            tRq = tR0 - tq; if (tRq < 0) tRq += N;
            const float2 yq = tw_mul<SGN>(src[(size_t)q * NR * es], P.tw[tq]);
            const float2 yr = tw_mul<SGN>(src[(size_t)(R - q) * NR * es], P.tw[tRq]);
            const float2 w = P.tw[rq];                                  // (cos, -sin) of 2 pi q r / R
            A.x = fmaf(yq.x + yr.x, w.x, A.x);
            A.y = fmaf(yq.y + yr.y, w.x, A.y);
            B.x = fmaf(yq.x - yr.x, -w.y, B.x);
            B.y = fmaf(yq.y - yr.y, -w.y, B.y);
        }
        float2* dst = out + seq * ss;
        const int o0 = (jhi * R) * Ns + k;                              // output r of the butterfly lives at (jhi * R + r) * Ns + k
        dst[(size_t)(o0 + r * Ns) * es] = make_float2(A.x - SGN * B.y, A.y + SGN * B.x);
        if (r > 0) dst[(size_t)(o0 + (R - r) * Ns) * es] = make_float2(A.x + SGN * B.y, A.y - SGN * B.x);
    };
    if (COLS) {
        const int total = (NR * npair) << lg;
        for (int idx = threadIdx.x; idx < total; idx += blockDim.x) one(idx & (nseq - 1), idx >> lg);
    } else {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
        for (int seq = warp; seq < nseq; seq += nwarps)
            for (int o = lane; o < NR * npair; o += 32) one(seq, o);
    }
}

// The same prime-radix butterfly as the FIRST pass of a transform (Ns = 1: no input twiddles, butterfly j reads x[j + q N/R] and writes
// X[j R + r]).  A thread produces G output pairs from one sweep over the inputs, so an input pair is read once per G outputs and
// the inner loop is one root lookup + four FMAs per output pair (the general pass above: two twiddle products and three table
// reads per term).  Work items are ordered (group, butterfly[, column]) so that lanes read consecutive inputs and the same root.
template <int SGN, bool COLS, int G>
__device__ __forceinline__ void stockham_pass_prime_first(const float2* __restrict__ in, float2* __restrict__ out, const FftPlanDev& P,
                                                          int R, int nseq, int ss, int es, int lg) {
    const int N = P.N, NR = N / R;                                      // NR butterflies; the R-th roots sit at stride NR in the table
    const int h = (R - 1) >> 1, npair = h + 1, ngrp = (npair + G - 1) / G;
    const float rcpNR = 1.0f / (float)NR;
    auto one = [&](int seq, int o) {                                    // o = g * NR + j
        const int g = fast_div(o, rcpNR), j = o - g * NR;
        const float2* src = in + seq * ss + j * es;
        int rq[G], rstep[G];
        float2 A[G], B[G];
        const float2 y0 = src[0];
#pragma unroll
        for (int i = 0; i < G; ++i) {
            const int r = min(g * G + i, h);
            rstep[i] = r * NR;
            rq[i] = 0;
            A[i] = y0;
            B[i] = make_float2(0.f, 0.f);
        }
        for (int q = 1; q <= h; ++q) {
            const float2 yq = src[(size_t)q * NR * es], yr = src[(size_t)(R - q) * NR * es];
            const float2 a = cadd(yq, yr), b = csub(yq, yr);
#pragma unroll
            for (int i = 0; i < G; ++i) {
                rq[i] += rstep[i];
                if (rq[i] >= N) rq[i] -= N;
                const float2 w = P.tw[rq[i]];                           // (cos, -sin) of 2 pi q r / R
                A[i].x = fmaf(a.x, w.x, A[i].x);
                A[i].y = fmaf(a.y, w.x, A[i].y);
                B[i].x = fmaf(b.x, -w.y, B[i].x);
                B[i].y = fmaf(b.y, -w.y, B[i].y);
            }
        }
        float2* dst = out + seq * ss;
#pragma unroll
        for (int i = 0; i < G; ++i) {
            const int r = g * G + i;
            if (r <= h) {
                dst[(size_t)(j * R + r) * es] = make_float2(A[i].x - SGN * B[i].y, A[i].y + SGN * B[i].x);
                if (r > 0) dst[(size_t)(j * R + R - r) * es] = make_float2(A[i].x + SGN * B[i].y, A[i].y - SGN * B[i].x);
            }
        }
    };
    if (COLS) {
        const int total = (NR * ngrp) << lg;
        for (int idx = threadIdx.x; idx < total; idx += blockDim.x) one(idx & (nseq - 1), idx >> lg);
    } else {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
        for (int seq = warp; seq < nseq; seq += nwarps)
            for (int o = lane; o < NR * ngrp; o += 32) one(seq, o);
    }
}

// Full FFT of nseq sequences; ping-pongs between a and b, returns the buffer holding the result.
// Must be called by all threads of the block; ends with a __syncthreads().
template <int SGN, bool COLS>
__device__ float2* fft_smem(const FftPlanDev& P, float2* a, float2* b, int nseq, int ss, int es) {
    int Ns = 1;
    int lg = 0;
    while ((1 << lg) < nseq) ++lg;           // COLS: nseq is a power of two
    for (int p = 0; p < P.npass; ++p) {
        const int R = P.radix[p];
        switch (R) {
            case 2: stockham_pass<2, SGN, COLS>(a, b, P, Ns, nseq, ss, es, lg); break;
            case 3: stockham_pass<3, SGN, COLS>(a, b, P, Ns, nseq, ss, es, lg); break;
            case 4: stockham_pass<4, SGN, COLS>(a, b, P, Ns, nseq, ss, es, lg); break;
            case 5: stockham_pass<5, SGN, COLS>(a, b, P, Ns, nseq, ss, es, lg); break;
            case 7: stockham_pass<7, SGN, COLS>(a, b, P, Ns, nseq, ss, es, lg); break;
            case 8: stockham_pass<8, SGN, COLS>(a, b, P, Ns, nseq, ss, es, lg); break;
            // unrolled odd radices of the 608x416 (13, 19), 3840x2176 (17) and fourier_fuse (11, 17, 19) transforms
            case 11: stockham_pass<11, SGN, COLS>(a, b, P, Ns, nseq, ss, es, lg); break;
            case 13: stockham_pass<13, SGN, COLS>(a, b, P, Ns, nseq, ss, es, lg); break;
            case 17: stockham_pass<17, SGN, COLS>(a, b, P, Ns, nseq, ss, es, lg); break;
            case 19: stockham_pass<19, SGN, COLS>(a, b, P, Ns, nseq, ss, es, lg); break;
            default:
                if (Ns == 1) stockham_pass_prime_first<SGN, COLS, 5>(a, b, P, R, nseq, ss, es, lg);
                else stockham_pass_generic<SGN, COLS>(a, b, P, R, Ns, nseq, ss, es, lg);
                break;
        }
        __syncthreads();
        Ns *= R;
        float2* t = a; a = b; b = t;
    }
    return a;
}

// ---------------------------------------------------------------------------------------------------
// rows: real -> half spectrum, W even.  The W real samples of a row are packed into M = W/2 complex numbers
// z[n] = x[2n] + i x[2n+1]; one length-M complex FFT gives Z, and
//   X[k] = E[k] + w^k O[k],  E[k] = (Z[k] + conj Z[M-k]) / 2,  O[k] = -i (Z[k] - conj Z[M-k]) / 2,  w = e^{-2 pi i / W}
// for k = 0..M (Z[M] = Z[0]).  Half the butterflies, shared memory and barriers of a length-W complex transform.
// ---------------------------------------------------------------------------------------------------
// in: nrows rows of W floats, row r at in + r*W.  out: row r at out + r*Wf, Wf = W/2+1.  PM = plan of length M, twW = table of W.
__global__ void __launch_bounds__(256) k_rows_r2c(const float* __restrict__ in, float2* __restrict__ out, FftPlanDev PM,
                                                  const float2* __restrict__ twW, int nrows, int rows_per_cta) {
    FDN_DYN_SMEM(smem);
    const int M = PM.N, Wf = M + 1, MS = (M + 1) & ~1;        // even row stride keeps the 128-bit stores of the first pass aligned
    float2* a = reinterpret_cast<float2*>(smem);
    float2* b = a + rows_per_cta * MS;
    float2* s_tw = b + rows_per_cta * MS;                       // pass twiddles staged in shared memory
    const int row0 = blockIdx.x * rows_per_cta;
    const int S = min(rows_per_cta, nrows - row0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int i = threadIdx.x; i < M; i += blockDim.x) s_tw[i] = PM.tw[i];
    for (int r = warp; r < S; r += nwarps) {
        const float2* src = reinterpret_cast<const float2*>(in + (size_t)(row0 + r) * 2 * M);   // rows are 8-byte aligned (W even)
        for (int n = lane; n < M; n += 32) a[r * MS + n] = src[n];
    }
    PM.tw = s_tw;
    __syncthreads();
    float2* res = fft_smem<-1, false>(PM, a, b, S, MS, 1);
    for (int r = warp; r < S; r += nwarps) {
        const float2* z = res + r * MS;
        float2* dst = out + (size_t)(row0 + r) * Wf;
        for (int k = lane; k < Wf; k += 32) {
            const float2 zk = z[k == M ? 0 : k];
            const float2 zc = z[k == 0 ? 0 : M - k];                     // conj applied below
            const float ex = 0.5f * (zk.x + zc.x), ey = 0.5f * (zk.y - zc.y);
            const float dx = zk.x - zc.x, dy = zk.y + zc.y;               // D = Z[k] - conj Z[M-k]
            const float ox = 0.5f * dy, oy = -0.5f * dx;                  // O = -i D / 2
            const float2 w = twW[k];
            float2 v = make_float2(ex + (w.x * ox - w.y * oy), ey + (w.x * oy + w.y * ox));
            if (k == 0 || k == M) v.y = 0.f;                              // exact for real input
            dst[k] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// rows: half spectrum -> real (inverse of the packing above), epilogue  out = img_scale[b] * (irfft * norm + res_coef * res)
//   Z[k] = E[k] + i O[k],  E[k] = (X[k] + conj X[M-k]) / 2,  O[k] = conj(w^k) (X[k] - conj X[M-k]) / 2,  k = 0..M-1
//   z = IDFT_M(Z) = M (x[2n] + i x[2n+1])
// The imaginary parts of X[0] and X[M] are ignored, as torch.fft.irfft does.
// ---------------------------------------------------------------------------------------------------
struct RowsC2RParams {
    const float2* in;     // [nrows][Wf]
    float* out;           // [nrows][W]
    const float* res;     // [nrows][W] or null
    const float* img_scale;   // [B] or null
    float res_coef;
    float norm;           // 1/(H*W)
    int nrows;
    int rows_per_cta;
    int rows_per_image;   // C*H, to find b for img_scale
};

__global__ void __launch_bounds__(256) k_rows_c2r(RowsC2RParams q, FftPlanDev PM, const float2* __restrict__ twW) {
    FDN_DYN_SMEM(smem);
    const int M = PM.N, Wf = M + 1, MS = (M + 1) & ~1;
    float2* a = reinterpret_cast<float2*>(smem);
    const int half_elems = (q.rows_per_cta * (M + 1) + 1) & ~1;       // keep both ping-pong buffers 16-byte aligned
    float2* b = a + half_elems;
    float2* s_tw = b + half_elems;
    const int row0 = blockIdx.x * q.rows_per_cta;
    const int S = min(q.rows_per_cta, q.nrows - row0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int i = threadIdx.x; i < M; i += blockDim.x) s_tw[i] = PM.tw[i];
    PM.tw = s_tw;
    // stage the half spectrum (M+1 bins per row) in b, then build Z in a
    for (int r = warp; r < S; r += nwarps) {
        const float2* src = q.in + (size_t)(row0 + r) * Wf;
        for (int k = lane; k < Wf; k += 32) {
            float2 v = src[k];
            if (k == 0 || k == M) v.y = 0.f;
            b[r * (M + 1) + k] = v;
        }
    }
    __syncthreads();
    for (int r = warp; r < S; r += nwarps) {
        const float2* x = b + r * (M + 1);
        for (int k = lane; k < M; k += 32) {
            const float2 xk = x[k], xc = x[M - k];
            const float ex = 0.5f * (xk.x + xc.x), ey = 0.5f * (xk.y - xc.y);
            const float tx = 0.5f * (xk.x - xc.x), ty = 0.5f * (xk.y + xc.y);     // T = (X[k] - conj X[M-k]) / 2
            const float2 w = twW[k];                                               // O = conj(w) T
            const float ox = w.x * tx + w.y * ty, oy = w.x * ty - w.y * tx;
            a[r * MS + k] = make_float2(ex - oy, ey + ox);                         // E + i O
        }
    }
    __syncthreads();
    float2* res = fft_smem<1, false>(PM, a, b, S, MS, 1);
    const float nrm = 2.0f * q.norm;                                               // IDFT_M gives (W/2) x
    for (int r = warp; r < S; r += nwarps) {
        const float2* z = res + r * MS;
        float2* dst = reinterpret_cast<float2*>(q.out + (size_t)(row0 + r) * 2 * M);
        const float2* rsrc = q.res ? reinterpret_cast<const float2*>(q.res + (size_t)(row0 + r) * 2 * M) : nullptr;
        const float sc = q.img_scale ? q.img_scale[(row0 + r) / q.rows_per_image] : 1.0f;
        for (int n = lane; n < M; n += 32) {
            float2 v = make_float2(z[n].x * nrm, z[n].y * nrm);
            if (rsrc) { const float2 t = rsrc[n]; v.x += q.res_coef * t.x; v.y += q.res_coef * t.y; }
            v.x *= sc; v.y *= sc;
            dst[n] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// columns
// ---------------------------------------------------------------------------------------------------
enum { COLS_FWD = 0, COLS_INV = 1, COLS_FWD_MOD_INV = 2, COLS_FWD_ANGLE = 3, COLS_FWD_ABS = 4 };

struct ColsParams {
    const float2* in;         // spectrum, element (plane, y, x) at in[plane*in_ps + y*in_rs + x]
    float2* out;              // complex output (modes 0..2)
    float* out_real;          // real output (modes 3,4), same indexing as out
    long long in_ps, out_ps;
    int in_rs, out_rs;
    int ncols;                // columns to process (W/2+1 of the *output* transform)
    int W;                    // real width (to locate the Nyquist column), 0 = do not force self-conjugate bins
    int tc;                   // columns per CTA
    int mode;
    // FCAFFN modulation (mode 2): plane = b*C + c
    int C;
    const float* amp;         // [B][3][H][ncols]
    const float* pha;         // [B][3][H][ncols]
    const float* w_xa;        // [C][3]
    const float* w_xp;        // [C][3]
};

__global__ void __launch_bounds__(256) k_cols(ColsParams q, FftPlanDev P) {
    FDN_DYN_SMEM(smem);
    const int H = P.N, tc = q.tc;
    float2* a = reinterpret_cast<float2*>(smem);
    float2* b = a + H * tc;
    float2* s_tw = b + H * tc;
    for (int i = threadIdx.x; i < H; i += blockDim.x) s_tw[i] = P.tw[i];
    P.tw = s_tw;
    const int plane = blockIdx.y;
    const int c0 = blockIdx.x * tc;
    const int nc = min(tc, q.ncols - c0);
    const float2* src = q.in + (size_t)plane * q.in_ps + c0;
    for (int i = threadIdx.x; i < H * tc; i += blockDim.x) {
        int y = i / tc, c = i - y * tc;
        a[i] = c < nc ? src[(size_t)y * q.in_rs + c] : make_float2(0.f, 0.f);
    }
    __syncthreads();
    float2* res;
    if (q.mode == COLS_INV) {
        res = fft_smem<1, true>(P, a, b, tc, 1, tc);
    } else {
        res = fft_smem<-1, true>(P, a, b, tc, 1, tc);
        float2* other = (res == a) ? b : a;
        // the four self-conjugate bins of a real signal's spectrum are exactly real
        if (q.W > 0) {
            for (int i = threadIdx.x; i < 2 * tc; i += blockDim.x) {
                int c = i % tc, yy = i / tc;
                int x = c0 + c;
                bool xs = (x == 0) || (2 * x == q.W);
                int y = yy == 0 ? 0 : H / 2;
                if (xs && (yy == 0 || (H % 2 == 0))) res[y * tc + c].y = 0.f;
            }
            __syncthreads();
        }
        if (q.mode == COLS_FWD_MOD_INV) {
            const int bimg = plane / q.C, ch = plane - bimg * q.C;
            const float a0 = q.w_xa[ch * 3 + 0], a1 = q.w_xa[ch * 3 + 1], a2 = q.w_xa[ch * 3 + 2];
            const float p0 = q.w_xp[ch * 3 + 0], p1 = q.w_xp[ch * 3 + 1], p2 = q.w_xp[ch * 3 + 2];
            const size_t mstride = (size_t)H * q.ncols;
            const float* ampb = q.amp + (size_t)bimg * 3 * mstride + c0;
            const float* phab = q.pha + (size_t)bimg * 3 * mstride + c0;
            for (int i = threadIdx.x; i < H * tc; i += blockDim.x) {
                int y = i / tc, c = i - y * tc;
                if (c < nc) {
                    size_t m = (size_t)y * q.ncols + c;
                    float A = a0 * ampb[m] + a1 * ampb[m + mstride] + a2 * ampb[m + 2 * mstride];
                    float Pp = p0 * phab[m] + p1 * phab[m + mstride] + p2 * phab[m + 2 * mstride];
                    float sn, cs;
                    sincosf(Pp, &sn, &cs);
                    float2 z = res[i];
                    z.x = fdn_rd(z.x);
                    z.y = fdn_rd(z.y);
                    // A * z * e^{-iP}
                    res[i] = make_float2(A * (z.x * cs + z.y * sn), A * (z.y * cs - z.x * sn));
                }
            }
            __syncthreads();
            res = fft_smem<1, true>(P, res, other, tc, 1, tc);
        }
    }
    if (q.mode == COLS_FWD_ANGLE || q.mode == COLS_FWD_ABS) {
        float* dst = q.out_real + (size_t)plane * q.out_ps + c0;
        for (int i = threadIdx.x; i < H * tc; i += blockDim.x) {
            int y = i / tc, c = i - y * tc;
            if (c < nc) {
                float2 z = res[i];
                float v;
                if (q.mode == COLS_FWD_ANGLE) v = atan2f(fdn_rd(z.y), fdn_rd(z.x));
                else v = sqrtf(z.x * z.x + z.y * z.y);
                dst[(size_t)y * q.out_rs + c] = v;
            }
        }
    } else {
        float2* dst = q.out + (size_t)plane * q.out_ps + c0;
        for (int i = threadIdx.x; i < H * tc; i += blockDim.x) {
            int y = i / tc, c = i - y * tc;
            if (c < nc) dst[(size_t)y * q.out_rs + c] = res[i];
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// MAR per-bin channel MLPs (FreBlock / fourier_fuse): mag, pha -> (1x1, LeakyReLU 0.1, 1x1) each -> polar
// ---------------------------------------------------------------------------------------------------
struct SpecMlpParams {
    float2* spec;             // [B][NC][H*Wf] complex, in place
    long long plane_stride;   // elements between channels
    long long nbins;          // H*Wf
    int B;
    const float* w;           // packed: W1m[NC*NC] b1m[NC] W2m[NC*NC] b2m[NC] W1p b1p W2p b2p   (row-major [out][in])
};

// A CTA of 128 threads owns 32 consecutive bins of one image: lane = bin, warp q computes output channels [q*NC/4, (q+1)*NC/4) of every
// layer, so a weight row is a warp-wide broadcast and the layer inputs (staged in shared memory as [channel][bin]) are read
// conflict free into registers once per layer.  The four NC x NC layers are 4 * NC/4 * NC FMAs per thread instead of 4 * NC * NC in a
// single thread with all channels in registers (255 registers and spills at NC = 48).
#define SM_BINS 32
template <int NC>
__device__ __forceinline__ void mlp_layer_q(const float* __restrict__ W, const float* __restrict__ b, const float* __restrict__ xin, int bin, int q,
                                            float (&y)[NC / 4]) {
    constexpr int NO = NC / 4;
    float x[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) x[k] = xin[k * SM_BINS + bin];
#pragma unroll
    for (int o = 0; o < NO; ++o) {
        const int n = q * NO + o;
        float acc = b[n];
        const float4* w4 = reinterpret_cast<const float4*>(W + n * NC);
#pragma unroll
        for (int k = 0; k < NC / 4; ++k) {
            const float4 w = w4[k];
            acc += w.x * x[4 * k];
            acc += w.y * x[4 * k + 1];
            acc += w.z * x[4 * k + 2];
            acc += w.w * x[4 * k + 3];
        }
        y[o] = acc;
    }
}

template <int NC>
__global__ void __launch_bounds__(128) k_spec_mlp(SpecMlpParams q) {
    static_assert(NC % 16 == 0 || NC == 12 || NC == 24, "NC/4 outputs per warp, rows read as float4");
    constexpr int NO = NC / 4;
    FDN_DYN_SMEM(smem);
    float* sw = reinterpret_cast<float*>(smem);                 // 4 * (NC*NC + NC) weights
    float* xm = sw + 4 * (NC * NC + NC);                        // [NC][32] magnitudes, later the hidden layer
    float* xp = xm + NC * SM_BINS;                              // [NC][32] phases
    float* hd = xp + NC * SM_BINS;                              // [NC][32] hidden
    for (int i = threadIdx.x; i < 4 * (NC * NC + NC); i += 128) sw[i] = q.w[i];
    const float* W1m = sw;
    const float* b1m = W1m + NC * NC;
    const float* W2m = b1m + NC;
    const float* b2m = W2m + NC * NC;
    const float* W1p = b2m + NC;
    const float* b1p = W1p + NC * NC;
    const float* W2p = b1p + NC;
    const float* b2p = W2p + NC * NC;
    const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
    const long long bin = (long long)blockIdx.x * SM_BINS + lane;
    const bool ok = bin < q.nbins;
    float2* z = q.spec + (size_t)blockIdx.y * NC * q.plane_stride + (ok ? bin : 0);
#pragma unroll
    for (int o = 0; o < NO; ++o) {
        const int c = wq * NO + o;
        const float2 v = ok ? z[(size_t)c * q.plane_stride] : make_float2(1.f, 0.f);
        xm[c * SM_BINS + lane] = sqrtf(v.x * v.x + v.y * v.y);
        xp[c * SM_BINS + lane] = atan2f(v.y, v.x);
    }
    __syncthreads();
    float t[NO], om[NO];
    mlp_layer_q<NC>(W1m, b1m, xm, lane, wq, t);
#pragma unroll
    for (int o = 0; o < NO; ++o) hd[(wq * NO + o) * SM_BINS + lane] = fdn_lrelu(t[o]);
    __syncthreads();
    mlp_layer_q<NC>(W2m, b2m, hd, lane, wq, om);
    mlp_layer_q<NC>(W1p, b1p, xp, lane, wq, t);
    __syncthreads();                                            // every warp is done reading hd
#pragma unroll
    for (int o = 0; o < NO; ++o) hd[(wq * NO + o) * SM_BINS + lane] = fdn_lrelu(t[o]);
    __syncthreads();
    mlp_layer_q<NC>(W2p, b2p, hd, lane, wq, t);
    if (ok) {
#pragma unroll
        for (int o = 0; o < NO; ++o) {
            float sn, cs;
            sincosf(t[o], &sn, &cs);
            z[(size_t)(wq * NO + o) * q.plane_stride] = make_float2(om[o] * cs, om[o] * sn);
        }
    }
}

template <class K>
static int set_smem(K kern, size_t bytes) {
    if (bytes > 227 * 1024) {
        fdn_set_error("FFT tile does not fit in shared memory");
        return -1;
    }
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) {
            fdn_set_error(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
            return (int)e;
        }
    }
    return 0;
}

#include "fft_fast.cuh"

// ---------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------
static int rows_per_cta_for(int W) { return max(1, min(16, 8192 / W)); }              // ~4096 packed complex points per CTA
static size_t rows_smem(int W, int rpc) { return ((size_t)2 * rpc * (W / 2 + 2) + W / 2) * sizeof(float2); }
static int cols_per_cta_for(int H) {
    int tc = 8;
    while (tc > 1 && (size_t)2 * H * tc * sizeof(float2) > 160 * 1024) tc >>= 1;
    return tc;
}

// Creates (and caches) the twiddle tables for lengths H and W.  Call once per shape before CUDA-graph capture.
FDN_API int fdn_fft_prepare(int H, int W) {
    FftPlanDev p;
    FDN_REQUIRE(H >= 1 && W >= 2, "bad FFT size");
    FDN_REQUIRE(fdn_fft_get_plan(H, &p) == 0, "plan creation failed");
    FDN_REQUIRE(fdn_fft_get_plan(W, &p) == 0, "plan creation failed");
    if (W % 2 == 0) FDN_REQUIRE(fdn_fft_get_plan(W / 2, &p) == 0, "plan creation failed");
    return 0;
}

// x [planes][H][W] real -> spec [planes][H][W/2+1] complex (interleaved re,im).  Rows pass only.
FDN_API int fdn_fft_rows_r2c(const float* x, float* spec, int planes, int H, int W, cudaStream_t st) {
    FDN_REQUIRE(x && spec && planes > 0 && H > 0 && W >= 2, "bad arguments");
    FDN_REQUIRE(W % 2 == 0, "the real-packed row transform needs an even width");
    FDN_REQUIRE((reinterpret_cast<uintptr_t>(x) & 7) == 0, "x must be 8-byte aligned");
    FftPlanDev PM, PW;
    FDN_REQUIRE(fdn_fft_get_plan(W / 2, &PM, st) == 0 && fdn_fft_get_plan(W, &PW, st) == 0, "plan creation failed");
    int nrows = planes * H;
    if (fft_fast_enabled()) {
        int frc = fft_fast_rows_r2c(x, reinterpret_cast<float2*>(spec), W / 2, PM.tw, PW.tw, nrows, st);
        if (frc != FFT_FAST_NONE) return frc;
    }
    int rpc = rows_per_cta_for(W);
    size_t smem = rows_smem(W, rpc);
    int rc = set_smem(k_rows_r2c, smem);
    if (rc) return rc;
    FDN_LAUNCH(k_rows_r2c, dim3(fdn_cdiv(nrows, rpc)), dim3(256), smem, st, x, reinterpret_cast<float2*>(spec), PM, PW.tw, nrows, rpc);
    return fdn_check_launch("k_rows_r2c");
}

// spec [planes][H][W/2+1] -> y [planes][H][W];  y = img_scale[b] * (irfft_rows(spec)/(norm_hw) + res_coef*res)
FDN_API int fdn_fft_rows_c2r(const float* spec, float* y, int planes, int H, int W, float inv_norm, const float* res,
                             float res_coef, const float* img_scale, int planes_per_image, cudaStream_t st) {
    FDN_REQUIRE(spec && y && planes > 0 && H > 0 && W >= 2, "bad arguments");
    FDN_REQUIRE(W % 2 == 0, "the real-packed row transform needs an even width");
    FDN_REQUIRE((reinterpret_cast<uintptr_t>(y) & 7) == 0 && (!res || (reinterpret_cast<uintptr_t>(res) & 7) == 0), "y/res must be 8-byte aligned");
    FftPlanDev PM, PW;
    FDN_REQUIRE(fdn_fft_get_plan(W / 2, &PM, st) == 0 && fdn_fft_get_plan(W, &PW, st) == 0, "plan creation failed");
    RowsC2RParams q;
    q.in = reinterpret_cast<const float2*>(spec);
    q.out = y;
    q.res = res;
    q.img_scale = img_scale;
    q.res_coef = res_coef;
    q.norm = inv_norm;
    q.nrows = planes * H;
    q.rows_per_cta = rows_per_cta_for(W);
    q.rows_per_image = max(1, planes_per_image) * H;
    if (fft_fast_enabled()) {
        int frc = fft_fast_rows_c2r(q, W / 2, PM.tw, PW.tw, st);
        if (frc != FFT_FAST_NONE) return frc;
    }
    size_t smem = rows_smem(W, q.rows_per_cta);
    int rc = set_smem(k_rows_c2r, smem);
    if (rc) return rc;
    FDN_LAUNCH(k_rows_c2r, dim3(fdn_cdiv(q.nrows, q.rows_per_cta)), dim3(256), smem, st, q, PM, PW.tw);
    return fdn_check_launch("k_rows_c2r");
}

// Column pass over a spectrum.  mode: 0 forward, 1 inverse (unscaled), 2 forward + FCAFFN modulation + inverse,
// 3 forward -> angle(replace_denormals(.)) real map, 4 forward -> abs real map.
// in/out are addressed as  base + plane*ps + y*rs + x  (complex elements for spectra, floats for real maps).
FDN_API int fdn_fft_cols(const float* in, long long in_ps, int in_rs, float* out, long long out_ps, int out_rs, int planes,
                         int H, int ncols, int W_real, int mode, int C, const float* amp, const float* pha,
                         const float* w_xa, const float* w_xp, cudaStream_t st) {
    FDN_REQUIRE(in && out && planes > 0 && H > 0 && ncols > 0, "bad arguments");
    FDN_REQUIRE(mode >= 0 && mode <= 4, "bad mode");
    if (mode == COLS_FWD_MOD_INV) FDN_REQUIRE(C > 0 && amp && pha && w_xa && w_xp && planes % C == 0, "modulation needs maps and weights");
    FftPlanDev P;
    FDN_REQUIRE(fdn_fft_get_plan(H, &P, st) == 0, "plan creation failed");
    ColsParams q;
    q.in = reinterpret_cast<const float2*>(in);
    q.out = reinterpret_cast<float2*>(out);
    q.out_real = out;
    q.in_ps = in_ps;
    q.out_ps = out_ps;
    q.in_rs = in_rs;
    q.out_rs = out_rs;
    q.ncols = ncols;
    q.W = (mode == COLS_INV) ? 0 : W_real;
    q.tc = cols_per_cta_for(H);
    q.mode = mode;
    q.C = C;
    q.amp = amp;
    q.pha = pha;
    q.w_xa = w_xa;
    q.w_xp = w_xp;
    if (fft_fast_enabled()) {
        int frc = fft_fast_cols(q, H, P.tw, planes, st);
        if (frc != FFT_FAST_NONE) return frc;
    }
    size_t smem = ((size_t)2 * H * q.tc + H) * sizeof(float2);
    int rc = set_smem(k_cols, smem);
    if (rc) return rc;
    FDN_LAUNCH(k_cols, dim3(fdn_cdiv(ncols, q.tc), planes), dim3(256), smem, st, q, P);
    return fdn_check_launch("k_cols");
}

// MAR spectral MLPs in place on spec [B][NC][nbins] complex (channel stride plane_stride complex elements).
// w = W1m b1m W2m b2m W1p b1p W2p b2p, each W row-major [out][in].
FDN_API int fdn_spec_mlp(float* spec, long long plane_stride, long long nbins, int B, int NC, const float* w, cudaStream_t st) {
    FDN_REQUIRE(spec && w && B > 0 && nbins > 0, "bad arguments");
    SpecMlpParams q;
    q.spec = reinterpret_cast<float2*>(spec);
    q.plane_stride = plane_stride;
    q.nbins = nbins;
    q.B = B;
    q.w = w;
    dim3 grid(fdn_cdiv(nbins, SM_BINS), B), block(128);
#define FDN_SPEC_MLP_CASE(NCV)                                                                         \
    {                                                                                                  \
        auto k = k_spec_mlp<NCV>;                                                                      \
        const size_t smem = (size_t)(4 * (NCV * NCV + NCV) + 3 * NCV * SM_BINS) * sizeof(float);       \
        int rc = set_smem(k, smem);                                                                    \
        if (rc) return rc;                                                                             \
        FDN_LAUNCH(k, grid, block, smem, st, q);                                                       \
    }
    if (NC == 12) FDN_SPEC_MLP_CASE(12)
    else if (NC == 24) FDN_SPEC_MLP_CASE(24)
    else if (NC == 48) FDN_SPEC_MLP_CASE(48)
    else FDN_REQUIRE(false, "unsupported channel count (12/24/48)");
#undef FDN_SPEC_MLP_CASE
    return fdn_check_launch("k_spec_mlp");
}
