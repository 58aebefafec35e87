// 1x1 convolutions of the FDformer blocks on the 5th-generation tensor cores (tcgen05, sm_100a).
//
//   Y[b, n, p] = epilogue( sum_k W[n, k] * prologue(X)[b, k, p] )       p = pixel, M = 128 pixels per tile
//
// These are the only dense contractions of the network (83 % of its FLOPs):
// FDSA to_hidden / project_out, FDFFN project_in / project_out, FCAFFN project_in / project_out, Fuse conv / conv2
// (FDN_arch.py:388-389, 451-452, 562, 566, 685-686).
//
// Persistent, warp-specialised kernel (20 warps per CTA, one CTA per SM, static round-robin over pixel tiles):
//   warps 17-19 loaders: stream raw [32 channels][128 pixels] blocks of X (and of the per-pixel side operand) from HBM
//               into a shared-memory ring, one TMA bulk copy (cp.async.bulk, 512 contiguous bytes) per channel row with
//               completion counted on the slot's mbarrier - several K blocks ahead of the consumers, which keeps enough
//               bytes in flight to cover HBM latency.  Issuing a bulk copy costs ~60 cycles of the issuing thread, so the
//               rows are interleaved over the lanes of the loader warps.  When the weights do not fit in shared memory the last
//               loader warp streams one weight panel per K block instead.  (FDN_MMA_BULK=0 selects 16-byte cp.async.)
//   warps 0-7   producers: two threads per pixel; take the LayerNorm statistics from shared memory (two-pass mean /
//               biased variance like the reference; for the gate prologue with E <= 40 also the three group statistics), apply
//               the per-pixel prologue, split each value into tf32 hi + lo (lo is left untruncated: the tensor core reads only its
//               tf32 bits) and store it into the canonical K-major SWIZZLE_128B operand stage.  No per-element validity select:
//               the raw ring starts as zeros and gamma / beta / weights are zero on padding rows
//   warp  16    MMA issuer: one elected lane issues tcgen05.mma (cta_group::1, kind::tf32, 128 x N x 8) with the weight
//               panel that is resident in shared memory (packed on the host into the UMMA image) and owns TMEM
//   warps 8-15  epilogue: thread = TMEM lane = pixel (two warps per lane quadrant split the columns); residual prefetched into registers, tcgen05.ld, IEEE sum of the
//               accumulators, bias / FiLM / residual, 128-byte coalesced NCHW stores
// Barriers: raw_full[r] (tx bytes) -> producers -> raw_empty[r]; a_full[s] (one arrival per producer warp) -> MMA; tcgen05.commit ->
// a_empty[s]; commit -> acc_full -> epilogue -> acc_empty (one arrival per epilogue warp) -> MMA of the next tile.
//
// Precision: passes = 3 ("3xTF32", default) splits both operands into tf32 hi + tf32 lo and accumulates
// hi*hi + hi*lo + lo*hi.  The split itself is exact to 7e-8; because the tensor core truncates its fp32 accumulator after
// every instruction, the large hi*hi sum and the small corrections live in separate TMEM accumulators and long K is
// spread round-robin over up to three main accumulators, which the epilogue adds with IEEE fp32 additions.  For Nc <= 128
// A_hi x [B_hi;B_lo] is issued as one MMA of N = 2*Nc into an adjacent (main, correction) accumulator pair.
// passes = 1 is single-pass TF32 (reported separately).
#include "fdn_common.cuh"
#include <type_traits>

#ifndef FDN_EMU

#define MMA_TP 128          // pixels per tile (UMMA M)
#define MMA_KB 32           // channels per K block (one 128-byte swizzle row of tf32)
#define MMA_PROD_THREADS 256
#define MMA_EPI_WARP0 8
#define MMA_EPI_THREADS 256
#define MMA_MMA_WARP 16
#define MMA_LOAD_WARP0 17     // three loader warps: a warp retires its lanes' bulk copies one after another (~60 cycles each), so the copy issue rate scales with the number of loader warps; 20 warps still get 96 registers each (a 21st caps them at 80 and spills)
#define MMA_LOAD_THREADS 96
#define MMA_THREADS 640
#define MMA_MAX_K 512        // LayerNorm gamma/beta staged in shared memory
#define MMA_MAX_RING 8
#define MMA_SLOT_BYTES (MMA_KB * MMA_TP * 4)   // 16 KB: one raw K block
// prologue 2 packs its ring slot: 30 gate rows, 10 v_value rows, 6 statistics rows (23 KB instead of 35 KB -> deeper ring)
#define MMA_P2_V_OFF (3 * MMA_EB * MMA_TP * 4)
#define MMA_P2_ST_OFF (4 * MMA_EB * MMA_TP * 4)
#define MMA_P2_SLOT ((4 * MMA_EB + 6) * MMA_TP * 4)
#define MMA_EB 10            // prologue 2: channels of each LayerNorm group per K block (3 x 10 = 30 of the 32 rows)

struct PwMmaParams {
    const float* src0;      // [B][C0][HW]
    const float* src1;      // [B][C1][HW] or null (channel concat)
    int C0, C1;
    int K;                  // C0 + C1
    int Kpad;               // K rounded up to 8
    int N;                  // real output channels
    int Nc;                 // padded output channels per chunk (multiple of 16, <= 256)
    int HW, B;
    const float* bpack;     // [nchunks][kblocks][2 (hi,lo)][Nc][32] floats, pre-swizzled
    int prologue;           // 0 none, 1 LayerNorm(K), 2 FDSA gate (3 LN groups, stats precomputed) x v_value, 3 FCAFFN mix
    const float* ln_w;      // [K] (prologue 1,3) or [3][E] (prologue 2)
    const float* ln_b;
    const float* aux;       // prologue 2: v_value, element (b,e,p) at aux[b*aux_bs + e*HW + p]; prologue 3: x1 [B][K][HW]
    long long aux_bs;
    const float* stats;     // prologue 2: [B][3][2][HW] (mean, 1/sqrt(var+eps)) per LayerNorm group
    const float* bias;      // [N] or null
    const float* film_mul;  // [B][N][HW] or null
    const float* film_add;
    const float* res;       // [B][N][HW] or null
    float res_coef;
    float* out;             // [B][N][HW]
    int passes;             // 3 = 3xTF32, 1 = TF32
    uint32_t idesc;
    int b_resident;         // whole weight chunk kept in shared memory (else streamed into the operand stage)
    int nstage;             // operand stages (1 or 2)
    int ring;               // raw ring slots
    int nmain;              // main accumulators (K blocks round-robin)
    int merge;              // 1: A_hi x [B_hi;B_lo] is one MMA of N = 2*Nc into a (main, correction) accumulator pair
    int set_cols;           // TMEM columns of one accumulator set
    uint32_t idesc2;        // instruction descriptor with N = 2*Nc (merge)
    int tmem_cols;          // power of two >= nbuf * (nmain + ncorr) * Nc
    int ncorr;              // 1 when the corrections have their own accumulator
    int nbuf;               // accumulator sets (2 = the epilogue of tile t overlaps the MMAs of tile t+1)
    int bulk;               // 1: one cp.async.bulk per 512-byte channel row (few rows per tile), 0: 16-byte cp.async by 128 threads
    int E;                  // prologue 2: channels per LayerNorm group (K is then laid out in blocks of 3 x 10 channels)
    int Kreal;              // number of real input channels (= K except for the grouped layout of prologue 2)
    unsigned long long* dbg;  // optional [8] counters: cycles each role spent waiting on each barrier (fdn_pw_mma_set_debug)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Round to tf32 (10 explicit mantissa bits), nearest with ties away from zero - what cvt.rna.tf32.f32 computes for finite inputs.
// ptxas expands that instruction into four (add, Inf/NaN test, select, mask); activations are finite, so two suffice.
__device__ __forceinline__ float to_tf32(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// Wait for the phase with the given parity.  A failed probe backs off with nanosleep so that polling warps do not steal
// issue slots from the role that is the bottleneck (measured with fdn_pw_mma_set_debug: idle roles were spinning 40-60 %).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.b32 %0, 1, 0, P1;\n"
        "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    while (!done) {
        asm volatile("nanosleep.u32 40;" ::: "memory");
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
            "selp.b32 %0, 1, 0, P1;\n"
            "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}
// mbar_wait that charges the waiting time to a debug counter (only one thread per role records)
#ifndef FDN_MMA_PROFILE
#define FDN_MMA_PROFILE 0      // 1: per-role wait counters (fdn_pw_mma_set_debug); costs registers, so off in the product build
#endif
__device__ __forceinline__ void mbar_wait_t(uint64_t* bar, uint32_t parity, unsigned long long* acc, bool rec) {
    if (!FDN_MMA_PROFILE || !rec) { mbar_wait(bar, parity); return; }
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    *acc += (unsigned long long)(clock64() - t0);
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
// the mbarrier receives this thread's arrival once all of its prior cp.async copies have landed
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major, SWIZZLE_128B, 8-row groups 1024 bytes apart, descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);          // start address
    d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset (leading byte offset unused: one atom along K)
    d |= (uint64_t)1 << 46;                          // version
    d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
    return d;
}
// byte offset of the 16-byte group `chunk` of row `row` inside a [rows][32 floats] K-major SWIZZLE_128B panel
__device__ __forceinline__ uint32_t sw128_off(int row, int chunk) {
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4));
}
// Position in a ring of n slots plus the parity of the current round: the roles walk their rings with an increment, a compare and
// an xor instead of it % n and it / n on run-time n (each ~20 instructions through MUFU.RCP; the single-warp MMA loop spent most of
// its ~270 instructions per K block on them and capped the kernel at ~2000 cycles per K block in round 1).
struct RingPos {
    int idx;
    uint32_t par;
    __device__ __forceinline__ void advance(int n) {
        if (++idx == n) { idx = 0; par ^= 1u; }
    }
};
__device__ __forceinline__ void prod_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// issue only; call tmem_ld_wait() before using the registers
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t r[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ const float* src_row(const PwMmaParams& q, int b, int k) {
    if (k < q.C0) return q.src0 + ((size_t)b * q.C0 + k) * q.HW;
    return q.src1 + ((size_t)b * q.C1 + (k - q.C0)) * q.HW;
}

// Epilogue of one 16-column group of one pixel (thread): planes are HW floats apart.  All loads are issued before the first
// store (out may alias res for in-place residuals); null pointers switch a stage off in the generic instantiation, offsets stay 32-bit (16 planes * HW * 4 B < 2^31 for any image we accept).
template <bool FULL, bool RES, bool FILM, bool BIAS>
__device__ __forceinline__ void epi_group(float (&acc)[16], float* op, const float* rp, const float* fm, const float* fa,
                                          const float* bias, const float (&rpre)[16], bool pre, float res_coef, uint32_t HW, int nvalid,
                                          const float* fmpre = nullptr, const float* fapre = nullptr) {
    if (BIAS && bias != nullptr) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (FULL || j < nvalid) acc[j] += bias[j];
    }
    if (FILM && fmpre != nullptr) {                    // FiLM maps fetched before the accumulator wait (registers)
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (FULL || j < nvalid) acc[j] = acc[j] * fmpre[j] + fapre[j];
    } else if (FILM && fm != nullptr) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (FULL || j < nvalid) acc[j] = acc[j] * fm[(size_t)HW * j] + fa[(size_t)HW * j];
    }
    if (RES) {
        if (pre) {
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] += res_coef * rpre[j];
        } else if (rp != nullptr) {
#pragma unroll
            for (int h = 0; h < 16; h += 8) {         // two batches of eight loads: enough memory parallelism, half the registers
                float r[8];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (FULL || h + j < nvalid) r[j] = rp[(size_t)HW * (h + j)];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (FULL || h + j < nvalid) acc[h + j] += res_coef * r[j];
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j)
        if (FULL || j < nvalid) op[(size_t)HW * j] = acc[j];
}

template <int PRO, int PASSES>
__global__ void __launch_bounds__(MMA_THREADS, 1) k_pw_mma(PwMmaParams q) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // align by pointer arithmetic on the shared array (an integer round trip would turn every access into a generic LD/ST)
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr uint32_t a_bytes = MMA_TP * 128;
    constexpr bool has_aux = PRO >= 2;
    constexpr uint32_t slot_bytes = PRO == 2 ? MMA_P2_SLOT : MMA_SLOT_BYTES * (has_aux ? 2 : 1);
    const uint32_t b_bytes = (uint32_t)q.Nc * 128;
    const int nkb = (q.Kpad + MMA_KB - 1) / MMA_KB;
    const uint32_t bres_bytes = q.b_resident ? (uint32_t)nkb * 2 * b_bytes : 0;
    const uint32_t stage_bytes = 2 * a_bytes + (q.b_resident ? 0 : 2 * b_bytes);
    unsigned char* s_bres = smem;
    unsigned char* s_stage = s_bres + bres_bytes;
    unsigned char* s_raw = s_stage + q.nstage * stage_bytes;
    float* s_part = reinterpret_cast<float*>(s_raw + q.ring * slot_bytes);       // [2][128] partial sums
    float* s_gam = s_part + 6 * MMA_TP;                                          // [MMA_MAX_K] LayerNorm gamma   (s_part: [2 halves][3 groups][128])
    float* s_bet = s_gam + MMA_MAX_K;                                            // [MMA_MAX_K] LayerNorm beta
    uint64_t* raw_full = reinterpret_cast<uint64_t*>(s_bet + MMA_MAX_K);
    uint64_t* raw_empty = raw_full + MMA_MAX_RING;
    uint64_t* a_full = raw_empty + MMA_MAX_RING;
    uint64_t* a_empty = a_full + 4;          // [4] operand stages
    uint64_t* acc_full = a_empty + 4;        // [2]
    uint64_t* acc_empty = acc_full + 2;      // [2]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int chunk = blockIdx.y;
    const int HW = q.HW;
    const int tiles_per_img = (HW + MMA_TP - 1) / MMA_TP;
    const int ntiles = tiles_per_img * q.B;
    const float* bsrc = q.bpack + (size_t)chunk * nkb * 2 * q.Nc * 32;
    constexpr int panels = PASSES == 3 ? 2 : 1;
    const int gsize = PRO == 2 ? q.E : q.K;

    if (tid == 0) {
        for (int i = 0; i < q.ring; ++i) { mbar_init(&raw_full[i], q.bulk ? 1 : (q.b_resident ? MMA_LOAD_THREADS : MMA_LOAD_THREADS - 32)); mbar_init(&raw_empty[i], MMA_PROD_THREADS / 32); }
        for (int i = 0; i < q.nstage; ++i) { mbar_init(&a_full[i], MMA_PROD_THREADS / 32 + (q.b_resident ? 0 : 1)); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], MMA_EPI_THREADS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"((uint32_t)q.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (PRO != 0)
        for (int i = tid; i < MMA_MAX_K; i += MMA_THREADS) {
            // gamma / beta in the order of the operand rows, zero on padding rows.  Prologue 2: row kk = g*10 + el of K block kb is
            // channel g*E + kb*10 + el (grouped layout), so the producers read them with the same 128-bit loads as the plain order
            int src = i;
            bool ok = i < q.Kreal;
            if (PRO == 2) {
                const int kb = i >> 5, kk = i & 31, g = kk / MMA_EB, e = kb * MMA_EB + (kk - g * MMA_EB);
                ok = g < 3 && e < q.E;
                src = g * q.E + e;
            }
            s_gam[i] = ok ? q.ln_w[src] : 0.f;
            s_bet[i] = ok ? q.ln_b[src] : 0.f;
        }
    // The raw ring starts out as zeros: rows or pixels a tile does not load then hold zeros or stale (finite) activations, never the
    // bit patterns a previous kernel left behind, and their contribution is removed by the zero gamma / beta and zero weight padding -
    // so the producers need no per-element validity select.
    for (int i = tid; i < (int)(q.ring * slot_bytes / 16); i += MMA_THREADS) reinterpret_cast<float4*>(s_raw)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (q.b_resident) {      // the whole weight chunk stays in shared memory for the lifetime of the CTA
        const float4* src = reinterpret_cast<const float4*>(bsrc);
        float4* dst = reinterpret_cast<float4*>(s_bres);
        const int per_kb = 2 * q.Nc * 8, used = panels * q.Nc * 8;
        for (int i = tid; i < nkb * per_kb; i += MMA_THREADS)
            if ((i % per_kb) < used) dst[i] = src[i];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *s_tmem;
    const int nmain = q.nmain;
    unsigned long long w0 = 0, w1 = 0;            // debug: cycles spent waiting (recorded by one thread per role)
    const bool rec = FDN_MMA_PROFILE && q.dbg != nullptr && (tid == 0 || tid == MMA_EPI_WARP0 * 32 || tid == MMA_MMA_WARP * 32 || tid == MMA_LOAD_WARP0 * 32);
    const long long t_start = FDN_MMA_PROFILE ? clock64() : 0;

    if (warp >= MMA_LOAD_WARP0) {
        // =============================================== loaders ====================================================
        // 128 threads stream 16-byte pieces with cp.async (L2 -> shared memory, no registers); each thread's arrival on
        // raw_full[r] is deferred by the hardware until its copies have landed
        const int lt = tid - MMA_LOAD_WARP0 * 32;
        uint32_t lit = 0;
        if (!q.b_resident && warp == MMA_LOAD_WARP0 + MMA_LOAD_THREADS / 32 - 1) {
            // weight panels that do not fit in shared memory: one contiguous TMA bulk copy per K block straight into the
            // operand stage (the packed image is contiguous in global memory); completion counts on a_full[s]
            if (lane == 0) {
                RingPos sp{0, 0};
                const uint32_t bytes = (uint32_t)panels * b_bytes;
                for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
                    for (int kb = 0; kb < nkb; ++kb) {
                        const int s = sp.idx;
                        mbar_wait(&a_empty[s], sp.par ^ 1u);        // first round: the preceding phase of a fresh barrier counts as complete
                        mbar_expect_tx(&a_full[s], bytes);
                        bulk_g2s(s_stage + s * stage_bytes + 2 * a_bytes, bsrc + (size_t)kb * 2 * q.Nc * 32, bytes, &a_full[s]);
                        sp.advance(q.nstage);
                    }
            }
        } else if (q.bulk) {
            // one TMA bulk copy (512 contiguous bytes) per channel row.  Issuing a bulk copy costs ~60 cycles of one thread's
            // uniform datapath, so the rows of every K block are interleaved over all loader warps that do not stream weights.
            // activation loader warps: two (a compile-time constant) except for the gate prologue with resident weights, whose 46 rows per
            // K block use all three; with streamed weights the last loader warp carries the weight panels
            const int nlw = PRO == 2 ? (q.b_resident ? MMA_LOAD_THREADS / 32 : MMA_LOAD_THREADS / 32 - 1) : 2;
            const int lw = warp - MMA_LOAD_WARP0;
            RingPos rp{0, 0};
            if (lw < nlw)
                for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                    const int b = tile / tiles_per_img, p0 = (tile - b * tiles_per_img) * MMA_TP;
                    const uint32_t len = (uint32_t)min(MMA_TP, HW - p0) * 4;
                    for (int kb = 0; kb < nkb; ++kb, rp.advance(q.ring)) {
                        const int r = rp.idx;
                        mbar_wait_t(&raw_empty[r], rp.par ^ 1u, &w0, rec);
                        unsigned char* slot = s_raw + (size_t)r * slot_bytes;
                        if (PRO == 2) {
                            // grouped layout: row kk = g*10 + el holds channel g*E + kb*10 + el; the 10 v_value rows are loaded once;
                            // virtual rows 0..29 x, 30..39 v_value, 40..45 the LayerNorm statistics (first K block only)
                            const int ne = min(MMA_EB, q.E - kb * MMA_EB);
                            if (lw == 0 && lane == 0) mbar_expect_tx(&raw_full[r], (uint32_t)(4 * ne + (kb == 0 && q.stats ? 6 : 0)) * len);
                            for (int vi = lane * nlw + lw; vi < 4 * MMA_EB + 6; vi += 32 * nlw) {      // virtual rows of this lane
                                const int g = vi / MMA_EB, el = vi - g * MMA_EB;
                                if (g < 3) {
                                    if (el < ne)
                                        bulk_g2s(slot + vi * (MMA_TP * 4), q.src0 + ((size_t)b * q.C0 + g * q.E + kb * MMA_EB + el) * HW + p0, len, &raw_full[r]);
                                } else if (g == 3) {
                                    if (el < ne)
                                        bulk_g2s(slot + MMA_P2_V_OFF + el * (MMA_TP * 4), q.aux + (size_t)b * q.aux_bs + (size_t)(kb * MMA_EB + el) * HW + p0, len, &raw_full[r]);
                                } else if (kb == 0 && q.stats) {
                                    const int sr = vi - 4 * MMA_EB;
                                    bulk_g2s(slot + MMA_P2_ST_OFF + sr * (MMA_TP * 4), q.stats + ((size_t)b * 6 + sr) * HW + p0, len, &raw_full[r]);
                                }
                            }
                        } else {
                            const int rows = min(MMA_KB, q.K - kb * MMA_KB);
                            if (lw == 0 && lane == 0) mbar_expect_tx(&raw_full[r], (uint32_t)rows * len * (has_aux ? 2 : 1));
                            for (int vi = lane * nlw + lw; vi < (has_aux ? 2 : 1) * rows; vi += 32 * nlw) {
                                if (vi < rows) {
                                    bulk_g2s(slot + vi * (MMA_TP * 4), src_row(q, b, kb * MMA_KB + vi) + p0, len, &raw_full[r]);
                                } else {
                                    const int ar = vi - rows;
                                    bulk_g2s(slot + MMA_SLOT_BYTES + ar * (MMA_TP * 4), q.aux + (size_t)b * q.aux_bs + (size_t)(kb * MMA_KB + ar) * HW + p0, len, &raw_full[r]);
                                }
                            }
                        }
                    }
                }
        } else {
        const int lthreads = q.b_resident ? MMA_LOAD_THREADS : MMA_LOAD_THREADS - 32;      // the last loader warp may stream weights
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int b = tile / tiles_per_img, p0 = (tile - b * tiles_per_img) * MMA_TP;
            const int vch = min(MMA_TP, HW - p0) >> 2;                        // valid 16-byte pieces per channel row
            for (int kb = 0; kb < nkb; ++kb, ++lit) {
                const int r = lit % q.ring;
                if (lit >= (uint32_t)q.ring) mbar_wait(&raw_empty[r], ((lit / q.ring) - 1) & 1);      // (FDN_MMA_BULK=0 dev path: not tuned)
                unsigned char* slot = s_raw + (size_t)r * slot_bytes;
                if (PRO == 2) {
                    const int ne = min(MMA_EB, q.E - kb * MMA_EB);
                    for (int i = lt; i < 4 * MMA_EB * 32; i += lthreads) {      // rows 0..29: x (3 groups x 10), rows 30..39: v_value
                        const int row = i >> 5, ch = i & 31;
                        const int g = row / MMA_EB, el = row - g * MMA_EB;
                        if (ch < vch && el < ne) {
                            if (g < 3)
                                cp_async16(slot + row * (MMA_TP * 4) + ch * 16,
                                           q.src0 + ((size_t)b * q.C0 + g * q.E + kb * MMA_EB + el) * HW + p0 + ch * 4);
                            else
                                cp_async16(slot + MMA_P2_V_OFF + el * (MMA_TP * 4) + ch * 16,
                                           q.aux + (size_t)b * q.aux_bs + (size_t)(kb * MMA_EB + el) * HW + p0 + ch * 4);
                        }
                    }
                } else {
                const int rows = min(MMA_KB, q.K - kb * MMA_KB);
                for (int i = lt; i < rows * 32; i += lthreads) {
                    const int row = i >> 5, ch = i & 31;
                    if (ch < vch) {
                        const int k = kb * MMA_KB + row;
                        cp_async16(slot + row * (MMA_TP * 4) + ch * 16, src_row(q, b, k) + p0 + ch * 4);
                        if (has_aux)
                            cp_async16(slot + MMA_SLOT_BYTES + row * (MMA_TP * 4) + ch * 16, q.aux + (size_t)b * q.aux_bs + (size_t)k * HW + p0 + ch * 4);
                    }
                }
                }
                if (PRO == 2 && kb == 0)
                    for (int i = lt; i < 6 * 32; i += lthreads) {
                        const int row = i >> 5, ch = i & 31;
                        if (ch < vch)
                            cp_async16(slot + MMA_P2_ST_OFF + row * (MMA_TP * 4) + ch * 16, q.stats + ((size_t)b * 6 + row) * HW + p0 + ch * 4);
                    }
                cp_async_arrive(&raw_full[r]);
            }
        }
        }
    } else if (warp < MMA_EPI_WARP0) {
        // =============================================== producers =================================================
        const int pix = tid & (MMA_TP - 1), half = tid >> 7;
        RingPos rr{0, 0}, sr{0, 0};       // raw ring slot / operand stage of the next K block (same sequence as the loader and the MMA warp)
        // raw slot of K block kb of the current tile (the statistics passes look ahead; ring >= nkb whenever they run)
        auto ahead = [&](int kb, uint32_t& par) {
            int i = rr.idx + kb;
            par = rr.par;
            if (i >= q.ring) { i -= q.ring; par ^= 1u; }
            return i;
        };
        // byte offsets of this thread's four 16-byte groups inside the swizzled operand panel
        uint32_t soff[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) soff[i] = sw128_off(pix, half + 2 * i);
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int b = tile / tiles_per_img, p0 = (tile - b * tiles_per_img) * MMA_TP;
            float mu = 0.f, rs = 1.f;
            float gmu0 = 0.f, gmu1 = 0.f, gmu2 = 0.f, grs0 = 1.f, grs1 = 1.f, grs2 = 1.f;
            if (PRO == 1 || PRO == 3) {
                // LayerNorm statistics over all K channels of this pixel, from the raw ring (all K blocks of the tile)
                float s = 0.f;
                for (int kb = 0; kb < nkb; ++kb) {
                    uint32_t par;
                    const int ri = ahead(kb, par);
                    mbar_wait_t(&raw_full[ri], par, &w0, rec);
                    const float* raw = reinterpret_cast<const float*>(s_raw + (size_t)ri * slot_bytes) + pix;
                    const int kmax = min(MMA_KB, q.K - kb * MMA_KB);
#pragma unroll
                    for (int kk = 0; kk < 16; ++kk) {
                        const int k2 = 2 * kk + half;
                        if (k2 < kmax) s += raw[k2 * MMA_TP];
                    }
                }
                prod_sync();                       // the previous tile's readers of s_part are done
                s_part[half * MMA_TP + pix] = s;
                prod_sync();
                mu = (s_part[pix] + s_part[MMA_TP + pix]) / (float)q.K;
                float v = 0.f;
                for (int kb = 0; kb < nkb; ++kb) {
                    uint32_t par;
                    const float* raw = reinterpret_cast<const float*>(s_raw + (size_t)ahead(kb, par) * slot_bytes) + pix;
                    const int kmax = min(MMA_KB, q.K - kb * MMA_KB);
#pragma unroll
                    for (int kk = 0; kk < 16; ++kk) {
                        const int k2 = 2 * kk + half;
                        if (k2 < kmax) { const float d = raw[k2 * MMA_TP] - mu; v += d * d; }
                    }
                }
                prod_sync();
                s_part[half * MMA_TP + pix] = v;
                prod_sync();
                rs = 1.0f / sqrtf((s_part[pix] + s_part[MMA_TP + pix]) / (float)q.K + 1e-5f);
            }
            if (PRO == 2 && q.stats == nullptr) {
                // LayerNorm statistics of the three groups from the raw ring (all K blocks of the tile are resident: the host only
                // selects this mode when they fit) - two passes like k_group_stats, which this replaces for small E
                auto group_pass = [&](auto half_c, float m0, float m1, float m2, bool second, float (&acc)[3]) {
                    constexpr int HALF = decltype(half_c)::value;
                    acc[0] = acc[1] = acc[2] = 0.f;
                    for (int kb = 0; kb < nkb; ++kb) {
                        uint32_t par;
                        const int ri = ahead(kb, par);
                        if (!second) mbar_wait_t(&raw_full[ri], par, &w0, rec);
                        const float* raw = reinterpret_cast<const float*>(s_raw + (size_t)ri * slot_bytes) + pix;
                        const int ne = min(MMA_EB, q.E - kb * MMA_EB);
#pragma unroll
                        for (int kk2 = 0; kk2 < 15; ++kk2) {
                            constexpr int dummy = 0; (void)dummy;
                            const int kk = 2 * kk2 + HALF;              // rows 0..29: group kk / 10, channel kk % 10 of this block
                            const int g = kk / MMA_EB, el = kk - g * MMA_EB;
                            if (el < ne) {
                                const float x = raw[kk * MMA_TP];
                                if (!second) acc[g] += x;
                                else { const float d = x - (g == 0 ? m0 : (g == 1 ? m1 : m2)); acc[g] += d * d; }
                            }
                        }
                    }
                };
                float acc[3];
                if (half == 0) group_pass(std::integral_constant<int, 0>{}, 0.f, 0.f, 0.f, false, acc);
                else group_pass(std::integral_constant<int, 1>{}, 0.f, 0.f, 0.f, false, acc);
                prod_sync();                       // the previous tile's readers of s_part are done
#pragma unroll
                for (int g = 0; g < 3; ++g) s_part[(half * 3 + g) * MMA_TP + pix] = acc[g];
                prod_sync();
                const float invE = 1.0f / (float)q.E;
                gmu0 = (s_part[0 * MMA_TP + pix] + s_part[3 * MMA_TP + pix]) * invE;
                gmu1 = (s_part[1 * MMA_TP + pix] + s_part[4 * MMA_TP + pix]) * invE;
                gmu2 = (s_part[2 * MMA_TP + pix] + s_part[5 * MMA_TP + pix]) * invE;
                if (half == 0) group_pass(std::integral_constant<int, 0>{}, gmu0, gmu1, gmu2, true, acc);
                else group_pass(std::integral_constant<int, 1>{}, gmu0, gmu1, gmu2, true, acc);
                prod_sync();
#pragma unroll
                for (int g = 0; g < 3; ++g) s_part[(half * 3 + g) * MMA_TP + pix] = acc[g];
                prod_sync();
                grs0 = 1.0f / sqrtf((s_part[0 * MMA_TP + pix] + s_part[3 * MMA_TP + pix]) * invE + 1e-5f);
                grs1 = 1.0f / sqrtf((s_part[1 * MMA_TP + pix] + s_part[4 * MMA_TP + pix]) * invE + 1e-5f);
                grs2 = 1.0f / sqrtf((s_part[2 * MMA_TP + pix] + s_part[5 * MMA_TP + pix]) * invE + 1e-5f);
            }
            for (int kb = 0; kb < nkb; ++kb, rr.advance(q.ring), sr.advance(q.nstage)) {
                const int r = rr.idx;
                const float* raw = reinterpret_cast<const float*>(s_raw + (size_t)r * slot_bytes) + pix;
                mbar_wait_t(&raw_full[r], rr.par, &w0, rec);
                if (PRO == 2 && kb == 0 && q.stats != nullptr) {
                    const float* st = raw + MMA_P2_ST_OFF / 4;
                    gmu0 = st[0 * MMA_TP]; grs0 = st[1 * MMA_TP]; gmu1 = st[2 * MMA_TP]; grs1 = st[3 * MMA_TP];
                    gmu2 = st[4 * MMA_TP]; grs2 = st[5 * MMA_TP];
                }
                const int s = sr.idx;
                unsigned char* stage = s_stage + s * stage_bytes;
                const int nchunks_used = min(MMA_KB, q.Kpad - kb * MMA_KB) >> 2;
                mbar_wait_t(&a_empty[s], sr.par ^ 1u, &w1, rec);      // first round: passes at once
                // The conversion is specialised on this thread's half (0/1) so that every per-element index (channel, LayerNorm
                // group, v_value row, smem offsets) is a compile-time constant: ~3x fewer instructions than runtime indexing.
                auto convert = [&](auto half_c) {
                    constexpr int HALF = decltype(half_c)::value;
                    const float* gam_kb = s_gam + kb * MMA_KB;
                    const float* bet_kb = s_bet + kb * MMA_KB;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int c = HALF + 2 * i;
                        if (c < nchunks_used) {
                            float hi[4], lo[4];
                            float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f), b4 = g4;
                            if (PRO != 0) {
                                g4 = *reinterpret_cast<const float4*>(gam_kb + 4 * c);
                                b4 = *reinterpret_cast<const float4*>(bet_kb + 4 * c);
                            }
                            // two channels per packed operation (FADD2 / FMUL2 / FFMA2 with the per-pixel statistics as broadcast
                            // operands): same operation order as the scalar form, so the operand bits are unchanged
                            const float2 gp[2] = {make_float2(g4.x, g4.y), make_float2(g4.z, g4.w)};
                            const float2 bp[2] = {make_float2(b4.x, b4.y), make_float2(b4.z, b4.w)};
#pragma unroll
                            for (int jp = 0; jp < 2; ++jp) {
                                const int kk = 4 * c + 2 * jp;
                                float2 x = make_float2(raw[kk * MMA_TP], raw[(kk + 1) * MMA_TP]);
                                if (PRO == 1) {
                                    x = f2fma(f2mul_s(f2add_s(x, -mu), rs), gp[jp], bp[jp]);
                                } else if (PRO == 2) {
                                    // grouped layout: kk = g*10 + el, channel g*E + kb*10 + el, v_value row el (kk even: a pair never straddles groups)
                                    const int g = kk / MMA_EB, el = kk - g * MMA_EB;
                                    if (g < 3) {
                                        const float gm = g == 0 ? gmu0 : (g == 1 ? gmu1 : gmu2), gr = g == 0 ? grs0 : (g == 1 ? grs1 : grs2);
                                        const float2 vv = make_float2(raw[MMA_P2_V_OFF / 4 + el * MMA_TP], raw[MMA_P2_V_OFF / 4 + (el + 1) * MMA_TP]);
                                        x = f2mul(f2fma(f2mul_s(f2add_s(x, -gm), gr), gp[jp], bp[jp]), vv);
                                    } else {
                                        x = make_float2(0.f, 0.f);          // rows 30, 31 of a grouped block
                                    }
                                } else if (PRO == 3) {
                                    const float2 x1 = make_float2(raw[(MMA_SLOT_BYTES / 4) + kk * MMA_TP], raw[(MMA_SLOT_BYTES / 4) + (kk + 1) * MMA_TP]);
                                    x = f2fma(f2fma(f2mul_s(f2add_s(x, -mu), rs), gp[jp], bp[jp]), x1, x1);
                                }
                                const float2 h = make_float2(to_tf32(x.x), to_tf32(x.y));
                                const float2 l = csub(x, h);      // exact; the tensor core reads only the tf32 bits of it (truncation: 2^-21 |x|, the size of the dropped lo*lo term)
                                hi[2 * jp] = h.x; hi[2 * jp + 1] = h.y;
                                lo[2 * jp] = l.x; lo[2 * jp + 1] = l.y;
                            }
                            *reinterpret_cast<float4*>(stage + soff[i]) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                            if (PASSES == 3) *reinterpret_cast<float4*>(stage + a_bytes + soff[i]) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                        }
                    }
                };
                if (half == 0) convert(std::integral_constant<int, 0>{}); else convert(std::integral_constant<int, 1>{});
                // One arrival per warp, not per thread: 256 arrivals on one shared-memory word serialise (two barriers per K block),
                // which was a fixed ~2 k cycles per K block.  Every lane makes its operand stores visible to the async proxy, the
                // warp converges, lane 0 signals both barriers.
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&raw_empty[r]);        // the warp is done with the raw slot
                    mbar_arrive(&a_full[s]);
                }
            }
        }
    } else if (warp == MMA_MMA_WARP) {
        // =============================================== MMA issuer ================================================
        // One lane issues everything, so this loop is serial, latency-bound scalar code: shared-memory descriptors are advanced with
        // integer adds from a constant template (the 14-bit address field cannot carry: shared memory ends below 256 KB) and the
        // stage / accumulator indices are counters.
        uint32_t titer = 0;
        const uint32_t set_cols = (uint32_t)q.set_cols, Nc = (uint32_t)q.Nc;
        const uint64_t desc0 = make_desc(0);
        const uint32_t stage_u = smem_u32(s_stage) >> 4, stage_step = stage_bytes >> 4;     // 16-byte units, as the descriptor counts
        const uint32_t a_step = a_bytes >> 4, b_step = b_bytes >> 4, bres_u = smem_u32(s_bres) >> 4;
        const bool merged = PASSES == 3 && q.merge;
        RingPos sp{0, 0};
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++titer) {
            const uint32_t buf = q.nbuf == 2 ? (titer & 1) : 0;
            const uint32_t use = q.nbuf == 2 ? (titer >> 1) : titer;          // how often this accumulator set was used before
            if (use >= 1) mbar_wait_t(&acc_empty[buf], (use - 1) & 1, &w1, rec);  // the epilogue has drained this accumulator set
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d_set = tmem_base + buf * set_cols;
            const uint32_t d_corr = d_set + (uint32_t)nmain * Nc;             // shared correction accumulator (unmerged mode)
            uint32_t d_cur = d_set;                                            // main accumulator (or pair) of this K block
            int m = 0;
            uint32_t acc_main = 0, acc_corr = q.ncorr ? 0u : 1u;              // 0: the first MMA into the accumulator overwrites it
            uint32_t bres_kb = bres_u;
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = sp.idx;
                mbar_wait_t(&a_full[s], sp.par, &w0, rec);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const uint32_t a_u = stage_u + (uint32_t)s * stage_step;
                    const uint64_t da_hi = desc0 + a_u, da_lo = da_hi + a_step;
                    const uint64_t db_hi = desc0 + (q.b_resident ? bres_kb : a_u + 2 * a_step), db_lo = db_hi + b_step;
                    const int ksteps = min(MMA_KB, q.Kpad - kb * MMA_KB) >> 3;
                    if (merged) {
                        // pair (main, correction) in adjacent TMEM columns: [B_hi;B_lo] are adjacent panels, so A_hi x both is one
                        // MMA of N = 2*Nc (a third fewer instructions and A_hi is read from shared memory once)
                        for (int t = 0; t < ksteps; ++t) {                    // 8 tf32 = 32 bytes = 2 descriptor units per k step
                            umma_tf32(d_cur, da_hi + 2 * t, db_hi + 2 * t, q.idesc2, acc_main | (uint32_t)(t > 0));
                            umma_tf32(d_cur + Nc, da_lo + 2 * t, db_hi + 2 * t, q.idesc, 1u);
                        }
                    } else {
                        for (int t = 0; t < ksteps; ++t) {
                            umma_tf32(d_cur, da_hi + 2 * t, db_hi + 2 * t, q.idesc, acc_main | (uint32_t)(t > 0));
                            if (PASSES == 3) {
                                umma_tf32(q.ncorr ? d_corr : d_cur, da_hi + 2 * t, db_lo + 2 * t, q.idesc, acc_corr);
                                umma_tf32(q.ncorr ? d_corr : d_cur, da_lo + 2 * t, db_hi + 2 * t, q.idesc, 1u);
                                acc_corr = 1u;
                            }
                        }
                    }
                    umma_commit(&a_empty[s]);
                    if (kb == nkb - 1) umma_commit(&acc_full[buf]);
                }
                __syncwarp();
                sp.advance(q.nstage);
                bres_kb += 2 * b_step;
                d_cur += merged ? 2 * Nc : Nc;
                if (++m == nmain) { m = 0; d_cur = d_set; acc_main = 1u; }    // K blocks go round-robin over the main accumulators
            }
        }
    } else {
        // =============================================== epilogue ==================================================
        // 8 warps: TMEM lane quadrant = warp % 4, the two warps of a quadrant split the 16-column groups.  Accumulator a of a set sits
        // at column a * Nc (merged: main0, corr0, main1, ...; unmerged: the mains, then the shared correction accumulator).
        const int lane_grp = warp & 3, col_half = (warp - MMA_EPI_WARP0) >> 2;
        const int row = lane_grp * 32 + lane;
        const int ncol16 = q.Nc >> 4;
        const int c16_mid = (ncol16 + 1) >> 1;
        const int c16_begin = col_half == 0 ? 0 : c16_mid, c16_end = col_half == 0 ? c16_mid : ncol16;
        const bool has_res = q.res != nullptr, has_film = q.film_mul != nullptr, has_bias = q.bias != nullptr;
        const float res_coef = q.res_coef;
        const int epi_mode = (has_film || has_bias) ? 2 : (has_res ? 1 : 0);
        float* const out_p = q.out;
        const float* const res_p = q.res;
        const int N_all = q.N;
        const uint32_t HWu = (uint32_t)HW, Ncu = (uint32_t)q.Nc;
        const uint32_t tlane = tmem_base + ((uint32_t)(lane_grp * 32) << 16);
        uint32_t titer = 0;
        const uint32_t set_cols = (uint32_t)q.set_cols;
        const int nacc = q.merge ? 2 * nmain : nmain + q.ncorr;
        const int nfirst = min(16, N_all - (chunk * q.Nc + c16_begin * 16));       // valid columns of this warp's first group
        // residual of this warp's first 16-column group, fetched one tile ahead (right after the previous tile consumed it) so its
        // HBM latency is hidden behind a tile
        float rnext[16];
        auto prefetch_res = [&](int tile_n) {
            const int bn = tile_n / tiles_per_img, pn = (tile_n - bn * tiles_per_img) * MMA_TP + row;
            const bool ok = has_res && tile_n < ntiles && pn < HW && c16_begin < c16_end;
            const float* rp = q.res + ((size_t)bn * q.N + (size_t)chunk * q.Nc + c16_begin * 16) * HW + (ok ? pn : 0);
            if (ok && nfirst >= 16) {
#pragma unroll
                for (int j = 0; j < 16; ++j) rnext[j] = rp[(size_t)HWu * j];
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) rnext[j] = (ok && j < nfirst) ? rp[(size_t)HWu * j] : 0.f;
            }
        };
        prefetch_res(blockIdx.x);
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++titer) {
            const int b = tile / tiles_per_img, p0 = (tile - b * tiles_per_img) * MMA_TP;
            const int pe = p0 + row;
            const bool valid = pe < HW;
            const size_t base = ((size_t)b * q.N + (size_t)chunk * q.Nc) * HW + (valid ? pe : 0);
            const uint32_t buf = q.nbuf == 2 ? (titer & 1) : 0;
            const uint32_t use = q.nbuf == 2 ? (titer >> 1) : titer;
            const uint32_t tacc = tlane + buf * set_cols;
            // FiLM maps of this warp's first 16-column group: independent of the accumulator, so they are requested BEFORE the wait for
            // the MMAs of the tile - their HBM latency overlaps the wait instead of following it (ncu: long_scoreboard 8.5 on the
            // FCAFFN project_in shapes)
            float fmv[16], fav[16];
            const bool film_pre = has_film && valid && c16_begin < c16_end && nfirst >= 16;
            if (film_pre) {
                const size_t g0 = base + (size_t)HWu * (uint32_t)(c16_begin * 16);
#pragma unroll
                for (int j = 0; j < 16; ++j) { fmv[j] = q.film_mul[g0 + (size_t)HWu * j]; fav[j] = q.film_add[g0 + (size_t)HWu * j]; }
            }
            mbar_wait_t(&acc_full[buf], use & 1, &w0, rec);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int c16 = c16_begin; c16 < c16_end; ++c16) {
                float acc[16];
                {
                    // accumulators are added in IEEE fp32 (one at a time: a second set of 16 registers in flight spills)
                    uint32_t ta = tacc + (uint32_t)(c16 * 16);
                    uint32_t r[16];
                    tmem_ld16(ta, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[j] = __uint_as_float(r[j]);
                    for (int a = 1; a < nacc; ++a) {
                        ta += Ncu;
                        tmem_ld16(ta, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) acc[j] += __uint_as_float(r[j]);
                    }
                }
                if (c16 == c16_end - 1) {        // this warp's TMEM reads of the tile are complete: hand the accumulators back
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[buf]);
                }
                const int n0 = chunk * q.Nc + c16 * 16;
                if (valid) {
                    const size_t goff = base + (size_t)HWu * (uint32_t)(c16 * 16);
                    const int nvalid = N_all - n0;            // >= 16: whole group, no per-column predicate
                    const bool pre = c16 == c16_begin;
                    if (epi_mode == 0) {
                        if (nvalid >= 16) epi_group<true, false, false, false>(acc, out_p + goff, nullptr, nullptr, nullptr, nullptr, rnext, false, res_coef, HWu, 16);
                        else epi_group<false, false, false, false>(acc, out_p + goff, nullptr, nullptr, nullptr, nullptr, rnext, false, res_coef, HWu, nvalid);
                    } else if (epi_mode == 1) {
                        if (nvalid >= 16) epi_group<true, true, false, false>(acc, out_p + goff, res_p + goff, nullptr, nullptr, nullptr, rnext, pre, res_coef, HWu, 16);
                        else epi_group<false, true, false, false>(acc, out_p + goff, res_p + goff, nullptr, nullptr, nullptr, rnext, pre, res_coef, HWu, nvalid);
                    } else {
                        const float* bn = has_bias ? q.bias + n0 : nullptr;
                        const float* fm = has_film ? q.film_mul + goff : nullptr;
                        const float* fa = has_film ? q.film_add + goff : nullptr;
                        // generic: null-guarded (rpre is zero when there is no residual)
                        const bool fpre = film_pre && c16 == c16_begin;
                        epi_group<false, true, true, true>(acc, out_p + goff, has_res ? res_p + goff : nullptr, fm, fa, bn, rnext, pre, res_coef, HWu, min(nvalid, 16),
                                                           fpre ? fmv : nullptr, fpre ? fav : nullptr);
                    }
                }
                // the prefetched residual has been consumed: refill the same registers for the next tile of this CTA
                if (c16 == c16_begin) prefetch_res(tile + gridDim.x);
            }
            if (c16_begin >= c16_end) {          // a warp without columns (Nc == 16) still takes part in the hand-back
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[buf]);
            }
        }
    }
    if (FDN_MMA_PROFILE && rec) {     // [0,1] producer (raw_full, a_empty) [2] epilogue (acc_full) [3,4] MMA (a_full, acc_empty) [5] loader (raw_empty) [6] total
        const int role = tid == 0 ? 0 : (tid == MMA_EPI_WARP0 * 32 ? 2 : (tid == MMA_MMA_WARP * 32 ? 3 : 5));
        atomicAdd(&q.dbg[role], w0);
        if (role == 0 || role == 3) atomicAdd(&q.dbg[role + 1], w1);
        if (role == 0) atomicAdd(&q.dbg[6], (unsigned long long)(clock64() - t_start));
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == MMA_MMA_WARP) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)q.tmem_cols) : "memory");
    }
}

// per-pixel LayerNorm statistics of G channel groups: stats[b][g][0] = mean, stats[b][g][1] = 1/sqrt(var + eps)
__global__ void __launch_bounds__(256) k_group_stats(const float* __restrict__ x, float* __restrict__ stats, int C, int HW, long long total) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // over B*G*HW
    if (i >= total) return;
    const int p = (int)(i % HW);
    const long long bg = i / HW;
    const float* xp = x + (size_t)bg * C * HW + p;
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += xp[(size_t)c * HW];
    const float mu = s / (float)C;
    float v = 0.f;
    for (int c = 0; c < C; ++c) {
        const float d = xp[(size_t)c * HW] - mu;
        v += d * d;
    }
    stats[(size_t)bg * 2 * HW + p] = mu;
    stats[((size_t)bg * 2 + 1) * HW + p] = 1.0f / sqrtf(v / (float)C + 1e-5f);
}

struct PwMmaPlan { int resident, nstage, ring; size_t smem; };

static bool pw_mma_plan(int Nc, int nkb, int prologue, bool ring_holds_tile, PwMmaPlan* out) {
    const size_t a_bytes = (size_t)MMA_TP * 128, b_bytes = (size_t)Nc * 128;
    const size_t slot = prologue == 2 ? (size_t)MMA_P2_SLOT : (size_t)MMA_SLOT_BYTES * (prologue == 3 ? 2 : 1);
    const size_t misc = 1024 + (6 * MMA_TP + 2 * MMA_MAX_K) * sizeof(float) + (2 * MMA_MAX_RING + 12) * sizeof(uint64_t) + 64;
    const size_t budget = 227 * 1024;
    const int ring_min = ring_holds_tile ? max(nkb, 2) : 2;  // LayerNorm statistics need all K blocks of a tile resident
    PwMmaPlan best;
    bool found = false;
    int ns_hi = 2, ns_lo = 1;                 // operand stages; FDN_MMA_NSTAGE pins the count (dev knob, up to 4)
    if (const char* e = getenv("FDN_MMA_NSTAGE")) { ns_hi = ns_lo = max(1, min(4, atoi(e))); }
    for (int resident = 1; resident >= 0; --resident)
        for (int nstage = ns_hi; nstage >= ns_lo; --nstage) {
            const size_t fixed = misc + (resident ? (size_t)nkb * 2 * b_bytes : 0) + nstage * (2 * a_bytes + (resident ? 0 : 2 * b_bytes));
            if (fixed + ring_min * slot > budget) continue;
            int ring = (int)((budget - fixed) / slot);
            if (ring > MMA_MAX_RING) ring = MMA_MAX_RING;
            PwMmaPlan p{resident, nstage, ring, fixed + ring * slot};
            // prefer a resident weight (no per-tile weight traffic, all loader warps stream activations) as long as the
            // ring can still prefetch one K block beyond the minimum; then the deeper ring; then two operand stages
            auto score_of = [&](const PwMmaPlan& c) {
                return (c.resident && c.ring >= ring_min + 1 ? 100 : 0) + min(c.ring - ring_min, 4) * 4 + c.resident * 2 + (c.nstage - 1);
            };
            const int score = score_of(p);
            const int best_score = found ? score_of(best) : -1;
            if (score > best_score) { best = p; found = true; }
        }
    if (found) *out = best;
    return found;
}
#endif  // !FDN_EMU

static unsigned long long* g_pw_mma_dbg = nullptr;
// Development aid: device buffer of 8 uint64 counters that every following fdn_pw_mma launch adds its per-role wait cycles to
// (NULL disables).  [0] producers on raw_full, [1] producers on a_empty, [2] epilogue on acc_full, [3] MMA on a_full,
// [4] MMA on acc_empty, [5] loader on raw_empty, [6] total cycles (summed over CTAs).
FDN_API int fdn_pw_mma_set_debug(void* counters) {
    g_pw_mma_dbg = reinterpret_cast<unsigned long long*>(counters);
    return 0;
}

FDN_API int fdn_has_tcgen05() {
#ifdef FDN_EMU
    return 0;
#else
    return 1;
#endif
}

// 1 if fdn_pw_mma can run a layer with K input channels (prologue 2: K = 3E), output chunks of Nc columns and the given prologue
// (stats_in_kernel: prologue 2 without a statistics pre-pass); 0 if its tile does not fit in shared memory or K exceeds the
// LayerNorm staging (callers then use fdn_pw_conv).  Pure host arithmetic.
FDN_API int fdn_pw_mma_supported(int K, int Nc, int prologue, int stats_in_kernel) {
#ifdef FDN_EMU
    return 0;
#else
    if (K <= 0 || K > MMA_MAX_K || Nc < 16 || Nc > 256 || (Nc & 15) || prologue < 0 || prologue > 3) return 0;
    const int Kpad = prologue == 2 ? ((K / 3 + MMA_EB - 1) / MMA_EB) * MMA_KB : ((K + 7) & ~7);
    const int nkb = (Kpad + MMA_KB - 1) / MMA_KB;
    if (nkb * MMA_KB > 1024) return 0;
    PwMmaPlan plan;
    return pw_mma_plan(Nc, nkb, prologue, prologue == 1 || prologue == 3 || (prologue == 2 && stats_in_kernel), &plan) ? 1 : 0;
#endif
}

// stats[b][g][0][p] = mean over the C channels of group g, stats[b][g][1][p] = 1/sqrt(biased var + 1e-5); x [B][G*C][HW]
FDN_API int fdn_group_stats(const float* x, float* stats, int B, int G, int C, int HW, cudaStream_t st) {
#ifdef FDN_EMU
    fdn_set_error("fdn_group_stats: not available in the host emulation build");
    return -1;
#else
    FDN_REQUIRE(x && stats && B > 0 && G > 0 && C > 0 && HW > 0, "bad arguments");
    long long total = (long long)B * G * HW;
    k_group_stats<<<fdn_cdiv(total, 256), 256, 0, st>>>(x, stats, C, HW, total);
    return fdn_check_launch("k_group_stats");
#endif
}

// Tensor-core 1x1 convolution.  bpack is the host-packed weight (see fdn_tip2025_b200/packing.py): for every output chunk of
// Nc (multiple of 16, <= 256) channels and every block of 32 input channels, a [Nc][32] tf32-hi panel followed by the tf32-lo
// panel, both in the K-major SWIZZLE_128B image.  nchunks*Nc >= N.  prologue: 0 none, 1 LayerNorm over the K inputs,
// 2 FDSA gate (three LayerNorm groups of K/3 channels - statistics precomputed in `stats`, or NULL: computed here, K/3 <= 40 - times v_value = aux), 3 FCAFFN mix
// LN(src)*aux + aux.  passes: 3 = 3xTF32 (fp32-level accuracy), 1 = single TF32.  HW must be a multiple of 4.
FDN_API int fdn_pw_mma(const float* src0, int c0, const float* src1, int c1, const float* bpack, int N, int Nc, int nchunks,
                       int prologue, const float* ln_w, const float* ln_b, const float* aux, long long aux_bs, const float* stats,
                       const float* bias, const float* film_mul, const float* film_add, const float* res, float res_coef,
                       float* out, int B, int HW, int passes, cudaStream_t st) {
#ifdef FDN_EMU
    fdn_set_error("fdn_pw_mma: tcgen05 kernels are not available in the host emulation build");
    return -1;
#else
    FDN_REQUIRE(src0 && bpack && out && B > 0 && HW > 0 && c0 > 0 && N > 0, "bad arguments");
    FDN_REQUIRE(Nc % 16 == 0 && Nc >= 16 && Nc <= 256 && (long long)nchunks * Nc >= N, "bad output chunking");
    FDN_REQUIRE(prologue >= 0 && prologue <= 3 && (passes == 1 || passes == 3), "bad mode");
    FDN_REQUIRE((film_mul == nullptr) == (film_add == nullptr), "film needs both maps");
    FDN_REQUIRE(HW % 4 == 0, "HW must be a multiple of 4 (16-byte bulk copies)");
    if (prologue) FDN_REQUIRE(ln_w && ln_b && !src1, "LayerNorm prologue needs gamma/beta and a single source");
    if (prologue >= 2) FDN_REQUIRE(aux != nullptr && fdn_aligned16(aux) && aux_bs % 4 == 0, "prologue needs a 16-byte aligned aux tensor");
    if (prologue == 2) FDN_REQUIRE(c0 % 3 == 0 && (!stats || fdn_aligned16(stats)), "FDSA gate needs 3 equal groups (statistics: precomputed, or NULL = in the kernel)");
    FDN_REQUIRE(fdn_aligned16(bpack) && fdn_aligned16(src0) && (!src1 || fdn_aligned16(src1)), "pointers must be 16-byte aligned");
    PwMmaParams q;
    q.src0 = src0; q.src1 = src1; q.C0 = c0; q.C1 = src1 ? c1 : 0;
    q.Kreal = q.C0 + q.C1;
    q.E = prologue == 2 ? c0 / 3 : 0;
    if (prologue == 2) {          // grouped layout: ceil(E/10) blocks of 3 x 10 channels (rows 30, 31 of each block are zero)
        q.K = ((q.E + MMA_EB - 1) / MMA_EB) * MMA_KB;
        q.Kpad = q.K;
    } else {
        q.K = q.Kreal;
        q.Kpad = (q.K + 7) & ~7;
    }
    q.N = N; q.Nc = Nc; q.HW = HW; q.B = B;
    q.bpack = bpack; q.prologue = prologue; q.ln_w = ln_w; q.ln_b = ln_b; q.aux = aux; q.aux_bs = aux_bs; q.stats = stats;
    q.bias = bias; q.film_mul = film_mul; q.film_add = film_add; q.res = res; q.res_coef = res_coef; q.out = out;
    q.passes = passes;
    q.dbg = g_pw_mma_dbg;
    // instruction descriptor: D=f32 (bit 4), A=B=tf32 (2<<7, 2<<10), both K-major, N>>3 at bit 17, M>>4 at bit 24
    q.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(Nc >> 3) << 17) | ((uint32_t)(MMA_TP >> 4) << 24);
    const int nkb = (q.Kpad + MMA_KB - 1) / MMA_KB;
    PwMmaPlan plan;
    FDN_REQUIRE(pw_mma_plan(Nc, nkb, prologue, prologue == 1 || prologue == 3 || (prologue == 2 && stats == nullptr), &plan), "tile does not fit in shared memory");
    q.b_resident = plan.resident; q.nstage = plan.nstage; q.ring = plan.ring;
    // Accumulators.  The tensor core truncates its fp32 accumulator after every instruction, so 3xTF32 keeps the large hi*hi sum and
    // the small corrections in separate TMEM accumulators and spreads long K over several main accumulators, which the epilogue adds
    // in IEEE fp32 - but only as far as TWO accumulator sets still fit the 512 TMEM columns: without the second set the epilogue of a
    // tile cannot overlap the MMAs of the next one, which costs far more (L2 / L3 to_hidden ran 1.6x slower than single-pass TF32
    // for that reason alone) than the ~1e-6 the extra accumulators buy.  One K block (<= 12 accumulations) needs no split at all.
    q.ncorr = (passes == 3 && nkb >= 2) ? 1 : 0;
    q.nmain = 1;
    q.merge = (passes == 3 && nkb >= 2 && 2 * Nc <= 256) ? 1 : 0;
    if (const char* e = getenv("FDN_MMA_MERGE")) q.merge = q.merge && atoi(e) != 0;
    // main accumulators wanted: one up to K = 128 (<= 16 truncating accumulations each), two up to 256, three beyond - every extra
    // accumulator costs the epilogue a TMEM load and 16 additions per thread and 16-column group
    const int want = nkb <= 4 ? 1 : (nkb <= 8 ? 2 : 3);
    int set_cols;
    if (q.merge) {
        // (main, correction) pairs of 2*Nc columns
        q.nmain = max(1, min(min(want, nkb), 256 / (2 * Nc)));
        q.ncorr = 0;
        set_cols = 2 * q.nmain * Nc;
        q.idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * Nc) >> 3) << 17) | ((uint32_t)(MMA_TP >> 4) << 24);
    } else {
        if (q.ncorr && 2 * Nc > 256) q.ncorr = 0;          // Nc > 128: main and corrections share one accumulator
        if (nkb >= 2) q.nmain = max(1, min(min(want, nkb), (256 / Nc) - q.ncorr));
        set_cols = (q.nmain + q.ncorr) * Nc;
        q.idesc2 = q.idesc;
    }
    q.set_cols = set_cols;
    q.nbuf = 2 * set_cols <= 512 ? 2 : 1;
    // copy requests per tile: the TMA unit handles ~1 small bulk request per 64 cycles, so many-row tiles use cp.async instead
    // TMA bulk copies (one per 512-byte channel row) measured equal or faster than 16-byte cp.async for every layer shape once
    // the gate prologue stopped re-loading v_value three times; cp.async stays selectable with FDN_MMA_BULK=0
    q.bulk = 1;
    if (const char* e = getenv("FDN_MMA_NBUF")) { if (atoi(e) == 1) { q.nbuf = 1; q.tmem_cols = 32; while (q.tmem_cols < set_cols) q.tmem_cols <<= 1; } }
    if (const char* e = getenv("FDN_MMA_BULK")) q.bulk = atoi(e);
    q.tmem_cols = 32;
    while (q.tmem_cols < q.nbuf * set_cols) q.tmem_cols <<= 1;
    FDN_REQUIRE(q.tmem_cols <= 512, "accumulators do not fit in tensor memory");
    FDN_REQUIRE(nkb * MMA_KB <= 1024 && q.Kreal <= MMA_MAX_K, "too many input channels");
    const int dev = fdn_device();
    static int num_sms_dev[FDN_MAX_DEVICES] = {0};
    if (num_sms_dev[dev] == 0) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        num_sms_dev[dev] = n > 0 ? n : 148;
    }
    const int num_sms = num_sms_dev[dev];
    const int ntiles = fdn_cdiv(HW, MMA_TP) * B;
    const int gx = min(ntiles, max(1, num_sms / nchunks));
    void (*kern)(PwMmaParams) = nullptr;
    switch (prologue * 2 + (passes == 3 ? 1 : 0)) {
        case 0: kern = k_pw_mma<0, 1>; break;
        case 1: kern = k_pw_mma<0, 3>; break;
        case 2: kern = k_pw_mma<1, 1>; break;
        case 3: kern = k_pw_mma<1, 3>; break;
        case 4: kern = k_pw_mma<2, 1>; break;
        case 5: kern = k_pw_mma<2, 3>; break;
        case 6: kern = k_pw_mma<3, 1>; break;
        default: kern = k_pw_mma<3, 3>; break;
    }
    static bool configured[FDN_MAX_DEVICES][8] = {};      // the opt-in is a per-device function attribute
    const int ki = prologue * 2 + (passes == 3 ? 1 : 0);
    if (!configured[dev][ki]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { fdn_set_error(cudaGetErrorString(e)); return (int)e; }
        configured[dev][ki] = true;
    }
    kern<<<dim3(gx, nchunks, 1), dim3(MMA_THREADS), plan.smem, st>>>(q);
    return fdn_check_launch("k_pw_mma");
#endif
}
