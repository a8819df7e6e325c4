// Log-semiring batched matmul on the 5th-generation tensor cores (tcgen05, accumulators in tensor memory).
//
// Reference semantics: Contraction[Logaddexp, Add, Tensor, Tensor] funsor/cnf.py:335-379 ->
// einsum/numpy_log.py:24-47 for the scan-level equation `abcd,abde->abce`:
//     out[n,i,j] = (sx[n,i] + sy[n,j]) + log sum_k exp(x[n,i,k] - sx[n,i]) * exp(y[n,k,j] - sy[n,j])
// with sx / sy the row / column maxima clamped to finfo.min.  The exp is applied while the operands are
// staged into shared memory, the inner contraction is a plain GEMM and runs as 3xTF32
// (hi*hi + lo*hi + hi*lo, fp32 accumulation in TMEM), the log and the shifts are the epilogue.
//
// One CTA = one 128 x 128 output tile of one batch element.  K is consumed in chunks of 32:
//   warps 0-3 stage X (one row per thread: exp, TF32 hi/lo split, K-major canonical layout),
//   warps 4-7 stage Y (4 x 8 patches, exp, split, register transpose into the same K-major canonical layout),
//   after a proxy fence + barrier one elected lane issues the 12 MMAs (4 k-steps x 3 products) of the chunk and
//   commits them to the buffer's mbarrier; staging of the next chunk overlaps the tensor pipe (2 buffers).
// Epilogue: all 8 warps read their quarter of the accumulator lanes with tcgen05.ld, apply log + shifts, store.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace fb2 {

constexpr int LM_TILE = 128;                            // output rows per CTA (MMA M)
constexpr int LM_KC = 32;                               // K per staged chunk
constexpr int LM_STAGES = 2;
template <int BN>
struct LmCfg {
    static constexpr int A_OPB = LM_TILE * LM_KC * 4;   // bytes of one A tile (hi or lo): 16 KB
    static constexpr int B_OPB = BN * LM_KC * 4;        // bytes of one B tile (hi or lo): 16 / 32 KB
    static constexpr int STAGEB = 2 * A_OPB + 2 * B_OPB;
    static constexpr int SMEM = LM_STAGES * STAGEB + 1024;  // + alignment slack
};

struct LogMatmulArgs {
    const float* x;
    const float* y;
    const float* sx;  // [n, I] clamped row maxima of x
    const float* sy;  // [n, J] clamped column maxima of y
    float* out;
    int64_t n1, os0;
    int64_t xs0, xs1, xsi;  // xsk == 1
    int64_t ys0, ys1, ysk;  // ysj == 1
    int I, K, J;
    float comp;  // expected relative truncation loss per accumulating MMA (0 = no compensation)
    int mode;  // FB2_LM_MODE experiments: 1 = stage only the first two chunks, 2 = do not issue MMAs
};

__device__ __forceinline__ uint32_t lm_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// TF32 "hi" part by truncation (a LOP3 on the ALU pipe; cvt.rna.tf32 is a quarter-rate conversion and two of them per
// element made the staging XU-bound).  lo = x - hi is exact in fp32 and the tensor core itself ignores its low 13 bits,
// so hi + lo carries 21+ mantissa bits either way.
__device__ __forceinline__ uint32_t lm_tf32(float x) { return __float_as_uint(x) & 0xffffe000u; }
// exp(d) for d <= 0 as one FMUL + MUFU.EX2 (flush-to-zero).  __expf without fast-math wraps the same MUFU in denormal
// handling (FSETP + two more FMULs per element: a quarter of all instructions the kernel issued).
__device__ __forceinline__ float lm_exp(float d) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d * 1.4426950408889634f));
    return r;
}
// K-major operand tile in the 128-byte-swizzle canonical layout: row r (an output row of X / an output column of Y) at
// r * 128 B, its eight 16-byte k-chunks XOR-swizzled with (r % 8); 8-row groups 1024 B apart (SBO); one staged chunk of
// K = 32 TF32 is exactly one swizzle row, and a K = 8 instruction advances the start address by 32 B.  (The no-swizzle
// "interleave" layout was measured first: ~600 cycles per 128x128x8 MMA, i.e. shared-memory-bound.)
__device__ __forceinline__ uint64_t lm_desc(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void lm_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void lm_tmem_ld16(uint32_t taddr, uint32_t (&d)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(d[8]), "=r"(d[9]),
          "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t lm_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}

// sx[n,i] = max(max_k x[n,i,k], -FLT_MAX): one warp per row (k contiguous)
__global__ void lm_rowmax(const float* x, int64_t n1, int64_t xs0, int64_t xs1, int64_t xsi, int I, int K, int64_t rows, float* sx) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    if (row >= rows) return;
    const int lane = threadIdx.x % 32;
    const int64_t n = row / I;
    const int i = (int)(row % I);
    const float* p = x + (n / n1) * xs0 + (n % n1) * xs1 + (int64_t)i * xsi;
    float m = -FLT_MAX;
    for (int k = lane; k < K; k += 32) m = fmaxf(m, p[k]);
    m = warp_max(m);
    if (lane == 0) sx[row] = m;
}
// sy[n,j] = max(max_k y[n,k,j], -FLT_MAX): threads over j (contiguous), 8-way split of k, combined through shared memory
__global__ void lm_colmax(const float* y, int64_t n1, int64_t ys0, int64_t ys1, int64_t ysk, int K, int J, float* sy) {
    __shared__ float red[8][32];
    const int64_t n = blockIdx.y;
    const int j = blockIdx.x * 32 + threadIdx.x % 32;
    const int part = threadIdx.x / 32;
    const float* p = y + (n / n1) * ys0 + (n % n1) * ys1 + j;
    float m = -FLT_MAX;
    if (j < J)
        for (int k = part; k < K; k += 8) m = fmaxf(m, p[(int64_t)k * ysk]);
    red[part][threadIdx.x % 32] = m;
    __syncthreads();
    if (part == 0 && j < J) {
#pragma unroll
        for (int q = 1; q < 8; ++q) m = fmaxf(m, red[q][threadIdx.x]);
        sy[n * J + j] = m;
    }
}

// BN = output columns per CTA (MMA N): 256 when J > 128 (each staged x row then feeds 256 columns, halving the L2 -> SM
// traffic per flop, which is what bounded the 128 x 128 version), else 128.
//
// NMAIN + 1 accumulators in tensor memory.  Measured on B200 (scripts/matmul_bias.py): every tcgen05.mma accumulation
// TRUNCATES the fp32 accumulator toward zero, a one-sided loss of ~2^-24 of the running sum per instruction; with all
// 3 K/8 instructions of a tile chained into one accumulator the result was low by 2.2e-8 K (relative) -- 1.1e-5 at K = 512,
// nearly deterministic, and it adds up along a chain of products.  Hence: the two correction products (lo*hi, hi*lo, 2^-11 of the
// magnitude) go to their own accumulator, where their truncation is negligible, and the hi*hi products of successive K chunks
// rotate over NMAIN accumulators; the NMAIN + 1 partial sums are added with round-to-nearest in the epilogue.  That divides
// the loss by 3 NMAIN.  What remains is one-sided and, with >= 8 accumulations, close to its mean (rms error ~ |mean error|
// in the measurement), so the epilogue multiplies every hi*hi partial sum by 1 + 5.82e-8 * (its number of accumulations):
// that centres the error (residual in scripts/matmul_bias.py / DESIGN.md 4.4).  BN * (NMAIN + 1) = 512 columns = all of the SM's tensor memory (one CTA per SM anyway).
//
// NTH threads per CTA: the staging work (loads, exp, split, swizzled stores) is what limits the kernel (ncu: tensor pipe 34 %
// busy, long-scoreboard stalls with 8 warps per SM), so the wide tile runs 16 warps that each stage half as much.
template <int BN, int NMAIN, int NTH>
__global__ void __launch_bounds__(NTH, 1) log_matmul_tc(LogMatmulArgs a) {
    constexpr int XM = 128 / (NTH / 8);      // X rows per thread (rows rbase + (NTH/8) m)
    constexpr int RP = NTH / 8;              // X rows staged per pass
    constexpr int YC = BN >= NTH / 8 * 8 ? 8 : BN / (NTH / 8);  // Y columns per thread (8, or 4 with 512 threads on the 256-column tile)
    constexpr int YV = YC / 4;               // float4 loads per k row
    static_assert(YC == 8 || YC == 4, "unsupported thread count for this tile");
    constexpr uint32_t NCOLS = BN * (NMAIN + 1);
    static_assert(NCOLS == 512 || NCOLS == 256 || NCOLS == 128, "tensor memory allocation must be a power of two <= 512");
    using Cfg = LmCfg<BN>;
    extern __shared__ unsigned char lm_smem_raw[];
    __shared__ __align__(8) unsigned long long s_bar[LM_STAGES];
    __shared__ __align__(8) unsigned long long s_done;
    __shared__ uint32_t s_tmem;
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(lm_smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n = blockIdx.z;
    const int i0 = blockIdx.y * LM_TILE, j0 = blockIdx.x * BN;
    const int64_t b0 = n / a.n1, b1 = n % a.n1;
    const float* xb = a.x + b0 * a.xs0 + b1 * a.xs1;
    const float* yb = a.y + b0 * a.ys0 + b1 * a.ys1;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < LM_STAGES; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(lm_smem_u32(&s_bar[s])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(lm_smem_u32(&s_done)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(lm_smem_u32(&s_tmem)), "n"(NCOLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    // ---- staging roles (every thread stages a part of X and a part of Y)
    // X: thread t owns k-chunk kc = t % 8 of rows t/8 + 32 m (m = 0..3): a warp instruction reads 4 full 128-byte lines.
    // Y: thread t (t < BN) owns k = 4 kq .. 4 kq + 3 (kq = t % 8) of the chunk and columns 8 jq .. 8 jq + 7 (jq = t / 8): a warp
    //    instruction reads 8 full lines; the 4 x 8 patch is transposed in registers.
    // Global loads for chunk c+1 are issued right after chunk c has been written to shared memory (register prefetch).
    const int kc = tid & 7, rbase = tid >> 3;
    const int kq = tid & 7, jq = tid >> 3;
    const bool has_y = jq * YC < BN;
    float sxr[XM], syr[YC];
#pragma unroll
    for (int m = 0; m < XM; ++m) {
        const int row = i0 + rbase + RP * m;
        sxr[m] = row < a.I ? a.sx[n * a.I + row] : 0.f;
    }
#pragma unroll
    for (int q = 0; q < YC; ++q) {
        const int j = j0 + jq * YC + q;
        syr[q] = (has_y && j < a.J) ? a.sy[n * a.J + j] : 0.f;
    }
    const bool x_vec = (a.xsi % 4 == 0) && (reinterpret_cast<uintptr_t>(xb) % 16 == 0);
    const bool y_vec = (a.ysk % 4 == 0) && (reinterpret_cast<uintptr_t>(yb) % 16 == 0);

    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((128u >> 4) << 24);  // A, B K-major
    const int nchunks = (a.K + LM_KC - 1) / LM_KC;

    float4 prex[XM], prey[4 * YV];  // prefetched raw values of the next chunk
    auto load_chunk = [&](int c) {
        const int k0 = c * LM_KC;
        {
            const int k = k0 + kc * 4;
#pragma unroll
            for (int m = 0; m < XM; ++m) {
                const int row = i0 + rbase + RP * m;
                const float* p = xb + (int64_t)(row < a.I ? row : 0) * a.xsi + k;
                if (row < a.I && k + 3 < a.K && x_vec) {
                    prex[m] = __ldg(reinterpret_cast<const float4*>(p));
                } else {
                    float v[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) v[q] = (row < a.I && k + q < a.K) ? __ldg(p + q) : -INFINITY;
                    prex[m] = make_float4(v[0], v[1], v[2], v[3]);
                }
            }
        }
        if (has_y) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const int k = k0 + kq * 4 + kk;
                const float* p = yb + (int64_t)(k < a.K ? k : 0) * a.ysk + j0 + jq * YC;
#pragma unroll
                for (int h = 0; h < YV; ++h) {
                    const int j = j0 + jq * YC + h * 4;
                    if (k < a.K && j + 3 < a.J && y_vec) {
                        prey[kk * YV + h] = __ldg(reinterpret_cast<const float4*>(p + h * 4));
                    } else {
                        float v[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) v[q] = (k < a.K && j + q < a.J) ? __ldg(p + h * 4 + q) : -INFINITY;
                        prey[kk * YV + h] = make_float4(v[0], v[1], v[2], v[3]);
                    }
                }
            }
        }
    };
    // Fast path (interior of the problem: vectorisable strides, all of this thread's rows / columns in range, chunk fully
    // inside K): running pointers and unpredicated 16-byte loads.  The generic lambda above cost ~2x the instructions of the
    // arithmetic itself (ncu: 5600 warp instructions per CTA-chunk against ~2800 for load + exp + split + store).
    const float* xp[XM];
    bool xrow[XM];
#pragma unroll
    for (int m = 0; m < XM; ++m) {
        const int row = i0 + rbase + RP * m;
        xrow[m] = row < a.I;
        xp[m] = xb + (int64_t)(xrow[m] ? row : 0) * a.xsi + kc * 4;
    }
    const float* yp = yb + (int64_t)(kq * 4) * a.ysk + j0 + jq * YC;
    const bool thread_fast = x_vec && y_vec && (!has_y || j0 + jq * YC + YC <= a.J);
    const int64_t ystep = a.ysk;
    auto load_chunk_fast = [&](int c) {
        const int k0 = c * LM_KC;
#pragma unroll
        for (int m = 0; m < XM; ++m)
            prex[m] = xrow[m] ? __ldg(reinterpret_cast<const float4*>(xp[m] + k0)) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        if (has_y) {
            const float* p = yp + (int64_t)k0 * ystep;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                for (int h = 0; h < YV; ++h) prey[kk * YV + h] = __ldg(reinterpret_cast<const float4*>(p + kk * ystep + h * 4));
        }
    };
    auto load_any = [&](int c) {
        if (thread_fast && (c + 1) * LM_KC <= a.K) load_chunk_fast(c);
        else load_chunk(c);
    };
    load_any(0);

    for (int c = 0; c < nchunks; ++c) {
        const int s = c % LM_STAGES;
        unsigned char* stage = smem + (size_t)s * Cfg::STAGEB;
        // the MMAs that read this buffer LM_STAGES chunks ago must have completed
        if (c >= LM_STAGES && !(a.mode & 2)) {
            const uint32_t parity = ((c / LM_STAGES) - 1) & 1;
            while (!lm_try_wait(lm_smem_u32(&s_bar[s]), parity)) {
            }
        }
        if (!((a.mode & 1) && c >= LM_STAGES)) {
            // A operand: 16-byte k-chunk kc of row r at  r * 128 + ((kc ^ (r % 8)) * 16)
#pragma unroll
            for (int m = 0; m < XM; ++m) {
                const int r = rbase + RP * m;
                const float v[4] = {prex[m].x, prex[m].y, prex[m].z, prex[m].w};
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float e = lm_exp(v[q] - sxr[m]);
                    hi[q] = lm_tf32(e);
                    lo[q] = __float_as_uint(e - __uint_as_float(hi[q]));
                }
                unsigned char* dst = stage + r * 128 + ((kc ^ (r & 7)) << 4);
                *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(dst + Cfg::A_OPB) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
            // B operand (same layout with j in the role of the row): column j = 8 jq + q, chunk kq at j * 128 + ((kq ^ q) * 16);
            // the eight threads of a quarter warp differ in kq, i.e. hit eight different bank groups.
            if (has_y) {
                unsigned char* Bh = stage + 2 * Cfg::A_OPB;
#pragma unroll
                for (int q = 0; q < YC; ++q) {
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const float4 pv = prey[kk * YV + (q >> 2)];
                        const float raw = (q & 3) == 0 ? pv.x : (q & 3) == 1 ? pv.y : (q & 3) == 2 ? pv.z : pv.w;
                        const float e = lm_exp(raw - syr[q]);
                        hi[kk] = lm_tf32(e);
                        lo[kk] = __float_as_uint(e - __uint_as_float(hi[kk]));
                    }
                    const int rr = jq * YC + q;  // row of the B tile = output column
                    unsigned char* dst = Bh + rr * 128 + ((kq ^ (rr & 7)) << 4);
                    *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4*>(dst + Cfg::B_OPB) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
            }
        }
        if (c + 1 < nchunks && !((a.mode & 8) && c >= 1)) load_any(c + 1);  // (mode 8: experiment, no operand loads after the second chunk)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (warp == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t elected;
            asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\tselp.u32 %0, 1, 0, e;\n\t}" : "=r"(elected));
            if (elected && !((a.mode & 2) && c < nchunks - 1)) {
                const uint32_t ah = lm_smem_u32(stage), al = ah + Cfg::A_OPB, bh = ah + 2 * Cfg::A_OPB, bl = bh + Cfg::B_OPB;
#pragma unroll
                for (int ks = 0; ks < LM_KC / 8; ++ks) {
                    const uint64_t dah = lm_desc(ah + ks * 32), dal = lm_desc(al + ks * 32);
                    const uint64_t dbh = lm_desc(bh + ks * 32), dbl = lm_desc(bl + ks * 32);
                    lm_mma(tmem + (uint32_t)(c % NMAIN) * BN, dah, dbh, IDESC, (c >= NMAIN || ks > 0) ? 1u : 0u);
                    if (!(a.mode & 4)) {  // (mode 4: experiment, hi*hi products only)
                        lm_mma(tmem + NMAIN * BN, dal, dbh, IDESC, (c > 0 || ks > 0) ? 1u : 0u);
                        lm_mma(tmem + NMAIN * BN, dah, dbl, IDESC, 1u);
                    }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(lm_smem_u32(&s_bar[s])) : "memory");
                if (c == nchunks - 1)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(lm_smem_u32(&s_done)) : "memory");
            }
            __syncwarp();
        }
    }

    // ---- epilogue: warp w reads lanes 32*(w%4).. and columns [BN/2*(w/4), +BN/2)
    while (!lm_try_wait(lm_smem_u32(&s_done), 0)) {
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
        constexpr int CW = BN / (NTH / 128);  // columns per warp: the NTH / 32 warps are 4 lane quarters x NTH / 128 column ranges
        const int row = i0 + (warp & 3) * 32 + lane;
        const int cbase = (warp >> 2) * CW;
        const float sxo = row < a.I ? a.sx[n * a.I + row] : 0.f;
        float* orow = a.out + b0 * a.os0 + b1 * (int64_t)a.I * a.J + (int64_t)row * a.J;
        const bool vec_out = (a.J % 4 == 0) && (reinterpret_cast<uintptr_t>(a.out) % 16 == 0);
#pragma unroll 1
        for (int cc = 0; cc < CW; cc += 16) {
            float d[16];
            {
                // the correction accumulator first, then the hi*hi partial sums that were actually written
                uint32_t r[16];
                lm_tmem_ld16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(NMAIN * BN + cbase + cc), r);
#pragma unroll
                for (int q = 0; q < 16; ++q) d[q] = __uint_as_float(r[q]);
#pragma unroll
                for (int mi = 0; mi < NMAIN; ++mi) {
                    if (mi < nchunks) {
                        lm_tmem_ld16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(mi * BN + cbase + cc), r);
                        // expected-loss compensation: this accumulator took (LM_KC / 8) * ceil((nchunks - mi) / NMAIN) truncating
                        // accumulations, each losing on average 0.976 * 2^-24 of the sum (calibrated, see above)
                        const float f = 1.f + a.comp * (float)((LM_KC / 8) * ((nchunks - mi + NMAIN - 1) / NMAIN));
#pragma unroll
                        for (int q = 0; q < 16; ++q) d[q] = fmaf(__uint_as_float(r[q]), f, d[q]);
                    }
                }
            }
            if (row < a.I) {
                const int jb = j0 + cbase + cc;
                float o[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) o[q] = (sxo + ((jb + q) < a.J ? __ldg(a.sy + n * a.J + jb + q) : 0.f)) + __logf(d[q]);
                if (jb + 15 < a.J && vec_out) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) *reinterpret_cast<float4*>(orow + jb + 4 * q) = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
                } else {
#pragma unroll
                    for (int q = 0; q < 16; ++q)
                        if (jb + q < a.J) orow[jb + q] = o[q];
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(NCOLS) : "memory");
}

// Returns FB2_ERR_UNSUPPORTED when the shapes / strides do not fit this kernel (caller falls back to the SIMT tile kernel).
int launch_log_matmul_tc(const float* x, const int64_t* xs, const float* y, const int64_t* ys, float* out, int64_t n0, int64_t n1, int I,
                         int K, int J, int64_t out_s0, float* sx_sy_workspace, cudaStream_t stream) {
    if (xs[3] != 1 || ys[3] != 1 || I < 64 || J < 64 || K < 32 || sx_sy_workspace == nullptr) return FB2_ERR_UNSUPPORTED;
    const int64_t n = n0 * n1;
    if (n <= 0 || n > 65535) return FB2_ERR_UNSUPPORTED;
    float* sx = sx_sy_workspace;
    float* sy = sx + n * I;
    const int64_t rows = n * I;
    lm_rowmax<<<(unsigned)ceil_div(rows, 8), 256, 0, stream>>>(x, n1, xs[0], xs[1], xs[2], I, K, rows, sx);
    FB2_LAUNCH_CHECK();
    lm_colmax<<<dim3((unsigned)ceil_div(J, 32), (unsigned)n), 256, 0, stream>>>(y, n1, ys[0], ys[1], ys[2], K, J, sy);
    FB2_LAUNCH_CHECK();
    LogMatmulArgs a;
    a.x = x; a.y = y; a.sx = sx; a.sy = sy; a.out = out; a.n1 = n1;
    a.os0 = out_s0 >= 0 ? out_s0 : n1 * (int64_t)I * J;
    a.xs0 = xs[0]; a.xs1 = xs[1]; a.xsi = xs[2];
    a.ys0 = ys[0]; a.ys1 = ys[1]; a.ysk = ys[2];
    a.I = I; a.K = K; a.J = J;
    {
        const char* m = getenv("FB2_LM_MODE");
        a.mode = m ? atoi(m) : 0;
        const char* c = getenv("FB2_LM_COMP");  // A/B switch: 0 disables the expected-loss compensation
        a.comp = (c && atoi(c) == 0) ? 0.f : 5.82e-8f;
    }
    // per call, not cached in a static: the attribute is per device and a process may drive more than one
    FB2_CUDA(cudaFuncSetAttribute(log_matmul_tc<128, 3, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, LmCfg<128>::SMEM));
    FB2_CUDA(cudaFuncSetAttribute(log_matmul_tc<256, 1, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, LmCfg<256>::SMEM));
    FB2_CUDA(cudaFuncSetAttribute(log_matmul_tc<256, 1, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, LmCfg<256>::SMEM));
    // Tile choice: 128 x 256 (one hi*hi accumulator, 110 TFLOP/s) is the faster tile, 128 x 128 (three, 78 TFLOP/s) has a third
    // of the accumulation loss to compensate; measured rms error after compensation 5.5e-7 vs 2.9e-7 at K = 1024, so only
    // longer contractions take the 128-column tile.  FB2_LM_BN=128|256 forces one.
    const char* bn_env = getenv("FB2_LM_BN");
    const int forced = bn_env ? atoi(bn_env) : 0;
    const bool wide = forced == 256 ? J > 128 : forced == 128 ? false : (J > 128 && K <= 1024);
    if (wide) {
        dim3 grid((unsigned)ceil_div(J, 256), (unsigned)ceil_div(I, LM_TILE), (unsigned)n);
        const char* th_env = getenv("FB2_LM_THREADS");  // A/B switch: 256 = eight staging warps
        if (th_env && atoi(th_env) == 256) log_matmul_tc<256, 1, 256><<<grid, 256, LmCfg<256>::SMEM, stream>>>(a);
        else log_matmul_tc<256, 1, 512><<<grid, 512, LmCfg<256>::SMEM, stream>>>(a);
    } else {
        dim3 grid((unsigned)ceil_div(J, 128), (unsigned)ceil_div(I, LM_TILE), (unsigned)n);
        log_matmul_tc<128, 3, 256><<<grid, 256, LmCfg<128>::SMEM, stream>>>(a);
    }
    FB2_LAUNCH_CHECK();
    return FB2_OK;
}

size_t log_matmul_tc_workspace_floats(int64_t n, int I, int J) { return (size_t)n * ((size_t)I + (size_t)J); }

}  // namespace fb2
