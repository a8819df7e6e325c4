// Linear-domain matrix-chain products on the 5th-generation tensor cores: TMA-fed, warp-specialised, persistent 3xTF32 GEMM.
//
// What it computes.  The chunk summary of a time-sharded HMM (mixed_sequential_sum_product sum_product.py:766-795 with
// segments = GPUs) and every large dense scan are chains of S x S matrix products in the log semiring (sequential_sum_product
// sum_product.py:652-701; one level = Contraction[Logaddexp, Add, Tensor, Tensor] cnf.py:335-379).  The reference's own trick
// (einsum/numpy_log.py:24-47: subtract a shift, exp, multiply, log, add the shift back) is applied ONCE per chain instead of once per
// level: all elements of the tree are kept in the LINEAR domain as non-negative fp32 values times a power of two per matrix, so that a
// level is a plain GEMM, computed as 3xTF32 (hi*hi + lo*hi + hi*lo, fp32 accumulation in tensor memory) for fp32-class accuracy.
//
// Data movement.  A matrix is ONE fp32 plane in HBM.  The tensor core reads an fp32 container as TF32 by ignoring the low 13
// mantissa bits, so the plane as loaded by TMA already is the "hi" operand; the "lo" operand, v - trunc_tf32(v), is computed in
// shared memory by transform warps (one LOP + one FSUB per element, same swizzled offset in the neighbouring buffer).  Measured on
// the first version of this kernel, which kept hi / lo planes in HBM: at S = 512 a product moved 6 MB for 0.8 GFLOP of issued tensor
// work and the kernel ran at 76 % of HBM bandwidth -- memory-bound; computing lo on chip halves every stream (and lets the leaf
// level scale its left operand by the observation weights on the fly instead of materialising scaled copies).
//
//     warp 0     producer : cp.async.bulk.tensor (TMA, 128-byte swizzle) of the raw A / B tiles of a K chunk into a 2-stage ring,
//                           mbarrier complete_tx; runs ahead of the MMA warp across tiles
//     warps 2-7  transform: lo = v - trunc(v) (optionally v *= w[k] first: the leaf level's diag(exp(obs))), fence.proxy.async,
//                           arrive on the stage's `ready` barrier
//     warp 1     MMA      : tcgen05.mma.cta_group::1.kind::tf32, M = 128, N = 256, K = 8; 12 instructions per chunk; tcgen05.commit
//                           releases the stage; the last commit of a tile hands the accumulators to the epilogue
//     warps 8-15 epilogue : tcgen05.ld of the warp's 32 x 128 block into REGISTERS, tensor memory handed back at once (the next
//                           tile's MMAs overlap the stores), optional column scaling, exact power-of-two renormalisation, 32 x 16
//                           sub-blocks staged in shared memory and written by TMA stores in the orientation the NEXT level needs,
//                           per-matrix max by atomicMax.  setmaxnreg moves registers from the other warpgroups to these two.
//
// Persistent: grid = min(#tiles, #SMs); tile t = product * tiles_per_product + tile, so neighbouring CTAs share a product's operands in L2.
//
// Orientation.  Both MMA operands are K-major in shared memory, so a matrix that will be a RIGHT operand is stored transposed.  In
// the reference's tree (pairs (2p, 2p+1), odd leftover appended, sum_product.py:690-701) the role of an element is the parity of its
// index at the level where it is finally paired; cg_role walks the carry chain of the last element.
//
// Scaling.  stored = true / 2^E with an integer E per matrix: leaves are <= 1; the epilogue multiplies by 2^-(ea + eb), ea / eb the
// binary exponents of the operands' largest stored entries (read from the level below), which bounds every stored value by 4 K and
// keeps E exact.  Entries more than ~2^-126 below the largest entry of their matrix flush to zero (the same statement as for the
// dataflow kernels).
//
// Accumulation bias: as in log_matmul_tc.cu the correction products have their own accumulator and the hi*hi sum is multiplied by
// 1 + 5.82e-8 * (number of accumulations) (measured one-sided truncation of the tensor-memory accumulator).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace fb2 {

// K chunk per stage: 32 fp32 (one 128-byte swizzle row, 2 stages of 96 KB; default) or 16 (64-byte swizzle rows, 4 stages of 48 KB;
// -DCG_KC_VALUE=16).  The deeper ring was built to test whether the main loop (2.3k cycles per 32-wide chunk against 1.57k of tensor
// work) was waiting for loads: it measured the same (1.60 vs 1.55 us/step), so it is not.  What the numbers add up to is SHARED-MEMORY
// BANDWIDTH: per chunk the tensor core reads 144 KB of operands (A 4 KB + B 8 KB per instruction, 12 instructions), the transform warps
// read and write 48 KB each and TMA writes 48 KB: 288 KB at 128 B/cycle = 2.25k cycles.
#ifndef CG_KC_VALUE
#define CG_KC_VALUE 32
#endif
constexpr int CG_BM = 128, CG_BN = 256, CG_KC = CG_KC_VALUE, CG_STAGES = CG_KC == 32 ? 2 : 4;
constexpr int CG_ROW_BYTES = CG_KC * 4;                        // 128 or 64: the swizzle span
constexpr int CG_PIECES = CG_ROW_BYTES / 16;                   // 16-byte pieces per row: 8 or 4
constexpr int CG_A_PLANE = CG_BM * CG_KC * 4;  // 16 / 8 KB
constexpr int CG_B_PLANE = CG_BN * CG_KC * 4;  // 32 / 16 KB
constexpr int CG_STAGE_BYTES = 2 * CG_A_PLANE + 2 * CG_B_PLANE;  // 96 / 48 KB: [A hi | A lo | B hi | B lo]
constexpr int CG_TX_BYTES = CG_A_PLANE + CG_B_PLANE;             // what TMA brings per stage: the two raw tiles
constexpr int CG_THREADS = 512;                                  // 4 warpgroups
constexpr int CG_XF_WARPS = 6;                                   // transform warps 2..7
constexpr int CG_EPI_WARPS = 8;                                  // epilogue warps 8..15
constexpr int CG_EPI_BYTES = 4096;                               // per epilogue warp: two [32 x 16] fp32 staging slots
constexpr int CG_SMEM_TOTAL = 1024 + CG_STAGES * CG_STAGE_BYTES + CG_EPI_WARPS * CG_EPI_BYTES;

struct ChainGemmArgs {
    float* out;              // [n_out_elems][S][S] fp32, orientation per element
    unsigned* zmax_out;      // [n_out_elems] bits of the largest stored entry (zeroed before the launch)
    int* E_out;              // [n_out_elems]
    const unsigned* zmax_a;  // scaling state of the operands: nullptr = (1.0, 0)
    const int* E_a;
    const unsigned* zmax_b;
    const int* E_b;
    const float* kscale;     // nullable: A[:, k] *= kscale[p * scale_stride + k]   (leaf level: diag(exp(obs_t0)) between the two P)
    const float* colscale;   // nullable: out[:, j] *= colscale[p * scale_stride + j] (leaf level: diag(exp(obs_t1)))
    int64_t scale_stride;
    int S;
    int a_shared;            // 1: every product uses element 0 of the A tensor map
    int a_stride, a_off;     // A element of pair q of chain ch = ch * a_cs + a_stride * q + a_off
    int b_stride, b_off;
    int last_transposed;     // orientation of the LAST output element of a chain (others: parity of the output index)
    int n_out;               // products in this launch (all chains)
    int ppc;                 // products per chain
    int a_cs, b_cs, o_cs;    // elements per chain in the A / B operand arrays and in the output array
    float comp;
    long long* prof;         // debugging aid (FB2_CG_PROF=1): clock64 stamps of CTA 0
    int mode;                // FB2_CG_MODE ablations: 1 = no TMA loads after the first two chunks, 2 = one MMA per tile, 4 = no stores
};

__device__ __forceinline__ uint32_t cg_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t cg_desc(uint32_t addr) {  // K-major; 128-byte (64-byte) swizzle, 8-row groups 1024 (512) B apart
    constexpr uint64_t layout = CG_KC == 32 ? 2ull : 4ull;  // cute::UMMA::LayoutType SWIZZLE_128B / SWIZZLE_64B
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)((8u * CG_ROW_BYTES) >> 4) << 32) | (1ull << 46) | (layout << 61);
}
__device__ __forceinline__ void cg_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void cg_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void cg_tma_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void cg_tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void cg_tmem_ld16(uint32_t taddr, uint32_t (&d)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(d[8]), "=r"(d[9]),
          "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ float cg_trunc_tf32(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }
__device__ __forceinline__ float4 cg_lo4(const float4 v) {
    return make_float4(v.x - cg_trunc_tf32(v.x), v.y - cg_trunc_tf32(v.y), v.z - cg_trunc_tf32(v.z), v.w - cg_trunc_tf32(v.w));
}

__global__ void __launch_bounds__(CG_THREADS, 1) chain_gemm_tc(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                                               const __grid_constant__ CUtensorMap map_o, const __grid_constant__ CUtensorMap map_ot,
                                                               ChainGemmArgs a) {
    extern __shared__ unsigned char cg_smem_raw[];
    __shared__ __align__(8) unsigned long long s_full[CG_STAGES], s_ready[CG_STAGES], s_empty[CG_STAGES], s_tfull, s_tempty;
    __shared__ uint32_t s_tmem;
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(cg_smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int S = a.S;
    const int tiles_n = (S + CG_BN - 1) / CG_BN;
    const int tiles = (S / CG_BM) * tiles_n;
    const int64_t total = (int64_t)tiles * a.n_out;
    const int nchunks = S / CG_KC;
    const bool do_prof = a.prof != nullptr && blockIdx.x == 0;
    if (do_prof && tid == 0) a.prof[0] = clock64();

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < CG_STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(cg_u32(&s_full[s])) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(cg_u32(&s_ready[s])), "n"(CG_XF_WARPS) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(cg_u32(&s_empty[s])) : "memory");
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(cg_u32(&s_tfull)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(cg_u32(&s_tempty)), "n"(CG_EPI_WARPS) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_o) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_ot) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(cg_u32(&s_tmem)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    if (do_prof && tid == 0) a.prof[1] = clock64();

    if (warp < 8) {
        // warpgroups 0 and 1 (producer, MMA issuer, transform) need few registers: give the rest to the epilogue warpgroups
        asm volatile("setmaxnreg.dec.sync.aligned.u32 88;" ::: "memory");
        if (warp == 0) {
            // ---------------- producer: one lane, runs ahead of the MMA warp by the depth of the ring (also across tiles)
            if (lane == 0) {
                int64_t gc = 0;  // chunks issued by this CTA so far
                for (int64_t t = blockIdx.x; t < total; t += gridDim.x) {
                    const int p = (int)(t / tiles), tile = (int)(t % tiles);
                    const int i0 = (tile / tiles_n) * CG_BM, j0 = (tile % tiles_n) * CG_BN;
                    const int ch = p / a.ppc, q = p % a.ppc;
                    const int ia = a.a_shared ? 0 : ch * a.a_cs + a.a_stride * q + a.a_off;
                    const int ib = ch * a.b_cs + a.b_stride * q + a.b_off;
                    for (int c = 0; c < nchunks; ++c, ++gc) {
                        const int s = (int)(gc % CG_STAGES);
                        if (gc >= CG_STAGES) cg_wait(cg_u32(&s_empty[s]), (uint32_t)(((gc / CG_STAGES) - 1) & 1));
                        const uint32_t bar = cg_u32(&s_full[s]);
                        if ((a.mode & 1) && gc >= CG_STAGES) {  // ablation: no data movement, just hand the stage on
                            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
                            continue;
                        }
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)CG_TX_BYTES) : "memory");
                        const uint32_t st = cg_u32(smem + (size_t)s * CG_STAGE_BYTES);
                        cg_tma_3d(st, &map_a, c * CG_KC, i0, ia, bar);
                        cg_tma_3d(st + 2 * CG_A_PLANE, &map_b, c * CG_KC, j0, ib, bar);
                    }
                }
            }
        } else if (warp == 1) {
            // ---------------- MMA issuer
            constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(CG_BN >> 3) << 17) | ((128u >> 4) << 24);
            int64_t gc = 0;
            int it = 0;
            for (int64_t t = blockIdx.x; t < total; t += gridDim.x, ++it) {
                if (it > 0) {  // the epilogue of the previous tile must have read the accumulators out of tensor memory
                    cg_wait(cg_u32(&s_tempty), (uint32_t)((it - 1) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                for (int c = 0; c < nchunks; ++c, ++gc) {
                    const int s = (int)(gc % CG_STAGES);
                    cg_wait(cg_u32(&s_ready[s]), (uint32_t)((gc / CG_STAGES) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (do_prof && lane == 0 && it < 2 && (c == 0 || c == nchunks - 1)) a.prof[2 + 2 * it + (c ? 1 : 0)] = clock64();
                    if (lane == 0) {
                        const uint32_t ah = cg_u32(smem + (size_t)s * CG_STAGE_BYTES), al = ah + CG_A_PLANE, bh = ah + 2 * CG_A_PLANE, bl = bh + CG_B_PLANE;
#pragma unroll
                        for (int ks = 0; ks < CG_KC / 8; ++ks) {
                            if ((a.mode & 2) && (c > 0 || ks > 0)) break;  // ablation: one MMA per tile
                            const uint64_t dah = cg_desc(ah + ks * 32), dal = cg_desc(al + ks * 32);
                            const uint64_t dbh = cg_desc(bh + ks * 32), dbl = cg_desc(bl + ks * 32);
                            cg_mma(tmem, dah, dbh, IDESC, (c > 0 || ks > 0) ? 1u : 0u);
                            cg_mma(tmem + CG_BN, dal, dbh, IDESC, (c > 0 || ks > 0) ? 1u : 0u);
                            cg_mma(tmem + CG_BN, dah, dbl, IDESC, 1u);
                        }
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(cg_u32(&s_empty[s])) : "memory");
                        if (c == nchunks - 1)
                            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(cg_u32(&s_tfull)) : "memory");
                    }
                    __syncwarp();
                }
            }
        } else {
            // ---------------- transform: lo = v - trunc_tf32(v) for the stage's raw tiles (A: 1024 16-byte pieces, B: 2048)
            const int xt = tid - 64;  // 0 .. 191
            int64_t gc = 0;
            for (int64_t t = blockIdx.x; t < total; t += gridDim.x) {
                const int p = (int)(t / tiles);
                const float* ksc = a.kscale ? a.kscale + (size_t)p * a.scale_stride : nullptr;
                for (int c = 0; c < nchunks; ++c, ++gc) {
                    const int s = (int)(gc % CG_STAGES);
                    cg_wait(cg_u32(&s_full[s]), (uint32_t)((gc / CG_STAGES) & 1));
                    unsigned char* st = smem + (size_t)s * CG_STAGE_BYTES;
                    if (ksc) {
                        // leaf level: the left operand is P diag(w): scale column k of the A tile.  The 16-byte piece at position j' of
                        // row r holds the logical k-chunk j' ^ (r % 8) of the 128-byte swizzle.
                        for (int ci = xt; ci < CG_A_PLANE / 16; ci += CG_XF_WARPS * 32) {
                            // logical k-piece of the physical piece: 128-byte swizzle XORs with (row % 8), 64-byte swizzle with ((row / 2) % 4)
                            const int r = ci / CG_PIECES;
                            const int jl = (ci % CG_PIECES) ^ (CG_KC == 32 ? (r & 7) : ((r >> 1) & 3));
                            const float4 w = __ldg(reinterpret_cast<const float4*>(ksc + c * CG_KC + 4 * jl));
                            float4 v = *reinterpret_cast<float4*>(st + ci * 16);
                            v.x *= w.x; v.y *= w.y; v.z *= w.z; v.w *= w.w;
                            *reinterpret_cast<float4*>(st + ci * 16) = v;
                            *reinterpret_cast<float4*>(st + CG_A_PLANE + ci * 16) = cg_lo4(v);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < (CG_A_PLANE / 16 + CG_XF_WARPS * 32 - 1) / (CG_XF_WARPS * 32); ++i) {  // all loads of the unrolled body issue first
                            const int ci = xt + i * CG_XF_WARPS * 32;
                            if (ci < CG_A_PLANE / 16) *reinterpret_cast<float4*>(st + CG_A_PLANE + ci * 16) = cg_lo4(*reinterpret_cast<const float4*>(st + ci * 16));
                        }
                    }
                    unsigned char* sb = st + 2 * CG_A_PLANE;
                    {
                        constexpr int NB = (CG_B_PLANE / 16 + CG_XF_WARPS * 32 - 1) / (CG_XF_WARPS * 32);
                        float4 vb[NB];
#pragma unroll
                        for (int i = 0; i < NB; ++i) {
                            const int ci = xt + i * CG_XF_WARPS * 32;
                            if (ci < CG_B_PLANE / 16) vb[i] = *reinterpret_cast<const float4*>(sb + ci * 16);
                        }
#pragma unroll
                        for (int i = 0; i < NB; ++i) {
                            const int ci = xt + i * CG_XF_WARPS * 32;
                            if (ci < CG_B_PLANE / 16) *reinterpret_cast<float4*>(sb + CG_B_PLANE + ci * 16) = cg_lo4(vb[i]);
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core's reads
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(cg_u32(&s_ready[s])) : "memory");
                }
            }
        }
    } else {
        // ---------------- epilogue: warp w owns tensor-memory lanes 32 (w % 4) .. and one half of the tile's columns
        asm volatile("setmaxnreg.inc.sync.aligned.u32 168;" ::: "memory");
        const int q4 = warp & 3;
        const int chalf = (warp - 8) >> 2;
        unsigned char* stg = smem + (size_t)CG_STAGES * CG_STAGE_BYTES + (size_t)(warp - 8) * CG_EPI_BYTES;
        const uint32_t stg_u32 = cg_u32(stg);
        const float fmain = 1.f + a.comp * (float)(nchunks * (CG_KC / 8));
        const uint32_t lane_base = tmem + ((uint32_t)(q4 * 32) << 16);
        int it = 0;
        int slot = 0;
        for (int64_t t = blockIdx.x; t < total; t += gridDim.x, ++it) {
            const int p = (int)(t / tiles), tile = (int)(t % tiles);
            const int i0 = (tile / tiles_n) * CG_BM, j0 = (tile % tiles_n) * CG_BN;
            const int ch = p / a.ppc, pq = p % a.ppc;
            const int ia = a.a_shared ? 0 : ch * a.a_cs + a.a_stride * pq + a.a_off;
            const int ib = ch * a.b_cs + a.b_stride * pq + a.b_off;
            const int po = ch * a.o_cs + pq;  // output element
            // scaling state of the operands -> exact power-of-two renormalisation
            const float za = a.zmax_a ? __uint_as_float(a.zmax_a[ia]) : 1.f;
            const float zb = a.zmax_b ? __uint_as_float(a.zmax_b[ib]) : 1.f;
            int ea = 0, eb = 0;
            if (za > 0.f) frexpf(za, &ea);
            if (zb > 0.f) frexpf(zb, &eb);
            const float scale = ldexpf(1.f, -(ea + eb));
            const bool transposed = pq == a.ppc - 1 ? a.last_transposed != 0 : (pq & 1) != 0;
            if (tile == 0 && q4 == 0 && chalf == 0 && lane == 0)
                a.E_out[po] = (a.E_a ? a.E_a[ia] : 0) + (a.E_b ? a.E_b[ib] : 0) + ea + eb;
            const float* cs = a.colscale ? a.colscale + (size_t)p * a.scale_stride : nullptr;
            cg_wait(cg_u32(&s_tfull), (uint32_t)(it & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (do_prof && tid == 256 && it < 2) a.prof[8 + 2 * it] = clock64();
            // Phase A: this warp's 32 x 128 block of the accumulators -> registers; tensor memory goes back to the MMA warp at once
            float vmax = 0.f;
            float v[128];
            const int cbase = chalf * (CG_BN / 2);
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                const int cc = cbase + g * 16;
                if (j0 + cc < S) {  // warp-uniform
                    uint32_t r0[16], r1[16];
                    cg_tmem_ld16(lane_base + (uint32_t)cc, r0);
                    cg_tmem_ld16(lane_base + (uint32_t)(CG_BN + cc), r1);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        float z = fmaf(__uint_as_float(r0[q]), fmain, __uint_as_float(r1[q])) * scale;
                        if (cs) z *= __ldg(cs + j0 + cc + q);
                        v[g * 16 + q] = fmaxf(z, 0.f);
                        vmax = fmaxf(vmax, v[g * 16 + q]);
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 16; ++q) v[g * 16 + q] = 0.f;
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(cg_u32(&s_tempty)) : "memory");
            if (do_prof && tid == 256 && it < 2) a.prof[13 + it] = clock64();
            // Phase B: stores, one 32 x 16 sub-block at a time through the warp's two staging slots
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                const int cc = cbase + g * 16;
                if (j0 + cc >= S || (a.mode & 4)) continue;  // warp-uniform
                // the slot written two sub-blocks ago may still be read by its TMA store
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncwarp();
                float* sg = reinterpret_cast<float*>(stg + slot * 2048);
                if (!transposed) {  // [32 rows][16 cols], row = lane
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        *reinterpret_cast<float4*>(sg + lane * 16 + 4 * q) = make_float4(v[g * 16 + 4 * q], v[g * 16 + 4 * q + 1], v[g * 16 + 4 * q + 2], v[g * 16 + 4 * q + 3]);
                } else {  // out^T: [16 cols][32 rows], the warp writes 32 consecutive floats per column
#pragma unroll
                    for (int q = 0; q < 16; ++q) sg[q * 32 + lane] = v[g * 16 + q];
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    if (!transposed) cg_tma_store_3d(&map_o, stg_u32 + slot * 2048, j0 + cc, i0 + q4 * 32, po);
                    else cg_tma_store_3d(&map_ot, stg_u32 + slot * 2048, i0 + q4 * 32, j0 + cc, po);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                slot ^= 1;
            }
            vmax = warp_max(vmax);
            if (lane == 0 && vmax > 0.f) atomicMax(a.zmax_out + po, __float_as_uint(vmax));
            if (do_prof && tid == 256 && it < 2) a.prof[9 + 2 * it] = clock64();
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
    if (do_prof && tid == 32) a.prof[12] = clock64();
}

// ------------------------------------------------------------------------------------------------ CTA-pair variant (cta_group::2)
// Same pipeline on a pair of CTAs (one thread-block cluster of 2 = the two SMs of a TPC) computing a 256 x 256 tile with
// tcgen05.mma.cta_group::2: each CTA stages its 128 rows of A and HALF of B (128 of the tile's 256 columns), the leader CTA issues the
// M = 256 instructions, each tensor core reads its own A and both halves of B, each CTA's tensor memory receives its 128 rows of the
// result.  What it buys: the 1-CTA kernel is bound by shared-memory bandwidth (288 KB per 32-wide K chunk: see above); here a CTA moves
// 32 KB by TMA, 64 KB through the transform warps and 96 KB into the tensor core per chunk (192 KB), and the stage shrinks to 64 KB,
// so three stages fit.
constexpr int CG2_STAGES = 3;
constexpr int CG2_KC = 32;
constexpr int CG2_A_PLANE = 128 * CG2_KC * 4;                 // 16 KB
constexpr int CG2_B_PLANE = 128 * CG2_KC * 4;                 // 16 KB: this CTA's half of the B tile
constexpr int CG2_STAGE_BYTES = 2 * CG2_A_PLANE + 2 * CG2_B_PLANE;  // 64 KB: [A hi | A lo | B hi | B lo]
constexpr int CG2_TX_BYTES = CG2_A_PLANE + CG2_B_PLANE;
constexpr int CG2_SMEM_TOTAL = 1024 + CG2_STAGES * CG2_STAGE_BYTES + CG_EPI_WARPS * CG_EPI_BYTES;

__device__ __forceinline__ uint64_t cg2_desc(uint32_t addr) {  // K-major, 128-byte swizzle, 8-row groups 1024 B apart
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void cg2_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// bf16 side planes: K-major rows of 32 bf16 = 64 bytes, 64-byte swizzle, 8-row groups 512 B apart
__device__ __forceinline__ uint64_t cg2_desc16(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(512u >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
__device__ __forceinline__ void cg2_mma16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ uint32_t cg_pack_bf16(float lo, float hi) {  // {upper half: hi, lower half: lo}, round to nearest
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// One 16-byte piece (4 fp32 = k 4 jl .. 4 jl + 3 of row r) of a raw tile -> 8 bytes in each bf16 side plane (v and lo = v - trunc_tf32(v)):
// rows of 64 bytes, 16-byte pieces xor-swizzled with (r >> 1) & 3.
__device__ __forceinline__ void cg2_side16(unsigned char* plane_v, unsigned char* plane_lo, int r, int jl, const float4 v) {
    const float4 lo = cg_lo4(v);
    const int off = r * 64 + (((jl >> 1) ^ ((r >> 1) & 3)) << 4) + ((jl & 1) << 3);
    *reinterpret_cast<uint2*>(plane_v + off) = make_uint2(cg_pack_bf16(v.x, v.y), cg_pack_bf16(v.z, v.w));
    *reinterpret_cast<uint2*>(plane_lo + off) = make_uint2(cg_pack_bf16(lo.x, lo.y), cg_pack_bf16(lo.z, lo.w));
}
__device__ __forceinline__ void cg2_wait_cluster(uint32_t bar, uint32_t parity) {  // arrivals may come from the peer CTA
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void cg2_arrive_on_leader(uint32_t local_bar) {  // arrive on the barrier at the same offset in CTA 0 of the pair
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(remote) : "r"(local_bar));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void cg2_commit_both(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void cg2_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CG_THREADS, 1)
    chain_gemm_tc2(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_o,
                   const __grid_constant__ CUtensorMap map_ot, ChainGemmArgs a) {
    extern __shared__ unsigned char cg_smem_raw[];
    __shared__ __align__(8) unsigned long long s_full[CG2_STAGES], s_ready[CG2_STAGES], s_empty[CG2_STAGES], s_tfull, s_tempty;
    __shared__ uint32_t s_tmem;
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(cg_smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint32_t cta_rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
    const bool leader = cta_rank == 0;
    const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
    const int S = a.S;
    const int tiles_n = S / 256;
    const int tiles = (S / 256) * tiles_n;  // 256 x 256 pair-tiles per product
    const int64_t total = (int64_t)tiles * a.n_out;
    const int nchunks = S / CG2_KC;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < CG2_STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(cg_u32(&s_full[s])) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(cg_u32(&s_ready[s])), "n"(2 * CG_XF_WARPS) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(cg_u32(&s_empty[s])) : "memory");
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(cg_u32(&s_tfull)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(cg_u32(&s_tempty)), "n"(2 * CG_EPI_WARPS) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_o) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_ot) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(cg_u32(&s_tmem)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cg2_cluster_sync();  // the peer's barriers are initialised before anybody arrives on them remotely
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    if (warp < 8) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 88;" ::: "memory");
        if (warp == 0) {
            // ---------------- producer (both CTAs): this CTA's 128 rows of A and its 128-column half of B
            if (lane == 0) {
                int64_t gc = 0;
                for (int64_t t = cluster_id; t < total; t += nclusters) {
                    const int p = (int)(t / tiles), tile = (int)(t % tiles);
                    const int i0 = (tile / tiles_n) * 256 + 128 * (int)cta_rank, jb = (tile % tiles_n) * 256 + 128 * (int)cta_rank;
                    const int ch = p / a.ppc, q = p % a.ppc;
                    const int ia = a.a_shared ? 0 : ch * a.a_cs + a.a_stride * q + a.a_off;
                    const int ib = ch * a.b_cs + a.b_stride * q + a.b_off;
                    for (int c = 0; c < nchunks; ++c, ++gc) {
                        const int s = (int)(gc % CG2_STAGES);
                        if (gc >= CG2_STAGES) cg_wait(cg_u32(&s_empty[s]), (uint32_t)(((gc / CG2_STAGES) - 1) & 1));
                        const uint32_t bar = cg_u32(&s_full[s]);
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)CG2_TX_BYTES) : "memory");
                        const uint32_t st = cg_u32(smem + (size_t)s * CG2_STAGE_BYTES);
                        cg_tma_3d(st, &map_a, c * CG2_KC, i0, ia, bar);
                        cg_tma_3d(st + 2 * CG2_A_PLANE, &map_b, c * CG2_KC, jb, ib, bar);
                    }
                }
            }
        } else if (warp == 1) {
            // ---------------- MMA issuer: the leader CTA only
            if (leader) {
                constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(256 >> 3) << 17) | ((256u >> 4) << 24);  // M = 256, N = 256
                int64_t gc = 0;
                int it = 0;
                for (int64_t t = cluster_id; t < total; t += nclusters, ++it) {
                    if (it > 0) {  // both CTAs' epilogues have read the previous tile's accumulators
                        cg2_wait_cluster(cg_u32(&s_tempty), (uint32_t)((it - 1) & 1));
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    }
                    for (int c = 0; c < nchunks; ++c, ++gc) {
                        const int s = (int)(gc % CG2_STAGES);
                        cg2_wait_cluster(cg_u32(&s_ready[s]), (uint32_t)((gc / CG2_STAGES) & 1));
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        if (lane == 0) {
                            const uint32_t ah = cg_u32(smem + (size_t)s * CG2_STAGE_BYTES), al = ah + CG2_A_PLANE, bh = ah + 2 * CG2_A_PLANE, bl = bh + CG2_B_PLANE;
                            if (!(a.mode & 8)) {  // default: 3xTF32 (lo planes in fp32)
#pragma unroll
                                for (int ks = 0; ks < CG2_KC / 8; ++ks) {
                                    const uint64_t dah = cg2_desc(ah + ks * 32), dal = cg2_desc(al + ks * 32);
                                    const uint64_t dbh = cg2_desc(bh + ks * 32), dbl = cg2_desc(bl + ks * 32);
                                    cg2_mma(tmem, dah, dbh, IDESC, (c > 0 || ks > 0) ? 1u : 0u);
                                    cg2_mma(tmem + CG_BN, dal, dbh, IDESC, (c > 0 || ks > 0) ? 1u : 0u);
                                    cg2_mma(tmem + CG_BN, dah, dbl, IDESC, 1u);
                                }
                            } else {
                                // FB2_CG_MODE=8 (measured, not the default): hi x hi in TF32 (4 instructions per chunk of 32 k), the two side
                                // products (2^-10 of the result) in bf16 with K = 16 per instruction: [A lo16 x B v16] + [A v16 x B lo16] -- 8 tensor
                                // instructions per chunk instead of 12.  Parity-green; alternating A/B: 1.58 vs 1.59 us per step at S = 512, 11.79 vs 12.07 at
                                // S = 1024: the loop is bound by the transform warps and shared-memory traffic, not by tensor issue.
                                constexpr uint32_t IDESC16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((256u >> 4) << 24);
                                const uint32_t av16 = al, al16 = al + CG2_A_PLANE / 2, bv16 = bl, bl16 = bl + CG2_B_PLANE / 2;
#pragma unroll
                                for (int ks = 0; ks < CG2_KC / 8; ++ks)
                                    cg2_mma(tmem, cg2_desc(ah + ks * 32), cg2_desc(bh + ks * 32), IDESC, (c > 0 || ks > 0) ? 1u : 0u);
#pragma unroll
                                for (int ks = 0; ks < CG2_KC / 16; ++ks) {
                                    cg2_mma16(tmem + CG_BN, cg2_desc16(al16 + ks * 32), cg2_desc16(bv16 + ks * 32), IDESC16, (c > 0 || ks > 0) ? 1u : 0u);
                                    cg2_mma16(tmem + CG_BN, cg2_desc16(av16 + ks * 32), cg2_desc16(bl16 + ks * 32), IDESC16, 1u);
                                }
                            }
                            cg2_commit_both(cg_u32(&s_empty[s]));
                            if (c == nchunks - 1) cg2_commit_both(cg_u32(&s_tfull));
                        }
                        __syncwarp();
                    }
                }
            }
        } else {
            // ---------------- transform (both CTAs): lo = v - trunc_tf32(v); arrive on the LEADER's ready barrier
            const int xt = tid - 64;
            int64_t gc = 0;
            for (int64_t t = cluster_id; t < total; t += nclusters) {
                const int p = (int)(t / tiles);
                const float* ksc = a.kscale ? a.kscale + (size_t)p * a.scale_stride : nullptr;
                for (int c = 0; c < nchunks; ++c, ++gc) {
                    const int s = (int)(gc % CG2_STAGES);
                    cg_wait(cg_u32(&s_full[s]), (uint32_t)((gc / CG2_STAGES) & 1));
                    unsigned char* st = smem + (size_t)s * CG2_STAGE_BYTES;
                    const bool side16 = (a.mode & 8) != 0;
                    if (ksc) {
                        for (int ci = xt; ci < CG2_A_PLANE / 16; ci += CG_XF_WARPS * 32) {
                            const int r = ci >> 3, jl = (ci & 7) ^ (r & 7);
                            const float4 w = __ldg(reinterpret_cast<const float4*>(ksc + c * CG2_KC + 4 * jl));
                            float4 v = *reinterpret_cast<float4*>(st + ci * 16);
                            v.x *= w.x; v.y *= w.y; v.z *= w.z; v.w *= w.w;
                            *reinterpret_cast<float4*>(st + ci * 16) = v;
                            if (side16) cg2_side16(st + CG2_A_PLANE, st + CG2_A_PLANE + CG2_A_PLANE / 2, r, jl, v);
                            else *reinterpret_cast<float4*>(st + CG2_A_PLANE + ci * 16) = cg_lo4(v);
                        }
                    } else if (side16) {
#pragma unroll
                        for (int i = 0; i < (CG2_A_PLANE / 16 + CG_XF_WARPS * 32 - 1) / (CG_XF_WARPS * 32); ++i) {
                            const int ci = xt + i * CG_XF_WARPS * 32;
                            if (ci < CG2_A_PLANE / 16) {
                                const int r = ci >> 3, jl = (ci & 7) ^ (r & 7);
                                cg2_side16(st + CG2_A_PLANE, st + CG2_A_PLANE + CG2_A_PLANE / 2, r, jl, *reinterpret_cast<const float4*>(st + ci * 16));
                            }
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < (CG2_A_PLANE / 16 + CG_XF_WARPS * 32 - 1) / (CG_XF_WARPS * 32); ++i) {
                            const int ci = xt + i * CG_XF_WARPS * 32;
                            if (ci < CG2_A_PLANE / 16) *reinterpret_cast<float4*>(st + CG2_A_PLANE + ci * 16) = cg_lo4(*reinterpret_cast<const float4*>(st + ci * 16));
                        }
                    }
                    unsigned char* sb = st + 2 * CG2_A_PLANE;
#pragma unroll
                    for (int i = 0; i < (CG2_B_PLANE / 16 + CG_XF_WARPS * 32 - 1) / (CG_XF_WARPS * 32); ++i) {
                        const int ci = xt + i * CG_XF_WARPS * 32;
                        if (ci < CG2_B_PLANE / 16) {
                            if (side16) {
                                const int r = ci >> 3, jl = (ci & 7) ^ (r & 7);
                                cg2_side16(sb + CG2_B_PLANE, sb + CG2_B_PLANE + CG2_B_PLANE / 2, r, jl, *reinterpret_cast<const float4*>(sb + ci * 16));
                            } else {
                                *reinterpret_cast<float4*>(sb + CG2_B_PLANE + ci * 16) = cg_lo4(*reinterpret_cast<const float4*>(sb + ci * 16));
                            }
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) cg2_arrive_on_leader(cg_u32(&s_ready[s]));
                }
            }
        }
    } else {
        // ---------------- epilogue (both CTAs): this CTA's 128 rows x 256 columns of the pair-tile
        asm volatile("setmaxnreg.inc.sync.aligned.u32 168;" ::: "memory");
        const int q4 = warp & 3;
        const int chalf = (warp - 8) >> 2;
        unsigned char* stg = smem + (size_t)CG2_STAGES * CG2_STAGE_BYTES + (size_t)(warp - 8) * CG_EPI_BYTES;
        const uint32_t stg_u32 = cg_u32(stg);
        const float fmain = 1.f + a.comp * (float)(nchunks * (CG2_KC / 8));
        const uint32_t lane_base = tmem + ((uint32_t)(q4 * 32) << 16);
        int it = 0;
        int slot = 0;
        for (int64_t t = cluster_id; t < total; t += nclusters, ++it) {
            const int p = (int)(t / tiles), tile = (int)(t % tiles);
            const int i0 = (tile / tiles_n) * 256 + 128 * (int)cta_rank, j0 = (tile % tiles_n) * 256;
            const int ch = p / a.ppc, pq = p % a.ppc;
            const int ia = a.a_shared ? 0 : ch * a.a_cs + a.a_stride * pq + a.a_off;
            const int ib = ch * a.b_cs + a.b_stride * pq + a.b_off;
            const int po = ch * a.o_cs + pq;
            const float za = a.zmax_a ? __uint_as_float(a.zmax_a[ia]) : 1.f;
            const float zb = a.zmax_b ? __uint_as_float(a.zmax_b[ib]) : 1.f;
            int ea = 0, eb = 0;
            if (za > 0.f) frexpf(za, &ea);
            if (zb > 0.f) frexpf(zb, &eb);
            const float scale = ldexpf(1.f, -(ea + eb));
            const bool transposed = pq == a.ppc - 1 ? a.last_transposed != 0 : (pq & 1) != 0;
            if (tile == 0 && leader && q4 == 0 && chalf == 0 && lane == 0)
                a.E_out[po] = (a.E_a ? a.E_a[ia] : 0) + (a.E_b ? a.E_b[ib] : 0) + ea + eb;
            const float* cs = a.colscale ? a.colscale + (size_t)p * a.scale_stride : nullptr;
            cg_wait(cg_u32(&s_tfull), (uint32_t)(it & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float vmax = 0.f;
            float v[128];
            const int cbase = chalf * 128;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                const int cc = cbase + g * 16;
                uint32_t r0[16], r1[16];
                cg_tmem_ld16(lane_base + (uint32_t)cc, r0);
                cg_tmem_ld16(lane_base + (uint32_t)(CG_BN + cc), r1);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    float z = fmaf(__uint_as_float(r0[q]), fmain, __uint_as_float(r1[q])) * scale;
                    if (cs) z *= __ldg(cs + j0 + cc + q);
                    v[g * 16 + q] = fmaxf(z, 0.f);
                    vmax = fmaxf(vmax, v[g * 16 + q]);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) cg2_arrive_on_leader(cg_u32(&s_tempty));
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                const int cc = cbase + g * 16;
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncwarp();
                float* sg = reinterpret_cast<float*>(stg + slot * 2048);
                if (!transposed) {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        *reinterpret_cast<float4*>(sg + lane * 16 + 4 * q) = make_float4(v[g * 16 + 4 * q], v[g * 16 + 4 * q + 1], v[g * 16 + 4 * q + 2], v[g * 16 + 4 * q + 3]);
                } else {
#pragma unroll
                    for (int q = 0; q < 16; ++q) sg[q * 32 + lane] = v[g * 16 + q];
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    if (!transposed) cg_tma_store_3d(&map_o, stg_u32 + slot * 2048, j0 + cc, i0 + q4 * 32, po);
                    else cg_tma_store_3d(&map_ot, stg_u32 + slot * 2048, i0 + q4 * 32, j0 + cc, po);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                slot ^= 1;
            }
            vmax = warp_max(vmax);
            if (lane == 0 && vmax > 0.f) atomicMax(a.zmax_out + po, __float_as_uint(vmax));
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cg2_cluster_sync();  // neither CTA may free tensor memory or exit while the pair's instructions can still touch it
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}

// ------------------------------------------------------------------------------------------------ leaf preparation and conversions
// role (0 = left / row-major, 1 = right / transposed) of the element at index `idx` of a level with `size` elements, following the
// carry chain of the odd leftover (sum_product.py:696-698); the final result of the tree is row-major
__host__ __device__ static int cg_role(int64_t idx, int64_t size) {
    while (size > 1) {
        if (idx == size - 1 && (size & 1)) {
            idx = size / 2;
            size = (size + 1) / 2;
            continue;
        }
        return (int)(idx & 1);
    }
    return 0;
}

// w[e][j] = exp(obs[e][j] - max_j obs[e][j]), shift[e] = that max (an impossible observation row, all -inf, gives w = 0, shift = 0)
__global__ void cg_obs_weights(const float* __restrict__ obs, int S, float* __restrict__ w, float* __restrict__ shift_out) {
    __shared__ float s_red[32];
    const int64_t e = blockIdx.x;
    const float* o = obs + e * S;
    float m = -FLT_MAX;
    for (int k = threadIdx.x; k < S; k += blockDim.x) m = fmaxf(m, o[k]);
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
    __syncthreads();
    m = -FLT_MAX;
    for (int q = 0; q < (int)(blockDim.x >> 5); ++q) m = fmaxf(m, s_red[q]);
    if (!(m > -FLT_MAX)) m = 0.f;
    for (int k = threadIdx.x; k < S; k += blockDim.x) w[e * S + k] = expf(o[k] - m);
    if (threadIdx.x == 0) shift_out[e] = m;
}
// P = exp(A) row-major and P^T (the shared operands of every leaf product); A row-major (prev, curr)
__global__ void cg_exp_both(const float* __restrict__ A, int S, float* __restrict__ P, float* __restrict__ PT) {
    __shared__ float tile[32][33];
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) {
        const float v = expf(A[(size_t)(r0 + rr) * S + c0 + threadIdx.x]);
        tile[rr][threadIdx.x] = v;
        P[(size_t)(r0 + rr) * S + c0 + threadIdx.x] = v;
    }
    __syncthreads();
    for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) PT[(size_t)(c0 + rr) * S + r0 + threadIdx.x] = tile[threadIdx.x][rr];
}
// a single leaf M = P diag(w) in the orientation `transposed` (0: M[i][j] = P[i][j] w[j]; 1: M^T[j][i] = P^T[j][i] w[j]); state (1.0, 0)
__global__ void cg_single_leaf(const float* __restrict__ P, const float* __restrict__ PT, const float* __restrict__ w, int S, int transposed, float* __restrict__ out,
                               unsigned* __restrict__ zmax, int* __restrict__ E) {
    const size_t SS = (size_t)S * S;
    const float* base = transposed ? PT : P;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < SS; idx += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(idx / S), c = (int)(idx % S);
        out[idx] = base[idx] * (transposed ? w[r] : w[c]);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *zmax = 0x3f800000u;
        *E = 0;
    }
}
// fp32 matrices (row-major) -> the orientation of their role among n elements (tree results re-entering a tree as elements)
__global__ void cg_orient(const float* __restrict__ M, int S, int64_t n, float* __restrict__ out) {
    __shared__ float tile[32][33];
    const int64_t e = blockIdx.z;
    const size_t SS = (size_t)S * S;
    const bool tr = cg_role(e, n) != 0;
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const float* m = M + (size_t)e * SS;
    float* o = out + (size_t)e * SS;
    for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) tile[rr][threadIdx.x] = m[(size_t)(r0 + rr) * S + c0 + threadIdx.x];
    __syncthreads();
    for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) {
        if (tr) o[(size_t)(c0 + rr) * S + r0 + threadIdx.x] = tile[threadIdx.x][rr];
        else o[(size_t)(r0 + rr) * S + c0 + threadIdx.x] = tile[rr][threadIdx.x];
    }
}
// rows[i][j] = log(M[i][j]) - log(zmax),  offset = E ln2 + log(zmax) + extra   (M = the final matrix of a tree, row-major)
__global__ void cg_finish_rows(const float* __restrict__ M, const unsigned* __restrict__ zmax, const int* __restrict__ E, const double* __restrict__ extra, int S,
                               float* __restrict__ rows, double* __restrict__ row_offset) {
    const float zm = __uint_as_float(zmax[0]);
    const size_t SS = (size_t)S * S;
    const float lz = zm > 0.f ? logf(zm) : 0.f;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < SS; idx += (size_t)gridDim.x * blockDim.x) {
        const float v = M[idx];
        rows[idx] = v > 0.f ? logf(v) - lz : -INFINITY;
    }
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < S; i += blockDim.x)
            row_offset[i] = zm > 0.f ? (double)E[0] * 0.69314718055994530942 + (double)lz + extra[0] : -INFINITY;
}
__global__ void cg_sum_shifts(const float* __restrict__ shift, int64_t n, double* __restrict__ acc) {
    __shared__ double s_red[32];
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s += (double)shift[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int q = 0; q < (int)(blockDim.x >> 5); ++q) t += s_red[q];
        atomicAdd(acc, t);
    }
}
// carry the odd leftover of every chain to the end of the next level (already in the orientation of its eventual role)
__global__ void cg_copy_leftover(const float* __restrict__ src, const unsigned* __restrict__ zsrc, const int* __restrict__ esrc, int64_t n, int64_t nd, size_t SS,
                                 float* __restrict__ dst, unsigned* __restrict__ zdst, int* __restrict__ edst) {
    const int64_t ch = blockIdx.y;
    const float* s_ = src + (size_t)(ch * n + n - 1) * SS;
    float* d_ = dst + (size_t)(ch * nd + nd - 1) * SS;
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < SS; i += (size_t)gridDim.x * blockDim.x * 4)
        *reinterpret_cast<float4*>(d_ + i) = *reinterpret_cast<const float4*>(s_ + i);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        zdst[ch * nd + nd - 1] = zsrc ? zsrc[ch * n + n - 1] : 0x3f800000u;  // leaves: largest stored entry 1.0, E = 0
        edst[ch * nd + nd - 1] = esrc ? esrc[ch * n + n - 1] : 0;
    }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn cg_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}
// [n][S][S] fp32 as a 3-D tensor (inner, row, element) with an (inner_box, row_box, 1) box
static int cg_make_map(CUtensorMap* map, const float* base, int64_t n_elems, int S, int box_inner, int box_rows, bool swizzle128) {
    EncodeTiledFn enc = cg_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return FB2_ERR_UNSUPPORTED;
    }
    cuuint64_t dims[3] = {(cuuint64_t)S, (cuuint64_t)S, (cuuint64_t)n_elems};
    cuuint64_t strides[2] = {(cuuint64_t)S * 4, (cuuint64_t)S * S * 4};
    cuuint32_t box[3] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle128 ? (CG_KC == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B) : CU_TENSOR_MAP_SWIZZLE_NONE,
                     swizzle128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d)", (int)r);
        return FB2_ERR_CUDA;
    }
    return FB2_OK;
}

bool chain_gemm_supported(int S) { return S >= 256 && S <= 1024 && S % 128 == 0 && cg_encode_fn() != nullptr; }

// One level.  n_a / n_b / n_o = elements behind the A / B / output base pointers.
static int cg_launch(const float* a_base, int64_t n_a, const float* b_base, int64_t n_b, int64_t n_o, const ChainGemmArgs& args, cudaStream_t stream) {
    CUtensorMap ma, mb, mo, mot;
    int st = cg_make_map(&ma, a_base, n_a, args.S, CG_KC, CG_BM, true);
    if (st != FB2_OK) return st;
    st = cg_make_map(&mb, b_base, n_b, args.S, CG_KC, CG_BN, true);
    if (st != FB2_OK) return st;
    st = cg_make_map(&mo, args.out, n_o, args.S, 16, 32, false);
    if (st != FB2_OK) return st;
    st = cg_make_map(&mot, args.out, n_o, args.S, 32, 16, false);
    if (st != FB2_OK) return st;
    FB2_CUDA(cudaFuncSetAttribute(chain_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, CG_SMEM_TOTAL));
    ChainGemmArgs margs = args;
    {
        const char* m = getenv("FB2_CG_MODE");
        margs.mode = m ? atoi(m) : 0;
    }
    static long long* d_prof = nullptr;
    const char* pe = getenv("FB2_CG_PROF");
    if (pe && pe[0] == '1') {
        if (!d_prof) FB2_CUDA(cudaMalloc(&d_prof, 32 * sizeof(long long)));
        FB2_CUDA(cudaMemsetAsync(d_prof, 0, 32 * sizeof(long long), stream));
        margs.prof = d_prof;
    }
    const int tiles = (args.S / CG_BM) * ((args.S + CG_BN - 1) / CG_BN);
    FB2_REQUIRE(args.n_out >= 1 && args.ppc >= 1, "chain_gemm: empty level");
    static int sm_count = 0;
    if (!sm_count) {
        int dev = 0;
        FB2_CUDA(cudaGetDevice(&dev));
        FB2_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    const int64_t total = (int64_t)tiles * args.n_out;
    const char* two_env = getenv("FB2_CG_2CTA");
    const bool two = CG_KC == 32 && args.S % 256 == 0 && !(two_env && two_env[0] == '0') && margs.prof == nullptr && (margs.mode & ~8) == 0;
    if (two) {
        // CTA pairs on 256 x 256 tiles (tcgen05 cta_group::2): the B operand needs a 128-row box
        st = cg_make_map(&mb, b_base, n_b, args.S, CG_KC, 128, true);
        if (st != FB2_OK) return st;
        FB2_CUDA(cudaFuncSetAttribute(chain_gemm_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, CG2_SMEM_TOTAL));
        const int64_t total2 = (int64_t)(args.S / 256) * (args.S / 256) * args.n_out;
        const int64_t clusters = total2 < sm_count / 2 ? total2 : sm_count / 2;
        chain_gemm_tc2<<<(unsigned)(2 * clusters), CG_THREADS, CG2_SMEM_TOTAL, stream>>>(ma, mb, mo, mot, margs);
        FB2_LAUNCH_CHECK();
        return FB2_OK;
    }
    const unsigned grid = (unsigned)(total < sm_count ? total : sm_count);
    chain_gemm_tc<<<grid, CG_THREADS, CG_SMEM_TOTAL, stream>>>(ma, mb, mo, mot, margs);
    FB2_LAUNCH_CHECK();
    if (margs.prof) {
        long long h[32];
        FB2_CUDA(cudaStreamSynchronize(stream));
        FB2_CUDA(cudaMemcpy(h, d_prof, sizeof(h), cudaMemcpyDeviceToHost));
        fprintf(stderr, "[cg prof] products %d cta 0: setup %lld | tile0 first chunk +%lld, last chunk +%lld | tile1 first chunk +%lld, last chunk +%lld | "
                        "epilogue0 start +%lld, tensor memory released after %lld, len %lld | epilogue1 start +%lld, released after %lld, len %lld | total %lld (cycles)\n",
                args.n_out, h[1] - h[0], h[2] - h[1], h[3] - h[2], h[4] - h[3], h[5] - h[4], h[8] - h[0], h[13] - h[8], h[9] - h[8], h[10] - h[0],
                h[14] - h[10], h[11] - h[10], h[12] - h[0]);
    }
    return FB2_OK;
}

// Balanced tree (the reference's order) over `chains` independent chains of `n` elements each, laid out [chains][n][S][S] in `src`
// (element roles by cg_role), scaling state (zsrc, esrc) or nullptr for leaves.  Ping-pongs between buf[0] / buf[1]
// (>= chains * ceil(n/2) and chains * ceil(n/4) elements); the final products land row-major in final_out[chains][S][S].
struct CgTreeBuffers {
    float* buf[2];
    unsigned* zmax[2];  // each >= chains * ceil(n/2)
    int* E[2];
};
static int cg_tree(const float* src, const unsigned* zsrc, const int* esrc, int64_t chains, int64_t n, int S, const CgTreeBuffers& tb, float* final_out,
                   unsigned* final_zmax, int* final_E, float comp, cudaStream_t stream) {
    const size_t SS = (size_t)S * S;
    int which = 0;
    while (n > 1) {
        const int64_t pairs = n / 2, nd = (n + 1) / 2;
        const bool is_final = nd == 1;
        ChainGemmArgs a{};
        a.S = S; a.comp = comp;
        a.a_stride = 2; a.a_off = 0; a.b_stride = 2; a.b_off = 1;
        a.zmax_a = zsrc; a.E_a = esrc; a.zmax_b = zsrc; a.E_b = esrc;
        a.ppc = (int)pairs; a.a_cs = (int)n; a.b_cs = (int)n; a.o_cs = (int)nd;
        a.n_out = (int)(chains * pairs);
        a.out = is_final ? final_out : tb.buf[which];
        a.zmax_out = is_final ? final_zmax : tb.zmax[which];
        a.E_out = is_final ? final_E : tb.E[which];
        a.last_transposed = cg_role(pairs - 1, nd);
        FB2_CUDA(cudaMemsetAsync(a.zmax_out, 0, (size_t)chains * nd * sizeof(unsigned), stream));
        int st = cg_launch(src, chains * n, src, chains * n, chains * nd, a, stream);
        if (st != FB2_OK) return st;
        if (n & 1) {
            cg_copy_leftover<<<dim3(32, (unsigned)chains), 256, 0, stream>>>(src, zsrc, esrc, n, nd, SS, a.out, a.zmax_out, a.E_out);
            FB2_LAUNCH_CHECK();
        }
        src = a.out; zsrc = a.zmax_out; esrc = a.E_out;
        n = nd;
        which ^= 1;
    }
    return FB2_OK;
}

// Blocks of `block_steps` (even) time steps; a single trailing step is absorbed by the last block, so every block has >= 2 steps.
static int64_t cg_block_count(int64_t T, int64_t bs) {
    int64_t n = 0;
    for (int64_t t = 0; t < T;) {
        int64_t nt = bs < T - t ? bs : T - t;
        if (T - (t + nt) == 1) nt += 1;
        t += nt;
        ++n;
    }
    return n;
}

size_t hmm_chunk_summary_gemm_workspace_bytes(int64_t T, int S, int64_t block_steps) {
    const size_t SS = (size_t)S * S;
    int64_t bs = block_steps < T ? block_steps : T;
    bs += bs & 1;
    const int64_t l1 = bs / 2 + 2;  // elements after the leaf level (+ leftover leaf)
    const int64_t nblocks = cg_block_count(T, bs);
    size_t w = 2 * align_up(SS * 4);                                                 // P, P^T
    w += align_up((size_t)(bs + 2) * S * 4) + align_up((size_t)(bs + 2) * 4);        // observation weights, shifts
    w += align_up((size_t)l1 * SS * 4) + align_up((size_t)(l1 / 2 + 1) * SS * 4) + align_up((size_t)(l1 / 4 + 2) * SS * 4);  // leaf outputs + ping-pong
    w += 6 * align_up((size_t)(l1 + 1) * 4);
    w += 2 * align_up((size_t)nblocks * SS * 4) + align_up((size_t)(nblocks / 2 + 1) * SS * 4) + align_up((size_t)(nblocks / 4 + 2) * SS * 4);
    w += 6 * align_up((size_t)(nblocks + 1) * 4);
    w += align_up(SS * 4) + 3 * 256;
    return w + 4096;
}

// out rows [S][S] fp32 (largest entry 0) + row_offset [S] float64 of  (x)_{t} (A + obs[t])  for ONE chain; obs [T][S] contiguous
int hmm_chunk_summary_gemm_device(const float* A, const float* obs, int64_t T, int S, int64_t block_steps, float* rows_out, double* row_offset,
                                  void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (!chain_gemm_supported(S) || T < 2) return FB2_ERR_UNSUPPORTED;
    const size_t SS = (size_t)S * S;
    int64_t bs = block_steps < T ? block_steps : T;
    bs += bs & 1;
    FB2_REQUIRE(bs >= 2 && bs <= (1 << 22), "hmm_chunk_summary (gemm): block of %lld steps out of range", (long long)bs);
    const int64_t l1 = bs / 2 + 2;
    const int64_t nblocks = cg_block_count(T, bs);
    Arena ar(workspace, workspace_bytes);
    float* P = ar.take<float>(SS);
    float* PT = ar.take<float>(SS);
    float* W = ar.take<float>((size_t)(bs + 2) * S);
    float* shift = ar.take<float>((size_t)bs + 2);
    float* L1 = ar.take<float>((size_t)l1 * SS);
    CgTreeBuffers tb{}, ub{};
    tb.buf[0] = ar.take<float>((size_t)(l1 / 2 + 1) * SS);
    tb.buf[1] = ar.take<float>((size_t)(l1 / 4 + 2) * SS);
    unsigned* z1 = ar.take<unsigned>((size_t)l1 + 1);
    int* e1 = ar.take<int>((size_t)l1 + 1);
    for (int q = 0; q < 2; ++q) {
        tb.zmax[q] = ar.take<unsigned>((size_t)l1 + 1);
        tb.E[q] = ar.take<int>((size_t)l1 + 1);
    }
    float* blockM = ar.take<float>((size_t)nblocks * SS);
    float* blockO = ar.take<float>((size_t)nblocks * SS);
    ub.buf[0] = ar.take<float>((size_t)(nblocks / 2 + 1) * SS);
    ub.buf[1] = ar.take<float>((size_t)(nblocks / 4 + 2) * SS);
    unsigned* block_zmax = ar.take<unsigned>((size_t)nblocks + 1);
    int* block_E = ar.take<int>((size_t)nblocks + 1);
    for (int q = 0; q < 2; ++q) {
        ub.zmax[q] = ar.take<unsigned>((size_t)nblocks + 1);
        ub.E[q] = ar.take<int>((size_t)nblocks + 1);
    }
    float* finalM = ar.take<float>(SS);
    unsigned* final_zmax = ar.take<unsigned>(64);
    int* final_E = ar.take<int>(64);
    double* shift_acc = ar.take<double>(32);
    if (!ar.ok()) {
        set_error("hmm_chunk_summary (gemm): workspace too small (%zu bytes)", workspace_bytes);
        return FB2_ERR_WORKSPACE;
    }
    const float comp = 5.82e-8f;
    FB2_CUDA(cudaMemsetAsync(shift_acc, 0, sizeof(double), stream));
    cg_exp_both<<<dim3((unsigned)(S / 32), (unsigned)(S / 32)), dim3(32, 8), 0, stream>>>(A, S, P, PT);
    FB2_LAUNCH_CHECK();
    int64_t t0 = 0;
    for (int64_t b = 0; b < nblocks; ++b) {
        int64_t nt = bs < T - t0 ? bs : T - t0;
        if (T - (t0 + nt) == 1) nt += 1;
        const int64_t pairs = nt / 2;
        const bool odd = nt & 1;
        float* dstM = nblocks == 1 ? finalM : blockM + (size_t)b * SS;
        unsigned* dstZ = nblocks == 1 ? final_zmax : block_zmax + b;
        int* dstE = nblocks == 1 ? final_E : block_E + b;
        // observation weights of every step of the block: w_t = exp(obs_t - max obs_t)
        cg_obs_weights<<<(unsigned)nt, 128, 0, stream>>>(obs + t0 * S, S, W, shift);
        FB2_LAUNCH_CHECK();
        cg_sum_shifts<<<32, 256, 0, stream>>>(shift, nt, shift_acc);
        FB2_LAUNCH_CHECK();
        // leaf level: product p = P diag(w_{2p}) P diag(w_{2p+1}): A = P scaled along K in shared memory, B = P^T, columns scaled in the epilogue
        const int64_t n1 = pairs + (odd ? 1 : 0);
        const bool leaf_final = n1 == 1;
        {
            ChainGemmArgs a{};
            a.S = S; a.comp = comp;
            a.a_shared = 1; a.b_stride = 0; a.b_off = 0;  // A = P, B = P^T for every product
            a.kscale = W; a.colscale = W + S; a.scale_stride = 2 * (int64_t)S;  // rows 2p and 2p + 1 of W
            a.ppc = (int)pairs; a.a_cs = 0; a.b_cs = 0; a.o_cs = (int)n1;
            a.n_out = (int)pairs;
            a.out = leaf_final ? dstM : L1;
            a.zmax_out = leaf_final ? dstZ : z1;
            a.E_out = leaf_final ? dstE : e1;
            a.last_transposed = cg_role(pairs - 1, n1);
            FB2_CUDA(cudaMemsetAsync(a.zmax_out, 0, (size_t)n1 * sizeof(unsigned), stream));
            int st = cg_launch(P, 1, PT, 1, n1, a, stream);
            if (st != FB2_OK) return st;
        }
        if (odd) {
            // the last step of an odd block is a leaf of its own, M = P D_t, in the orientation of its eventual role
            cg_single_leaf<<<296, 256, 0, stream>>>(P, PT, W + (size_t)(nt - 1) * S, S, cg_role(pairs, n1), L1 + (size_t)pairs * SS, z1 + pairs, e1 + pairs);
            FB2_LAUNCH_CHECK();
        }
        if (!leaf_final) {
            int st = cg_tree(L1, z1, e1, 1, n1, S, tb, dstM, dstZ, dstE, comp, stream);
            if (st != FB2_OK) return st;
        }
        t0 += nt;
    }
    if (nblocks > 1) {
        // the block results re-enter a tree as elements, in the orientation of their role
        cg_orient<<<dim3((unsigned)(S / 32), (unsigned)(S / 32), (unsigned)nblocks), dim3(32, 8), 0, stream>>>(blockM, S, nblocks, blockO);
        FB2_LAUNCH_CHECK();
        int st = cg_tree(blockO, block_zmax, block_E, 1, nblocks, S, ub, finalM, final_zmax, final_E, comp, stream);
        if (st != FB2_OK) return st;
    }
    cg_finish_rows<<<296, 256, 0, stream>>>(finalM, final_zmax, final_E, shift_acc, S, rows_out, row_offset);
    FB2_LAUNCH_CHECK();
    return FB2_OK;
}

// ------------------------------------------------------------------------------------------------ dense chains: sequential_sum_product
// trans[B][T][S][S] (log domain, contiguous) -> out[B][S][S] = trans[b,0] (x) ... (x) trans[b,T-1] in the reference's tree order
// (sum_product.py:690-701).  Every element is shifted by its largest entry and exponentiated ONCE (the shift trick of
// einsum/numpy_log.py:24-47 hoisted out of the levels), the tree runs in the linear domain on the GEMM kernel above, the result is
// taken back to the log domain with all shifts and binary exponents added in double precision.
__global__ void cg_elem_max(const float* __restrict__ x, int64_t SS, float* __restrict__ shift) {
    __shared__ float s_red[32];
    const float* xe = x + (int64_t)blockIdx.x * SS;
    float m = -FLT_MAX;
    for (int64_t i = threadIdx.x * 4; i < SS; i += blockDim.x * 4) {
        const float4 v = *reinterpret_cast<const float4*>(xe + i);
        m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
    }
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = -FLT_MAX;
        for (int q = 0; q < (int)(blockDim.x >> 5); ++q) t = fmaxf(t, s_red[q]);
        shift[blockIdx.x] = t > -FLT_MAX ? t : 0.f;  // an all -inf element stays all zero
    }
}
// leaves[e] = exp(x[e] - shift[e]) in the orientation of element e's role in its chain of T elements
__global__ void cg_exp_orient(const float* __restrict__ x, const float* __restrict__ shift, int S, int64_t T, float* __restrict__ leaves) {
    __shared__ float tile[32][33];
    const int64_t e = blockIdx.z;
    const size_t SS = (size_t)S * S;
    const bool tr = cg_role(e % T, T) != 0;
    const float sh = shift[e];
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const float* xe = x + e * SS;
    float* o = leaves + e * SS;
    for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) tile[rr][threadIdx.x] = expf(xe[(size_t)(r0 + rr) * S + c0 + threadIdx.x] - sh);
    __syncthreads();
    for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) {
        if (tr) o[(size_t)(c0 + rr) * S + r0 + threadIdx.x] = tile[threadIdx.x][rr];
        else o[(size_t)(r0 + rr) * S + c0 + threadIdx.x] = tile[rr][threadIdx.x];
    }
}
// out[b] = log(M[b]) + E[b] ln2 + sum_t shift[b,t]
__global__ void cg_finish_dense(const float* __restrict__ M, const int* __restrict__ E, const float* __restrict__ shift, int64_t T, size_t SS, float* __restrict__ out) {
    __shared__ double s_off;
    const int64_t b = blockIdx.y;
    if (threadIdx.x == 0) {
        double acc = (double)E[b] * 0.69314718055994530942;
        for (int64_t t = 0; t < T; ++t) acc += (double)shift[b * T + t];
        s_off = acc;
    }
    __syncthreads();
    const double off = s_off;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < SS; i += (size_t)gridDim.x * blockDim.x) {
        const float v = M[b * SS + i];
        out[b * SS + i] = v > 0.f ? (float)((double)logf(v) + off) : -INFINITY;
    }
}

size_t chain_product_gemm_workspace_bytes(int64_t B, int64_t T, int S) {
    const size_t SS = (size_t)S * S;
    const int64_t t1 = (T + 1) / 2, t2 = (t1 + 1) / 2;
    return align_up((size_t)B * T * SS * 4) + align_up((size_t)B * t1 * SS * 4) + align_up((size_t)B * t2 * SS * 4) + align_up((size_t)B * SS * 4) +
           align_up((size_t)B * T * 4) + 4 * align_up((size_t)B * (t1 + 1) * 4) + 2 * align_up((size_t)B * 4) + 4096;
}

int chain_product_gemm_device(const float* trans, int64_t B, int64_t T, int S, float* out, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (!chain_gemm_supported(S) || T < 2 || B < 1) return FB2_ERR_UNSUPPORTED;
    const size_t SS = (size_t)S * S;
    const int64_t t1 = (T + 1) / 2, t2 = (t1 + 1) / 2;
    if (B * T > (1 << 22) || T > 65535) return FB2_ERR_UNSUPPORTED;
    Arena ar(workspace, workspace_bytes);
    float* leaves = ar.take<float>((size_t)B * T * SS);
    CgTreeBuffers tb{};
    tb.buf[0] = ar.take<float>((size_t)B * t1 * SS);
    tb.buf[1] = ar.take<float>((size_t)B * t2 * SS);
    float* finalM = ar.take<float>((size_t)B * SS);
    float* shift = ar.take<float>((size_t)B * T);
    for (int q = 0; q < 2; ++q) {
        tb.zmax[q] = ar.take<unsigned>((size_t)B * (t1 + 1));
        tb.E[q] = ar.take<int>((size_t)B * (t1 + 1));
    }
    unsigned* final_zmax = ar.take<unsigned>((size_t)B);
    int* final_E = ar.take<int>((size_t)B);
    if (!ar.ok()) return FB2_ERR_UNSUPPORTED;  // the caller falls back to the level-by-level log-domain kernels
    cg_elem_max<<<(unsigned)(B * T), 256, 0, stream>>>(trans, (int64_t)SS, shift);
    FB2_LAUNCH_CHECK();
    for (int64_t b0 = 0; b0 < B;) {  // grid.z limit: whole chains per launch
        int64_t nb = 65535 / T;
        if (nb > B - b0) nb = B - b0;
        cg_exp_orient<<<dim3((unsigned)(S / 32), (unsigned)(S / 32), (unsigned)(nb * T)), dim3(32, 8), 0, stream>>>(trans + b0 * T * SS, shift + b0 * T, S, T,
                                                                                                                  leaves + b0 * T * SS);
        FB2_LAUNCH_CHECK();
        b0 += nb;
    }
    int st = cg_tree(leaves, nullptr, nullptr, B, T, S, tb, finalM, final_zmax, final_E, 5.82e-8f, stream);
    if (st != FB2_OK) return st;
    cg_finish_dense<<<dim3(64, (unsigned)B), 256, 0, stream>>>(finalM, final_E, shift, T, SS, out);
    FB2_LAUNCH_CHECK();
    return FB2_OK;
}

}  // namespace fb2

using namespace fb2;

extern "C" size_t fb2_hmm_chunk_summary_gemm_workspace_bytes(int64_t T, int32_t S, int64_t block_steps) {
    if (T < 2 || S < 1 || block_steps < 2) return 0;
    return hmm_chunk_summary_gemm_workspace_bytes(T, S, block_steps);
}

extern "C" int fb2_hmm_chunk_summary_gemm_f32(const float* A, const float* obs, int64_t T, int32_t S, int64_t block_steps, float* rows,
                                              double* row_offset, void* workspace, size_t workspace_bytes, void* stream) {
    int st = check_device();
    if (st != FB2_OK) return st;
    FB2_REQUIRE(A && obs && rows && row_offset && workspace, "null pointer");
    FB2_REQUIRE(T >= 2 && block_steps >= 2, "bad extents T=%lld block_steps=%lld", (long long)T, (long long)block_steps);
    if (!chain_gemm_supported(S)) {
        set_error("hmm_chunk_summary (gemm): S=%d not supported (multiples of 128 in [256, 1024])", S);
        return FB2_ERR_UNSUPPORTED;
    }
    return hmm_chunk_summary_gemm_device(A, obs, T, S, block_steps, rows, row_offset, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}
