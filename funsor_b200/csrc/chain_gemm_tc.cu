// Linear-domain matrix-chain products on the 5th-generation tensor cores: TMA-fed, warp-specialised 3xTF32 GEMM.
//
// What it computes.  The chunk summary of a time-sharded HMM (mixed_sequential_sum_product sum_product.py:766-795 with
// segments = GPUs) and every large dense scan are chains of S x S matrix products in the log semiring (sequential_sum_product
// sum_product.py:652-701; one level = Contraction[Logaddexp, Add, Tensor, Tensor] cnf.py:335-379).  The reference's own trick
// (einsum/numpy_log.py:24-47: subtract a shift, exp, multiply, log, add the shift back) is applied ONCE per chain instead of once per
// level: all elements of the tree are kept in the LINEAR domain as non-negative fp32 values times a power of two per matrix, so that a
// level is a plain GEMM.  A matrix lives as two planes, hi = the TF32 truncation of the value and lo = value - hi (exact in fp32, its
// own TF32 truncation carries 21+ mantissa bits together with hi), and a product is hi*hi + lo*hi + hi*lo with fp32 accumulation in
// tensor memory.  The planes are written by the epilogue of the level below, so the main loop does nothing but move tiles and issue
// tensor instructions:
//
//     warp 0  producer : cp.async.bulk.tensor (TMA, 128-byte swizzle) of the A / B tiles of one K chunk into a 2-stage ring,
//                        mbarrier complete_tx
//     warp 1  MMA      : tcgen05.mma.cta_group::1.kind::tf32, M = 128, N = 256, K = 8; 12 instructions per chunk; tcgen05.commit
//                        releases the stage; the last commit hands the accumulators to the epilogue
//     warps 2-5 epilogue: tcgen05.ld, optional column scaling (the observation of the second step of a leaf pair), exact power-of-two
//                        renormalisation, hi / lo split, stores in the orientation the NEXT level needs, per-matrix max by atomicMax
//
// Orientation.  Both MMA operands are K-major in shared memory, so a matrix that will be a RIGHT operand is stored transposed.  In
// the reference's tree (pairs (2p, 2p+1), odd leftover appended, sum_product.py:690-701) the role of an element is the parity of its
// index at the level where it is finally paired; the host walks the carry chain of the last element and tells the kernel.
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

constexpr int CG_BM = 128, CG_BN = 256, CG_KC = 32, CG_STAGES = 2;
constexpr int CG_A_PLANE = CG_BM * CG_KC * 4;  // 16 KB
constexpr int CG_B_PLANE = CG_BN * CG_KC * 4;  // 32 KB
constexpr int CG_STAGE_BYTES = 2 * CG_A_PLANE + 2 * CG_B_PLANE;  // 96 KB

constexpr int CG_THREADS = 320;  // warp 0 producer, warp 1 MMA, warps 2-9 epilogue (4 lane quarters x 2 column halves)


struct ChainGemmArgs {
    // outputs of this level
    float* out;            // planes [n_out][2][S][S] (orientation per element), or [n_out][S][S] fp32 when final_sum
    unsigned* zmax_out;    // [n_out] bits of the largest stored entry (zeroed before the launch)
    int* E_out;            // [n_out]
    // scaling state of the operands
    const unsigned* zmax_a;  // [n_a] or nullptr (= 1.0)
    const int* E_a;          // [n_a] or nullptr (= 0)
    const unsigned* zmax_b;
    const int* E_b;
    const float* colscale;   // nullable [pairs][S]: out[:, j] *= colscale[p][j]
    int S;
    int a_shared;          // 1: every pair uses element 0 of the A tensor map (the transition matrix P)
    int a_stride, a_off;   // A element of pair p = a_stride * p + a_off (in the A map's element index)
    int b_stride, b_off;
    int last_transposed;   // orientation of the LAST output element of a chain (others: parity of the output index)
    int n_out;             // products in this launch (all chains)
    int ppc;               // products per chain; product p belongs to chain p / ppc and is pair q = p % ppc of it
    int a_cs, b_cs, o_cs;  // elements per chain in the A / B operand arrays and in the output array
    int final_sum;         // 1: write hi + lo as one fp32 matrix, row-major (the result of a tree)
    float comp;
    long long stagger;  // start delay of odd CTAs in cycles (0 = none)
    long long* prof;  // debugging aid (FB2_CG_PROF=1): clock64 stamps of CTA (0, 0) and of the last CTA
    int mode;  // FB2_CG_MODE ablations: 1 = no TMA loads after the first two chunks, 2 = no MMAs, 4 = no epilogue stores
};

__device__ __forceinline__ uint32_t cg_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t cg_desc(uint32_t addr) {  // K-major, 128-byte swizzle, 8-row groups 1024 B apart
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
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
__device__ __forceinline__ void cg_tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void cg_tmem_ld16(uint32_t taddr, uint32_t (&d)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(d[8]), "=r"(d[9]),
          "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15])
        : "r"(taddr)
        : "memory");
}

// Persistent: grid = min(#tiles, #SMs) CTAs, each walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... (t = pair * tiles_per_product +
// tile, so neighbouring CTAs work on the same product and share its operands in L2).  Measured with clock64 stamps on the first,
// non-persistent version (S = 512): first TMA load of a CTA 6-12k cycles, main loop 1.7k cycles per K chunk (tensor-bound), epilogue
// 10-22k cycles of poorly coalesced direct stores -- i.e. half of a CTA's life was not tensor work.  Hence (i) the producer runs ahead
// into the next tile while the epilogue drains tensor memory, (ii) the epilogue stages 32 x 16 sub-blocks in shared memory and
// writes them with TMA stores (cp.async.bulk.tensor ... global.shared::cta) in the orientation the next level needs.
constexpr int CG_EPI_WARPS = 8;
constexpr int CG_EPI_BYTES = 4096;  // per epilogue warp: hi [32 x 16] + lo [32 x 16] fp32
constexpr int CG_SMEM_TOTAL = 1024 + CG_STAGES * CG_STAGE_BYTES + CG_EPI_WARPS * CG_EPI_BYTES;

__device__ __forceinline__ void cg_tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

__global__ void __launch_bounds__(CG_THREADS, 1) chain_gemm_tc(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                                               const __grid_constant__ CUtensorMap map_o, const __grid_constant__ CUtensorMap map_ot,
                                                               ChainGemmArgs a) {
    extern __shared__ unsigned char cg_smem_raw[];
    __shared__ __align__(8) unsigned long long s_full[CG_STAGES], s_empty[CG_STAGES], s_tfull, s_tempty;
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
    // De-phase the CTAs.  All persistent CTAs start together and would stay in lock step: every SM reading operands during the
    // main loop, then every SM writing its 256 KB of planes during the epilogue, so that neither the HBM read nor the write
    // stream is ever shared with tensor work of another SM.  Odd CTAs start half a tile late (a.stagger cycles).
    if (a.stagger > 0 && (blockIdx.x & 1)) {
        const long long t0 = clock64();
        while (clock64() - t0 < a.stagger) __nanosleep(200);
    }

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
                    if ((a.mode & 1) && gc >= CG_STAGES) {  // ablation: no data movement, just hand the stage back
                        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
                        continue;
                    }
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)CG_STAGE_BYTES) : "memory");
                    const uint32_t st = cg_u32(smem + (size_t)s * CG_STAGE_BYTES);
                    const int k0 = c * CG_KC;
                    cg_tma_3d(st, &map_a, k0, i0, 2 * ia, bar);
                    cg_tma_3d(st + CG_A_PLANE, &map_a, k0, i0, 2 * ia + 1, bar);
                    cg_tma_3d(st + 2 * CG_A_PLANE, &map_b, k0, j0, 2 * ib, bar);
                    cg_tma_3d(st + 2 * CG_A_PLANE + CG_B_PLANE, &map_b, k0, j0, 2 * ib + 1, bar);
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
                cg_wait(cg_u32(&s_full[s]), (uint32_t)((gc / CG_STAGES) & 1));
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
        // ---------------- epilogue: warp w owns tensor-memory lanes 32 (w % 4) .. and one half of the tile's columns
        const int q4 = warp & 3;
        const int chalf = (warp - 2) >> 2;
        unsigned char* stg = smem + (size_t)CG_STAGES * CG_STAGE_BYTES + (size_t)(warp - 2) * CG_EPI_BYTES;
        float* stg_hi = reinterpret_cast<float*>(stg);
        float* stg_lo = stg_hi + 512;
        const uint32_t stg_u32 = cg_u32(stg);
        const float fmain = 1.f + a.comp * (float)(nchunks * (CG_KC / 8));
        const size_t SS = (size_t)S * S;
        const uint32_t lane_base = tmem + ((uint32_t)(q4 * 32) << 16);
        int it = 0;
        for (int64_t t = blockIdx.x; t < total; t += gridDim.x, ++it) {
            const int p = (int)(t / tiles), tile = (int)(t % tiles);
            const int i0 = (tile / tiles_n) * CG_BM, j0 = (tile % tiles_n) * CG_BN;
            const int ch = p / a.ppc, pq = p % a.ppc;
            const int ia = a.a_shared ? 0 : ch * a.a_cs + a.a_stride * pq + a.a_off;
            const int ib = ch * a.b_cs + a.b_stride * pq + a.b_off;
            const int po = ch * a.o_cs + pq;  // output element
            const int row = i0 + q4 * 32 + lane;
            // scaling state of the operands -> exact power-of-two renormalisation
            const float za = a.zmax_a ? __uint_as_float(a.zmax_a[ia]) : 1.f;
            const float zb = a.zmax_b ? __uint_as_float(a.zmax_b[ib]) : 1.f;
            int ea = 0, eb = 0;
            if (za > 0.f) frexpf(za, &ea);
            if (zb > 0.f) frexpf(zb, &eb);
            const float scale = ldexpf(1.f, -(ea + eb));
            const bool transposed = a.final_sum ? false : (pq == a.ppc - 1 ? a.last_transposed != 0 : (pq & 1) != 0);
            if (tile == 0 && q4 == 0 && chalf == 0 && lane == 0)
                a.E_out[po] = (a.E_a ? a.E_a[ia] : 0) + (a.E_b ? a.E_b[ib] : 0) + ea + eb;
            const float* cs = a.colscale ? a.colscale + (size_t)p * S : nullptr;
            cg_wait(cg_u32(&s_tfull), (uint32_t)(it & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (do_prof && tid == 64 && it < 2) a.prof[8 + 2 * it] = clock64();
            // Phase A: pull this warp's 32 x 128 block of the accumulators into REGISTERS (128 values per thread) and hand tensor memory
            // back at once, so that the next tile's MMAs run while phase B writes the planes out.  (With the single 2 x 256-column
            // accumulator set there is no second buffer in tensor memory; measured before this split: the MMA warp sat idle for the
            // whole 19-21k-cycle epilogue of every tile.)
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
            if (do_prof && tid == 64 && it < 2) a.prof[13 + it] = clock64();
            // Phase B: hi / lo split and stores, one 32 x 16 sub-block at a time through the warp's staging slot
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                const int cc = cbase + g * 16;
                if (j0 + cc >= S || (a.mode & 4)) continue;  // warp-uniform
                if (a.final_sum) {
                    float* o = a.out + (size_t)po * SS + (size_t)row * S + j0 + cc;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        *reinterpret_cast<float4*>(o + 4 * q) = make_float4(v[g * 16 + 4 * q], v[g * 16 + 4 * q + 1], v[g * 16 + 4 * q + 2], v[g * 16 + 4 * q + 3]);
                    continue;
                }
                // the staging slot may still be read by the TMA stores of the previous sub-block
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
                if (!transposed) {  // [32 rows][16 cols], row = lane
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float h[4], l[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            h[u] = __uint_as_float(__float_as_uint(v[g * 16 + 4 * q + u]) & 0xffffe000u);
                            l[u] = v[g * 16 + 4 * q + u] - h[u];
                        }
                        *reinterpret_cast<float4*>(stg_hi + lane * 16 + 4 * q) = make_float4(h[0], h[1], h[2], h[3]);
                        *reinterpret_cast<float4*>(stg_lo + lane * 16 + 4 * q) = make_float4(l[0], l[1], l[2], l[3]);
                    }
                } else {  // out^T: [16 cols][32 rows], the warp writes 32 consecutive floats per column
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        const float h = __uint_as_float(__float_as_uint(v[g * 16 + q]) & 0xffffe000u);
                        stg_hi[q * 32 + lane] = h;
                        stg_lo[q * 32 + lane] = v[g * 16 + q] - h;
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    if (!transposed) {
                        cg_tma_store_3d(&map_o, stg_u32, j0 + cc, i0 + q4 * 32, 2 * po);
                        cg_tma_store_3d(&map_o, stg_u32 + 2048, j0 + cc, i0 + q4 * 32, 2 * po + 1);
                    } else {
                        cg_tma_store_3d(&map_ot, stg_u32, i0 + q4 * 32, j0 + cc, 2 * po);
                        cg_tma_store_3d(&map_ot, stg_u32 + 2048, i0 + q4 * 32, j0 + cc, 2 * po + 1);
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            vmax = warp_max(vmax);
            if (lane == 0 && vmax > 0.f) atomicMax(a.zmax_out + po, __float_as_uint(vmax));
            if (do_prof && tid == 64 && it < 2) a.prof[9 + 2 * it] = clock64();
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
    if (do_prof && tid == 32) a.prof[12] = clock64();
}

// ------------------------------------------------------------------------------------------------ leaf preparation and conversions
// base_hi/lo: planes of a matrix M [S][S]; out[e] planes = M[r][c] * w[e][row_scale ? r : c] split into hi / lo, w = exp(obs - shift).
// Used for: N_t^T[j][k] = P^T[j][k] e_t[k] (right operand of a leaf pair), single leaves P D_t (either orientation).
__global__ void cg_scale_planes(const float* __restrict__ base, const float* __restrict__ obs, int64_t obs_stride, int S, int row_scale, float* __restrict__ out,
                                float* __restrict__ shift_out) {
    __shared__ float s_w[1024];
    __shared__ float s_red[32];
    const int64_t e = blockIdx.y;
    const float* o = obs + e * obs_stride;
    float m = -FLT_MAX;
    for (int k = threadIdx.x; k < S; k += blockDim.x) m = fmaxf(m, o[k]);
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
    __syncthreads();
    m = -FLT_MAX;
    for (int q = 0; q < (int)(blockDim.x >> 5); ++q) m = fmaxf(m, s_red[q]);
    if (!(m > -FLT_MAX)) m = 0.f;  // an impossible observation row stays all-zero
    for (int k = threadIdx.x; k < S; k += blockDim.x) s_w[k] = expf(o[k] - m);
    if (shift_out && blockIdx.x == 0 && threadIdx.x == 0) shift_out[e] = m;
    __syncthreads();
    const size_t SS = (size_t)S * S;
    float* oh = out + (size_t)e * 2 * SS;
    float* ol = oh + SS;
    const int rows_per_block = (S + gridDim.x - 1) / gridDim.x;
    const int r0 = blockIdx.x * rows_per_block, r1 = min(S, r0 + rows_per_block);
    for (int idx = r0 * S + threadIdx.x * 4; idx < r1 * S; idx += blockDim.x * 4) {
        const int r = idx / S, c = idx % S;
        const float4 b = *reinterpret_cast<const float4*>(base + idx);
        float4 w4;
        if (row_scale) w4 = make_float4(s_w[r], s_w[r], s_w[r], s_w[r]);
        else w4 = make_float4(s_w[c], s_w[c + 1], s_w[c + 2], s_w[c + 3]);
        const float v[4] = {b.x * w4.x, b.y * w4.y, b.z * w4.z, b.w * w4.w};
        float h[4], l[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            h[q] = __uint_as_float(__float_as_uint(v[q]) & 0xffffe000u);
            l[q] = v[q] - h[q];
        }
        *reinterpret_cast<float4*>(oh + idx) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(ol + idx) = make_float4(l[0], l[1], l[2], l[3]);
    }
}
// w[e][j] = exp(obs[e][j] - max_j obs[e][j]), shift[e] = that max: the column scaling of a leaf pair's second step
__global__ void cg_obs_weights(const float* __restrict__ obs, int64_t obs_stride, int S, float* __restrict__ w, float* __restrict__ shift_out) {
    __shared__ float s_red[32];
    const int64_t e = blockIdx.x;
    const float* o = obs + e * obs_stride;
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
// planes of exp(A - colshift) in both orientations: P[i][j] (left operand) and P^T[j][i]; A row-major (prev, curr), entries <= 0 after the shift
__global__ void cg_exp_planes(const float* __restrict__ A, int S, float* __restrict__ P, float* __restrict__ PT_single) {
    const size_t SS = (size_t)S * S;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < SS; idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx / S), j = (int)(idx % S);
        const float v = expf(A[idx]);
        const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        P[idx] = h;
        P[SS + idx] = v - h;
        PT_single[(size_t)j * S + i] = v;  // one fp32 plane of P^T: the base of the right-operand leaves
    }
}
// fp32 matrix (row-major) -> planes in the requested orientation (tree results re-entering a tree as elements)
__global__ void cg_split_planes(const float* __restrict__ M, int S, const int* __restrict__ transposed, float* __restrict__ out) {
    const int64_t e = blockIdx.y;
    const size_t SS = (size_t)S * S;
    const float* m = M + (size_t)e * SS;
    float* oh = out + (size_t)e * 2 * SS;
    float* ol = oh + SS;
    const bool tr = transposed[e] != 0;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < SS; idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx / S), j = (int)(idx % S);
        const float v = tr ? m[(size_t)j * S + i] : m[idx];
        const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        oh[idx] = h;
        ol[idx] = v - h;
    }
}
// planes of one element -> planes in the other orientation (a carried leftover whose role changes)
__global__ void cg_transpose_planes(const float* __restrict__ in, int S, float* __restrict__ out) {
    const size_t SS = (size_t)S * S;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < 2 * SS; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t pl = idx / SS, r = idx % SS;
        const int i = (int)(r / S), j = (int)(r % S);
        out[pl * SS + (size_t)j * S + i] = in[idx];
    }
}
// rows[i][j] = log(M[i][j]) - log(zmax),  offset = E ln2 + log(zmax) + extra   (M = the final fp32 matrix of a tree)
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
__global__ void cg_fill_u32(unsigned* p, int64_t n, unsigned v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
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
// planes buffer [n][2][S][S] fp32 as a 3-D tensor (k, row, 2 n) with a (32, box_rows, 1) box and 128-byte swizzle
static int cg_make_map(CUtensorMap* map, const float* base, int64_t n_elems, int S, int box_rows) {
    EncodeTiledFn enc = cg_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return FB2_ERR_UNSUPPORTED;
    }
    cuuint64_t dims[3] = {(cuuint64_t)S, (cuuint64_t)S, (cuuint64_t)(2 * n_elems)};
    cuuint64_t strides[2] = {(cuuint64_t)S * 4, (cuuint64_t)S * S * 4};
    cuuint32_t box[3] = {CG_KC, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d)", (int)r);
        return FB2_ERR_CUDA;
    }
    return FB2_OK;
}

bool chain_gemm_supported(int S) { return S >= 256 && S <= 1024 && S % 128 == 0 && cg_encode_fn() != nullptr; }

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

__global__ void cg_roles_kernel(int* roles, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) roles[i] = cg_role(i, n);
}

// planes buffer [n][2][S][S] as the destination of the epilogue's TMA stores: (inner, rows) box without swizzle
static int cg_make_store_map(CUtensorMap* map, const float* base, int64_t n_elems, int S, int box_inner, int box_rows) {
    EncodeTiledFn enc = cg_encode_fn();
    if (!enc) return FB2_ERR_UNSUPPORTED;
    cuuint64_t dims[3] = {(cuuint64_t)S, (cuuint64_t)S, (cuuint64_t)(2 * n_elems)};
    cuuint64_t strides[2] = {(cuuint64_t)S * 4, (cuuint64_t)S * S * 4};
    cuuint32_t box[3] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (store map) failed (%d)", (int)r);
        return FB2_ERR_CUDA;
    }
    return FB2_OK;
}

static int cg_launch(const float* a_planes, int64_t n_a, const float* b_planes, int64_t n_b, const ChainGemmArgs& args, int64_t pairs, cudaStream_t stream) {
    CUtensorMap ma, mb, mo, mot;
    int st = cg_make_map(&ma, a_planes, n_a, args.S, CG_BM);
    if (st != FB2_OK) return st;
    st = cg_make_map(&mb, b_planes, n_b, args.S, CG_BN);
    if (st != FB2_OK) return st;
    // final_sum writes one fp32 plane per product with direct stores; the maps then only need to be valid descriptors
    const float* obase = args.final_sum ? a_planes : args.out;
    const int64_t on = args.final_sum ? n_a : (args.ppc > 0 ? (int64_t)(pairs / args.ppc) * args.o_cs : pairs);
    st = cg_make_store_map(&mo, obase, on, args.S, 16, 32);
    if (st != FB2_OK) return st;
    st = cg_make_store_map(&mot, obase, on, args.S, 32, 16);
    if (st != FB2_OK) return st;
    FB2_CUDA(cudaFuncSetAttribute(chain_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, CG_SMEM_TOTAL));
    ChainGemmArgs margs = args;
    if (margs.ppc <= 0) {  // single chain
        margs.ppc = (int)pairs;
        margs.a_cs = margs.b_cs = margs.o_cs = 0;
    }
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
    FB2_REQUIRE(pairs >= 1 && pairs <= (1 << 24), "chain_gemm: %lld products in one level", (long long)pairs);
    static int sm_count = 0;
    if (!sm_count) {
        int dev = 0;
        FB2_CUDA(cudaGetDevice(&dev));
        FB2_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    const int64_t total = (int64_t)tiles * pairs;
    const unsigned grid = (unsigned)(total < sm_count ? total : sm_count);
    {
        const char* se = getenv("FB2_CG_STAGGER");
        const long long cyc = se ? atoll(se) : (long long)(args.S / CG_KC) * 1000;  // ~ half of a tile's main loop
        margs.stagger = (total >= 4 * (int64_t)grid) ? cyc : 0;
    }
    chain_gemm_tc<<<grid, CG_THREADS, CG_SMEM_TOTAL, stream>>>(ma, mb, mo, mot, margs);
    FB2_LAUNCH_CHECK();
    if (margs.prof) {
        long long h[32];
        FB2_CUDA(cudaStreamSynchronize(stream));
        FB2_CUDA(cudaMemcpy(h, d_prof, sizeof(h), cudaMemcpyDeviceToHost));
        fprintf(stderr, "[cg prof] pairs %lld cta 0: setup %lld | tile0 first chunk +%lld, last chunk +%lld | tile1 first chunk +%lld, last chunk +%lld | "
                        "epilogue0 start +%lld, tensor memory released after %lld, len %lld | epilogue1 start +%lld, released after %lld, len %lld | total %lld (cycles)\n",
                (long long)pairs, h[1] - h[0], h[2] - h[1], h[3] - h[2], h[4] - h[3], h[5] - h[4], h[8] - h[0], h[13] - h[8], h[9] - h[8], h[10] - h[0],
                h[14] - h[10], h[11] - h[10], h[12] - h[0]);
    }
    return FB2_OK;
}

// Balanced tree (the reference's order) over `n` elements given as planes in `cur` (roles by cg_role) with scaling state
// (zmax, E); ping-pongs between two plane buffers; writes the final product as ONE fp32 matrix + its (zmax, E).
struct CgTreeBuffers {
    float* planes[2];     // each >= ceil(n/2) elements of 2 S^2 floats
    unsigned* zmax[2];    // each >= n
    int* E[2];
};
static int cg_tree(const float* first_a, int64_t first_na, int first_a_shared, const float* first_b, int64_t first_nb, const float* first_colscale,
                   int64_t n_pairs_first, const float* odd_leaf /*planes of a leftover leaf, role-correct, or nullptr*/, int S, const CgTreeBuffers& buf,
                   float* final_out, unsigned* final_zmax, int* final_E, float comp, cudaStream_t stream) {
    // level 0 -> 1: n_pairs_first products (A from first_a, B from first_b), optional leftover leaf appended
    int64_t n = n_pairs_first + (odd_leaf ? 1 : 0);  // elements at the level being produced
    const size_t SS2 = (size_t)2 * S * S;
    int which = 0;
    {
        ChainGemmArgs a{};
        a.S = S; a.comp = comp;
        a.a_shared = first_a_shared; a.a_stride = first_a_shared ? 0 : 1; a.a_off = 0; a.b_stride = 1; a.b_off = 0;
        a.colscale = first_colscale;
        a.n_out = (int)n_pairs_first;
        const bool is_final = n == 1;
        a.final_sum = is_final;
        a.out = is_final ? final_out : buf.planes[which];
        a.zmax_out = is_final ? final_zmax : buf.zmax[which];
        a.E_out = is_final ? final_E : buf.E[which];
        a.last_transposed = cg_role(n_pairs_first - 1, n);
        FB2_CUDA(cudaMemsetAsync(a.zmax_out, 0, (size_t)n * sizeof(unsigned), stream));
        int st = cg_launch(first_a, first_na, first_b, first_nb, a, n_pairs_first, stream);
        if (st != FB2_OK) return st;
        if (odd_leaf) {  // append the leftover leaf (already in its role's orientation; stored <= 1, E = 0)
            FB2_CUDA(cudaMemcpyAsync(buf.planes[which] + (size_t)n_pairs_first * SS2, odd_leaf, SS2 * 4, cudaMemcpyDeviceToDevice, stream));
            cg_fill_u32<<<1, 32, 0, stream>>>(buf.zmax[which] + n_pairs_first, 1, 0x3f800000u);  // bits of 1.0f
            FB2_LAUNCH_CHECK();
            FB2_CUDA(cudaMemsetAsync(buf.E[which] + n_pairs_first, 0, sizeof(int), stream));
        }
        if (is_final) return FB2_OK;
    }
    while (n > 1) {
        const int64_t pairs = n / 2, nd = (n + 1) / 2;
        const bool odd = n & 1;
        const int src = which, dst = which ^ 1;
        ChainGemmArgs a{};
        a.S = S; a.comp = comp;
        a.a_shared = 0; a.a_stride = 2; a.a_off = 0; a.b_stride = 2; a.b_off = 1;
        a.zmax_a = buf.zmax[src]; a.E_a = buf.E[src]; a.zmax_b = buf.zmax[src]; a.E_b = buf.E[src];
        a.n_out = (int)pairs;
        const bool is_final = nd == 1;
        a.final_sum = is_final;
        a.out = is_final ? final_out : buf.planes[dst];
        a.zmax_out = is_final ? final_zmax : buf.zmax[dst];
        a.E_out = is_final ? final_E : buf.E[dst];
        a.last_transposed = cg_role(pairs - 1, nd);
        FB2_CUDA(cudaMemsetAsync(a.zmax_out, 0, (size_t)nd * sizeof(unsigned), stream));
        int st = cg_launch(buf.planes[src], n, buf.planes[src], n, a, pairs, stream);
        if (st != FB2_OK) return st;
        if (odd) {  // carry the leftover: it was stored row-major (its index was even); transpose if its eventual role is "right"
            const float* lsrc = buf.planes[src] + (size_t)(n - 1) * SS2;
            float* ldst = buf.planes[dst] + (size_t)pairs * SS2;
            // the leftover's CURRENT orientation is its eventual role as computed when it was written (cg_role follows the whole carry
            // chain), so a plain copy is correct
            FB2_CUDA(cudaMemcpyAsync(ldst, lsrc, SS2 * 4, cudaMemcpyDeviceToDevice, stream));
            FB2_CUDA(cudaMemcpyAsync(buf.zmax[dst] + pairs, buf.zmax[src] + (n - 1), sizeof(unsigned), cudaMemcpyDeviceToDevice, stream));
            FB2_CUDA(cudaMemcpyAsync(buf.E[dst] + pairs, buf.E[src] + (n - 1), sizeof(int), cudaMemcpyDeviceToDevice, stream));
        }
        n = nd;
        which = dst;
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
    const int64_t pairs = bs / 2 + 2;
    const int64_t nblocks = cg_block_count(T, bs);
    size_t w = 0;
    w += align_up(2 * SS * 4) + align_up(SS * 4);                       // P planes, P^T single plane
    w += align_up((size_t)pairs * 2 * SS * 4);                           // leaf right operands N (+ leftover leaf slot)
    w += align_up((size_t)pairs * S * 4);                                // column weights
    w += 2 * align_up((size_t)(bs + 2) * 4);                             // shifts
    w += align_up((size_t)pairs * 2 * SS * 4) + align_up((size_t)(pairs / 2 + 1) * 2 * SS * 4);  // tree ping-pong
    w += 4 * align_up((size_t)(pairs + 1) * 4);                          // zmax / E ping-pong
    w += align_up((size_t)nblocks * SS * 4) + 3 * align_up((size_t)nblocks * 4);  // block results, zmax, E, roles
    w += align_up((size_t)nblocks * 2 * SS * 4) + align_up((size_t)(nblocks / 2 + 1) * 2 * SS * 4);  // upper tree
    w += 4 * align_up((size_t)(nblocks + 1) * 4);
    w += align_up(SS * 4) + 3 * 256;                                     // final matrix, zmax, E, offset accumulator
    return w + 4096;
}

// out rows [S][S] fp32 (largest entry 0) + row_offset [S] float64 of  (x)_{t} (A + obs[t])  for ONE chain; obs [T][S] contiguous
int hmm_chunk_summary_gemm_device(const float* A, const float* obs, int64_t T, int S, int64_t block_steps, float* rows_out, double* row_offset,
                                  void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (!chain_gemm_supported(S) || T < 2) return FB2_ERR_UNSUPPORTED;
    const size_t SS = (size_t)S * S;
    int64_t bs = block_steps < T ? block_steps : T;
    bs += bs & 1;
    FB2_REQUIRE(bs >= 2 && bs <= 131068, "hmm_chunk_summary (gemm): block of %lld steps outside [2, 131068]", (long long)bs);
    const int64_t max_pairs = bs / 2 + 2;
    const int64_t nblocks = cg_block_count(T, bs);
    Arena ar(workspace, workspace_bytes);
    float* P = ar.take<float>(2 * SS);
    float* PT = ar.take<float>(SS);
    float* N = ar.take<float>((size_t)max_pairs * 2 * SS);
    float* W = ar.take<float>((size_t)max_pairs * S);
    float* shiftN = ar.take<float>((size_t)bs + 2);
    float* shiftW = ar.take<float>((size_t)bs + 2);
    CgTreeBuffers tb{}, ub{};
    tb.planes[0] = ar.take<float>((size_t)max_pairs * 2 * SS);
    tb.planes[1] = ar.take<float>((size_t)(max_pairs / 2 + 1) * 2 * SS);
    for (int q = 0; q < 2; ++q) {
        tb.zmax[q] = ar.take<unsigned>((size_t)max_pairs + 1);
        tb.E[q] = ar.take<int>((size_t)max_pairs + 1);
    }
    float* blockM = ar.take<float>((size_t)nblocks * SS);
    unsigned* block_zmax = ar.take<unsigned>((size_t)nblocks);
    int* block_E = ar.take<int>((size_t)nblocks);
    int* block_role = ar.take<int>((size_t)nblocks);
    ub.planes[0] = ar.take<float>((size_t)nblocks * 2 * SS);
    ub.planes[1] = ar.take<float>((size_t)(nblocks / 2 + 1) * 2 * SS);
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
    cg_exp_planes<<<296, 256, 0, stream>>>(A, S, P, PT);
    FB2_LAUNCH_CHECK();
    int64_t t0 = 0;
    for (int64_t b = 0; b < nblocks; ++b) {
        int64_t nt = bs < T - t0 ? bs : T - t0;
        if (T - (t0 + nt) == 1) nt += 1;
        const int64_t pairs = nt / 2;
        const bool odd = nt & 1;
        const float* ob = obs + t0 * S;
        float* dstM = nblocks == 1 ? finalM : blockM + (size_t)b * SS;
        unsigned* dstZ = nblocks == 1 ? final_zmax : block_zmax + b;
        int* dstE = nblocks == 1 ? final_E : block_E + b;
        // right operands of the leaf pairs: N_p^T[j][k] = P^T[j][k] * e_{2p}[k]  (column scaling of P^T), one element per pair
        cg_scale_planes<<<dim3(8, (unsigned)pairs), 256, 0, stream>>>(PT, ob, 2 * (int64_t)S, S, 0, N, shiftN);
        FB2_LAUNCH_CHECK();
        // column weights of the second step of every pair (applied in the epilogue of the leaf products)
        cg_obs_weights<<<(unsigned)pairs, 128, 0, stream>>>(ob + S, 2 * (int64_t)S, S, W, shiftW);
        FB2_LAUNCH_CHECK();
        cg_sum_shifts<<<32, 256, 0, stream>>>(shiftN, pairs, shift_acc);
        FB2_LAUNCH_CHECK();
        cg_sum_shifts<<<32, 256, 0, stream>>>(shiftW, pairs, shift_acc);
        FB2_LAUNCH_CHECK();
        const float* odd_leaf = nullptr;
        if (odd) {
            // the last step of an odd block is a leaf of its own, M = P D_t, stored in the orientation of its eventual role.
            // Row scaling of P^T gives M^T (the "right operand" orientation); the "left operand" orientation is its transpose.
            const int role = cg_role(pairs, pairs + 1);
            float* leaf = N + (size_t)pairs * 2 * SS;  // spare slot behind the pair operands
            cg_scale_planes<<<dim3(8, 1), 256, 0, stream>>>(PT, ob + (nt - 1) * S, S, S, 1, leaf, shiftN + pairs);
            FB2_LAUNCH_CHECK();
            cg_sum_shifts<<<1, 32, 0, stream>>>(shiftN + pairs, 1, shift_acc);
            FB2_LAUNCH_CHECK();
            if (role == 0) {
                float* tmp = tb.planes[1];  // not in use before the tree starts
                cg_transpose_planes<<<296, 256, 0, stream>>>(leaf, S, tmp);
                FB2_LAUNCH_CHECK();
                FB2_CUDA(cudaMemcpyAsync(leaf, tmp, 2 * SS * 4, cudaMemcpyDeviceToDevice, stream));
            }
            odd_leaf = leaf;
        }
        int st = cg_tree(P, 1, 1, N, pairs + (odd ? 1 : 0), W, pairs, odd_leaf, S, tb, dstM, dstZ, dstE, comp, stream);
        if (st != FB2_OK) return st;
        t0 += nt;
    }
    if (nblocks > 1) {
        // the block results re-enter a tree as elements: split into planes in the orientation of their role
        cg_roles_kernel<<<1, 256, 0, stream>>>(block_role, nblocks);
        FB2_LAUNCH_CHECK();
        cg_split_planes<<<dim3(64, (unsigned)nblocks), 256, 0, stream>>>(blockM, S, block_role, ub.planes[0]);
        FB2_LAUNCH_CHECK();
        FB2_CUDA(cudaMemcpyAsync(ub.zmax[0], block_zmax, (size_t)nblocks * sizeof(unsigned), cudaMemcpyDeviceToDevice, stream));
        FB2_CUDA(cudaMemcpyAsync(ub.E[0], block_E, (size_t)nblocks * sizeof(int), cudaMemcpyDeviceToDevice, stream));
        int64_t n = nblocks;
        int which = 0;
        const size_t SS2 = 2 * SS;
        while (n > 1) {
            const int64_t pairs = n / 2, nd = (n + 1) / 2;
            const int src = which, dst = which ^ 1;
            ChainGemmArgs a{};
            a.S = S; a.comp = comp;
            a.a_stride = 2; a.a_off = 0; a.b_stride = 2; a.b_off = 1;
            a.zmax_a = ub.zmax[src]; a.E_a = ub.E[src]; a.zmax_b = ub.zmax[src]; a.E_b = ub.E[src];
            a.n_out = (int)pairs;
            const bool is_final = nd == 1;
            a.final_sum = is_final;
            a.out = is_final ? finalM : ub.planes[dst];
            a.zmax_out = is_final ? final_zmax : ub.zmax[dst];
            a.E_out = is_final ? final_E : ub.E[dst];
            a.last_transposed = cg_role(pairs - 1, nd);
            FB2_CUDA(cudaMemsetAsync(a.zmax_out, 0, (size_t)nd * sizeof(unsigned), stream));
            int st = cg_launch(ub.planes[src], n, ub.planes[src], n, a, pairs, stream);
            if (st != FB2_OK) return st;
            if (n & 1) {
                FB2_CUDA(cudaMemcpyAsync(ub.planes[dst] + (size_t)pairs * SS2, ub.planes[src] + (size_t)(n - 1) * SS2, SS2 * 4, cudaMemcpyDeviceToDevice, stream));
                FB2_CUDA(cudaMemcpyAsync(ub.zmax[dst] + pairs, ub.zmax[src] + (n - 1), sizeof(unsigned), cudaMemcpyDeviceToDevice, stream));
                FB2_CUDA(cudaMemcpyAsync(ub.E[dst] + pairs, ub.E[src] + (n - 1), sizeof(int), cudaMemcpyDeviceToDevice, stream));
            }
            n = nd;
            which = dst;
        }
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
// planes[e] = hi / lo split of exp(x[e] - shift[e]) in the orientation of element e's role in its chain of T elements
__global__ void cg_exp_split_tiles(const float* __restrict__ x, const float* __restrict__ shift, int S, int64_t T, float* __restrict__ planes) {
    __shared__ float tile[32][33];
    const int64_t e = blockIdx.z;
    const size_t SS = (size_t)S * S;
    const bool tr = cg_role(e % T, T) != 0;
    const float sh = shift[e];
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const float* xe = x + e * SS;
    for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) tile[rr][threadIdx.x] = expf(xe[(size_t)(r0 + rr) * S + c0 + threadIdx.x] - sh);
    __syncthreads();
    float* oh = planes + e * 2 * SS;
    float* ol = oh + SS;
    for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) {
        const float v = tr ? tile[threadIdx.x][rr] : tile[rr][threadIdx.x];
        const size_t o = tr ? (size_t)(c0 + rr) * S + r0 + threadIdx.x : (size_t)(r0 + rr) * S + c0 + threadIdx.x;
        const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        oh[o] = h;
        ol[o] = v - h;
    }
}
__global__ void cg_copy_leftover(const float* __restrict__ src, const unsigned* __restrict__ zsrc, const int* __restrict__ esrc, int64_t n, int64_t nd, size_t SS2,
                                 float* __restrict__ dst, unsigned* __restrict__ zdst, int* __restrict__ edst, int first_level) {
    const int64_t ch = blockIdx.y;
    const float* s_ = src + (size_t)(ch * n + n - 1) * SS2;
    float* d_ = dst + (size_t)(ch * nd + nd - 1) * SS2;
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < SS2; i += (size_t)gridDim.x * blockDim.x * 4)
        *reinterpret_cast<float4*>(d_ + i) = *reinterpret_cast<const float4*>(s_ + i);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        zdst[ch * nd + nd - 1] = first_level ? 0x3f800000u : zsrc[ch * n + n - 1];
        edst[ch * nd + nd - 1] = first_level ? 0 : esrc[ch * n + n - 1];
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
    return align_up((size_t)B * T * 2 * SS * 4) + align_up((size_t)B * t1 * 2 * SS * 4) + align_up((size_t)B * t2 * 2 * SS * 4) + align_up((size_t)B * SS * 4) +
           align_up((size_t)B * T * 4) + 4 * align_up((size_t)B * (t1 + 1) * 4) + 4096;
}

int chain_product_gemm_device(const float* trans, int64_t B, int64_t T, int S, float* out, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (!chain_gemm_supported(S) || T < 2 || B < 1) return FB2_ERR_UNSUPPORTED;
    const size_t SS = (size_t)S * S, SS2 = 2 * SS;
    const int64_t t1 = (T + 1) / 2, t2 = (t1 + 1) / 2;
    if (B * T > (1 << 22)) return FB2_ERR_UNSUPPORTED;
    Arena ar(workspace, workspace_bytes);
    float* leaves = ar.take<float>((size_t)B * T * SS2);
    float* buf[2] = {ar.take<float>((size_t)B * t1 * SS2), ar.take<float>((size_t)B * t2 * SS2)};
    float* finalM = ar.take<float>((size_t)B * SS);
    float* shift = ar.take<float>((size_t)B * T);
    unsigned* zm[2];
    int* Ex[2];
    for (int q = 0; q < 2; ++q) {
        zm[q] = ar.take<unsigned>((size_t)B * (t1 + 1));
        Ex[q] = ar.take<int>((size_t)B * (t1 + 1));
    }
    if (!ar.ok()) return FB2_ERR_UNSUPPORTED;  // the caller falls back to the level-by-level log-domain kernels
    cg_elem_max<<<(unsigned)(B * T), 256, 0, stream>>>(trans, (int64_t)SS, shift);
    FB2_LAUNCH_CHECK();
    {
        int64_t done = 0;
        while (done < B * T) {  // grid.z limit
            const int64_t nz = B * T - done < 65535 ? B * T - done : 65535;
            FB2_REQUIRE(done % T == 0 || nz == B * T - done || true, "unreachable");
            // roles depend on e % T, so offsetting the element pointer by `done` needs the same offset inside the kernel: pass via pointer arithmetic
            // only when `done` is a multiple of T; otherwise launch per chain
            if (done % T != 0) return FB2_ERR_UNSUPPORTED;
            int64_t take = nz - nz % T;
            if (take == 0) return FB2_ERR_UNSUPPORTED;  // one chain longer than 65535 elements
            cg_exp_split_tiles<<<dim3((unsigned)(S / 32), (unsigned)(S / 32), (unsigned)take), dim3(32, 8), 0, stream>>>(trans + done * SS, shift + done, S, T,
                                                                                                                  leaves + done * SS2);
            FB2_LAUNCH_CHECK();
            done += take;
        }
    }
    const float comp = 5.82e-8f;
    const float* src = leaves;
    const unsigned* zsrc = nullptr;
    const int* esrc = nullptr;
    int64_t n = T;
    int which = 0;
    bool first = true;
    while (n > 1) {
        const int64_t pairs = n / 2, nd = (n + 1) / 2;
        const bool is_final = nd == 1;
        ChainGemmArgs a{};
        a.S = S; a.comp = comp;
        a.a_stride = 2; a.a_off = 0; a.b_stride = 2; a.b_off = 1;
        a.zmax_a = zsrc; a.E_a = esrc; a.zmax_b = zsrc; a.E_b = esrc;
        a.ppc = (int)pairs; a.a_cs = (int)n; a.b_cs = (int)n; a.o_cs = (int)nd;
        a.n_out = (int)(B * pairs);
        a.final_sum = is_final;
        a.out = is_final ? finalM : buf[which];
        a.zmax_out = zm[which];
        a.E_out = Ex[which];
        a.last_transposed = cg_role(pairs - 1, nd);
        FB2_CUDA(cudaMemsetAsync(a.zmax_out, 0, (size_t)B * nd * sizeof(unsigned), stream));
        int st = cg_launch(src, B * n, src, B * n, a, B * pairs, stream);
        if (st != FB2_OK) return st;
        if (n & 1) {
            cg_copy_leftover<<<dim3(32, (unsigned)B), 256, 0, stream>>>(src, zsrc, esrc, n, nd, SS2, buf[which], zm[which], Ex[which], first ? 1 : 0);
            FB2_LAUNCH_CHECK();
        }
        src = buf[which]; zsrc = zm[which]; esrc = Ex[which];
        n = nd;
        which ^= 1;
        first = false;
    }
    cg_finish_dense<<<dim3(64, (unsigned)B), 256, 0, stream>>>(finalM, esrc, shift, T, SS, out);
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
