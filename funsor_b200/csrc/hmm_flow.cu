// Dataflow forward pass for the log semiring with a time-homogeneous transition matrix.
//
// Reference semantics: DiscreteHMM.log_prob funsor/pyro/hmm.py:129-135, i.e.
// sequential_sum_product(logaddexp, add, A + obs) (sum_product.py:652-701) followed by
// init + reduce(curr), reduce(state), evaluated as the vector recurrence
//     alpha_t[j] = logsumexp_i(alpha_{t-1}[i] + A[i,j]) + obs[t,j]
// with the reference's own max-subtract / exp trick (einsum/numpy_log.py:24-47): the contraction
// runs in the linear domain on the tensor pipe and only O(S) exp / O(1) log per step remain.
//
// The step-to-step dependency chain is T long, so the kernel is built around the latency of one
// step, not around throughput:
//   * P = exp(A - colmax(A)) never moves: every CTA keeps its [Ks x 64] block as split-TF32
//     (hi + lo) mma.sync B-fragments in REGISTERS for the whole pass.
//   * 2-D split.  A cluster of 8 CTAs owns 64 output columns; inside the cluster the contraction
//     index is split 8 ways (Ks = Sp/8 rows per CTA).  A CTA therefore reads only its K-slice of the
//     message (16 chains x Ks values) instead of the whole vector, which keeps the per-step L2
//     traffic at Sp*16*8 B per cluster instead of per CTA.
//   * The 8 partial sums of a column group are reduced over distributed shared memory
//     (st.shared::cluster + a remote mbarrier arrive); CTA q of the cluster is the final owner of
//     columns [8q, 8q+8) of the cluster's 64.
//   * Messages travel between clusters through L2 as self-validating 8-byte packets
//     {fp32 mantissa, step stamp | block exponent}: a consumer polls the data itself, so there is no
//     grid barrier, no separate flag and no fence on the critical path.
//   * Block floating point.  Every group of 8 columns carries its own integer exponent, so a
//     producer normalises with purely local information; consumers align the exponents of their
//     K-slice (exact power-of-two scaling).  The absolute exponent is tracked by the final owner in
//     an int, so the log-likelihood keeps fp32 mantissa accuracy no matter how large |log Z| grows.
//   * 3xTF32 (hi*hi + lo*hi + hi*lo) keeps the contraction at fp32-class accuracy.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cooperative_groups.h>
#include <stdlib.h>

#include <mutex>

#include "common.cuh"

namespace fb2 {

constexpr int FLOW_CLUSTER = 8;
constexpr int FLOW_THREADS = 256;
constexpr int FLOW_CHAINS = 16;  // chains per group = M of the mma tile
constexpr float kLog2e = 1.4426950408889634f;
constexpr int FLOW_PROF_CTAS = 160;
constexpr int FLOW_OBS_DEPTH = 8;     // observation prefetch distance in steps (HBM latency << 7 steps)
constexpr long long FLOW_SPIN_LIMIT = 2000000000ll;  // cycles (~1 s) before a wait gives up and aborts the pass
#ifndef FLOW_EXPO_FIRST_KW
#define FLOW_EXPO_FIRST_KW 8  // single-hop kernel: from this many fragments per warp on, spin on the exponent words first (see the fetch loop)
#endif

struct FlowArgs {
    const float* init;  // [B,S] log domain, nullable (= 0)
    int64_t init_sb;
    const float* A;     // [S,S] log domain
    const float* obs;   // [B,T,S], nullable (= 0)
    const float* cA;    // [Sp] clamped column maxima of A (0 in the padding)
    int64_t B, T;
    int S;
    unsigned long long* pkt;  // [nsets][2][16][Sp] packets
    float* fin_u;             // [groups*16][Sp] final mantissas
    int* fin_e;               // [groups*16][Sp/8] final absolute exponents (INT_MIN = zero group)
    int* abort_flag;
    // optional history of the messages after every step (forward-backward): mantissas [B][hist_T][Sp] and absolute block
    // exponents [B][hist_T][Sp/8] (INT_MIN = zero group)
    float* hist_u;
    int* hist_e;
    int64_t hist_T;
    int64_t obs_T;   // time extent of the obs array (rows per chain); the pass runs p.T steps over obs rows [0, p.T)
    const unsigned* ready_flags;  // nullable: obs rows [c*rows_per_chunk, ...) of every chain are resident once ready_flags[c] != 0
    int64_t rows_per_chunk;       //   (host-buffer entry point: the H2D copy of obs is overlapped with the pass)
    int cpo;         // chains per obs entry: 1 normally; S for chunk summaries (chain c = b*S + row reads obs[b] and starts from
                     // the semiring identity row `row` instead of init)
    int transpose;   // 1: contract with A^T (backward pass)
    int reverse;     // 1: step t consumes obs row p.T-1-t and writes history slot p.T-1-t
    long long* prof;  // nullable: [gridDim][8] accumulated cycles per phase (FB2_FLOW_PROF=1)
    int nsets, groups_total;
    int mode;  // experiment switches (FB2_FLOW_MODE): 1 publish with atom.exch (default), 2/4 sleep 0.5/1 us before polling, 16 four staggered probes
};

// ------------------------------------------------------------------------------------------ PTX helpers
// fp32 -> TF32, round to nearest, ties away from zero (= cvt.rna.tf32.f32 for every finite input; inf / NaN stay non-finite).  Two integer
// instructions: ptxas expands cvt.rna.tf32 into ~6 (the staging phase of the dataflow kernels is ALU-issue bound).
__device__ __forceinline__ uint32_t cvt_tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_bf16_k16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {  // {upper half: hi, lower half: lo}, round to nearest
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// upper halves of two fp32 registers as one bf16 pair {hi: b, lo: a}; volatile: recomputed where it is used (ptxas otherwise hoists the 32
// loop-invariant results of the contraction loop out of the time loop and holds them in registers for the whole pass)
__device__ __forceinline__ uint32_t upper_halves(uint32_t a, uint32_t b) {
    uint32_t r;
    asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// Remote shared-memory stores that signal the destination CTA's mbarrier by transaction bytes: data and
// notification travel together, no release/acquire fence at cluster scope on the critical path.
__device__ __forceinline__ void st_async_v2f(uint32_t addr, float a, float b, uint32_t mbar) {
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(addr), "f"(a),
                 "f"(b), "r"(mbar)
                 : "memory");
}
__device__ __forceinline__ void st_async_u32(uint32_t addr, uint32_t v, uint32_t mbar) {
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.u32 [%0], %1, [%2];" ::"r"(addr), "r"(v), "r"(mbar)
                 : "memory");
}
__device__ __forceinline__ void mbar_arm(uint32_t addr, uint32_t tx_bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(tx_bytes) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t addr, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t addr, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void ld_pkt2(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ void st_pkt(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// Wait until the packet pair at `addr` carries stamp `want`.  Four probe loads are kept in flight a quarter of a
// hand-over apart (one L2 store->load hand-over is ~936 cycles, scripts/microbench/pingpong.cu), so a new packet is
// seen ~230 cycles after it lands instead of up to a full round trip later.  The step time is the maximum over all
// polling threads of the chip, so this quantisation was paid in full on every step.
__device__ __forceinline__ void spin_cycles(int n) {
    const long long c = clock64();
    while (clock64() - c < n) {
    }
}
__device__ __forceinline__ bool probe_wait(const unsigned long long* addr, unsigned want, long long start) {
    unsigned long long a0, b0, a1, b1, a2, b2, a3, b3;
    ld_pkt2(addr, a0, b0);
    if ((unsigned)(b0 >> 48) == want) return true;
    spin_cycles(200);
    ld_pkt2(addr, a1, b1);
    spin_cycles(200);
    ld_pkt2(addr, a2, b2);
    spin_cycles(200);
    ld_pkt2(addr, a3, b3);
    for (;;) {
        if ((unsigned)(b0 >> 48) == want) return true;
        ld_pkt2(addr, a0, b0);
        if ((unsigned)(b1 >> 48) == want) return true;
        ld_pkt2(addr, a1, b1);
        if ((unsigned)(b2 >> 48) == want) return true;
        ld_pkt2(addr, a2, b2);
        if ((unsigned)(b3 >> 48) == want) return true;
        ld_pkt2(addr, a3, b3);
        if (clock64() - start > FLOW_SPIN_LIMIT) return false;
    }
}
__device__ __forceinline__ int sext15(int x) { return (int)((unsigned)x << 17) >> 17; }
__device__ __forceinline__ float pow2i(int d) {  // 2^d for d <= 0, flushed to 0 below 2^-126
    return d < -126 ? 0.f : __int_as_float((127 + d) << 23);
}
__device__ __forceinline__ float group8_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return v;
}

// Normalise the 8 values of one column group (one per lane of an aligned 8-lane group): returns the
// mantissa of this lane and the new wrapped exponent / zero flag of the group.
//   true value = s * 2^(x) * 2^(e_in)   ->   u * 2^(E)
__device__ __forceinline__ float flow_normalise(float s, float x, int e_in15, int& E15, bool& zero, int& dE, bool check_range = true) {
    dE = 0;
    const float xm = group8_max(x);
    float nf = 0.f, e2 = 0.f;
    if (xm > -1e30f) {
        nf = ceilf(xm);
        e2 = exp2f(x - nf);
        // a per-step |obs + colmax| beyond the 15-bit wrapped block-exponent range poisons the result (NaN)
        if (check_range && !(fabsf(nf) <= 8000.f)) e2 = __int_as_float(0x7fc00000);
    }
    float u = s * e2;
    float mu = group8_max(u);
    zero = !(mu > 0.f);
    if (!(mu <= FLT_MAX)) {  // inf or NaN input: poison the group, the log-likelihood becomes NaN
        E15 = e_in15 & 0x7FFF;
        zero = false;
        return __int_as_float(0x7fc00000);
    }
    if (zero) {
        E15 = e_in15 & 0x7FFF;
        return 0.f;
    }
    int extra = 0;
    if (mu < 1e-20f) {
        u *= 0x1p64f;
        mu *= 0x1p64f;
        extra = -64;
    }
    const int q = ((__float_as_int(mu) >> 23) & 0xFF) - 127;
    u *= __int_as_float((127 - q) << 23);
    dE = (int)nf + q + extra;
    E15 = (e_in15 + dE) & 0x7FFF;
    return u;
}

// The same normalisation split in two, so that the observation-only half (group max of x, exp2) runs BEFORE the thread
// waits for the partial sums and only a short dependent chain remains after the wait.
struct FlowObs1 {
    float e2;
    int nf;
    int bad;
};
__device__ __forceinline__ FlowObs1 flow_obs1(float x, bool check_range) {
    FlowObs1 o;
    const float xm = group8_max(x);
    float nf = 0.f;
    o.e2 = 0.f;
    if (xm > -1e30f) {
        nf = ceilf(xm);
        o.e2 = exp2f(x - nf);
    }
    o.bad = (check_range && xm > -1e30f && !(fabsf(nf) <= 8000.f)) ? 1 : 0;
    o.nf = (int)nf;
    return o;
}
__device__ __forceinline__ float flow_finish1(float s, const FlowObs1& o, int e_in15, int& E15, bool& zero, int& dE) {
    float u = s * o.e2;
    int bad = o.bad | (!(u <= FLT_MAX) ? 1 : 0);
    float mu = group8_max(u);
    {
        unsigned b = __ballot_sync(0xffffffffu, bad);
        bad = (b >> ((threadIdx.x & 31) & ~7)) & 0xFF ? 1 : 0;  // any lane of this 8-lane group
    }
    dE = 0;
    E15 = e_in15 & 0x7FFF;
    zero = !(mu > 0.f);
    if (bad) {
        zero = false;
        return __int_as_float(0x7fc00000);
    }
    if (zero) return 0.f;
    int extra = 0;
    if (mu < 1e-20f) {
        u *= 0x1p64f;
        mu *= 0x1p64f;
        extra = -64;
    }
    const int q = ((__float_as_int(mu) >> 23) & 0xFF) - 127;
    u *= __int_as_float((127 - q) << 23);
    dE = o.nf + q + extra;
    E15 = (e_in15 + dE) & 0x7FFF;
    return u;
}

// ------------------------------------------------------------------------------------------ the kernel
// K16 = k16-steps per CTA (Ks = 16*K16 rows of P per CTA, Sp = 8*Ks = 128*K16 padded states);
// NT  = 8-column n-tiles per warp (a cluster owns 64*NT columns); grid = nsets * (2*K16/NT) clusters of 8.
//
// Contraction arithmetic: 3xTF32.  v = vh + vl and P = Ph + Pl with TF32 halves; vh*Ph + vl*Ph + vh*Pl on
// mma.sync.m16n8k8 with fp32 accumulation gives fp32-class accuracy over the full fp32 exponent range
// (a split-fp16 variant was measured and rejected: its 2^40 dynamic range loses the small entries of
// sparse messages).  Pl is kept as packed fp16 of 2^24*Pl (exactly representable in TF32 after
// conversion) to halve its register footprint; the 2^-24 is applied to the accumulator.
constexpr float kScalePl = 16777216.f, kUnscalePl = 1.f / 16777216.f;

// CL = CTAs per cluster: 8 (each warp contracts the whole K-slice for its n-tile(s)) or 4 ("half-K": the K-slice is twice as
// long, a cluster owns 32 columns, warp w contracts n-tile w & 3 over K-half w >> 2, so a warp issues half the MMAs and a CTA
// ships 1.5 KB instead of 7 KB of partials per step; 2 x 4 = 8 partial sources per column group either way).
//
// PK4 (single chain group, S in (512, 1024]): the message travels as 4-byte self-validating words in mma A-fragment order instead
// of 8-byte packets -- the format of the single-hop kernel below: every fp32 mantissa carries a 1-bit step stamp in its LSB (rounded
// to even at bit 1 first, cleared by the consumer: <= 1 ulp, unbiased), block exponents travel as {stamp16 | zero | E15} words.  A
// consumer CTA's K-slice is one contiguous 8 KB block [k8-step][lane][4] + 1 KB of exponent words, fetched by all 256 threads with
// fully coalesced 16-byte loads (2 per thread and round): 272 L2 sectors per CTA and poll round instead of 2048 (the 8-byte packets
// were read as 16-byte pieces 64 bytes apart, every sector twice).
template <int K16, int NT, int CL = 8, bool PK4 = false>
__global__ void __launch_bounds__(FLOW_THREADS, 1) hmm_flow_forward(FlowArgs p) {
    constexpr bool HALFK = CL == 4;
    static_assert(!PK4 || (CL == 8 && (K16 == 8 || K16 == 4)), "4-byte words: clusters of 8, 16 or 8 k8-steps per K-slice");
    constexpr int SP = K16 * 128;         // padded state count
    constexpr int KS = SP / CL;           // rows of P (= message entries) per CTA
    constexpr int NG = KS / 8;            // producer groups (8 columns each) in this CTA's K-slice
    constexpr int NKS = HALFK ? NG / 2 : NG;  // k8-steps contracted by one warp
    constexpr int NA = 8 * NG;            // threads that fetch packets: (chain pair ga, producer group pa)
    constexpr int CW = CL * 8 * NT;       // columns owned by a cluster
    constexpr int NCC = SP / CW;          // clusters per chain-group set
    static_assert(CL == 8 || (CL == 4 && NT == 1), "cluster of 8, or cluster of 4 with one n-tile per warp");
    constexpr int ROWB = 1024 + 16;       // bytes per k8-step row of A fragments: 32 lanes x (16 B hi + 16 B lo) + pad
    constexpr int C0 = FLOW_THREADS - 128 * NT;  // first final-owner thread
    static_assert(K16 >= 1 && K16 <= 16 && (K16 & (K16 - 1)) == 0, "K16 must be a power of two in [1,16]");
    static_assert(NT == 1 || NT == 2, "NT must be 1 or 2");
    static_assert(NA <= FLOW_THREADS && NG <= 32, "K-slice too long");

    extern __shared__ __align__(16) unsigned char s_frag_raw[];  // [2][NG][ROWB]
    unsigned char(*s_frag)[NG][ROWB] = reinterpret_cast<unsigned char(*)[NG][ROWB]>(s_frag_raw);
    __shared__ __align__(16) float s_part[2][FLOW_CLUSTER][NT][FLOW_CHAINS][8];
    __shared__ int s_eref[2][FLOW_CLUSTER][FLOW_CHAINS];
    __shared__ int s_erefloc[2][FLOW_CHAINS];
    __shared__ __align__(8) unsigned long long s_bar[2];
    __shared__ int s_abort;
    __shared__ float s_obs[FLOW_OBS_DEPTH][FLOW_THREADS];  // cp.async ring of obs[b, t, colC], private per thread
    // PK4: the slice's 16 block exponents per chain, as E15 << 17 (wrapping differences are then plain int subtractions), bit 0 set = zero
    // group (no value of E15 << 17 is free to serve as a marker: 0x4000 << 17 is INT_MIN); column 16 = E15 << 17 of group 0 whatever its
    // zero flag (the reference)
    __shared__ __align__(16) int s_ex[PK4 ? FLOW_CHAINS : 1][20];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tig = lane & 3;
    const int rank = (int)cluster_ctarank();
    const int cid = blockIdx.x / CL;
    const int set = cid / NCC, cc = cid % NCC;
    const int S = p.S;
    const int khalf = HALFK ? warp >> 2 : 0;            // which half of the K-slice this warp contracts
    const int wtile = HALFK ? (warp & 3) : warp;        // n-tile of this warp = cluster rank that finally owns its columns
    const int src_slot = HALFK ? rank * 2 + khalf : rank;  // this warp's slot among the 8 partial sources of a column group

    // bytes one step delivers into this CTA: 8 source CTAs x (NT tiles x 16 chains x 8 columns fp32 + 16 exponents)
    constexpr uint32_t TX_BYTES = FLOW_CLUSTER * (NT * FLOW_CHAINS * 8 * 4 + FLOW_CHAINS * 4);
    if (tid == 0) {
        mbar_init(smem_u32(&s_bar[0]), 1);
        mbar_init(smem_u32(&s_bar[1]), 1);
        s_abort = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_arm(smem_u32(&s_bar[0]), TX_BYTES);  // armed for steps 0 and 1; re-armed two steps ahead in phase C
        mbar_arm(smem_u32(&s_bar[1]), TX_BYTES);
    }

    // ---- resident operand: B fragments of P[k, col] = exp(A[k, col] - cA[col]); k-permutation inside a
    //      k8-step: fragment slot (tig, h) holds k = 2*tig + h
    uint32_t Ph[NT][NKS][2], Pl[NT][NKS];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int col = cc * CW + (HALFK ? 0 : nt * 64) + wtile * 8 + g;
        const float ca = col < S ? p.cA[col] : 0.f;
#pragma unroll
        for (int ks = 0; ks < NKS; ++ks) {
            float lo2[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = rank * KS + (khalf * NKS + ks) * 8 + 2 * tig + h;
                float v = 0.f;
                if (k < S && col < S) v = expf((p.transpose ? p.A[(int64_t)col * S + k] : p.A[(int64_t)k * S + col]) - ca);
                const uint32_t hi = cvt_tf32(v);
                Ph[nt][ks][h] = hi;
                lo2[h] = (v - __uint_as_float(hi)) * kScalePl;
            }
            if constexpr (PK4) {
                Pl[nt][ks] = pack_bf16(lo2[0] * kUnscalePl, lo2[1] * kUnscalePl);  // plain bf16 pair: B fragment half of the k16 side products
            } else {
                const __half2 l = __floats2half2_rn(lo2[0], lo2[1]);
                Pl[nt][ks] = *reinterpret_cast<const uint32_t*>(&l);
            }
        }
    }

    // ---- roles
    // phase A (tid < NA): chain pair (ga, ga+8), producer group pa of this CTA's K-slice
    const int ga = tid / NG, pa = tid % NG;
    // phase C (tid >= C0): final owner of column colC for chain bc
    const int tc = tid - C0;
    const int ntc = tc >> 7, bc = (tc >> 3) & 15, jc = tc & 7;
    const int colC = cc * CW + ntc * 64 + rank * 8 + jc;
    const float caC = (tc >= 0 && colC < S) ? p.cA[colC] : 0.f;
    const uint32_t bar_addr0 = smem_u32(&s_bar[0]);
    // remote addresses for the reduce-scatter: warp w sends its partials to cluster rank w
    const uint32_t r_part = mapa_u32(smem_u32(&s_part[0][src_slot][0][0][0]), (uint32_t)wtile);
    const uint32_t r_eref = mapa_u32(smem_u32(&s_eref[0][src_slot][0]), (uint32_t)wtile);
    const uint32_t r_bar0 = mapa_u32(bar_addr0, (uint32_t)wtile);

    cluster_sync_all();  // barriers initialised cluster-wide before any remote arrive

    unsigned long long* const pkt_set = p.pkt + (size_t)set * 2 * FLOW_CHAINS * SP;
    // PK4 layout inside the same allocation: data [parity][K-slice][k8-step][32 lanes][4] fp32, then exponents [parity][K-slice][k8-step][16 chains]
    uint32_t* const data_set = reinterpret_cast<uint32_t*>(pkt_set);
    uint32_t* const expo_set = data_set + 2 * FLOW_CLUSTER * NG * 128;
    // where the value of (chain bc, column colC) goes: K-slice colC / KS, k8-step (colC % KS) / 8, lane (bc & 7, jc >> 1), word (bc >> 3) + 2 (jc & 1)
    const int pub_grp = tc >= 0 ? (colC / KS) * NG + (colC % KS) / 8 : 0;
    const int pub_word = (((bc & 7) * 4 + (jc >> 1)) * 4) + (bc >> 3) + 2 * (jc & 1);
    auto publish4 = [&](unsigned t_stamp, float u, int E15, bool zero) {
        uint32_t bits = __float_as_uint(u);
        bits = ((bits & 3u) == 3u) ? bits + 1u : (bits & ~1u);  // round to even at bit 1: bit 0 is free for the stamp
        bits |= (t_stamp >> 1) & 1u;
        uint32_t* dst = data_set + ((size_t)(t_stamp & 1) * FLOW_CLUSTER * NG + pub_grp) * 128 + pub_word;
        if (p.mode & 1) atomicExch(dst, bits);
        else asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(dst), "r"(bits) : "memory");
        if (jc == 0) {
            const uint32_t w = ((t_stamp & 0xFFFFu) << 16) | (zero ? 0x8000u : 0u) | (unsigned)E15;
            uint32_t* de = expo_set + ((size_t)(t_stamp & 1) * FLOW_CLUSTER * NG + pub_grp) * 16 + bc;
            if (p.mode & 1) atomicExch(de, w);
            else asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(de), "r"(w) : "memory");
        }
    };
    long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long c_prev_top = 0;
    unsigned nstep = 0;  // steps executed by this CTA (drives barrier parity and smem buffer choice)
    unsigned tau = 0;    // packet stamp counter (monotone across chain groups)
    bool aborted = false;

    for (int grp = set; grp < p.groups_total && !aborted; grp += p.nsets) {
        const int64_t b0 = (int64_t)grp * FLOW_CHAINS;
        int E_prev = 0, E_abs = 0;
        bool zero_last = true;
        float u_last = 0.f;
        const bool chain_ok = tc >= 0 && (b0 + bc) < p.B && colC < S;
        const float* obs_c = (chain_ok && p.obs) ? p.obs + (((b0 + bc) / p.cpo) * p.obs_T) * S + colC : nullptr;
        auto obs_row = [&](int64_t step) { return p.reverse ? (p.T - 1 - step) : step; };
        int64_t chunk_lo = 0, chunk_hi = 0;  // rows [chunk_lo, chunk_hi) are known to have landed (no 64-bit division on the per-step path)
        auto wait_rows = [&](int64_t row) {  // overlapped H2D: block until the chunk holding `row` has landed
            if (!p.ready_flags) return;
            if (row >= chunk_lo && row < chunk_hi) return;
            const int64_t c = row / p.rows_per_chunk;
            const long long start = clock64();
            unsigned v;
            do {
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p.ready_flags + c) : "memory");
            } while (v == 0 && clock64() - start < 20 * FLOW_SPIN_LIMIT);
            if (v == 0) {  // the host-to-device copy of this chunk never landed: the rows about to be read are stale -> poison the pass
                if (atomicCAS(p.abort_flag + 1, 0, 4) == 0) {
                    p.abort_flag[2] = (int)blockIdx.x;
                    p.abort_flag[3] = tid;
                    p.abort_flag[4] = (int)row;
                }
                atomicExch(p.abort_flag, 1);
            }
            chunk_lo = c * p.rows_per_chunk;
            chunk_hi = chunk_lo + p.rows_per_chunk;
        };

        // ---- message before the first step: alpha_{-1} = init
        if (tc >= 0) {
            float x = -INFINITY;
            if (chain_ok) x = (p.init ? p.init[(b0 + bc) * p.init_sb + colC] : 0.f) * kLog2e;
            if (chain_ok && p.cpo > 1) x = (colC == (int)((b0 + bc) % p.cpo)) ? 0.f : -INFINITY;  // identity row
            int E15;
            bool zero;
            int dE;
            const float u = flow_normalise(1.f, x, 0, E15, zero, dE, false);  // |init| may be large (hand-over of alpha)
            E_abs = dE;
            E_prev = E15;
            u_last = u;
            zero_last = zero;
            const unsigned stamp = ((tau & 0xFFFFu) << 16) | (zero ? 0x8000u : 0u) | (unsigned)E15;
            if constexpr (PK4) publish4(tau, u, E15, zero);
            else
                st_pkt(pkt_set + ((size_t)(tau & 1) * FLOW_CHAINS + bc) * SP + colC,
                       ((unsigned long long)stamp << 32) | (unsigned long long)__float_as_uint(u));
        }
        // obs prefetch ring: one cp.async group per step, FLOW_OBS_DEPTH-1 steps ahead
        const uint32_t obs_ring = smem_u32(&s_obs[0][tid]);
#pragma unroll
        for (int q = 0; q < FLOW_OBS_DEPTH - 1; ++q) {
            if (obs_c && q < p.T) {
                wait_rows(obs_row(q));
                cp_async4(obs_ring + (uint32_t)q * (FLOW_THREADS * 4), obs_c + obs_row(q) * S);
            }
            cp_async_commit();
        }

        for (int64_t t = 0; t < p.T; ++t) {
            const int buf = nstep & 1, par = nstep & 1;
            const unsigned ph = (nstep >> 1) & 1;
            int timed_out = 0;
            const long long c_top = p.prof ? clock64() : 0;
            if (PK4 && p.prof && tid == FLOW_THREADS - 1) {  // PK4 report: the step period (top of this step - top of the previous one)
                if (t > 0) prof_acc[7] += c_top - c_prev_top;
                c_prev_top = c_top;
            }
            long long c_pkt = c_top, c_bar1 = 0;

            // ================= phase A: fetch this CTA's K-slice of the message =================
            if constexpr (PK4) {
                // thread tid < 16 NG holds lane `lane` of the A fragments of k8-steps warp and warp + NG / 2, and exponent word tid = (k8-step
                // tid >> 4, chain tid & 15); with 8 k8-steps per slice (S <= 512) warps 4-7 only take part in the barriers
                const uint32_t* dsl = data_set + ((size_t)(tau & 1) * FLOW_CLUSTER + rank) * NG * 128;
                const uint32_t* esl = expo_set + ((size_t)(tau & 1) * FLOW_CLUSTER + rank) * NG * 16;
                const uint32_t want_bit = (tau >> 1) & 1u, want16 = tau & 0xFFFFu;
                uint4 f0 = make_uint4(0, 0, 0, 0), f1 = f0;
                uint32_t ew = 0;
                const long long start = clock64();
                const bool stager = tid < NG * 16;
                auto fetch = [&](uint4& a0, uint4& a1, uint32_t& e) {
                    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a0.x), "=r"(a0.y), "=r"(a0.z), "=r"(a0.w) : "l"(dsl + tid * 4) : "memory");
                    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a1.x), "=r"(a1.y), "=r"(a1.z), "=r"(a1.w) : "l"(dsl + (tid + NG * 16) * 4) : "memory");
                    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(e) : "l"(esl + tid) : "memory");
                };
                auto valid = [&](const uint4& a0, const uint4& a1, uint32_t e) {
                    const uint32_t all_and = a0.x & a0.y & a0.z & a0.w & a1.x & a1.y & a1.z & a1.w;
                    const uint32_t all_or = a0.x | a0.y | a0.z | a0.w | a1.x | a1.y | a1.z | a1.w;
                    return ((want_bit ? all_and : ~all_or) & 1u) && (e >> 16) == want16;
                };
                // (Two probe sets kept in flight half a round trip apart were measured with this format too: no change, 3.26 vs 3.25 us/step.)
                if (stager)
                for (;;) {
                    fetch(f0, f1, ew);
                    if (valid(f0, f1, ew)) break;
                    if (clock64() - start > FLOW_SPIN_LIMIT) {
                        timed_out = 1;
                        if (atomicCAS(p.abort_flag + 1, 0, 1) == 0) {  // first timeout wins: leave a post-mortem record
                            p.abort_flag[2] = (int)blockIdx.x;
                            p.abort_flag[3] = tid;
                            p.abort_flag[4] = (int)t;
                            p.abort_flag[5] = (int)want16;
                            p.abort_flag[6] = (int)f0.x;
                            p.abort_flag[7] = (int)ew;
                        }
                        break;
                    }
                }
                if (p.prof) c_pkt = clock64();
                if (stager) s_ex[tid & 15][tid >> 4] = (int)(ew << 17) | (int)((ew >> 15) & 1u);
                if (tid < 16) s_ex[tid][16] = (int)(ew << 17);
                __syncthreads();
                if (p.prof) c_bar1 = clock64();
                if (stager) {
                // align the block exponents of the slice's 16 column groups, per chain.  Every warp: lane -> (chain lane & 15, half lane >> 4)
                // takes the maximum over 8 groups relative to group 0, the halves are merged by one shuffle, and a lane fetches (ref, max) of
                // the chains g and g + 8 of its fragments by shuffle (the first version let every thread scan all 16 words of both chains:
                // ~190 ALU instructions per thread, ~770 cycles of ALU issue per step).
                float sc[2][2];
                {
                    constexpr int GH = NG / 2;  // column groups per half-warp
                    const int chn = lane & 15, hf = lane >> 4;
                    const int ref_l = s_ex[chn][16];
                    const int own[2][2] = {{s_ex[g][warp], s_ex[g][warp + GH]}, {s_ex[g + 8][warp], s_ex[g + 8][warp + GH]}};
                    int m_l = INT_MIN;
#pragma unroll
                    for (int q4 = 0; q4 < GH / 4; ++q4) {
                        const int4 e4 = *reinterpret_cast<const int4*>(&s_ex[chn][GH * hf + 4 * q4]);
                        const int d0 = (e4.x & 1) ? INT_MIN : e4.x - ref_l, d1 = (e4.y & 1) ? INT_MIN : e4.y - ref_l;  // = sext15(E - ref) << 17
                        const int d2 = (e4.z & 1) ? INT_MIN : e4.z - ref_l, d3 = (e4.w & 1) ? INT_MIN : e4.w - ref_l;
                        m_l = max(m_l, max(max(d0, d1), max(d2, d3)));
                    }
                    m_l = max(m_l, __shfl_xor_sync(0xffffffffu, m_l, 16));
                    if (warp == 0 && lane < 16)
                        s_erefloc[buf][lane] = (m_l == INT_MIN) ? (0x8000 | (int)((unsigned)ref_l >> 17)) : (int)((unsigned)(ref_l + m_l) >> 17);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int ch = g + 8 * h;
                        const int ref = __shfl_sync(0xffffffffu, ref_l, ch), m = __shfl_sync(0xffffffffu, m_l, ch) >> 17;
                        sc[h][0] = (own[h][0] & 1) ? 0.f : pow2i(((own[h][0] - ref) >> 17) - m);
                        sc[h][1] = (own[h][1] & 1) ? 0.f : pow2i(((own[h][1] - ref) >> 17) - m);
                    }
                }
                // Staged per lane: the TF32 hi fragments of k8-steps warp and warp + 8, and for the k16 unit made of these two k8-steps
                // the bf16 A fragments of lo = v - hi and of v (rows g / g + 8, k slots (2 tig, 2 tig + 1) of either k8-step):
                //   v P = vh Ph (TF32, 2 mma.k8) + [vl Ph + vh Pl] (bf16, 2 mma.k16): 4 tensor instructions per 16 k instead of 6; the
                //   side products are 2^-11 of the result, so bf16's 2^-9 keeps them at 2^-20.
                const int swz = (g & 1) * 16;
                float v[2][4], lo[2][4];
                uint32_t hi[2][4];
#pragma unroll
                for (int f = 0; f < 2; ++f) {
                    const uint4 q = f ? f1 : f0;
                    v[f][0] = __uint_as_float(q.x & ~1u) * sc[0][f];
                    v[f][1] = __uint_as_float(q.y & ~1u) * sc[1][f];
                    v[f][2] = __uint_as_float(q.z & ~1u) * sc[0][f];
                    v[f][3] = __uint_as_float(q.w & ~1u) * sc[1][f];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        hi[f][c] = cvt_tf32(v[f][c]);
                        lo[f][c] = v[f][c] - __uint_as_float(hi[f][c]);
                    }
                }
                const uint4 bl = make_uint4(pack_bf16(lo[0][0], lo[0][2]), pack_bf16(lo[0][1], lo[0][3]), pack_bf16(lo[1][0], lo[1][2]),
                                            pack_bf16(lo[1][1], lo[1][3]));
                const uint4 bh = make_uint4(pack_bf16(v[0][0], v[0][2]), pack_bf16(v[0][1], v[0][3]), pack_bf16(v[1][0], v[1][2]),
                                            pack_bf16(v[1][1], v[1][3]));
                unsigned char* d0 = &s_frag[buf][warp][0] + lane * 32;
                unsigned char* d1 = &s_frag[buf][warp + NG / 2][0] + lane * 32;
                *reinterpret_cast<uint4*>(d0 + swz) = make_uint4(hi[0][0], hi[0][1], hi[0][2], hi[0][3]);
                *reinterpret_cast<uint4*>(d0 + (16 ^ swz)) = bl;
                *reinterpret_cast<uint4*>(d1 + swz) = make_uint4(hi[1][0], hi[1][1], hi[1][2], hi[1][3]);
                *reinterpret_cast<uint4*>(d1 + (16 ^ swz)) = bh;
                }
            } else if (tid < NA) {
                constexpr unsigned amask = NA >= 32 ? 0xffffffffu : ((1u << (NA & 31)) - 1u);
                const unsigned long long* src0 = pkt_set + ((size_t)(tau & 1) * FLOW_CHAINS + ga) * SP + rank * KS + pa * 8;
                const unsigned long long* src1 = src0 + (size_t)8 * SP;
                unsigned long long q0[8], q1[8];
                const unsigned want = tau & 0xFFFFu;
                const long long start = clock64();
                if (p.mode & 2) __nanosleep(500);
                if (p.mode & 4) __nanosleep(1000);
                for (;;) {
                    if (p.mode & 16) probe_wait(src1 + 6, want, start);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        ld_pkt2(src0 + 2 * i, q0[2 * i], q0[2 * i + 1]);
                        ld_pkt2(src1 + 2 * i, q1[2 * i], q1[2 * i + 1]);
                    }
                    bool ok = true;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        ok = ok && ((unsigned)(q0[i] >> 48) == want) && ((unsigned)(q1[i] >> 48) == want);
                    if (ok) break;
                    if (clock64() - start > FLOW_SPIN_LIMIT) {
                        timed_out = 1;
                        if (atomicCAS(p.abort_flag + 1, 0, 1) == 0) {  // first timeout wins: leave a post-mortem record
                            p.abort_flag[2] = (int)blockIdx.x;
                            p.abort_flag[3] = tid;
                            p.abort_flag[4] = (int)t;
                            p.abort_flag[5] = (int)want;
                            p.abort_flag[6] = (int)(q0[0] >> 32);
                            p.abort_flag[7] = (int)(q1[7] >> 32);
                        }
                        break;
                    }
                }
                if (p.prof) c_pkt = clock64();
                // align block exponents across the producer groups of the slice (lanes sharing ga are contiguous)
                const int e0 = (int)(q0[0] >> 32) & 0xFFFF, e1 = (int)(q1[0] >> 32) & 0xFFFF;
                const int base_lane = lane & ~(NG - 1);
                const int ref0 = __shfl_sync(amask, e0, base_lane) & 0x7FFF;
                const int ref1 = __shfl_sync(amask, e1, base_lane) & 0x7FFF;
                const int d0 = (e0 & 0x8000) ? -40000 : sext15(e0 - ref0);
                const int d1 = (e1 & 0x8000) ? -40000 : sext15(e1 - ref1);
                int m0 = d0, m1 = d1;
#pragma unroll
                for (int o = NG / 2; o > 0; o >>= 1) {
                    m0 = max(m0, __shfl_xor_sync(amask, m0, o));
                    m1 = max(m1, __shfl_xor_sync(amask, m1, o));
                }
                const float sc0 = (d0 == -40000) ? 0.f : pow2i(d0 - m0);
                const float sc1 = (d1 == -40000) ? 0.f : pow2i(d1 - m1);
                if (pa == 0) {
                    s_erefloc[buf][ga] = (m0 == -40000) ? (0x8000 | ref0) : ((ref0 + m0) & 0x7FFF);
                    s_erefloc[buf][ga + 8] = (m1 == -40000) ? (0x8000 | ref1) : ((ref1 + m1) & 0x7FFF);
                }
                // split to TF32 hi/lo and lay out as mma A fragments of k8-step pa: lane (ga, tig') gets
                // {v0[2tig'], v1[2tig'], v0[2tig'+1], v1[2tig'+1]} (chains ga / ga+8), 16 B hi + 16 B lo per
                // lane; the two halves are xor-swizzled with bit 2 of the lane so that the writes here and
                // the reads in phase B are bank-conflict free.
                unsigned char* row = &s_frag[buf][pa][0];
                const int swz = (ga & 1) * 16;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float v[4] = {__uint_as_float((unsigned)q0[2 * i]) * sc0, __uint_as_float((unsigned)q1[2 * i]) * sc1,
                                        __uint_as_float((unsigned)q0[2 * i + 1]) * sc0,
                                        __uint_as_float((unsigned)q1[2 * i + 1]) * sc1};
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        hi[c] = cvt_tf32(v[c]);
                        lo[c] = __float_as_uint(v[c] - __uint_as_float(hi[c]));  // exact and signed; the tensor core drops its low 13 bits (unbiased), one quarter-rate cvt less
                    }
                    unsigned char* dst = row + (ga * 4 + i) * 32;
                    *reinterpret_cast<uint4*>(dst + swz) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4*>(dst + (16 ^ swz)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
            }
            if (__syncthreads_or(timed_out | *(volatile int*)&s_abort)) {
                aborted = true;
                break;
            }

            const long long c_sync = p.prof ? clock64() : 0;
            // ================= phase B: partial contraction + reduce-scatter over DSMEM =================
            {
                float acc[NT][3][4];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int q = 0; q < 3; ++q)
#pragma unroll
                        for (int c = 0; c < 4; ++c) acc[nt][q][c] = 0.f;
                const int swz = ((lane >> 2) & 1) * 16;
                if constexpr (PK4) {
                    uint4 nh0, nbl, nh1, nbh;  // fragments of the next k16 unit travel while this one is contracted
                    {
                        const unsigned char* r0 = &s_frag[buf][0][0] + lane * 32;
                        const unsigned char* r1 = &s_frag[buf][NKS / 2][0] + lane * 32;
                        nh0 = *reinterpret_cast<const uint4*>(r0 + swz), nbl = *reinterpret_cast<const uint4*>(r0 + (16 ^ swz));
                        nh1 = *reinterpret_cast<const uint4*>(r1 + swz), nbh = *reinterpret_cast<const uint4*>(r1 + (16 ^ swz));
                    }
#pragma unroll
                    for (int j = 0; j < NKS / 2; ++j) {
                        const uint4 h0 = nh0, bl = nbl, h1 = nh1, bh = nbh;
                        if (j + 1 < NKS / 2) {
                            const unsigned char* r0 = &s_frag[buf][j + 1][0] + lane * 32;
                            const unsigned char* r1 = &s_frag[buf][j + 1 + NKS / 2][0] + lane * 32;
                            nh0 = *reinterpret_cast<const uint4*>(r0 + swz), nbl = *reinterpret_cast<const uint4*>(r0 + (16 ^ swz));
                            nh1 = *reinterpret_cast<const uint4*>(r1 + swz), nbh = *reinterpret_cast<const uint4*>(r1 + (16 ^ swz));
                        }
                        const uint32_t ah0[4] = {h0.x, h0.y, h0.z, h0.w}, ah1[4] = {h1.x, h1.y, h1.z, h1.w};
                        const uint32_t abl[4] = {bl.x, bl.y, bl.z, bl.w}, abh[4] = {bh.x, bh.y, bh.z, bh.w};
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt) mma_tf32(acc[nt][0], ah0, Ph[nt][j][0], Ph[nt][j][1]);
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt)  // bf16 B fragment of Ph: the upper halves of the TF32 registers (slots h = 0, 1 of either k8-step)
                            mma_bf16_k16(acc[nt][2], abl, upper_halves(Ph[nt][j][0], Ph[nt][j][1]),
                                         upper_halves(Ph[nt][j + NKS / 2][0], Ph[nt][j + NKS / 2][1]));
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt) mma_tf32(acc[nt][1], ah1, Ph[nt][j + NKS / 2][0], Ph[nt][j + NKS / 2][1]);
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt) mma_bf16_k16(acc[nt][2], abh, Pl[nt][j], Pl[nt][j + NKS / 2]);
                    }
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                        for (int c = 0; c < 4; ++c) acc[nt][2][c] *= kScalePl;  // the common epilogue below multiplies by 2^-24
                } else
#pragma unroll
                for (int ks = 0; ks < NKS; ++ks) {
                    const unsigned char* rowp = &s_frag[buf][khalf * NKS + ks][0] + lane * 32;
                    const uint4 fh = *reinterpret_cast<const uint4*>(rowp + swz);
                    const uint4 fl = *reinterpret_cast<const uint4*>(rowp + (16 ^ swz));
                    const uint32_t ah[4] = {fh.x, fh.y, fh.z, fh.w};
                    const uint32_t al[4] = {fl.x, fl.y, fl.z, fl.w};
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        const float2 pl = __half22float2(*reinterpret_cast<const __half2*>(&Pl[nt][ks]));
                        mma_tf32(acc[nt][0], ah, Ph[nt][ks][0], Ph[nt][ks][1]);
                        mma_tf32(acc[nt][1], al, Ph[nt][ks][0], Ph[nt][ks][1]);
                        mma_tf32(acc[nt][2], ah, __float_as_uint(pl.x), __float_as_uint(pl.y));
                    }
                }
                const uint32_t off = (uint32_t)par * (FLOW_CLUSTER * NT * FLOW_CHAINS * 8 * 4);
                const uint32_t r_bar = r_bar0 + (uint32_t)par * 8;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const uint32_t o2 = off + (uint32_t)nt * (FLOW_CHAINS * 8 * 4);
                    st_async_v2f(r_part + o2 + (uint32_t)((g * 8 + 2 * tig) * 4),
                                 acc[nt][0][0] + (acc[nt][1][0] + kUnscalePl * acc[nt][2][0]),
                                 acc[nt][0][1] + (acc[nt][1][1] + kUnscalePl * acc[nt][2][1]), r_bar);
                    st_async_v2f(r_part + o2 + (uint32_t)(((g + 8) * 8 + 2 * tig) * 4),
                                 acc[nt][0][2] + (acc[nt][1][2] + kUnscalePl * acc[nt][2][2]),
                                 acc[nt][0][3] + (acc[nt][1][3] + kUnscalePl * acc[nt][2][3]), r_bar);
                }
                if (lane < FLOW_CHAINS)
                    st_async_u32(r_eref + (uint32_t)par * (FLOW_CLUSTER * FLOW_CHAINS * 4) + (uint32_t)lane * 4,
                                 (uint32_t)s_erefloc[buf][lane], r_bar);
            }
            const long long c_sent = p.prof ? clock64() : 0;
            if (p.prof && tid == 0) {
                prof_acc[0] += c_pkt - c_top;    // waiting for / loading packets
                prof_acc[1] += c_sync - c_pkt;   // exponent alignment, split, STS, bar.sync
                prof_acc[2] += c_sent - c_sync;  // LDS + mma + DSMEM stores + arrive
                if (c_bar1) prof_acc[6] += c_bar1 - c_pkt;  // PK4: own words valid -> every thread's words valid (barrier 1)
            }

            // ================= phase C: final owners finish their columns and publish =================
            if (tc >= 0) {
                // observation-only work first: it overlaps the flight time of the partial sums
                cp_async_wait<FLOW_OBS_DEPTH - 2>();
                const float ob = obs_c ? s_obs[t % FLOW_OBS_DEPTH][tid] : 0.f;
                const FlowObs1 oterm = flow_obs1(chain_ok ? (ob + caC) * kLog2e : -INFINITY, true);
                const long long start = clock64();
                auto wait_bar = [&](uint32_t addr, int kind) {
                    while (!mbar_try_wait(addr, ph)) {
                        if (clock64() - start > FLOW_SPIN_LIMIT) {
                            s_abort = 1;
                            if (atomicCAS(p.abort_flag + 1, 0, kind) == 0) {
                                p.abort_flag[2] = (int)blockIdx.x;
                                p.abort_flag[3] = tid;
                                p.abort_flag[4] = (int)t;
                            }
                            break;
                        }
                    }
                };
                // The 8 lanes (jc = 0..7) that share a chain split the 8 sources between them: lane jc aligns source jc, the maximum is an
                // 8-lane butterfly and the weights are handed round by shuffle (every thread scanning all 8 exponent words cost ~110 ALU
                // instructions per thread: ~450 cycles of ALU issue per step).  (Sending the exponents ahead of the contraction under their own
                // mbarrier, so that the weights are ready when the partial sums land, was measured: no change in the step time.)
                wait_bar(bar_addr0 + (uint32_t)par * 8, 2);
                const long long c_got = p.prof ? clock64() : 0;
                const int e_mine = s_eref[par][jc][bc];
                const int e0 = __shfl_sync(0xffffffffu, e_mine, lane & ~7) & 0x7FFF;
                const int d_mine = (e_mine & 0x8000) ? -40000 : sext15(e_mine - e0);
                int dmax = d_mine;
                dmax = max(dmax, __shfl_xor_sync(0xffffffffu, dmax, 4));
                dmax = max(dmax, __shfl_xor_sync(0xffffffffu, dmax, 2));
                dmax = max(dmax, __shfl_xor_sync(0xffffffffu, dmax, 1));
                const float w_mine = (d_mine == -40000) ? 0.f : pow2i(d_mine - dmax);
                float wr[FLOW_CLUSTER];
#pragma unroll
                for (int r = 0; r < FLOW_CLUSTER; ++r) wr[r] = __shfl_sync(0xffffffffu, w_mine, (lane & ~7) + r);
                const int emax = (dmax == -40000) ? e0 : ((e0 + dmax) & 0x7FFF);
                float s = 0.f;
#pragma unroll
                for (int r = 0; r < FLOW_CLUSTER; ++r) s += s_part[par][r][ntc][bc][jc] * wr[r];
                // this phase is complete (all threads of the CTA that wait on it see the same parity); arm the
                // barriers for the step after next.  No sender can be two steps ahead of this CTA.
                if (tc == 0) mbar_arm(bar_addr0 + (uint32_t)par * 8, TX_BYTES);
                const long long c_red = p.prof ? clock64() : 0;
                int E15;
                bool zero;
                const long long c_obs = p.prof ? clock64() : 0;
                int dE;
                const float u = flow_finish1(s, oterm, emax, E15, zero, dE);
                const long long c_nrm = p.prof ? clock64() : 0;
                E_abs += sext15(E15 - E_prev);
                E_prev = E15;
                u_last = u;
                zero_last = zero;
                const unsigned tn = tau + 1;
                const unsigned stamp = ((tn & 0xFFFFu) << 16) | (zero ? 0x8000u : 0u) | (unsigned)E15;
                if constexpr (PK4)
                    publish4(tn, u, E15, zero);
                else if (p.mode & 1)
                    atomicExch(pkt_set + ((size_t)(tn & 1) * FLOW_CHAINS + bc) * SP + colC,
                               ((unsigned long long)stamp << 32) | (unsigned long long)__float_as_uint(u));
                else
                    st_pkt(pkt_set + ((size_t)(tn & 1) * FLOW_CHAINS + bc) * SP + colC,
                           ((unsigned long long)stamp << 32) | (unsigned long long)__float_as_uint(u));
                if (p.hist_u && (b0 + bc) < p.B) {
                    const int64_t slot = (b0 + bc) * p.hist_T + obs_row(t);
                    p.hist_u[slot * SP + colC] = u;
                    if (jc == 0) p.hist_e[slot * (SP / 8) + colC / 8] = zero ? INT_MIN : E_abs;
                }
                // refill the slot consumed one step ago
                {
                    const int64_t tn2 = t + FLOW_OBS_DEPTH - 1;
                    if (obs_c && tn2 < p.T) {
                        wait_rows(obs_row(tn2));
                        cp_async4(obs_ring + (uint32_t)(tn2 % FLOW_OBS_DEPTH) * (FLOW_THREADS * 4), obs_c + obs_row(tn2) * S);
                    }
                    cp_async_commit();
                }
                if (p.prof && tid == FLOW_THREADS - 1) {
                    prof_acc[3] += c_got - c_sent;       // waiting for the 8 partials of my columns
                    prof_acc[4] += clock64() - c_got;    // reduce, exp2, normalise, publish
                    prof_acc[5] += PK4 ? c_sent - c_sync : c_red - c_got;  // PK4 report: contraction + send of the LAST warp
                    prof_acc[6] += c_obs - c_red;
                    if (!PK4) prof_acc[7] += c_nrm - c_obs;
                }
            }
            ++nstep;
            ++tau;
        }
        cp_async_wait<0>();
        // ---- final message of the group (alpha_{T-1}) for the log-likelihood reduction
        if (tc >= 0 && (b0 + bc) < p.B) {
            p.fin_u[(b0 + bc) * SP + colC] = u_last;
            if (jc == 0) p.fin_e[(b0 + bc) * (SP / 8) + colC / 8] = zero_last ? INT_MIN : E_abs;
        }
        // The next group's initial message overwrites this group's last message (same buffer parity, which
        // nobody reads any more) under a stamp that differs from it.
        tau += 2;
    }
    if (p.prof) {
        if (tid == 0) {
            for (int q = 0; q < 3; ++q) p.prof[(size_t)blockIdx.x * 8 + q] = prof_acc[q];
            if (PK4) p.prof[(size_t)blockIdx.x * 8 + 7] = prof_acc[6];  // PK4: the "obs" column of the report is the wait at barrier 1
        }
        if (tid == FLOW_THREADS - 1) {
            p.prof[(size_t)blockIdx.x * 8 + 3] = prof_acc[3];
            p.prof[(size_t)blockIdx.x * 8 + 4] = prof_acc[4];
            p.prof[(size_t)blockIdx.x * 8 + 5] = (long long)nstep;
            p.prof[(size_t)blockIdx.x * 8 + 6] = prof_acc[5];
            if (!PK4) p.prof[(size_t)blockIdx.x * 8 + 7] = prof_acc[6];
            p.prof[(size_t)blockIdx.x * 8 + 2] += 0;
            p.prof[(size_t)(FLOW_PROF_CTAS / 2 + blockIdx.x / 2) * 8 + (blockIdx.x & 1)] = prof_acc[7];
        }
    }
    if (aborted && tid == 0) atomicExch(p.abort_flag, 1);
    cluster_sync_all();  // nobody may exit while a peer can still write into its shared memory
}

// ------------------------------------------------------------------------------------------ single-hop variant
// hmm_direct_forward<KW>: the same recurrence with ONE chip-wide hand-over per step instead of two.
//
// CTA c of a set owns the 8 columns [8c, 8c+8) and keeps the FULL-K block P[:, 8c:8c+8] resident (K = Sp rows split over
// its 8 warps, KW = Sp/64 k8-steps each: 3 KW registers per thread).  The 8 columns a CTA produces are exactly one k8-step
// of every consumer's contraction, so the producer writes them straight in mma A-fragment order (32 lanes x 16 B) and every
// consumer warp loads its KW fragments with 16-byte loads: no shared-memory staging, no DSMEM reduce-scatter; the K-reduction
// happens between the 8 warps of a CTA through shared memory.  L2 read traffic rises to 16 Sp 4 B per CTA-step (8 MB per
// step chip-wide at S = 1024, all hits), which is what buys the removal of the second hop (measured step budget of the
// two-hop kernel: packet wait 1450 + DSMEM partial wait 1100 of 6400 cycles).
//
// Self-validating words: every fp32 mantissa carries a 1-bit step stamp in its LSB (the value is first rounded to even at
// that bit, the consumer clears it: <= 1 ulp, unbiased); block exponents travel as {stamp16 | zero | E15} words.  Stale
// content of a slot is always two steps old, whose stamp bit differs.  Each 4-byte word validates itself, so no 16-byte
// atomicity is assumed.  Chain groups beyond the first are separated by a grid-wide barrier (cooperative launch).
template <int KW>
__global__ void __launch_bounds__(FLOW_THREADS, 1) hmm_direct_forward(FlowArgs p) {
    constexpr int SP = KW * 64;
    constexpr int NC = SP / 8;  // CTAs per set = k8-steps of the contraction
    constexpr int NW = 8;       // warps
    __shared__ __align__(16) float s_part[2][NW][FLOW_CHAINS][8];
    __shared__ int s_eref[2][NW][FLOW_CHAINS];
    __shared__ float s_obs[FLOW_OBS_DEPTH][128];
    __shared__ int s_abort;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tig = lane & 3;
    const int set = blockIdx.x / NC, c = blockIdx.x % NC;
    const int S = p.S;
    uint32_t* const data_set = reinterpret_cast<uint32_t*>(p.pkt) + (size_t)set * (2 * NC * 128 + 2 * NC * 16);
    uint32_t* const expo_set = data_set + 2 * NC * 128;
    if (tid == 0) s_abort = 0;

    // resident operand: B fragments of P[k, col] = exp(A[k, col] - cA[col]), col = 8c + g, k8-step (warp, ks); slot (tig, h) <-> k = 2 tig + h
    uint32_t Ph[KW][2], Pl[KW];
    {
        const int col = c * 8 + g;
        const float ca = col < S ? p.cA[col] : 0.f;
#pragma unroll
        for (int ks = 0; ks < KW; ++ks) {
            float lo2[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = (warp * KW + ks) * 8 + 2 * tig + h;
                float v = 0.f;
                if (k < S && col < S) v = expf((p.transpose ? p.A[(int64_t)col * S + k] : p.A[(int64_t)k * S + col]) - ca);
                const uint32_t hi = cvt_tf32(v);
                Ph[ks][h] = hi;
                lo2[h] = (v - __uint_as_float(hi)) * kScalePl;
            }
            const __half2 l = __floats2half2_rn(lo2[0], lo2[1]);
            Pl[ks] = *reinterpret_cast<const uint32_t*>(&l);
        }
    }
    // final-owner role (tid < 128): chain bc, column jc of this CTA
    const bool owner = tid < 128;
    const int bc = (tid >> 3) & 15, jc = tid & 7;
    const int colC = c * 8 + jc;
    const float caC = (owner && colC < S) ? p.cA[colC] : 0.f;
    // where this thread's value goes in the consumers' fragment: lane (bc & 7, jc >> 1), slot (bc >> 3) + 2 (jc & 1)
    const int pub_word = (((bc & 7) * 4 + (jc >> 1)) * 4) + (bc >> 3) + 2 * (jc & 1);
    __syncthreads();

    unsigned tau = 0;
    unsigned nstep = 0;
    bool aborted = false;
    const int ngroups_iter = (p.groups_total + p.nsets - 1) / p.nsets;
    for (int it = 0; it < ngroups_iter; ++it) {
        const int grp = it * p.nsets + set;
        const bool grp_ok = grp < p.groups_total;
        if (grp_ok && !aborted) {
            const int64_t b0 = (int64_t)grp * FLOW_CHAINS;
            int E_prev = 0, E_abs = 0;
            bool zero_last = true;
            float u_last = 0.f;
            const bool chain_ok = owner && (b0 + bc) < p.B && colC < S;
            const float* obs_c = (chain_ok && p.obs) ? p.obs + (((b0 + bc) / p.cpo) * p.obs_T) * S + colC : nullptr;
            auto obs_row = [&](int64_t step) { return p.reverse ? (p.T - 1 - step) : step; };
            int64_t chunk_lo = 0, chunk_hi = 0;  // rows [chunk_lo, chunk_hi) are known to have landed
            auto wait_rows = [&](int64_t row) {
                if (!p.ready_flags) return;
                if (row >= chunk_lo && row < chunk_hi) return;
                const int64_t ch = row / p.rows_per_chunk;
                const long long start = clock64();
                unsigned v;
                do {
                    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p.ready_flags + ch) : "memory");
                } while (v == 0 && clock64() - start < 20 * FLOW_SPIN_LIMIT);
                if (v == 0) {  // see hmm_flow_forward: a chunk that never landed poisons the pass instead of being read stale
                    if (atomicCAS(p.abort_flag + 1, 0, 4) == 0) {
                        p.abort_flag[2] = (int)blockIdx.x;
                        p.abort_flag[3] = tid;
                        p.abort_flag[4] = (int)row;
                    }
                    atomicExch(p.abort_flag, 1);
                }
                chunk_lo = ch * p.rows_per_chunk;
                chunk_hi = chunk_lo + p.rows_per_chunk;
            };
            auto publish = [&](unsigned t_stamp, float u, int E15, bool zero) {
                uint32_t bits = __float_as_uint(u);
                bits = ((bits & 3u) == 3u) ? bits + 1u : (bits & ~1u);  // round to even at bit 1: bit 0 is free for the stamp
                bits |= (t_stamp >> 1) & 1u;
                uint32_t* dst = data_set + ((size_t)(t_stamp & 1) * NC + c) * 128 + pub_word;
                asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(dst), "r"(bits) : "memory");
                if (jc == 0) {
                    const uint32_t w = ((t_stamp & 0xFFFFu) << 16) | (zero ? 0x8000u : 0u) | (unsigned)E15;
                    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(expo_set + ((size_t)(t_stamp & 1) * NC + c) * 16 + bc), "r"(w)
                                 : "memory");
                }
            };
            // ---- message before the first step: alpha_{-1} = init
            if (owner) {
                float x = -INFINITY;
                if (chain_ok) x = (p.init ? p.init[(b0 + bc) * p.init_sb + colC] : 0.f) * kLog2e;
                if (chain_ok && p.cpo > 1) x = (colC == (int)((b0 + bc) % p.cpo)) ? 0.f : -INFINITY;
                int E15, dE;
                bool zero;
                const float u = flow_normalise(1.f, x, 0, E15, zero, dE, false);
                E_abs = dE;
                E_prev = E15;
                u_last = u;
                zero_last = zero;
                publish(tau, u, E15, zero);
            }
            const uint32_t obs_ring = smem_u32(&s_obs[0][tid & 127]);
            if (owner) {
#pragma unroll
                for (int q = 0; q < FLOW_OBS_DEPTH - 1; ++q) {
                    if (obs_c && q < p.T) {
                        wait_rows(obs_row(q));
                        cp_async4(obs_ring + (uint32_t)q * (128 * 4), obs_c + obs_row(q) * S);
                    }
                    cp_async_commit();
                }
            }

            for (int64_t t = 0; t < p.T; ++t) {
                const int par = nstep & 1;
                int timed_out = 0;
                // ================= fetch the KW fragments of this warp's K-range =================
                uint4 q[KW];
                uint32_t ew[KW];
                {
                    const uint32_t* dsrc = data_set + ((size_t)(tau & 1) * NC + warp * KW) * 128 + lane * 4;
                    const uint32_t* esrc = expo_set + ((size_t)(tau & 1) * NC + warp * KW) * 16 + ((tig & 1) ? g + 8 : g);
                    const uint32_t want_bit = (tau >> 1) & 1u, want16 = tau & 0xFFFFu;
                    const long long start = clock64();
                    // Spin on the (small) exponent words only; the 16-byte fragments are requested once all KW producer groups
                    // of this warp's K-range have published (polling the fragments themselves multiplied the L2 traffic).
                    // KW <= 4: the fragments are polled together with the exponent words (one round trip; measured 1.32 vs
                    // 1.58 us/step at S = 128).
                    for (;;) {
                        bool ok = true;
                        if constexpr (KW >= FLOW_EXPO_FIRST_KW) {
#pragma unroll
                            for (int ks = 0; ks < KW; ++ks)
                                asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(ew[ks]) : "l"(esrc + ks * 16) : "memory");
#pragma unroll
                            for (int ks = 0; ks < KW; ++ks) ok = ok && ((ew[ks] >> 16) == want16);
                            if (__all_sync(0xffffffffu, ok)) {
#pragma unroll
                                for (int ks = 0; ks < KW; ++ks)
                                    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                                                 : "=r"(q[ks].x), "=r"(q[ks].y), "=r"(q[ks].z), "=r"(q[ks].w)
                                                 : "l"(dsrc + ks * 128)
                                                 : "memory");
#pragma unroll
                                for (int ks = 0; ks < KW; ++ks) {
                                    const uint32_t all_and = q[ks].x & q[ks].y & q[ks].z & q[ks].w, all_or = q[ks].x | q[ks].y | q[ks].z | q[ks].w;
                                    ok = ok && ((want_bit ? all_and : ~all_or) & 1u);
                                }
                                if (__all_sync(0xffffffffu, ok)) break;
                            }
                        } else {
#pragma unroll
                            for (int ks = 0; ks < KW; ++ks) {
                                asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                                             : "=r"(q[ks].x), "=r"(q[ks].y), "=r"(q[ks].z), "=r"(q[ks].w)
                                             : "l"(dsrc + ks * 128)
                                             : "memory");
                                asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(ew[ks]) : "l"(esrc + ks * 16) : "memory");
                            }
#pragma unroll
                            for (int ks = 0; ks < KW; ++ks) {
                                const uint32_t all_and = q[ks].x & q[ks].y & q[ks].z & q[ks].w, all_or = q[ks].x | q[ks].y | q[ks].z | q[ks].w;
                                ok = ok && ((want_bit ? all_and : ~all_or) & 1u) && ((ew[ks] >> 16) == want16);
                            }
                            if (__all_sync(0xffffffffu, ok)) break;
                        }
                        if (clock64() - start > FLOW_SPIN_LIMIT) {
                            timed_out = 1;
                            if (lane == 0 && atomicCAS(p.abort_flag + 1, 0, 3) == 0) {
                                p.abort_flag[2] = (int)blockIdx.x;
                                p.abort_flag[3] = tid;
                                p.abort_flag[4] = (int)t;
                                p.abort_flag[5] = (int)want16;
                                p.abort_flag[6] = (int)q[0].x;
                                p.abort_flag[7] = (int)ew[0];
                            }
                            break;
                        }
                    }
                }
                // ================= align block exponents over the K-range, contract =================
                {
                    int e0[KW], e1[KW];  // exponent words of chains g and g+8
#pragma unroll
                    for (int ks = 0; ks < KW; ++ks) {
                        const uint32_t other = __shfl_xor_sync(0xffffffffu, ew[ks], 1);
                        e0[ks] = (int)(((tig & 1) ? other : ew[ks]) & 0xFFFFu);
                        e1[ks] = (int)(((tig & 1) ? ew[ks] : other) & 0xFFFFu);
                    }
                    const int ref0 = e0[0] & 0x7FFF, ref1 = e1[0] & 0x7FFF;
                    int m0 = -40000, m1 = -40000;
#pragma unroll
                    for (int ks = 0; ks < KW; ++ks) {
                        e0[ks] = (e0[ks] & 0x8000) ? -40000 : sext15(e0[ks] - ref0);
                        e1[ks] = (e1[ks] & 0x8000) ? -40000 : sext15(e1[ks] - ref1);
                        m0 = max(m0, e0[ks]);
                        m1 = max(m1, e1[ks]);
                    }
                    float acc[3][4];
#pragma unroll
                    for (int a = 0; a < 3; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
#pragma unroll
                    for (int ks = 0; ks < KW; ++ks) {
                        const float sc0 = (e0[ks] == -40000) ? 0.f : pow2i(e0[ks] - m0);
                        const float sc1 = (e1[ks] == -40000) ? 0.f : pow2i(e1[ks] - m1);
                        const float v[4] = {__uint_as_float(q[ks].x & ~1u) * sc0, __uint_as_float(q[ks].y & ~1u) * sc1,
                                            __uint_as_float(q[ks].z & ~1u) * sc0, __uint_as_float(q[ks].w & ~1u) * sc1};
                        uint32_t ah[4], al[4];
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            ah[b] = cvt_tf32(v[b]);
                            al[b] = __float_as_uint(v[b] - __uint_as_float(ah[b]));  // exact, signed; the tensor core drops its low 13 bits (unbiased)
                        }
                        const float2 pl = __half22float2(*reinterpret_cast<const __half2*>(&Pl[ks]));
                        mma_tf32(acc[0], ah, Ph[ks][0], Ph[ks][1]);
                        mma_tf32(acc[1], al, Ph[ks][0], Ph[ks][1]);
                        mma_tf32(acc[2], ah, __float_as_uint(pl.x), __float_as_uint(pl.y));
                    }
                    float* dst = &s_part[par][warp][0][0];
                    *reinterpret_cast<float2*>(dst + g * 8 + 2 * tig) =
                        make_float2(acc[0][0] + (acc[1][0] + kUnscalePl * acc[2][0]), acc[0][1] + (acc[1][1] + kUnscalePl * acc[2][1]));
                    *reinterpret_cast<float2*>(dst + (g + 8) * 8 + 2 * tig) =
                        make_float2(acc[0][2] + (acc[1][2] + kUnscalePl * acc[2][2]), acc[0][3] + (acc[1][3] + kUnscalePl * acc[2][3]));
                    if (tig == 0) {
                        s_eref[par][warp][g] = (m0 == -40000) ? (0x8000 | ref0) : ((ref0 + m0) & 0x7FFF);
                        s_eref[par][warp][g + 8] = (m1 == -40000) ? (0x8000 | ref1) : ((ref1 + m1) & 0x7FFF);
                    }
                }
                // observation-only work before the barrier (owners)
                FlowObs1 oterm;
                if (owner) {
                    cp_async_wait<FLOW_OBS_DEPTH - 2>();
                    const float ob = obs_c ? s_obs[t % FLOW_OBS_DEPTH][tid] : 0.f;
                    oterm = flow_obs1(chain_ok ? (ob + caC) * kLog2e : -INFINITY, true);
                }
                if (__syncthreads_or(timed_out | *(volatile int*)&s_abort)) {
                    aborted = true;
                    break;
                }
                // ================= owners: reduce the 8 warp partials, finish, publish =================
                if (owner) {
                    const int er0 = s_eref[par][0][bc] & 0x7FFF;
                    int dr[NW], dmax = -40000;
#pragma unroll
                    for (int r = 0; r < NW; ++r) {
                        const int e = s_eref[par][r][bc];
                        dr[r] = (e & 0x8000) ? -40000 : sext15(e - er0);
                        dmax = max(dmax, dr[r]);
                    }
                    float s = 0.f;
#pragma unroll
                    for (int r = 0; r < NW; ++r) s += (dr[r] == -40000) ? 0.f : s_part[par][r][bc][jc] * pow2i(dr[r] - dmax);
                    const int emax = (dmax == -40000) ? er0 : ((er0 + dmax) & 0x7FFF);
                    int E15, dE;
                    bool zero;
                    const float u = flow_finish1(s, oterm, emax, E15, zero, dE);
                    E_abs += sext15(E15 - E_prev);
                    E_prev = E15;
                    u_last = u;
                    zero_last = zero;
                    publish(tau + 1, u, E15, zero);
                    if (p.hist_u && (b0 + bc) < p.B) {
                        const int64_t slot = (b0 + bc) * p.hist_T + obs_row(t);
                        p.hist_u[slot * SP + colC] = u;
                        if (jc == 0) p.hist_e[slot * (SP / 8) + colC / 8] = zero ? INT_MIN : E_abs;
                    }
                    const int64_t tn2 = t + FLOW_OBS_DEPTH - 1;
                    if (obs_c && tn2 < p.T) {
                        wait_rows(obs_row(tn2));
                        cp_async4(obs_ring + (uint32_t)(tn2 % FLOW_OBS_DEPTH) * (128 * 4), obs_c + obs_row(tn2) * S);
                    }
                    cp_async_commit();
                }
                ++nstep;
                ++tau;
            }
            cp_async_wait<0>();
            if (owner && (b0 + bc) < p.B && !aborted) {
                p.fin_u[(b0 + bc) * SP + colC] = u_last;
                if (jc == 0) p.fin_e[(b0 + bc) * (SP / 8) + colC / 8] = zero_last ? INT_MIN : E_abs;
            }
        }
        if (ngroups_iter > 1) {
            // every CTA of the grid is done with this group's packets before the next group's initial message may overwrite them
            __syncthreads();
            cooperative_groups::this_grid().sync();
            // slots now hold stamps tau (unread) and tau - 1; continuing with tau + 1 keeps every overwrite two stamps apart
            if (grp_ok && !aborted) tau += 1;
            else tau = 0;
        }
    }
    if (aborted && tid == 0) atomicExch(p.abort_flag, 1);
}

// ------------------------------------------------------------------------------------------ tcgen05 variant
// Same dataflow, but the per-step contraction runs on the 5th-generation tensor cores:
//   D[m, n] (TMEM, 128 lanes x 16 columns fp32) += P^T[m, k] (TMEM, resident for the whole pass) * v[n, k] (SMEM)
// with m = the 128 columns of P owned by the cluster, n = the 16 chains, k = this CTA's K-slice, kind::tf32,
// M=128, N=16, K=8 per instruction, 3xTF32 (Ph*vh + Ph*vl + Pl*vh accumulate into one D).  One elected thread
// issues the 3*KSTEPS MMAs and commits to an mbarrier; warps 4-7 read D with tcgen05.ld (one lane = one column of
// P) and push their 16 chain values straight to the column's final owner with st.async.  The legacy
// HMMA.1688.TF32 pipe needed ~1900 cycles per step for the same product (measured); this needs a few hundred.
//
// KSTEPS = k8-steps per CTA (Ks = 8*KSTEPS rows, Sp = 64*KSTEPS padded states, Sp/128 clusters of 8 per set).
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void st_async_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t mbar) {
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(addr), "r"(a),
                 "r"(b), "r"(c), "r"(d), "r"(mbar)
                 : "memory");
}
// K-major, no-swizzle shared-memory operand: core matrix = 8 rows x 16 B contiguous; SBO = 128 B between the two
// 8-row groups of the 16 chains, LBO = 256 B between the two 16-byte k-chunks of one K=8 instruction.
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(256u >> 4) << 16) | ((uint64_t)(128u >> 4) << 32) | (1ull << 46);
}

// Issue the 3*KSTEPS MMAs of one step.  Everything the instructions need is a compile-time constant or the
// (warp-uniform) dynamic shared memory base: the tensor-memory allocation is the whole 512 columns, so its base
// is 0 (checked once after the allocation), and BUF selects the B tiles / accumulators.  That keeps the operands
// in uniform registers; an earlier version that computed them from per-thread registers inside `if (tid == 0)`
// paid ~64 cycles per MMA in R2UR / elect loops (measured).
template <int KSTEPS, int NACC, int BUF>
__device__ __forceinline__ void tc_issue_step(uint32_t smem_base, uint32_t mma_bar) {
    constexpr int KS = KSTEPS * 8;
    constexpr int STEPB = 512 + 16;
    constexpr int TILEB = KSTEPS * STEPB;
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
    constexpr uint32_t D0 = (uint32_t)(2 * KS + NACC * 16 * BUF);
    const uint32_t bh = smem_base + (uint32_t)((BUF * 2 + 0) * TILEB), bl = smem_base + (uint32_t)((BUF * 2 + 1) * TILEB);
    uint32_t elected;
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\tselp.u32 %0, 1, 0, e;\n\t}" : "=r"(elected));
    if (elected) {
#pragma unroll
        for (int i = 0; i < 3 * KSTEPS; ++i) {
            const int pass = i / KSTEPS, ks = i % KSTEPS;
            const uint32_t a = (uint32_t)((pass == 2 ? KS : 0) + ks * 8);
            const uint32_t b = (pass == 1 ? bl : bh) + (uint32_t)(ks * STEPB);
            tc_mma_tf32_ts(D0 + (uint32_t)((i % NACC) * 16), a, tc_smem_desc(b), IDESC, i >= NACC ? 1u : 0u);
        }
        tc_commit(mma_bar);
    }
    __syncwarp();
}

__device__ __forceinline__ void st_pkt2(unsigned long long* p, unsigned long long a, unsigned long long b) {
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const float* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// Normalise one column group held by a PAIR of lanes (lane, lane^16), 4 columns each:
//   true value = s[i] * 2^(x[i]) * 2^(e_in)   ->   u[i] * 2^(E)
// Split in two so that everything that depends only on the observations (group max of x, exp2) is done
// BEFORE the thread waits for the partial sums; only a short dependent chain remains after the wait.
struct FlowObsTerm {
    float e2[4];  // 2^(x[i] - nf)
    int nf;       // ceil(group max of x)
    int bad;      // per-step magnitude outside the wrapped-exponent range
};
__device__ __forceinline__ FlowObsTerm flow_obs_term(const float (&x)[4], bool check_range) {
    FlowObsTerm o;
    float xm = fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3]));
    xm = fmaxf(xm, __shfl_xor_sync(0xffffffffu, xm, 16));
    float nf = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) o.e2[i] = 0.f;
    if (xm > -1e30f) {
        nf = ceilf(xm);
#pragma unroll
        for (int i = 0; i < 4; ++i) o.e2[i] = exp2f(x[i] - nf);
    }
    o.bad = (check_range && xm > -1e30f && !(fabsf(nf) <= 8000.f)) ? 1 : 0;
    o.nf = (int)nf;
    return o;
}
__device__ __forceinline__ void flow_finish4(const float (&s)[4], const FlowObsTerm& o, int e_in15, float (&u)[4], int& E15, bool& zero,
                                             int& dE) {
    int bad = o.bad;
    float mu = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        u[i] = s[i] * o.e2[i];
        bad |= !(u[i] <= FLT_MAX) ? 1 : 0;  // NaN or +inf
        mu = fmaxf(mu, u[i]);
    }
    mu = fmaxf(mu, __shfl_xor_sync(0xffffffffu, mu, 16));
    bad |= __shfl_xor_sync(0xffffffffu, bad, 16);
    dE = 0;
    E15 = e_in15 & 0x7FFF;
    zero = !(mu > 0.f);
    if (bad) {  // poison the group: the log-likelihood becomes NaN
        zero = false;
#pragma unroll
        for (int i = 0; i < 4; ++i) u[i] = __int_as_float(0x7fc00000);
        return;
    }
    if (zero) {
#pragma unroll
        for (int i = 0; i < 4; ++i) u[i] = 0.f;
        return;
    }
    int extra = 0;
    float pre = 1.f;
    if (mu < 1e-20f) {
        pre = 0x1p64f;
        mu *= 0x1p64f;
        extra = -64;
    }
    const int q = ((__float_as_int(mu) >> 23) & 0xFF) - 127;
    const float sc = __int_as_float((127 - q) << 23);  // two multiplications: pre * sc can overflow for denormal mu
#pragma unroll
    for (int i = 0; i < 4; ++i) u[i] = (u[i] * pre) * sc;
    dE = o.nf + q + extra;
    E15 = (e_in15 + dE) & 0x7FFF;
}

template <int KSTEPS, int NACC>
__global__ void __launch_bounds__(FLOW_THREADS, 1) hmm_flow_forward_tc(FlowArgs p) {
    constexpr int KS = KSTEPS * 8;
    constexpr int SP = KSTEPS * 64;
    constexpr int NG = KSTEPS;
    constexpr int NA = 8 * NG;
    constexpr int NCC = SP / 128;
    constexpr int CW = 128;
    constexpr int NT = 2;
    // NACC = independent TMEM accumulators, used round-robin (2*KS + 2*NACC*16 <= 512 columns)
    constexpr int STEPB = 512 + 16;        // bytes between consecutive k-steps of a B tile (+16: conflict-free STS)
    constexpr int TILEB = KSTEPS * STEPB;  // one B tile (hi or lo)
    constexpr uint32_t TX_BYTES = FLOW_CLUSTER * (NT * FLOW_CHAINS * 8 * 4 + FLOW_CHAINS * 4);
    static_assert(KSTEPS >= 2 && KSTEPS <= 16 && (KSTEPS & (KSTEPS - 1)) == 0, "KSTEPS must be a power of two in [2,16]");

    extern __shared__ __align__(128) unsigned char s_b[];  // [2 buf][2 (hi, lo)][TILEB]
    __shared__ __align__(16) float s_part[2][FLOW_CLUSTER][NT][8][FLOW_CHAINS];  // [parity][src][tile][column][chain]
    __shared__ __align__(16) int s_eref[2][FLOW_CLUSTER][FLOW_CHAINS];
    __shared__ __align__(16) int s_erefloc[2][FLOW_CHAINS];
    __shared__ __align__(8) unsigned long long s_bar[2];
    __shared__ __align__(8) unsigned long long s_mma_bar;
    __shared__ uint32_t s_tmem_base;
    __shared__ int s_abort;
    __shared__ long long s_ts[8], s_ts2[8];  // FB2_FLOW_PROF: event timestamps of the current step (one SM, one clock)
    __shared__ __align__(16) float s_obs[FLOW_OBS_DEPTH][64][4];  // cp.async ring of obs[b, t, 4 columns], private per phase-C thread

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rank = (int)cluster_ctarank();
    const int cid = blockIdx.x / FLOW_CLUSTER;
    const int set = cid / NCC, cc = cid % NCC;
    const int S = p.S;

    if (tid == 0) {
        mbar_init(smem_u32(&s_bar[0]), 1);
        mbar_init(smem_u32(&s_bar[1]), 1);
        mbar_init(smem_u32(&s_mma_bar), 1);
        s_abort = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_arm(smem_u32(&s_bar[0]), TX_BYTES);
        mbar_arm(smem_u32(&s_bar[1]), TX_BYTES);
    }
    if (warp == 0) {  // tensor memory: P hi [KS cols] | P lo [KS cols] | two D buffers of 32 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = 0;  // a 512-column allocation is the whole tensor memory
    if (s_tmem_base != 0) __trap();

    // ---- resident operand: lane m of TMEM holds column (cc*128 + m) of P for this CTA's K-slice (hi, then lo)
    if (warp >= 4) {
        const int m = tid - 128;
        const int col = cc * CW + m;
        const float ca = col < S ? p.cA[col] : 0.f;
        const uint32_t lane_base = tmem + ((uint32_t)((warp - 4) * 32) << 16);
#pragma unroll 1
        for (int k0 = 0; k0 < KS; k0 += 16) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int k = rank * KS + k0 + i;
                float v = 0.f;
                if (k < S && col < S) v = expf(p.A[(int64_t)k * S + col] - ca);
                hi[i] = cvt_tf32(v);
                lo[i] = cvt_tf32(v - __uint_as_float(hi[i]));
            }
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(
                    lane_base + (uint32_t)k0),
                "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]), "r"(hi[4]), "r"(hi[5]), "r"(hi[6]), "r"(hi[7]), "r"(hi[8]), "r"(hi[9]),
                "r"(hi[10]), "r"(hi[11]), "r"(hi[12]), "r"(hi[13]), "r"(hi[14]), "r"(hi[15])
                : "memory");
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(
                    lane_base + (uint32_t)(KS + k0)),
                "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]), "r"(lo[4]), "r"(lo[5]), "r"(lo[6]), "r"(lo[7]), "r"(lo[8]), "r"(lo[9]),
                "r"(lo[10]), "r"(lo[11]), "r"(lo[12]), "r"(lo[13]), "r"(lo[14]), "r"(lo[15])
                : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // ---- roles
    const int ga = tid / NG, pa = tid % NG;        // phase A (tid < NA): chain pair (ga, ga+8), k-step pa
    const int me = tid - 128;                      // epilogue (warps 4-7): TMEM lane = column me of the cluster's 128
    // phase C (warps 4 and 5): warp = tile, lane = chain + 16*half; a thread finishes 4 of the 8 columns of its group
    const bool isC = (warp == 4 || warp == 5);
    const int ntc = warp & 1, bc = lane & 15, jh = lane >> 4;
    const int tcc = tid - 128;  // 0..63 for the phase-C threads
    const int colC = cc * CW + ntc * 64 + rank * 8 + 4 * jh;
    float caC[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) caC[i] = (colC + i) < S ? p.cA[colC + i] : 0.f;
    const bool vec_obs = (S % 4) == 0 && (reinterpret_cast<uintptr_t>(p.obs) % 16) == 0;
    const uint32_t bar_addr0 = smem_u32(&s_bar[0]);
    const uint32_t mma_bar = smem_u32(&s_mma_bar);
    // epilogue thread me sends its 16 chain values to the final owner of its column
    uint32_t r_part_t = 0, r_eref_t = 0, r_bar_t = 0;
    if (me >= 0) {
        const uint32_t owner = (uint32_t)((me & 63) >> 3);
        r_part_t = mapa_u32(smem_u32(&s_part[0][rank][me >> 6][me & 7][0]), owner);
        r_eref_t = mapa_u32(smem_u32(&s_eref[0][rank][0]), owner);
        r_bar_t = mapa_u32(bar_addr0, owner);
    }

    cluster_sync_all();

    unsigned long long* const pkt_set = p.pkt + (size_t)set * 2 * FLOW_CHAINS * SP;
    long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned nstep = 0, tau = 0;
    bool aborted = false;

    for (int grp = set; grp < p.groups_total && !aborted; grp += p.nsets) {
        const int64_t b0 = (int64_t)grp * FLOW_CHAINS;
        int E_prev = 0, E_abs = 0;
        bool zero_last = true;
        float u_last[4] = {0.f, 0.f, 0.f, 0.f};
        const bool chain_ok = isC && (b0 + bc) < p.B;
        const float* obs_c = (chain_ok && p.obs && colC < S) ? p.obs + ((b0 + bc) * p.obs_T) * S + colC : nullptr;
        const uint32_t obs_ring = smem_u32(&s_obs[0][isC ? tcc : 0][0]);
        auto prefetch_obs = [&](int64_t tt) {
            if (obs_c && tt < p.T) {
                const uint32_t dst = obs_ring + (uint32_t)(tt % FLOW_OBS_DEPTH) * (64 * 16);
                const float* src = obs_c + tt * S;
                if (vec_obs) cp_async16(dst, src);
                else
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (colC + i < S) cp_async4(dst + 4 * i, src + i);
            }
            cp_async_commit();
        };
        if (isC) {
            float x[4], ones[4] = {1.f, 1.f, 1.f, 1.f};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                x[i] = -INFINITY;
                if (chain_ok && colC + i < S) x[i] = (p.init ? p.init[(b0 + bc) * p.init_sb + colC + i] : 0.f) * kLog2e;
            }
            int E15, dE;
            bool zero;
            flow_finish4(ones, flow_obs_term(x, false), 0, u_last, E15, zero, dE);  // |init| may be large (hand-over of alpha)
            E_abs = dE;
            E_prev = E15;
            zero_last = zero;
            const unsigned long long stamp = (unsigned long long)(((tau & 0xFFFFu) << 16) | (zero ? 0x8000u : 0u) | (unsigned)E15) << 32;
            unsigned long long* dst = pkt_set + ((size_t)(tau & 1) * FLOW_CHAINS + bc) * SP + colC;
            st_pkt2(dst, stamp | __float_as_uint(u_last[0]), stamp | __float_as_uint(u_last[1]));
            st_pkt2(dst + 2, stamp | __float_as_uint(u_last[2]), stamp | __float_as_uint(u_last[3]));
        }
#pragma unroll
        for (int q = 0; q < FLOW_OBS_DEPTH - 1; ++q) prefetch_obs(q);

        for (int64_t t = 0; t < p.T; ++t) {
            const int buf = nstep & 1, par = nstep & 1;
            const unsigned ph = (nstep >> 1) & 1;
            int timed_out = 0;
            if (p.prof && tid == 0) s_ts[0] = clock64();
            if (false) {
                if (nstep > 0) {
                    prof_acc[0] += s_ts[1] - s_ts[0];  // top -> packets valid
                    prof_acc[1] += s_ts[2] - s_ts[1];  // align, split, STS, proxy fence, bar.sync
                    prof_acc[2] += s_ts[3] - s_ts[2];  // MMA issue + commit
                    prof_acc[3] += s_ts[4] - s_ts[2];  // bar.sync -> accumulators ready (epilogue thread)
                    prof_acc[4] += s_ts[5] - s_ts[4];  // tcgen05.ld + sends
                    prof_acc[5] += s_ts[6] - s_ts[5];  // sends -> all partials here
                    prof_acc[6] += s_ts[7] - s_ts[6];  // reduce, normalise, publish
                }
            }
            unsigned char* const tile_hi = s_b + (size_t)(buf * 2 + 0) * TILEB;
            unsigned char* const tile_lo = s_b + (size_t)(buf * 2 + 1) * TILEB;

            // ================= phase A: fetch the K-slice, align exponents, write the B operand =================
            if (tid < NA) {
                constexpr unsigned amask = NA >= 32 ? 0xffffffffu : ((1u << (NA & 31)) - 1u);
                const unsigned long long* src0 = pkt_set + ((size_t)(tau & 1) * FLOW_CHAINS + ga) * SP + rank * KS + pa * 8;
                const unsigned long long* src1 = src0 + (size_t)8 * SP;
                unsigned long long q0[8], q1[8];
                const unsigned want = tau & 0xFFFFu;
                const long long start = clock64();
                for (;;) {
                    // light probe first: spin on ONE 16-byte load (the last packets this thread needs) so that the
                    // 8 CTAs x 128 threads polling the same lines do not flood the L2 slices the producers write to
                    // light probe first: spin on ONE 16-byte load (the last packets this thread needs); measured faster
                    // than polling all eight loads and than keeping several staggered probes in flight (mode 16), which
                    // load the L2 slices the producers are writing to
                    if (p.mode & 16) probe_wait(src1 + 6, want, start);
                    else
                        for (;;) {
                            ld_pkt2(src1 + 6, q1[6], q1[7]);
                            if ((unsigned)(q1[7] >> 48) == want || clock64() - start > FLOW_SPIN_LIMIT) break;
                        }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        ld_pkt2(src0 + 2 * i, q0[2 * i], q0[2 * i + 1]);
                        ld_pkt2(src1 + 2 * i, q1[2 * i], q1[2 * i + 1]);
                    }
                    bool ok = true;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        ok = ok && ((unsigned)(q0[i] >> 48) == want) && ((unsigned)(q1[i] >> 48) == want);
                    if (ok) break;
                    if (clock64() - start > FLOW_SPIN_LIMIT) {
                        timed_out = 1;
                        if (atomicCAS(p.abort_flag + 1, 0, 1) == 0) {
                            p.abort_flag[2] = (int)blockIdx.x;
                            p.abort_flag[3] = tid;
                            p.abort_flag[4] = (int)t;
                            p.abort_flag[5] = (int)want;
                            p.abort_flag[6] = (int)(q0[0] >> 32);
                            p.abort_flag[7] = (int)(q1[7] >> 32);
                        }
                        break;
                    }
                }
                if (p.prof && tid == 0) {
                    s_ts[1] = clock64();
                }
                const int e0 = (int)(q0[0] >> 32) & 0xFFFF, e1 = (int)(q1[0] >> 32) & 0xFFFF;
                const int base_lane = lane & ~(NG - 1);
                const int ref0 = __shfl_sync(amask, e0, base_lane) & 0x7FFF;
                const int ref1 = __shfl_sync(amask, e1, base_lane) & 0x7FFF;
                const int d0 = (e0 & 0x8000) ? -40000 : sext15(e0 - ref0);
                const int d1 = (e1 & 0x8000) ? -40000 : sext15(e1 - ref1);
                int m0 = d0, m1 = d1;
#pragma unroll
                for (int o = NG / 2; o > 0; o >>= 1) {
                    m0 = max(m0, __shfl_xor_sync(amask, m0, o));
                    m1 = max(m1, __shfl_xor_sync(amask, m1, o));
                }
                const float sc0 = (d0 == -40000) ? 0.f : pow2i(d0 - m0);
                const float sc1 = (d1 == -40000) ? 0.f : pow2i(d1 - m1);
                if (pa == 0) {
                    s_erefloc[buf][ga] = (m0 == -40000) ? (0x8000 | ref0) : ((ref0 + m0) & 0x7FFF);
                    s_erefloc[buf][ga + 8] = (m1 == -40000) ? (0x8000 | ref1) : ((ref1 + m1) & 0x7FFF);
                }
                // B operand, K-major: k-step pa at pa*STEPB; chunk j (4 k's) at +j*256; 8-row group at +128; row at +16*ga
                const uint32_t off = (uint32_t)pa * STEPB + (uint32_t)ga * 16;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    uint32_t h0[4], l0[4], h1[4], l1[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float v0 = __uint_as_float((unsigned)q0[4 * j + c]) * sc0;
                        const float v1 = __uint_as_float((unsigned)q1[4 * j + c]) * sc1;
                        h0[c] = cvt_tf32(v0);
                        l0[c] = cvt_tf32(v0 - __uint_as_float(h0[c]));
                        h1[c] = cvt_tf32(v1);
                        l1[c] = cvt_tf32(v1 - __uint_as_float(h1[c]));
                    }
                    *reinterpret_cast<uint4*>(tile_hi + off + j * 256) = make_uint4(h0[0], h0[1], h0[2], h0[3]);
                    *reinterpret_cast<uint4*>(tile_hi + off + j * 256 + 128) = make_uint4(h1[0], h1[1], h1[2], h1[3]);
                    *reinterpret_cast<uint4*>(tile_lo + off + j * 256) = make_uint4(l0[0], l0[1], l0[2], l0[3]);
                    *reinterpret_cast<uint4*>(tile_lo + off + j * 256 + 128) = make_uint4(l1[0], l1[1], l1[2], l1[3]);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core reads
            }
            if (__syncthreads_or(timed_out | *(volatile int*)&s_abort)) {
                aborted = true;
                break;
            }
            if (p.prof && tid == 0) {
                if (nstep > 0) {  // everybody is past the barrier: the previous step's timeline (s_ts2[]) is complete
                    prof_acc[2] += s_ts2[3] - s_ts2[2];  // MMA issue + commit
                    prof_acc[3] += s_ts2[4] - s_ts2[2];  // bar.sync -> accumulators ready (epilogue thread)
                    prof_acc[4] += s_ts2[5] - s_ts2[4];  // tcgen05.ld + sends
                    prof_acc[5] += s_ts2[6] - s_ts2[5];  // sends -> all partials here
                    prof_acc[6] += s_ts2[7] - s_ts2[6];  // reduce, normalise, publish
                    prof_acc[7] += s_ts[1] - s_ts2[7];   // own publish -> next packets valid (exchange + skew)
                }
                prof_acc[0] += s_ts[1] - s_ts[0];  // top -> packets valid
                s_ts[2] = clock64();
                prof_acc[1] += s_ts[2] - s_ts[1];  // align, split, STS, proxy fence, bar.sync
                s_ts2[2] = s_ts[2];
            }
            const uint32_t d_tmem = tmem + (uint32_t)(2 * KS + NACC * 16 * (int)(nstep & 1));

            // ================= phase B: one thread issues the MMAs =================
            if (warp == 0) {  // whole warp, warp-uniform operands; one elected lane issues
                tc_fence_after();
                if (nstep & 1) tc_issue_step<KSTEPS, NACC, 1>(smem_u32(s_b), mma_bar);
                else tc_issue_step<KSTEPS, NACC, 0>(smem_u32(s_b), mma_bar);
                if (p.prof && tid == 0) s_ts2[3] = clock64();
            }
            // ================= epilogue: D -> registers -> final owners (reduce-scatter over DSMEM) =================
            if (warp >= 4) {
                const long long start = clock64();
                while (!mbar_try_wait(mma_bar, nstep & 1)) {
                    if (clock64() - start > FLOW_SPIN_LIMIT) {
                        s_abort = 1;
                        if (atomicCAS(p.abort_flag + 1, 0, 3) == 0) {
                            p.abort_flag[2] = (int)blockIdx.x;
                            p.abort_flag[3] = tid;
                            p.abort_flag[4] = (int)t;
                        }
                        break;
                    }
                }
                tc_fence_after();
                if (p.prof && tid == 128) s_ts2[4] = clock64();
                uint32_t d[16];
                {
                    const uint32_t taddr = d_tmem + ((uint32_t)((warp - 4) * 32) << 16);
                    float acc[16];
#pragma unroll
                    for (int a = 0; a < NACC; ++a) {
                        uint32_t r[16];
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                            : "r"(taddr + (uint32_t)(a * 16))
                            : "memory");
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int q = 0; q < 16; ++q) acc[q] = a == 0 ? __uint_as_float(r[q]) : acc[q] + __uint_as_float(r[q]);
                    }
#pragma unroll
                    for (int q = 0; q < 16; ++q) d[q] = __float_as_uint(acc[q]);
                }
                tc_fence_before();
                const uint32_t r_bar = r_bar_t + (uint32_t)par * 8;
                const uint32_t dst = r_part_t + (uint32_t)par * (FLOW_CLUSTER * NT * 8 * FLOW_CHAINS * 4);
#pragma unroll
                for (int q = 0; q < 4; ++q) st_async_v4(dst + q * 16, d[4 * q], d[4 * q + 1], d[4 * q + 2], d[4 * q + 3], r_bar);
                if ((me & 71) == 0) {  // column 0 of tile 0 of each owner: also deliver the K-slice exponents
                    const uint32_t de = r_eref_t + (uint32_t)par * (FLOW_CLUSTER * FLOW_CHAINS * 4);
                    const uint4* src = reinterpret_cast<const uint4*>(&s_erefloc[buf][0]);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint4 e = src[q];
                        st_async_v4(de + q * 16, e.x, e.y, e.z, e.w, r_bar);
                    }
                }
            }
            if (p.prof && tid == 128) s_ts2[5] = clock64();

            // ================= phase C: final owners finish their columns and publish =================
            if (isC) {
                // observation-only work first: it overlaps the flight time of the partial sums
                cp_async_wait<FLOW_OBS_DEPTH - 2>();
                float x[4];
                {
                    const float4 ob = obs_c ? *reinterpret_cast<const float4*>(&s_obs[t % FLOW_OBS_DEPTH][tcc][0]) : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float obv[4] = {ob.x, ob.y, ob.z, ob.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) x[i] = (chain_ok && colC + i < S) ? (obv[i] + caC[i]) * kLog2e : -INFINITY;
                }
                const FlowObsTerm oterm = flow_obs_term(x, true);
                prefetch_obs(t + FLOW_OBS_DEPTH - 1);
                const long long start = clock64();
                while (!mbar_try_wait(bar_addr0 + (uint32_t)par * 8, ph)) {
                    if (clock64() - start > FLOW_SPIN_LIMIT) {
                        s_abort = 1;
                        if (atomicCAS(p.abort_flag + 1, 0, 2) == 0) {
                            p.abort_flag[2] = (int)blockIdx.x;
                            p.abort_flag[3] = tid;
                            p.abort_flag[4] = (int)t;
                        }
                        break;
                    }
                }
                if (p.prof && tid == 191) s_ts2[6] = clock64();
                const int e0 = s_eref[par][0][bc] & 0x7FFF;
                int dr[FLOW_CLUSTER], dmax = -40000;
#pragma unroll
                for (int r = 0; r < FLOW_CLUSTER; ++r) {
                    const int e = s_eref[par][r][bc];
                    dr[r] = (e & 0x8000) ? -40000 : sext15(e - e0);
                    dmax = max(dmax, dr[r]);
                }
                float sacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int r = 0; r < FLOW_CLUSTER; ++r) {
                    const float w = (dr[r] == -40000) ? 0.f : pow2i(dr[r] - dmax);
#pragma unroll
                    for (int i = 0; i < 4; ++i) sacc[i] = fmaf(s_part[par][r][ntc][4 * jh + i][bc], w, sacc[i]);
                }
                const int emax = (dmax == -40000) ? e0 : ((e0 + dmax) & 0x7FFF);
                // this phase is complete; arm the barrier for the step after next (no sender can be two steps ahead)
                if (tcc == 0) mbar_arm(bar_addr0 + (uint32_t)par * 8, TX_BYTES);
                int E15, dE;
                bool zero;
                flow_finish4(sacc, oterm, emax, u_last, E15, zero, dE);
                E_abs += sext15(E15 - E_prev);
                E_prev = E15;
                zero_last = zero;
                const unsigned tn = tau + 1;
                const unsigned long long stamp = (unsigned long long)(((tn & 0xFFFFu) << 16) | (zero ? 0x8000u : 0u) | (unsigned)E15) << 32;
                unsigned long long* dst = pkt_set + ((size_t)(tn & 1) * FLOW_CHAINS + bc) * SP + colC;
                if (p.mode & 1) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) atomicExch(dst + i, stamp | __float_as_uint(u_last[i]));
                } else {
                    st_pkt2(dst, stamp | __float_as_uint(u_last[0]), stamp | __float_as_uint(u_last[1]));
                    st_pkt2(dst + 2, stamp | __float_as_uint(u_last[2]), stamp | __float_as_uint(u_last[3]));
                }
                if (p.prof && tid == 191) s_ts2[7] = clock64();
            }
            ++nstep;
            ++tau;
        }
        cp_async_wait<0>();
        if (chain_ok) {
#pragma unroll
            for (int i = 0; i < 4; ++i) p.fin_u[(b0 + bc) * SP + colC + i] = u_last[i];
            if (jh == 0) p.fin_e[(b0 + bc) * (SP / 8) + colC / 8] = zero_last ? INT_MIN : E_abs;
        }
        tau += 2;
    }
    if (p.prof && tid == 0) {
        for (int q = 0; q < 8; ++q) p.prof[(size_t)blockIdx.x * 8 + q] = prof_acc[q];
        p.prof[(size_t)(FLOW_PROF_CTAS / 2 + blockIdx.x / 2) * 8 + (blockIdx.x & 1)] = (long long)nstep;
    }
    if (aborted && tid == 0) atomicExch(p.abort_flag, 1);
    tc_fence_before();
    cluster_sync_all();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// cA[j] = max(max_i A[i,j], -FLT_MAX) for j < S, 0 in the padding (clamped like numpy_log.py:31-32)
__global__ void flow_colmax(const float* A, int S, int Sp, float* cA, int transpose) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Sp) return;
    float m = -FLT_MAX;
    if (j < S)
        for (int i = 0; i < S; ++i) m = fmaxf(m, transpose ? A[(int64_t)j * S + i] : A[(int64_t)i * S + j]);
    cA[j] = j < S ? m : 0.f;
}

// logZ[b] = log sum_j u[b,j] 2^{E[b,j/8]}  (one warp per chain, double accumulation)
__global__ void flow_finalise(const float* fin_u, const int* fin_e, const int* abort_flag, int64_t B, int S, int Sp,
                              float* logZ, double* logZ_d, int* sticky) {
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (sticky && blockIdx.x == 0 && threadIdx.x == 0 && *abort_flag) atomicExch_system(sticky, abort_flag[1] ? abort_flag[1] : 9);
    if (b >= B) return;
    const int ngrp = Sp / 8;
    int emax = INT_MIN;
    for (int q = lane; q < ngrp; q += 32) emax = max(emax, fin_e[b * ngrp + q]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) emax = max(emax, __shfl_xor_sync(0xffffffffu, emax, o));
    double sum = 0.0;
    if (emax != INT_MIN)
        for (int j = lane; j < S; j += 32) {
            const int e = fin_e[b * ngrp + j / 8];
            if (e != INT_MIN) sum += ldexp((double)fin_u[b * Sp + j], max(e - emax, -2000));
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) {
        double z = (emax == INT_MIN) ? -INFINITY : (double)emax * 0.69314718055994530942 + log(sum);
        if (*abort_flag) z = __longlong_as_double(0x7ff8000000000000ll);
        logZ[b] = (float)z;
        if (logZ_d) logZ_d[b] = z;
    }
}

// gamma[b,t,j] = log alpha_t[j] + log beta_t[j] - logZ from the two recorded passes.  The backward pass records
// w_t = obs_t (.) beta_t (for t < T-1; w_{T-1} = obs_{T-1}), so log beta_t = log w_t - obs_t.  The large parts
// (block exponents, logZ) are combined in double; everything else is O(1).
__global__ void flow_gamma(const float* ua, const int* ea, const float* uw, const int* ew, const float* obs, const double* logZ,
                           int64_t B, int64_t T, int S, int Sp, float* gamma) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * T * (int64_t)S) return;
    const int j = (int)(idx % S);
    const int64_t bt = idx / S, t = bt % T, b = bt / T;
    const int G = Sp / 8;
    const float a = ua[bt * Sp + j];
    const int e1 = ea[bt * G + j / 8];
    float out = -INFINITY;
    if (e1 != INT_MIN && a > 0.f) {
        if (t == T - 1) {
            out = (float)((double)e1 * 0.69314718055994530942 - logZ[b]) + logf(a);
        } else {
            const float w = uw[bt * Sp + j];
            const int e2 = ew[bt * G + j / 8];
            if (e2 != INT_MIN && w > 0.f)
                out = (float)(((double)e1 + (double)e2) * 0.69314718055994530942 - logZ[b]) + (logf(a) + logf(w) - obs[idx]);
        }
    }
    gamma[idx] = out;
}

// ---- summed pairwise posteriors  xi_sum[i,j] = sum_t alpha_{t-1}[i] A[i,j] w_t[j] / Z  from the two recorded passes.
// Steps 1 .. T-2 are an outer-product accumulation over time (a [S x T] . [T x S] GEMM, 2 T S^2 flops) whose rows are
// rescaled on the fly: per step t the alpha side is aligned to its largest block exponent Ma_t, the w side to Mw_t, and
// the scalar 2^(Ma_t + Mw_t - log2 Z) (a probability-sized number) multiplies the alpha side.  The two end steps
// (alpha_{-1} = init, w_{T-1} = obs_{T-1}) are added by the finalising kernel.
constexpr int XI_TILE = 64, XI_KC = 16;

__global__ void xi_prep(const int* ea, const int* ew, const double* logZ, int64_t B, int64_t T, int G, int* Ma, int* Mw, float* st) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * T) return;
    const int64_t t = idx % T, b = idx / T;
    int ma = INT_MIN, mw = INT_MIN;
    float sc = 0.f;
    if (t >= 1 && t <= T - 2) {
        for (int g = 0; g < G; ++g) {
            ma = max(ma, ea[(idx - 1) * G + g]);
            mw = max(mw, ew[idx * G + g]);
        }
        if (ma != INT_MIN && mw != INT_MIN) {
            const double l2 = (double)ma + (double)mw - logZ[b] * 1.4426950408889634074;
            sc = (float)exp2(fmin(l2, 100.0));
        }
    }
    Ma[idx] = ma;
    Mw[idx] = mw;
    st[idx] = sc;
}

__global__ void __launch_bounds__(256) xi_gemm(const float* ua, const int* ea, const float* uw, const int* ew, const int* Ma, const int* Mw,
                                               const float* st, int64_t T, int S, int Sp, int ksplit, float* partial) {
    __shared__ __align__(16) float sa[XI_KC][XI_TILE + 4];
    __shared__ __align__(16) float sb[XI_KC][XI_TILE + 4];
    const int G = Sp / 8;
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int i0 = blockIdx.y * XI_TILE, j0 = blockIdx.x * XI_TILE;
    const int64_t b = blockIdx.z / ksplit;
    const int ks = blockIdx.z % ksplit;
    const int64_t span = (T - 2 + ksplit - 1) / ksplit;  // steps 1 .. T-2
    const int64_t t_lo = 1 + ks * span, t_hi = min(t_lo + span, T - 1);
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
    const int lt = tid / 16, lq = (tid % 16) * 4;  // staging: time row lt of the chunk, 4 consecutive states from lq
    for (int64_t t0 = t_lo; t0 < t_hi; t0 += XI_KC) {
        const int64_t t = t0 + lt;
        float4 av = make_float4(0.f, 0.f, 0.f, 0.f), bv = av;
        if (t < t_hi) {
            const int64_t ra = b * T + t - 1, rb = b * T + t;
            const float sc = st[rb];
            if (sc > 0.f) {
                if (i0 + lq < Sp) {
                    const int e = ea[ra * G + (i0 + lq) / 8];
                    const float f = (e == INT_MIN) ? 0.f : sc * pow2i(e - Ma[rb]);
                    const float4 v = *reinterpret_cast<const float4*>(ua + ra * Sp + i0 + lq);
                    av = make_float4(v.x * f, v.y * f, v.z * f, v.w * f);
                }
                if (j0 + lq < Sp) {
                    const int e = ew[rb * G + (j0 + lq) / 8];
                    const float f = (e == INT_MIN) ? 0.f : pow2i(e - Mw[rb]);
                    const float4 v = *reinterpret_cast<const float4*>(uw + rb * Sp + j0 + lq);
                    bv = make_float4(v.x * f, v.y * f, v.z * f, v.w * f);
                }
            }
        }
        *reinterpret_cast<float4*>(&sa[lt][lq]) = av;
        *reinterpret_cast<float4*>(&sb[lt][lq]) = bv;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < XI_KC; ++k) {
            const float4 xv = *reinterpret_cast<const float4*>(&sa[k][ty * 4]);
            const float4 yv = *reinterpret_cast<const float4*>(&sb[k][tx * 4]);
            const float xr[4] = {xv.x, xv.y, xv.z, xv.w}, yr[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(xr[r], yr[c], acc[r][c]);
        }
        __syncthreads();
    }
    float* out = partial + ((size_t)blockIdx.z * S) * S;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = i0 + ty * 4 + r;
        if (i >= S) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int j = j0 + tx * 4 + c;
            if (j < S) out[(size_t)i * S + j] = acc[r][c];
        }
    }
}

__global__ void xi_final(const float* init, int64_t init_sb, const float* A, const float* obs, const float* ua, const int* ea, const float* uw,
                         const int* ew, const double* logZ, const float* partial, int ksplit, int64_t B, int64_t T, int S, int Sp,
                         float* xi) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t per = (int64_t)S * S;
    if (idx >= B * per) return;
    const int64_t b = idx / per;
    const int i = (int)((idx % per) / S), j = (int)(idx % S);
    const int G = Sp / 8;
    const float a = A[(int64_t)i * S + j];
    const double z = logZ[b];
    const double LN2 = 0.69314718055994530942;
    float sum = 0.f;
    if (T >= 3) {
        float m = 0.f;
        for (int k = 0; k < ksplit; ++k) m += partial[((size_t)(b * ksplit + k) * S + i) * S + j];
        sum = __expf(a) * m;
    }
    const float in_i = init ? init[b * init_sb + i] : 0.f;
    if (T == 1) {
        sum += __expf(in_i + a + obs[(b * T) * S + j] - (float)z);
    } else {
        // t = 0: alpha_{-1} = init, w_0 from the backward history
        const int e0 = ew[(b * T) * G + j / 8];
        const float w0 = uw[(b * T) * Sp + j];
        if (e0 != INT_MIN && w0 > 0.f) sum += __expf(in_i + a + logf(w0) + (float)((double)e0 * LN2 - z));
        // t = T-1: alpha_{T-2} from the forward history, w_{T-1} = obs_{T-1}
        const int e1 = ea[(b * T + T - 2) * G + i / 8];
        const float a1 = ua[(b * T + T - 2) * Sp + i];
        if (e1 != INT_MIN && a1 > 0.f) sum += __expf(logf(a1) + a + obs[(b * T + T - 1) * S + j] + (float)((double)e1 * LN2 - z));
    }
    xi[idx] = sum;
}

// alpha[b,t,j] = log of the recorded message (mantissa, block exponent) as fp32
__global__ void flow_alpha_from_history(const float* u, const int* e, int64_t BT, int S, int Sp, float* alpha) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= BT * (int64_t)S) return;
    const int j = (int)(idx % S);
    const int64_t bt = idx / S;
    const float m = u[bt * Sp + j];
    const int ex = e[bt * (Sp / 8) + j / 8];
    alpha[idx] = (ex == INT_MIN || !(m > 0.f)) ? -INFINITY : (float)((double)ex * 0.69314718055994530942 + (double)logf(m));
}

int flow_padded_states(int S);

// rows_out[c, j] = log(final message of chain c)[j] - row_offset[c],  row_offset[c] = ln2 * (largest block exponent of chain c)
__global__ void flow_rows_out(const float* fin_u, const int* fin_e, const int* abort_flag, int64_t C, int S, int Sp, float* rows,
                              double* row_offset, int* sticky) {
    const int64_t c = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (sticky && blockIdx.x == 0 && threadIdx.x == 0 && *abort_flag) atomicExch_system(sticky, abort_flag[1] ? abort_flag[1] : 9);
    if (c >= C) return;
    const int ngrp = Sp / 8;
    int emax = INT_MIN;
    for (int q = lane; q < ngrp; q += 32) emax = max(emax, fin_e[c * ngrp + q]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) emax = max(emax, __shfl_xor_sync(0xffffffffu, emax, o));
    const bool bad = *abort_flag != 0;
    for (int j = lane; j < S; j += 32) {
        const int e = fin_e[c * ngrp + j / 8];
        const float u = fin_u[c * Sp + j];
        float v = -INFINITY;
        if (e != INT_MIN && u > 0.f) v = logf(u) + (float)(e - emax) * 0.69314718055994530942f;
        rows[c * S + j] = bad ? __int_as_float(0x7fc00000) : v;
    }
    if (lane == 0) row_offset[c] = emax == INT_MIN ? 0.0 : (double)emax * 0.69314718055994530942;
}

// ------------------------------------------------------------------------------------------ host side
struct FlowShape {
    int k16, nt;
};
static FlowShape flow_shape(int S) {
    if (S <= 64 || S > 1024) return {0, 0};
    if (S <= 128) return {1, 1};
    if (S <= 256) return {2, 1};
    if (S <= 512) return {4, 1};
    return {8, 2};
}

int flow_padded_states(int S) { return flow_shape(S).k16 * 128; }

bool hmm_flow_supported(int semiring, int64_t a_sb, int64_t a_st, int S, bool wants_alpha) {
    return semiring == FB2_SEMIRING_LOG && a_sb == 0 && a_st == 0 && !wants_alpha && flow_shape(S).k16 != 0;
}

constexpr int FLOW_MAX_SETS = 8;

size_t hmm_flow_workspace_bytes(int64_t B, int S) {
    const FlowShape fs = flow_shape(S);
    if (!fs.k16) return 0;
    const size_t Sp = (size_t)fs.k16 * 128;
    const size_t groups = (size_t)ceil_div(B, FLOW_CHAINS);
    return align_up(Sp * 4) + align_up((size_t)FLOW_MAX_SETS * 2 * FLOW_CHAINS * Sp * 8) + align_up(groups * FLOW_CHAINS * Sp * 4) +
           align_up(groups * FLOW_CHAINS * (Sp / 8) * 4) + align_up(32) + align_up((size_t)FLOW_PROF_CTAS * 64) + 256;
}

template <int K16, int NT, int CL = 8, bool PK4 = false>
static int flow_launch(FlowArgs& a, cudaStream_t stream) {
    auto kernel = hmm_flow_forward<K16, NT, CL, PK4>;
    constexpr int NCC = K16 * 128 / (CL * 8 * NT);
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(FLOW_THREADS);
    cfg.dynamicSmemBytes = (size_t)2 * (K16 * 128 / CL / 8) * (1024 + 16);
    FB2_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes));
    cfg.stream = stream;
    cudaLaunchAttribute attrs[1];
    attrs[0].id = cudaLaunchAttributeClusterDimension;
    attrs[0].val.clusterDim.x = CL;
    attrs[0].val.clusterDim.y = 1;
    attrs[0].val.clusterDim.z = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = 1;
    // how many clusters of 8 can be co-resident?  (the CTAs of one set wait on each other)
    cfg.gridDim = dim3(CL * NCC);
    int max_clusters = 0;
    FB2_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, kernel, &cfg));
    int nsets = max_clusters / NCC;
    if (nsets > FLOW_MAX_SETS) nsets = FLOW_MAX_SETS;
    if (nsets > a.groups_total) nsets = a.groups_total;
    if (nsets < 1) {
        set_error("hmm flow: only %d clusters of %d CTAs can be co-resident, %d needed", max_clusters, CL, NCC);
        return FB2_ERR_UNSUPPORTED;
    }
    a.nsets = nsets;
    cfg.gridDim = dim3(CL * NCC * nsets);
    // Plain cluster launch.  The occupancy query above proves that every cluster of the grid can be resident at
    // once, which is what the inter-CTA waits need.  cudaLaunchAttributeCooperative is deliberately NOT added:
    // measured on B200 / driver 580, a cooperative + cluster launch is not co-scheduled under ncu / CUPTI kernel
    // serialisation (the first packet wait of every CTA timed out), while the plain cluster launch runs there.
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, a);
    if (getenv("FB2_DEBUG")) fprintf(stderr, "[fb2 flow] launch grid %d max_clusters %d -> %s\n", cfg.gridDim.x, max_clusters, cudaGetErrorString(e));
    if (e != cudaSuccess) {
        set_error("hmm flow launch failed: %s", cudaGetErrorString(e));
        return FB2_ERR_CUDA;
    }
    count_launch();
    return FB2_OK;
}

template <int KW>
static int direct_launch(FlowArgs& a, cudaStream_t stream) {
    auto kernel = hmm_direct_forward<KW>;
    constexpr int NC = KW * 8;
    int per_sm = 0, sms = 0, dev = 0;
    FB2_CUDA(cudaGetDevice(&dev));
    FB2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    FB2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, FLOW_THREADS, 0));
    int nsets = per_sm * sms / NC;
    if (nsets > FLOW_MAX_SETS) nsets = FLOW_MAX_SETS;
    if (nsets > a.groups_total) nsets = a.groups_total;
    if (nsets < 1) {
        set_error("hmm direct: %d CTAs cannot be co-resident (%d per SM x %d SMs)", NC, per_sm, sms);
        return FB2_ERR_UNSUPPORTED;
    }
    a.nsets = nsets;
    void* args[] = {&a};
    // cooperative launch: all CTAs of the grid are resident together (they wait on each other's packets)
    cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kernel), dim3((unsigned)(nsets * NC)), dim3(FLOW_THREADS), args, 0, stream);
    if (getenv("FB2_DEBUG")) fprintf(stderr, "[fb2 direct] launch grid %d (nsets %d) -> %s\n", nsets * NC, nsets, cudaGetErrorString(e));
    if (e != cudaSuccess) {
        set_error("hmm direct launch failed: %s", cudaGetErrorString(e));
        return FB2_ERR_CUDA;
    }
    count_launch();
    return FB2_OK;
}

template <int KSTEPS, int NACC>
static int flow_launch_tc(FlowArgs& a, cudaStream_t stream) {
    auto kernel = hmm_flow_forward_tc<KSTEPS, NACC>;
    constexpr int NCC = KSTEPS / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(FLOW_THREADS);
    cfg.dynamicSmemBytes = (size_t)4 * KSTEPS * (512 + 16);
    FB2_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes));
    cfg.stream = stream;
    cudaLaunchAttribute attrs[1];
    attrs[0].id = cudaLaunchAttributeClusterDimension;
    attrs[0].val.clusterDim.x = FLOW_CLUSTER;
    attrs[0].val.clusterDim.y = 1;
    attrs[0].val.clusterDim.z = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = 1;
    cfg.gridDim = dim3(FLOW_CLUSTER * NCC);
    int max_clusters = 0;
    FB2_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, kernel, &cfg));
    int nsets = max_clusters / NCC;
    if (nsets > FLOW_MAX_SETS) nsets = FLOW_MAX_SETS;
    if (nsets > a.groups_total) nsets = a.groups_total;
    if (nsets < 1) {
        set_error("hmm flow (tcgen05): only %d clusters of %d CTAs can be co-resident, %d needed", max_clusters, FLOW_CLUSTER, NCC);
        return FB2_ERR_UNSUPPORTED;
    }
    a.nsets = nsets;
    cfg.gridDim = dim3(FLOW_CLUSTER * NCC * nsets);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, a);
    if (getenv("FB2_DEBUG")) fprintf(stderr, "[fb2 flow tc] launch grid %d max_clusters %d -> %s\n", cfg.gridDim.x, max_clusters, cudaGetErrorString(e));
    if (e != cudaSuccess) {
        set_error("hmm flow (tcgen05) launch failed: %s", cudaGetErrorString(e));
        return FB2_ERR_CUDA;
    }
    count_launch();
    return FB2_OK;
}

int hmm_flow_pass_device(const float* init, int64_t init_sb, const float* A, const float* obs, int64_t obs_T, int64_t B, int64_t T, int S,
                         float* logZ, double* logZ_d, float* hist_u, int* hist_e, int64_t hist_T, int transpose, int reverse,
                         void* workspace, size_t workspace_bytes, cudaStream_t stream, const unsigned* ready_flags = nullptr,
                         int64_t rows_per_chunk = 0, int cpo = 1, float* rows_out = nullptr, double* row_offset = nullptr);

int hmm_flow_forward_device(const float* init, int64_t init_sb, const float* A, const float* obs, int64_t B, int64_t T, int S,
                            float* logZ, double* logZ_d, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    return hmm_flow_pass_device(init, init_sb, A, obs, T, B, T, S, logZ, logZ_d, nullptr, nullptr, 0, 0, 0, workspace, workspace_bytes,
                                stream);
}

// General pass: T steps over obs rows [0, T) of an array with obs_T rows per chain, optionally with A transposed and
// time reversed (the backward recursion is the forward recursion on (A^T, reversed obs), see fb2_hmm_marginals_f32),
// optionally recording every message (mantissas + block exponents).
// set by hmm_flow_marginals_device around its two passes: take the two-tiles-per-warp variant (half the clusters) even for a
// single chain group, so that the forward and the backward pass fit on the GPU side by side
static thread_local int tl_force_wide = 0;
// set by hmm_flow_marginals_device while it runs its two passes side by side: the pair is ONE unit of the per-device serialisation
static thread_local int tl_in_pair = 0;

// Per-device state of the dataflow kernels.
//  * Every pass needs all of its CTAs co-resident (they spin on each other's packets).  cudaOccupancyMaxActiveClusters proves that
//    for an otherwise empty GPU only, so two passes launched from different host threads / streams must not overlap: each pass
//    waits (on the device, cudaStreamWaitEvent -- no host synchronisation) for the completion event of the previous pass on this
//    device and records its own.  Foreign kernels can still delay residency; they end, and the spin limit (~1 s) is far beyond that.
//  * A pass that does give up (lost co-residency, a host-to-device chunk that never landed) poisons its outputs with NaN and raises
//    a sticky word in mapped host memory: fb2_async_status() reports it, the next dataflow call on the device refuses to run until
//    it has been read, and the host-buffer entry points (which synchronise anyway) retry on the cooperative kernel.
struct FlowDevice {
    std::mutex mu;
    cudaEvent_t last_done = nullptr;
    bool has_last = false;
    int* sticky_host = nullptr;
    int* sticky_dev = nullptr;
};
static int flow_device(FlowDevice** out) {
    static FlowDevice table[64];
    int dev = 0;
    FB2_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return FB2_ERR_UNSUPPORTED;
    FlowDevice& d = table[dev];
    std::lock_guard<std::mutex> lock(d.mu);
    if (!d.last_done) {
        FB2_CUDA(cudaEventCreateWithFlags(&d.last_done, cudaEventDisableTiming));
        FB2_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&d.sticky_host), sizeof(int), cudaHostAllocMapped));
        *d.sticky_host = 0;
        FB2_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&d.sticky_dev), d.sticky_host, 0));
    }
    *out = &d;
    return FB2_OK;
}
static void flow_serialise_begin(FlowDevice* d, cudaStream_t stream) {
    std::lock_guard<std::mutex> lock(d->mu);
    if (d->has_last) cudaStreamWaitEvent(stream, d->last_done, 0);
}
static void flow_serialise_end(FlowDevice* d, cudaStream_t stream) {
    std::lock_guard<std::mutex> lock(d->mu);
    cudaEventRecord(d->last_done, stream);
    d->has_last = true;
}
// 0 = no dataflow pass on the current device has aborted since the last clear; otherwise the kind of the first wait that gave up
// (1 packets, 2/3 partials, 4 input chunk, 9 unknown).
int hmm_flow_async_status(int clear) {
    FlowDevice* d = nullptr;
    if (flow_device(&d) != FB2_OK) return 0;
    const int v = *static_cast<volatile int*>(d->sticky_host);
    if (clear) *static_cast<volatile int*>(d->sticky_host) = 0;
    return v;
}

int hmm_flow_pass_device(const float* init, int64_t init_sb, const float* A, const float* obs, int64_t obs_T, int64_t B, int64_t T, int S,
                         float* logZ, double* logZ_d, float* hist_u, int* hist_e, int64_t hist_T, int transpose, int reverse,
                         void* workspace, size_t workspace_bytes, cudaStream_t stream, const unsigned* ready_flags,
                         int64_t rows_per_chunk, int cpo, float* rows_out, double* row_offset) {
    const FlowShape fs = flow_shape(S);
    if (!fs.k16) {
        set_error("hmm flow: S=%d outside the supported range", S);
        return FB2_ERR_UNSUPPORTED;
    }
    const int Sp = fs.k16 * 128;
    const int64_t groups = ceil_div(B, FLOW_CHAINS);
    FlowDevice* dv = nullptr;
    {
        const int dst = flow_device(&dv);
        if (dst != FB2_OK) return dst;
    }
    if (const int kind = *static_cast<volatile int*>(dv->sticky_host)) {
        set_error("an earlier dataflow pass on this device aborted (kind %d: its CTAs lost co-residency or an input chunk never arrived); "
                  "its outputs are NaN.  Read and clear the condition with fb2_async_status(1) before launching another pass", kind);
        return FB2_ERR_ABORTED;
    }
    Arena arena(workspace, workspace_bytes);
    float* cA = arena.take<float>((size_t)Sp);
    unsigned long long* pkt = arena.take<unsigned long long>((size_t)FLOW_MAX_SETS * 2 * FLOW_CHAINS * Sp);
    float* fin_u = arena.take<float>((size_t)groups * FLOW_CHAINS * Sp);
    int* fin_e = arena.take<int>((size_t)groups * FLOW_CHAINS * (Sp / 8));
    int* abort_flag = arena.take<int>(8);  // [0] abort, [1..7] post-mortem of the first wait that gave up
    long long* prof = arena.take<long long>((size_t)FLOW_PROF_CTAS * 8);
    if (!arena.ok()) {
        set_error("hmm flow: workspace too small (%zu bytes)", workspace_bytes);
        return FB2_ERR_WORKSPACE;
    }
    // stale packets must never carry a valid stamp: 0xFF.. = stamp 0xFFFF, overwritten after two steps
    if (!tl_in_pair) flow_serialise_begin(dv, stream);
    FB2_CUDA(cudaMemsetAsync(pkt, 0xFF, (size_t)FLOW_MAX_SETS * 2 * FLOW_CHAINS * Sp * 8, stream));
    FB2_CUDA(cudaMemsetAsync(abort_flag, 0, 32, stream));
    if (const char* fa = getenv("FB2_FLOW_FORCE_ABORT")) {  // test hook: behave as if a wait had given up (outputs NaN, sticky status raised)
        if (fa[0] == '1') FB2_CUDA(cudaMemsetAsync(abort_flag, 1, 1, stream));
    }
    flow_colmax<<<(unsigned)ceil_div(Sp, 128), 128, 0, stream>>>(A, S, Sp, cA, transpose);
    FB2_LAUNCH_CHECK();
    FlowArgs a{};
    a.init = init; a.init_sb = init_sb; a.A = A; a.obs = obs; a.cA = cA; a.B = B; a.T = T; a.S = S;
    a.pkt = pkt; a.fin_u = fin_u; a.fin_e = fin_e; a.abort_flag = abort_flag; a.groups_total = (int)groups;
    a.hist_u = hist_u; a.hist_e = hist_e; a.hist_T = hist_T; a.obs_T = obs_T; a.transpose = transpose; a.reverse = reverse;
    a.ready_flags = ready_flags; a.rows_per_chunk = rows_per_chunk > 0 ? rows_per_chunk : 1;
    a.cpo = cpo > 1 ? cpo : 1;
    const char* mode_env = getenv("FB2_FLOW_MODE");
    a.mode = mode_env ? atoi(mode_env) : 1;  // default: publish packets with atom.exch (store->load hand-over measured ~10 % faster)
    const char* prof_env = getenv("FB2_FLOW_PROF");
    a.prof = (prof_env && prof_env[0] == '1') ? prof : nullptr;
    if (a.prof) FB2_CUDA(cudaMemsetAsync(prof, 0, (size_t)FLOW_PROF_CTAS * 64, stream));
    int st;
    // Two implementations of the per-step contraction, same dataflow around them: mma.sync 3xTF32 with P in registers
    // (default: measured 3.4 us/step at S=1024) and tcgen05 kind::tf32 with P in tensor memory (FB2_FLOW_TC=1: 3.7 us/step;
    // its shorter MMA phase is hidden behind the DSMEM reduce-scatter, which moves the same 8 KB per CTA-step either way).
    const char* tc_env = getenv("FB2_FLOW_TC");
    const bool legacy = !(tc_env && tc_env[0] == '1') || hist_u != nullptr || transpose || reverse || ready_flags != nullptr || cpo > 1;
    // single-hop variant (hmm_direct_forward): measured faster than the two-hop kernel for S <= 512 (1.32 / 1.86 / 2.35 vs 2.08 / 2.42 /
    // 2.66 us per step at S = 128 / 256 / 512) and slower at S = 1024 (3.68 vs 3.37: 8 MB of L2 reads per step).  FB2_FLOW_DIRECT=0|1 forces.
    // Round 2: with 4-byte words, bf16 side products and the shared alignment work the two-hop kernel overtakes it for a single chain group
    // at 256 < S <= 512 when a pass runs alone (2.05 vs 2.28 us per step, forward); side by side (forward-backward, two passes of 4 wide
    // clusters each) the single-hop kernel stays ahead (2.35 vs 2.29 / 2.46 us at B = 1 / 16).
    const char* direct_env = getenv("FB2_FLOW_DIRECT");
    const char* pk4_off = getenv("FB2_FLOW_PK4");
    const bool two_hop_512 = fs.k16 == 4 && groups == 1 && !tl_in_pair && !(pk4_off && pk4_off[0] == '0');
    const bool direct = direct_env ? direct_env[0] == '1' : (fs.k16 <= 4 && !two_hop_512);
    if (direct) {
        if (fs.k16 == 1) st = direct_launch<2>(a, stream);
        else if (fs.k16 == 2) st = direct_launch<4>(a, stream);
        else if (fs.k16 == 4) st = direct_launch<8>(a, stream);
        else st = direct_launch<16>(a, stream);
    } else if (!legacy) {
        if (fs.k16 == 1) st = flow_launch_tc<2, 1>(a, stream);
        else if (fs.k16 == 2) st = flow_launch_tc<4, 1>(a, stream);
        else if (fs.k16 == 4) st = flow_launch_tc<8, 1>(a, stream);
        else if (a.mode & 16) st = flow_launch_tc<16, 4>(a, stream);
        else st = flow_launch_tc<16, 1>(a, stream);
    } else {
        // Many chain groups (B > 16, chunk summaries): two n-tiles per warp halve the clusters a set needs, so two to
        // three times as many groups run side by side (at most 15 clusters of 8 are co-resident); a single group is
        // faster with one tile per warp (less tensor work on its critical path).
        const bool wide = groups > 1 || tl_force_wide;
        const char* pk4_env = getenv("FB2_FLOW_PK4");  // 0: 8-byte packets also for a single chain group
        const bool pk4 = groups == 1 && !(pk4_env && pk4_env[0] == '0');
        if (fs.k16 == 1) st = wide ? flow_launch<1, 2>(a, stream) : flow_launch<1, 1>(a, stream);
        else if (fs.k16 == 2) st = wide ? flow_launch<2, 2>(a, stream) : flow_launch<2, 1>(a, stream);
        else if (fs.k16 == 4) {
            if (pk4) st = wide ? flow_launch<4, 2, 8, true>(a, stream) : flow_launch<4, 1, 8, true>(a, stream);
            else st = wide ? flow_launch<4, 2>(a, stream) : flow_launch<4, 1>(a, stream);
        } else {
            const char* cl4 = getenv("FB2_FLOW_CL4");  // experiment: clusters of 4 with half-K warps
            if (cl4 && cl4[0] == '1') st = flow_launch<8, 1, 4>(a, stream);
            else if (pk4) st = flow_launch<8, 2, 8, true>(a, stream);
            else st = flow_launch<8, 2>(a, stream);
        }
    }
    if (st != FB2_OK) {
        if (!tl_in_pair) flow_serialise_end(dv, stream);
        return st;
    }
    if (rows_out) {  // chunk summary: the final messages themselves, as rows relative to a float64 offset
        flow_rows_out<<<(unsigned)ceil_div(B, 4), 128, 0, stream>>>(fin_u, fin_e, abort_flag, B, S, Sp, rows_out, row_offset, dv->sticky_dev);
        FB2_LAUNCH_CHECK();
    } else {
        flow_finalise<<<(unsigned)ceil_div(B, 4), 128, 0, stream>>>(fin_u, fin_e, abort_flag, B, S, Sp, logZ, logZ_d, dv->sticky_dev);
        FB2_LAUNCH_CHECK();
    }
    if (!tl_in_pair) flow_serialise_end(dv, stream);
    if (getenv("FB2_DEBUG")) {
        int h[8];
        FB2_CUDA(cudaStreamSynchronize(stream));
        FB2_CUDA(cudaMemcpy(h, abort_flag, sizeof(h), cudaMemcpyDeviceToHost));
        if (h[0] || h[1])
            fprintf(stderr, "[fb2 flow] ABORT=%d first timeout: kind %d (1 packets, 2 partials) cta %d tid %d t %d want %#x saw %#x %#x\n",
                    h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
    }
    if (a.prof && !legacy) {  // debugging aid only: per-step timeline of a few CTAs (tcgen05 variant)
        static long long h[FLOW_PROF_CTAS * 8];
        FB2_CUDA(cudaStreamSynchronize(stream));
        FB2_CUDA(cudaMemcpy(h, prof, sizeof(h), cudaMemcpyDeviceToHost));
        for (int c = 0; c < FLOW_PROF_CTAS / 2; c += 9) {
            const double n = (double)h[(FLOW_PROF_CTAS / 2 + c / 2) * 8 + (c & 1)] - 1.0;
            if (n > 0)
                fprintf(stderr, "[fb2 flow tc prof] cta %3d | top->pkts %.0f  A-proc %.0f | issue %.0f  sync->acc-ready %.0f  ld+send %.0f | "
                        "send->partials %.0f  C-proc %.0f | publish->next pkts %.0f cycles/step\n",
                        c, h[c * 8 + 0] / n, h[c * 8 + 1] / n, h[c * 8 + 2] / n, h[c * 8 + 3] / n, h[c * 8 + 4] / n, h[c * 8 + 5] / n,
                        h[c * 8 + 6] / n, h[c * 8 + 7] / n);
        }
    } else if (a.prof) {  // debugging aid only: synchronises and prints the per-phase cycle averages of a few CTAs
        static long long h[FLOW_PROF_CTAS * 8];
        FB2_CUDA(cudaStreamSynchronize(stream));
        FB2_CUDA(cudaMemcpy(h, prof, sizeof(h), cudaMemcpyDeviceToHost));
        for (int c = 0; c < FLOW_PROF_CTAS / 2; c += 9) {
            const double n = h[c * 8 + 5] > 0 ? (double)h[c * 8 + 5] : 1.0;
            if (h[c * 8 + 5] > 0)
                fprintf(stderr, "[fb2 flow prof] cta %3d steps %lld | pkt-wait %.0f  A-proc+sync %.0f  B(mma+send) %.0f | C-wait %.0f  C-proc %.0f (reduce|B-last-warp %.0f obs|bar1 %.0f normalise|step %.0f) cycles/step\n",
                        c, h[c * 8 + 5], h[c * 8 + 0] / n, h[c * 8 + 1] / n, h[c * 8 + 2] / n, h[c * 8 + 3] / n, h[c * 8 + 4] / n,
                        h[c * 8 + 6] / n, h[c * 8 + 7] / n, h[(FLOW_PROF_CTAS / 2 + c / 2) * 8 + (c & 1)] / n);
        }
    }
    return FB2_OK;
}


static int xi_ksplit(int64_t T) {
    int64_t k = (T + 2047) / 2048;
    return (int)(k < 1 ? 1 : (k > 64 ? 64 : k));
}

size_t hmm_flow_marginals_workspace_bytes(int64_t B, int64_t T, int S) {
    const size_t Sp = (size_t)flow_padded_states(S);
    if (!Sp) return 0;
    return 2 * (align_up((size_t)B * T * Sp * 4) + align_up((size_t)B * T * (Sp / 8) * 4)) + 2 * align_up((size_t)B * 8) +
           align_up((size_t)B * 4) + 3 * align_up((size_t)B * T * 4) + align_up((size_t)B * xi_ksplit(T) * S * S * 4) +
           2 * hmm_flow_workspace_bytes(B, S) + 512;
}

// A side stream per device for the backward pass of small-batch forward-backward (see below).
struct SideStream {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};
static int side_stream_for_current_device(SideStream** out) {
    // per host thread: the fork / join events belong to one call at a time (two threads sharing them could wait on each other's fork)
    static thread_local SideStream table[64];
    int dev = 0;
    FB2_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return FB2_ERR_UNSUPPORTED;
    SideStream& s = table[dev];
    if (!s.stream) {
        FB2_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        FB2_CUDA(cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming));
        FB2_CUDA(cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming));
    }
    *out = &s;
    return FB2_OK;
}

// Forward-backward posteriors on the dataflow kernel: a forward pass and a backward pass (= the same kernel on A^T and
// time-reversed observations, started from obs_{T-1}) record their messages; flow_gamma combines them.
int hmm_flow_marginals_device(const float* init, int64_t init_sb, const float* A, const float* obs, int64_t B, int64_t T, int S,
                              float* logZ, float* gamma, float* xi_sum, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    const int Sp = flow_padded_states(S);
    if (!Sp) return FB2_ERR_UNSUPPORTED;
    Arena arena(workspace, workspace_bytes);
    float* ua = arena.take<float>((size_t)B * T * Sp);
    int* ea = arena.take<int>((size_t)B * T * (Sp / 8));
    float* uw = arena.take<float>((size_t)B * T * Sp);
    int* ew = arena.take<int>((size_t)B * T * (Sp / 8));
    double* z_d = arena.take<double>((size_t)B);
    double* zb_d = arena.take<double>((size_t)B);
    float* zb = arena.take<float>((size_t)B);
    const int ksplit = xi_ksplit(T);
    int* Ma = arena.take<int>((size_t)B * T);
    int* Mw = arena.take<int>((size_t)B * T);
    float* stt = arena.take<float>((size_t)B * T);
    float* partial = arena.take<float>((size_t)B * ksplit * S * S);
    if (!arena.ok()) {
        set_error("hmm flow marginals: workspace too small (%zu bytes)", workspace_bytes);
        return FB2_ERR_WORKSPACE;
    }
    char* rest = static_cast<char*>(workspace) + arena.used;
    const size_t rest_bytes = workspace_bytes - arena.used;
    // One chain group (B <= 16) occupies 2 K16 / NT clusters of the 15 a B200 can hold, and the two passes are independent:
    // run the backward pass on a side stream next to the forward pass.  Each pass needs all its clusters resident, so this
    // is only done when both fit together with room to spare (S <= 512: 4 + 4 clusters, using the two-tiles-per-warp variant at
    // K16 = 4); S = 1024 needs 8 clusters per pass and stays sequential.  FB2_FLOW_DUAL=0 switches it off.
    const FlowShape fs = flow_shape(S);
    const char* dual_env = getenv("FB2_FLOW_DUAL");
    const size_t per_pass = hmm_flow_workspace_bytes(B, S);
    const bool dual = B <= FLOW_CHAINS && fs.k16 <= 4 && T >= 64 && rest_bytes >= 2 * per_pass && !(dual_env && dual_env[0] == '0');
    int st;
    if (dual) {
        SideStream* side = nullptr;
        st = side_stream_for_current_device(&side);
        if (st != FB2_OK) return st;
        FlowDevice* dv = nullptr;
        st = flow_device(&dv);
        if (st != FB2_OK) return st;
        flow_serialise_begin(dv, stream);
        FB2_CUDA(cudaEventRecord(side->fork, stream));
        FB2_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
        tl_force_wide = fs.k16 == 4;
        tl_in_pair = 1;
        st = hmm_flow_pass_device(obs + (T - 1) * (int64_t)S, T * (int64_t)S, A, obs, T, B, T - 1, S, zb, zb_d, uw, ew, T, 1, 1, rest + per_pass,
                                  rest_bytes - per_pass, side->stream);
        if (st == FB2_OK) st = hmm_flow_pass_device(init, init_sb, A, obs, T, B, T, S, logZ, z_d, ua, ea, T, 0, 0, rest, per_pass, stream);
        tl_force_wide = 0;
        tl_in_pair = 0;
        // join even on failure so that the caller's stream never runs ahead of the side stream
        cudaEventRecord(side->join, side->stream);
        cudaStreamWaitEvent(stream, side->join, 0);
        flow_serialise_end(dv, stream);
        if (st != FB2_OK) return st;
    } else {
        st = hmm_flow_pass_device(init, init_sb, A, obs, T, B, T, S, logZ, z_d, ua, ea, T, 0, 0, rest, rest_bytes, stream);
        if (st != FB2_OK) return st;
        // backward: w_{T-1} = obs_{T-1}; w_{t-1}[i] = obs_{t-1}[i] * sum_j A[i,j] w_t[j]  ->  T-1 steps over obs rows T-2 .. 0
        st = hmm_flow_pass_device(obs + (T - 1) * (int64_t)S, T * (int64_t)S, A, obs, T, B, T - 1, S, zb, zb_d, uw, ew, T, 1, 1, rest, rest_bytes,
                                  stream);
        if (st != FB2_OK) return st;
    }
    const int64_t n = B * T * (int64_t)S;
    flow_gamma<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(ua, ea, uw, ew, obs, z_d, B, T, S, Sp, gamma);
    FB2_LAUNCH_CHECK();
    if (xi_sum) {
        if (T >= 3) {
            xi_prep<<<(unsigned)ceil_div(B * T, 256), 256, 0, stream>>>(ea, ew, z_d, B, T, Sp / 8, Ma, Mw, stt);
            FB2_LAUNCH_CHECK();
            dim3 grid((unsigned)ceil_div(S, XI_TILE), (unsigned)ceil_div(S, XI_TILE), (unsigned)(B * ksplit));
            xi_gemm<<<grid, 256, 0, stream>>>(ua, ea, uw, ew, Ma, Mw, stt, T, S, Sp, ksplit, partial);
            FB2_LAUNCH_CHECK();
        }
        const int64_t m = B * (int64_t)S * S;
        xi_final<<<(unsigned)ceil_div(m, 256), 256, 0, stream>>>(init, init_sb, A, obs, ua, ea, uw, ew, z_d, partial, ksplit, B, T, S, Sp, xi_sum);
        FB2_LAUNCH_CHECK();
    }
    return FB2_OK;
}


size_t hmm_flow_alpha_workspace_bytes(int64_t B, int64_t T, int S) {
    const size_t Sp = (size_t)flow_padded_states(S);
    if (!Sp) return 0;
    return align_up((size_t)B * T * Sp * 4) + align_up((size_t)B * T * (Sp / 8) * 4) + hmm_flow_workspace_bytes(B, S) + 512;
}

// forward pass that also returns every filtered message alpha[B,T,S] (log domain)
int hmm_flow_forward_alpha_device(const float* init, int64_t init_sb, const float* A, const float* obs, int64_t B, int64_t T, int S,
                                  float* logZ, float* alpha, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    const int Sp = flow_padded_states(S);
    if (!Sp) return FB2_ERR_UNSUPPORTED;
    Arena arena(workspace, workspace_bytes);
    float* ua = arena.take<float>((size_t)B * T * Sp);
    int* ea = arena.take<int>((size_t)B * T * (Sp / 8));
    if (!arena.ok()) {
        set_error("hmm flow forward+alpha: workspace too small (%zu bytes)", workspace_bytes);
        return FB2_ERR_WORKSPACE;
    }
    char* rest = static_cast<char*>(workspace) + arena.used;
    int st = hmm_flow_pass_device(init, init_sb, A, obs, T, B, T, S, logZ, nullptr, ua, ea, T, 0, 0, rest, workspace_bytes - arena.used, stream);
    if (st != FB2_OK) return st;
    const int64_t n = B * T * (int64_t)S;
    flow_alpha_from_history<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(ua, ea, B * T, S, Sp, alpha);
    FB2_LAUNCH_CHECK();
    return FB2_OK;
}


// chunk summary on the dataflow kernel: S chains per batch entry, started from the identity rows, all reading obs[b]
size_t hmm_flow_summary_workspace_bytes(int64_t B, int S) { return hmm_flow_workspace_bytes(B * S, S); }
int hmm_flow_summary_device(const float* A, const float* obs, int64_t B, int64_t T, int S, float* rows_out, double* row_offset, void* workspace,
                            size_t workspace_bytes, cudaStream_t stream) {
    return hmm_flow_pass_device(nullptr, 0, A, obs, T, B * S, T, S, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, workspace, workspace_bytes, stream,
                                nullptr, 0, S, rows_out, row_offset);
}

}  // namespace fb2
