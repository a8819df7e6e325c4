// Forward pass for small state spaces (S <= 64), log semiring: one warp per chain, no inter-CTA traffic at all.
//
// Reference semantics: DiscreteHMM.log_prob funsor/pyro/hmm.py:129-135 (the regime of examples/discrete_hmm.py,
// hidden_dim = 16) evaluated as the vector recurrence with the exp trick of einsum/numpy_log.py:24-47:
//     v_t[j] = (sum_i v_{t-1}[i] P[i,j]) * exp(obs[t,j] + colmax_j),   P = exp(A - colmax)
// The whole chain lives in one warp: lane j owns state j (and j + 32), P sits in shared memory ([i][j], conflict-free), the
// message is exchanged with shuffles, and after every step the warp renormalises by an exact power of two whose exponent is
// accumulated in an integer -- so log Z keeps fp32-mantissa accuracy for any chain length.  B chains run fully in parallel
// (4 warps per CTA, as many CTAs as chains / 4); the generic cooperative kernel served 16 chains per grid barrier.
#include "common.cuh"

namespace fb2 {

constexpr int SM_WARPS = 4;
constexpr float kLog2eS = 1.4426950408889634f;

struct SmallArgs {
    const float* init;  // nullable
    int64_t init_sb;
    const float* A;
    int64_t a_sb;       // 0 = shared transition matrix
    const float* obs;   // nullable
    int64_t B, T;
    int S;
    float* logZ;        // nullable
    float* alpha_all;   // nullable [B,OT,S]
    // forward-backward support: with ell_all the messages are written NORMALISED (log of the mantissa) and their offsets go to
    // ell_all[B,OT]; reverse = 1 runs the backward recursion as a forward pass on A^T (transpose = 1) over rows T-1 .. 0 and
    // writes beta_hat (the contraction WITHOUT the row's observation) with the offset of the incoming message
    double* ell_all;
    double* logZ_d;     // nullable
    int64_t obs_T;      // rows per chain of obs / alpha_all / ell_all (0 = T)
    int transpose, reverse;
};

// NJ = columns per lane: 1 (S <= 32) or 2 (S <= 64); NL = source lanes visited per column block (8, 16 or 32 >= S when NJ == 1).
// All loops are static: rows / lanes beyond S hold zeros, so no shuffle sits behind a run-time condition.
template <int NJ, int NL>
__global__ void __launch_bounds__(SM_WARPS * 32) hmm_small_forward(SmallArgs p) {
    extern __shared__ __align__(16) float sm_raw[];
    constexpr int SP = NJ * 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.S;
    const bool shared_A = p.a_sb == 0;
    float* sP = sm_raw + (shared_A ? 0 : (size_t)warp * SP * SP);  // [i][j], i < SP, j < SP
    float* sCA = sm_raw + (shared_A ? SP * SP : SM_WARPS * SP * SP) + (shared_A ? 0 : warp * SP);
    const int64_t b = (int64_t)blockIdx.x * SM_WARPS + warp;
    const bool live = b < p.B;

    // ---- P = exp(A - colmax) and the clamped column maxima (numpy_log.py:31-32)
    {
        const float* Ab = p.A + (shared_A ? 0 : (live ? b : 0) * p.a_sb);
        const int nthreads = shared_A ? SM_WARPS * 32 : 32;
        const int t0 = shared_A ? threadIdx.x : lane;
        for (int j = t0; j < SP; j += nthreads) {
            float m = -FLT_MAX;
            if (j < S)
                for (int i = 0; i < S; ++i) m = fmaxf(m, p.transpose ? Ab[(int64_t)j * S + i] : Ab[(int64_t)i * S + j]);
            sCA[j] = j < S ? m : 0.f;
        }
        if (shared_A) __syncthreads();
        else __syncwarp();
        for (int idx = t0; idx < SP * SP; idx += nthreads) {
            const int i = idx / SP, j = idx % SP;
            sP[idx] = (i < S && j < S) ? expf((p.transpose ? Ab[(int64_t)j * S + i] : Ab[(int64_t)i * S + j]) - sCA[j]) : 0.f;
        }
        if (shared_A) __syncthreads();
        else __syncwarp();
    }
    if (!live) return;

    float ca[NJ], v[NJ];
    long long E = 0;  // alpha_t[j] = ln2 * E + log v[j]
    const int64_t OT = p.obs_T ? p.obs_T : p.T;
    const float* obs_b = p.obs ? p.obs + b * OT * S : nullptr;
    auto row_of = [&](int64_t t) { return p.reverse ? (p.T - 1 - t) : t; };
    // normalise the lane values x (log2 domain) into v = 2^(x - n), n = ceil(max x); returns n (0 if everything is -inf)
    auto renorm_pow2 = [&](float (&val)[NJ]) -> int {  // scale the warp's message by 2^-q so that its max lies in [1, 2)
        float m = val[0];
#pragma unroll
        for (int q = 1; q < NJ; ++q) m = fmaxf(m, val[q]);
        m = warp_max(m);
        if (!(m > 0.f) || !(m <= FLT_MAX)) return 0;
        int extra = 0;
        float pre = 1.f;
        if (m < 1e-20f) {
            pre = 0x1p64f;
            m *= 0x1p64f;
            extra = -64;
        }
        const int q = ((__float_as_int(m) >> 23) & 0xFF) - 127;
        const float sc = __int_as_float((127 - q) << 23);
#pragma unroll
        for (int k = 0; k < NJ; ++k) val[k] = (val[k] * pre) * sc;
        return q + extra;
    };
    {
        float x[NJ];
        float xm = -INFINITY;
#pragma unroll
        for (int q = 0; q < NJ; ++q) {
            const int j = lane + 32 * q;
            ca[q] = j < S ? sCA[j] : 0.f;
            x[q] = j < S ? (p.init ? p.init[b * p.init_sb + j] : 0.f) * kLog2eS : -INFINITY;
            xm = fmaxf(xm, x[q]);
        }
        xm = warp_max(xm);
        const float nf = xm > -1e30f ? ceilf(xm) : 0.f;
#pragma unroll
        for (int q = 0; q < NJ; ++q) v[q] = xm > -1e30f ? exp2f(x[q] - nf) : 0.f;
        E = (long long)nf + renorm_pow2(v);
    }

    constexpr int PF = 8;  // observation rows are loaded PF steps at a time, one block ahead of their use
    float ob_next[PF][NJ];
    auto load_block = [&](int64_t t0) {
#pragma unroll
        for (int u = 0; u < PF; ++u)
#pragma unroll
            for (int q = 0; q < NJ; ++q) {
                const int j = lane + 32 * q;
                ob_next[u][q] = (obs_b && t0 + u < p.T && j < S) ? __ldg(obs_b + row_of(t0 + u) * S + j) : 0.f;
            }
    };
    load_block(0);
    for (int64_t t0 = 0; t0 < p.T; t0 += PF) {
        float ob[PF][NJ];
#pragma unroll
        for (int u = 0; u < PF; ++u)
#pragma unroll
            for (int q = 0; q < NJ; ++q) ob[u][q] = ob_next[u][q];
        if (t0 + PF < p.T) load_block(t0 + PF);
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            const int64_t t = t0 + u;
            if (t >= p.T) break;
            // s[j] = sum_i v[i] P[i,j]: the message comes from the other lanes by shuffle, P from shared memory
            float acc[NJ][2];
#pragma unroll
            for (int q = 0; q < NJ; ++q) acc[q][0] = acc[q][1] = 0.f;
#pragma unroll
            for (int qi = 0; qi < NJ; ++qi) {
#pragma unroll
                for (int l = 0; l < NL; ++l) {
                    const int i = l + 32 * qi;
                    const float vi = __shfl_sync(0xffffffffu, v[qi], l);
#pragma unroll
                    for (int q = 0; q < NJ; ++q) acc[q][l & 1] = fmaf(vi, sP[i * SP + lane + 32 * q], acc[q][l & 1]);
                }
            }
            // times exp(obs + colmax), with a common power-of-two shift taken from the largest exponent argument
            float x[NJ], xm = -INFINITY;
#pragma unroll
            for (int q = 0; q < NJ; ++q) {
                const int j = lane + 32 * q;
                x[q] = j < S ? (ob[u][q] + ca[q]) * kLog2eS : -INFINITY;
                xm = fmaxf(xm, x[q]);
            }
            xm = warp_max(xm);
            const float nf = xm > -1e30f ? ceilf(xm) : 0.f;
#pragma unroll
            for (int q = 0; q < NJ; ++q) v[q] = (acc[q][0] + acc[q][1]) * (xm > -1e30f ? exp2f(x[q] - nf) : 0.f);
            const long long E_in = E;
            E += (long long)nf + renorm_pow2(v);
            if (p.alpha_all) {
                const int64_t slot = b * OT + row_of(t);
#pragma unroll
                for (int q = 0; q < NJ; ++q) {
                    const int j = lane + 32 * q;
                    if (j < S) {
                        float outv;
                        if (p.reverse) {
                            const float sacc = acc[q][0] + acc[q][1];
                            outv = sacc > 0.f ? logf(sacc) + ca[q] : -INFINITY;  // beta_hat: no observation of this row, offset E_in
                        } else if (p.ell_all) {
                            outv = v[q] > 0.f ? logf(v[q]) : -INFINITY;
                        } else {
                            outv = v[q] > 0.f ? (float)((double)E * 0.69314718055994530942 + (double)logf(v[q])) : -INFINITY;
                        }
                        p.alpha_all[slot * S + j] = outv;
                    }
                }
                if (p.ell_all && lane == 0) p.ell_all[slot] = (double)(p.reverse ? E_in : E) * 0.69314718055994530942;
            }
        }
    }
    float tot = 0.f;
#pragma unroll
    for (int q = 0; q < NJ; ++q) tot += v[q];
    tot = warp_sum(tot);
    if (lane == 0) {
        const double z = tot > 0.f ? (double)E * 0.69314718055994530942 + log((double)tot) : (tot == tot ? -INFINITY : (double)tot);
        if (p.logZ) p.logZ[b] = (float)z;
        if (p.logZ_d) p.logZ_d[b] = z;
    }
}

// Max-plus twin (Viterbi for small state spaces): lane j scans the predecessors i in ascending order with a strict '>' on
//   delta[i] + (A[i,j] + obs[t,j])      (the reference's evaluation order, hmm.py:130; first maximiser wins)
// so values and back-pointers are bit-identical to the serial fp32 recurrence.  A is kept in shared memory untransformed.
template <int NJ, int NL>
__global__ void __launch_bounds__(SM_WARPS * 32) hmm_small_maxplus(SmallArgs p, int32_t* backptr, int32_t* last) {
    extern __shared__ __align__(16) float sm_raw[];
    constexpr int SP = NJ * 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.S;
    const bool shared_A = p.a_sb == 0;
    float* sA = sm_raw + (shared_A ? 0 : (size_t)warp * SP * SP);
    const int64_t b = (int64_t)blockIdx.x * SM_WARPS + warp;
    const bool live = b < p.B;
    {
        const float* Ab = p.A + (shared_A ? 0 : (live ? b : 0) * p.a_sb);
        const int nthreads = shared_A ? SM_WARPS * 32 : 32;
        const int t0 = shared_A ? threadIdx.x : lane;
        for (int idx = t0; idx < SP * SP; idx += nthreads) {
            const int i = idx / SP, j = idx % SP;
            sA[idx] = (i < S && j < S) ? Ab[(int64_t)i * S + j] : kNegInf;
        }
        if (shared_A) __syncthreads();
        else __syncwarp();
    }
    if (!live) return;
    float d[NJ];
#pragma unroll
    for (int q = 0; q < NJ; ++q) {
        const int j = lane + 32 * q;
        d[q] = j < S ? (p.init ? p.init[b * p.init_sb + j] : 0.f) : kNegInf;
    }
    const int64_t OT = p.obs_T ? p.obs_T : p.T;
    const float* obs_b = p.obs ? p.obs + b * OT * S : nullptr;
    auto row_of = [&](int64_t t) { return p.reverse ? (p.T - 1 - t) : t; };
    for (int64_t t = 0; t < p.T; ++t) {
        float ob[NJ], best[NJ];
        int arg[NJ];
#pragma unroll
        for (int q = 0; q < NJ; ++q) {
            const int j = lane + 32 * q;
            ob[q] = (obs_b && j < S) ? __ldg(obs_b + t * S + j) : 0.f;
            best[q] = kNegInf;
            arg[q] = 0;
        }
#pragma unroll
        for (int qi = 0; qi < NJ; ++qi)
#pragma unroll
            for (int l = 0; l < NL; ++l) {
                const int i = l + 32 * qi;
                const float di = __shfl_sync(0xffffffffu, d[qi], l);
#pragma unroll
                for (int q = 0; q < NJ; ++q) {
                    const float v = di + (sA[i * SP + lane + 32 * q] + ob[q]);
                    if (v > best[q]) {
                        best[q] = v;
                        arg[q] = i;
                    }
                }
            }
#pragma unroll
        for (int q = 0; q < NJ; ++q) {
            const int j = lane + 32 * q;
            d[q] = j < S ? best[q] : kNegInf;
            if (j < S) {
                if (backptr) backptr[(b * p.T + t) * S + j] = arg[q];
                if (p.alpha_all) p.alpha_all[(b * p.T + t) * S + j] = best[q];
            }
        }
    }
    // best final state, first index on ties
    float m = kNegInf;
    int am = 0x7fffffff;
#pragma unroll
    for (int q = 0; q < NJ; ++q) {
        const int j = lane + 32 * q;
        if (j < S && (d[q] > m || (d[q] == m && j < am))) {
            m = d[q];
            am = j;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, m, o);
        const int oa = __shfl_xor_sync(0xffffffffu, am, o);
        if (om > m || (om == m && oa < am)) {
            m = om;
            am = oa;
        }
    }
    if (lane == 0) {
        p.logZ[b] = m;
        if (last) last[b] = am == 0x7fffffff ? 0 : am;
    }
}

bool hmm_small_supported(int semiring, int64_t a_st, int S) {
    return (semiring == FB2_SEMIRING_LOG || semiring == FB2_SEMIRING_MAX) && a_st == 0 && S >= 1 && S <= 64;
}

int hmm_small_forward_device(int semiring, const float* init, int64_t init_sb, const float* A, int64_t a_sb, const float* obs, int64_t B,
                             int64_t T, int S, float* logZ, float* alpha_all, int32_t* backptr, int32_t* last, cudaStream_t stream,
                             double* ell_all, double* logZ_d, int64_t obs_T, int transpose, int reverse) {
    SmallArgs a;
    a.init = init; a.init_sb = init_sb; a.A = A; a.a_sb = a_sb; a.obs = obs; a.B = B; a.T = T; a.S = S; a.logZ = logZ; a.alpha_all = alpha_all;
    a.ell_all = ell_all; a.logZ_d = logZ_d; a.obs_T = obs_T; a.transpose = transpose; a.reverse = reverse;
    const int NJ = S <= 32 ? 1 : 2;
    const int SP = NJ * 32;
    const size_t smem = ((size_t)(a_sb == 0 ? 1 : SM_WARPS) * (SP * SP + SP)) * sizeof(float);
    const unsigned grid = (unsigned)ceil_div(B, SM_WARPS);
#define FB2_SMALL(NJv, NLv)                                                                                                    \
    do {                                                                                                                        \
        FB2_CUDA(cudaFuncSetAttribute(hmm_small_forward<NJv, NLv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
        hmm_small_forward<NJv, NLv><<<grid, SM_WARPS * 32, smem, stream>>>(a);                                                  \
    } while (0)
#define FB2_SMALLMAX(NJv, NLv)                                                                                                 \
    do {                                                                                                                        \
        FB2_CUDA(cudaFuncSetAttribute(hmm_small_maxplus<NJv, NLv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
        hmm_small_maxplus<NJv, NLv><<<grid, SM_WARPS * 32, smem, stream>>>(a, backptr, last);                                   \
    } while (0)
    if (semiring == FB2_SEMIRING_MAX) {
        if (S <= 8) FB2_SMALLMAX(1, 8);
        else if (S <= 16) FB2_SMALLMAX(1, 16);
        else if (S <= 32) FB2_SMALLMAX(1, 32);
        else FB2_SMALLMAX(2, 32);
    } else if (S <= 8) FB2_SMALL(1, 8);
    else if (S <= 16) FB2_SMALL(1, 16);
    else if (S <= 32) FB2_SMALL(1, 32);
    else FB2_SMALL(2, 32);
#undef FB2_SMALL
#undef FB2_SMALLMAX
    FB2_LAUNCH_CHECK();
    return FB2_OK;
}

}  // namespace fb2
