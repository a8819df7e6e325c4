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

// hmm_small_forward_tc: the same recurrence for MANY chains with a shared transition matrix and only log Z wanted (the batched
// DiscreteHMM.log_prob of the reference's examples): a warp owns 16 chains = the 16 rows of mma.sync.m16n8k8 tiles, the messages live in
// C-fragment layout and a step is the small GEMM  V' = V P  [16 x S] [S x S]  in 3xTF32 against P held in shared memory as split B
// fragments; the C fragments of one step are the A fragments of the next without moving data (k slots permuted: slot tig <-> column
// 2 tig, slot tig + 4 <-> column 2 tig + 1).  The warp-per-chain kernel above spends one shuffle, one shared-memory load and one FMA
// per (i, j) pair; here 16 chains share each operand load and the products run on the tensor pipe.
struct SmallFrag {
    uint32_t hi[4], lo[4];
};
__device__ __forceinline__ SmallFrag small_split(float c0, float c1, float c2, float c3) {  // C fragment -> split A fragment (permuted slots)
    SmallFrag f;
    const float v[4] = {c0, c2, c1, c3};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f.hi[i] = (__float_as_uint(v[i]) + 0x1000u) & 0xFFFFE000u;  // round to nearest TF32
        f.lo[i] = __float_as_uint(v[i] - __uint_as_float(f.hi[i]));
    }
    return f;
}
__device__ __forceinline__ void small_mma(float (&d)[4], const uint32_t (&a)[4], float b0, float b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}
__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}
constexpr int SMT_WARPS = 4;
// NB = padded state count / 8 (1, 2, 4 or 8)
template <int NB>
__global__ void __launch_bounds__(SMT_WARPS * 32) hmm_small_forward_tc(SmallArgs p) {
    extern __shared__ __align__(16) float sm_raw[];
    constexpr int SP = NB * 8;
    float4* sB = reinterpret_cast<float4*>(sm_raw);   // [kb][nb][lane] {hi(k0), hi(k0 + 1), lo(k0), lo(k0 + 1)}, B[k = prev][col = curr] = P[k][col]
    float* sCA = sm_raw + NB * NB * 32 * 4;           // [SP] clamped column maxima
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tig = lane & 3;
    const int S = p.S;
    for (int j = tid; j < SP; j += SMT_WARPS * 32) {
        float m = -FLT_MAX;
        if (j < S)
            for (int i = 0; i < S; ++i) m = fmaxf(m, p.transpose ? p.A[(int64_t)j * S + i] : p.A[(int64_t)i * S + j]);
        sCA[j] = j < S ? m : 0.f;
    }
    __syncthreads();
    for (int idx = tid; idx < NB * NB * 32; idx += SMT_WARPS * 32) {
        const int blk = idx >> 5, ln = idx & 31, kb = blk / NB, nb = blk % NB;
        const int k0 = 8 * kb + 2 * (ln & 3), col = 8 * nb + (ln >> 2);
        float v[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = k0 + h;
            v[h] = (k < S && col < S) ? expf((p.transpose ? p.A[(int64_t)col * S + k] : p.A[(int64_t)k * S + col]) - sCA[col]) : 0.f;
        }
        const float h0 = __uint_as_float((__float_as_uint(v[0]) + 0x1000u) & 0xFFFFE000u), h1 = __uint_as_float((__float_as_uint(v[1]) + 0x1000u) & 0xFFFFE000u);
        sB[idx] = make_float4(h0, h1, v[0] - h0, v[1] - h1);
    }
    __syncthreads();
    const int64_t b_base = ((int64_t)blockIdx.x * SMT_WARPS + warp) * 16;
    if (b_base >= p.B) return;
    const int64_t br[2] = {b_base + g, b_base + g + 8};
    const bool live[2] = {br[0] < p.B, br[1] < p.B};
    const int64_t OT = p.obs_T ? p.obs_T : p.T;
    const float* obs_r[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) obs_r[h] = p.obs ? p.obs + (live[h] ? br[h] : b_base) * OT * S : nullptr;  // a missing chain repeats chain 0 of the warp (not stored)
    auto row_of = [&](int64_t t) { return p.reverse ? (p.T - 1 - t) : t; };
    float ca[NB][2];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
        ca[nb][0] = sCA[8 * nb + 2 * tig];
        ca[nb][1] = sCA[8 * nb + 2 * tig + 1];
    }
    // x (log2 domain, rows g / g + 8 of every column block) -> v = w * 2^(x - n), n = ceil(row max); E += n + renormalisation exponent
    float v[NB][4];
    long long E[2] = {0, 0};
    auto apply = [&](const float (&w)[NB][4], const float (&x)[NB][4]) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float xm = -INFINITY;
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) xm = fmaxf(xm, fmaxf(x[nb][2 * h], x[nb][2 * h + 1]));
            xm = quad_max(xm);
            const bool any = xm > -1e30f;
            const float nf = any ? ceilf(xm) : 0.f;
            float m = 0.f;
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float val = w[nb][2 * h + e] * (any ? exp2f(x[nb][2 * h + e] - nf) : 0.f);
                    v[nb][2 * h + e] = val;
                    m = fmaxf(m, val);
                }
            }
            // NaN / inf anywhere in the row must not be hidden by fmaxf: test the sum as well
            float chk = 0.f;
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) chk += v[nb][2 * h] + v[nb][2 * h + 1];
            chk = quad_sum(chk);
            m = quad_max(m);
            int q = 0;
            if (m > 0.f && chk <= FLT_MAX) {
                int extra = 0;
                float pre = 1.f;
                if (m < 1e-20f) {
                    pre = 0x1p64f;
                    m *= 0x1p64f;
                    extra = -64;
                }
                const int qq = ((__float_as_int(m) >> 23) & 0xFF) - 127;
                const float sc = __int_as_float((127 - qq) << 23);
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) {
                    v[nb][2 * h] = (v[nb][2 * h] * pre) * sc;
                    v[nb][2 * h + 1] = (v[nb][2 * h + 1] * pre) * sc;
                }
                q = qq + extra;
            }
            E[h] += (long long)nf + q;
        }
    };
    {
        float ones[NB][4], x[NB][4];
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int h = e >> 1, col = 8 * nb + 2 * tig + (e & 1);
                ones[nb][e] = 1.f;
                x[nb][e] = col < S ? (p.init ? p.init[(live[h] ? br[h] : b_base) * p.init_sb + col] : 0.f) * kLog2eS : -INFINITY;
            }
        apply(ones, x);
    }
    const bool ovec = (S % 2 == 0) && p.obs && (reinterpret_cast<uintptr_t>(p.obs) % 8 == 0);
    float ob_next[NB][4];
    auto load_obs = [&](int64_t t) {
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int col = 8 * nb + 2 * tig;
                float o0 = 0.f, o1 = 0.f;
                if (obs_r[h] && t < p.T) {
                    const float* src = obs_r[h] + row_of(t) * S + col;
                    if (ovec) {
                        if (col < S) {
                            const float2 o = __ldg(reinterpret_cast<const float2*>(src));
                            o0 = o.x;
                            o1 = o.y;
                        }
                    } else {
                        if (col < S) o0 = __ldg(src);
                        if (col + 1 < S) o1 = __ldg(src + 1);
                    }
                }
                ob_next[nb][2 * h] = o0;
                ob_next[nb][2 * h + 1] = o1;
            }
    };
    load_obs(0);
#pragma unroll 1
    for (int64_t t = 0; t < p.T; ++t) {
        float x[NB][4];
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int col = 8 * nb + 2 * tig + (e & 1);
                x[nb][e] = col < S ? (ob_next[nb][e] + ca[nb][e & 1]) * kLog2eS : -INFINITY;
            }
        load_obs(t + 1);  // the next row travels while this step computes
        float acc[NB][4];
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[nb][e] = 0.f;
#pragma unroll
        for (int kb = 0; kb < NB; ++kb) {
            const SmallFrag a = small_split(v[kb][0], v[kb][1], v[kb][2], v[kb][3]);
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
                const float4 bf = sB[(kb * NB + nb) * 32 + lane];
                small_mma(acc[nb], a.lo, bf.x, bf.y);
                small_mma(acc[nb], a.hi, bf.z, bf.w);
                small_mma(acc[nb], a.hi, bf.x, bf.y);
            }
        }
        apply(acc, x);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float tot = 0.f;
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) tot += v[nb][2 * h] + v[nb][2 * h + 1];
        tot = quad_sum(tot);
        if (live[h] && tig == 0) {
            const double z = tot > 0.f ? (double)E[h] * 0.69314718055994530942 + log((double)tot) : (tot == tot ? -INFINITY : (double)tot);
            if (p.logZ) p.logZ[br[h]] = (float)z;
            if (p.logZ_d) p.logZ_d[br[h]] = z;
        }
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
    // 16 chains per warp on the tensor pipe: a win once there are far more chains than the GPU holds warps (measured, T = 1000: S = 16,
    // B = 16384: 2.27 -> 0.72 ms; S = 8, B = 65536: 1.52 -> 0.24 ms; but S = 16, B = 1024: 0.37 -> 0.61 ms and S >= 32 at B = 4096 no gain: a
    // step of a 16-chain warp is a longer dependent chain than a step of a one-chain warp).  FB2_SMALL_TC=1 forces it, 0 switches it off.
    const char* tc_env = getenv("FB2_SMALL_TC");
    const bool tc_forced = tc_env && tc_env[0] == '1', tc_off = tc_env && tc_env[0] == '0';
    if (semiring == FB2_SEMIRING_LOG && a_sb == 0 && !alpha_all && !ell_all && !tc_off && (tc_forced || (S <= 16 && B >= 8192))) {
        const int NB = S <= 8 ? 1 : S <= 16 ? 2 : S <= 32 ? 4 : 8;
        const size_t smem_tc = ((size_t)NB * NB * 32 * 4 + NB * 8) * sizeof(float);
        const unsigned grid_tc = (unsigned)ceil_div(B, SMT_WARPS * 16);
        if (NB == 1) hmm_small_forward_tc<1><<<grid_tc, SMT_WARPS * 32, smem_tc, stream>>>(a);
        else if (NB == 2) hmm_small_forward_tc<2><<<grid_tc, SMT_WARPS * 32, smem_tc, stream>>>(a);
        else if (NB == 4) hmm_small_forward_tc<4><<<grid_tc, SMT_WARPS * 32, smem_tc, stream>>>(a);
        else hmm_small_forward_tc<8><<<grid_tc, SMT_WARPS * 32, smem_tc, stream>>>(a);
        FB2_LAUNCH_CHECK();
        return FB2_OK;
    }
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
