// Factored HMM passes: forward (log / max-plus), Viterbi back-pointers + backtrace, backward.
//
// Reference semantics: funsor/pyro/hmm.py:129-135 evaluated as the vector recurrence
//   alpha_t[j] = (+)_i ( alpha_{t-1}[i] (x) A[i,j] ) (x) obs[t,j],   alpha_{-1} = init
// instead of materialising trans + obs and calling sequential_sum_product (sum_product.py:652-701).
//
// Layout / execution model ("generic" path, any S, both semirings):
//   * chains are processed in groups of R (R in {1,4,16}); a group is served by Gg co-resident CTAs
//     (cooperative launch), CTA c owning the W = ceil(S/Gg) output states [c*W, c*W+W).
//   * the message vector lives in a double-buffered global (L2-resident) array vec[2][B][S]; after
//     every step the group's CTAs meet at a counter barrier and re-read the full vector with ld.cg.
//   * the CTA's slice of A is re-read every step: it is served by L1/L2 (S*W*4 bytes per CTA).
//   * obs[b,t,slice] is prefetched one step ahead into shared memory, so HBM latency is off the
//     step-to-step critical path.
#include "common.cuh"

#include <stdlib.h>

namespace fb2 {

// hmm_flow.cu
bool hmm_flow_supported(int semiring, int64_t a_sb, int64_t a_st, int S, bool wants_alpha);
size_t hmm_flow_workspace_bytes(int64_t B, int S);
int hmm_flow_forward_device(const float* init, int64_t init_sb, const float* A, const float* obs, int64_t B, int64_t T, int S,
                            float* logZ, double* logZ_d, void* workspace, size_t workspace_bytes, cudaStream_t stream);
size_t hmm_flow_marginals_workspace_bytes(int64_t B, int64_t T, int S);
int hmm_flow_marginals_device(const float* init, int64_t init_sb, const float* A, const float* obs, int64_t B, int64_t T, int S,
                              float* logZ, float* gamma, float* xi_sum, void* workspace, size_t workspace_bytes, cudaStream_t stream);
size_t hmm_flow_alpha_workspace_bytes(int64_t B, int64_t T, int S);
int hmm_flow_forward_alpha_device(const float* init, int64_t init_sb, const float* A, const float* obs, int64_t B, int64_t T, int S,
                                  float* logZ, float* alpha, void* workspace, size_t workspace_bytes, cudaStream_t stream);
size_t hmm_flow_summary_workspace_bytes(int64_t B, int S);
int hmm_flow_summary_device(const float* A, const float* obs, int64_t B, int64_t T, int S, float* rows_out, double* row_offset, void* workspace,
                            size_t workspace_bytes, cudaStream_t stream);
// hmm_small.cu
bool hmm_small_supported(int semiring, int64_t a_st, int S);
int hmm_small_forward_device(int semiring, const float* init, int64_t init_sb, const float* A, int64_t a_sb, const float* obs, int64_t B,
                             int64_t T, int S, float* logZ, float* alpha_all, int32_t* backptr, int32_t* last, cudaStream_t stream,
                             double* ell_all = nullptr, double* logZ_d = nullptr, int64_t obs_T = 0, int transpose = 0, int reverse = 0);
// FB2_DISABLE_FLOW=1 forces the generic kernel (A/B testing of the two GPU paths)
static bool stream_disabled() {
    const char* e = getenv("FB2_DISABLE_STREAM");
    return e && e[0] == '1';
}
static thread_local int tl_force_coop = 0;
// retry path of the host-buffer entry points after an aborted dataflow pass: only kernels launched cooperatively
void hmm_force_cooperative_kernels(int on) { tl_force_coop = on; }
static bool flow_disabled() {
    const char* e = getenv("FB2_DISABLE_FLOW");
    return tl_force_coop || (e && e[0] == '1');
}

struct HmmArgs {
    const float* init;  // nullable -> semiring one
    int64_t init_sb;
    const float* A;
    int64_t a_sb, a_st;
    const float* obs;  // nullable -> 0
    int64_t B, T;
    int S;
    float* vec;          // [2][B][S]
    unsigned* counters;  // [groups_conc]
    float* alpha_all;    // nullable [B,T,S]
    int32_t* backptr;    // nullable [B,T,S]
    float* out;          // [B]   (+)_j alpha_{T-1}[j]
    int32_t* last;       // nullable [B] argmax_j alpha_{T-1}[j]
    const float* final_; // nullable [B,S] folded into out (adjoint use): out = (+)_j alpha[j] (x) final[j]
    int Gg, W, LW;
    int groups_total, groups_conc;
    double* ell_all;     // nullable [B,T]: when given, alpha_all holds NORMALISED messages and ell_all their offsets
    double* out_d;       // nullable [B]: out in double (LOG)
    float* alpha_last;   // nullable [B,S]: alpha_{T-1} (forward) with its offset folded in (or relative to ell_last)
    double* ell_last;    // nullable [B]: when given, alpha_last is written WITHOUT the running offset and the offset goes here
    int cpo;             // chains per obs/A batch entry (1 normally; S for chunk summaries: chain = b*S + row)
    int init_identity;   // 1: alpha_{-1} of chain c is the semiring identity row (c % cpo)
    int64_t obs_T;       // hmm_logsum_stream: rows per chain of obs / alpha_all / ell_all (0 = T)
    int reverse;         // hmm_logsum_stream: step t consumes row T-1-t (backward pass run as a forward pass on A^T)
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// All CTAs of one group meet here. `target` is the cumulative arrival count to wait for.
__device__ __forceinline__ void group_barrier(unsigned* ctr, unsigned target, int Gg) {
    __syncthreads();
    if (Gg > 1 && threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        while ((int)(ld_acquire_u32(ctr) - target) < 0) {
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// forward recurrence, generic path
// smem: s_alpha[R][S] | s_red_a[256*R] | s_red_b[256*R] | s_obs[2][R][W]
template <int SEMI, int R, bool ARG>
__global__ void __launch_bounds__(256) hmm_forward_generic(HmmArgs p) {
    extern __shared__ __align__(16) float smem[];
    const int S = p.S, W = p.W, LW = p.LW, NSUB = 256 / LW;
    float* s_alpha = smem;
    float* s_red_a = s_alpha + (size_t)R * S;
    float* s_red_b = s_red_a + 256 * R;
    float* s_obs = s_red_b + 256 * R;
    // LOG only: messages are kept relative to a running offset ell (double), so fp32 resolution does not
    // degrade as |alpha| grows with t:  alpha_t = vec_t + ell_t,  ell_t = sum_{tau<=t} max_i vec_{tau-1}[i]
    __shared__ double s_ell[16];
    const int tid = threadIdx.x;
    const int gconc = blockIdx.x / p.Gg, slice = blockIdx.x % p.Gg;
    const int j0 = slice * W;
    const int jw = max(0, min(W, S - j0));
    const int lw = tid % LW, isub = tid / LW;
    unsigned* ctr = p.counters + gconc;
    unsigned epoch = 0;

    for (int g = gconc; g < p.groups_total; g += p.groups_conc) {
        const int64_t b0 = (int64_t)g * R;
        const int nb = (int)min((int64_t)R, p.B - b0);
        if (tid < 16) s_ell[tid] = 0.0;
        // ---- alpha_{-1} = init  -> vec[0]
        for (int idx = tid; idx < nb * jw; idx += 256) {
            int r = idx / jw, j = j0 + idx % jw;
            float v = p.init ? p.init[(b0 + r) * p.init_sb + j] : 0.f;
            if (p.init_identity) v = (j == (int)((b0 + r) % p.cpo)) ? 0.f : kNegInf;
            p.vec[((b0 + r) * S) + j] = v;
        }
        // prefetch obs for t = 0
        if (p.obs)
            for (int idx = tid; idx < nb * jw; idx += 256) {
                int r = idx / jw, jj = idx % jw;
                s_obs[(0 * R + r) * W + jj] = p.obs[((b0 + r) / p.cpo * p.T + 0) * S + j0 + jj];
            }
        group_barrier(ctr, (++epoch) * p.Gg, p.Gg);

        for (int64_t t = 0; t < p.T; ++t) {
            const float* vin = p.vec + (size_t)(t & 1) * p.B * S;
            float* vout = p.vec + (size_t)((t + 1) & 1) * p.B * S;
            // ---- stage the full previous vector of every chain of the group
            for (int idx = tid; idx < R * S; idx += 256) {
                int r = idx / S, i = idx % S;
                s_alpha[idx] = (r < nb) ? __ldcg(vin + (b0 + r) * S + i) : kNegInf;
            }
            // ---- prefetch obs for t+1 (latency hidden behind this step's compute + barrier)
            if (p.obs && t + 1 < p.T)
                for (int idx = tid; idx < nb * jw; idx += 256) {
                    int r = idx / jw, jj = idx % jw;
                    s_obs[(((t + 1) & 1) * R + r) * W + jj] = p.obs[((b0 + r) / p.cpo * p.T + t + 1) * S + j0 + jj];
                }
            __syncthreads();
            if (SEMI == FB2_SEMIRING_LOG) {
                const int warp = tid / 32, lane = tid % 32;
                for (int r = warp; r < R; r += 8) {
                    float m = -FLT_MAX;
                    for (int i = lane; i < S; i += 32) m = fmaxf(m, s_alpha[r * S + i]);
                    m = warp_max(m);
                    for (int i = lane; i < S; i += 32) s_alpha[r * S + i] -= m;
                    if (lane == 0 && r < nb) s_ell[r] += (double)m;
                }
                __syncthreads();
            }
            const float* s_ob = s_obs + (size_t)((t & 1) * R) * W;

            for (int c0 = 0; c0 < jw; c0 += LW) {
                const int jj = c0 + lw;
                const bool colok = jj < jw;
                float acc_a[R], acc_b[R];  // LOG: (max, sum)   MAX: (best, arg as float bits)
                if (colok) {
                    float ob[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) ob[r] = (p.obs && r < nb) ? s_ob[r * W + jj] : 0.f;
                    const bool per_chain = p.a_sb != 0;
                    const float* Abase = p.A + t * p.a_st + j0 + jj;
                    int64_t aoff[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) aoff[r] = per_chain ? ((b0 + (r < nb ? r : nb - 1)) / p.cpo) * p.a_sb : 0;
                    if (SEMI == FB2_SEMIRING_LOG) {
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            acc_a[r] = -FLT_MAX;
                            acc_b[r] = 0.f;
                        }
                        for (int i = isub; i < S; i += NSUB) {
                            const float a0 = per_chain ? 0.f : __ldg(Abase + (int64_t)i * S);
#pragma unroll
                            for (int r = 0; r < R; ++r) {
                                const float a = per_chain ? __ldg(Abase + aoff[r] + (int64_t)i * S) : a0;
                                acc_a[r] = fmaxf(acc_a[r], s_alpha[r * S + i] + a);
                            }
                        }
                        for (int i = isub; i < S; i += NSUB) {
                            const float a0 = per_chain ? 0.f : __ldg(Abase + (int64_t)i * S);
#pragma unroll
                            for (int r = 0; r < R; ++r) {
                                const float a = per_chain ? __ldg(Abase + aoff[r] + (int64_t)i * S) : a0;
                                acc_b[r] += __expf(s_alpha[r * S + i] + a - acc_a[r]);
                            }
                        }
                    } else {
                        int arg[R];
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            acc_a[r] = kNegInf;
                            arg[r] = isub < S ? isub : 0;
                        }
                        for (int i = isub; i < S; i += NSUB) {
                            const float a0 = per_chain ? 0.f : __ldg(Abase + (int64_t)i * S);
#pragma unroll
                            for (int r = 0; r < R; ++r) {
                                const float a = per_chain ? __ldg(Abase + aoff[r] + (int64_t)i * S) : a0;
                                const float v = s_alpha[r * S + i] + (a + ob[r]);  // (A + obs) first: hmm.py:130
                                if (v > acc_a[r]) {
                                    acc_a[r] = v;
                                    arg[r] = i;
                                }
                            }
                        }
#pragma unroll
                        for (int r = 0; r < R; ++r) acc_b[r] = __int_as_float(arg[r]);
                    }
                }
                // ---- combine the NSUB partials of every (chain, column) in a fixed order
                if (colok) {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        s_red_a[(isub * R + r) * LW + lw] = acc_a[r];
                        s_red_b[(isub * R + r) * LW + lw] = acc_b[r];
                    }
                }
                __syncthreads();
                // (R * LW can exceed the 256 threads -- 16 chains x 32 columns when many groups share the GPU -- hence the loop:
                //  a plain `if (tid < R * LW)` left chains 8..15 of every group without a result for S * B > ~38000)
                for (int fin = tid; fin < R * LW; fin += 256) {
                    const int r = fin / LW, l = fin % LW;
                    const int col = c0 + l;
                    if (r < nb && col < jw) {
                        const int64_t b = b0 + r;
                        float res;
                        if (SEMI == FB2_SEMIRING_LOG) {
                            float m = -FLT_MAX;
                            for (int q = 0; q < NSUB; ++q) m = fmaxf(m, s_red_a[(q * R + r) * LW + l]);
                            float s = 0.f;
                            for (int q = 0; q < NSUB; ++q)
                                s += s_red_b[(q * R + r) * LW + l] * __expf(s_red_a[(q * R + r) * LW + l] - m);
                            res = m + logf(s) + (p.obs ? s_ob[r * W + col] : 0.f);
                        } else {
                            float best = kNegInf;
                            int arg = 0x7fffffff;
                            for (int q = 0; q < NSUB; ++q) {
                                float v = s_red_a[(q * R + r) * LW + l];
                                int ai = __float_as_int(s_red_b[(q * R + r) * LW + l]);
                                if (v > best || (v == best && ai < arg)) {
                                    best = v;
                                    arg = ai;
                                }
                            }
                            res = best;
                            if (ARG) p.backptr[(b * p.T + t) * S + j0 + col] = arg;
                        }
                        vout[b * S + j0 + col] = res;
                        if (p.alpha_all) {
                            float off = (SEMI == FB2_SEMIRING_LOG && !p.ell_all) ? (float)s_ell[r] : 0.f;
                            p.alpha_all[(b * p.T + t) * S + j0 + col] = res + off;
                        }
                        if (SEMI == FB2_SEMIRING_LOG && p.ell_all && j0 + col == 0) p.ell_all[b * p.T + t] = s_ell[r];
                    }
                }
                __syncthreads();
            }
            group_barrier(ctr, (++epoch) * p.Gg, p.Gg);
        }
        // ---- out[b] = (+)_j alpha_{T-1}[j] (x) final[j]  (slice 0 of the group does it)
        if (slice == 0) {
            const float* vfin = p.vec + (size_t)(p.T & 1) * p.B * S;
            const int warp = tid / 32, lane = tid % 32;
            for (int r = warp; r < nb; r += 8) {
                const int64_t b = b0 + r;
                float m = (SEMI == FB2_SEMIRING_LOG) ? -FLT_MAX : kNegInf;
                int arg = 0x7fffffff;
                for (int j = lane; j < S; j += 32) {
                    float v = __ldcg(vfin + b * S + j) + (p.final_ ? p.final_[b * S + j] : 0.f);
                    if (SEMI == FB2_SEMIRING_LOG) {
                        m = fmaxf(m, v);
                    } else if (v > m || (v == m && j < arg)) {
                        m = v;
                        arg = j;
                    }
                }
                if (SEMI == FB2_SEMIRING_LOG) {
                    m = warp_max(m);
                    float s = 0.f;
                    for (int j = lane; j < S; j += 32)
                        s += __expf(__ldcg(vfin + b * S + j) + (p.final_ ? p.final_[b * S + j] : 0.f) - m);
                    s = warp_sum(s);
                    if (lane == 0) {
                        double z = s_ell[r] + (double)(m + logf(s));
                        p.out[b] = (float)z;
                        if (p.out_d) p.out_d[b] = z;
                    }
                } else {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        float om = __shfl_xor_sync(0xffffffffu, m, o);
                        int oa = __shfl_xor_sync(0xffffffffu, arg, o);
                        if (om > m || (om == m && oa < arg)) {
                            m = om;
                            arg = oa;
                        }
                    }
                    if (lane == 0) {
                        p.out[b] = m;
                        if (p.last) p.last[b] = arg;
                    }
                }
            }
        }
        if (p.alpha_last) {
            const float* vfin = p.vec + (size_t)(p.T & 1) * p.B * S;
            for (int idx = tid; idx < nb * jw; idx += 256) {
                int r = idx / jw, j = j0 + idx % jw;
                float off = (SEMI == FB2_SEMIRING_LOG && !p.ell_last) ? (float)s_ell[r] : 0.f;
                p.alpha_last[(b0 + r) * S + j] = __ldcg(vfin + (b0 + r) * S + j) + off;
            }
            if (p.ell_last && slice == 0 && tid < nb) p.ell_last[b0 + tid] = (SEMI == FB2_SEMIRING_LOG) ? s_ell[tid] : 0.0;
        }
        // the next wave re-uses vec[] of other chains only, but s_obs/s_alpha/s_ell are re-used: sync
        group_barrier(ctr, (++epoch) * p.Gg, p.Gg);
    }
}

// ---------------------------------------------------------------------------------------------
// max-plus forward, streaming path: shared (time-homogeneous) A, few chains, large S (Viterbi, BASELINE config 3).
//   delta_t[j] = max_i( delta_{t-1}[i] + (A[i,j] + obs[t,j]) ),  first-index argmax, same fp32 order as the generic path.
// A no longer fits anywhere on chip (64 MB at S=4096) and is re-read every step, so the step is bound by L2 -> SM
// bandwidth.  Each CTA owns 32 columns = one 128-byte line per row of A; warp w streams rows w, w+8, ... with 16
// independent line loads in flight per thread, keeps (best, arg) for R chains in registers (A is read once for all of
// them), and the 8 warps are combined through shared memory.  One grid barrier per step.
// A[S,S] row-major -> strips[tile][row][32] (columns beyond S padded with -inf): every CTA then streams ONE contiguous
// 128*S-byte strip per step instead of 128-byte pieces 4*S bytes apart (DRAM page locality; A does not stay in L2 at
// S=4096: 64 MB against 2 x 63 MB of die-local L2).
// With RP > 1 a tile is TW = 32 / RP columns wide and a warp instruction covers RP consecutive rows of it (lane = rp * TW + cl), so
// that S / TW CTAs share the work instead of S / 32: strips[tile][row][TW], rows padded with -inf to a multiple of RP.
__global__ void stream_retile_A(const float* A, int S, float* strips, int RP) {
    const int TW = 32 / RP;
    const int SR = (S + RP - 1) / RP * RP;  // padded rows
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int ntiles = (S + TW - 1) / TW;
    if (idx >= (int64_t)ntiles * SR * TW) return;
    const int l = (int)(idx % TW);
    const int64_t r = idx / TW;
    const int row = (int)(r % SR), tile = (int)(r / SR);
    const int col = tile * TW + l;
    strips[idx] = (col < S && row < S) ? A[(int64_t)row * S + col] : kNegInf;
}
constexpr int STREAM_WARPS = 8;
constexpr int STREAM_THREADS = STREAM_WARPS * 32;
// One column per lane (RP = 1): the original form of the kernel, kept for S > 2368 (BASELINE C3: S = 4096), where the general
// row-phase version below measured 23 % slower per step (16.2 vs 13.2 us).
template <int R, bool ARG>
__global__ void __launch_bounds__(STREAM_THREADS) hmm_maxplus_stream32(HmmArgs p) {
    constexpr int NW = STREAM_WARPS, NTH = STREAM_THREADS;
    extern __shared__ __align__(16) float smem[];
    const int S = p.S;
    float* s_delta = smem;                                     // [R][S]
    float* s_best = s_delta + (((size_t)R * S + 3) & ~(size_t)3);  // [NW][R][32] (16-byte aligned)
    int* s_arg = reinterpret_cast<int*>(s_best + NW * R * 32);  // [NW][R][32]
    float* s_strip = reinterpret_cast<float*>(s_arg + NW * R * 32);  // [keep][32]: the first `keep` rows of this CTA's strip
    // When every CTA owns exactly one strip, the part of it that fits in the remaining shared memory stays on chip for
    // the whole pass and only the rest is streamed each step.
    const int keep = p.W;  // rows resident in shared memory (0 if a CTA serves several strips)
    if (keep > 0 && (int)blockIdx.x < (S + 31) / 32) {
        const float4* src = reinterpret_cast<const float4*>(p.A + (size_t)blockIdx.x * S * 32);
        float4* dst = reinterpret_cast<float4*>(s_strip);
        for (int idx = threadIdx.x; idx < keep * 8; idx += NTH) dst[idx] = __ldg(src + idx);
    }
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ntiles = (S + 31) / 32;
    unsigned epoch = 0;
    unsigned* ctr = p.counters;

    for (int64_t b0 = 0; b0 < p.B; b0 += R) {
        const int nb = (int)min((int64_t)R, p.B - b0);
        // ---- delta_{-1} = init -> vec[0]
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
            for (int idx = tid; idx < nb * 32; idx += NTH) {
                const int r = idx >> 5, j = tile * 32 + (idx & 31);
                if (j < S) p.vec[(b0 + r) * S + j] = p.init ? p.init[(b0 + r) * p.init_sb + j] : 0.f;
            }
        group_barrier(ctr, (++epoch) * gridDim.x, gridDim.x);

        for (int64_t t = 0; t < p.T; ++t) {
            const float* vin = p.vec + (size_t)(t & 1) * p.B * S;
            float* vout = p.vec + (size_t)((t + 1) & 1) * p.B * S;
            for (int i = tid; i < S; i += NTH) {  // R independent loads per iteration, no division
#pragma unroll
                for (int r = 0; r < R; ++r) s_delta[r * S + i] = (r < nb) ? __ldcg(vin + (b0 + r) * S + i) : kNegInf;
            }
            __syncthreads();
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int col = tile * 32 + lane;
                const bool colok = col < S;
                float ob[R], best[R];
                int arg[R];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    ob[r] = (p.obs && colok && r < nb) ? __ldg(p.obs + ((b0 + r) * p.T + t) * S + col) : 0.f;
                    best[r] = kNegInf;
                    arg[r] = warp < S ? warp : 0;
                }
                const float* Acol = p.A + (size_t)tile * S * 32 + lane;  // strip-major copy of A: [tile][row][32]
                int i = warp;
                for (; i < keep; i += NW * 8) {  // resident rows (keep is a multiple of 8*NW)
                    float a[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) a[u] = s_strip[(i + NW * u) * 32 + lane];
#pragma unroll
                    for (int u = 0; u < 8; ++u)
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const float v = s_delta[r * S + i + NW * u] + (a[u] + ob[r]);
                            if (v > best[r]) {
                                best[r] = v;
                                arg[r] = i + NW * u;
                            }
                        }
                }
                for (; i + NW * 15 < S; i += NW * 16) {
                    float a[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) a[u] = __ldg(Acol + (size_t)(i + NW * u) * 32);
#pragma unroll
                    for (int u = 0; u < 16; ++u)
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const float v = s_delta[r * S + i + NW * u] + (a[u] + ob[r]);  // (A + obs) first: hmm.py:130
                            if (v > best[r]) {
                                best[r] = v;
                                arg[r] = i + NW * u;
                            }
                        }
                }
                for (; i < S; i += NW) {
                    const float a = __ldg(Acol + (size_t)i * 32);
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const float v = s_delta[r * S + i] + (a + ob[r]);
                        if (v > best[r]) {
                            best[r] = v;
                            arg[r] = i;
                        }
                    }
                }
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    s_best[(warp * R + r) * 32 + lane] = best[r];
                    s_arg[(warp * R + r) * 32 + lane] = arg[r];
                }
                __syncthreads();
                for (int fin = tid; fin < R * 32; fin += NTH) {  // R * 32 may exceed the thread count (R = 16)
                    const int r = fin >> 5, l = fin & 31, c = tile * 32 + l;
                    if (r < nb && c < S) {
                        float bb = kNegInf;
                        int aa = 0x7fffffff;
                        for (int w = 0; w < NW; ++w) {
                            const float v = s_best[(w * R + r) * 32 + l];
                            const int ai = s_arg[(w * R + r) * 32 + l];
                            if (v > bb || (v == bb && ai < aa)) {
                                bb = v;
                                aa = ai;
                            }
                        }
                        vout[(b0 + r) * S + c] = bb;
                        if (ARG) p.backptr[((b0 + r) * p.T + t) * S + c] = aa;
                        if (p.alpha_all) p.alpha_all[((b0 + r) * p.T + t) * S + c] = bb;
                    }
                }
                __syncthreads();
            }
            group_barrier(ctr, (++epoch) * gridDim.x, gridDim.x);
        }
        // ---- out[b] = max_j delta_{T-1}[j], first-index argmax
        if (blockIdx.x == 0) {
            const float* vfin = p.vec + (size_t)(p.T & 1) * p.B * S;
            for (int r = warp; r < nb; r += NW) {
                float m = kNegInf;
                int arg = 0x7fffffff;
                for (int j = lane; j < S; j += 32) {
                    const float v = __ldcg(vfin + (b0 + r) * S + j);
                    if (v > m || (v == m && j < arg)) {
                        m = v;
                        arg = j;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float om = __shfl_xor_sync(0xffffffffu, m, o);
                    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
                    if (om > m || (om == m && oa < arg)) {
                        m = om;
                        arg = oa;
                    }
                }
                if (lane == 0) {
                    p.out[b0 + r] = m;
                    if (p.last) p.last[b0 + r] = arg;
                }
            }
        }
        group_barrier(ctr, (++epoch) * gridDim.x, gridDim.x);  // vec[] is reused by the next chains
    }
}

template <int R, bool ARG, int RP>
__global__ void __launch_bounds__(STREAM_THREADS) hmm_maxplus_stream(HmmArgs p) {
    constexpr int NW = STREAM_WARPS, NTH = STREAM_THREADS;
    constexpr int TW = 32 / RP;  // columns per tile; lane = rp * TW + cl scans rows rp, rp + RP, ... of column cl
    extern __shared__ __align__(16) float smem[];
    const int S = p.S;
    const int SR = (S + RP - 1) / RP * RP;  // rows incl. padding
    const int NS = SR / RP;                 // row slots (one 128-byte line of the strip each)
    float* s_delta = smem;                                      // [R][SR]
    float* s_best = s_delta + (((size_t)R * SR + 3) & ~(size_t)3);  // [NW][R][TW] (16-byte aligned)
    int* s_arg = reinterpret_cast<int*>(s_best + NW * R * TW);   // [NW][R][TW]
    float* s_strip = reinterpret_cast<float*>(s_arg + NW * R * TW);  // [keep][32]: the first `keep` row slots of this CTA's strip
    const int ntiles = (S + TW - 1) / TW;
    // When every CTA owns exactly one strip, the part of it that fits in the remaining shared memory stays on chip for
    // the whole pass and only the rest is streamed each step.
    const int keep = p.W;  // row slots resident in shared memory (0 if a CTA serves several strips)
    if (keep > 0 && (int)blockIdx.x < ntiles) {
        const float4* src = reinterpret_cast<const float4*>(p.A + (size_t)blockIdx.x * SR * TW);
        float4* dst = reinterpret_cast<float4*>(s_strip);
        for (int idx = threadIdx.x; idx < keep * 8; idx += NTH) dst[idx] = __ldg(src + idx);
    }
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rp = lane / TW, cl = lane % TW;
    unsigned epoch = 0;
    unsigned* ctr = p.counters;

    for (int64_t b0 = 0; b0 < p.B; b0 += R) {
        const int nb = (int)min((int64_t)R, p.B - b0);
        // ---- delta_{-1} = init -> vec[0]
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
            for (int idx = tid; idx < nb * TW; idx += NTH) {
                const int r = idx / TW, j = tile * TW + (idx % TW);
                if (j < S) p.vec[(b0 + r) * S + j] = p.init ? p.init[(b0 + r) * p.init_sb + j] : 0.f;
            }
        group_barrier(ctr, (++epoch) * gridDim.x, gridDim.x);

        for (int64_t t = 0; t < p.T; ++t) {
            const float* vin = p.vec + (size_t)(t & 1) * p.B * S;
            float* vout = p.vec + (size_t)((t + 1) & 1) * p.B * S;
            // R independent loads per iteration and no division (the flat loop with idx / SR cost ~17 us per step at R = 16, S = 1024)
            for (int i = tid; i < SR; i += NTH) {
#pragma unroll
                for (int r = 0; r < R; ++r) s_delta[r * SR + i] = (r < nb && i < S) ? __ldcg(vin + (b0 + r) * S + i) : kNegInf;
            }
            __syncthreads();
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int col = tile * TW + cl;
                const bool colok = col < S;
                float ob[R], best[R];
                int arg[R];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    ob[r] = (p.obs && colok && r < nb) ? __ldg(p.obs + ((b0 + r) * p.T + t) * S + col) : 0.f;
                    best[r] = kNegInf;
                    arg[r] = warp * RP + rp < S ? warp * RP + rp : 0;
                }
                const float* Acol = p.A + (size_t)tile * SR * TW + lane;  // strip-major copy of A: slot sl at + sl * 32
                int sl = warp;
                for (; sl < keep; sl += NW * 8) {  // resident slots (keep is a multiple of 8*NW)
                    float a[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) a[u] = s_strip[(sl + NW * u) * 32 + lane];
#pragma unroll
                    for (int u = 0; u < 8; ++u)
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const int i = (sl + NW * u) * RP + rp;
                            const float v = s_delta[r * SR + i] + (a[u] + ob[r]);
                            if (v > best[r]) {
                                best[r] = v;
                                arg[r] = i;
                            }
                        }
                }
                for (; sl + NW * 15 < NS; sl += NW * 16) {
                    float a[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) a[u] = __ldg(Acol + (size_t)(sl + NW * u) * 32);
#pragma unroll
                    for (int u = 0; u < 16; ++u)
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const int i = (sl + NW * u) * RP + rp;
                            const float v = s_delta[r * SR + i] + (a[u] + ob[r]);  // (A + obs) first: hmm.py:130
                            if (v > best[r]) {
                                best[r] = v;
                                arg[r] = i;
                            }
                        }
                }
                for (; sl < NS; sl += NW) {
                    const float a = __ldg(Acol + (size_t)sl * 32);
                    const int i = sl * RP + rp;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const float v = s_delta[r * SR + i] + (a + ob[r]);
                        if (v > best[r]) {
                            best[r] = v;
                            arg[r] = i;
                        }
                    }
                }
                // combine the RP row phases of a column (lanes cl, cl + TW, ...): larger value, on ties the smaller row
#pragma unroll
                for (int o = TW; o < 32; o <<= 1) {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const float ov = __shfl_xor_sync(0xffffffffu, best[r], o);
                        const int oa = __shfl_xor_sync(0xffffffffu, arg[r], o);
                        if (ov > best[r] || (ov == best[r] && oa < arg[r])) {
                            best[r] = ov;
                            arg[r] = oa;
                        }
                    }
                }
                if (rp == 0) {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        s_best[(warp * R + r) * TW + cl] = best[r];
                        s_arg[(warp * R + r) * TW + cl] = arg[r];
                    }
                }
                __syncthreads();
                for (int fin = tid; fin < R * TW; fin += NTH) {
                    const int r = fin / TW, l = fin % TW, c = tile * TW + l;
                    if (r < nb && c < S) {
                        float bb = kNegInf;
                        int aa = 0x7fffffff;
                        for (int w = 0; w < NW; ++w) {
                            const float v = s_best[(w * R + r) * TW + l];
                            const int ai = s_arg[(w * R + r) * TW + l];
                            if (v > bb || (v == bb && ai < aa)) {
                                bb = v;
                                aa = ai;
                            }
                        }
                        vout[(b0 + r) * S + c] = bb;
                        if (ARG) p.backptr[((b0 + r) * p.T + t) * S + c] = aa;
                        if (p.alpha_all) p.alpha_all[((b0 + r) * p.T + t) * S + c] = bb;
                    }
                }
                __syncthreads();
            }
            group_barrier(ctr, (++epoch) * gridDim.x, gridDim.x);
        }
        // ---- out[b] = max_j delta_{T-1}[j], first-index argmax
        if (blockIdx.x == 0) {
            const float* vfin = p.vec + (size_t)(p.T & 1) * p.B * S;
            for (int r = warp; r < nb; r += NW) {
                float m = kNegInf;
                int arg = 0x7fffffff;
                for (int j = lane; j < S; j += 32) {
                    const float v = __ldcg(vfin + (b0 + r) * S + j);
                    if (v > m || (v == m && j < arg)) {
                        m = v;
                        arg = j;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float om = __shfl_xor_sync(0xffffffffu, m, o);
                    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
                    if (om > m || (om == m && oa < arg)) {
                        m = om;
                        arg = oa;
                    }
                }
                if (lane == 0) {
                    p.out[b0 + r] = m;
                    if (p.last) p.last[b0 + r] = arg;
                }
            }
        }
        group_barrier(ctr, (++epoch) * gridDim.x, gridDim.x);  // vec[] is reused by the next chains
    }
}

// ---------------------------------------------------------------------------------------------
// backward recurrence, generic path (log / max):
//   beta_{T-1} = final (or one);  beta_{t-1}[i] = (+)_j A_t[i,j] (x) obs[t,j] (x) beta_t[j]
// CTA c owns rows i in [c*W, c*W+W); one warp per row, lanes stride over j (coalesced rows of A).
// beta_all[b,t,:] = beta_t.  smem: s_w[R][S] (= obs_t + beta_t)
template <int SEMI, int R>
__global__ void __launch_bounds__(256) hmm_backward_generic(HmmArgs p) {
    extern __shared__ __align__(16) float smem[];
    const int S = p.S, W = p.W;
    float* s_w = smem;
    __shared__ double s_ell[16];  // LOG: beta_t = vec_t + ell_t (same scheme as the forward kernel)
    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    const int gconc = blockIdx.x / p.Gg, slice = blockIdx.x % p.Gg;
    const int i0 = slice * W;
    const int iw = max(0, min(W, S - i0));
    unsigned* ctr = p.counters + gconc;
    unsigned epoch = 0;

    for (int g = gconc; g < p.groups_total; g += p.groups_conc) {
        const int64_t b0 = (int64_t)g * R;
        const int nb = (int)min((int64_t)R, p.B - b0);
        // buffer index convention: vec[t&1] holds beta_t
        if (tid < 16) s_ell[tid] = 0.0;
        for (int idx = tid; idx < nb * iw; idx += 256) {
            int r = idx / iw, i = i0 + idx % iw;
            float v = p.final_ ? p.final_[(b0 + r) * S + i] : 0.f;
            p.vec[(size_t)((p.T - 1) & 1) * p.B * S + (b0 + r) * S + i] = v;
            if (p.alpha_all) p.alpha_all[((b0 + r) * p.T + p.T - 1) * S + i] = v;
            if (p.ell_all && i == 0) p.ell_all[(b0 + r) * p.T + p.T - 1] = 0.0;
        }
        group_barrier(ctr, (++epoch) * p.Gg, p.Gg);
        for (int64_t t = p.T - 1; t >= 1; --t) {
            const float* vin = p.vec + (size_t)(t & 1) * p.B * S;
            float* vout = p.vec + (size_t)((t - 1) & 1) * p.B * S;
            for (int idx = tid; idx < R * S; idx += 256) {
                int r = idx / S, j = idx % S;
                float v = kNegInf;
                if (r < nb) v = __ldcg(vin + (b0 + r) * S + j) + (p.obs ? p.obs[((b0 + r) * p.T + t) * S + j] : 0.f);
                s_w[idx] = v;
            }
            __syncthreads();
            if (SEMI == FB2_SEMIRING_LOG) {
                for (int r = warp; r < R; r += 8) {
                    float m = -FLT_MAX;
                    for (int j = lane; j < S; j += 32) m = fmaxf(m, s_w[r * S + j]);
                    m = warp_max(m);
                    for (int j = lane; j < S; j += 32) s_w[r * S + j] -= m;
                    if (lane == 0 && r < nb) s_ell[r] += (double)m;
                }
                __syncthreads();
            }
            for (int ii = warp; ii < iw; ii += 8) {
                const int i = i0 + ii;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (r >= nb) break;
                    const float* Ar = p.A + (b0 + r) * p.a_sb + t * p.a_st + (int64_t)i * S;
                    float m = (SEMI == FB2_SEMIRING_LOG) ? -FLT_MAX : kNegInf;
                    for (int j = lane; j < S; j += 32) m = fmaxf(m, __ldg(Ar + j) + s_w[r * S + j]);
                    m = warp_max(m);
                    float res = m;
                    if (SEMI == FB2_SEMIRING_LOG) {
                        float s = 0.f;
                        for (int j = lane; j < S; j += 32) s += __expf(__ldg(Ar + j) + s_w[r * S + j] - m);
                        s = warp_sum(s);
                        res = m + logf(s);
                    }
                    if (lane == 0) {
                        vout[(b0 + r) * S + i] = res;
                        if (p.alpha_all) {
                            float off = (SEMI == FB2_SEMIRING_LOG && !p.ell_all) ? (float)s_ell[r] : 0.f;
                            p.alpha_all[((b0 + r) * p.T + t - 1) * S + i] = res + off;
                        }
                        if (SEMI == FB2_SEMIRING_LOG && p.ell_all && i == 0) p.ell_all[(b0 + r) * p.T + t - 1] = s_ell[r];
                    }
                }
            }
            group_barrier(ctr, (++epoch) * p.Gg, p.Gg);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Viterbi backtrace: path[b,T] = last[b]; path[b,t] = backptr[b,t,path[b,t+1]]
__global__ void viterbi_backtrace(const int32_t* backptr, const int32_t* last, int64_t B, int64_t T, int S,
                                  int64_t* path) {
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    int s = last[b];
    path[b * (T + 1) + T] = s;
    for (int64_t t = T - 1; t >= 0; --t) {
        s = backptr[(b * T + t) * S + s];
        path[b * (T + 1) + t] = s;
    }
}

// gamma[b,t,j] = alpha_hat[b,t,j] + beta_hat[b,t,j] + (ell_a[b,t] + ell_b[b,t] - logZ[b])   (in place over alpha)
__global__ void marginals_combine(float* alpha, const float* beta, const double* ell_a, const double* ell_b,
                                  const double* logZ, int64_t B, int64_t T, int S) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * T * (int64_t)S) return;
    int64_t bt = idx / S;
    alpha[idx] = alpha[idx] + beta[idx] + (float)(ell_a[bt] + ell_b[bt] - logZ[bt / T]);
}

// xi_sum[b,i,j] = sum_t exp(alpha_{t-1}[i] + A_t[i,j] + obs[t,j] + beta_t[j] - logZ)
// one thread per (b,i,j); alpha_{-1} = init.  Simple O(T S^2) accumulation (T-serial per thread).
__global__ void xi_sum_kernel(const float* init, int64_t init_sb, const float* A, int64_t a_sb, int64_t a_st,
                              const float* obs, const float* alpha, const float* beta, const double* ell_a,
                              const double* ell_b, const double* logZ, int64_t B, int64_t T, int S, float* xi) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t per = (int64_t)S * S;
    if (idx >= B * per) return;
    int64_t b = idx / per;
    int i = (int)((idx % per) / S), j = (int)(idx % S);
    double z = logZ[b];
    float acc = 0.f;
    for (int64_t t = 0; t < T; ++t) {
        float ap = t == 0 ? (init ? init[b * init_sb + i] : 0.f) : alpha[(b * T + t - 1) * S + i];
        float off = (float)((t == 0 ? 0.0 : ell_a[b * T + t - 1]) + ell_b[b * T + t] - z);
        float v = ap + A[b * a_sb + t * a_st + (int64_t)i * S + j] + obs[(b * T + t) * S + j] + beta[(b * T + t) * S + j] + off;
        acc += __expf(v);
    }
    xi[idx] = acc;
}

// adj[b,t,i,j] = alphaprev[b,t,i] + beta[b,t,j]  where alphaprev[b,0] = init (or 0) and alphaprev[b,t] = alpha[b,t-1];
// ell_a / ell_b (nullable) are the offsets of normalised messages (LOG).
__global__ void adjoint_outer(const float* init, const float* alpha, const float* beta, const double* ell_a,
                              const double* ell_b, int64_t B, int64_t T, int S, float* adj) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t per = (int64_t)S * S;
    if (idx >= B * T * per) return;
    int64_t bt = idx / per;
    int64_t b = bt / T, t = bt % T;
    int i = (int)((idx % per) / S), j = (int)(idx % S);
    float ap = t == 0 ? (init ? init[b * S + i] : 0.f) : alpha[(b * T + t - 1) * S + i];
    float off = 0.f;
    if (ell_a) off = (float)((t == 0 ? 0.0 : ell_a[bt - 1]) + ell_b[bt]);
    adj[idx] = (ap + beta[(b * T + t) * S + j]) + off;
}

// ---------------------------------------------------------------------------------------------
struct PassPlan {
    int R, Gg, W, LW, groups_total, groups_conc, grid;
    size_t smem;
};

static int g_sm_count = 0;
static size_t g_smem_optin = 0;

static int device_facts() {
    if (g_sm_count) return FB2_OK;
    int dev = 0;
    FB2_CUDA(cudaGetDevice(&dev));
    int sm = 0, optin = 0;
    FB2_CUDA(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev));
    FB2_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    g_sm_count = sm;
    g_smem_optin = (size_t)optin;
    return FB2_OK;
}

static int make_plan(int64_t B, int S, bool backward, PassPlan* pl) {
    int st = device_facts();
    if (st != FB2_OK) return st;
    int R = B >= 16 ? 16 : (B > 1 ? 4 : 1);
    if (backward && R > 4) R = 4;
    // shared memory needed independent of W: alpha staging
    for (;;) {
        size_t fixed = (size_t)R * S * 4 + (backward ? 0 : 2 * 256 * (size_t)R * 4);
        if (fixed + 4096 <= g_smem_optin || R == 1) break;
        R = R == 16 ? 4 : 1;
    }
    int groups_total = (int)ceil_div(B, R);
    int groups_conc = groups_total < g_sm_count ? groups_total : g_sm_count;
    int Gg = g_sm_count / groups_conc;
    if (Gg > S) Gg = S;
    if (Gg < 1) Gg = 1;
    int W = (int)ceil_div(S, Gg);
    if (W > 4) W = (W + 7) / 8 * 8;  // whole 32-byte sectors per row segment
    Gg = (int)ceil_div(S, W);        // drop CTAs that would own nothing
    int LW = 1;
    while (LW < W && LW < 32) LW <<= 1;
    pl->R = R;
    pl->Gg = Gg;
    pl->W = W;
    pl->LW = LW;
    pl->groups_total = groups_total;
    pl->groups_conc = groups_conc;
    pl->grid = groups_conc * Gg;
    pl->smem = (size_t)R * S * 4 + (backward ? 0 : (2 * 256 * (size_t)R * 4 + 2 * (size_t)R * W * 4));
    if (pl->smem > g_smem_optin) {
        set_error("hmm pass: state dimension S=%d needs %zu bytes of shared memory (> %zu)", S, pl->smem, g_smem_optin);
        return FB2_ERR_UNSUPPORTED;
    }
    return FB2_OK;
}

template <typename K>
static int coop_launch(K kernel, const PassPlan& pl, HmmArgs& args, cudaStream_t stream, int threads = 256) {
    FB2_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    int per_sm = 0;
    FB2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, pl.smem));
    if (per_sm * g_sm_count < pl.grid) {
        set_error("hmm pass: grid of %d CTAs cannot be co-resident (%d per SM)", pl.grid, per_sm);
        return FB2_ERR_UNSUPPORTED;
    }
    void* kargs[] = {&args};
    FB2_CUDA(cudaLaunchCooperativeKernel((void*)kernel, dim3(pl.grid), dim3(threads), kargs, pl.smem, stream));
    count_launch();
    return FB2_OK;
}

static int run_forward(int semiring, bool want_arg, const PassPlan& pl, HmmArgs& a, cudaStream_t stream) {
#define FB2_FWD(SEMI, RR, ARGF) return coop_launch(hmm_forward_generic<SEMI, RR, ARGF>, pl, a, stream)
    if (semiring == FB2_SEMIRING_LOG) {
        if (pl.R == 16) FB2_FWD(FB2_SEMIRING_LOG, 16, false);
        if (pl.R == 4) FB2_FWD(FB2_SEMIRING_LOG, 4, false);
        FB2_FWD(FB2_SEMIRING_LOG, 1, false);
    } else if (want_arg) {
        if (pl.R == 16) FB2_FWD(FB2_SEMIRING_MAX, 16, true);
        if (pl.R == 4) FB2_FWD(FB2_SEMIRING_MAX, 4, true);
        FB2_FWD(FB2_SEMIRING_MAX, 1, true);
    } else {
        if (pl.R == 16) FB2_FWD(FB2_SEMIRING_MAX, 16, false);
        if (pl.R == 4) FB2_FWD(FB2_SEMIRING_MAX, 4, false);
        FB2_FWD(FB2_SEMIRING_MAX, 1, false);
    }
#undef FB2_FWD
}

static int run_backward(int semiring, const PassPlan& pl, HmmArgs& a, cudaStream_t stream) {
    if (semiring == FB2_SEMIRING_LOG) {
        if (pl.R == 4) return coop_launch(hmm_backward_generic<FB2_SEMIRING_LOG, 4>, pl, a, stream);
        return coop_launch(hmm_backward_generic<FB2_SEMIRING_LOG, 1>, pl, a, stream);
    }
    if (pl.R == 4) return coop_launch(hmm_backward_generic<FB2_SEMIRING_MAX, 4>, pl, a, stream);
    return coop_launch(hmm_backward_generic<FB2_SEMIRING_MAX, 1>, pl, a, stream);
}

// ---------------------------------------------------------------------------------------------
// Log-semiring forward, streaming path: the same shape class as hmm_maxplus_stream (shared A, S >= 256, few chains) -- in
// particular S > 1024, where the dataflow kernels do not apply and the generic kernel needs S^2 exp per chain-step (222 us per
// step at S = 4096).  Exp trick: P = exp(A - colmax A) is formed ONCE, strip-major, and a step is a GEMV in the linear domain
//     w = v_{t-1} / max v_{t-1},   u_j = (sum_i w_i P_ij) * exp(obs_j + cA_j - m_t),   m_t = max_j (obs_j + cA_j)
// with the log of every removed scale accumulated in double: L_t = L_{t-1} + log max v_{t-1} + m_t.  Both maxima are computed
// redundantly (and identically: max is order-independent) by every CTA, so no extra exchange is needed; one grid barrier per
// step.  Dynamic range per step is that of fp32 (entries ~87 nats below the step maximum flush to zero).
__global__ void stream_retile_P(const float* A, const float* cA, int S, float* strips, int transpose) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int ntiles = (S + 31) / 32;
    if (idx >= (int64_t)ntiles * S * 32) return;
    const int l = (int)(idx & 31);
    const int64_t r = idx >> 5;
    const int row = (int)(r % S), tile = (int)(r / S);
    const int col = tile * 32 + l;
    strips[idx] = col < S ? expf((transpose ? A[(int64_t)col * S + row] : A[(int64_t)row * S + col]) - cA[col]) : 0.f;
}
__global__ void stream_colmax(const float* A, int S, float* cA, int transpose) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= S) return;
    float m = -FLT_MAX;
    for (int i = 0; i < S; ++i) m = fmaxf(m, transpose ? A[(int64_t)j * S + i] : A[(int64_t)i * S + j]);
    cA[j] = m;
}

template <int R>
__global__ void __launch_bounds__(STREAM_THREADS) hmm_logsum_stream(HmmArgs p, const float* cA) {
    constexpr int NW = STREAM_WARPS, NTH = STREAM_THREADS;
    extern __shared__ __align__(16) float smem[];
    const int S = p.S;
    float* s_w = smem;                                        // [R][S] scaled previous message
    float* s_part = s_w + (((size_t)R * S + 3) & ~(size_t)3);  // [NW][R][32]
    float* s_strip = s_part + NW * R * 32;                     // [keep][32]
    __shared__ float s_red[2][R][NW];
    __shared__ double s_L[R];
    __shared__ float s_vm[R], s_mt[R];
    const int keep = p.W;
    if (keep > 0 && (int)blockIdx.x < (S + 31) / 32) {
        const float4* src = reinterpret_cast<const float4*>(p.A + (size_t)blockIdx.x * S * 32);
        float4* dst = reinterpret_cast<float4*>(s_strip);
        for (int idx = threadIdx.x; idx < keep * 8; idx += NTH) dst[idx] = __ldg(src + idx);
    }
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ntiles = (S + 31) / 32;
    const int64_t OT = p.obs_T ? p.obs_T : p.T;  // rows per chain of obs / alpha_all / ell_all
    unsigned epoch = 0;
    unsigned* ctr = p.counters;

    for (int64_t b0 = 0; b0 < p.B; b0 += R) {
        const int nb = (int)min((int64_t)R, p.B - b0);
        // ---- v_{-1} = exp(init - max init), L_{-1} = max init (every CTA computes the maxima itself)
        for (int r = 0; r < R; ++r) {
            float m = -INFINITY;
            if (r < nb)
                for (int j = tid; j < S; j += NTH) m = fmaxf(m, p.init ? p.init[(b0 + r) * p.init_sb + j] : 0.f);
            m = warp_max(m);
            if (lane == 0) s_red[0][r][warp] = m;
        }
        __syncthreads();
        if (tid < R) {
            float m = -INFINITY;
            for (int w = 0; w < NW; ++w) m = fmaxf(m, s_red[0][tid][w]);
            s_L[tid] = (double)m;  // -inf for a dead chain, NaN/+inf propagate
            s_vm[tid] = m;
        }
        __syncthreads();
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
            for (int idx = tid; idx < nb * 32; idx += NTH) {
                const int r = idx >> 5, j = tile * 32 + (idx & 31);
                if (j < S) {
                    const float x = p.init ? p.init[(b0 + r) * p.init_sb + j] : 0.f;
                    p.vec[(b0 + r) * S + j] = (s_vm[r] == -INFINITY) ? 0.f : __expf(x - s_vm[r]);
                }
            }
        group_barrier(ctr, (++epoch) * gridDim.x, gridDim.x);

        for (int64_t t = 0; t < p.T; ++t) {
            const int64_t row = p.reverse ? (p.T - 1 - t) : t;
            const float* vin = p.vec + (size_t)(t & 1) * p.B * S;
            float* vout = p.vec + (size_t)((t + 1) & 1) * p.B * S;
            // ---- stage v_{t-1}, its maximum, and m_t = max_j (obs_j + cA_j)
            for (int r = 0; r < R; ++r) {
                float vm = 0.f, mt = -INFINITY;
                if (r < nb)
                    for (int i = tid; i < S; i += NTH) {
                        const float v = __ldcg(vin + (b0 + r) * S + i);
                        s_w[r * S + i] = v;
                        vm = fmaxf(vm, v);
                        mt = fmaxf(mt, (p.obs ? __ldg(p.obs + ((b0 + r) * OT + row) * S + i) : 0.f) + __ldg(cA + i));
                    }
                else
                    for (int i = tid; i < S; i += NTH) s_w[r * S + i] = 0.f;
                vm = warp_max(vm);
                mt = warp_max(mt);
                if (lane == 0) {
                    s_red[0][r][warp] = vm;
                    s_red[1][r][warp] = mt;
                }
            }
            __syncthreads();
            if (tid < R) {
                float vm = 0.f, mt = -INFINITY;
                for (int w = 0; w < NW; ++w) {
                    vm = fmaxf(vm, s_red[0][tid][w]);
                    mt = fmaxf(mt, s_red[1][tid][w]);
                }
                s_vm[tid] = vm;
                s_mt[tid] = mt;
                if (tid < nb) {
                    s_L[tid] += (vm > 0.f ? (double)logf(vm) : -INFINITY) + (double)mt;
                    if (p.ell_all && blockIdx.x == 0) p.ell_all[(b0 + tid) * OT + row] = s_L[tid];
                }
            }
            __syncthreads();
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                float acc[R];
#pragma unroll
                for (int r = 0; r < R; ++r) acc[r] = 0.f;
                const float* Pcol = p.A + (size_t)tile * S * 32 + lane;  // strip-major P: [tile][row][32]
                int i = warp;
                for (; i < keep; i += NW * 8) {
                    float a[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) a[u] = s_strip[(i + NW * u) * 32 + lane];
#pragma unroll
                    for (int u = 0; u < 8; ++u)
#pragma unroll
                        for (int r = 0; r < R; ++r) acc[r] = fmaf(s_w[r * S + i + NW * u], a[u], acc[r]);
                }
                for (; i + NW * 15 < S; i += NW * 16) {
                    float a[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) a[u] = __ldg(Pcol + (size_t)(i + NW * u) * 32);
#pragma unroll
                    for (int u = 0; u < 16; ++u)
#pragma unroll
                        for (int r = 0; r < R; ++r) acc[r] = fmaf(s_w[r * S + i + NW * u], a[u], acc[r]);
                }
                for (; i < S; i += NW) {
                    const float a = __ldg(Pcol + (size_t)i * 32);
#pragma unroll
                    for (int r = 0; r < R; ++r) acc[r] = fmaf(s_w[r * S + i], a, acc[r]);
                }
#pragma unroll
                for (int r = 0; r < R; ++r) s_part[(warp * R + r) * 32 + lane] = acc[r];
                __syncthreads();
                for (int fin = tid; fin < R * 32; fin += NTH) {  // R * 32 may exceed the thread count (R = 16)
                    const int r = fin >> 5, l = fin & 31, c = tile * 32 + l;
                    if (r < nb && c < S) {
                        float sum = 0.f;
                        for (int w = 0; w < NW; ++w) sum += s_part[(w * R + r) * 32 + l];
                        const float vm = s_vm[r], mt = s_mt[r];
                        const float ob = (p.obs ? __ldg(p.obs + ((b0 + r) * OT + row) * S + c) : 0.f) + __ldg(cA + c);
                        float u = 0.f;
                        if (vm > 0.f && mt > -INFINITY) u = (sum / vm) * __expf(ob - mt);
                        if (!(mt < INFINITY) || !(vm < INFINITY)) u = __int_as_float(0x7fc00000);  // inf / NaN input poisons the chain
                        vout[(b0 + r) * S + c] = u;
                        if (p.alpha_all) {
                            float* dst = p.alpha_all + ((b0 + r) * OT + row) * S + c;
                            if (p.reverse)  // beta_row = (A exp-contracted with the message) without the observation of this row
                                *dst = (vm > 0.f ? __logf(sum / vm) : -INFINITY) + __ldg(cA + c) - mt;
                            else if (p.ell_all)
                                *dst = __logf(u);  // normalised: the offset goes to ell_all
                            else
                                *dst = (float)(s_L[r] + (double)__logf(u));
                        }
                    }
                }
                __syncthreads();
            }
            group_barrier(ctr, (++epoch) * gridDim.x, gridDim.x);
        }
        // ---- out[b] = L_{T-1} + log sum_j v_{T-1}[j]
        if (blockIdx.x == 0) {
            const float* vfin = p.vec + (size_t)(p.T & 1) * p.B * S;
            for (int r = warp; r < nb; r += NW) {
                double sum = 0.0;
                for (int j = lane; j < S; j += 32) sum += (double)__ldcg(vfin + (b0 + r) * S + j);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                if (lane == 0) {
                    const double z = s_L[r] + log(sum);
                    p.out[b0 + r] = (float)z;
                    if (p.out_d) p.out_d[b0 + r] = z;
                }
            }
        }
        group_barrier(ctr, (++epoch) * gridDim.x, gridDim.x);  // vec[] and s_L are reused by the next chains
    }
}

static int check_hmm_args(int semiring, const float* A, int64_t B, int64_t T, int S) {
    FB2_REQUIRE(semiring == FB2_SEMIRING_LOG || semiring == FB2_SEMIRING_MAX, "unknown semiring %d", semiring);
    FB2_REQUIRE(B >= 0 && T >= 1 && S >= 1, "bad extents B=%lld T=%lld S=%d", (long long)B, (long long)T, S);
    FB2_REQUIRE(A != nullptr, "A is null");
    return FB2_OK;
}

// workspace layout helpers ----------------------------------------------------------------------
static size_t vec_bytes(int64_t B, int S) { return align_up((size_t)2 * B * S * 4); }
static size_t ctr_bytes() { return align_up(1024 * sizeof(unsigned)); }

// beta_{T-1} = 1 (log 0) with offset 0
__global__ void stream_fill_last_row(float* beta_all, double* ell_all, int64_t B, int64_t T, int S) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * S) return;
    const int64_t b = idx / S;
    beta_all[(b * T + T - 1) * S + idx % S] = 0.f;
    if (idx % S == 0) ell_all[b * T + T - 1] = 0.0;
}

// Runs hmm_logsum_stream if the shape qualifies and the workspace holds its buffers; FB2_ERR_UNSUPPORTED otherwise.
// reverse = 1: the backward recursion as a forward pass on A^T over rows T_steps-1 .. 0 (beta_hat into alpha_all, offsets into ell_all).
static int log_stream_pass(const float* init, int64_t init_sb, const float* A, const float* obs, int64_t obs_T, int64_t B, int64_t T_steps, int S,
                           float* out, double* out_d, float* alpha_all, double* ell_all, int reverse, void* workspace, size_t workspace_bytes,
                           cudaStream_t stream) {
    if (S < 256 || B > 64 || stream_disabled() || !g_sm_count) return FB2_ERR_UNSUPPORTED;
    Arena arena(workspace, workspace_bytes);
    float* vec = arena.take<float>((size_t)2 * B * S);
    unsigned* ctr = arena.take<unsigned>(1024);
    const size_t strip_elems = (size_t)((S + 31) / 32) * S * 32;
    float* strips = arena.take<float>(strip_elems);
    float* cA = arena.take<float>((size_t)S);
    float* out_scratch = arena.take<float>((size_t)B);
    // chains per pass: the strips are re-read for every pass, so as many chains as shared memory (R * S staged entries) allows
    int R = B >= 16 ? 16 : (B >= 4 ? 4 : 1);
    auto stream_smem = [&](int r) { return ((((size_t)r * S + 3) & ~(size_t)3) + STREAM_WARPS * r * 32) * 4; };
    if (R == 16 && stream_smem(16) > 96 * 1024) R = 4;
    size_t smem = stream_smem(R);
    if (!arena.ok() || smem > g_smem_optin) return FB2_ERR_UNSUPPORTED;
    FB2_CUDA(cudaMemsetAsync(ctr, 0, 1024 * sizeof(unsigned), stream));
    stream_colmax<<<(unsigned)ceil_div(S, 128), 128, 0, stream>>>(A, S, cA, reverse);
    FB2_LAUNCH_CHECK();
    stream_retile_P<<<(unsigned)ceil_div((int64_t)strip_elems, 256), 256, 0, stream>>>(A, cA, S, strips, reverse);
    FB2_LAUNCH_CHECK();
    HmmArgs a{};
    a.init = init; a.init_sb = init_sb; a.A = strips; a.obs = obs; a.B = B; a.T = T_steps; a.S = S; a.vec = vec; a.counters = ctr;
    a.alpha_all = alpha_all; a.ell_all = ell_all; a.out = out ? out : out_scratch; a.out_d = out_d; a.obs_T = obs_T; a.reverse = reverse;
    const int ntiles = (S + 31) / 32;
    const int grid = ntiles < g_sm_count ? ntiles : g_sm_count;
    int keep = 0;
    if (ntiles <= g_sm_count) {
        keep = (int)((g_smem_optin - smem - 2048) / 128);
        if (keep > S) keep = S;
        keep = keep / (8 * STREAM_WARPS) * (8 * STREAM_WARPS);
        if (keep < 0) keep = 0;
    }
    a.W = keep;
    smem += (size_t)keep * 128;
    const float* cAc = cA;
    void* args[] = {&a, &cAc};
    const void* fn = R == 16 ? reinterpret_cast<const void*>(hmm_logsum_stream<16>)
                             : (R == 4 ? reinterpret_cast<const void*>(hmm_logsum_stream<4>) : reinterpret_cast<const void*>(hmm_logsum_stream<1>));
    FB2_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3((unsigned)grid), dim3(STREAM_THREADS), args, smem, stream);
    if (e != cudaSuccess) {
        set_error("hmm_logsum_stream launch failed: %s", cudaGetErrorString(e));
        return FB2_ERR_CUDA;
    }
    count_launch();
    return FB2_OK;
}

int hmm_forward_device(int semiring, const float* init, int64_t init_sb, const float* A, int64_t a_sb, int64_t a_st,
                       const float* obs, const float* final_, int64_t B, int64_t T, int S, float* out,
                       float* alpha_all, int32_t* backptr, int32_t* last, void* workspace, size_t workspace_bytes,
                       cudaStream_t stream, int cpo = 1, float* alpha_last = nullptr, double* ell_all = nullptr,
                       double* out_d = nullptr, double* ell_last = nullptr) {
    PassPlan pl;
    int st = make_plan(B, S, false, &pl);
    const bool stream_ok = semiring == FB2_SEMIRING_MAX && a_sb == 0 && a_st == 0 && cpo == 1 && !final_ && !alpha_last &&
                           !ell_all && S > 64 && (B <= 16 || (S >= 512 && B <= 64)) && !stream_disabled();
    // (S <= 64: warp-per-chain kernel.  The streaming kernel runs on S/32 SMs only; measured, the generic kernel is faster per chain
    //  once there is more than one 16-chain pass at S < 512: e.g. S=256 B=64: 19 vs 13.5 us/step, scripts/cliff_time.py)
    if (semiring == FB2_SEMIRING_LOG && a_sb == 0 && a_st == 0 && cpo == 1 && !final_ && !alpha_last && !ell_last && !backptr && !last) {
        const int ls = log_stream_pass(init, init_sb, A, obs, T, B, T, S, out, out_d, alpha_all, ell_all, 0, workspace, workspace_bytes, stream);
        if (ls != FB2_ERR_UNSUPPORTED) return ls;
    }
    if (st != FB2_OK && !(stream_ok && st == FB2_ERR_UNSUPPORTED)) return st;
    Arena arena(workspace, workspace_bytes);
    float* vec = arena.take<float>((size_t)2 * B * S);
    unsigned* ctr = arena.take<unsigned>(1024);
    if (!arena.ok()) {
        set_error("hmm forward: workspace too small (%zu bytes)", workspace_bytes);
        return FB2_ERR_WORKSPACE;
    }
    FB2_CUDA(cudaMemsetAsync(ctr, 0, 1024 * sizeof(unsigned), stream));
    if (stream_ok) {
        // row phases: as many as keep the number of tiles (= CTAs) within the SM count
        int RP = 1;
        while (RP < 8 && (S + (32 / (2 * RP)) - 1) / (32 / (2 * RP)) <= g_sm_count) RP *= 2;
        const int TW = 32 / RP, SR = (S + RP - 1) / RP * RP;
        const int ntiles = (S + TW - 1) / TW;
        const size_t strip_elems = (size_t)ntiles * SR * TW;
        float* strips = arena.take<float>(strip_elems);
        if (!arena.ok()) {
            set_error("hmm forward (max-plus stream): workspace too small (%zu bytes)", workspace_bytes);
            return FB2_ERR_WORKSPACE;
        }
        stream_retile_A<<<(unsigned)ceil_div((int64_t)strip_elems, 256), 256, 0, stream>>>(A, S, strips, RP);
        FB2_LAUNCH_CHECK();
        HmmArgs a{};
        a.init = init; a.init_sb = init_sb; a.A = strips; a.obs = obs; a.B = B; a.T = T; a.S = S; a.vec = vec; a.counters = ctr;
        a.alpha_all = alpha_all; a.backptr = backptr; a.out = out; a.last = last;
        int R = B >= 16 ? 16 : (B >= 4 ? 4 : 1);  // chains per pass (A is re-read for every pass)
        auto mp_smem = [&](int r) { return ((((size_t)r * SR + 3) & ~(size_t)3) + 2 * STREAM_WARPS * r * TW) * 4; };
        if (R == 16 && mp_smem(16) > 96 * 1024) R = 4;
        size_t smem = mp_smem(R);
        if (smem <= g_smem_optin) {
            int grid = ntiles < g_sm_count ? ntiles : g_sm_count;
            int keep = 0;
            const int NS = SR / RP;
            if (ntiles <= g_sm_count) {  // one strip per CTA: keep as many of its row slots on chip as fit
                keep = (int)((g_smem_optin - smem - 1024) / 128);
                if (keep > NS) keep = NS;
                keep = keep / (8 * STREAM_WARPS) * (8 * STREAM_WARPS);
                if (keep < 0) keep = 0;
            }
            a.W = keep;
            smem += (size_t)keep * 128;
            PassPlan sp{};
            sp.grid = grid;
            sp.smem = smem;
#define FB2_MPS(Rv, RPv)                                                                                           \
    return backptr ? coop_launch(hmm_maxplus_stream<Rv, true, RPv>, sp, a, stream, STREAM_THREADS)                  \
                   : coop_launch(hmm_maxplus_stream<Rv, false, RPv>, sp, a, stream, STREAM_THREADS)
#define FB2_MPS_R(RPv)                 \
    do {                               \
        if (R == 16) { FB2_MPS(16, RPv); } \
        if (R == 4) { FB2_MPS(4, RPv); }   \
        FB2_MPS(1, RPv);               \
    } while (0)
            if (RP == 8) FB2_MPS_R(8);
            if (RP == 4) FB2_MPS_R(4);
            if (RP == 2) FB2_MPS_R(2);
            if (R == 16) return backptr ? coop_launch(hmm_maxplus_stream32<16, true>, sp, a, stream, STREAM_THREADS)
                                        : coop_launch(hmm_maxplus_stream32<16, false>, sp, a, stream, STREAM_THREADS);
            if (R == 4) return backptr ? coop_launch(hmm_maxplus_stream32<4, true>, sp, a, stream, STREAM_THREADS)
                                       : coop_launch(hmm_maxplus_stream32<4, false>, sp, a, stream, STREAM_THREADS);
            return backptr ? coop_launch(hmm_maxplus_stream32<1, true>, sp, a, stream, STREAM_THREADS)
                           : coop_launch(hmm_maxplus_stream32<1, false>, sp, a, stream, STREAM_THREADS);
#undef FB2_MPS_R
#undef FB2_MPS
        }
        if (st != FB2_OK) return st;
    }
    HmmArgs a{};
    a.init = init; a.init_sb = init_sb; a.A = A; a.a_sb = a_sb; a.a_st = a_st; a.obs = obs;
    a.B = B; a.T = T; a.S = S; a.vec = vec; a.counters = ctr; a.alpha_all = alpha_all; a.backptr = backptr;
    a.out = out; a.last = last; a.final_ = final_;
    a.Gg = pl.Gg; a.W = pl.W; a.LW = pl.LW; a.groups_total = pl.groups_total; a.groups_conc = pl.groups_conc;
    a.cpo = cpo; a.init_identity = cpo > 1 ? 1 : 0;
    a.alpha_last = alpha_last; a.ell_all = ell_all; a.out_d = out_d; a.ell_last = ell_last;
    return run_forward(semiring, backptr != nullptr, pl, a, stream);
}

int hmm_backward_device(int semiring, const float* A, int64_t a_sb, int64_t a_st, const float* obs, const float* final_,
                        int64_t B, int64_t T, int S, float* beta_all, void* workspace, size_t workspace_bytes,
                        cudaStream_t stream, double* ell_all = nullptr) {
    PassPlan pl;
    int st = make_plan(B, S, true, &pl);
    // large shared-A chains: the backward recursion as a streaming forward pass on A^T (beta_hat + offsets), see hmm_logsum_stream
    if (semiring == FB2_SEMIRING_LOG && a_sb == 0 && a_st == 0 && !final_ && ell_all && obs && T >= 2) {
        const int ls = log_stream_pass(obs + (T - 1) * (int64_t)S, T * (int64_t)S, A, obs, T, B, T - 1, S, nullptr, nullptr, beta_all, ell_all, 1,
                                       workspace, workspace_bytes, stream);
        if (ls == FB2_OK) {
            stream_fill_last_row<<<(unsigned)ceil_div(B * (int64_t)S, 256), 256, 0, stream>>>(beta_all, ell_all, B, T, S);
            FB2_LAUNCH_CHECK();
            return FB2_OK;
        }
        if (ls != FB2_ERR_UNSUPPORTED) return ls;
    }
    if (st != FB2_OK) return st;
    Arena arena(workspace, workspace_bytes);
    float* vec = arena.take<float>((size_t)2 * B * S);
    unsigned* ctr = arena.take<unsigned>(1024);
    if (!arena.ok()) {
        set_error("hmm backward: workspace too small (%zu bytes)", workspace_bytes);
        return FB2_ERR_WORKSPACE;
    }
    FB2_CUDA(cudaMemsetAsync(ctr, 0, 1024 * sizeof(unsigned), stream));
    HmmArgs a{};
    a.A = A; a.a_sb = a_sb; a.a_st = a_st; a.obs = obs; a.final_ = final_;
    a.B = B; a.T = T; a.S = S; a.vec = vec; a.counters = ctr; a.alpha_all = beta_all;
    a.Gg = pl.Gg; a.W = pl.W; a.LW = pl.LW; a.groups_total = pl.groups_total; a.groups_conc = pl.groups_conc;
    a.cpo = 1; a.ell_all = ell_all;
    return run_backward(semiring, pl, a, stream);
}

}  // namespace fb2

using namespace fb2;

extern "C" size_t fb2_hmm_workspace_bytes(int op, int64_t B, int64_t T, int32_t S) {
    size_t base = vec_bytes(B, S) + ctr_bytes() + 512;
    if (op == 0) base += hmm_flow_workspace_bytes(B, S);
    if (op == 3) {  // forward with the alpha history
        const size_t fa = hmm_flow_alpha_workspace_bytes(B, T, S);
        base += hmm_flow_workspace_bytes(B, S);
        if (fa > base) base = fa;
    }
    if ((op <= 1 || op == 3) && S > 64)  // strip-major copy of A (+ column maxima, scratch)
        base += align_up((size_t)((S + 31) / 32) * (S + 8) * 32 * sizeof(float)) + align_up((size_t)S * 4) + align_up((size_t)B * 4);
    if (op == 1) base += align_up((size_t)B * T * S * sizeof(int32_t)) + align_up((size_t)B * sizeof(int32_t));
    if (op == 2) {
        base += align_up((size_t)B * T * S * sizeof(float)) + 2 * align_up((size_t)B * T * 8) + align_up((size_t)B * 8);
        if (S >= 256) base += align_up((size_t)((S + 31) / 32) * S * 32 * sizeof(float)) + align_up((size_t)S * 4) + align_up((size_t)B * 4);  // log stream passes
        const size_t fm = hmm_flow_marginals_workspace_bytes(B, T, S);
        if (fm > base) base = fm;
    }
    return base;
}

extern "C" int fb2_hmm_forward_f32(int semiring, const float* init, int64_t init_sb, const float* A, int64_t a_sb,
                                   int64_t a_st, const float* obs, int64_t B, int64_t T, int32_t S, float* logZ,
                                   float* alpha_all, void* workspace, size_t workspace_bytes, void* stream) {
    int st = check_device();
    if (st != FB2_OK) return st;
    st = check_hmm_args(semiring, A, B, T, S);
    if (st != FB2_OK) return st;
    FB2_REQUIRE(logZ != nullptr, "logZ is null");
    if (B == 0) return FB2_OK;
    // time-homogeneous log-semiring chains take the dataflow kernel (hmm_flow.cu); everything else, and
    // shapes it declines, run the generic cooperative kernel above.  Both are GPU paths.
    // small state spaces: one warp per chain, no inter-CTA traffic
    if (hmm_small_supported(semiring, a_st, S) && !flow_disabled())
        return hmm_small_forward_device(semiring, init, init_sb, A, a_sb, obs, B, T, S, logZ, alpha_all, nullptr, nullptr,
                                        static_cast<cudaStream_t>(stream));
    if (alpha_all != nullptr && hmm_flow_supported(semiring, a_sb, a_st, S, false) && !flow_disabled() &&
        workspace_bytes >= hmm_flow_alpha_workspace_bytes(B, T, S)) {
        st = hmm_flow_forward_alpha_device(init, init_sb, A, obs, B, T, S, logZ, alpha_all, workspace, workspace_bytes,
                                           static_cast<cudaStream_t>(stream));
        if (st != FB2_ERR_UNSUPPORTED) return st;
    }
    if (hmm_flow_supported(semiring, a_sb, a_st, S, alpha_all != nullptr) && !flow_disabled()) {
        const size_t fw = hmm_flow_workspace_bytes(B, S);
        if (workspace_bytes >= fw) {
            st = hmm_flow_forward_device(init, init_sb, A, obs, B, T, S, logZ, nullptr, workspace, workspace_bytes,
                                         static_cast<cudaStream_t>(stream));
            if (st != FB2_ERR_UNSUPPORTED) return st;
            if (getenv("FB2_DEBUG")) fprintf(stderr, "[fb2] dataflow kernel declined: %s\n", fb2_last_error());
        }
    }
    return hmm_forward_device(semiring, init, init_sb, A, a_sb, a_st, obs, nullptr, B, T, S, logZ, alpha_all, nullptr,
                              nullptr, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

extern "C" int fb2_hmm_viterbi_f32(const float* init, int64_t init_sb, const float* A, int64_t a_sb, int64_t a_st,
                                   const float* obs, int64_t B, int64_t T, int32_t S, float* best, int64_t* path,
                                   void* workspace, size_t workspace_bytes, void* stream_) {
    int st = check_device();
    if (st != FB2_OK) return st;
    st = check_hmm_args(FB2_SEMIRING_MAX, A, B, T, S);
    if (st != FB2_OK) return st;
    FB2_REQUIRE(best && path, "null output");
    if (B == 0) return FB2_OK;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    Arena arena(workspace, workspace_bytes);
    int32_t* bp = arena.take<int32_t>((size_t)B * T * S);
    int32_t* last = arena.take<int32_t>((size_t)B);
    if (!arena.ok()) {
        set_error("viterbi: workspace too small (%zu bytes)", workspace_bytes);
        return FB2_ERR_WORKSPACE;
    }
    if (hmm_small_supported(FB2_SEMIRING_MAX, a_st, S) && !flow_disabled())
        st = hmm_small_forward_device(FB2_SEMIRING_MAX, init, init_sb, A, a_sb, obs, B, T, S, best, nullptr, bp, last, stream);
    else
        st = hmm_forward_device(FB2_SEMIRING_MAX, init, init_sb, A, a_sb, a_st, obs, nullptr, B, T, S, best, nullptr, bp, last,
                                static_cast<char*>(workspace) + arena.used, workspace_bytes - arena.used, stream);
    if (st != FB2_OK) return st;
    viterbi_backtrace<<<(unsigned)ceil_div(B, 64), 64, 0, stream>>>(bp, last, B, T, S, path);
    FB2_LAUNCH_CHECK();
    return FB2_OK;
}

extern "C" int fb2_hmm_marginals_f32(const float* init, int64_t init_sb, const float* A, int64_t a_sb, int64_t a_st,
                                     const float* obs, int64_t B, int64_t T, int32_t S, float* logZ, float* gamma,
                                     float* xi_sum, void* workspace, size_t workspace_bytes, void* stream_) {
    int st = check_device();
    if (st != FB2_OK) return st;
    st = check_hmm_args(FB2_SEMIRING_LOG, A, B, T, S);
    if (st != FB2_OK) return st;
    FB2_REQUIRE(logZ && gamma && obs, "null pointer");
    if (B == 0) return FB2_OK;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    // homogeneous chains run both passes on the dataflow kernel (gamma from the recorded messages, xi_sum as a GEMM over time)
    if (T >= 2 && hmm_flow_supported(FB2_SEMIRING_LOG, a_sb, a_st, S, false) && !flow_disabled() &&
        workspace_bytes >= hmm_flow_marginals_workspace_bytes(B, T, S)) {
        st = hmm_flow_marginals_device(init, init_sb, A, obs, B, T, S, logZ, gamma, xi_sum, workspace, workspace_bytes, stream);
        if (st != FB2_ERR_UNSUPPORTED) return st;
    }
    Arena arena(workspace, workspace_bytes);
    float* beta = arena.take<float>((size_t)B * T * S);
    double* ell_a = arena.take<double>((size_t)B * T);
    double* ell_b = arena.take<double>((size_t)B * T);
    double* z_d = arena.take<double>((size_t)B);
    if (!arena.ok()) {
        set_error("marginals: workspace too small (%zu bytes)", workspace_bytes);
        return FB2_ERR_WORKSPACE;
    }
    char* rest = static_cast<char*>(workspace) + arena.used;
    size_t rest_bytes = workspace_bytes - arena.used;
    // small state spaces: both recursions on the warp-per-chain kernel (backward = forward on A^T over reversed rows)
    if (hmm_small_supported(FB2_SEMIRING_LOG, a_st, S) && !flow_disabled()) {
        st = hmm_small_forward_device(FB2_SEMIRING_LOG, init, init_sb, A, a_sb, obs, B, T, S, logZ, gamma, nullptr, nullptr, stream, ell_a, z_d);
        if (st != FB2_OK) return st;
        if (T >= 2) {
            st = hmm_small_forward_device(FB2_SEMIRING_LOG, obs + (T - 1) * (int64_t)S, T * (int64_t)S, A, a_sb, obs, B, T - 1, S, nullptr, beta,
                                          nullptr, nullptr, stream, ell_b, nullptr, T, 1, 1);
            if (st != FB2_OK) return st;
        }
        stream_fill_last_row<<<(unsigned)ceil_div(B * (int64_t)S, 256), 256, 0, stream>>>(beta, ell_b, B, T, S);
        FB2_LAUNCH_CHECK();
        if (xi_sum) {
            const int64_t nx = B * (int64_t)S * S;
            xi_sum_kernel<<<(unsigned)ceil_div(nx, 256), 256, 0, stream>>>(init, init_sb, A, a_sb, a_st, obs, gamma, beta, ell_a, ell_b, z_d, B, T, S,
                                                                         xi_sum);
            FB2_LAUNCH_CHECK();
        }
        const int64_t ng = B * T * (int64_t)S;
        marginals_combine<<<(unsigned)ceil_div(ng, 256), 256, 0, stream>>>(gamma, beta, ell_a, ell_b, z_d, B, T, S);
        FB2_LAUNCH_CHECK();
        return FB2_OK;
    }
    // normalised alpha goes straight into gamma, then gamma = alpha + beta + offsets - logZ in place
    st = hmm_forward_device(FB2_SEMIRING_LOG, init, init_sb, A, a_sb, a_st, obs, nullptr, B, T, S, logZ, gamma, nullptr,
                            nullptr, rest, rest_bytes, stream, 1, nullptr, ell_a, z_d);
    if (st != FB2_OK) return st;
    st = hmm_backward_device(FB2_SEMIRING_LOG, A, a_sb, a_st, obs, nullptr, B, T, S, beta, rest, rest_bytes, stream, ell_b);
    if (st != FB2_OK) return st;
    if (xi_sum) {
        int64_t n = B * (int64_t)S * S;
        xi_sum_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(init, init_sb, A, a_sb, a_st, obs, gamma, beta, ell_a,
                                                                     ell_b, z_d, B, T, S, xi_sum);
        FB2_LAUNCH_CHECK();
    }
    int64_t n = B * T * (int64_t)S;
    marginals_combine<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(gamma, beta, ell_a, ell_b, z_d, B, T, S);
    FB2_LAUNCH_CHECK();
    return FB2_OK;
}

extern "C" size_t fb2_chain_adjoint_workspace_bytes(int64_t B, int64_t T, int32_t S) {
    return 2 * align_up((size_t)B * T * S * 4) + 2 * align_up((size_t)B * T * 8) + vec_bytes(B, S) + ctr_bytes() + 512;
}

extern "C" int fb2_chain_adjoint_f32(int semiring, const float* trans, const int64_t ts[4], const float* init,
                                     const float* final_, int64_t B, int64_t T, int32_t S, float* Z, float* adj,
                                     void* workspace, size_t workspace_bytes, void* stream_) {
    int st = check_device();
    if (st != FB2_OK) return st;
    st = check_hmm_args(semiring, trans, B, T, S);
    if (st != FB2_OK) return st;
    FB2_REQUIRE(Z && adj, "null output");
    FB2_REQUIRE(ts[2] == S && ts[3] == 1, "chain_adjoint: each trans[b,t] must be a contiguous row-major S x S matrix");
    if (B == 0) return FB2_OK;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    Arena arena(workspace, workspace_bytes);
    float* alpha = arena.take<float>((size_t)B * T * S);
    float* beta = arena.take<float>((size_t)B * T * S);
    double* ell_a = arena.take<double>((size_t)B * T);
    double* ell_b = arena.take<double>((size_t)B * T);
    if (semiring != FB2_SEMIRING_LOG) ell_a = ell_b = nullptr;
    if (!arena.ok()) {
        set_error("chain_adjoint: workspace too small (%zu bytes)", workspace_bytes);
        return FB2_ERR_WORKSPACE;
    }
    char* rest = static_cast<char*>(workspace) + arena.used;
    size_t rest_bytes = workspace_bytes - arena.used;
    st = hmm_forward_device(semiring, init, S, trans, ts[0], ts[1], nullptr, final_, B, T, S, Z, alpha, nullptr, nullptr,
                            rest, rest_bytes, stream, 1, nullptr, ell_a, nullptr);
    if (st != FB2_OK) return st;
    st = hmm_backward_device(semiring, trans, ts[0], ts[1], nullptr, final_, B, T, S, beta, rest, rest_bytes, stream, ell_b);
    if (st != FB2_OK) return st;
    int64_t n = B * T * (int64_t)S * S;
    FB2_REQUIRE(ceil_div(n, 256) < (1ll << 31), "chain_adjoint: output too large for one launch");
    adjoint_outer<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(init, alpha, beta, ell_a, ell_b, B, T, S, adj);
    FB2_LAUNCH_CHECK();
    return FB2_OK;
}

extern "C" size_t fb2_hmm_chunk_summary_workspace_bytes(int64_t B, int64_t T, int32_t S) {
    const size_t generic = vec_bytes(B * S, S) + ctr_bytes() + align_up((size_t)B * S * 4) + 512;
    const size_t flow = hmm_flow_summary_workspace_bytes(B, S);
    return generic > flow ? generic : flow;
}

extern "C" int fb2_hmm_chunk_summary_f32(int semiring, const float* A, int64_t a_sb, int64_t a_st, const float* obs,
                                         int64_t B, int64_t T, int32_t S, float* out, double* row_offset, void* workspace,
                                         size_t workspace_bytes, void* stream) {
    int st = check_device();
    if (st != FB2_OK) return st;
    st = check_hmm_args(semiring, A, B, T, S);
    if (st != FB2_OK) return st;
    FB2_REQUIRE(out != nullptr, "out is null");
    if (B == 0) return FB2_OK;
    // log-semiring, homogeneous A, offsets requested: the S identity-started chains of every batch entry run on the dataflow kernel
    if (row_offset && obs && hmm_flow_supported(semiring, a_sb, a_st, S, false) && !flow_disabled() &&
        workspace_bytes >= hmm_flow_summary_workspace_bytes(B, S)) {
        st = hmm_flow_summary_device(A, obs, B, T, S, out, row_offset, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
        if (st != FB2_ERR_UNSUPPORTED) return st;
    }
    Arena arena(workspace, workspace_bytes);
    float* scratch_out = arena.take<float>((size_t)B * S);  // per-row reductions, unused by the caller
    if (!arena.ok()) {
        set_error("chunk_summary: workspace too small (%zu bytes)", workspace_bytes);
        return FB2_ERR_WORKSPACE;
    }
    // S chains per batch entry, chain (b, r) starts from the identity row r: the final vectors are the
    // rows of the chunk's transfer matrix.
    return hmm_forward_device(semiring, nullptr, 0, A, a_sb, a_st, obs, nullptr, B * S, T, S, scratch_out, nullptr, nullptr,
                              nullptr, static_cast<char*>(workspace) + arena.used, workspace_bytes - arena.used,
                              static_cast<cudaStream_t>(stream), S, out, nullptr, nullptr, row_offset);
}
