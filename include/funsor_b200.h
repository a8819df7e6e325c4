/*
 * funsor_b200 -- C ABI of the B200-native Markov-chain sum-product path.
 *
 * This is the drop-in boundary: plain pointers, sizes and strides, no
 * torch types.  Every entry point returns an int status (FB2_OK == 0) and never throws; a message
 * for the last non-zero status of the calling thread is available from fb2_last_error().
 * All `const float*` / `float*` arguments are DEVICE pointers unless the function name ends in
 * `_host`.  Strides are in ELEMENTS.  `stream` is a cudaStream_t passed as void* (NULL = default
 * stream).  Kernels neither allocate nor free: scratch memory is passed in as `workspace`, sized by
 * the matching *_workspace_bytes() query, and must be 256-byte aligned.  No host synchronisation
 * happens inside the device-pointer entry points.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the funsor tree).
 */
#ifndef FUNSOR_B200_H
#define FUNSOR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FB2_VERSION 100

/* status codes */
#define FB2_OK 0
#define FB2_ERR_INVALID_ARGUMENT 1 /* -> Python ValueError                                   */
#define FB2_ERR_UNSUPPORTED 2      /* -> delegate to the captured reference rule / NotImplementedError */
#define FB2_ERR_CUDA 3             /* -> RuntimeError                                        */
#define FB2_ERR_WORKSPACE 4        /* workspace too small or misaligned                      */
#define FB2_ERR_NO_DEVICE 5        /* no sm_100 device: the product path fails loudly, no CPU fallback */
#define FB2_ERR_NUMERIC 6          /* e.g. non-positive pivot: "Too little information to marginalize" (gaussian.py:897-901) */
#define FB2_ERR_ABORTED 7          /* an earlier asynchronous dataflow pass on this device gave up (see fb2_async_status)      */

/* semirings: (sum_op, prod_op) pairs of funsor/ops/builtin.py:266-292 that the path supports */
#define FB2_SEMIRING_LOG 0 /* (ops.logaddexp, ops.add) */
#define FB2_SEMIRING_MAX 1 /* (ops.max, ops.add)       */

int fb2_version(void);
const char* fb2_status_string(int status);
const char* fb2_last_error(void);
/* Asynchronous failures.  The device-pointer entry points never synchronise, so a dataflow pass that gives up while running
 * (its co-resident CTAs were starved by other work for ~1 s, or a host-to-device chunk of the *_host path never landed) cannot
 * change the status of the call that launched it.  Such a pass (i) writes NaN to ALL of its outputs and (ii) raises a sticky
 * per-device condition: fb2_async_status(clear) returns 0 if none is pending on the current device, otherwise a small positive
 * code (which wait gave up), and clears it when `clear` is non-zero.  While the condition is pending every further dataflow
 * launch on the device returns FB2_ERR_ABORTED.  Call it after synchronising the stream.  The *_host entry points (which
 * synchronise themselves) check it and transparently recompute on the cooperative-launch kernel. */
int fb2_async_status(int clear);
/* Device facts the host side needs for grid sizing / roofline arithmetic. */
int fb2_device_query(int device, int* sm_count, int* cc_major, int* cc_minor, size_t* smem_optin_bytes);

/* ------------------------------------------------------------------------------------------------
 * (1) Strided batched semiring matmul -- one level of the scan.
 * Replaces: Contraction[Logaddexp, Add, Tensor, Tensor] funsor/cnf.py:335-379 (-> einsum/numpy_log.py:24-47)
 *           Contraction[Max, Add, Tensor, Tensor]       funsor/cnf.py:312-324 (-> tensor.py:700-726, 313-330)
 *   out[n0,n1,i,j] = (+)_k x[n0,n1,i,k] (x) y[n0,n1,k,j]
 * x, y are arbitrary strided views (the reference passes even/odd time slices);
 * out is contiguous [n0,n1,I,J].  LOG uses the reference's clamped row/column max shifts.
 * argk (nullable, MAX only): int32 [n0,n1,I,J], first maximising k (approximations.py:115-127).
 * workspace (nullable): scratch for the LOG shifts; when given, contiguous-k / contiguous-j operands with
 * I, J >= 64 and K >= 32 run on the tcgen05 tensor cores (3xTF32, fp32-class accuracy), otherwise on CUDA cores.
 */
size_t fb2_semiring_matmul_workspace_bytes(int64_t n0, int64_t n1, int32_t I, int32_t K, int32_t J);
int fb2_semiring_matmul_f32(int semiring, const float* x, const int64_t x_strides[4], const float* y,
                            const int64_t y_strides[4], float* out, int32_t* argk, int64_t n0, int64_t n1,
                            int32_t I, int32_t K, int32_t J, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (2) Chain product over time -- the whole scan.
 * Replaces: sequential_sum_product funsor/sum_product.py:652-701, naive_/mixed_ 625-649/704-795,
 *           eager MarkovProduct 1049-1064 for Tensor factors.
 *   out[b] = trans[b,0] (x) trans[b,1] (x) ... (x) trans[b,T-1]   (matrix products in the semiring)
 * Association order is the reference's balanced even/odd tree (sum_product.py:690-701), so MAX
 * results are bit-identical to the reference evaluated in fp32.  trans is a strided [B,T,S,S] view,
 * out is contiguous [B,S,S].
 */
size_t fb2_chain_product_workspace_bytes(int64_t B, int64_t T, int32_t S);
/* Optional larger workspace: when the caller provides at least this many bytes (0 = not applicable), a contiguous LOG-semiring chain
 * with S a multiple of 128 in [256, 1024] and T >= 4 runs as a linear-domain GEMM tree (TMA + tcgen05, csrc/chain_gemm_tc.cu): the
 * shift / exp / log of einsum/numpy_log.py:24-47 are applied once per chain instead of once per level.  Entries more than ~87 nats
 * below the largest entry of their matrix flush to -inf on this path. */
size_t fb2_chain_product_gemm_workspace_bytes(int64_t B, int64_t T, int32_t S);
int fb2_chain_product_f32(int semiring, const float* trans, const int64_t trans_strides[4], int64_t B,
                          int64_t T, int32_t S, float* out, void* workspace, size_t workspace_bytes,
                          void* stream);

/* ------------------------------------------------------------------------------------------------
 * (3) Factored HMM passes: trans[b,t,i,j] = A[b,t,i,j] + obs[b,t,j] is never materialised.
 * Replaces: DiscreteHMM.log_prob funsor/pyro/hmm.py:129-135 (trans + obs -> sequential_sum_product
 *           -> init + reduce(curr) -> reduce(state)) evaluated as a vector recurrence.
 * init  [B,S]   batch stride init_sb (0 = shared)
 * A     [S,S]   row-major (prev, curr); batch stride a_sb and time stride a_st (0 = shared / homogeneous)
 * obs   [B,T,S] contiguous
 * forward : logZ[B] (required); alpha_all (nullable) [B,T,S] log-domain filtered messages.
 * viterbi : best[B] max-plus value and path[B,T+1] int64 (path[:,0] = state before step 0); first-index
 *           tie-break; recurrence delta_t[j] = max_i(delta_{t-1}[i] + (A[i,j] + obs[t,j])) evaluated in
 *           that fp32 order, so values are bit-identical to a serial fp32 evaluation.
 * marginals (examples/forward_backward.py:58-107,162): logZ[B], gamma[B,T,S] = log p(z_t | obs),
 *           xi_sum (nullable) [B,S,S] = sum_t p(z_{t-1}=i, z_t=j | obs).
 *
 * Tolerance statement.  logZ / best agree with the reference to 1e-5 relative (measured at the BASELINE sizes against float64:
 * 1.4e-6 at T=65536, S=1024; 6.5e-7 at T=1048576, S=512); Viterbi values and paths are bit-identical to a serial fp32 evaluation.
 * One documented deviation: for the LOG semiring with a shared A the per-state outputs (alpha_all, gamma) come from exp-domain kernels
 * that align every message to its largest entry, so a state more than ~87 nats (2^-126) below the largest entry of its message is
 * reported as -inf where the reference's log-space arithmetic keeps a finite value (e.g. -120); such states hold < 1e-37 of the mass
 * and logZ is unaffected.  Set FB2_DISABLE_FLOW=1 to run the log-domain cooperative kernels, which keep the reference's range.
 */
size_t fb2_hmm_workspace_bytes(int op /*0 fwd, 1 viterbi, 2 marginals, 3 fwd with alpha_all*/, int64_t B, int64_t T, int32_t S);
int fb2_hmm_forward_f32(int semiring, const float* init, int64_t init_sb, const float* A, int64_t a_sb,
                        int64_t a_st, const float* obs, int64_t B, int64_t T, int32_t S, float* logZ,
                        float* alpha_all, void* workspace, size_t workspace_bytes, void* stream);
int fb2_hmm_viterbi_f32(const float* init, int64_t init_sb, const float* A, int64_t a_sb, int64_t a_st,
                        const float* obs, int64_t B, int64_t T, int32_t S, float* best, int64_t* path,
                        void* workspace, size_t workspace_bytes, void* stream);
int fb2_hmm_marginals_f32(const float* init, int64_t init_sb, const float* A, int64_t a_sb, int64_t a_st,
                          const float* obs, int64_t B, int64_t T, int32_t S, float* logZ, float* gamma,
                          float* xi_sum, void* workspace, size_t workspace_bytes, void* stream);

/* Time-sharded pieces: one rank owns obs[:, t0:t1].
 * chunk_summary: out[b] = (x)_{t in chunk} (A + obs[b,t]) as an S x S matrix (the summary that is
 *                all-gathered); it is the forward recurrence started from the semiring identity matrix.
 * The fold of gathered summaries into boundary vectors uses fb2_semiring_matmul_f32 with I = 1. */
size_t fb2_hmm_chunk_summary_workspace_bytes(int64_t B, int64_t T, int32_t S);
/* row_offset (nullable) [B,S] double: when given, row i of out[b] is stored RELATIVE to row_offset[b,i] (entries O(1)), so
 * that summaries of long chunks (log-magnitudes ~1e5) keep their fp32 resolution; the true entry is out + row_offset. */
int fb2_hmm_chunk_summary_f32(int semiring, const float* A, int64_t a_sb, int64_t a_st, const float* obs,
                              int64_t B, int64_t T, int32_t S, float* out /*[B,S,S]*/, double* row_offset, void* workspace,
                              size_t workspace_bytes, void* stream);

/* chunk_summary_gemm: the same summary for ONE chain with a shared log-semiring transition matrix A[S,S] (S a multiple of 128 in
 * [256, 1024]) as a blocked balanced tree of S x S products (the reference's own sequential_sum_product order, sum_product.py:690-701)
 * kept in the LINEAR domain: TMA-fed 3xTF32 tcgen05 GEMMs whose epilogue writes the TF32 hi / lo planes of the next level
 * (csrc/chain_gemm_tc.cu).  obs[T,S] contiguous, T >= 2; time is processed in blocks of `block_steps` steps (bounds the workspace:
 * ~ 3 * block_steps * S * S * 4 bytes).  rows[S,S] fp32 with largest entry 0, row_offset[S] float64 (true entry = rows + row_offset). */
size_t fb2_hmm_chunk_summary_gemm_workspace_bytes(int64_t T, int32_t S, int64_t block_steps);
int fb2_hmm_chunk_summary_gemm_f32(const float* A, const float* obs, int64_t T, int32_t S, int64_t block_steps, float* rows,
                                   double* row_offset, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (4) Dense adjoint of the chain (backward pass).
 * Replaces: AdjointTape.adjoint over a scan, funsor/adjoint.py:72-131, 231-294 (+ Scatter tensor.py:639-683).
 *   adj[b,t,i,j] = alpha_{t-1}[b,i] (x) beta_t[b,j];  Z[b] = (+)_j alpha_{T-1}[b,j] (x) final[b,j]
 * init/final nullable (= semiring unit).  trans strided [B,T,S,S]; adj contiguous [B,T,S,S].
 */
size_t fb2_chain_adjoint_workspace_bytes(int64_t B, int64_t T, int32_t S);
int fb2_chain_adjoint_f32(int semiring, const float* trans, const int64_t trans_strides[4], const float* init,
                          const float* final_, int64_t B, int64_t T, int32_t S, float* Z, float* adj,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (5) Gaussian factors in information form (Lam[n,n], eta[n], c): f(x) = c - 1/2 x Lam x' + x eta,
 *     n = 2D over rows (prev; curr).  Lam = P P', eta = P w, c = s - |w|^2/2 of the reference's
 *     (white_vec w, prec_sqrt P, scalar Tensor s) triple (funsor/gaussian.py:419-498, 545, 569).
 * sqrt_to_info : that conversion (fused P P' / P w), elements strided over [N].
 * pair_combine : Contraction[Logaddexp, Add, Gaussian|GaussianMixture x2] funsor/cnf.py:385-389 =
 *                eager_add_gaussian_gaussian gaussian.py:1117-1132 + Gaussian.eager_reduce 841-906 +
 *                _marginalize_after_split 1061-1090, as a Schur complement on the shared block.
 * chain        : sequential_sum_product over Gaussian elements, reference tree order.
 * Element n of a batch lives at base + n*stride; Lam row-major [2D,2D].  D <= 32.
 */
int fb2_gaussian_sqrt_to_info_f32(const float* w, const float* P, const float* s, int64_t N, int32_t dim,
                                  int32_t rank, float* Lam, float* eta, float* c, void* stream);
/* info_to_sqrt : the way back, (Lam, eta, c) -> (white_vec w[N,dim], prec_sqrt P[N,dim,dim], scalar s[N]) with P P' = Lam, P w = eta,
 *                s - |w|^2/2 = c, so that a rule can hand funsor a Gaussian(white_vec, prec_sqrt) (gaussian.py:419-498).  Lam is positive
 *                SEMI-definite in general: diagonally pivoted Cholesky, columns beyond the numerical rank (remaining diagonal
 *                <= rel_tol * max diagonal; rel_tol <= 0 selects dim * 1.2e-7) are zero.  rank (nullable) int32 [N]. */
int fb2_gaussian_info_to_sqrt_f32(const float* Lam, const float* eta, const float* c, int64_t N, int32_t dim, float rel_tol,
                                  float* w, float* P, float* s, int32_t* rank, void* stream);
/* moment_match : collapse a mixture of K Gaussians per element into ONE Gaussian with the mixture's mass, mean and covariance.
 *                Replaces: moment_matching_contract_joint funsor/joint.py:102-150 (the Contraction[Logaddexp, Add, {discrete}, Tensor,
 *                Gaussian] rule of the `moment_matching` interpretation; SwitchingLinearHMM pyro/hmm.py:527-551).
 *                logits[N,K], Lam[N,K,dim,dim], eta[N,K,dim], c[N,K] (component k is exp(logit_k + c_k - x Lam_k x'/2 + x eta_k)) ->
 *                Lo[N,dim,dim], eo[N,dim], co[N]: a normalised Gaussian density scaled by the total mass.  An empty mixture (all
 *                logits -inf) yields co = -inf with unit covariance (joint.py:141-143).  status_flag (nullable) is raised on a
 *                non-positive pivot. */
size_t fb2_gaussian_moment_match_workspace_bytes(int64_t N, int32_t K, int32_t dim);
int fb2_gaussian_moment_match_f32(const float* logits, const float* Lam, const float* eta, const float* c, int64_t N, int32_t K,
                                  int32_t dim, float* Lo, float* eo, float* co, int32_t* status_flag, void* workspace,
                                  size_t workspace_bytes, void* stream);
int fb2_gaussian_pair_combine_f32(const float* Lx, const float* ex, const float* cx, int64_t x_stride,
                                  const float* Ly, const float* ey, const float* cy, int64_t y_stride,
                                  int64_t n0, int64_t n1, int64_t x_s0, int64_t y_s0, int32_t D, float* Lo,
                                  float* eo, float* co, int32_t* status_flag, void* stream);
/* eta          : eta[b,t] = P[b] w[b,t], c[b,t] = s[b,t] - |w[b,t]|^2/2 for a TIME-INVARIANT prec_sqrt P[b] (GaussianHMM with
 *                homogeneous dynamics, pyro/hmm.py:258-270: only white_vec depends on t).
 * chain_homog  : the same scan for Lam[B,2D,2D] shared by all time steps of a chain (eta[B,T,2D], c[B,T] per step): the
 *                T-sized precision array is never materialised; per-level precision operators once per chain, then the tile walk of
 *                kalman_chain below with the elements read from eta / c (T >= 64; FB2_GAUSS_HOMOG_TILES=0: level-by-level sweeps).
 */
/* kalman_eta   : the same per-step (eta, c) for a GaussianHMM with fixed dynamics, computed straight from the observations y[B,T,O]
 *                (pyro/hmm.py:258-270 with matrix_and_mvn_to_funsor pyro/convert.py:278-298 and the value substitution gaussian.py:723-744):
 *                u_t = wo + y_t Pr, eta_t = a0 + Po u_t, c_t = c0 - |u_t|^2 / 2, with the per-chain constants Pr[B,O,O] (obs prec_sqrt),
 *                wo[B,O], a0[B,2D] = P[:, :D] w_trans, Po[B,2D,O] = P[:, D:], c0[B].  No [B,T,D+O] white_vec array is materialised. */
int fb2_gaussian_kalman_eta_f32(const float* y, const float* Pr, const float* wo, const float* a0, const float* Po, const float* c0, int64_t B,
                                int64_t T, int32_t D, int32_t O, float* eta, float* c, void* stream);
/* kalman_constants : everything of the GaussianHMM chain element that does not depend on t, one CTA per chain (pyro/hmm.py:258-270,
 *                     matrix_and_mvn_to_funsor pyro/convert.py:278-298): F[B,D,D], q_loc[B,D], q_tril[B,D,D], H[B,D,O], r_loc[B,O],
 *                     r_tril[B,O,O] -> Lam[B,2D,2D] = P P' (the precision shared by all steps), Pr, wo, a0, Po, c0 as kalman_eta takes them.
 * kalman_chain     : the whole scan of pyro/hmm.py:269 (sequential_sum_product over the T chain elements, sum_product.py:652-701) from the
 *                     observations y[B,T,O] and those constants to (Lam[B,2D,2D], eta[B,2D], c[B]) over (state_-1; state_T-1).  The per-step
 *                     elements and the first K levels of the reference's tree (FB2_KALMAN_K) are evaluated per tile of 2^K steps
 *                     without touching memory.  D <= 16: one thread per tile, fp32 FMA, bit-identical to kalman_eta + chain_homog (K = 4);
 *                     D > 16: 16 tiles per warp as 3xTF32 mma.sync GEMMs against the shared operators (K = 5; within 5e-6 of that path,
 *                     FB2_KALMAN_TC=0 selects the FMA form).  O <= 16, T >= 2. */
int fb2_gaussian_kalman_constants_f32(const float* F, const float* q_loc, const float* q_tril, const float* H, const float* r_loc,
                                      const float* r_tril, int64_t B, int32_t D, int32_t O, float* Lam, float* Pr, float* wo, float* a0,
                                      float* Po, float* c0, void* stream);
size_t fb2_gaussian_kalman_chain_workspace_bytes(int64_t B, int64_t T, int32_t D);
int fb2_gaussian_kalman_chain_f32(const float* y, const float* Lam, const float* Pr, const float* wo, const float* a0, const float* Po,
                                  const float* c0, int64_t B, int64_t T, int32_t D, int32_t O, float* Lo, float* eo, float* co,
                                  void* workspace, size_t workspace_bytes, void* stream);
int fb2_gaussian_eta_f32(const float* w, const float* P, const float* s, int64_t B, int64_t T, int32_t dim, int32_t rank,
                         float* eta, float* c, void* stream);
int fb2_gaussian_chain_homog_f32(const float* Lam, const float* eta, const float* c, int64_t B, int64_t T, int32_t D,
                                 float* Lo, float* eo, float* co, void* workspace, size_t workspace_bytes, void* stream);
/* Backward sampling of a homogeneous Gaussian chain (forward-filter backward-sample).
 * Replaces: recipes.forward_filter_backward_rsample funsor/recipes.py:16-90 specialised to a chain = the adjoint of the scan under
 *           MonteCarlo (montecarlo.py:48-61) with Gaussian._sample (gaussian.py:946-1059) drawing every eliminated state given its two
 *           neighbours in the scan tree (sum_product.py:690-701).
 * chain_homog_keep            : fb2_gaussian_chain_homog_f32 that also keeps every level's eta and the per-level operators in `workspace`
 *                               (sized by fb2_gaussian_chain_sample_workspace_bytes); T >= 2.
 * chain_homog_backward_sample : walks the tree top-down.  x[B,N,T+1,D]: on entry x[..,0,:] and x[..,T,:] hold draws of the two END states
 *                               (from the joint of the chain result and the initial distribution, a [B] problem the caller solves); on
 *                               exit every interior state t = 1..T-1 is filled with  x_m = L^-T (noise_m + v - U [x_s; x_e])  for the pair
 *                               that eliminated it.  noise[B,N,T+1,D] standard normal (rows 0 and T unused): the result is a deterministic
 *                               function of (inputs, noise).  `eta` and `workspace` must be the ones the keep call just used. */
size_t fb2_gaussian_chain_sample_workspace_bytes(int64_t B, int64_t T, int32_t D);
int fb2_gaussian_chain_homog_keep_f32(const float* Lam, const float* eta, const float* c, int64_t B, int64_t T, int32_t D,
                                      float* Lo, float* eo, float* co, void* workspace, size_t workspace_bytes, void* stream);
int fb2_gaussian_chain_homog_backward_sample_f32(const float* eta, int64_t B, int64_t T, int32_t D, int64_t N, const float* noise,
                                                 float* x, void* workspace, size_t workspace_bytes, void* stream);
size_t fb2_gaussian_chain_workspace_bytes(int64_t B, int64_t T, int32_t D);
int fb2_gaussian_chain_f32(const float* Lam, const float* eta, const float* c, int64_t B, int64_t T, int32_t D,
                           float* Lo, float* eo, float* co, void* workspace, size_t workspace_bytes,
                           void* stream);

/* ------------------------------------------------------------------------------------------------
 * (7) Backward sampling of state paths (forward-filter backward-sample).
 * Replaces: recipes.forward_filter_backward_rsample funsor/recipes.py:16-90 specialised to a discrete chain
 *           (sum_product + adjoint.forward_backward under MonteCarlo, montecarlo.py:48-61; Tensor._sample tensor.py:332-427).
 * alpha    [B,T,S] log-domain forward messages (fb2_hmm_forward_f32, alpha_all)
 * uniforms [B,N,T+1] in [0,1): one per draw; draw t uses inverse-CDF lookup in state order, so the result is a
 *          deterministic function of (inputs, uniforms).
 * paths    [B,N,T+1] int64, paths[..,0] = state before the first transition (as fb2_hmm_viterbi_f32)
 * score    [B,N] float64: init[z0] + sum_t A_t[z_t,z_{t+1}] + obs[t,z_{t+1}]  (log q(path) = score - logZ)
 */
size_t fb2_hmm_backward_sample_workspace_bytes(int64_t B, int64_t T, int32_t S, int64_t a_sb, int64_t a_st);
int fb2_hmm_backward_sample_f32(const float* init, int64_t init_sb, const float* A, int64_t a_sb, int64_t a_st,
                                const float* obs, const float* alpha, const float* uniforms, int64_t B, int64_t T,
                                int32_t S, int64_t N, int64_t* paths, double* score, void* workspace,
                                size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Host-buffer entry points (what a CPU-side caller of the reference binds): inputs and outputs are
 * HOST pointers; the call stages them through device memory it owns (one cached arena per thread),
 * runs the device path on its own stream and synchronises before returning.  These are the `e2e`
 * numbers of bench.py.
 */
int fb2_hmm_forward_f32_host(int semiring, const float* init, int64_t init_sb, const float* A, int64_t a_sb,
                             int64_t a_st, const float* obs, int64_t B, int64_t T, int32_t S, float* logZ);
int fb2_chain_product_f32_host(int semiring, const float* trans, int64_t B, int64_t T, int32_t S, float* out);
int fb2_host_arena_release(void);

/* Counters: kernels launched by this library on the calling thread since the last reset (bench.py's
 * gpu_launches). */
int64_t fb2_launch_count(void);
void fb2_launch_count_reset(void);

#ifdef __cplusplus
}
#endif
#endif /* FUNSOR_B200_H */
