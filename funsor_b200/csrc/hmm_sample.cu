// Backward sampling of state paths from the filtered messages of a discrete chain (forward-filter backward-sample).
//
// Reference: recipes.forward_filter_backward_rsample funsor/recipes.py:16-90 = sum_product + adjoint.forward_backward under the
// MonteCarlo interpretation (montecarlo.py:48-61), whose discrete leaf is Tensor._sample tensor.py:332-427 (a categorical draw
// from the normalised logits).  For a chain this is the classical recursion
//     z_T ~ softmax(alpha_{T-1}),   z_t ~ softmax_i( alpha_{t-1}[i] + A_t[i, z_{t+1}] ),   z_0 ~ softmax_i( init[i] + A_0[i, z_1] )
// with alpha the log-domain forward messages (fb2_hmm_forward_f32, alpha_all) and path index t+1 = state after transition t.
//
// One warp per (chain, sample); a CTA holds WARPS samples of the SAME chain so that the alpha rows they stream are shared in
// L1/L2.  Each draw is an inverse-CDF lookup with a caller-provided uniform u[b,n,t]: weights w_i are staged in the warp's
// shared-memory slice, p_i = exp(w_i - max), and the first index whose running sum reaches u * sum(p) is found by a warp scan
// over 32-wide tiles (ballot).  Deterministic given the uniforms, so the oracle can check every conditional draw.
// A is read through a transposed copy (row z of A^T = column z of A, contiguous) when it does not vary with time.
#include "common.cuh"

namespace fb2 {

constexpr int SAMPLE_WARPS = 8;

struct SampleArgs {
    const float* init;  // nullable (= zeros)
    int64_t init_sb;
    const float* A;     // [.., S, S] (prev, curr)
    const float* AT;    // nullable: transposed copy [Bt, S, S] (curr, prev), batch stride at_sb
    int64_t a_sb, a_st, at_sb;
    const float* obs;    // [B,T,S]
    const float* alpha;  // [B,T,S]
    const float* u;      // [B,N,T+1]
    int64_t B, T, N;
    int S, Sp;
    int64_t* paths;  // [B,N,T+1]
    double* score;   // [B,N]: init[z0] + sum_t A_t[z_t, z_{t+1}] + obs[t, z_{t+1}]
};

__global__ void transpose_batched(const float* A, float* AT, int S, int64_t a_sb) {
    __shared__ float tile[32][33];
    const float* a = A + blockIdx.z * a_sb;
    float* at = AT + (int64_t)blockIdx.z * S * S;
    const int x = blockIdx.x * 32 + threadIdx.x, y0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y)
        if (x < S && y0 + r < S) tile[r][threadIdx.x] = a[(int64_t)(y0 + r) * S + x];
    __syncthreads();
    const int xo = blockIdx.y * 32 + threadIdx.x, yo0 = blockIdx.x * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y)
        if (xo < S && yo0 + r < S) at[(int64_t)(yo0 + r) * S + xo] = tile[threadIdx.x][r];
}

__global__ void __launch_bounds__(SAMPLE_WARPS * 32) hmm_backward_sample(SampleArgs a) {
    extern __shared__ float sample_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n = (int64_t)blockIdx.x * SAMPLE_WARPS + warp, b = blockIdx.y;
    if (n >= a.N) return;  // whole warp leaves together; no block-wide barriers below
    float* w = sample_smem + (size_t)warp * a.Sp;
    const int S = a.S;
    const float* u = a.u + (b * a.N + n) * (a.T + 1);
    int64_t* path = a.paths + (b * a.N + n) * (a.T + 1);
    bool dead = false;
    int z = -1;  // state already drawn at path index t + 1
    float u_next = u[a.T];
    for (int64_t t = a.T; t >= 0; --t) {
        const float* base = t > 0 ? a.alpha + (b * a.T + (t - 1)) * S : (a.init ? a.init + b * a.init_sb : nullptr);
        const float ut = u_next;
        if (t > 0) {
            u_next = u[t - 1];
            // the alpha row of the NEXT draw does not depend on this one: pull its lines towards the SM now (one 128-byte line per lane)
            const float* nb = t > 1 ? a.alpha + (b * a.T + (t - 2)) * S : (a.init ? a.init + b * a.init_sb : nullptr);
            if (nb)
                for (int off = lane * 32; off < S; off += 32 * 32) asm volatile("prefetch.global.L1 [%0];" ::"l"(nb + off));
        }
        // column z of the transition used between path[t] and path[t+1] (time index t)
        const float* col = nullptr;
        int64_t col_stride = 1;
        if (z >= 0) {
            if (a.AT) {
                col = a.AT + b * a.at_sb + (int64_t)z * S;
            } else {
                col = a.A + b * a.a_sb + t * a.a_st + z;
                col_stride = S;
            }
        }
        float m = -INFINITY;
        for (int i = lane; i < S; i += 32) {
            float v = base ? base[i] : 0.f;
            if (col) v += col[(int64_t)i * col_stride];
            w[i] = v;
            m = fmaxf(m, v);
        }
        m = warp_max(m);
        int chosen = 0;
        if (m > -INFINITY) {
            float tot = 0.f;
            for (int i = lane; i < S; i += 32) {
                const float p = __expf(w[i] - m);
                w[i] = p;
                tot += p;
            }
            tot = warp_sum(tot);
            __syncwarp();
            const float target = ut * tot;
            float run = 0.f;
            int found = -1, last_pos = 0;
            for (int i0 = 0; i0 < S && found < 0; i0 += 32) {
                const int i = i0 + lane;
                const float p = i < S ? w[i] : 0.f;
                float c = p;  // inclusive scan within the tile
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const float q = __shfl_up_sync(0xffffffffu, c, o);
                    if (lane >= o) c += q;
                }
                c += run;
                const unsigned hit = __ballot_sync(0xffffffffu, p > 0.f && c >= target);
                const unsigned pos = __ballot_sync(0xffffffffu, p > 0.f);
                if (pos) last_pos = i0 + 31 - __clz(pos);
                if (hit) found = i0 + __ffs(hit) - 1;
                run = __shfl_sync(0xffffffffu, c, 31);
            }
            chosen = found >= 0 ? found : last_pos;  // rounding left the target above the final running sum
        } else {
            dead = true;  // no path has positive mass
        }
        __syncwarp();
        if (lane == 0) path[t] = chosen;
        z = chosen;
    }
    // score of the drawn path, after the fact and in parallel over t (the gathers would otherwise sit on every step's critical
    // path): init[z0] + sum_t A_t[z_t, z_{t+1}] + obs[t, z_{t+1}]
    __syncwarp();
    double score = 0.0;
    for (int64_t t = lane; t < a.T; t += 32) {
        const int64_t z0 = path[t], z1 = path[t + 1];
        score += (double)a.A[b * a.a_sb + t * a.a_st + z0 * S + z1] + (double)a.obs[(b * a.T + t) * S + z1];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) score += __shfl_xor_sync(0xffffffffu, score, o);
    if (a.init) score += (double)a.init[b * a.init_sb + path[0]];
    if (dead) score = -INFINITY;
    if (lane == 0) a.score[b * a.N + n] = score;
}

size_t hmm_sample_workspace_bytes(int64_t B, int64_t T, int S, int64_t a_sb, int64_t a_st) {
    if (a_st != 0) return 256;
    return align_up((size_t)(a_sb ? B : 1) * S * S * sizeof(float)) + 256;
}

int hmm_backward_sample_device(const float* init, int64_t init_sb, const float* A, int64_t a_sb, int64_t a_st, const float* obs,
                               const float* alpha, const float* uniforms, int64_t B, int64_t T, int S, int64_t N, int64_t* paths,
                               double* score, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    SampleArgs a;
    a.init = init; a.init_sb = init_sb; a.A = A; a.a_sb = a_sb; a.a_st = a_st; a.obs = obs; a.alpha = alpha; a.u = uniforms;
    a.B = B; a.T = T; a.N = N; a.S = S; a.Sp = (S + 31) / 32 * 32; a.paths = paths; a.score = score;
    a.AT = nullptr; a.at_sb = 0;
    if (a_st == 0) {
        Arena arena(workspace, workspace_bytes);
        const int64_t nb = a_sb ? B : 1;
        float* AT = arena.take<float>((size_t)nb * S * S);
        if (!arena.ok()) {
            set_error("hmm backward sample: workspace too small (%zu bytes)", workspace_bytes);
            return FB2_ERR_WORKSPACE;
        }
        if (nb > 65535) return FB2_ERR_UNSUPPORTED;
        transpose_batched<<<dim3((unsigned)ceil_div(S, 32), (unsigned)ceil_div(S, 32), (unsigned)nb), dim3(32, 8), 0, stream>>>(A, AT, S, a_sb);
        FB2_LAUNCH_CHECK();
        a.AT = AT;
        a.at_sb = a_sb ? (int64_t)S * S : 0;
    }
    const size_t smem = (size_t)SAMPLE_WARPS * a.Sp * sizeof(float);
    if (smem > 200 * 1024) {
        set_error("hmm backward sample: S=%d exceeds the shared-memory staging limit (%d states)", S, 200 * 1024 / 4 / SAMPLE_WARPS);
        return FB2_ERR_UNSUPPORTED;
    }
    if (B > 65535) return FB2_ERR_UNSUPPORTED;
    FB2_CUDA(cudaFuncSetAttribute(hmm_backward_sample, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hmm_backward_sample<<<dim3((unsigned)ceil_div(N, SAMPLE_WARPS), (unsigned)B), SAMPLE_WARPS * 32, smem, stream>>>(a);
    FB2_LAUNCH_CHECK();
    return FB2_OK;
}

}  // namespace fb2

extern "C" size_t fb2_hmm_backward_sample_workspace_bytes(int64_t B, int64_t T, int32_t S, int64_t a_sb, int64_t a_st) {
    return fb2::hmm_sample_workspace_bytes(B, T, S, a_sb, a_st);
}

extern "C" int fb2_hmm_backward_sample_f32(const float* init, int64_t init_sb, const float* A, int64_t a_sb, int64_t a_st, const float* obs,
                                           const float* alpha, const float* uniforms, int64_t B, int64_t T, int32_t S, int64_t N,
                                           int64_t* paths, double* score, void* workspace, size_t workspace_bytes, void* stream) {
    using namespace fb2;
    int st = check_device();
    if (st != FB2_OK) return st;
    FB2_REQUIRE(A && obs && alpha && uniforms && paths && score, "hmm backward sample: null pointer argument");
    FB2_REQUIRE(B >= 1 && T >= 1 && S >= 1 && N >= 1, "hmm backward sample: B, T, S, N must be positive");
    return hmm_backward_sample_device(init, init_sb, A, a_sb, a_st, obs, alpha, uniforms, B, T, S, N, paths, score, workspace,
                                      workspace_bytes, static_cast<cudaStream_t>(stream));
}
