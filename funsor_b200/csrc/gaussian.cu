// Gaussian chain elements in information form and their pair contraction.
//
// Reference semantics (funsor tree): product = concatenate along rank gaussian.py:1117-1132,
// marginalise the shared real block gaussian.py:841-906 + 1061-1090, pair contraction cnf.py:385-389,
// scan order sum_product.py:690-701.  The reference carries (white_vec w, prec_sqrt P, scalar s);
// only Lam = P P', eta = P w and c = s - |w|^2/2 are representation independent (testing.py:149-153),
// and those are what these kernels carry.  In that form the contraction of x over (a;b) with y over
// (b;c) is a Schur complement on the shared block b:
//     M = Lx_bb + Ly_aa = L L',  G = [Lx_ab ; Ly_ba],  U = L^-1 G',  v = L^-1 (ex_b + ey_a)
//     Lam' = diag(Lx_aa, Ly_bb) - U'U,  eta' = [ex_a ; ey_b] - U'v,
//     c'   = cx + cy + D/2 log 2pi - sum log diag L + |v|^2/2
// One CTA of 256 threads per pair; everything between the loads and the stores stays in shared memory.
#include "common.cuh"

namespace fb2 {

constexpr int GMAXD = 32;
constexpr float kHalfLog2Pi = 0.91893853320467274178f;

struct GaussPairArgs {
    const float *Lx, *ex, *cx, *Ly, *ey, *cy;
    float *Lo, *eo, *co;
    int32_t* flag;
    int64_t n1;                     // pairs per outer batch index
    int64_t x_s0, x_s1, y_s0, y_s1; // element strides (units of one element) for (outer, inner)
    int64_t o_s0;                   // output element stride of the outer index (inner contiguous)
    int D;
};

__global__ void __launch_bounds__(256) gaussian_pair_combine(GaussPairArgs a) {
    const int D = a.D, n = 2 * D, NC = n + 1;  // NC columns of the augmented right-hand side
    __shared__ float sM[GMAXD][GMAXD + 1];
    __shared__ float sU[GMAXD][2 * GMAXD + 2];  // U (D x 2D) with v appended as column 2D
    __shared__ float s_red[8];
    __shared__ int s_bad;
    const int tid = threadIdx.x;
    const int64_t pair = blockIdx.x;
    const int64_t q0 = pair / a.n1, q1 = pair % a.n1;
    const int64_t xe = q0 * a.x_s0 + q1 * a.x_s1, ye = q0 * a.y_s0 + q1 * a.y_s1, oe = q0 * a.o_s0 + q1;
    const float* Lx = a.Lx + xe * n * n;
    const float* Ly = a.Ly + ye * n * n;
    const float* ex = a.ex + xe * n;
    const float* ey = a.ey + ye * n;
    if (tid == 0) s_bad = 0;

    // ---- stage M = Lx_bb + Ly_aa and the augmented right-hand side [G' | e_b]
    for (int idx = tid; idx < D * D; idx += 256) {
        int r = idx / D, k = idx % D;
        sM[r][k] = Lx[(D + r) * n + D + k] + Ly[r * n + k];
    }
    for (int idx = tid; idx < D * NC; idx += 256) {
        int k = idx / NC, col = idx % NC;  // row k of G' = column k of G (shared-block index)
        float v;
        if (col < D) v = Lx[col * n + D + k];                    // Lx_ab[col, k]
        else if (col < n) v = Ly[(D + (col - D)) * n + k];        // Ly_ba[col-D, k]  (y: rows curr, cols prev)
        else v = ex[D + k] + ey[k];
        sU[k][col] = v;
    }
    __syncthreads();

    // ---- Cholesky M = L L' (right-looking, in place, lower triangle) fused with the forward solve
    //      L U = [G' | e_b]: after column j of L is final, row j of U is final and is eliminated
    //      from the rows below.
    float logdet = 0.f;
    for (int j = 0; j < D; ++j) {
        float d = sM[j][j];
        if (!(d > 0.f)) {
            if (tid == 0) s_bad = 1;
            d = 1.f;
        }
        const float ljj = sqrtf(d);
        const float inv = 1.f / ljj;
        logdet += logf(ljj);
        __syncthreads();  // everyone has read sM[j][j]
        for (int i = j + 1 + tid; i < D; i += 256) sM[i][j] *= inv;
        for (int col = tid; col < NC; col += 256) sU[j][col] *= inv;
        if (tid == 0) sM[j][j] = ljj;
        __syncthreads();
        // trailing update of M (lower triangle) and of U rows below j
        const int rem = D - j - 1;
        for (int idx = tid; idx < rem * rem; idx += 256) {
            int i = j + 1 + idx / rem, k = j + 1 + idx % rem;
            if (k <= i) sM[i][k] -= sM[i][j] * sM[k][j];
        }
        for (int idx = tid; idx < rem * NC; idx += 256) {
            int i = j + 1 + idx / NC, col = idx % NC;
            sU[i][col] -= sM[i][j] * sU[j][col];
        }
        __syncthreads();
    }

    // ---- outputs
    float* Lo = a.Lo + oe * n * n;
    float* eo = a.eo + oe * n;
    for (int idx = tid; idx < n * n; idx += 256) {
        int r = idx / n, s = idx % n;
        float base = 0.f;
        if (r < D && s < D) base = Lx[r * n + s];
        else if (r >= D && s >= D) base = Ly[r * n + s];  // Ly_bb sits at rows/cols D..2D of y
        float acc = 0.f;
#pragma unroll 8
        for (int k = 0; k < D; ++k) acc = fmaf(sU[k][r], sU[k][s], acc);
        Lo[idx] = base - acc;
    }
    float vv = 0.f;
    if (tid < n) {
        float base = tid < D ? ex[tid] : ey[tid];
        float acc = 0.f;
        for (int k = 0; k < D; ++k) acc = fmaf(sU[k][tid], sU[k][n], acc);
        eo[tid] = base - acc;
    }
    if (tid < 32) {
        for (int k = tid; k < D; k += 32) vv += sU[k][n] * sU[k][n];
        vv = warp_sum(vv);
        if (tid == 0) {
            a.co[oe] = a.cx[xe] + a.cy[ye] + D * kHalfLog2Pi - logdet + 0.5f * vv;
            if (s_bad && a.flag) atomicExch(a.flag, 1);
        }
    }
    (void)s_red;
}

// (w, P, s) -> (Lam, eta, c): one CTA per element, P staged in shared memory in chunks of rank.
__global__ void __launch_bounds__(256) gaussian_sqrt_to_info(const float* w, const float* P, const float* s, int dim,
                                                             int rank, float* Lam, float* eta, float* c) {
    extern __shared__ float sm[];  // P chunk [dim][RC+1], w chunk [RC]
    constexpr int RC = 32;
    float* sP = sm;
    float* sw = sm + dim * (RC + 1);
    const int tid = threadIdx.x;
    const int64_t e = blockIdx.x;
    const float* Pe = P + e * (int64_t)dim * rank;
    const float* we = w + e * (int64_t)rank;
    // each thread owns up to ceil((dim*dim + dim + 1)/256) outputs: Lam entries, eta entries, |w|^2
    const int nout = dim * dim + dim + 1;
    float acc[20];
#pragma unroll
    for (int q = 0; q < 20; ++q) acc[q] = 0.f;
    for (int r0 = 0; r0 < rank; r0 += RC) {
        const int rc = min(RC, rank - r0);
        for (int idx = tid; idx < dim * rc; idx += 256) {
            int i = idx / rc, k = idx % rc;
            sP[i * (RC + 1) + k] = Pe[(int64_t)i * rank + r0 + k];
        }
        for (int k = tid; k < rc; k += 256) sw[k] = we[r0 + k];
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 20; ++q) {
            int o = tid + 256 * q;
            if (o >= nout) break;
            float a = acc[q];
            if (o < dim * dim) {
                int i = o / dim, j = o % dim;
                for (int k = 0; k < rc; ++k) a = fmaf(sP[i * (RC + 1) + k], sP[j * (RC + 1) + k], a);
            } else if (o < dim * dim + dim) {
                int i = o - dim * dim;
                for (int k = 0; k < rc; ++k) a = fmaf(sP[i * (RC + 1) + k], sw[k], a);
            } else {
                for (int k = 0; k < rc; ++k) a = fmaf(sw[k], sw[k], a);
            }
            acc[q] = a;
        }
        __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < 20; ++q) {
        int o = tid + 256 * q;
        if (o >= nout) break;
        if (o < dim * dim) Lam[e * (int64_t)dim * dim + o] = acc[q];
        else if (o < dim * dim + dim) eta[e * (int64_t)dim + (o - dim * dim)] = acc[q];
        else c[e] = (s ? s[e] : 0.f) - 0.5f * acc[q];
    }
}

// copy one strided element per outer index: dst[q0*o_s0 + slot] = src[q0*s_s0 + which]
__global__ void gaussian_copy_elements(const float* L, const float* e, const float* c, int64_t s_s0, int64_t which, float* Lo,
                                       float* eo, float* co, int64_t o_s0, int64_t slot, int64_t B, int n) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int per = n * n + n + 1;
    if (idx >= B * per) return;
    int64_t b = idx / per;
    int o = (int)(idx % per);
    int64_t se = b * s_s0 + which, de = b * o_s0 + slot;
    if (o < n * n) Lo[de * n * n + o] = L[se * n * n + o];
    else if (o < n * n + n) eo[de * n + (o - n * n)] = e[se * n + (o - n * n)];
    else co[de] = c[se];
}

__global__ void gaussian_poison_if_flag(const int32_t* flag, float* co, int64_t B) {
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B && *flag) co[b] = __int_as_float(0x7fc00000);
}

static int launch_pair(const float* Lx, const float* ex, const float* cx, int64_t x_s0, int64_t x_s1, const float* Ly,
                       const float* ey, const float* cy, int64_t y_s0, int64_t y_s1, int64_t n0, int64_t n1, int D,
                       float* Lo, float* eo, float* co, int64_t o_s0, int32_t* flag, cudaStream_t stream) {
    if (n0 * n1 == 0) return FB2_OK;
    FB2_REQUIRE(n0 * n1 < (1ll << 31), "gaussian_pair_combine: too many pairs for one launch");
    GaussPairArgs a;
    a.Lx = Lx; a.ex = ex; a.cx = cx; a.Ly = Ly; a.ey = ey; a.cy = cy; a.Lo = Lo; a.eo = eo; a.co = co; a.flag = flag;
    a.n1 = n1; a.x_s0 = x_s0; a.x_s1 = x_s1; a.y_s0 = y_s0; a.y_s1 = y_s1; a.o_s0 = o_s0; a.D = D;
    gaussian_pair_combine<<<(unsigned)(n0 * n1), 256, 0, stream>>>(a);
    FB2_LAUNCH_CHECK();
    return FB2_OK;
}

}  // namespace fb2

using namespace fb2;

extern "C" int fb2_gaussian_sqrt_to_info_f32(const float* w, const float* P, const float* s, int64_t N, int32_t dim,
                                             int32_t rank, float* Lam, float* eta, float* c, void* stream) {
    int st = check_device();
    if (st != FB2_OK) return st;
    FB2_REQUIRE(w && P && Lam && eta && c, "null pointer");
    FB2_REQUIRE(N >= 0 && dim >= 1 && dim <= 2 * GMAXD && rank >= 0, "bad extents N=%lld dim=%d rank=%d", (long long)N, dim, rank);
    FB2_REQUIRE(dim * dim + dim + 1 <= 20 * 256, "dim too large");
    if (N == 0) return FB2_OK;
    FB2_REQUIRE(N < (1ll << 31), "too many elements for one launch");
    size_t smem = ((size_t)dim * 33 + 32) * sizeof(float);
    gaussian_sqrt_to_info<<<(unsigned)N, 256, smem, static_cast<cudaStream_t>(stream)>>>(w, P, s, dim, rank, Lam, eta, c);
    FB2_LAUNCH_CHECK();
    return FB2_OK;
}

extern "C" int fb2_gaussian_pair_combine_f32(const float* Lx, const float* ex, const float* cx, int64_t x_stride,
                                             const float* Ly, const float* ey, const float* cy, int64_t y_stride,
                                             int64_t n0, int64_t n1, int64_t x_s0, int64_t y_s0, int32_t D, float* Lo,
                                             float* eo, float* co, int32_t* status_flag, void* stream) {
    int st = check_device();
    if (st != FB2_OK) return st;
    FB2_REQUIRE(Lx && ex && cx && Ly && ey && cy && Lo && eo && co, "null pointer");
    FB2_REQUIRE(D >= 1 && D <= GMAXD, "state dimension D=%d outside [1, %d]", D, GMAXD);
    FB2_REQUIRE(n0 >= 0 && n1 >= 0, "bad extents");
    return launch_pair(Lx, ex, cx, x_s0, x_stride, Ly, ey, cy, y_s0, y_stride, n0, n1, D, Lo, eo, co, n1, status_flag,
                       static_cast<cudaStream_t>(stream));
}

static size_t gauss_elem_floats(int n) { return (size_t)n * n + n + 1; }

extern "C" size_t fb2_gaussian_chain_workspace_bytes(int64_t B, int64_t T, int32_t D) {
    int n = 2 * D;
    int64_t t1 = (T + 1) / 2, t2 = (t1 + 1) / 2;
    size_t per = 0;
    for (int64_t tt : {t1, t2})
        per += align_up((size_t)B * tt * n * n * 4) + align_up((size_t)B * tt * n * 4) + align_up((size_t)B * tt * 4);
    (void)gauss_elem_floats;
    return per + align_up(4) + 256;
}

extern "C" int fb2_gaussian_chain_f32(const float* Lam, const float* eta, const float* c, int64_t B, int64_t T, int32_t D,
                                      float* Lo, float* eo, float* co, void* workspace, size_t workspace_bytes,
                                      void* stream_) {
    int st = check_device();
    if (st != FB2_OK) return st;
    FB2_REQUIRE(Lam && eta && c && Lo && eo && co, "null pointer");
    FB2_REQUIRE(D >= 1 && D <= GMAXD, "state dimension D=%d outside [1, %d]", D, GMAXD);
    FB2_REQUIRE(B >= 0 && T >= 1, "bad extents");
    if (B == 0) return FB2_OK;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int n = 2 * D;
    const int per = n * n + n + 1;
    if (T == 1) {
        gaussian_copy_elements<<<(unsigned)ceil_div(B * per, 256), 256, 0, stream>>>(Lam, eta, c, 1, 0, Lo, eo, co, 1, 0, B, n);
        FB2_LAUNCH_CHECK();
        return FB2_OK;
    }
    int64_t t1 = (T + 1) / 2, t2 = (t1 + 1) / 2;
    Arena arena(workspace, workspace_bytes);
    float *bL[2], *be[2], *bc[2];
    int64_t tt[2] = {t1, t2};
    for (int q = 0; q < 2; ++q) {
        bL[q] = arena.take<float>((size_t)B * tt[q] * n * n);
        be[q] = arena.take<float>((size_t)B * tt[q] * n);
        bc[q] = arena.take<float>((size_t)B * tt[q]);
    }
    int32_t* flag = arena.take<int32_t>(1);
    if (!arena.ok()) {
        set_error("gaussian_chain: workspace too small (%zu bytes)", workspace_bytes);
        return FB2_ERR_WORKSPACE;
    }
    FB2_CUDA(cudaMemsetAsync(flag, 0, sizeof(int32_t), stream));
    const float *cL = Lam, *ce = eta, *cc = c;
    int64_t cur_s0 = T;  // elements per batch row of the current level
    int64_t d = T;
    int which = 0;
    while (d > 1) {
        int64_t half = d / 2, nd = (d + 1) / 2;
        float* dL = nd == 1 ? Lo : bL[which];
        float* de = nd == 1 ? eo : be[which];
        float* dc = nd == 1 ? co : bc[which];
        int64_t o_s0 = nd;  // when nd == 1 outputs are [B] contiguous: stride 1 == nd
        st = launch_pair(cL, ce, cc, cur_s0, 2, cL + (size_t)n * n, ce + n, cc + 1, cur_s0, 2, B, half, D, dL, de, dc, o_s0,
                         flag, stream);
        if (st != FB2_OK) return st;
        if (nd != half) {
            gaussian_copy_elements<<<(unsigned)ceil_div(B * per, 256), 256, 0, stream>>>(cL, ce, cc, cur_s0, d - 1, dL, de, dc,
                                                                                       o_s0, half, B, n);
            FB2_LAUNCH_CHECK();
        }
        cL = dL; ce = de; cc = dc;
        cur_s0 = nd;
        d = nd;
        which ^= 1;
    }
    // non-positive pivot ("Too little information to marginalize", gaussian.py:897-901) -> NaN in the scalar
    // term, so the caller can detect it without a host synchronisation inside this call
    gaussian_poison_if_flag<<<(unsigned)ceil_div(B, 256), 256, 0, stream>>>(flag, co, B);
    FB2_LAUNCH_CHECK();
    return FB2_OK;
}
