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
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

namespace fb2 {

constexpr int GMAXD = 32;
constexpr float kHalfLog2Pi = 0.91893853320467274178f;

struct GaussPairArgs {
    const float *Lx, *ex, *cx, *Ly, *ey, *cy;
    float *Lo, *eo, *co;
    int32_t* flag;
    int64_t n1;                     // pairs per outer batch index
    int64_t x_s0, x_s1, y_s0, y_s1; // element strides (units of one element) for (outer, inner) of eta / c
    int64_t lx_s0, lx_s1, ly_s0, ly_s1;  // the same for the Lam arrays (stride 0 over time = time-invariant precision)
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
    const float* Lx = a.Lx + (q0 * a.lx_s0 + q1 * a.lx_s1) * n * n;
    const float* Ly = a.Ly + (q0 * a.ly_s0 + q1 * a.ly_s1) * n * n;
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

// ---------------------------------------------------------------------------------------------
// Warp-per-pair version of the same contraction.  The CTA-per-pair kernel above spends its time in ~100 block
// barriers around a 32-step Cholesky (measured: 1.07e7 pairs/s, ~5 % of the fp32 roofline of the arithmetic);
// here one warp owns a pair, nothing but __syncwarp separates the phases, and 8 warps per CTA (2 CTAs per SM) keep
// 16 pairs in flight per SM.
//   phase 1  lane i holds row i of M = Lx_bb + Ly_aa in registers; right-looking Cholesky with shuffles
//   phase 2  L goes to the warp's shared-memory slot; lane l forward-substitutes right-hand-side columns l, l+32 (, 64)
//            of [G' | e_b] reading L as broadcast loads -> U columns in registers and in shared memory
//   phase 3  lane l produces output rows l and l+32:  Lam'[r, :] = base[r, :] - U[:, r]' U,  eta'[r] likewise
// warps per CTA: D = 32 runs one 12-warp CTA per SM (168 registers per thread) with all warps in step, smaller D three 4-warp CTAs
template <int DT>
__host__ __device__ constexpr int gw_warps() { return DT >= 32 ? 12 : 4; }
template <int DT>
struct GaussWarpSmem {
    float L[DT][DT + 1];
    float U[DT][(2 * DT + 4) < 20 ? 20 : (2 * DT + 4)];  // column 2*DT holds v; rows are 16-byte aligned and at least 16 + 4 wide
};

// DT = compile-time padded state dimension (D <= DT): the shared block is padded with an identity and zero right-hand
// sides, so every loop has static bounds, every array stays in registers and the shuffles are unconditional (a first
// version with run-time bounds put the row array in local memory and wrapped each shuffle in divergence handling).
// One pair contraction by one warp (see the phases above).  SYNC: the warps of the CTA step through the phases together;
// EXPORT: also write the operators (U and L^-1) that update eta and c of any other pair with the same two precision matrices.
template <int DT, bool SYNC, bool EXPORT>
__device__ __forceinline__ void gauss_pair_body(const float* Lx, const float* Ly, const float* ex, const float* ey, float cxv, float cyv, int D,
                                                float* Lo, float* eo, float* co, int32_t* flag, GaussWarpSmem<DT>& sm, int lane,
                                                float* opU, float* opLinv) {
    const int n = 2 * D;
    constexpr int NCOL = (2 * DT + 31) / 32;  // right-hand-side columns / output rows per lane
    // right-hand sides of phase 2 (G' columns, lane = column).  For D <= 16 they are requested before the Cholesky, whose
    // serial chain hides the latency; at D = 32 the 64 extra live registers spill, so they are loaded at phase 2.
    float u[NCOL][DT];
    auto load_rhs = [&]() {
#pragma unroll
        for (int q = 0; q < NCOL; ++q) {
            const int c = lane + 32 * q;
            const int blk = c / DT, cc = c % DT;  // 0: a-block, 1: c-block
#pragma unroll
            for (int k = 0; k < DT; ++k) {
                float g = 0.f;
                if (c < 2 * DT && cc < D && k < D) g = blk == 0 ? __ldg(Lx + (D + k) * n + cc) : __ldg(Ly + k * n + D + cc);
                u[q][k] = g;
            }
        }
    };
    if (DT <= 16) load_rhs();

    // ---- phase 1: Cholesky of M = Lx_bb + Ly_aa, lane = row (M is symmetric: column `lane` is read, coalesced, as row `lane`)
    float m[DT];
#pragma unroll
    for (int k = 0; k < DT; ++k)
        m[k] = (lane < D && k < D) ? __ldg(Lx + (D + k) * n + D + lane) + __ldg(Ly + k * n + lane) : (k == lane ? 1.f : 0.f);
    float logdet = 0.f, dinv = 1.f;  // dinv: 1 / L[lane][lane]
    int bad = 0;
#pragma unroll
    for (int j = 0; j < DT; ++j) {
        float d = __shfl_sync(0xffffffffu, m[j], j);
        if (!(d > 0.f)) {
            bad = 1;
            d = 1.f;
        }
        // 1/sqrt(d): MUFU.RSQ + one Newton step (full fp32 accuracy, no slow-path call; sqrtf and the divisions below
        // cost ~25 instructions each and the kernel is instruction-fetch bound)
        float inv = rsqrtf(d);
        inv = inv * fmaf(-0.5f * d * inv, inv, 1.5f);
        const float ljj = d * inv;
        logdet += __logf(ljj);
        if (lane == j) dinv = inv;
        const float lij = lane > j ? m[j] * inv : (lane == j ? ljj : 0.f);
        m[j] = lij;
#pragma unroll
        for (int k = j + 1; k < DT; ++k) m[k] = fmaf(-lij, __shfl_sync(0xffffffffu, lij, k), m[k]);
    }
    // v = L^-1 (ex_b + ey_a), still lane = row: lane k ends up with v[k]
    float val = lane < D ? ex[D + lane] + ey[lane] : 0.f;
#pragma unroll
    for (int j = 0; j < DT; ++j) {
        const float vj = __shfl_sync(0xffffffffu, val, j) * __shfl_sync(0xffffffffu, dinv, j);
        val = lane == j ? vj : (lane > j ? fmaf(-m[j], vj, val) : val);
    }
    if (lane < DT) {
#pragma unroll
        for (int k = 0; k < DT; ++k) sm.L[lane][k] = m[k];
        sm.L[lane][lane] = dinv;  // phase 2 only needs the inverse of the diagonal
        sm.U[lane][2 * DT] = val;
    }
    __syncwarp();

    if (SYNC) __syncthreads();
    // ---- phase 2: forward substitution, lane = right-hand-side columns c = lane + 32 q of G' (padded layout: the a-block
    // occupies columns [0, DT), the c-block columns [DT, 2 DT)); G'[k, c] is read through its symmetric counterpart so
    // that the lanes coalesce: Lx_ab[c, k] = Lx_ba[k, c], Ly_ba[c, k] = Ly_ab[k, c]
    if (DT > 16) load_rhs();
#pragma unroll
    for (int k = 0; k < DT; ++k) {
        float acc[NCOL];
#pragma unroll
        for (int q = 0; q < NCOL; ++q) acc[q] = u[q][k];
#pragma unroll
        for (int mm = 0; mm < k; ++mm) {
            const float l = sm.L[k][mm];
#pragma unroll
            for (int q = 0; q < NCOL; ++q) acc[q] = fmaf(-l, u[q][mm], acc[q]);
        }
        const float inv = sm.L[k][k];  // holds 1 / L[k][k]
#pragma unroll
        for (int q = 0; q < NCOL; ++q) {
            u[q][k] = acc[q] * inv;
            const int c = lane + 32 * q;
            if (c < 2 * DT) sm.U[k][c] = u[q][k];
        }
    }
    __syncwarp();

    if (EXPORT) {
        // operators for the eta-only update of a time-homogeneous chain: U (padded layout, DT x 2DT) and Linv = L^-1, obtained
        // by forward substitution on the identity with lane = column
#pragma unroll
        for (int k = 0; k < DT; ++k) {
#pragma unroll
            for (int q = 0; q < NCOL; ++q)
                if (lane + 32 * q < 2 * DT) opU[k * 2 * DT + lane + 32 * q] = u[q][k];
        }
        float x[DT];
#pragma unroll
        for (int k = 0; k < DT; ++k) {
            float acc = k == lane ? 1.f : 0.f;
#pragma unroll
            for (int mm = 0; mm < k; ++mm) acc = fmaf(-sm.L[k][mm], x[mm], acc);
            x[k] = acc * sm.L[k][k];
        }
        if (lane < DT) {
#pragma unroll
            for (int k = 0; k < DT; ++k) opLinv[k * DT + lane] = x[k];
        }
    }
    if (SYNC) __syncthreads();
    // ---- phase 3: outputs, lane = output rows (padded index) r = lane + 32 q; inputs and output are symmetric, so element
    // (r, s) is read from / written to (s, r) and the lanes coalesce
    // The diagonal blocks of the result start from the operands' own blocks (Lx_aa, Ly_cc).  Those reads are issued one
    // 16-column strip AHEAD of the strip being accumulated: measured with ncu (source page), reading them next to the
    // store that consumes them serialised 128 global-load latencies per pair and was 55 % of all stall samples.
    auto load_base = [&](int s0, float (&dst)[NCOL][16]) {
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const int sp = s0 + t, sblk = sp / DT, sc = sp % DT;
#pragma unroll
            for (int q = 0; q < NCOL; ++q) {
                const int rp = lane + 32 * q, rblk = rp / DT, rc = rp % DT;
                float base = 0.f;
                if (sp < 2 * DT && sc < D && rp < 2 * DT && rc < D && rblk == sblk) {
                    const int sidx = sblk * D + sc, r = rblk * D + rc;
                    base = __ldg((rblk == 0 ? Lx : Ly) + sidx * n + r);
                }
                dst[q][t] = base;
            }
        }
    };
    float nb[NCOL][16];
    load_base(0, nb);
#pragma unroll 1
    for (int s0 = 0; s0 < 2 * DT; s0 += 16) {
        float acc[NCOL][16];
#pragma unroll
        for (int q = 0; q < NCOL; ++q)
#pragma unroll
            for (int t = 0; t < 16; ++t) acc[q][t] = nb[q][t];
        if (s0 + 16 < 2 * DT) load_base(s0 + 16, nb);
#pragma unroll
        for (int k = 0; k < DT; ++k) {
            const float4* up = reinterpret_cast<const float4*>(&sm.U[k][s0]);
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
                const float4 uv = up[q4];
#pragma unroll
                for (int q = 0; q < NCOL; ++q) {
                    acc[q][4 * q4 + 0] = fmaf(-u[q][k], uv.x, acc[q][4 * q4 + 0]);
                    acc[q][4 * q4 + 1] = fmaf(-u[q][k], uv.y, acc[q][4 * q4 + 1]);
                    acc[q][4 * q4 + 2] = fmaf(-u[q][k], uv.z, acc[q][4 * q4 + 2]);
                    acc[q][4 * q4 + 3] = fmaf(-u[q][k], uv.w, acc[q][4 * q4 + 3]);
                }
            }
        }
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const int sp = s0 + t;  // padded column index
            const int sblk = sp / DT, sc = sp % DT;
            if (sp < 2 * DT && sc < D) {
                const int sidx = sblk * D + sc;  // real column index
#pragma unroll
                for (int q = 0; q < NCOL; ++q) {
                    const int rp = lane + 32 * q, rblk = rp / DT, rc = rp % DT;
                    if (rp < 2 * DT && rc < D) Lo[sidx * n + rblk * D + rc] = acc[q][t];
                }
            }
        }
    }
    // eta' and the scalar need column v of U
    float e[NCOL], vv = 0.f;
#pragma unroll
    for (int q = 0; q < NCOL; ++q) e[q] = 0.f;
#pragma unroll
    for (int k = 0; k < DT; ++k) {
        const float v = sm.U[k][2 * DT];
#pragma unroll
        for (int q = 0; q < NCOL; ++q) e[q] = fmaf(u[q][k], v, e[q]);
        vv = fmaf(v, v, vv);
    }
#pragma unroll
    for (int q = 0; q < NCOL; ++q) {
        const int rp = lane + 32 * q, rblk = rp / DT, rc = rp % DT;
        if (rp < 2 * DT && rc < D) {
            const int r = rblk * D + rc;
            eo[r] = (rblk == 0 ? ex[r] : ey[r]) - e[q];
        }
    }
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
        *co = cxv + cyv + D * kHalfLog2Pi - logdet + 0.5f * vv;
        if (bad && flag) atomicExch(flag, 1);
    }
}

template <int DT>
__global__ void __launch_bounds__(gw_warps<DT>() * 32, DT >= 32 ? 1 : 3) gaussian_pair_combine_warp(GaussPairArgs a, int64_t npairs) {
    constexpr int GW_WARPS = gw_warps<DT>();
    extern __shared__ __align__(16) unsigned char gw_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    GaussWarpSmem<DT>& sm = reinterpret_cast<GaussWarpSmem<DT>*>(gw_raw)[warp];
    const int D = a.D, n = 2 * D;
    // The loop is CTA-uniform and the warps of a CTA meet at a barrier per pair: the body is ~150 KB of straight-line code
    // (everything is unrolled so that the register arrays are statically indexed), and warps drifting apart through it made
    // instruction fetch the top stall reason (ncu: no_instructions 31 %).  In step, the four warps share each fetched line.
    for (int64_t pair0 = (int64_t)blockIdx.x * GW_WARPS; pair0 < npairs; pair0 += (int64_t)gridDim.x * GW_WARPS) {
        __syncthreads();
        // warps past the end redo the last pair (identical values to identical addresses) so that every warp reaches
        // the barriers inside the body
        const int64_t pair = pair0 + warp < npairs ? pair0 + warp : npairs - 1;
        const int64_t q0 = pair / a.n1, q1 = pair % a.n1;
        const int64_t xe = q0 * a.x_s0 + q1 * a.x_s1, ye = q0 * a.y_s0 + q1 * a.y_s1, oe = q0 * a.o_s0 + q1;
        const float* Lx = a.Lx + (q0 * a.lx_s0 + q1 * a.lx_s1) * n * n;
        const float* Ly = a.Ly + (q0 * a.ly_s0 + q1 * a.ly_s1) * n * n;
        const float* ex = a.ex + xe * n;
        const float* ey = a.ey + ye * n;

        // L2 prefetch of the operands of the pair this warp takes next (one 128-byte line per lane and instruction), so that
        // its loads find them in L2 instead of HBM.  Measured: helps at D <= 16, costs 4 % at D = 32 -> compiled out there.
        if (DT <= 16) {
            const int64_t nxt = pair + (int64_t)gridDim.x * GW_WARPS;
            if (nxt < npairs) {
                const int64_t r0 = nxt / a.n1, r1 = nxt % a.n1;
                const char* px = reinterpret_cast<const char*>(a.Lx + (r0 * a.lx_s0 + r1 * a.lx_s1) * n * n);
                const char* py = reinterpret_cast<const char*>(a.Ly + (r0 * a.ly_s0 + r1 * a.ly_s1) * n * n);
                const int bytes = n * n * 4;
                for (int off = lane * 128; off < bytes; off += 32 * 128) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(px + off));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(py + off));
                }
            }
        }
        gauss_pair_body<DT, true, false>(Lx, Ly, ex, ey, a.cx[xe], a.cy[ye], D, a.Lo + oe * n * n, a.eo + oe * n, a.co + oe, a.flag, sm, lane,
                                         nullptr, nullptr);
        __syncwarp();  // the shared-memory slot is reused by the next pair
    }
}

template <int DT>
static int launch_pair_warp(const GaussPairArgs& a, int64_t npairs, cudaStream_t stream) {
    constexpr int GW_WARPS = gw_warps<DT>();
    const size_t smem = GW_WARPS * sizeof(GaussWarpSmem<DT>);
    FB2_CUDA(cudaFuncSetAttribute(gaussian_pair_combine_warp<DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = ceil_div(npairs, GW_WARPS);
    if (blocks > 148 * 96 / GW_WARPS) blocks = 148 * 96 / GW_WARPS;  // persistent-ish: 8 waves of 12 warps per SM
    gaussian_pair_combine_warp<DT><<<(unsigned)blocks, GW_WARPS * 32, smem, stream>>>(a, npairs);
    FB2_LAUNCH_CHECK();
    return FB2_OK;
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

// eta[b,t,:] = P[b] w[b,t,:],  c[b,t] = s[b,t] - |w[b,t]|^2 / 2   for a time-invariant prec_sqrt P[b] (n x r).
// One thread per (b,t): P[b] sits in shared memory (rows padded to a multiple of 4, read as broadcast LDS.128), the thread
// streams its w row four entries at a time and keeps all n outputs in registers; a warp reads / writes contiguous
// 32 r / 32 n floats.  (The first version -- one warp per element with per-lane strided reads of P -- took 38.6 ms of the
// 47 ms of C4; this one is bandwidth-bound.)
constexpr int ETA_THREADS = 128;
template <int NMAX>
__global__ void __launch_bounds__(ETA_THREADS) gaussian_eta_c(const float* w, const float* P, const float* s, int64_t B, int64_t T, int n, int r,
                                                              float* eta, float* c) {
    extern __shared__ __align__(16) float eta_sP[];  // [n][r4]
    const int r4 = (r + 3) / 4 * 4;
    const int64_t b = blockIdx.y;
    const float* Pb = P + b * (int64_t)n * r;
    for (int idx = threadIdx.x; idx < n * r4; idx += ETA_THREADS) {
        const int i = idx / r4, k = idx % r4;
        eta_sP[idx] = k < r ? Pb[(int64_t)i * r + k] : 0.f;
    }
    __syncthreads();
    const int64_t t = (int64_t)blockIdx.x * ETA_THREADS + threadIdx.x;
    if (t >= T) return;
    const int64_t e = b * T + t;
    const float* we = w + e * r;
    const bool vec = (r % 4 == 0) && (reinterpret_cast<uintptr_t>(w) % 16 == 0);
    float acc[NMAX];
#pragma unroll
    for (int i = 0; i < NMAX; ++i) acc[i] = 0.f;
    float ww = 0.f;
    for (int k0 = 0; k0 < r4; k0 += 4) {
        float4 wv;
        if (vec) {
            wv = __ldg(reinterpret_cast<const float4*>(we + k0));
        } else {
            wv.x = k0 + 0 < r ? __ldg(we + k0 + 0) : 0.f;
            wv.y = k0 + 1 < r ? __ldg(we + k0 + 1) : 0.f;
            wv.z = k0 + 2 < r ? __ldg(we + k0 + 2) : 0.f;
            wv.w = k0 + 3 < r ? __ldg(we + k0 + 3) : 0.f;
        }
        ww = fmaf(wv.x, wv.x, fmaf(wv.y, wv.y, fmaf(wv.z, wv.z, fmaf(wv.w, wv.w, ww))));
#pragma unroll
        for (int i = 0; i < NMAX; ++i) {
            if (i < n) {
                const float4 p = *reinterpret_cast<const float4*>(&eta_sP[i * r4 + k0]);
                acc[i] = fmaf(p.x, wv.x, fmaf(p.y, wv.y, fmaf(p.z, wv.z, fmaf(p.w, wv.w, acc[i]))));
            }
        }
    }
    float* eo = eta + e * n;
    if (n % 4 == 0 && reinterpret_cast<uintptr_t>(eta) % 16 == 0) {
#pragma unroll
        for (int i = 0; i < NMAX; i += 4)
            if (i < n) *reinterpret_cast<float4*>(eo + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
    } else {
#pragma unroll
        for (int i = 0; i < NMAX; ++i)
            if (i < n) eo[i] = acc[i];
    }
    c[e] = (s ? s[e] : 0.f) - 0.5f * ww;
}

// copy one strided element per outer index: dst[q0*o_s0 + slot] = src[q0*s_s0 + which]
__global__ void gaussian_copy_elements(const float* L, const float* e, const float* c, int64_t s_s0, int64_t which, float* Lo,
                                       float* eo, float* co, int64_t o_s0, int64_t slot, int64_t B, int n, int64_t l_s0 = -1,
                                       int64_t l_which = 0) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int per = n * n + n + 1;
    if (idx >= B * per) return;
    int64_t b = idx / per;
    int o = (int)(idx % per);
    int64_t se = b * s_s0 + which, de = b * o_s0 + slot;
    const int64_t le = l_s0 < 0 ? se : b * l_s0 + l_which;
    if (o < n * n) Lo[de * n * n + o] = L[le * n * n + o];
    else if (o < n * n + n) eo[de * n + (o - n * n)] = e[se * n + (o - n * n)];
    else co[de] = c[se];
}

__global__ void gaussian_poison_if_flag(const int32_t* flag, float* co, int64_t B) {
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B && *flag) co[b] = __int_as_float(0x7fc00000);
}

static int launch_pair(const float* Lx, const float* ex, const float* cx, int64_t x_s0, int64_t x_s1, const float* Ly,
                       const float* ey, const float* cy, int64_t y_s0, int64_t y_s1, int64_t n0, int64_t n1, int D,
                       float* Lo, float* eo, float* co, int64_t o_s0, int32_t* flag, cudaStream_t stream, bool lam_time_invariant = false) {
    if (n0 * n1 == 0) return FB2_OK;
    FB2_REQUIRE(n0 * n1 < (1ll << 31), "gaussian_pair_combine: too many pairs for one launch");
    GaussPairArgs a;
    a.Lx = Lx; a.ex = ex; a.cx = cx; a.Ly = Ly; a.ey = ey; a.cy = cy; a.Lo = Lo; a.eo = eo; a.co = co; a.flag = flag;
    a.n1 = n1; a.x_s0 = x_s0; a.x_s1 = x_s1; a.y_s0 = y_s0; a.y_s1 = y_s1; a.o_s0 = o_s0; a.D = D;
    a.lx_s0 = x_s0; a.lx_s1 = x_s1; a.ly_s0 = y_s0; a.ly_s1 = y_s1;
    if (lam_time_invariant) {  // Lam[b] shared by every time step of chain b: both operands read the same matrix
        a.lx_s0 = a.ly_s0 = 1;
        a.lx_s1 = a.ly_s1 = 0;
    }
    const char* old_env = getenv("FB2_GAUSS_CTA");  // A/B switch: the CTA-per-pair kernel
    if (old_env && old_env[0] == '1') {
        gaussian_pair_combine<<<(unsigned)(n0 * n1), 256, 0, stream>>>(a);
        FB2_LAUNCH_CHECK();
        return FB2_OK;
    }
    const int64_t npairs = n0 * n1;
    if (D <= 4) return launch_pair_warp<4>(a, npairs, stream);
    if (D <= 8) return launch_pair_warp<8>(a, npairs, stream);
    if (D <= 16) return launch_pair_warp<16>(a, npairs, stream);
    return launch_pair_warp<32>(a, npairs, stream);
}


// ---------------------------------------------------------------------------------------------
// Time-homogeneous chains (GaussianHMM with fixed dynamics, pyro/hmm.py:258-270: only white_vec depends on t).
// Every element of a level carries the SAME precision matrix per chain -- except the odd leftover of the reference's tree
// (sum_product.py:696-698), which travels at the end of the sequence as a "tail" with an older precision.  So the O(D^3) part of
// a pair contraction (Cholesky of M, U = L^-1 G', Lam' = base - U'U) is needed once per level for the (regular, regular) pair
// and at most once for the (regular, tail) pair; for all other pairs only eta and c change:
//     z = eta_x[b] + eta_y[a],  v = L^-1 z,  eta' = [eta_x[a]; eta_y[c]] - U' v,  c' = c_x + c_y + const + |v|^2 / 2
// which is O(D^2) flops and ~12 D bytes per pair.  At C4 (T=100000, B=64, D=32) this replaces 6.4e6 Schur complements by
// 17 + <=17 of them per chain plus a bandwidth-bound sweep over eta.
//
// gaussian_homog_levels: one CTA per chain, warp 0 = the regular pair of each level, warp 1 = the tail pair; writes the
// per-level operators (U, L^-1, const) and the level precisions (every buffer is written once and read afterwards, so the
// non-coherent loads of the pair body never see a stale line).
struct HomogLevelsArgs {
    const float* Lam;  // [B, n, n]
    float* Lreg;       // [B, NL, n, n]
    float* Ltail;      // [B, NL, n, n]
    float* opU;        // [B, NL, 2, DT, 2 DT]
    float* opLinv;     // [B, NL, 2, DT, DT]
    float* konst;      // [B, NL, 2]
    const float* zeros;  // [n]
    float* scratch;      // [B, 2, n] unused eta outputs
    float* Lo;           // [B, n, n] precision of the result
    int32_t* flag;
    int64_t T;
    int D, NL;
    int l_begin, l_end;  // levels [l_begin, l_end) are computed by this launch (earlier ones are replayed for their bookkeeping only)
};

template <int DT>
__global__ void __launch_bounds__(64) gaussian_homog_levels(HomogLevelsArgs a) {
    __shared__ __align__(16) GaussWarpSmem<DT> sm2[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t b = blockIdx.x;
    const int D = a.D, n = 2 * D;
    const size_t nn = (size_t)n * n;
    const float* cur_reg = a.Lam + b * nn;
    const float* cur_tail = nullptr;
    bool has_tail = false;
    int64_t d = a.T;
    for (int l = 0; d > 1 && l < a.l_end; ++l) {
        const int64_t npairs = d / 2;
        const bool odd = d & 1;
        const bool tail_pair = has_tail && !odd;
        const int64_t reg_pairs = npairs - (tail_pair ? 1 : 0);
        float* next_reg = a.Lreg + (b * a.NL + l) * nn;
        float* next_tail = a.Ltail + (b * a.NL + l) * nn;
        const size_t op = ((size_t)b * a.NL + l) * 2 + warp;
        if (l >= a.l_begin && ((warp == 0 && reg_pairs > 0) || (warp == 1 && tail_pair))) {
            gauss_pair_body<DT, false, true>(cur_reg, warp == 0 ? cur_reg : cur_tail, a.zeros, a.zeros, 0.f, 0.f, D, warp == 0 ? next_reg : next_tail,
                                             a.scratch + (b * 2 + warp) * n, a.konst + op, a.flag, sm2[warp], lane, a.opU + op * DT * 2 * DT,
                                             a.opLinv + op * DT * DT);
        }
        __syncthreads();
        if (tail_pair) {
            cur_tail = next_tail;
        } else if (!has_tail && odd) {
            cur_tail = cur_reg;
            has_tail = true;
        }
        cur_reg = next_reg;
        d = (d + 1) / 2;
    }
    if (d > 1) return;  // the launch that reaches the top of the tree writes the result
    const float* res = has_tail ? cur_tail : cur_reg;
    for (size_t i = threadIdx.x; i < nn; i += blockDim.x) a.Lo[b * nn + i] = res[i];
}

// gaussian_homog_levels_cta: the same chain of per-level pair contractions with a whole CTA (256 threads) per chain and the level
// precisions resident in shared memory.  The warp-wide version above spends ~25 us per level (one warp walks the O(D^3) work of a pair,
// and every level goes through global memory): 0.43 ms of BASELINE config 4, a quarter of the whole scan once the eta work moved to the
// tensor pipe.  Here warp 0 factorises M (the same shuffle Cholesky, so the operators stay bit-identical), warps 0-1 run the forward
// substitution for the 2 DT right-hand-side columns while warp 2 inverts L, and all 8 warps share Lam' = base - U'U (16 outputs per
// thread, U broadcast from shared memory).  Same operation order per output as gauss_pair_body.
template <int DT>
struct HomogLevelsSmem {
    float Lam[4][2 * DT][2 * DT + 1];  // padded layout (a-block rows/cols [0, DT), c-block [DT, 2 DT)): cur regular, cur tail, next regular, next tail
    float L[DT][DT + 1];               // Cholesky factor, diagonal holds 1 / L[k][k]
    float U[DT][2 * DT + 4];           // U = L^-1 G'
    float col[2][DT < 4 ? 4 : DT];     // column j of L as the warp computes it (broadcast through shared memory instead of 31 - j shuffles)
    float logdet;
    int bad;
};
template <int DT>
__device__ void homog_pair_cta(const float (*X)[2 * DT + 1], const float (*Y)[2 * DT + 1], float (*Z)[2 * DT + 1], HomogLevelsSmem<DT>& sm, int D,
                               float* opU, float* opLinv, float* konst, int32_t* flag) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // ---- phase 1 (warp 0): Cholesky of M = X_cc + Y_aa, lane = row.  Column j of L goes to the other lanes through shared memory: the
    //      31 - j shuffles per column of the warp-wide kernel are shuffle-throughput bound (8.4k cycles per pair when every warp ran them,
    //      7.8-13k inside a divergent region, where ptxas wraps each shuffle in a WARPSYNC.COLLECTIVE sequence); same arithmetic.
    if (warp == 0) {
        float m[DT];
#pragma unroll
        for (int k = 0; k < DT; ++k) m[k] = (lane < D && k < D) ? X[DT + k][DT + lane] + Y[k][lane] : (k == lane ? 1.f : 0.f);
        float logdet = 0.f, dinv = 1.f;
        int bad = 0;
#pragma unroll
        for (int j = 0; j < DT; ++j) {
            float* colj = sm.col[j & 1];
            if (lane == j) colj[0] = m[j];  // slot 0 first carries the pivot
            __syncwarp();
            float d = colj[0];
            __syncwarp();
            if (!(d > 0.f)) {
                bad = 1;
                d = 1.f;
            }
            float inv = rsqrtf(d);
            inv = inv * fmaf(-0.5f * d * inv, inv, 1.5f);
            const float ljj = d * inv;
            logdet += __logf(ljj);
            if (lane == j) dinv = inv;
            const float lij = lane > j ? m[j] * inv : (lane == j ? ljj : 0.f);
            m[j] = lij;
            if (lane < DT) colj[lane] = lij;
            __syncwarp();
#pragma unroll
            for (int k = j + 1; k < DT; ++k) m[k] = fmaf(-lij, colj[k], m[k]);
        }
        bad = __any_sync(0xffffffffu, bad);
        if (lane < DT) {
#pragma unroll
            for (int k = 0; k < DT; ++k) sm.L[lane][k] = m[k];
            sm.L[lane][lane] = dinv;
        }
        if (lane == 0) {
            sm.logdet = logdet;
            if (bad && flag) atomicExch(flag, 1);
        }
    }
    __syncthreads();
    // ---- phase 2: threads 0 .. 2 DT - 1: forward substitution for column c of G' (a-block column cc: X_ca[k][cc], c-block: Y_ac[k][cc]);
    //      the next warp: L^-1 by forward substitution on the identity (lane = column)
    if (tid < 2 * DT) {
        const int c = tid, blk = c / DT, cc = c % DT;
        float u[DT];
#pragma unroll
        for (int k = 0; k < DT; ++k) u[k] = (cc < D && k < D) ? (blk == 0 ? X[DT + k][cc] : Y[k][DT + cc]) : 0.f;
        // right-looking order: once u[k] is final it is subtracted from every later entry -- per entry the same operations in the same
        // order as the row-by-row form of gauss_pair_body (u[i] - L[i][0] u[0] - L[i][1] u[1] - ...), but independent across i
#pragma unroll
        for (int k = 0; k < DT; ++k) {
            u[k] = u[k] * sm.L[k][k];
            sm.U[k][c] = u[k];
            opU[k * 2 * DT + c] = u[k];
#pragma unroll
            for (int i = k + 1; i < DT; ++i) u[i] = fmaf(-sm.L[i][k], u[k], u[i]);
        }
    } else if (warp == (2 * DT + 31) / 32) {
        const int col = lane;
        float x[DT];
#pragma unroll
        for (int k = 0; k < DT; ++k) x[k] = k == col ? 1.f : 0.f;
#pragma unroll
        for (int k = 0; k < DT; ++k) {
            x[k] = x[k] * sm.L[k][k];
#pragma unroll
            for (int i = k + 1; i < DT; ++i) x[i] = fmaf(-sm.L[i][k], x[k], x[i]);
        }
        if (col < DT) {
#pragma unroll
            for (int k = 0; k < DT; ++k) opLinv[k * DT + col] = x[k];
        }
        if (col == 0) *konst = __fsub_rn(__fmul_rn((float)D, kHalfLog2Pi), sm.logdet);  // two roundings, as gauss_pair_body evaluates it (no fused multiply-add)
    }
    __syncthreads();
    // ---- phase 3: Z = diag(X_aa, Y_cc) - U'U over the padded layout, 16 entries of one row per thread (rows of 2 DT = 64: 4 threads each)
    constexpr int PER = (2 * DT) * (2 * DT) / 256 < 1 ? 1 : (2 * DT) * (2 * DT) / 256;  // entries per thread
    constexpr int TPR = 2 * DT / PER;                                                       // threads per row
    if constexpr (DT == 32) {
        // 4 x 4 register tile per thread: two LDS.128 of U feed 16 FMAs per k (a 1 x 16 strip needed five loads); same k order per output
        const int r0 = (tid >> 4) * 4, s0 = (tid & 15) * 4;
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int r = r0 + i, sp = s0 + t, rblk = r / DT, rc = r % DT, sblk = sp / DT, sc = sp % DT;
                acc[i][t] = (rc < D && sc < D && rblk == sblk) ? (rblk == 0 ? X[sc][rc] : Y[DT + sc][DT + rc]) : 0.f;
            }
#pragma unroll
        for (int k = 0; k < DT; ++k) {
            const float4 ur = *reinterpret_cast<const float4*>(&sm.U[k][r0]);
            const float4 us = *reinterpret_cast<const float4*>(&sm.U[k][s0]);
            const float ura[4] = {ur.x, ur.y, ur.z, ur.w}, usa[4] = {us.x, us.y, us.z, us.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int t = 0; t < 4; ++t) acc[i][t] = fmaf(-ura[i], usa[t], acc[i][t]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int t = 0; t < 4; ++t) Z[s0 + t][r0 + i] = acc[i][t];
    } else if (tid < 2 * DT * TPR) {
        const int r = tid / TPR, s0 = (tid % TPR) * PER;
        const int rblk = r / DT, rc = r % DT;
        float acc[PER];
#pragma unroll
        for (int t = 0; t < PER; ++t) {
            const int sp = s0 + t, sblk = sp / DT, sc = sp % DT;
            acc[t] = (rc < D && sc < D && rblk == sblk) ? (rblk == 0 ? X[sc][rc] : Y[DT + sc][DT + rc]) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < DT; ++k) {
            const float ur = sm.U[k][r];
            if constexpr (PER % 4 == 0) {  // rows of U are 16-byte aligned (2 DT + 4 floats) and s0 is a multiple of PER
#pragma unroll
                for (int t = 0; t < PER; t += 4) {
                    const float4 uv = *reinterpret_cast<const float4*>(&sm.U[k][s0 + t]);
                    acc[t] = fmaf(-ur, uv.x, acc[t]);
                    acc[t + 1] = fmaf(-ur, uv.y, acc[t + 1]);
                    acc[t + 2] = fmaf(-ur, uv.z, acc[t + 2]);
                    acc[t + 3] = fmaf(-ur, uv.w, acc[t + 3]);
                }
            } else {
#pragma unroll
                for (int t = 0; t < PER; ++t) acc[t] = fmaf(-ur, sm.U[k][s0 + t], acc[t]);
            }
        }
#pragma unroll
        for (int t = 0; t < PER; ++t) Z[s0 + t][r] = acc[t];  // written through the transpose like gauss_pair_body (the result is symmetric up to rounding)
    }
    __syncthreads();
}

template <int DT>
__global__ void __launch_bounds__(256) gaussian_homog_levels_cta(HomogLevelsArgs a) {
    extern __shared__ __align__(16) unsigned char hl_raw[];
    HomogLevelsSmem<DT>& sm = *reinterpret_cast<HomogLevelsSmem<DT>*>(hl_raw);
    const int64_t b = blockIdx.x;
    const int D = a.D, n = 2 * D, tid = threadIdx.x;
    // Lam[b] -> padded layout, slot 0
    for (int idx = tid; idx < 2 * DT * 2 * DT; idx += blockDim.x) {
        const int r = idx / (2 * DT), c = idx % (2 * DT);
        const int rb = r / DT, rc = r % DT, cb = c / DT, cc = c % DT;
        sm.Lam[0][r][c] = (rc < D && cc < D) ? a.Lam[b * n * n + (rb * D + rc) * n + cb * D + cc] : 0.f;
    }
    __syncthreads();
    int reg = 0, tail = -1;  // slots of the current regular / tail precision
    bool has_tail = false;
    int64_t d = a.T;
    for (int l = 0; d > 1; ++l) {
        const int64_t npairs = d / 2;
        const bool odd = d & 1;
        const bool tail_pair = has_tail && !odd;
        const int64_t reg_pairs = npairs - (tail_pair ? 1 : 0);
        // two free slots for this level's results
        int free_slots[2], nf = 0;
        for (int q = 0; q < 4 && nf < 2; ++q)
            if (q != reg && q != tail) free_slots[nf++] = q;
        const size_t op = ((size_t)b * a.NL + l) * 2;
        if (reg_pairs > 0)
            homog_pair_cta<DT>(sm.Lam[reg], sm.Lam[reg], sm.Lam[free_slots[0]], sm, D, a.opU + op * DT * 2 * DT, a.opLinv + op * DT * DT, a.konst + op, a.flag);
        if (tail_pair)
            homog_pair_cta<DT>(sm.Lam[reg], sm.Lam[tail], sm.Lam[free_slots[1]], sm, D, a.opU + (op + 1) * DT * 2 * DT, a.opLinv + (op + 1) * DT * DT,
                               a.konst + op + 1, a.flag);
        if (tail_pair) {
            tail = free_slots[1];
        } else if (!has_tail && odd) {
            tail = reg;
            has_tail = true;
        }
        if (reg_pairs > 0) reg = free_slots[0];
        d = (d + 1) / 2;
    }
    const int res = has_tail ? tail : reg;
    for (int idx = tid; idx < n * n; idx += blockDim.x) {
        const int r = idx / n, c = idx % n;
        a.Lo[b * n * n + idx] = sm.Lam[res][(r / D) * DT + r % D][(c / D) * DT + c % D];
    }
}

// eta / c update of `count` pairs per chain starting at pair index p0, all with the operator (opU, opLinv, konst) of their
// chain; one warp per pair, the operator lives in registers.  `leftover`: also copy input element d-1 to output slot `lslot`.
struct HomogEtaArgs {
    const float* ein;  // [B, in_T, n]
    const float* cin;  // [B, in_T]
    float* eout;       // [B, out_T, n]
    float* cout;       // [B, out_T]
    const float* opU;  // + b * op_stride * DT * 2 DT
    const float* opLinv;
    const float* konst;
    int64_t op_stride;  // operators per chain (= 2 NL)
    int64_t in_T, out_T, p0, count;
    int64_t leftover_src, leftover_dst;  // -1: none
    int64_t out_p_off;  // pair p of the input range is written to output slot p + out_p_off (0 except for the remainder tile of the fused Kalman path)
    int D;
};
constexpr int HE_THREADS = 128;

// One THREAD per pair: the operator of the chain (L^-1: DT x DT, U: DT x 2DT) sits in shared memory and is read as broadcast
// LDS.128; z, v and the 2 DT outputs live in registers.  3 DT^2 FMAs and 3 DT^2 / 4 shared loads per pair, no shuffles (a
// warp-per-pair version with the operator in registers spent 64 shuffles per pair and took 6.7 ms over the levels of C4).
// A thread reads the 2 n contiguous floats of its pair; neighbouring threads are 2 n floats apart, the lines are shared through L1.
// eta / c update of one pair (elements 2p, 2p + 1 of `ein` -> element `po` of `eout`) with the operator in shared memory
// NC: the inputs are read-only for the whole kernel (non-coherent loads); otherwise they were written earlier by this kernel (L2 loads).
template <int DT, bool NC>
__device__ __forceinline__ void homog_eta_pair(const float* ex, float cxy, float* eo, float* co, const float* sLi, const float* sU, float konst, int D,
                                               bool vec) {
    auto ld4 = [](const float* p) { return NC ? __ldg(reinterpret_cast<const float4*>(p)) : __ldcg(reinterpret_cast<const float4*>(p)); };
    auto ld1 = [](const float* p) { return NC ? __ldg(p) : __ldcg(p); };
    const int n = 2 * D;
    const float* ey = ex + n;
    float z[DT], out[2 * DT];
    if (vec) {
#pragma unroll
        for (int k = 0; k < DT; k += 4) {
            if (k < D) {
                const float4 xa = ld4(ex + k), xb = ld4(ex + D + k);
                const float4 ya = ld4(ey + k), yc = ld4(ey + D + k);
                z[k] = xb.x + ya.x; z[k + 1] = xb.y + ya.y; z[k + 2] = xb.z + ya.z; z[k + 3] = xb.w + ya.w;
                out[k] = xa.x; out[k + 1] = xa.y; out[k + 2] = xa.z; out[k + 3] = xa.w;
                out[DT + k] = yc.x; out[DT + k + 1] = yc.y; out[DT + k + 2] = yc.z; out[DT + k + 3] = yc.w;
            } else {
                z[k] = z[k + 1] = z[k + 2] = z[k + 3] = 0.f;
                out[k] = out[k + 1] = out[k + 2] = out[k + 3] = 0.f;
                out[DT + k] = out[DT + k + 1] = out[DT + k + 2] = out[DT + k + 3] = 0.f;
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < DT; ++k) {
            z[k] = k < D ? ld1(ex + D + k) + ld1(ey + k) : 0.f;
            out[k] = k < D ? ld1(ex + k) : 0.f;
            out[DT + k] = k < D ? ld1(ey + D + k) : 0.f;
        }
    }
    float vv = 0.f;
#pragma unroll
    for (int r = 0; r < DT; ++r) {
        // v_r = (L^-1 z)_r; L^-1 is lower triangular: columns beyond r are zero (and beyond D the padding is the identity)
        float v = 0.f;
#pragma unroll
        for (int k = 0; k < DT; k += 4) {
            if (k <= r) {
                const float4 l = *reinterpret_cast<const float4*>(&sLi[r * DT + k]);
                v = fmaf(l.x, z[k], fmaf(l.y, z[k + 1], fmaf(l.z, z[k + 2], fmaf(l.w, z[k + 3], v))));
            }
        }
        vv = fmaf(v, v, vv);
#pragma unroll
        for (int c = 0; c < 2 * DT; c += 4) {
            const float4 u = *reinterpret_cast<const float4*>(&sU[r * 2 * DT + c]);
            out[c] = fmaf(-u.x, v, out[c]);
            out[c + 1] = fmaf(-u.y, v, out[c + 1]);
            out[c + 2] = fmaf(-u.z, v, out[c + 2]);
            out[c + 3] = fmaf(-u.w, v, out[c + 3]);
        }
    }
    if (vec) {
#pragma unroll
        for (int k = 0; k < DT; k += 4) {
            if (k < D) {
                *reinterpret_cast<float4*>(eo + k) = make_float4(out[k], out[k + 1], out[k + 2], out[k + 3]);
                *reinterpret_cast<float4*>(eo + D + k) = make_float4(out[DT + k], out[DT + k + 1], out[DT + k + 2], out[DT + k + 3]);
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < DT; ++k) {
            if (k < D) {
                eo[k] = out[k];
                eo[D + k] = out[DT + k];
            }
        }
    }
    *co = cxy + konst + 0.5f * vv;
}

template <int DT>
__global__ void __launch_bounds__(HE_THREADS) gaussian_homog_eta(HomogEtaArgs a) {
    __shared__ __align__(16) float sLi[DT * DT];
    __shared__ __align__(16) float sU[DT * 2 * DT];
    const int64_t b = blockIdx.y;
    const int D = a.D, n = 2 * D;
    {
        const float* U = a.opU + b * a.op_stride * DT * 2 * DT;
        const float* Li = a.opLinv + b * a.op_stride * DT * DT;
        for (int i = threadIdx.x; i < DT * DT; i += HE_THREADS) sLi[i] = __ldg(Li + i);
        for (int i = threadIdx.x; i < DT * 2 * DT; i += HE_THREADS) sU[i] = __ldg(U + i);
    }
    __syncthreads();
    const float konst = a.konst[b * a.op_stride];
    const float* ein = a.ein + b * a.in_T * n;
    const float* cin = a.cin + b * a.in_T;
    float* eout = a.eout + b * a.out_T * n;
    float* cout = a.cout + b * a.out_T;
    const bool vec = (D % 4 == 0) && (reinterpret_cast<uintptr_t>(a.ein) % 16 == 0) && (reinterpret_cast<uintptr_t>(a.eout) % 16 == 0);
    for (int64_t i = (int64_t)blockIdx.x * HE_THREADS + threadIdx.x; i < a.count; i += (int64_t)gridDim.x * HE_THREADS) {
        const int64_t p = a.p0 + i;
        homog_eta_pair<DT, true>(ein + 2 * p * n, cin[2 * p] + cin[2 * p + 1], eout + (p + a.out_p_off) * n, cout + p + a.out_p_off, sLi, sU, konst, D, vec);
    }
    if (a.leftover_src >= 0 && blockIdx.x == 0 && threadIdx.x < 32) {
        const int lane = threadIdx.x;
        for (int i = lane; i < n; i += 32) eout[a.leftover_dst * n + i] = ein[a.leftover_src * n + i];
        if (lane == 0) cout[a.leftover_dst] = cin[a.leftover_src];
    }
}

// gaussian_homog_upper: every level from `l0` to the top of the tree in ONE launch, one CTA per chain (the levels above the tile depth
// hold a few thousand elements per chain and shrink by two each time: as 16 separate launches they were launch-latency bound, ~0.3 ms of
// BASELINE config 4).  Same per-pair arithmetic (homog_eta_pair) and the same regular / tail / leftover bookkeeping as the host loop of
// homog_chain_run; the elements ping-pong between two global buffers (L2 resident), the level's operator is reloaded into shared memory.
struct HomogUpperArgs {
    const float* e0;  // level-l0 elements [B, d0, n], [B, d0]
    const float* c0;
    float *eb[2], *cb[2];  // ping-pong buffers: capacity (d0 + 1) / 2 and ((d0 + 1) / 2 + 1) / 2 elements per chain
    float *eo, *co;        // chain result [B, n], [B]
    const float *opU, *opLinv, *konst;
    int64_t d0, cap[2];
    int l0, NL, D, has_tail;
};
constexpr int HU_THREADS = 256;
template <int DT>
__global__ void __launch_bounds__(HU_THREADS) gaussian_homog_upper(HomogUpperArgs a) {
    __shared__ __align__(16) float sLi[2][DT * DT];
    __shared__ __align__(16) float sU[2][DT * 2 * DT];
    __shared__ float sk[2];
    const int64_t b = blockIdx.x;
    const int D = a.D, n = 2 * D, tid = threadIdx.x;
    const float* ein = a.e0 + b * a.d0 * n;
    const float* cin = a.c0 + b * a.d0;
    int64_t d = a.d0;
    bool has_tail = a.has_tail != 0;
    int which = 0;
    for (int l = a.l0; d > 1; ++l) {
        const int64_t npairs = d / 2, nd = (d + 1) / 2;
        const bool odd = d & 1;
        const bool tail_pair = has_tail && !odd;
        const int64_t reg_pairs = npairs - (tail_pair ? 1 : 0);
        float* eout = nd == 1 ? a.eo + b * n : a.eb[which] + b * a.cap[which] * n;
        float* cout = nd == 1 ? a.co + b : a.cb[which] + b * a.cap[which];
        __syncthreads();  // the previous level's operator and results are no longer in use / are complete
        for (int w = 0; w < 2; ++w) {
            if ((w == 0 && reg_pairs > 0) || (w == 1 && tail_pair)) {
                const size_t op = ((size_t)b * a.NL + l) * 2 + w;
                const float4* Li = reinterpret_cast<const float4*>(a.opLinv + op * DT * DT);
                const float4* U = reinterpret_cast<const float4*>(a.opU + op * DT * 2 * DT);
                for (int i = tid; i < DT * DT / 4; i += HU_THREADS) reinterpret_cast<float4*>(sLi[w])[i] = __ldg(Li + i);
                for (int i = tid; i < DT * 2 * DT / 4; i += HU_THREADS) reinterpret_cast<float4*>(sU[w])[i] = __ldg(U + i);
                if (tid == 0) sk[w] = a.konst[op];
            }
        }
        __syncthreads();
        const bool vec = (D % 4 == 0) && (reinterpret_cast<uintptr_t>(ein) % 16 == 0) && (reinterpret_cast<uintptr_t>(eout) % 16 == 0);
        for (int64_t p = tid; p < reg_pairs; p += HU_THREADS)
            homog_eta_pair<DT, false>(ein + 2 * p * n, __ldcg(cin + 2 * p) + __ldcg(cin + 2 * p + 1), eout + p * n, cout + p, sLi[0], sU[0], sk[0], D, vec);
        if (tail_pair && tid == HU_THREADS - 1) {
            const int64_t p = npairs - 1;
            homog_eta_pair<DT, false>(ein + 2 * p * n, __ldcg(cin + 2 * p) + __ldcg(cin + 2 * p + 1), eout + p * n, cout + p, sLi[1], sU[1], sk[1], D, vec);
        }
        if (odd && tid < 32) {  // the leftover passes through
            for (int i = tid; i < n; i += 32) eout[(nd - 1) * n + i] = __ldcg(ein + (d - 1) * n + i);
            if (tid == 0) cout[nd - 1] = __ldcg(cin + d - 1);
        }
        if (!tail_pair && !has_tail && odd) has_tail = true;
        __threadfence_block();
        ein = eout; cin = cout; d = nd;
        which ^= 1;
    }
}

template <int DT>
static int launch_homog_levels(HomogLevelsArgs& la, int64_t B, cudaStream_t stream) {
    const char* warp_env = getenv("FB2_GAUSS_LEVELS_WARP");  // 1: the warp-per-pair form (one warp walks a level's pair contraction)
    if ((warp_env && warp_env[0] == '1') || la.l_begin > 0 || la.l_end <= la.NL) {
        gaussian_homog_levels<DT><<<(unsigned)B, 64, 0, stream>>>(la);
    } else {
        const size_t smem = sizeof(HomogLevelsSmem<DT>);
        FB2_CUDA(cudaFuncSetAttribute(gaussian_homog_levels_cta<DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gaussian_homog_levels_cta<DT><<<(unsigned)B, 256, smem, stream>>>(la);
    }
    FB2_LAUNCH_CHECK();
    return FB2_OK;
}

static int homog_levels_count(int64_t T) {
    int nl = 0;
    for (int64_t d = T; d > 1; d = (d + 1) / 2) ++nl;
    return nl;
}
// Workspace layout of the homogeneous chain.  keep = every level's eta is kept (levels 1 .. NL-1 in their own buffers) so that
// gaussian_homog_sample can walk the tree back down; otherwise two ping-pong buffers.
struct HomogLayout {
    float *Lreg, *Ltail, *opU, *opLinv, *konst, *zeros, *scratch;
    int32_t* flag;
    float* be[2];
    float* bc[2];
    float* lev_e[48];  // lev_e[l] = eta entering level l (l >= 1); lev_e[0] is the caller's array
    bool ok;
};
static HomogLayout homog_layout(void* workspace, size_t workspace_bytes, int64_t B, int64_t T, int D, int DT, bool keep) {
    const int n = 2 * D, NL = homog_levels_count(T);
    const int64_t t1 = (T + 1) / 2, t2 = (t1 + 1) / 2;
    Arena arena(workspace, workspace_bytes);
    HomogLayout L{};
    L.Lreg = arena.take<float>((size_t)B * NL * n * n);
    L.Ltail = arena.take<float>((size_t)B * NL * n * n);
    L.opU = arena.take<float>((size_t)B * NL * 2 * DT * 2 * DT);
    L.opLinv = arena.take<float>((size_t)B * NL * 2 * DT * DT);
    L.konst = arena.take<float>((size_t)B * NL * 2);
    L.zeros = arena.take<float>((size_t)n);
    L.scratch = arena.take<float>((size_t)B * 2 * n);
    L.flag = arena.take<int32_t>(1);
    const int64_t tt[2] = {t1, t2};
    for (int q = 0; q < 2; ++q) {
        L.be[q] = keep ? nullptr : arena.take<float>((size_t)B * tt[q] * n);
        L.bc[q] = arena.take<float>((size_t)B * tt[q]);
    }
    if (keep) {
        int64_t d = T;
        for (int l = 1; l < NL && l < 48; ++l) {
            d = (d + 1) / 2;
            L.lev_e[l] = arena.take<float>((size_t)B * d * n);
        }
    }
    L.ok = arena.ok() && NL < 48;
    return L;
}
static size_t homog_workspace_bytes(int64_t B, int64_t T, int D, bool keep = false) {
    const int n = 2 * D, NL = homog_levels_count(T);
    const int DT = D <= 4 ? 4 : D <= 8 ? 8 : D <= 16 ? 16 : 32;
    const int64_t t1 = (T + 1) / 2, t2 = (t1 + 1) / 2;
    size_t w = 2 * align_up((size_t)B * NL * n * n * 4) + align_up((size_t)B * NL * 2 * DT * 2 * DT * 4) + align_up((size_t)B * NL * 2 * DT * DT * 4) +
               align_up((size_t)B * NL * 2 * 4) + align_up((size_t)n * 4) + align_up((size_t)B * 2 * n * 4) + align_up(4);
    for (int64_t tt : {t1, t2}) w += (keep ? 0 : align_up((size_t)B * tt * n * 4)) + align_up((size_t)B * tt * 4);
    if (keep) {
        int64_t d = T;
        for (int l = 1; l < NL; ++l) {
            d = (d + 1) / 2;
            w += align_up((size_t)B * d * n * 4);
        }
    }
    return w + 256;
}

template <int DT>
static int homog_chain_run(const float* Lam, const float* eta, const float* c, int64_t B, int64_t T, int D, float* Lo, float* eo, float* co,
                           void* workspace, size_t workspace_bytes, cudaStream_t stream, bool keep = false) {
    const int NL = homog_levels_count(T);
    const HomogLayout lay = homog_layout(workspace, workspace_bytes, B, T, D, DT, keep);
    const int n = 2 * D;
    HomogLevelsArgs la;
    la.Lam = Lam; la.T = T; la.D = D; la.NL = NL; la.Lo = Lo; la.l_begin = 0; la.l_end = NL + 1;
    la.Lreg = lay.Lreg; la.Ltail = lay.Ltail; la.opU = lay.opU; la.opLinv = lay.opLinv; la.konst = lay.konst;
    float* zeros = lay.zeros;
    la.zeros = zeros;
    la.scratch = lay.scratch;
    la.flag = lay.flag;
    if (!lay.ok) {
        set_error("gaussian_chain (homogeneous): workspace too small (%zu bytes)", workspace_bytes);
        return FB2_ERR_WORKSPACE;
    }
    FB2_REQUIRE(B <= 65535, "gaussian_chain (homogeneous): at most 65535 chains per call");
    FB2_CUDA(cudaMemsetAsync(la.flag, 0, sizeof(int32_t), stream));
    FB2_CUDA(cudaMemsetAsync(zeros, 0, (size_t)n * 4, stream));
    {
        const int lst = launch_homog_levels<DT>(la, B, stream);
        if (lst != FB2_OK) return lst;
    }
    const float *ce = eta, *cc = c;
    int64_t d = T, in_T = T;
    bool has_tail = false;
    int which = 0;
    const char* up_env = getenv("FB2_KALMAN_UPPER");
    const bool upper_ok = !keep && !(up_env && up_env[0] == '0');
    for (int l = 0; d > 1; ++l) {
        // once a level is small, the rest of the tree runs in one launch (gaussian_homog_upper, one CTA per chain).  The current elements
        // sit in the caller's array (l = 0) or in be[which ^ 1] with stride d; odd levels from here go to be[which], even ones back over
        // the chain's own region of be[which ^ 1]
        if (upper_ok && d <= 1024) {
            HomogUpperArgs ua;
            ua.e0 = ce; ua.c0 = cc; ua.d0 = d;
            ua.eb[0] = lay.be[which]; ua.cb[0] = lay.bc[which]; ua.cap[0] = (d + 1) / 2;
            ua.eb[1] = lay.be[which ^ 1]; ua.cb[1] = lay.bc[which ^ 1]; ua.cap[1] = l == 0 ? ((d + 1) / 2 + 1) / 2 : d;
            ua.eo = eo; ua.co = co; ua.opU = la.opU; ua.opLinv = la.opLinv; ua.konst = la.konst;
            ua.l0 = l; ua.NL = NL; ua.D = D; ua.has_tail = has_tail ? 1 : 0;
            gaussian_homog_upper<DT><<<(unsigned)B, HU_THREADS, 0, stream>>>(ua);
            FB2_LAUNCH_CHECK();
            break;
        }
        const int64_t npairs = d / 2, nd = (d + 1) / 2;
        const bool odd = d & 1;
        const bool tail_pair = has_tail && !odd;
        const int64_t reg_pairs = npairs - (tail_pair ? 1 : 0);
        float* de = nd == 1 ? eo : (keep ? lay.lev_e[l + 1] : lay.be[which]);
        float* dc = nd == 1 ? co : lay.bc[which];
        HomogEtaArgs ea;
        ea.ein = ce; ea.cin = cc; ea.eout = de; ea.cout = dc; ea.in_T = in_T; ea.out_T = nd; ea.D = D; ea.op_stride = (int64_t)NL * 2;
        ea.leftover_src = -1; ea.leftover_dst = -1; ea.out_p_off = 0;
        if (reg_pairs > 0) {
            ea.opU = la.opU + (size_t)(l * 2 + 0) * DT * 2 * DT;
            ea.opLinv = la.opLinv + (size_t)(l * 2 + 0) * DT * DT;
            ea.konst = la.konst + (size_t)(l * 2 + 0);
            ea.p0 = 0; ea.count = reg_pairs;
            if (odd) {
                ea.leftover_src = d - 1;
                ea.leftover_dst = nd - 1;
            }
            int64_t blocks = ceil_div(reg_pairs, HE_THREADS);
            if (blocks > 65535) blocks = 65535;
            gaussian_homog_eta<DT><<<dim3((unsigned)blocks, (unsigned)B), HE_THREADS, 0, stream>>>(ea);
            FB2_LAUNCH_CHECK();
        }
        if (tail_pair) {
            ea.opU = la.opU + (size_t)(l * 2 + 1) * DT * 2 * DT;
            ea.opLinv = la.opLinv + (size_t)(l * 2 + 1) * DT * DT;
            ea.konst = la.konst + (size_t)(l * 2 + 1);
            ea.p0 = npairs - 1; ea.count = 1;
            ea.leftover_src = -1; ea.leftover_dst = -1;
            gaussian_homog_eta<DT><<<dim3(1, (unsigned)B), HE_THREADS, 0, stream>>>(ea);
            FB2_LAUNCH_CHECK();
        }
        if (!tail_pair && !has_tail && odd) has_tail = true;
        ce = de; cc = dc; in_T = nd; d = nd;
        which ^= 1;
    }
    gaussian_poison_if_flag<<<(unsigned)ceil_div(B, 256), 256, 0, stream>>>(la.flag, co, B);
    FB2_LAUNCH_CHECK();
    return FB2_OK;
}

// ---------------------------------------------------------------------------------------------
// Backward sampling through the same tree (forward-filter backward-sample for Gaussian chains: recipes.py:16-90 runs the adjoint of
// the scan under MonteCarlo, montecarlo.py:48-61, and Gaussian._sample gaussian.py:946-1059 draws each eliminated variable given its
// two neighbours).  The variable eliminated by pair p of level l is the state at time m = (2p+1) 2^l, its neighbours are the states at
// s = 2p 2^l and e = min((2p+2) 2^l, T).  With M = L L' the pair's shared-block precision, U = L^-1 G', v = L^-1 (eta_x[b] + eta_y[a])
// (exactly the operators the forward sweep used),   x_m | x_s, x_e  ~  N(M^-1 (z - G'[x_s; x_e]), M^-1),   drawn as
//     x_m = L^-T (eps_m + v - U [x_s; x_e])
// with caller-supplied white noise eps: the draw is a deterministic function of (inputs, noise), which is what lets the oracle
// check the joint law exactly.  Levels are walked top-down, all pairs and samples of a level in parallel; one thread per (pair, sample).
struct HomogSampleArgs {
    const float* e_lvl;  // eta entering level l: [B, d_l, n]
    const float* opU;
    const float* opLinv;
    int64_t op_stride;
    const float* noise;  // [B, N, T+1, D]
    float* x;            // [B, N, T+1, D]
    int64_t d_l, T, N, p0, count, span;
    int D;
};
template <int DT>
__global__ void __launch_bounds__(HE_THREADS) gaussian_homog_sample(HomogSampleArgs a) {
    __shared__ __align__(16) float sLi[DT * DT];
    __shared__ __align__(16) float sU[DT * 2 * DT];
    const int64_t b = blockIdx.y;
    const int D = a.D, n = 2 * D;
    {
        const float* U = a.opU + b * a.op_stride * DT * 2 * DT;
        const float* Li = a.opLinv + b * a.op_stride * DT * DT;
        for (int i = threadIdx.x; i < DT * DT; i += HE_THREADS) sLi[i] = __ldg(Li + i);
        for (int i = threadIdx.x; i < DT * 2 * DT; i += HE_THREADS) sU[i] = __ldg(U + i);
    }
    __syncthreads();
    const float* ein = a.e_lvl + b * a.d_l * n;
    const int64_t row = (a.T + 1) * (int64_t)D;
    for (int64_t idx = (int64_t)blockIdx.x * HE_THREADS + threadIdx.x; idx < a.count * a.N; idx += (int64_t)gridDim.x * HE_THREADS) {
        const int64_t p = a.p0 + idx / a.N, smp = idx % a.N;
        const int64_t ts = 2 * p * a.span, tm = ts + a.span;
        int64_t te = ts + 2 * a.span;
        if (te > a.T) te = a.T;
        float* xb = a.x + (b * a.N + smp) * row;
        const float* nb = a.noise + (b * a.N + smp) * row;
        const float* ex = ein + 2 * p * n;
        const float* ey = ex + n;
        float z[DT], xac[2 * DT];
#pragma unroll
        for (int k = 0; k < DT; ++k) {
            z[k] = k < D ? ex[D + k] + ey[k] : 0.f;
            xac[k] = k < D ? xb[ts * D + k] : 0.f;
            xac[DT + k] = k < D ? xb[te * D + k] : 0.f;
        }
        float w[DT];
#pragma unroll
        for (int r = 0; r < DT; ++r) {
            float v = r < D ? nb[tm * D + r] : 0.f;
#pragma unroll
            for (int k = 0; k < DT; k += 4) {
                if (k <= r) {
                    const float4 l = *reinterpret_cast<const float4*>(&sLi[r * DT + k]);
                    v = fmaf(l.x, z[k], fmaf(l.y, z[k + 1], fmaf(l.z, z[k + 2], fmaf(l.w, z[k + 3], v))));
                }
            }
#pragma unroll
            for (int c = 0; c < 2 * DT; c += 4) {
                const float4 u = *reinterpret_cast<const float4*>(&sU[r * 2 * DT + c]);
                v = fmaf(-u.x, xac[c], fmaf(-u.y, xac[c + 1], fmaf(-u.z, xac[c + 2], fmaf(-u.w, xac[c + 3], v))));
            }
            w[r] = v;
        }
        // x_m = L^-T w:  (L^-T)[i][r] = Linv[r][i], non-zero for r >= i; padded rows (r >= D) carry the identity and w = 0
        float out[DT];
#pragma unroll
        for (int i = 0; i < DT; ++i) out[i] = 0.f;
#pragma unroll
        for (int r = 0; r < DT; ++r) {
#pragma unroll
            for (int i = 0; i < DT; i += 4) {
                if (i <= r) {
                    const float4 l = *reinterpret_cast<const float4*>(&sLi[r * DT + i]);
                    out[i] = fmaf(l.x, w[r], out[i]);
                    out[i + 1] = fmaf(l.y, w[r], out[i + 1]);
                    out[i + 2] = fmaf(l.z, w[r], out[i + 2]);
                    out[i + 3] = fmaf(l.w, w[r], out[i + 3]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < DT; ++k)
            if (k < D) xb[tm * D + k] = out[k];
    }
}

template <int DT>
static int homog_sample_run(const float* eta, int64_t B, int64_t T, int D, int64_t N, const float* noise, float* x, void* workspace,
                            size_t workspace_bytes, cudaStream_t stream) {
    const int NL = homog_levels_count(T);
    const HomogLayout lay = homog_layout(workspace, workspace_bytes, B, T, D, DT, true);
    if (!lay.ok) {
        set_error("gaussian_chain sample: workspace too small (%zu bytes)", workspace_bytes);
        return FB2_ERR_WORKSPACE;
    }
    // replay the forward bookkeeping: which pairs of each level used the regular / the tail operator
    struct Lvl { int64_t d, reg_pairs; bool tail_pair; };
    Lvl lv[48];
    {
        int64_t d = T;
        bool has_tail = false;
        for (int l = 0; d > 1; ++l) {
            const int64_t npairs = d / 2;
            const bool odd = d & 1;
            const bool tail_pair = has_tail && !odd;
            lv[l] = {d, npairs - (tail_pair ? 1 : 0), tail_pair};
            if (!tail_pair && !has_tail && odd) has_tail = true;
            d = (d + 1) / 2;
        }
    }
    for (int l = NL - 1; l >= 0; --l) {
        HomogSampleArgs sa;
        sa.e_lvl = l == 0 ? eta : lay.lev_e[l];
        sa.op_stride = (int64_t)NL * 2;
        sa.noise = noise; sa.x = x; sa.d_l = lv[l].d; sa.T = T; sa.N = N; sa.span = (int64_t)1 << l; sa.D = D;
        if (lv[l].reg_pairs > 0) {
            sa.opU = lay.opU + (size_t)(l * 2 + 0) * DT * 2 * DT;
            sa.opLinv = lay.opLinv + (size_t)(l * 2 + 0) * DT * DT;
            sa.p0 = 0; sa.count = lv[l].reg_pairs;
            int64_t blocks = ceil_div(sa.count * N, HE_THREADS);
            if (blocks > 65535) blocks = 65535;
            gaussian_homog_sample<DT><<<dim3((unsigned)blocks, (unsigned)B), HE_THREADS, 0, stream>>>(sa);
            FB2_LAUNCH_CHECK();
        }
        if (lv[l].tail_pair) {
            sa.opU = lay.opU + (size_t)(l * 2 + 1) * DT * 2 * DT;
            sa.opLinv = lay.opLinv + (size_t)(l * 2 + 1) * DT * DT;
            sa.p0 = lv[l].d / 2 - 1; sa.count = 1;
            int64_t blocks = ceil_div(N, HE_THREADS);
            gaussian_homog_sample<DT><<<dim3((unsigned)blocks, (unsigned)B), HE_THREADS, 0, stream>>>(sa);
            FB2_LAUNCH_CHECK();
        }
    }
    return FB2_OK;
}


// (Lam, eta, c) -> (w, P, s): a square root of a positive SEMI-definite precision (chain products are rank deficient in general:
// transition factors carry rank D + O < 2D, gaussian.py:402-406) by diagonally pivoted Cholesky, one warp per element.
// P[n][n] = rows of the factor in the ORIGINAL row order, columns in pivot order, zero columns beyond the numerical rank; any P with
// P P' = Lam is a valid prec_sqrt (gaussian.py:419-498; testing.py:149-153 compares Lam and eta only).  w solves P w = eta on the
// range of Lam, s = c + |w|^2 / 2.  Pivoting makes the dropped remainder small ENTRYWISE (every remaining diagonal is below the
// threshold, and |m_ij| <= sqrt(m_ii m_jj) for a PSD remainder); without it one small leading pivot would drop sqrt(tol)-sized entries.
constexpr int I2S_WARPS = 4;
__global__ void __launch_bounds__(I2S_WARPS * 32) gaussian_info_to_sqrt(const float* __restrict__ Lam, const float* __restrict__ eta,
                                                                          const float* __restrict__ c, int64_t N, int n, float rel_tol,
                                                                          float* __restrict__ w, float* __restrict__ P, float* __restrict__ s,
                                                                          int32_t* __restrict__ rank_out) {
    extern __shared__ float sm[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int ld = n + 1;
    float* L = sm + (size_t)wid * (n * ld + n);  // L[i][j] (row i, pivot column j), then wv[n]
    float* wv = L + n * ld;
    const int64_t e = (int64_t)blockIdx.x * I2S_WARPS + wid;
    if (e >= N) return;
    const float* A = Lam + e * (int64_t)n * n;
    const int r0 = lane, r1 = lane + 32;
    float d0 = r0 < n ? A[(int64_t)r0 * n + r0] : -1.f, d1 = r1 < n ? A[(int64_t)r1 * n + r1] : -1.f;
    bool done0 = r0 >= n, done1 = r1 >= n;
    float e0 = r0 < n ? eta[e * n + r0] : 0.f, e1 = r1 < n ? eta[e * n + r1] : 0.f;  // running right-hand side eta - L[:, :j] w[:j]
    const float tol = fmaxf(warp_max(fmaxf(d0, d1)), 0.f) * rel_tol;
    for (int idx = lane; idx < n * ld + n; idx += 32) L[idx] = 0.f;
    __syncwarp();
    int rank = 0;
    float ww = 0.f;
    for (int j = 0; j < n; ++j) {
        // pivot = largest remaining diagonal, smallest index on ties
        float bv = done0 ? -INFINITY : d0;
        int bi = r0;
        if (!done1 && d1 > bv) bv = d1, bi = r1;
        if (done0 && done1) bi = 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) bv = ov, bi = oi;
        }
        if (!(bv > tol)) break;
        const int p = bi;
        // column of the remainder through the pivot: v_i = Lam[p][i] - sum_k L[i][k] L[p][k]   (row p of Lam: a coalesced read)
        float v0 = (r0 < n) ? A[(int64_t)p * n + r0] : 0.f, v1 = (r1 < n) ? A[(int64_t)p * n + r1] : 0.f;
        for (int k = 0; k < j; ++k) {
            const float lp = L[p * ld + k];
            if (r0 < n) v0 = fmaf(-L[r0 * ld + k], lp, v0);
            if (r1 < n) v1 = fmaf(-L[r1 * ld + k], lp, v1);
        }
        const float vp = __shfl_sync(0xffffffffu, (p & 32) ? v1 : v0, p & 31);
        const float piv = sqrtf(fmaxf(vp, tol * 0.5f));
        const float inv = 1.f / piv;
        const float ep = __shfl_sync(0xffffffffu, (p & 32) ? e1 : e0, p & 31);
        const float wj = ep * inv;
        ww = fmaf(wj, wj, ww);
        if (lane == 0) wv[j] = wj;
        if (r0 < n) {
            float l = (r0 == p) ? piv : (done0 ? 0.f : v0 * inv);
            L[r0 * ld + j] = l;
            d0 -= l * l;
            e0 -= l * wj;
            if (r0 == p) done0 = true;
        }
        if (r1 < n) {
            float l = (r1 == p) ? piv : (done1 ? 0.f : v1 * inv);
            L[r1 * ld + j] = l;
            d1 -= l * l;
            e1 -= l * wj;
            if (r1 == p) done1 = true;
        }
        rank = j + 1;
        __syncwarp();
    }
    float* Po = P + e * (int64_t)n * n;
    for (int idx = lane; idx < n * n; idx += 32) Po[idx] = L[(idx / n) * ld + (idx % n)];
    for (int k = lane; k < n; k += 32) w[e * n + k] = wv[k];
    if (lane == 0) {
        s[e] = c[e] + 0.5f * ww;
        if (rank_out) rank_out[e] = rank;
    }
}


// ---------------------------------------------------------------------------------------------
// Moment matching of a Gaussian mixture (joint.py:102-150: the Contraction[Logaddexp, Add, {discrete}, Tensor, Gaussian] rule of the
// `moment_matching` interpretation, which is what makes the SwitchingLinearHMM scan tractable, pyro/hmm.py:527-551).
// For every output element e: K components (logit d_k, Lam_k, eta_k, c_k) over n reals ->
//     logn_k = c_k + n/2 log 2pi - sum log diag L_k + |L_k^-1 eta_k|^2 / 2          (L_k L_k' = Lam_k: the component's mass)
//     Z = logsumexp_k (d_k + logn_k),  p_k = exp(d_k + logn_k - Z),  mu_k = Lam_k^-1 eta_k,  Sig_k = Lam_k^-1
//     mu = sum p_k mu_k,   Sig = sum p_k (Sig_k + (mu_k - mu)(mu_k - mu)')              (an empty mixture gets Sig = I, joint.py:141-143)
//     Lam' = Sig^-1,  eta' = Lam' mu,  c' = Z - n/2 log 2pi - 1/2 log det Sig - mu' eta' / 2   (a normalised density of mass exp Z)
// One warp per element; every matrix lives in the warp's shared-memory slot (Cholesky in place, inverse through two triangular
// solves with one right-hand-side column per lane).  Two passes over the components: the weights need Z first, and the centred form
// (mu_k - mu) avoids the cancellation of  sum p_k mu_k mu_k' - mu mu'.
constexpr int MM_WARPS = 2;
struct MMSlot {  // layout inside the dynamic shared memory of one warp, n <= 64, ld = n + 1
    float* A;    // [n][ld]  working matrix / Cholesky factor
    float* S;    // [n][ld]  accumulated covariance
    float* Y;    // [n][64]  right-hand sides of the inverse, one column per (lane, half)
    float* vec;  // [4][n]   eta / v / mu_k / mu
};
__device__ __forceinline__ bool mm_cholesky(float* A, int n, int ld, int lane, float& half_logdet) {
    // right-looking Cholesky, lane owns rows lane and lane + 32; returns false on a non-positive pivot
    half_logdet = 0.f;
    bool ok = true;
    for (int j = 0; j < n; ++j) {
        const float d = A[j * ld + j];
        if (!(d > 0.f)) ok = false;
        const float piv = sqrtf(fmaxf(d, 1e-30f));
        half_logdet += logf(piv);
        const float inv = 1.f / piv;
        __syncwarp();
        for (int i = lane; i < n; i += 32) {
            if (i == j) A[j * ld + j] = piv;
            else if (i > j) A[i * ld + j] *= inv;
        }
        __syncwarp();
        for (int i = j + 1 + lane; i < n; i += 32) {
            const float lij = A[i * ld + j];
            for (int c = j + 1; c <= i; ++c) A[i * ld + c] = fmaf(-lij, A[c * ld + j], A[i * ld + c]);
        }
    }
    __syncwarp();
    return ok;
}
// x = L^-1 b (in place in `b`, shared memory), then optionally x = L^-T x
__device__ __forceinline__ void mm_solve_vec(const float* L, int n, int ld, float* b, int lane, bool back) {
    for (int j = 0; j < n; ++j) {
        float part = 0.f;
        for (int m = lane; m < j; m += 32) part = fmaf(L[j * ld + m], b[m], part);
        part = warp_sum(part);
        __syncwarp();
        if (lane == 0) b[j] = (b[j] - part) / L[j * ld + j];
        __syncwarp();
    }
    if (!back) return;
    for (int j = n - 1; j >= 0; --j) {
        float part = 0.f;
        for (int m = j + 1 + lane; m < n; m += 32) part = fmaf(L[m * ld + j], b[m], part);
        part = warp_sum(part);
        __syncwarp();
        if (lane == 0) b[j] = (b[j] - part) / L[j * ld + j];
        __syncwarp();
    }
}
// Y[:, col] = (L L')^-1 e_col for the columns this lane owns (col = lane, lane + 32); all lanes walk the same (j, m) so L is a broadcast read
__device__ __forceinline__ void mm_inverse_columns(const float* L, int n, int ld, float* Y, int lane) {
    for (int h = 0; h < 2; ++h) {
        const int col = lane + 32 * h;
        float* y = Y + col;  // stride 64
        if (col < n) {
            for (int j = 0; j < n; ++j) {
                float acc = j == col ? 1.f : 0.f;
                for (int m = col; m < j; ++m) acc = fmaf(-L[j * ld + m], y[m * 64], acc);  // y[m] = 0 for m < col
                y[j * 64] = j < col ? 0.f : acc / L[j * ld + j];
            }
            for (int j = n - 1; j >= 0; --j) {
                float acc = y[j * 64];
                for (int m = j + 1; m < n; ++m) acc = fmaf(-L[m * ld + j], y[m * 64], acc);
                y[j * 64] = acc / L[j * ld + j];
            }
        }
    }
    __syncwarp();
}
__global__ void __launch_bounds__(MM_WARPS * 32) gaussian_moment_match(const float* __restrict__ logits, const float* __restrict__ Lam,
                                                                        const float* __restrict__ eta, const float* __restrict__ c, int64_t N, int K,
                                                                        int n, float* __restrict__ scratch /*[N,K,n+1]: mu_k, logit_k + logn_k*/,
                                                                        float* __restrict__ Lo, float* __restrict__ eo, float* __restrict__ co,
                                                                        int32_t* __restrict__ flag) {
    extern __shared__ float sm[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int ld = n + 1;
    const size_t per_warp = (size_t)2 * n * ld + (size_t)n * 64 + 4 * n;
    float* base = sm + wid * per_warp;
    float *A = base, *S = A + n * ld, *Y = S + n * ld, *vec = Y + n * 64;
    float *v0 = vec, *mu = vec + n;
    const int64_t e = (int64_t)blockIdx.x * MM_WARPS + wid;
    if (e >= N) return;
    float* sc = scratch + e * (int64_t)K * (n + 1);
    // ---- pass 1: mass and mean of every component
    float zmax = -INFINITY;
    for (int k = 0; k < K; ++k) {
        const float* Lk = Lam + (e * K + k) * (int64_t)n * n;
        for (int idx = lane; idx < n * n; idx += 32) A[(idx / n) * ld + idx % n] = Lk[idx];
        for (int i = lane; i < n; i += 32) v0[i] = eta[(e * K + k) * (int64_t)n + i];
        __syncwarp();
        float hl;
        if (!mm_cholesky(A, n, ld, lane, hl) && lane == 0 && flag) atomicExch(flag, 1);
        mm_solve_vec(A, n, ld, v0, lane, false);  // v = L^-1 eta
        float q = 0.f;
        for (int i = lane; i < n; i += 32) q = fmaf(v0[i], v0[i], q);
        q = warp_sum(q);
        const float lg = logits[e * K + k] + c[e * K + k] + n * kHalfLog2Pi - hl + 0.5f * q;
        // mu_k = L^-T v
        for (int j = n - 1; j >= 0; --j) {
            float part = 0.f;
            for (int m = j + 1 + lane; m < n; m += 32) part = fmaf(A[m * ld + j], v0[m], part);
            part = warp_sum(part);
            __syncwarp();
            if (lane == 0) v0[j] = (v0[j] - part) / A[j * ld + j];
            __syncwarp();
        }
        for (int i = lane; i < n; i += 32) sc[k * (n + 1) + i] = v0[i];
        if (lane == 0) sc[k * (n + 1) + n] = lg;
        zmax = fmaxf(zmax, lg);
        __syncwarp();
    }
    float wsum = 0.f;
    for (int k = 0; k < K; ++k) wsum += zmax == -INFINITY ? 0.f : __expf(sc[k * (n + 1) + n] - zmax);
    const float Z = zmax == -INFINITY ? -INFINITY : zmax + logf(wsum);
    for (int i = lane; i < n; i += 32) {
        float m = 0.f;
        for (int k = 0; k < K; ++k) {
            const float pk = Z == -INFINITY ? 0.f : __expf(sc[k * (n + 1) + n] - Z);
            m = fmaf(pk, sc[k * (n + 1) + i], m);
        }
        mu[i] = m;
    }
    for (int idx = lane; idx < n * ld; idx += 32) S[idx] = 0.f;
    __syncwarp();
    // ---- pass 2: Sig = sum p_k (Lam_k^-1 + (mu_k - mu)(mu_k - mu)')
    for (int k = 0; k < K; ++k) {
        const float pk = Z == -INFINITY ? 0.f : __expf(sc[k * (n + 1) + n] - Z);
        if (pk == 0.f) continue;  // warp-uniform
        const float* Lk = Lam + (e * K + k) * (int64_t)n * n;
        for (int idx = lane; idx < n * n; idx += 32) A[(idx / n) * ld + idx % n] = Lk[idx];
        __syncwarp();
        float hl;
        mm_cholesky(A, n, ld, lane, hl);
        mm_inverse_columns(A, n, ld, Y, lane);
        for (int i = lane; i < n; i += 32) v0[i] = sc[k * (n + 1) + i] - mu[i];
        __syncwarp();
        for (int idx = lane; idx < n * n; idx += 32) {
            const int i = idx / n, j = idx % n;
            S[i * ld + j] += pk * (Y[i * 64 + j] + v0[i] * v0[j]);
        }
        __syncwarp();
    }
    if (Z == -INFINITY) {  // empty mixture: bogus unit covariance (joint.py:141-143)
        for (int i = lane; i < n; i += 32) S[i * ld + i] = 1.f;
        __syncwarp();
    }
    // ---- Lam' = Sig^-1, eta' = Lam' mu, c'
    for (int idx = lane; idx < n * n; idx += 32) {  // symmetrise against rounding before factorising
        const int i = idx / n, j = idx % n;
        A[i * ld + j] = 0.5f * (S[i * ld + j] + S[j * ld + i]);
    }
    __syncwarp();
    float hl;
    if (!mm_cholesky(A, n, ld, lane, hl) && lane == 0 && flag) atomicExch(flag, 1);
    mm_inverse_columns(A, n, ld, Y, lane);
    float quad = 0.f;
    for (int i = lane; i < n; i += 32) {
        float acc = 0.f;
        for (int j = 0; j < n; ++j) acc = fmaf(0.5f * (Y[i * 64 + j] + Y[j * 64 + i]), mu[j], acc);
        eo[e * n + i] = acc;
        quad = fmaf(acc, mu[i], quad);
    }
    quad = warp_sum(quad);
    for (int idx = lane; idx < n * n; idx += 32) {
        const int i = idx / n, j = idx % n;
        Lo[e * (int64_t)n * n + idx] = 0.5f * (Y[i * 64 + j] + Y[j * 64 + i]);
    }
    if (lane == 0) co[e] = Z - n * kHalfLog2Pi - hl - 0.5f * quad;
}


// Chain elements of a time-homogeneous linear-Gaussian state-space model straight from the observations (GaussianHMM, pyro/hmm.py:258-270:
// element_t = trans + obs(value = y_t); matrix_and_mvn_to_funsor pyro/convert.py:278-298; gaussian.py:723-744 for the substitution).
// Only white_vec depends on t: w_t = [w_trans ; w_o + y_t Pr], so with u_t = w_o + y_t Pr (O values)
//     eta_t = P w_t = a0 + Po u_t        (a0 = P[:, :D] w_trans, Po = P[:, D:], both per chain)
//     c_t   = s - |w_t|^2 / 2 = c0 - |u_t|^2 / 2
// One warp per time step (strided over the chain): lane o < O owns u[o], lane i owns eta[i] and eta[i + 32]; y is read once (64 B per
// step at O = 16) and eta / c are written once, coalesced -- the [B, T, D + O] white_vec array of the reference's data flow (and the
// torch cat / matmul / expand that built it: 6 of the 10.3 ms of BASELINE config 4) never exists.
// One THREAD per time step for the arithmetic (Pr / Po sit in shared memory and are read as broadcast LDS.128: every lane of a warp
// multiplies by the same matrix entry), one WARP per tile of 32 consecutive steps for the memory traffic: the tile's observations are
// loaded coalesced into shared memory, the 32 x 2D block of results goes back through shared memory so that every global store is a
// full 128-byte line.  (A warp-per-step version with shuffles was measured first: 6.7 ms at C4 -- a chain of ~50 dependent
// shuffle / FMA pairs per step; this one is bound by the 1280 FMAs per step and the 330 bytes per step of traffic.)
template <int OT, int KE_WARPS>
__global__ void __launch_bounds__(KE_WARPS * 32) gaussian_kalman_eta(const float* __restrict__ y, const float* __restrict__ Pr, const float* __restrict__ wo,
                                                                      const float* __restrict__ a0, const float* __restrict__ Po, const float* __restrict__ c0,
                                                                      int64_t T, int D, int O, float* __restrict__ eta, float* __restrict__ c,
                                                                      int64_t y_T, int64_t t_begin) {
    // y_T = time extent of y per chain, t_begin = first step of the range; the T steps [t_begin, t_begin + T) are written to eta[B,T,2D], c[B,T]
    __shared__ __align__(16) float sPr[OT * OT];     // [o_in][o_out]
    __shared__ __align__(16) float sPo[64 * OT];     // [i][o], rows beyond 2D are zero
    __shared__ float sa0[64], swo[OT];
    __shared__ __align__(16) float sy[KE_WARPS][32 * OT];
    __shared__ float se[KE_WARPS][32 * 65];
    const int64_t b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = 2 * D;
    for (int i = threadIdx.x; i < OT * OT; i += blockDim.x) {
        const int oi = i / OT, oo = i % OT;
        sPr[i] = (oi < O && oo < O) ? Pr[(b * O + oi) * O + oo] : 0.f;
    }
    for (int i = threadIdx.x; i < 64 * OT; i += blockDim.x) {
        const int r = i / OT, o = i % OT;
        sPo[i] = (r < n && o < O) ? Po[(b * n + r) * O + o] : 0.f;
    }
    for (int i = threadIdx.x; i < 64; i += blockDim.x) sa0[i] = i < n ? a0[b * n + i] : 0.f;
    for (int i = threadIdx.x; i < OT; i += blockDim.x) swo[i] = i < O ? wo[b * O + i] : 0.f;
    __syncthreads();
    const float cc = c0[b];
    const float* yb = y + (b * y_T + t_begin) * O;
    float* eb = eta + b * T * n;
    float* cb = c + b * T;
    float* myy = sy[warp];
    float* mye = se[warp];
    const int64_t ntiles = (T + 31) / 32;
    for (int64_t tile = (int64_t)blockIdx.x * KE_WARPS + warp; tile < ntiles; tile += (int64_t)gridDim.x * KE_WARPS) {
        const int64_t t0 = tile * 32;
        const int nt = (int)(T - t0 < 32 ? T - t0 : 32);
        // observations of the tile: nt * O contiguous floats
        for (int i = lane; i < 32 * OT; i += 32) {
            const int st = i / OT, o = i % OT;
            myy[i] = (st < nt && o < O) ? __ldg(yb + (t0 + st) * O + o) : 0.f;
        }
        __syncwarp();
        float u[OT];
#pragma unroll
        for (int o = 0; o < OT; ++o) u[o] = swo[o];
#pragma unroll
        for (int oi = 0; oi < OT; ++oi) {
            const float yv = myy[lane * OT + oi];
#pragma unroll
            for (int o4 = 0; o4 < OT; o4 += 4) {
                const float4 p = *reinterpret_cast<const float4*>(&sPr[oi * OT + o4]);
                u[o4] = fmaf(yv, p.x, u[o4]);
                u[o4 + 1] = fmaf(yv, p.y, u[o4 + 1]);
                u[o4 + 2] = fmaf(yv, p.z, u[o4 + 2]);
                u[o4 + 3] = fmaf(yv, p.w, u[o4 + 3]);
            }
        }
        float uu = 0.f;
#pragma unroll
        for (int o = 0; o < OT; ++o) uu = fmaf(u[o], u[o], uu);
        if (lane < nt) cb[t0 + lane] = cc - 0.5f * uu;
#pragma unroll 1
        for (int i0 = 0; i0 < n; i0 += 8) {
            float e[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float acc = sa0[i0 + q];
#pragma unroll
                for (int o4 = 0; o4 < OT; o4 += 4) {
                    const float4 p = *reinterpret_cast<const float4*>(&sPo[(i0 + q) * OT + o4]);
                    acc = fmaf(u[o4], p.x, fmaf(u[o4 + 1], p.y, fmaf(u[o4 + 2], p.z, fmaf(u[o4 + 3], p.w, acc))));
                }
                e[q] = acc;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) mye[lane * 65 + i0 + q] = e[q];
        }
        __syncwarp();
        // coalesced stores: the tile's nt * n results are contiguous in global memory
        for (int i = lane; i < nt * n; i += 32) eb[t0 * n + i] = mye[(i / n) * 65 + i % n];
        __syncwarp();
    }
}


// ---------------------------------------------------------------------------------------------
// BASELINE config 4 in one library call: GaussianHMM with fixed dynamics, from the model parameters and the observations to the chain
// result.  Three pieces replace the [B]-sized torch linear algebra, the level-0 eta array and the first K tree levels of the eta sweep:
//
// gaussian_kalman_constants: one CTA per chain.  Everything of the chain element that does not depend on t (pyro/hmm.py:258-270,
//   matrix_and_mvn_to_funsor pyro/convert.py:278-298): Pq = (q_tril^-1)', Pr = (r_tril^-1)', the prec_sqrt P = [[F Pq, 0], [-Pq, H Pr]],
//   Lam = P P', w_trans = -q_loc Pq, wo = -r_loc Pr, a0 = P[:, :D] w_trans, Po = P[:, D:], c0 = s - |w_trans|^2 / 2.
struct KalmanConstArgs {
    const float *F, *q_loc, *q_tril, *H, *r_loc, *r_tril;  // [B,D,D] [B,D] [B,D,D] [B,D,O] [B,O] [B,O,O]
    float *Lam, *Pr, *wo, *a0, *Po, *c0;                   // [B,2D,2D] [B,O,O] [B,O] [B,2D] [B,2D,O] [B]
    int D, O;
};
__global__ void __launch_bounds__(256) gaussian_kalman_constants(KalmanConstArgs a) {
    __shared__ float Xq[32][33], Xr[32][33];  // inverses of q_tril / r_tril (lower triangular)
    __shared__ float sP[64][65];              // prec_sqrt, rows (state; state'), columns (transition noise; observation noise)
    __shared__ float swt[32], slog[64];
    const int64_t b = blockIdx.x;
    const int D = a.D, O = a.O, n = 2 * D, r = D + O, tid = threadIdx.x;
    const float* qt = a.q_tril + b * D * D;
    const float* rt = a.r_tril + b * O * O;
    // column j of the inverse of a lower-triangular matrix by forward substitution; thread j < D: q_tril, thread 32 + j, j < O: r_tril
    if (tid < 64) {
        const bool isq = tid < 32;
        const int j = tid & 31, m = isq ? D : O;
        const float* L = isq ? qt : rt;
        float(*X)[33] = isq ? Xq : Xr;
        if (j < m) {
            for (int i = 0; i < j; ++i) X[i][j] = 0.f;
            const float dj = L[j * m + j];
            X[j][j] = 1.f / dj;
            slog[tid] = logf(dj);
            for (int i = j + 1; i < m; ++i) {
                float acc = 0.f;
                for (int k = j; k < i; ++k) acc = fmaf(L[i * m + k], X[k][j], acc);
                X[i][j] = -acc / L[i * m + i];
            }
        } else {
            slog[tid] = 0.f;
        }
    }
    __syncthreads();
    // P[i][j];  Pq[k][j] = Xq[j][k],  Pr[o][j] = Xr[j][o]
    const float* F = a.F + b * D * D;
    const float* H = a.H + b * D * O;
    for (int idx = tid; idx < n * r; idx += blockDim.x) {
        const int i = idx / r, j = idx % r;
        float v = 0.f;
        if (j < D) {
            if (i < D) {
                for (int k = 0; k < D; ++k) v = fmaf(F[i * D + k], Xq[j][k], v);
            } else {
                v = -Xq[j][i - D];
            }
        } else if (i >= D) {
            const int jo = j - D;
            for (int o = 0; o < O; ++o) v = fmaf(H[(i - D) * O + o], Xr[jo][o], v);
        }
        sP[i][j] = v;
    }
    if (tid < D) {  // w_trans[j] = -sum_k q_loc[k] Pq[k][j]
        float v = 0.f;
        for (int k = 0; k < D; ++k) v = fmaf(a.q_loc[b * D + k], Xq[tid][k], v);
        swt[tid] = -v;
    }
    if (tid >= 32 && tid < 32 + O) {  // wo[j] = -sum_k r_loc[k] Pr[k][j]
        const int j = tid - 32;
        float v = 0.f;
        for (int k = 0; k < O; ++k) v = fmaf(a.r_loc[b * O + k], Xr[j][k], v);
        a.wo[b * O + j] = -v;
    }
    __syncthreads();
    for (int idx = tid; idx < n * n; idx += blockDim.x) {
        const int i = idx / n, j = idx % n;
        float v = 0.f;
        for (int k = 0; k < r; ++k) v = fmaf(sP[i][k], sP[j][k], v);
        a.Lam[b * n * n + idx] = v;
    }
    for (int idx = tid; idx < n * O; idx += blockDim.x) a.Po[b * n * O + idx] = sP[idx / O][D + idx % O];
    for (int idx = tid; idx < O * O; idx += blockDim.x) a.Pr[b * O * O + idx] = Xr[idx % O][idx / O];
    if (tid < n) {
        float v = 0.f;
        for (int k = 0; k < D; ++k) v = fmaf(sP[tid][k], swt[k], v);
        a.a0[b * n + tid] = v;
    }
    if (tid == 0) {
        float sl = 0.f, ww = 0.f;
        for (int k = 0; k < 64; ++k) sl += slog[k];
        for (int k = 0; k < D; ++k) ww = fmaf(swt[k], swt[k], ww);
        a.c0[b] = -(float)(D + O) * kHalfLog2Pi - sl - 0.5f * ww;
    }
}

// gaussian_kalman_tile: the per-step elements AND the first K levels of the reference's tree (sum_product.py:690-701) for the full
//   tiles of 2^K consecutive steps, one THREAD per tile.  Level-l element j of the tree covers the steps [j 2^l, (j+1) 2^l) (clipped
//   at T), so an aligned tile is a complete subtree whose pairs all use the level's regular operator: the thread walks its 2^K steps
//   in order (eta_t, c_t straight from y_t as in gaussian_kalman_eta), and folds finished subtrees like a binary counter -- after step
//   i it combines once per trailing one-bit of i; the pending left operands (at most K) sit in a per-thread stack.  Same arithmetic,
//   same order as gaussian_kalman_eta + gaussian_homog_eta, so the results are bit-identical to the level-by-level path, but the
//   level-0 .. level-(K-1) eta arrays (1.6 GB + 0.8 GB + ... at BASELINE config 4) never exist: per tile y is read once (2^K O floats)
//   and one element (2D + 1 floats) is written.  The operators of the K levels sit in shared memory as broadcast LDS.128 operands.
struct KalmanTileArgs {
    const float *eta_in, *c_in;  // non-null: the per-step elements are given (eta[B,T,2D], c[B,T]) instead of being formed from y
    const float *y, *Pr, *wo, *a0, *Po, *c0;
    const float *opU, *opLinv, *konst;  // per-level operators of gaussian_homog_levels: regular operator of level l at ((b NL + l) 2)
    float *eout, *cout;                 // level-K arrays [B, out_T, 2D], [B, out_T]
    int64_t T, ntiles, out_T;
    int D, O, K, NL;
};
constexpr int KT_THREADS = 128;
constexpr int KT_KMAX = 6;
template <int DT, int OT>
__host__ __device__ constexpr int kt_const_floats() { return OT * OT + 2 * DT * OT + 2 * DT + OT + 8; }
template <int DT>
__host__ __device__ constexpr int kt_level_floats() { return DT * DT + DT * 2 * DT; }

// NTL = tiles per thread (consecutive tiles, processed in lock step): every broadcast LDS.128 of operator entries then feeds 4 NTL FMAs.
// With one tile per thread the kernel is bound by shared-memory operand bandwidth (one LDS.128 per 4 FMAs: ~35 % of the FMA rate measured).
template <int DT, int OT, int NTL>
__global__ void __launch_bounds__(KT_THREADS, 2) gaussian_kalman_tile(KalmanTileArgs a) {
    extern __shared__ __align__(16) float kt_smem[];
    float* sPr = kt_smem;                 // [o_in][o_out]
    float* sPo = sPr + OT * OT;           // [blk * DT + k][o]: padded rows are zero
    float* sa0 = sPo + 2 * DT * OT;
    float* swo = sa0 + 2 * DT;
    float* sk = swo + OT;                 // konst per level
    float* sops = kt_smem + kt_const_floats<DT, OT>();  // per level: L^-1 [DT][DT], U [DT][2 DT]
    const int64_t b = blockIdx.y;
    const int D = a.D, O = a.O, n = 2 * D, K = a.K;
    const bool given = a.eta_in != nullptr;
    if (!given) {
        for (int i = threadIdx.x; i < OT * OT; i += KT_THREADS) {
            const int oi = i / OT, oo = i % OT;
            sPr[i] = (oi < O && oo < O) ? a.Pr[(b * O + oi) * O + oo] : 0.f;
        }
        for (int i = threadIdx.x; i < 2 * DT * OT; i += KT_THREADS) {
            const int row = i / OT, o = i % OT, blk = row / DT, k = row % DT;
            sPo[i] = (k < D && o < O) ? a.Po[(b * n + blk * D + k) * O + o] : 0.f;
        }
        for (int i = threadIdx.x; i < 2 * DT; i += KT_THREADS) sa0[i] = (i % DT) < D ? a.a0[b * n + (i / DT) * D + i % DT] : 0.f;
        for (int i = threadIdx.x; i < OT; i += KT_THREADS) swo[i] = i < O ? a.wo[b * O + i] : 0.f;
    }
    for (int l = 0; l < K; ++l) {
        const size_t op = ((size_t)b * a.NL + l) * 2;
        const float4* Li = reinterpret_cast<const float4*>(a.opLinv + op * DT * DT);
        const float4* U = reinterpret_cast<const float4*>(a.opU + op * DT * 2 * DT);
        float4* dLi = reinterpret_cast<float4*>(sops + (size_t)l * kt_level_floats<DT>());
        float4* dU = dLi + DT * DT / 4;
        for (int i = threadIdx.x; i < DT * DT / 4; i += KT_THREADS) dLi[i] = __ldg(Li + i);
        for (int i = threadIdx.x; i < DT * 2 * DT / 4; i += KT_THREADS) dU[i] = __ldg(U + i);
        if (threadIdx.x == 0) sk[l] = a.konst[op];
    }
    __syncthreads();
    const int64_t tile0 = ((int64_t)blockIdx.x * KT_THREADS + threadIdx.x) * NTL;
    if (tile0 >= a.ntiles) return;
    const float c0 = given ? 0.f : a.c0[b];
    const float* yp[NTL];
    const float* ep[NTL];  // given elements: eta / c of the tile's first step
    const float* cp[NTL];
    bool live[NTL];
#pragma unroll
    for (int q = 0; q < NTL; ++q) {
        live[q] = tile0 + q < a.ntiles;
        const int64_t t_first = b * a.T + ((live[q] ? tile0 + q : tile0) << K);  // a missing last tile repeats its neighbour (not stored)
        yp[q] = given ? nullptr : a.y + t_first * O;
        ep[q] = given ? a.eta_in + t_first * n : nullptr;
        cp[q] = given ? a.c_in + t_first : nullptr;
    }
    const bool yvec = !given && (O == OT) && (reinterpret_cast<uintptr_t>(a.y) % 16 == 0);
    const bool evec = given && (D % 4 == 0) && (reinterpret_cast<uintptr_t>(a.eta_in) % 16 == 0);
    float4 stk[NTL][KT_KMAX][2 * DT / 4];  // pending left operands, one per level (local memory: indexed by the run-time level)
    float stc[NTL][KT_KMAX];
    float cur[NTL][2 * DT], cc[NTL];
    float yv[NTL][OT];
    auto load_y = [&](int step) {
#pragma unroll
        for (int q = 0; q < NTL; ++q) {
            const float* src = yp[q] + (size_t)step * O;
            if (yvec) {
#pragma unroll
                for (int o = 0; o < OT; o += 4) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(src + o));
                    yv[q][o] = v.x; yv[q][o + 1] = v.y; yv[q][o + 2] = v.z; yv[q][o + 3] = v.w;
                }
            } else {
#pragma unroll
                for (int o = 0; o < OT; ++o) yv[q][o] = o < O ? __ldg(src + o) : 0.f;
            }
        }
    };
    if (!given) load_y(0);
    const int nsteps = 1 << K;
#pragma unroll 1
    for (int i = 0; i < nsteps; ++i) {
      if (given) {
        // ---- element of step i as given: eta in the padded register layout, c
#pragma unroll
        for (int q = 0; q < NTL; ++q) {
            const float* e = ep[q] + (size_t)i * n;
            if (evec) {
#pragma unroll
                for (int k = 0; k < DT; k += 4) {
                    if (k < D) {
                        const float4 ea = __ldg(reinterpret_cast<const float4*>(e + k)), ec = __ldg(reinterpret_cast<const float4*>(e + D + k));
                        cur[q][k] = ea.x; cur[q][k + 1] = ea.y; cur[q][k + 2] = ea.z; cur[q][k + 3] = ea.w;
                        cur[q][DT + k] = ec.x; cur[q][DT + k + 1] = ec.y; cur[q][DT + k + 2] = ec.z; cur[q][DT + k + 3] = ec.w;
                    } else {
                        cur[q][k] = cur[q][k + 1] = cur[q][k + 2] = cur[q][k + 3] = 0.f;
                        cur[q][DT + k] = cur[q][DT + k + 1] = cur[q][DT + k + 2] = cur[q][DT + k + 3] = 0.f;
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < DT; ++k) {
                    cur[q][k] = k < D ? __ldg(e + k) : 0.f;
                    cur[q][DT + k] = k < D ? __ldg(e + D + k) : 0.f;
                }
            }
            cc[q] = __ldg(cp[q] + i);
        }
      } else {
        // ---- element of step i: u = wo + y Pr, eta = a0 + Po u, c = c0 - |u|^2 / 2
        float u[NTL][OT];
#pragma unroll
        for (int q = 0; q < NTL; ++q)
#pragma unroll
            for (int o = 0; o < OT; ++o) u[q][o] = swo[o];
#pragma unroll
        for (int oi = 0; oi < OT; ++oi) {
#pragma unroll
            for (int o4 = 0; o4 < OT; o4 += 4) {
                const float4 p = *reinterpret_cast<const float4*>(&sPr[oi * OT + o4]);
#pragma unroll
                for (int q = 0; q < NTL; ++q) {
                    u[q][o4] = fmaf(yv[q][oi], p.x, u[q][o4]);
                    u[q][o4 + 1] = fmaf(yv[q][oi], p.y, u[q][o4 + 1]);
                    u[q][o4 + 2] = fmaf(yv[q][oi], p.z, u[q][o4 + 2]);
                    u[q][o4 + 3] = fmaf(yv[q][oi], p.w, u[q][o4 + 3]);
                }
            }
        }
        if (i + 1 < nsteps) load_y(i + 1);  // next step's observations travel while this step computes
#pragma unroll
        for (int q = 0; q < NTL; ++q) {
            float uu = 0.f;
#pragma unroll
            for (int o = 0; o < OT; ++o) uu = fmaf(u[q][o], u[q][o], uu);
            cc[q] = c0 - 0.5f * uu;
        }
#pragma unroll
        for (int r = 0; r < 2 * DT; ++r) {
            float acc[NTL];
#pragma unroll
            for (int q = 0; q < NTL; ++q) acc[q] = sa0[r];
#pragma unroll
            for (int o4 = 0; o4 < OT; o4 += 4) {
                const float4 p = *reinterpret_cast<const float4*>(&sPo[r * OT + o4]);
#pragma unroll
                for (int q = 0; q < NTL; ++q)
                    acc[q] = fmaf(u[q][o4], p.x, fmaf(u[q][o4 + 1], p.y, fmaf(u[q][o4 + 2], p.z, fmaf(u[q][o4 + 3], p.w, acc[q]))));
            }
#pragma unroll
            for (int q = 0; q < NTL; ++q) cur[q][r] = acc[q];
        }
      }
        // ---- fold the finished subtrees: one combine per trailing one-bit of i
        int lvl = 0;
#pragma unroll 1
        for (int m = i; m & 1; m >>= 1, ++lvl) {
            const float* sLi = sops + (size_t)lvl * kt_level_floats<DT>();
            const float* sU = sLi + DT * DT;
            float z[NTL][DT];
            float cxy[NTL], vv[NTL];
#pragma unroll
            for (int q = 0; q < NTL; ++q) {
#pragma unroll
                for (int k4 = 0; k4 < DT / 4; ++k4) {
                    const float4 xa = stk[q][lvl][k4], xb = stk[q][lvl][DT / 4 + k4];
                    z[q][4 * k4] = xb.x + cur[q][4 * k4];
                    z[q][4 * k4 + 1] = xb.y + cur[q][4 * k4 + 1];
                    z[q][4 * k4 + 2] = xb.z + cur[q][4 * k4 + 2];
                    z[q][4 * k4 + 3] = xb.w + cur[q][4 * k4 + 3];
                    cur[q][4 * k4] = xa.x; cur[q][4 * k4 + 1] = xa.y; cur[q][4 * k4 + 2] = xa.z; cur[q][4 * k4 + 3] = xa.w;
                }
                cxy[q] = stc[q][lvl] + cc[q];
                vv[q] = 0.f;
            }
#pragma unroll
            for (int r = 0; r < DT; ++r) {
                float v[NTL];
#pragma unroll
                for (int q = 0; q < NTL; ++q) v[q] = 0.f;
#pragma unroll
                for (int k = 0; k < DT; k += 4) {
                    if (k <= r) {
                        const float4 l = *reinterpret_cast<const float4*>(&sLi[r * DT + k]);
#pragma unroll
                        for (int q = 0; q < NTL; ++q)
                            v[q] = fmaf(l.x, z[q][k], fmaf(l.y, z[q][k + 1], fmaf(l.z, z[q][k + 2], fmaf(l.w, z[q][k + 3], v[q]))));
                    }
                }
#pragma unroll
                for (int q = 0; q < NTL; ++q) vv[q] = fmaf(v[q], v[q], vv[q]);
#pragma unroll
                for (int c = 0; c < 2 * DT; c += 4) {
                    const float4 uo = *reinterpret_cast<const float4*>(&sU[r * 2 * DT + c]);
#pragma unroll
                    for (int q = 0; q < NTL; ++q) {
                        cur[q][c] = fmaf(-uo.x, v[q], cur[q][c]);
                        cur[q][c + 1] = fmaf(-uo.y, v[q], cur[q][c + 1]);
                        cur[q][c + 2] = fmaf(-uo.z, v[q], cur[q][c + 2]);
                        cur[q][c + 3] = fmaf(-uo.w, v[q], cur[q][c + 3]);
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < NTL; ++q) cc[q] = cxy[q] + sk[lvl] + 0.5f * vv[q];
        }
        if (i + 1 < nsteps) {
#pragma unroll
            for (int q = 0; q < NTL; ++q) {
#pragma unroll
                for (int k4 = 0; k4 < 2 * DT / 4; ++k4)
                    stk[q][lvl][k4] = make_float4(cur[q][4 * k4], cur[q][4 * k4 + 1], cur[q][4 * k4 + 2], cur[q][4 * k4 + 3]);
                stc[q][lvl] = cc[q];
            }
        }
    }
#pragma unroll
    for (int q = 0; q < NTL; ++q) {
        if (!live[q]) continue;
        float* eo = a.eout + (b * a.out_T + tile0 + q) * n;
        if (D == DT && reinterpret_cast<uintptr_t>(a.eout) % 16 == 0) {
#pragma unroll
            for (int k4 = 0; k4 < 2 * DT / 4; ++k4)
                reinterpret_cast<float4*>(eo)[k4] = make_float4(cur[q][4 * k4], cur[q][4 * k4 + 1], cur[q][4 * k4 + 2], cur[q][4 * k4 + 3]);
        } else {
#pragma unroll
            for (int k = 0; k < DT; ++k) {
                if (k < D) {
                    eo[k] = cur[q][k];
                    eo[D + k] = cur[q][DT + k];
                }
            }
        }
        a.cout[b * a.out_T + tile0 + q] = cc[q];
    }
}

// gaussian_kalman_tile_tc (D = 17..32, O <= 16): the same tile walk with the arithmetic on the tensor pipe.  The FFMA form above is bound
// by shared-memory operand bandwidth (one broadcast LDS.128 per 4 FMAs).  Here a WARP owns 16 tiles = the 16 rows of mma.sync.m16n8k8
// tiles; every per-tile vector lives in C-fragment layout (rows g / g + 8 of a lane quad, columns 2 tig, 2 tig + 1 of each 8-column block),
// and every product of the step is a small GEMM against an operator that all tiles share:
//     U_t = wo + Y Pr          [16 x 16] [16 x 16]      eta = a0 + U_t Po'     [16 x 16] [16 x 64]
//     V   = Z Linv'            [16 x 32] [32 x 32] (upper-triangular blocks only)      cur -= V U     [16 x 32] [32 x 64]
// A C fragment becomes the A fragment of the next product without moving data: with the k slots of a k8-step permuted as (tig <-> column
// 2 tig, tig + 4 <-> column 2 tig + 1) a thread already holds exactly its A entries; the operators are stored once per CTA in shared
// memory as B fragments in that slot order, split into TF32 hi and lo (one LDS.128 per (k8, n8) block and lane).  3xTF32
// (Ah Bh + Al Bh + Ah Bl) keeps fp32-class accuracy.  Not bit-identical to the level-by-level path (different summation order); the
// tests hold it to the same 1e-5 against the float64 oracle.
constexpr int KTC_CONST_BLOCKS = 2 * 2 + 2 * 8;   // Pr: 2 x 2 blocks, Po': 2 x 8
constexpr int KTC_LEVEL_BLOCKS = 4 * 4 + 4 * 8;   // Linv': 4 x 4, -U: 4 x 8
__host__ __device__ constexpr size_t ktc_smem_bytes(int K) {
    return (size_t)(KTC_CONST_BLOCKS + K * KTC_LEVEL_BLOCKS) * 512 + (64 + 16 + 8) * 4;
}
__device__ __forceinline__ void ktc_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
struct KtcFrag {
    uint32_t hi[4], lo[4];
};
// C fragment (rows g, g + 8; columns 2 tig, 2 tig + 1) -> split A fragment in the permuted slot order
__device__ __forceinline__ KtcFrag ktc_split(float c0, float c1, float c2, float c3) {
    KtcFrag f;
    const float v[4] = {c0, c2, c1, c3};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f.hi[i] = (__float_as_uint(v[i]) + 0x1000u) & 0xFFFFE000u;  // round to nearest TF32 (ties away), see hmm_flow.cu
        f.lo[i] = __float_as_uint(v[i] - __uint_as_float(f.hi[i]));
    }
    return f;
}
__device__ __forceinline__ void ktc_mma3(float (&d)[4], const KtcFrag& a, const float4& b) {
    ktc_mma(d, a.lo, __float_as_uint(b.x), __float_as_uint(b.y));
    ktc_mma(d, a.hi, __float_as_uint(b.z), __float_as_uint(b.w));
    ktc_mma(d, a.hi, __float_as_uint(b.x), __float_as_uint(b.y));
}
__device__ __forceinline__ float ktc_quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

template <int KTC_WARPS>
__global__ void __launch_bounds__(KTC_WARPS * 32, KTC_WARPS <= 6 ? 2 : 1) gaussian_kalman_tile_tc(KalmanTileArgs a) {
    constexpr int DT = 32, OT = 16, KTC_THREADS = KTC_WARPS * 32;
    extern __shared__ __align__(16) unsigned char ktc_smem[];
    float4* sB = reinterpret_cast<float4*>(ktc_smem);  // [block][lane] {hi(k0), hi(k0 + 1), lo(k0), lo(k0 + 1)}: B fragments
    const int K = a.K;
    float* sa0 = reinterpret_cast<float*>(ktc_smem + (size_t)(KTC_CONST_BLOCKS + K * KTC_LEVEL_BLOCKS) * 512);  // [64] padded
    float* swo = sa0 + 64;                                                                                      // [16]
    float* sk = swo + 16;                                                                                       // konst per level
    const int64_t b = blockIdx.y;
    const int D = a.D, O = a.O, n = 2 * D;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tig = lane & 3;
    // ---- operators -> B fragments.  Block order: Pr (kb, nb) 2 x 2 | Po' 2 x 8 | per level: Linv' 4 x 4 | -U 4 x 8
    const int nblocks = KTC_CONST_BLOCKS + K * KTC_LEVEL_BLOCKS;
    for (int idx = tid; idx < nblocks * 32; idx += KTC_THREADS) {
        const int blk = idx >> 5, ln = idx & 31, gg = ln >> 2, tt = ln & 3;
        float v0 = 0.f, v1 = 0.f;
        auto fetch = [&](auto elem, int kb, int nb) {
            const int k0 = 8 * kb + 2 * tt, col = 8 * nb + gg;
            v0 = elem(k0, col);
            v1 = elem(k0 + 1, col);
        };
        if (blk < KTC_CONST_BLOCKS && a.eta_in) {  // given elements: the observation operators are not used
        } else if (blk < 4) {  // B[k = o_in][n = o_out] = Pr[o_in][o_out]
            fetch([&](int k, int c) { return (k < O && c < O) ? a.Pr[(b * O + k) * O + c] : 0.f; }, blk >> 1, blk & 1);
        } else if (blk < KTC_CONST_BLOCKS) {  // B[k = o][n = padded row i] = Po[i][o]
            const int q = blk - 4;
            fetch([&](int k, int c) {
                const int half = c / DT, kk = c % DT;
                return (kk < D && k < O) ? a.Po[(b * n + half * D + kk) * O + k] : 0.f;
            }, q >> 3, q & 7);
        } else {
            const int q = blk - KTC_CONST_BLOCKS, l = q / KTC_LEVEL_BLOCKS, r = q % KTC_LEVEL_BLOCKS;
            const size_t op = ((size_t)b * a.NL + l) * 2;
            if (r < 16) {  // B[k][n = r] = Linv[r][k]
                const float* Li = a.opLinv + op * DT * DT;
                fetch([&](int k, int c) { return __ldg(Li + c * DT + k); }, r >> 2, r & 3);
            } else {  // B[k = r][n = c] = -U[r][c]
                const float* U = a.opU + op * DT * 2 * DT;
                fetch([&](int k, int c) { return -__ldg(U + k * 2 * DT + c); }, (r - 16) >> 3, (r - 16) & 7);
            }
        }
        const float h0 = __uint_as_float((__float_as_uint(v0) + 0x1000u) & 0xFFFFE000u), h1 = __uint_as_float((__float_as_uint(v1) + 0x1000u) & 0xFFFFE000u);
        sB[idx] = make_float4(h0, h1, v0 - h0, v1 - h1);
    }
    const bool given = a.eta_in != nullptr;
    if (!given) {
        for (int i = tid; i < 64; i += KTC_THREADS) sa0[i] = (i % DT) < D ? a.a0[b * n + (i / DT) * D + i % DT] : 0.f;
        for (int i = tid; i < 16; i += KTC_THREADS) swo[i] = i < O ? a.wo[b * O + i] : 0.f;
    }
    if (tid < K) sk[tid] = a.konst[((size_t)b * a.NL + tid) * 2];
    __syncthreads();

    const int64_t tile_base = ((int64_t)blockIdx.x * KTC_WARPS + warp) * 16;
    if (tile_base >= a.ntiles) return;
    const int64_t tile_r[2] = {tile_base + g, tile_base + g + 8};
    const bool live[2] = {tile_r[0] < a.ntiles, tile_r[1] < a.ntiles};
    const float* yp[2];
    const float* ep[2];
    const float* cp[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int64_t t_first = b * a.T + ((live[h] ? tile_r[h] : tile_base) << K);  // a missing tile repeats tile 0 of the warp (not stored)
        yp[h] = given ? nullptr : a.y + t_first * O;
        ep[h] = given ? a.eta_in + t_first * n : nullptr;
        cp[h] = given ? a.c_in + t_first : nullptr;
    }
    const float c0 = given ? 0.f : a.c0[b];
    const bool yvec = !given && (O == OT) && (reinterpret_cast<uintptr_t>(a.y) % 8 == 0);
    const bool evec = given && (D % 2 == 0) && (reinterpret_cast<uintptr_t>(a.eta_in) % 8 == 0);
    const float4* sPr = sB;
    const float4* sPo = sB + 4 * 32;
    float stk[KT_KMAX][34];  // pending left operands per level: 8 C fragments + the two scalars (local memory, run-time level index)
    float cur[8][4], cc[2];
    float yv[2][2][2];  // [row half][k8-step][column 2 tig, 2 tig + 1]
    auto load_y = [&](int step) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const float* src = yp[h] + (size_t)step * O + 8 * s + 2 * tig;
                if (yvec) {
                    const float2 v = __ldg(reinterpret_cast<const float2*>(src));
                    yv[h][s][0] = v.x;
                    yv[h][s][1] = v.y;
                } else {
                    yv[h][s][0] = (8 * s + 2 * tig) < O ? __ldg(src) : 0.f;
                    yv[h][s][1] = (8 * s + 2 * tig + 1) < O ? __ldg(src + 1) : 0.f;
                }
            }
    };
    if (!given) load_y(0);
    const int nsteps = 1 << K;
#pragma unroll 1
    for (int i = 0; i < nsteps; ++i) {
      if (given) {
        // ---- element of step i as given: rows g / g + 8 of the C fragments straight from eta (padded column 32 half + kk <-> half D + kk)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float* e = ep[h] + (size_t)i * n;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int pc = 8 * j + 2 * tig, half = pc / DT, kk = pc % DT;
                float x0 = 0.f, x1 = 0.f;
                if (evec) {
                    if (kk < D) {
                        const float2 v = __ldg(reinterpret_cast<const float2*>(e + half * D + kk));
                        x0 = v.x;
                        x1 = v.y;
                    }
                } else {
                    if (kk < D) x0 = __ldg(e + half * D + kk);
                    if (kk + 1 < D) x1 = __ldg(e + half * D + kk + 1);
                }
                cur[j][2 * h] = x0;
                cur[j][2 * h + 1] = x1;
            }
            cc[h] = __ldg(cp[h] + i);
        }
      } else {
        // ---- u = wo + y Pr  (C fragments of two 8-column blocks)
        float u[2][4];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            u[j][0] = u[j][2] = swo[8 * j + 2 * tig];
            u[j][1] = u[j][3] = swo[8 * j + 2 * tig + 1];
        }
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const KtcFrag ya = ktc_split(yv[0][s][0], yv[0][s][1], yv[1][s][0], yv[1][s][1]);
#pragma unroll
            for (int j = 0; j < 2; ++j) ktc_mma3(u[j], ya, sPr[(s * 2 + j) * 32 + lane]);
        }
        if (i + 1 < nsteps) load_y(i + 1);  // next step's observations travel while this step computes
        {
            float uu0 = 0.f, uu1 = 0.f;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                uu0 = fmaf(u[j][0], u[j][0], fmaf(u[j][1], u[j][1], uu0));
                uu1 = fmaf(u[j][2], u[j][2], fmaf(u[j][3], u[j][3], uu1));
            }
            cc[0] = c0 - 0.5f * ktc_quad_sum(uu0);
            cc[1] = c0 - 0.5f * ktc_quad_sum(uu1);
        }
        // ---- eta = a0 + u Po'
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            cur[j][0] = cur[j][2] = sa0[8 * j + 2 * tig];
            cur[j][1] = cur[j][3] = sa0[8 * j + 2 * tig + 1];
        }
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const KtcFrag ua = ktc_split(u[s][0], u[s][1], u[s][2], u[s][3]);
#pragma unroll
            for (int j = 0; j < 8; ++j) ktc_mma3(cur[j], ua, sPo[(s * 8 + j) * 32 + lane]);
        }
      }
        // ---- fold the finished subtrees: one combine per trailing one-bit of i
        int lvl = 0;
#pragma unroll 1
        for (int m = i; m & 1; m >>= 1, ++lvl) {
            const float4* sLi = sB + (size_t)(KTC_CONST_BLOCKS + lvl * KTC_LEVEL_BLOCKS) * 32;
            const float4* sU = sLi + 16 * 32;
            // z = x_b + y_a: blocks 4..7 of the stacked element + blocks 0..3 of cur; blocks 0..3 of the result start as x_a
            KtcFrag za[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float z[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    z[e] = stk[lvl][(4 + j) * 4 + e] + cur[j][e];
                    cur[j][e] = stk[lvl][j * 4 + e];
                }
                za[j] = ktc_split(z[0], z[1], z[2], z[3]);
            }
            const float cxy0 = stk[lvl][32] + cc[0], cxy1 = stk[lvl][33] + cc[1];
            // v = z Linv' (Linv lower triangular: k-block <= r-block)
            float v[4][4];
#pragma unroll
            for (int rb = 0; rb < 4; ++rb)
#pragma unroll
                for (int e = 0; e < 4; ++e) v[rb][e] = 0.f;
#pragma unroll
            for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                for (int rb = kb; rb < 4; ++rb) ktc_mma3(v[rb], za[kb], sLi[(kb * 4 + rb) * 32 + lane]);
            float vv0 = 0.f, vv1 = 0.f;
#pragma unroll
            for (int rb = 0; rb < 4; ++rb) {
                vv0 = fmaf(v[rb][0], v[rb][0], fmaf(v[rb][1], v[rb][1], vv0));
                vv1 = fmaf(v[rb][2], v[rb][2], fmaf(v[rb][3], v[rb][3], vv1));
            }
            // cur -= v U
#pragma unroll
            for (int rb = 0; rb < 4; ++rb) {
                const KtcFrag va = ktc_split(v[rb][0], v[rb][1], v[rb][2], v[rb][3]);
#pragma unroll
                for (int j = 0; j < 8; ++j) ktc_mma3(cur[j], va, sU[(rb * 8 + j) * 32 + lane]);
            }
            cc[0] = cxy0 + sk[lvl] + 0.5f * ktc_quad_sum(vv0);
            cc[1] = cxy1 + sk[lvl] + 0.5f * ktc_quad_sum(vv1);
        }
        if (i + 1 < nsteps) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) stk[lvl][j * 4 + e] = cur[j][e];
            stk[lvl][32] = cc[0];
            stk[lvl][33] = cc[1];
        }
    }
    // ---- one level-K element per tile: row g / g + 8, columns 8 j + 2 tig, + 1 (padded column 32 half + k <-> 2D-layout column half D + k)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if (!live[h]) continue;
        float* eo = a.eout + (b * a.out_T + tile_r[h]) * n;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int pc = 8 * j + 2 * tig, half = pc / DT, kk = pc % DT;
            if (kk < D) eo[half * D + kk] = cur[j][2 * h];
            if (kk + 1 < D) eo[half * D + kk + 1] = cur[j][2 * h + 1];
        }
        if (tig == 0) a.cout[b * a.out_T + tile_r[h]] = cc[h];
    }
}

// gaussian_eta_tc: eta[b,t] = P[b] w[b,t], c[b,t] = s[b,t] - |w[b,t]|^2 / 2 for dim, rank <= 64 as a batched [T x rank] [rank x dim] GEMM on
// the tensor pipe (3xTF32, same fragment conventions as gaussian_kalman_tile_tc): a warp owns 16 steps, P' sits in shared memory as split B
// fragments, the A fragments come straight from w (float2 per k8-step and row).  The thread-per-step FFMA kernel (gaussian_eta_c) is bound
// by its broadcast operand loads: 4.0 ms at BASELINE config 4's size.
constexpr int ETC_WARPS = 8;
constexpr int ETC_REPS = 8;
__global__ void __launch_bounds__(ETC_WARPS * 32) gaussian_eta_tc(const float* __restrict__ w, const float* __restrict__ P, const float* __restrict__ sc, int64_t T,
                                                                  int n, int r, float* __restrict__ eta, float* __restrict__ c) {
    __shared__ float4 sB[8 * 8 * 32];  // [kb][nb][lane] {hi(k0), hi(k0 + 1), lo(k0), lo(k0 + 1)}, B[k = rank index][col = row of P] = P[col][k]
    const int64_t b = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tig = lane & 3;
    const int KB = (r + 7) / 8, NB = (n + 7) / 8;
    const float* Pb = P + b * (int64_t)n * r;
    for (int idx = tid; idx < KB * NB * 32; idx += ETC_WARPS * 32) {
        const int blk = idx >> 5, ln = idx & 31, kb = blk / NB, nb = blk % NB;
        const int k0 = 8 * kb + 2 * (ln & 3), col = 8 * nb + (ln >> 2);
        const float v0 = (col < n && k0 < r) ? Pb[(int64_t)col * r + k0] : 0.f, v1 = (col < n && k0 + 1 < r) ? Pb[(int64_t)col * r + k0 + 1] : 0.f;
        const float h0 = __uint_as_float((__float_as_uint(v0) + 0x1000u) & 0xFFFFE000u), h1 = __uint_as_float((__float_as_uint(v1) + 0x1000u) & 0xFFFFE000u);
        sB[(kb * 8 + nb) * 32 + ln] = make_float4(h0, h1, v0 - h0, v1 - h1);
    }
    __syncthreads();
    // ETC_REPS tiles of 16 steps per warp: the operator fragments above are formed once per CTA
#pragma unroll 1
    for (int rep = 0; rep < ETC_REPS; ++rep) {
    const int64_t t0 = (((int64_t)blockIdx.x * ETC_REPS + rep) * ETC_WARPS + warp) * 16;
    if (t0 >= T) return;
    const int64_t tr[2] = {t0 + g, t0 + g + 8};
    const bool live[2] = {tr[0] < T, tr[1] < T};
    const float* wp[2] = {w + (b * T + (live[0] ? tr[0] : t0)) * r, w + (b * T + (live[1] ? tr[1] : t0)) * r};
    const bool wvec = (r % 2 == 0) && (reinterpret_cast<uintptr_t>(w) % 8 == 0);
    float acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
    float ww0 = 0.f, ww1 = 0.f;
#pragma unroll
    for (int kb = 0; kb < 8; ++kb) {
        if (kb < KB) {
            float x[2][2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k0 = 8 * kb + 2 * tig;
                if (wvec && k0 + 1 < r) {
                    const float2 v = __ldg(reinterpret_cast<const float2*>(wp[h] + k0));
                    x[h][0] = v.x;
                    x[h][1] = v.y;
                } else {
                    x[h][0] = k0 < r ? __ldg(wp[h] + k0) : 0.f;
                    x[h][1] = k0 + 1 < r ? __ldg(wp[h] + k0 + 1) : 0.f;
                }
            }
            ww0 = fmaf(x[0][0], x[0][0], fmaf(x[0][1], x[0][1], ww0));
            ww1 = fmaf(x[1][0], x[1][0], fmaf(x[1][1], x[1][1], ww1));
            const KtcFrag a = ktc_split(x[0][0], x[0][1], x[1][0], x[1][1]);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < NB) ktc_mma3(acc[j], a, sB[(kb * 8 + j) * 32 + lane]);
        }
    }
    ww0 = ktc_quad_sum(ww0);
    ww1 = ktc_quad_sum(ww1);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if (!live[h]) continue;
        const int64_t e = b * T + tr[h];
        float* eo = eta + e * n;
        const bool st2 = (n % 2 == 0) && (reinterpret_cast<uintptr_t>(eta) % 8 == 0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = 8 * j + 2 * tig;
            if (st2 && col + 1 < n) {
                *reinterpret_cast<float2*>(eo + col) = make_float2(acc[j][2 * h], acc[j][2 * h + 1]);  // a quad writes one 32-byte sector
            } else {
                if (col < n) eo[col] = acc[j][2 * h];
                if (col + 1 < n) eo[col + 1] = acc[j][2 * h + 1];
            }
        }
        if (tig == 0) c[e] = (sc ? sc[e] : 0.f) - 0.5f * (h ? ww1 : ww0);
    }
    }
}

// A side stream per device and host thread (the fork / join events belong to one call at a time).
struct GaussSide {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};
static int gauss_side_for_current_device(GaussSide** out) {
    static thread_local GaussSide table[64];
    int dev = 0;
    FB2_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return FB2_ERR_UNSUPPORTED;
    GaussSide& s = table[dev];
    if (!s.stream) {
        FB2_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        FB2_CUDA(cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming));
        FB2_CUDA(cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming));
    }
    *out = &s;
    return FB2_OK;
}

static bool kalman_tc_tiles(int D) {
    const char* tc_env = getenv("FB2_KALMAN_TC");  // 0: the FFMA tile kernel also for D > 16
    return D > 16 && !(tc_env && tc_env[0] == '0');
}
static int kalman_pref_levels(int D) {
    const char* e = getenv("FB2_KALMAN_K");  // tree levels fused into the tile kernel (tile = 2^K steps); measured best: 4 (FFMA tiles), 5 (tensor-core tiles)
    int k = e ? atoi(e) : (kalman_tc_tiles(D) ? 5 : 4);
    return k < 1 ? 1 : (k > KT_KMAX ? KT_KMAX : k);
}
static int kalman_levels_for(int64_t T, int D) {
    int k = kalman_pref_levels(D);
    while (k > 1 && (T >> k) < 1) --k;
    return k;
}
struct KalmanLayout {
    float *Lreg, *Ltail, *opU, *opLinv, *konst, *zeros, *scratch;
    int32_t* flag;
    float *lkE[2], *lkC[2], *remE[2], *remC[2];
    bool ok;
};
static KalmanLayout kalman_layout(void* workspace, size_t workspace_bytes, int64_t B, int64_t T, int D, int DT, bool count_only, size_t* used) {
    const int n = 2 * D, NL = homog_levels_count(T), K = KT_KMAX;
    // sized for the largest tile the environment can select, so that the size query does not depend on FB2_KALMAN_K
    const int64_t dK = (T >> 1) + 1, dK1 = (dK + 1) / 2, rem = (int64_t)1 << K;
    static char dummy;
    Arena arena(count_only ? static_cast<void*>(&dummy) : workspace, count_only ? ~(size_t)0 >> 1 : workspace_bytes);
    KalmanLayout L{};
    L.Lreg = arena.take<float>((size_t)B * NL * n * n);
    L.Ltail = arena.take<float>((size_t)B * NL * n * n);
    L.opU = arena.take<float>((size_t)B * NL * 2 * DT * 2 * DT);
    L.opLinv = arena.take<float>((size_t)B * NL * 2 * DT * DT);
    L.konst = arena.take<float>((size_t)B * NL * 2);
    L.zeros = arena.take<float>((size_t)n);
    L.scratch = arena.take<float>((size_t)B * 2 * n);
    L.flag = arena.take<int32_t>(1);
    const int64_t tt[2] = {dK, dK1};
    for (int q = 0; q < 2; ++q) {
        L.lkE[q] = arena.take<float>((size_t)B * tt[q] * n);
        L.lkC[q] = arena.take<float>((size_t)B * tt[q]);
        L.remE[q] = arena.take<float>((size_t)B * rem * n);
        L.remC[q] = arena.take<float>((size_t)B * rem);
    }
    L.ok = arena.ok() && NL < 48;
    if (used) *used = arena.used;
    return L;
}

// Source of the per-step elements: the observations + per-chain constants (GaussianHMM, eta_in == nullptr) or the elements themselves
// (eta_in[B,T,2D], c_in[B,T]: any time-homogeneous chain, e.g. the funsor rule for a Gaussian MarkovProduct).
template <int DT>
static int kalman_chain_run(const float* y, const float* Lam, const float* Pr, const float* wo, const float* a0, const float* Po, const float* c0,
                            int64_t B, int64_t T, int D, int O, float* Lo, float* eo, float* co, void* workspace, size_t workspace_bytes,
                            cudaStream_t stream, const float* eta_in = nullptr, const float* c_in = nullptr) {
    const int NL = homog_levels_count(T), n = 2 * D;
    const int K = kalman_levels_for(T, D);
    const int64_t nfull = T >> K, rem = T - (nfull << K), dK = nfull + (rem > 0 ? 1 : 0);
    const KalmanLayout lay = kalman_layout(workspace, workspace_bytes, B, T, D, DT, false, nullptr);
    if (!lay.ok) {
        set_error("gaussian_kalman_chain: workspace too small (%zu bytes)", workspace_bytes);
        return FB2_ERR_WORKSPACE;
    }
    HomogLevelsArgs la;
    la.Lam = Lam; la.T = T; la.D = D; la.NL = NL; la.Lo = Lo;
    la.Lreg = lay.Lreg; la.Ltail = lay.Ltail; la.opU = lay.opU; la.opLinv = lay.opLinv; la.konst = lay.konst;
    la.zeros = lay.zeros; la.scratch = lay.scratch; la.flag = lay.flag;
    FB2_CUDA(cudaMemsetAsync(la.flag, 0, sizeof(int32_t), stream));
    FB2_CUDA(cudaMemsetAsync(lay.zeros, 0, (size_t)n * 4, stream));
    // The precision operators are a serial chain of NL small factorisations per chain (one CTA each): the first K levels are all the tile
    // kernel needs, the rest runs on a side stream next to it and is joined before the level-K sweeps.
    GaussSide* side = nullptr;
    const char* side_env = getenv("FB2_KALMAN_SIDE");  // off by default: measured 2.92 -> 2.91 ms at C4 (the tile kernel fills the SMs either way)
    const bool use_side = NL > K && side_env && side_env[0] == '1' && gauss_side_for_current_device(&side) == FB2_OK;
    la.l_begin = 0; la.l_end = use_side ? K : NL + 1;
    {
        const int lst = launch_homog_levels<DT>(la, B, stream);
        if (lst != FB2_OK) return lst;
    }
    if (use_side) {
        FB2_CUDA(cudaEventRecord(side->fork, stream));
        FB2_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
        la.l_begin = K; la.l_end = NL + 1;
        gaussian_homog_levels<DT><<<(unsigned)B, 64, 0, side->stream>>>(la);
        cudaError_t le = cudaGetLastError();
        cudaEventRecord(side->join, side->stream);  // joined below even on failure: the caller's stream never runs ahead of the side stream
        if (le != cudaSuccess) {
            cudaStreamWaitEvent(stream, side->join, 0);
            set_error("gaussian_kalman_chain: side-stream launch failed: %s", cudaGetErrorString(le));
            return FB2_ERR_CUDA;
        }
    }

    // level-K arrays: [B, dK, n] / [B, dK] (the chain result itself when dK == 1)
    float* kE = dK == 1 ? eo : lay.lkE[0];
    float* kC = dK == 1 ? co : lay.lkC[0];
    {
        KalmanTileArgs ta;
        ta.y = y; ta.Pr = Pr; ta.wo = wo; ta.a0 = a0; ta.Po = Po; ta.c0 = c0; ta.eta_in = eta_in; ta.c_in = c_in;
        ta.opU = la.opU; ta.opLinv = la.opLinv; ta.konst = la.konst;
        ta.eout = kE; ta.cout = kC; ta.T = T; ta.ntiles = nfull; ta.out_T = dK; ta.D = D; ta.O = O; ta.K = K; ta.NL = NL;
        const size_t smem = ((size_t)kt_const_floats<DT, 16>() + (size_t)K * kt_level_floats<DT>()) * sizeof(float);
        if (DT == 32 && kalman_tc_tiles(D)) {
            const size_t smem_tc = ktc_smem_bytes(K);
            const char* w_env = getenv("FB2_KALMAN_TC_WARPS");
            const int nw = w_env ? atoi(w_env) : 12;  // C4: 4 warps 2.06 ms, 6: 1.88, 8: 1.85 (K = 5), 12: 1.75 (K = 5)
#define FB2_KTC_LAUNCH(NWv)                                                                                                              \
    do {                                                                                                                                 \
        FB2_CUDA(cudaFuncSetAttribute(gaussian_kalman_tile_tc<NWv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tc));         \
        gaussian_kalman_tile_tc<NWv><<<dim3((unsigned)ceil_div(nfull, NWv * 16), (unsigned)B), NWv * 32, smem_tc, stream>>>(ta);         \
    } while (0)
            if (nw == 6) FB2_KTC_LAUNCH(6);
            else if (nw == 8) FB2_KTC_LAUNCH(8);
            else if (nw == 12) FB2_KTC_LAUNCH(12);
            else FB2_KTC_LAUNCH(4);
#undef FB2_KTC_LAUNCH
        } else {
            // One tile per thread.  Two tiles per thread (NTL = 2: every operator load feeds 8 FMAs) was measured at C4: 4.1 ms against 2.9 ms --
            // 255 registers with spills and half as many threads cost more than the halved shared-memory traffic gains.
            FB2_CUDA(cudaFuncSetAttribute(gaussian_kalman_tile<DT, 16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            gaussian_kalman_tile<DT, 16, 1><<<dim3((unsigned)ceil_div(nfull, KT_THREADS), (unsigned)B), KT_THREADS, smem, stream>>>(ta);
        }
        const cudaError_t le = cudaGetLastError();
        // join here: everything enqueued from now on waits for the side stream, the tile kernel above runs next to it
        if (use_side) cudaStreamWaitEvent(stream, side->join, 0);
        if (le != cudaSuccess) {
            set_error("gaussian_kalman_chain: tile kernel launch failed: %s", cudaGetErrorString(le));
            return FB2_ERR_CUDA;
        }
    }
    HomogEtaArgs ea;
    ea.D = D; ea.op_stride = (int64_t)NL * 2;
    auto op_at = [&](int l, int which) {
        ea.opU = la.opU + (size_t)(l * 2 + which) * DT * 2 * DT;
        ea.opLinv = la.opLinv + (size_t)(l * 2 + which) * DT * DT;
        ea.konst = la.konst + (size_t)(l * 2 + which);
    };
    bool has_tail = false;
    // ---- the remainder tile (the last rem < 2^K steps): level by level with the operators of the whole chain; its parities are those of
    //      the whole chain (d_l = nfull 2^(K-l) + dl_l), its single level-K element lands in slot nfull
    if (rem > 0) {
        if (!eta_in) {
            gaussian_kalman_eta<16, 4><<<dim3((unsigned)ceil_div(rem, 128), (unsigned)B), 128, 0, stream>>>(y, Pr, wo, a0, Po, c0, rem, D, O, lay.remE[0],
                                                                                                        lay.remC[0], T, nfull << K);
            FB2_LAUNCH_CHECK();
        }
        int64_t dl = rem;
        int which = 0;
        for (int l = 0; l < K; ++l) {
            const bool odd = dl & 1;
            const int64_t lp = dl / 2, ndl = (dl + 1) / 2;
            const bool tail_pair = has_tail && !odd && lp >= 1;
            const int64_t lreg = lp - (tail_pair ? 1 : 0);
            const bool last = l == K - 1;
            ea.ein = lay.remE[which]; ea.cin = lay.remC[which]; ea.in_T = dl;
            if (l == 0 && eta_in) {  // given elements: the remainder's level-0 elements are the tail of the caller's arrays (chain stride T)
                ea.ein = eta_in + (nfull << K) * n; ea.cin = c_in + (nfull << K); ea.in_T = T;
            }
            ea.eout = last ? kE : lay.remE[which ^ 1];
            ea.cout = last ? kC : lay.remC[which ^ 1];
            ea.out_T = last ? dK : ndl;
            ea.out_p_off = last ? nfull : 0;
            if (lreg > 0 || odd) {
                op_at(l, 0);
                ea.p0 = 0; ea.count = lreg;
                ea.leftover_src = odd ? dl - 1 : -1;
                ea.leftover_dst = odd ? ndl - 1 + ea.out_p_off : -1;
                int64_t blocks = ceil_div(lreg > 0 ? lreg : 1, HE_THREADS);
                gaussian_homog_eta<DT><<<dim3((unsigned)blocks, (unsigned)B), HE_THREADS, 0, stream>>>(ea);
                FB2_LAUNCH_CHECK();
            }
            if (tail_pair) {
                op_at(l, 1);
                ea.p0 = lp - 1; ea.count = 1;
                ea.leftover_src = -1; ea.leftover_dst = -1;
                gaussian_homog_eta<DT><<<dim3(1, (unsigned)B), HE_THREADS, 0, stream>>>(ea);
                FB2_LAUNCH_CHECK();
            }
            if (!tail_pair && !has_tail && odd) has_tail = true;
            dl = ndl;
            which ^= 1;
        }
    }
    // ---- levels K .. NL-1 over the level-K array: level by level while a level is large (many CTAs per chain), then the rest of the tree in
    //      one launch (gaussian_homog_upper, one CTA per chain; FB2_KALMAN_UPPER=0: off, FB2_KALMAN_UPPER_MAX: largest level it takes)
    const char* up_env = getenv("FB2_KALMAN_UPPER");
    const char* upmax_env = getenv("FB2_KALMAN_UPPER_MAX");
    const bool upper_ok = !(up_env && up_env[0] == '0');
    const int64_t upper_max = upmax_env ? atoll(upmax_env) : 1024;
    const float *ce = kE, *cc = kC;
    int64_t d = dK, in_T = dK;
    int which = 1;
    for (int l = K; d > 1; ++l) {
        if (upper_ok && d <= upper_max) {
            // per-chain strides: a chain only ever touches its own regions (the CTAs of different chains are at different levels): the input
            // sits in lkE[which ^ 1] with stride d, odd levels above it go to lkE[which] (stride (d + 1) / 2), even ones back over the chain's own input
            HomogUpperArgs ua;
            ua.e0 = ce; ua.c0 = cc; ua.d0 = d;
            ua.eb[0] = lay.lkE[which]; ua.cb[0] = lay.lkC[which]; ua.cap[0] = (d + 1) / 2;
            ua.eb[1] = lay.lkE[which ^ 1]; ua.cb[1] = lay.lkC[which ^ 1]; ua.cap[1] = d;
            ua.eo = eo; ua.co = co; ua.opU = la.opU; ua.opLinv = la.opLinv; ua.konst = la.konst;
            ua.l0 = l; ua.NL = NL; ua.D = D; ua.has_tail = has_tail ? 1 : 0;
            gaussian_homog_upper<DT><<<(unsigned)B, HU_THREADS, 0, stream>>>(ua);
            FB2_LAUNCH_CHECK();
            break;
        }
        const int64_t npairs = d / 2, nd = (d + 1) / 2;
        const bool odd = d & 1;
        const bool tail_pair = has_tail && !odd;
        const int64_t reg_pairs = npairs - (tail_pair ? 1 : 0);
        float* de = nd == 1 ? eo : lay.lkE[which];
        float* dc = nd == 1 ? co : lay.lkC[which];
        ea.ein = ce; ea.cin = cc; ea.eout = de; ea.cout = dc; ea.in_T = in_T; ea.out_T = nd;
        ea.leftover_src = -1; ea.leftover_dst = -1; ea.out_p_off = 0;
        if (reg_pairs > 0) {
            op_at(l, 0);
            ea.p0 = 0; ea.count = reg_pairs;
            if (odd) {
                ea.leftover_src = d - 1;
                ea.leftover_dst = nd - 1;
            }
            int64_t blocks = ceil_div(reg_pairs, HE_THREADS);
            if (blocks > 65535) blocks = 65535;
            gaussian_homog_eta<DT><<<dim3((unsigned)blocks, (unsigned)B), HE_THREADS, 0, stream>>>(ea);
            FB2_LAUNCH_CHECK();
        }
        if (tail_pair) {
            op_at(l, 1);
            ea.p0 = npairs - 1; ea.count = 1;
            ea.leftover_src = -1; ea.leftover_dst = -1;
            gaussian_homog_eta<DT><<<dim3(1, (unsigned)B), HE_THREADS, 0, stream>>>(ea);
            FB2_LAUNCH_CHECK();
        }
        if (!tail_pair && !has_tail && odd) has_tail = true;
        ce = de; cc = dc; in_T = nd; d = nd;
        which ^= 1;
    }
    gaussian_poison_if_flag<<<(unsigned)ceil_div(B, 256), 256, 0, stream>>>(la.flag, co, B);
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

extern "C" int fb2_gaussian_info_to_sqrt_f32(const float* Lam, const float* eta, const float* c, int64_t N, int32_t dim, float rel_tol,
                                             float* w, float* P, float* s, int32_t* rank, void* stream) {
    int st = check_device();
    if (st != FB2_OK) return st;
    FB2_REQUIRE(Lam && eta && c && w && P && s, "null pointer");
    FB2_REQUIRE(N >= 0 && dim >= 1 && dim <= 2 * GMAXD, "bad extents N=%lld dim=%d", (long long)N, dim);
    if (N == 0) return FB2_OK;
    if (!(rel_tol > 0.f)) rel_tol = dim * 1.2e-7f;
    size_t smem = (size_t)I2S_WARPS * ((size_t)dim * (dim + 1) + dim) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set && smem > 48 * 1024) {
        FB2_CUDA(cudaFuncSetAttribute(gaussian_info_to_sqrt, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        attr_set = true;
    }
    int64_t blocks = ceil_div(N, I2S_WARPS);
    FB2_REQUIRE(blocks < (1ll << 31), "too many elements for one launch");
    gaussian_info_to_sqrt<<<(unsigned)blocks, I2S_WARPS * 32, smem, static_cast<cudaStream_t>(stream)>>>(Lam, eta, c, N, dim, rel_tol, w, P, s, rank);
    FB2_LAUNCH_CHECK();
    return FB2_OK;
}

extern "C" size_t fb2_gaussian_moment_match_workspace_bytes(int64_t N, int32_t K, int32_t dim) {
    return align_up((size_t)N * K * (dim + 1) * sizeof(float)) + 256;
}

extern "C" int fb2_gaussian_moment_match_f32(const float* logits, const float* Lam, const float* eta, const float* c, int64_t N, int32_t K,
                                             int32_t dim, float* Lo, float* eo, float* co, int32_t* status_flag, void* workspace,
                                             size_t workspace_bytes, void* stream) {
    int st = check_device();
    if (st != FB2_OK) return st;
    FB2_REQUIRE(logits && Lam && eta && c && Lo && eo && co, "null pointer");
    FB2_REQUIRE(N >= 0 && K >= 1 && dim >= 1 && dim <= 2 * GMAXD, "bad extents N=%lld K=%d dim=%d", (long long)N, K, dim);
    if (N == 0) return FB2_OK;
    Arena arena(workspace, workspace_bytes);
    float* scratch = arena.take<float>((size_t)N * K * (dim + 1));
    if (!arena.ok()) {
        set_error("gaussian_moment_match: workspace too small (%zu bytes)", workspace_bytes);
        return FB2_ERR_WORKSPACE;
    }
    const size_t smem = (size_t)MM_WARPS * ((size_t)2 * dim * (dim + 1) + (size_t)dim * 64 + 4 * dim) * sizeof(float);
    FB2_CUDA(cudaFuncSetAttribute(gaussian_moment_match, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = ceil_div(N, MM_WARPS);
    FB2_REQUIRE(blocks < (1ll << 31), "too many elements for one launch");
    gaussian_moment_match<<<(unsigned)blocks, MM_WARPS * 32, smem, static_cast<cudaStream_t>(stream)>>>(logits, Lam, eta, c, N, K, dim, scratch, Lo, eo,
                                                                                                       co, status_flag);
    FB2_LAUNCH_CHECK();
    return FB2_OK;
}

extern "C" int fb2_gaussian_kalman_eta_f32(const float* y, const float* Pr, const float* wo, const float* a0, const float* Po, const float* c0, int64_t B,
                                           int64_t T, int32_t D, int32_t O, float* eta, float* c, void* stream) {
    int st = check_device();
    if (st != FB2_OK) return st;
    FB2_REQUIRE(y && Pr && wo && a0 && Po && c0 && eta && c, "null pointer");
    FB2_REQUIRE(D >= 1 && D <= GMAXD && O >= 1 && O <= 32, "dimensions D=%d O=%d outside [1, 32]", D, O);
    FB2_REQUIRE(B >= 0 && B <= 65535 && T >= 1, "bad extents");
    if (B == 0) return FB2_OK;
    const int warps = O <= 16 ? 4 : 2;
    int64_t blocks = ceil_div(T, (int64_t)warps * 32 * 4);  // ~4 tiles of 32 steps per warp
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    if (O <= 16) gaussian_kalman_eta<16, 4><<<dim3((unsigned)blocks, (unsigned)B), 128, 0, static_cast<cudaStream_t>(stream)>>>(y, Pr, wo, a0, Po, c0, T, D, O, eta, c, T, 0);
    else gaussian_kalman_eta<32, 2><<<dim3((unsigned)blocks, (unsigned)B), 64, 0, static_cast<cudaStream_t>(stream)>>>(y, Pr, wo, a0, Po, c0, T, D, O, eta, c, T, 0);
    FB2_LAUNCH_CHECK();
    return FB2_OK;
}

extern "C" int fb2_gaussian_kalman_constants_f32(const float* F, const float* q_loc, const float* q_tril, const float* H, const float* r_loc,
                                                 const float* r_tril, int64_t B, int32_t D, int32_t O, float* Lam, float* Pr, float* wo, float* a0,
                                                 float* Po, float* c0, void* stream) {
    int st = check_device();
    if (st != FB2_OK) return st;
    FB2_REQUIRE(F && q_loc && q_tril && H && r_loc && r_tril && Lam && Pr && wo && a0 && Po && c0, "null pointer");
    FB2_REQUIRE(D >= 1 && D <= GMAXD && O >= 1 && O <= 32, "dimensions D=%d O=%d outside [1, 32]", D, O);
    FB2_REQUIRE(B >= 0 && B < (1ll << 31), "bad extents");
    if (B == 0) return FB2_OK;
    KalmanConstArgs a;
    a.F = F; a.q_loc = q_loc; a.q_tril = q_tril; a.H = H; a.r_loc = r_loc; a.r_tril = r_tril;
    a.Lam = Lam; a.Pr = Pr; a.wo = wo; a.a0 = a0; a.Po = Po; a.c0 = c0; a.D = D; a.O = O;
    gaussian_kalman_constants<<<(unsigned)B, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    FB2_LAUNCH_CHECK();
    return FB2_OK;
}

extern "C" size_t fb2_gaussian_kalman_chain_workspace_bytes(int64_t B, int64_t T, int32_t D) {
    if (B < 1 || T < 2 || D < 1 || D > GMAXD) return 0;
    const int DT = D <= 4 ? 4 : D <= 8 ? 8 : D <= 16 ? 16 : 32;
    size_t used = 0;
    kalman_layout(nullptr, 0, B, T, D, DT, true, &used);
    return used + 256;
}

extern "C" int fb2_gaussian_kalman_chain_f32(const float* y, const float* Lam, const float* Pr, const float* wo, const float* a0, const float* Po,
                                             const float* c0, int64_t B, int64_t T, int32_t D, int32_t O, float* Lo, float* eo, float* co,
                                             void* workspace, size_t workspace_bytes, void* stream_) {
    int st = check_device();
    if (st != FB2_OK) return st;
    FB2_REQUIRE(y && Lam && Pr && wo && a0 && Po && c0 && Lo && eo && co, "null pointer");
    FB2_REQUIRE(D >= 1 && D <= GMAXD && O >= 1 && O <= 16, "dimensions D=%d (1..32) O=%d (1..16) out of range", D, O);
    FB2_REQUIRE(B >= 0 && B <= 65535 && T >= 2, "bad extents B=%lld T=%lld (T >= 2)", (long long)B, (long long)T);
    if (B == 0) return FB2_OK;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (D <= 4) return kalman_chain_run<4>(y, Lam, Pr, wo, a0, Po, c0, B, T, D, O, Lo, eo, co, workspace, workspace_bytes, stream);
    if (D <= 8) return kalman_chain_run<8>(y, Lam, Pr, wo, a0, Po, c0, B, T, D, O, Lo, eo, co, workspace, workspace_bytes, stream);
    if (D <= 16) return kalman_chain_run<16>(y, Lam, Pr, wo, a0, Po, c0, B, T, D, O, Lo, eo, co, workspace, workspace_bytes, stream);
    return kalman_chain_run<32>(y, Lam, Pr, wo, a0, Po, c0, B, T, D, O, Lo, eo, co, workspace, workspace_bytes, stream);
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

extern "C" int fb2_gaussian_eta_f32(const float* w, const float* P, const float* s, int64_t B, int64_t T, int32_t dim, int32_t rank,
                                    float* eta, float* c, void* stream) {
    int st = check_device();
    if (st != FB2_OK) return st;
    FB2_REQUIRE(w && P && eta && c, "null pointer");
    FB2_REQUIRE(B >= 0 && T >= 0 && dim >= 1 && rank >= 0, "bad extents");
    if (B * T == 0) return FB2_OK;
    FB2_REQUIRE(dim <= 2 * GMAXD && B <= 65535, "gaussian_eta: dim=%d or B=%lld out of range", dim, (long long)B);
    {
        const char* tc_env = getenv("FB2_GAUSS_ETA_TC");  // 0: the thread-per-step FFMA kernel
        if (dim <= 64 && rank <= 64 && T >= 64 && !(tc_env && tc_env[0] == '0')) {
            gaussian_eta_tc<<<dim3((unsigned)ceil_div(T, ETC_WARPS * 16 * ETC_REPS), (unsigned)B), ETC_WARPS * 32, 0, static_cast<cudaStream_t>(stream)>>>(w, P, s, T, dim,
                                                                                                                                 rank, eta, c);
            FB2_LAUNCH_CHECK();
            return FB2_OK;
        }
    }
    const size_t smem = (size_t)dim * ((rank + 3) / 4 * 4) * sizeof(float);
    FB2_REQUIRE(smem <= 200 * 1024, "gaussian_eta: prec_sqrt of %d x %d does not fit in shared memory", dim, rank);
    const dim3 grid((unsigned)ceil_div(T, ETA_THREADS), (unsigned)B);
    cudaStream_t st_ = static_cast<cudaStream_t>(stream);
#define FB2_ETA_LAUNCH(NMAXv)                                                                                                   \
    do {                                                                                                                        \
        FB2_CUDA(cudaFuncSetAttribute(gaussian_eta_c<NMAXv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));          \
        gaussian_eta_c<NMAXv><<<grid, ETA_THREADS, smem, st_>>>(w, P, s, B, T, dim, rank, eta, c);                             \
    } while (0)
    if (dim <= 8) FB2_ETA_LAUNCH(8);
    else if (dim <= 16) FB2_ETA_LAUNCH(16);
    else if (dim <= 32) FB2_ETA_LAUNCH(32);
    else FB2_ETA_LAUNCH(64);
#undef FB2_ETA_LAUNCH
    FB2_LAUNCH_CHECK();
    return FB2_OK;
}

extern "C" size_t fb2_gaussian_chain_workspace_bytes(int64_t B, int64_t T, int32_t D) {
    int n = 2 * D;
    int64_t t1 = (T + 1) / 2, t2 = (t1 + 1) / 2;
    size_t per = 0;
    for (int64_t tt : {t1, t2})
        per += align_up((size_t)B * tt * n * n * 4) + align_up((size_t)B * tt * n * 4) + align_up((size_t)B * tt * 4);
    (void)gauss_elem_floats;
    const size_t tree = per + align_up(4) + 256;
    size_t homog = T >= 2 ? homog_workspace_bytes(B, T, D) : 0;
    if (T >= 2 && B >= 1) {  // the tile walk of the homogeneous chain (kalman_chain_run with given elements)
        const int DT = D <= 4 ? 4 : D <= 8 ? 8 : D <= 16 ? 16 : 32;
        size_t used = 0;
        kalman_layout(nullptr, 0, B, T, D, DT, true, &used);
        if (used + 256 > homog) homog = used + 256;
    }
    return tree > homog ? tree : homog;
}

static int gaussian_chain_impl(const float* Lam, const float* eta, const float* c, int64_t B, int64_t T, int32_t D, float* Lo, float* eo,
                               float* co, void* workspace, size_t workspace_bytes, void* stream_, bool homog);

extern "C" int fb2_gaussian_chain_f32(const float* Lam, const float* eta, const float* c, int64_t B, int64_t T, int32_t D,
                                      float* Lo, float* eo, float* co, void* workspace, size_t workspace_bytes,
                                      void* stream_) {
    return gaussian_chain_impl(Lam, eta, c, B, T, D, Lo, eo, co, workspace, workspace_bytes, stream_, false);
}

extern "C" size_t fb2_gaussian_chain_sample_workspace_bytes(int64_t B, int64_t T, int32_t D) {
    return homog_workspace_bytes(B, T, D, true);
}

extern "C" int fb2_gaussian_chain_homog_keep_f32(const float* Lam, const float* eta, const float* c, int64_t B, int64_t T, int32_t D,
                                                 float* Lo, float* eo, float* co, void* workspace, size_t workspace_bytes, void* stream_) {
    int st = check_device();
    if (st != FB2_OK) return st;
    FB2_REQUIRE(Lam && eta && c && Lo && eo && co, "null pointer");
    FB2_REQUIRE(D >= 1 && D <= GMAXD, "state dimension D=%d outside [1, %d]", D, GMAXD);
    FB2_REQUIRE(B >= 1 && B <= 65535 && T >= 2, "bad extents B=%lld T=%lld (the sampling path needs T >= 2)", (long long)B, (long long)T);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (D <= 4) return homog_chain_run<4>(Lam, eta, c, B, T, D, Lo, eo, co, workspace, workspace_bytes, stream, true);
    if (D <= 8) return homog_chain_run<8>(Lam, eta, c, B, T, D, Lo, eo, co, workspace, workspace_bytes, stream, true);
    if (D <= 16) return homog_chain_run<16>(Lam, eta, c, B, T, D, Lo, eo, co, workspace, workspace_bytes, stream, true);
    return homog_chain_run<32>(Lam, eta, c, B, T, D, Lo, eo, co, workspace, workspace_bytes, stream, true);
}

extern "C" int fb2_gaussian_chain_homog_backward_sample_f32(const float* eta, int64_t B, int64_t T, int32_t D, int64_t N, const float* noise,
                                                            float* x, void* workspace, size_t workspace_bytes, void* stream_) {
    int st = check_device();
    if (st != FB2_OK) return st;
    FB2_REQUIRE(eta && noise && x && workspace, "null pointer");
    FB2_REQUIRE(D >= 1 && D <= GMAXD, "state dimension D=%d outside [1, %d]", D, GMAXD);
    FB2_REQUIRE(B >= 1 && B <= 65535 && T >= 2 && N >= 1, "bad extents");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (D <= 4) return homog_sample_run<4>(eta, B, T, D, N, noise, x, workspace, workspace_bytes, stream);
    if (D <= 8) return homog_sample_run<8>(eta, B, T, D, N, noise, x, workspace, workspace_bytes, stream);
    if (D <= 16) return homog_sample_run<16>(eta, B, T, D, N, noise, x, workspace, workspace_bytes, stream);
    return homog_sample_run<32>(eta, B, T, D, N, noise, x, workspace, workspace_bytes, stream);
}

extern "C" int fb2_gaussian_chain_homog_f32(const float* Lam, const float* eta, const float* c, int64_t B, int64_t T, int32_t D,
                                            float* Lo, float* eo, float* co, void* workspace, size_t workspace_bytes,
                                            void* stream_) {
    return gaussian_chain_impl(Lam, eta, c, B, T, D, Lo, eo, co, workspace, workspace_bytes, stream_, true);
}

static int gaussian_chain_impl(const float* Lam, const float* eta, const float* c, int64_t B, int64_t T, int32_t D, float* Lo, float* eo,
                               float* co, void* workspace, size_t workspace_bytes, void* stream_, bool homog) {
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
    // time-homogeneous chains: per-level precision operators + eta-only sweeps (FB2_GAUSS_HOMOG_TREE=1 keeps the full tree)
    {
        const char* tree_env = getenv("FB2_GAUSS_HOMOG_TREE");
        if (homog && T >= 4 && B <= 65535 && !(tree_env && tree_env[0] == '1')) {
            // the tile walk (elements read from eta / c): first K levels per tile without touching memory (FB2_GAUSS_HOMOG_TILES=0: level by level)
            const char* tiles_env = getenv("FB2_GAUSS_HOMOG_TILES");
            if (T >= 64 && !(tiles_env && tiles_env[0] == '0')) {
                if (D <= 4) return kalman_chain_run<4>(nullptr, Lam, nullptr, nullptr, nullptr, nullptr, nullptr, B, T, D, 1, Lo, eo, co, workspace, workspace_bytes, stream, eta, c);
                if (D <= 8) return kalman_chain_run<8>(nullptr, Lam, nullptr, nullptr, nullptr, nullptr, nullptr, B, T, D, 1, Lo, eo, co, workspace, workspace_bytes, stream, eta, c);
                if (D <= 16) return kalman_chain_run<16>(nullptr, Lam, nullptr, nullptr, nullptr, nullptr, nullptr, B, T, D, 1, Lo, eo, co, workspace, workspace_bytes, stream, eta, c);
                return kalman_chain_run<32>(nullptr, Lam, nullptr, nullptr, nullptr, nullptr, nullptr, B, T, D, 1, Lo, eo, co, workspace, workspace_bytes, stream, eta, c);
            }
            if (D <= 4) return homog_chain_run<4>(Lam, eta, c, B, T, D, Lo, eo, co, workspace, workspace_bytes, stream);
            if (D <= 8) return homog_chain_run<8>(Lam, eta, c, B, T, D, Lo, eo, co, workspace, workspace_bytes, stream);
            if (D <= 16) return homog_chain_run<16>(Lam, eta, c, B, T, D, Lo, eo, co, workspace, workspace_bytes, stream);
            return homog_chain_run<32>(Lam, eta, c, B, T, D, Lo, eo, co, workspace, workspace_bytes, stream);
        }
    }
    bool first_level = true;
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
        const bool inv = homog && first_level;  // level 1 of a homogeneous chain: Lam[b] is shared by all time steps
        st = launch_pair(cL, ce, cc, cur_s0, 2, inv ? cL : cL + (size_t)n * n, ce + n, cc + 1, cur_s0, 2, B, half, D, dL, de, dc, o_s0,
                         flag, stream, inv);
        if (st != FB2_OK) return st;
        if (nd != half) {
            if (inv) {  // odd leftover: Lam from [b], eta / c from [b, d-1]
                gaussian_copy_elements<<<(unsigned)ceil_div(B * per, 256), 256, 0, stream>>>(cL, ce, cc, cur_s0, d - 1, dL, de, dc, o_s0,
                                                                                           half, B, n, 1, 0);
            } else {
                gaussian_copy_elements<<<(unsigned)ceil_div(B * per, 256), 256, 0, stream>>>(cL, ce, cc, cur_s0, d - 1, dL, de, dc, o_s0,
                                                                                           half, B, n, cur_s0, d - 1);
            }
            FB2_LAUNCH_CHECK();
        }
        first_level = false;
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
