// Strided batched semiring matmul (one scan level) and the chain product built from it.
// Reference semantics: funsor/cnf.py:335-379 -> einsum/numpy_log.py:24-47 (LOG),
// cnf.py:312-324 -> tensor.py:700-726,313-330 (MAX), sum_product.py:690-701 (tree order).
#include <stdlib.h>

#include "common.cuh"

namespace fb2 {

struct MatmulArgs {
    const float* x;
    const float* y;
    float* out;
    int32_t* argk;
    int64_t n1;                  // inner batch extent (n = n0*n1 linear)
    int64_t os0;                 // out element stride of the outer batch index (inner is I*J)
    int64_t xs0, xs1, xsi, xsk;  // element strides
    int64_t ys0, ys1, ysk, ysj;
    int I, K, J;
    int tiles_i, tiles_j;
};

// ------------------------------------------------------------------ small matrices: one thread per output
template <int SEMI, bool ARG>
__global__ void __launch_bounds__(256) semiring_matmul_small(MatmulArgs a, int64_t total) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int j = (int)(idx % a.J);
    int64_t r = idx / a.J;
    int i = (int)(r % a.I);
    int64_t n = r / a.I;
    int64_t b0 = n / a.n1, b1 = n % a.n1;
    const float* xr = a.x + b0 * a.xs0 + b1 * a.xs1 + (int64_t)i * a.xsi;
    const float* yc = a.y + b0 * a.ys0 + b1 * a.ys1 + (int64_t)j * a.ysj;
    if (SEMI == FB2_SEMIRING_LOG) {
        float sx = -FLT_MAX, sy = -FLT_MAX;
        for (int k = 0; k < a.K; ++k) {
            sx = fmaxf(sx, xr[(int64_t)k * a.xsk]);
            sy = fmaxf(sy, yc[(int64_t)k * a.ysk]);
        }
        float acc = 0.f;
        for (int k = 0; k < a.K; ++k) acc += expf(xr[(int64_t)k * a.xsk] - sx) * expf(yc[(int64_t)k * a.ysk] - sy);
        a.out[b0 * a.os0 + (b1 * a.I + i) * (int64_t)a.J + j] = (sx + sy) + logf(acc);
    } else {
        float best = kNegInf;
        int arg = 0;
        for (int k = 0; k < a.K; ++k) {
            float v = xr[(int64_t)k * a.xsk] + yc[(int64_t)k * a.ysk];
            if (k == 0 || v > best) {
                best = v;
                arg = k;
            }
        }
        int64_t o = b0 * a.os0 + (b1 * a.I + i) * (int64_t)a.J + j;
        a.out[o] = best;
        if (ARG) a.argk[o] = arg;
    }
}

// ------------------------------------------------------------------ tiled kernel: 64x64 outputs per CTA
constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;

template <int SEMI, bool ARG>
__global__ void __launch_bounds__(256) semiring_matmul_tiled(MatmulArgs a) {
    __shared__ __align__(16) float xs[BK][BM + 4];
    __shared__ __align__(16) float ys[BK][BN + 4];
    __shared__ float sx[BM], sy[BN];

    int64_t blk = blockIdx.x;
    int tj = (int)(blk % a.tiles_j);
    blk /= a.tiles_j;
    int ti = (int)(blk % a.tiles_i);
    int64_t n = blk / a.tiles_i;
    int64_t b0 = n / a.n1, b1 = n % a.n1;
    const float* xb = a.x + b0 * a.xs0 + b1 * a.xs1;
    const float* yb = a.y + b0 * a.ys0 + b1 * a.ys1;
    const int i0 = ti * BM, j0 = tj * BN;
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;  // tx -> columns, ty -> rows

    if (SEMI == FB2_SEMIRING_LOG) {
        // pass 1: clamped row maxima of x and column maxima of y (numpy_log.py:27-32)
        // 4 threads per row (x) / per column (y)
        {
            int r = tid / 4, part = tid % 4;
            float m = -FLT_MAX;
            if (i0 + r < a.I) {
                const float* p = xb + (int64_t)(i0 + r) * a.xsi;
                for (int k = part; k < a.K; k += 4) m = fmaxf(m, p[(int64_t)k * a.xsk]);
            }
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
            if (part == 0) sx[r] = m;
        }
        {
            // columns: consecutive threads -> consecutive j for coalescing; 4 k-parts
            int c = tid % 64, part = tid / 64;
            float m = -FLT_MAX;
            if (j0 + c < a.J) {
                const float* p = yb + (int64_t)(j0 + c) * a.ysj;
                for (int k = part; k < a.K; k += 4) m = fmaxf(m, p[(int64_t)k * a.ysk]);
            }
            // reduce the 4 parts through shared memory (reuse ys rows 0..3)
            ys[part][c] = m;
            __syncthreads();
            if (tid < 64) sy[tid] = fmaxf(fmaxf(ys[0][tid], ys[1][tid]), fmaxf(ys[2][tid], ys[3][tid]));
        }
        __syncthreads();
    }

    float acc[TM][TN];
    int arg[TM][TN];
#pragma unroll
    for (int r = 0; r < TM; ++r)
#pragma unroll
        for (int c = 0; c < TN; ++c) {
            acc[r][c] = (SEMI == FB2_SEMIRING_LOG) ? 0.f : kNegInf;
            arg[r][c] = 0;
        }

    for (int k0 = 0; k0 < a.K; k0 += BK) {
        // stage tiles (out-of-range -> semiring zero)
#pragma unroll
        for (int r = 0; r < (BM * BK) / 256; ++r) {
            int idx = tid + 256 * r;
            int k = idx % BK, i = idx / BK;
            float v;
            if (i0 + i < a.I && k0 + k < a.K) {
                v = xb[(int64_t)(i0 + i) * a.xsi + (int64_t)(k0 + k) * a.xsk];
                if (SEMI == FB2_SEMIRING_LOG) v = expf(v - sx[i]);
            } else {
                v = (SEMI == FB2_SEMIRING_LOG) ? 0.f : kNegInf;
            }
            xs[k][i] = v;
        }
#pragma unroll
        for (int r = 0; r < (BN * BK) / 256; ++r) {
            int idx = tid + 256 * r;
            int j = idx % BN, k = idx / BN;
            float v;
            if (j0 + j < a.J && k0 + k < a.K) {
                v = yb[(int64_t)(k0 + k) * a.ysk + (int64_t)(j0 + j) * a.ysj];
                if (SEMI == FB2_SEMIRING_LOG) v = expf(v - sy[j]);
            } else {
                v = (SEMI == FB2_SEMIRING_LOG) ? 0.f : kNegInf;
            }
            ys[k][j] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float4 xv = *reinterpret_cast<const float4*>(&xs[k][ty * TM]);
            float4 yv = *reinterpret_cast<const float4*>(&ys[k][tx * TN]);
            float xr[TM] = {xv.x, xv.y, xv.z, xv.w};
            float yr[TN] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
            for (int r = 0; r < TM; ++r)
#pragma unroll
                for (int c = 0; c < TN; ++c) {
                    if (SEMI == FB2_SEMIRING_LOG) {
                        acc[r][c] = fmaf(xr[r], yr[c], acc[r][c]);
                    } else {
                        float v = xr[r] + yr[c];
                        if (ARG) {
                            // strict '>' keeps the first maximiser; k0+k==0 seeds the slot
                            if (v > acc[r][c] || (k0 + k) == 0) {
                                acc[r][c] = v;
                                arg[r][c] = k0 + k;
                            }
                        } else {
                            acc[r][c] = fmaxf(acc[r][c], v);
                        }
                    }
                }
        }
        __syncthreads();
    }

    const int64_t obase = b0 * a.os0 + b1 * (int64_t)a.I * a.J;
    float* ob = a.out + obase;
#pragma unroll
    for (int r = 0; r < TM; ++r) {
        int i = i0 + ty * TM + r;
        if (i >= a.I) continue;
#pragma unroll
        for (int c = 0; c < TN; ++c) {
            int j = j0 + tx * TN + c;
            if (j >= a.J) continue;
            float v = acc[r][c];
            if (SEMI == FB2_SEMIRING_LOG) v = (sx[ty * TM + r] + sy[tx * TN + c]) + logf(v);
            ob[(int64_t)i * a.J + j] = v;
            if (ARG) a.argk[obase + (int64_t)i * a.J + j] = arg[r][c];
        }
    }
}

// log_matmul_tc.cu: tcgen05 path for the log semiring (needs (n0*n1)*(I+J) floats of scratch for the shifts)
int launch_log_matmul_tc(const float* x, const int64_t* xs, const float* y, const int64_t* ys, float* out, int64_t n0, int64_t n1, int I,
                         int K, int J, int64_t out_s0, float* sx_sy_workspace, cudaStream_t stream);
size_t log_matmul_tc_workspace_floats(int64_t n, int I, int J);
static bool tc_matmul_disabled() {
    const char* e = getenv("FB2_DISABLE_TC_MATMUL");
    return e && e[0] == '1';
}

int launch_semiring_matmul(int semiring, const float* x, const int64_t* xs, const float* y, const int64_t* ys,
                           float* out, int32_t* argk, int64_t n0, int64_t n1, int I, int K, int J,
                           cudaStream_t stream, int64_t out_s0 = -1, float* shift_ws = nullptr) {
    if (n0 * n1 == 0 || I == 0 || J == 0) return FB2_OK;
    if (semiring == FB2_SEMIRING_LOG && shift_ws != nullptr && !tc_matmul_disabled()) {
        int st = launch_log_matmul_tc(x, xs, y, ys, out, n0, n1, I, K, J, out_s0, shift_ws, stream);
        if (st != FB2_ERR_UNSUPPORTED) return st;
    }
    MatmulArgs a;
    a.x = x; a.y = y; a.out = out; a.argk = argk; a.n1 = n1;
    a.os0 = out_s0 >= 0 ? out_s0 : n1 * (int64_t)I * J;
    a.xs0 = xs[0]; a.xs1 = xs[1]; a.xsi = xs[2]; a.xsk = xs[3];
    a.ys0 = ys[0]; a.ys1 = ys[1]; a.ysk = ys[2]; a.ysj = ys[3];
    a.I = I; a.K = K; a.J = J;
    a.tiles_i = (int)ceil_div(I, BM);
    a.tiles_j = (int)ceil_div(J, BN);
    int64_t n = n0 * n1;
    bool small = (I < 32 || J < 32);
    if (small) {
        int64_t total = n * I * J;
        int64_t blocks = ceil_div(total, 256);
        FB2_REQUIRE(blocks < (1ll << 31), "semiring_matmul: problem too large for one launch");
        if (semiring == FB2_SEMIRING_LOG)
            semiring_matmul_small<FB2_SEMIRING_LOG, false><<<(unsigned)blocks, 256, 0, stream>>>(a, total);
        else if (argk)
            semiring_matmul_small<FB2_SEMIRING_MAX, true><<<(unsigned)blocks, 256, 0, stream>>>(a, total);
        else
            semiring_matmul_small<FB2_SEMIRING_MAX, false><<<(unsigned)blocks, 256, 0, stream>>>(a, total);
    } else {
        int64_t blocks = n * a.tiles_i * a.tiles_j;
        FB2_REQUIRE(blocks < (1ll << 31), "semiring_matmul: problem too large for one launch");
        if (semiring == FB2_SEMIRING_LOG)
            semiring_matmul_tiled<FB2_SEMIRING_LOG, false><<<(unsigned)blocks, 256, 0, stream>>>(a);
        else if (argk)
            semiring_matmul_tiled<FB2_SEMIRING_MAX, true><<<(unsigned)blocks, 256, 0, stream>>>(a);
        else
            semiring_matmul_tiled<FB2_SEMIRING_MAX, false><<<(unsigned)blocks, 256, 0, stream>>>(a);
    }
    FB2_LAUNCH_CHECK();
    return FB2_OK;
}

// strided batched matrix copy: dst[b] (contiguous S*S, batch stride dst_sb) = src[b] (strided)
__global__ void copy_matrices(const float* src, int64_t sb, int64_t si, int64_t sj, float* dst, int64_t dst_sb,
                              int64_t B, int S) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t per = (int64_t)S * S;
    if (idx >= B * per) return;
    int64_t b = idx / per;
    int r = (int)(idx % per);
    dst[b * dst_sb + r] = src[b * sb + (int64_t)(r / S) * si + (int64_t)(r % S) * sj];
}

int launch_copy_matrices(const float* src, int64_t sb, int64_t si, int64_t sj, float* dst, int64_t dst_sb, int64_t B,
                         int S, cudaStream_t stream) {
    int64_t total = B * (int64_t)S * S;
    if (total == 0) return FB2_OK;
    copy_matrices<<<(unsigned)ceil_div(total, 256), 256, 0, stream>>>(src, sb, si, sj, dst, dst_sb, B, S);
    FB2_LAUNCH_CHECK();
    return FB2_OK;
}

}  // namespace fb2

using namespace fb2;

extern "C" size_t fb2_semiring_matmul_workspace_bytes(int64_t n0, int64_t n1, int32_t I, int32_t K, int32_t J) {
    (void)K;
    return align_up(log_matmul_tc_workspace_floats(n0 * n1, I, J) * sizeof(float)) + 256;
}

extern "C" int fb2_semiring_matmul_f32(int semiring, const float* x, const int64_t x_strides[4], const float* y,
                                       const int64_t y_strides[4], float* out, int32_t* argk, int64_t n0,
                                       int64_t n1, int32_t I, int32_t K, int32_t J, void* workspace,
                                       size_t workspace_bytes, void* stream) {
    int st = check_device();
    if (st != FB2_OK) return st;
    FB2_REQUIRE(semiring == FB2_SEMIRING_LOG || semiring == FB2_SEMIRING_MAX, "unknown semiring %d", semiring);
    FB2_REQUIRE(n0 >= 0 && n1 >= 0 && I >= 0 && J >= 0 && K >= 1, "bad extents n0=%lld n1=%lld I=%d K=%d J=%d",
                (long long)n0, (long long)n1, I, K, J);
    FB2_REQUIRE(x && y && out, "null pointer");
    FB2_REQUIRE(!(argk && semiring != FB2_SEMIRING_MAX), "argk is only defined for the max-plus semiring");
    float* shift_ws = nullptr;
    if (workspace != nullptr && workspace_bytes >= log_matmul_tc_workspace_floats(n0 * n1, I, J) * sizeof(float))
        shift_ws = static_cast<float*>(workspace);
    return launch_semiring_matmul(semiring, x, x_strides, y, y_strides, out, argk, n0, n1, I, K, J,
                                  static_cast<cudaStream_t>(stream), -1, shift_ws);
}

// ---------------------------------------------------------------------------------------------
// Fused tail of the scan: once the elements of a chain fit in shared memory (d S^2 fp32 x 1.5), ONE CTA per chain finishes
// all remaining levels of the reference's tree (sum_product.py:690-701) without leaving the SM.  Small problems -- BASELINE C1
// (T=100, S=16: 102 KB) is one -- are launch-latency bound on the level-by-level path (~3 launches x 7 levels); here they are a
// single launch.  Arithmetic per pair as einsum/numpy_log.py:24-47: row / column maxima clamped to finfo.min, exp applied in
// place (an element plays exactly one role per level), plain dot products, log + shifts.  Max-plus: max_k (x + y), exact.
constexpr size_t TAIL_SMEM_LIMIT = 200 * 1024;
constexpr int TAIL_THREADS = 1024;
static size_t tail_smem_bytes(int64_t d, int S) {
    const size_t per = (size_t)S * S;
    return ((size_t)d * per + (size_t)((d + 1) / 2) * per + 2 * (size_t)(d / 2) * S) * sizeof(float);
}

// index helper: walks idx = tid, tid + NT, ... as (p, i, j) with idx = (p * S + i) * S + j without a division per step
struct TailIdx {
    int p, i, j, di, dj;
    __device__ TailIdx(int tid, int S, int nt) {
        const int r = tid / S;
        j = tid % S;
        i = r % S;
        p = r / S;
        di = nt / S;
        dj = nt % S;
    }
    __device__ void next(int S) {
        j += dj;
        i += di;
        if (j >= S) {
            j -= S;
            ++i;
        }
        while (i >= S) {
            i -= S;
            ++p;
        }
    }
};

template <int SEMI>
__global__ void __launch_bounds__(TAIL_THREADS) chain_product_tail(const float* cur, int64_t s_b, int64_t s_t, int64_t s_i, int64_t s_j, int d0, int S,
                                                                   float* out) {
    extern __shared__ __align__(16) float tail_smem[];
    const int per = S * S;
    float* A = tail_smem;
    float* Bf = A + (size_t)d0 * per;
    float* sx = Bf + (size_t)((d0 + 1) / 2) * per;
    float* sy = sx + (size_t)(d0 / 2) * S;
    const int tid = threadIdx.x;
    const float* src = cur + blockIdx.x * s_b;
    {
        // gather the level into shared memory; four independent loads in flight per thread
        const int total = d0 * per;
        TailIdx w(tid, S, TAIL_THREADS);
        int idx = tid;
        for (; idx + 3 * TAIL_THREADS < total; idx += 4 * TAIL_THREADS) {
            float v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                v[q] = __ldg(src + (int64_t)w.p * s_t + (int64_t)w.i * s_i + (int64_t)w.j * s_j);
                w.next(S);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) A[idx + q * TAIL_THREADS] = v[q];
        }
        for (; idx < total; idx += TAIL_THREADS) {
            A[idx] = __ldg(src + (int64_t)w.p * s_t + (int64_t)w.i * s_i + (int64_t)w.j * s_j);
            w.next(S);
        }
    }
    __syncthreads();
    int d = d0;
    while (d > 1) {
        const int half = d / 2, nd = (d + 1) / 2;
        if (SEMI == FB2_SEMIRING_LOG) {
            // shifts: sx[p][i] = max_k x[i][k], sy[p][j] = max_k y[k][j], clamped to finfo.min
            for (int r = tid; r < half * S; r += TAIL_THREADS) {
                const int p = r / S, q = r % S;
                const float* x = A + (size_t)(2 * p) * per + q * S;
                const float* y = A + (size_t)(2 * p + 1) * per + q;
                float mx = -FLT_MAX, my = -FLT_MAX;
                for (int k = 0; k < S; ++k) {
                    mx = fmaxf(mx, x[k]);
                    my = fmaxf(my, y[k * S]);
                }
                sx[r] = mx;
                sy[r] = my;
            }
            __syncthreads();
            TailIdx w(tid, S, TAIL_THREADS);  // here p counts elements: even = x (row shift), odd = y (column shift)
            for (int idx = tid; idx < 2 * half * per; idx += TAIL_THREADS) {
                const float shift = (w.p & 1) ? sy[(w.p >> 1) * S + w.j] : sx[(w.p >> 1) * S + w.i];
                A[idx] = __expf(A[idx] - shift);
                w.next(S);
            }
            __syncthreads();
        }
        {
            TailIdx w(tid, S, TAIL_THREADS);
            for (int idx = tid; idx < half * per; idx += TAIL_THREADS) {
                const float* x = A + (size_t)(2 * w.p) * per + w.i * S;
                const float* y = A + (size_t)(2 * w.p + 1) * per + w.j;
                float v;
                if (SEMI == FB2_SEMIRING_LOG) {
                    float acc0 = 0.f, acc1 = 0.f;
                    int k = 0;
                    for (; k + 1 < S; k += 2) {
                        acc0 = fmaf(x[k], y[k * S], acc0);
                        acc1 = fmaf(x[k + 1], y[(k + 1) * S], acc1);
                    }
                    if (k < S) acc0 = fmaf(x[k], y[k * S], acc0);
                    v = (sx[w.p * S + w.i] + sy[w.p * S + w.j]) + __logf(acc0 + acc1);
                } else {
                    v = kNegInf;
                    for (int k = 0; k < S; ++k) v = fmaxf(v, x[k] + y[k * S]);
                }
                Bf[idx] = v;
                w.next(S);
            }
        }
        if (nd != half)
            for (int r = tid; r < per; r += TAIL_THREADS) Bf[(size_t)half * per + r] = A[(size_t)(d - 1) * per + r];
        __syncthreads();
        float* t = A;
        A = Bf;
        Bf = t;  // capacities alternate between d0 and (d0 + 1) / 2 elements; level sizes halve, so they always fit
        d = nd;
    }
    for (int r = tid; r < per; r += TAIL_THREADS) out[(int64_t)blockIdx.x * per + r] = A[r];
}

static bool chain_tail_disabled() {
    const char* e = getenv("FB2_DISABLE_CHAIN_TAIL");
    return e && e[0] == '1';
}

static void level_sizes(int64_t T, int64_t* t1, int64_t* t2) {
    *t1 = (T + 1) / 2;
    *t2 = (*t1 + 1) / 2;
}

namespace fb2 {  // chain_gemm_tc.cu
bool chain_gemm_supported(int S);
size_t chain_product_gemm_workspace_bytes(int64_t B, int64_t T, int S);
int chain_product_gemm_device(const float* trans, int64_t B, int64_t T, int S, float* out, void* workspace, size_t workspace_bytes, cudaStream_t stream);
}  // namespace fb2
static bool chain_gemm_enabled() {
    const char* e = getenv("FB2_CHAIN_GEMM");
    return !(e && e[0] == '0');
}

extern "C" size_t fb2_chain_product_gemm_workspace_bytes(int64_t B, int64_t T, int32_t S) {
    if (B < 1 || T < 4 || !chain_gemm_supported(S) || !chain_gemm_enabled()) return 0;
    return chain_product_gemm_workspace_bytes(B, T, S);
}

extern "C" size_t fb2_chain_product_workspace_bytes(int64_t B, int64_t T, int32_t S) {
    int64_t t1, t2;
    level_sizes(T, &t1, &t2);
    size_t per = (size_t)S * S * sizeof(float);
    return align_up((size_t)B * t1 * per) + align_up((size_t)B * t2 * per) +
           align_up(log_matmul_tc_workspace_floats(B * (T / 2 > 0 ? T / 2 : 1), S, S) * sizeof(float)) + 256;
}

extern "C" int fb2_chain_product_f32(int semiring, const float* trans, const int64_t ts[4], int64_t B, int64_t T,
                                     int32_t S, float* out, void* workspace, size_t workspace_bytes, void* stream_) {
    int st = check_device();
    if (st != FB2_OK) return st;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    FB2_REQUIRE(semiring == FB2_SEMIRING_LOG || semiring == FB2_SEMIRING_MAX, "unknown semiring %d", semiring);
    FB2_REQUIRE(B >= 0 && T >= 1 && S >= 1, "bad extents B=%lld T=%lld S=%d", (long long)B, (long long)T, S);
    FB2_REQUIRE(trans && out, "null pointer");
    if (B == 0) return FB2_OK;
    int64_t per = (int64_t)S * S;
    if (T == 1) return launch_copy_matrices(trans, ts[0], ts[2], ts[3], out, per, B, S, stream);
    // Large state spaces, log semiring, contiguous input, a workspace that holds the TF32 planes of every level: the whole tree runs in
    // the linear domain on the TMA + tcgen05 GEMM kernel (chain_gemm_tc.cu); exp / log are applied once per chain instead of per level.
    if (semiring == FB2_SEMIRING_LOG && T >= 4 && chain_gemm_enabled() && chain_gemm_supported(S) && ts[3] == 1 && ts[2] == S && ts[1] == per &&
        ts[0] == T * per && workspace_bytes >= chain_product_gemm_workspace_bytes(B, T, S)) {
        st = chain_product_gemm_device(trans, B, T, S, out, workspace, workspace_bytes, stream);
        if (st != FB2_ERR_UNSUPPORTED) return st;
    }
    int64_t t1, t2;
    level_sizes(T, &t1, &t2);
    Arena arena(workspace, workspace_bytes);
    float* buf[2];
    buf[0] = arena.take<float>((size_t)B * t1 * per);
    buf[1] = arena.take<float>((size_t)B * t2 * per);
    float* shift_ws = arena.take<float>(log_matmul_tc_workspace_floats(B * (T / 2 > 0 ? T / 2 : 1), S, S));
    if (!arena.ok()) {
        set_error("chain_product: workspace too small (%zu bytes given)", workspace_bytes);
        return FB2_ERR_WORKSPACE;
    }
    // current level view
    const float* cur = trans;
    int64_t cs[4] = {ts[0], ts[1], ts[2], ts[3]};
    int64_t d = T;
    int which = 0;
    while (d > 1) {
        // (d <= 128: beyond that a level still has enough independent pairs to fill the GPU and one CTA per chain is slower)
        if (S <= 32 && d <= 128 && B <= 65535 && !chain_tail_disabled() && tail_smem_bytes(d, S) <= TAIL_SMEM_LIMIT) {
            const size_t smem = tail_smem_bytes(d, S);
            if (semiring == FB2_SEMIRING_LOG) {
                FB2_CUDA(cudaFuncSetAttribute(chain_product_tail<FB2_SEMIRING_LOG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                chain_product_tail<FB2_SEMIRING_LOG><<<(unsigned)B, TAIL_THREADS, smem, stream>>>(cur, cs[0], cs[1], cs[2], cs[3], (int)d, S, out);
            } else {
                FB2_CUDA(cudaFuncSetAttribute(chain_product_tail<FB2_SEMIRING_MAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                chain_product_tail<FB2_SEMIRING_MAX><<<(unsigned)B, TAIL_THREADS, smem, stream>>>(cur, cs[0], cs[1], cs[2], cs[3], (int)d, S, out);
            }
            FB2_LAUNCH_CHECK();
            return FB2_OK;
        }
        int64_t half = d / 2, nd = (d + 1) / 2;
        float* dst = (nd == 1) ? out : buf[which];
        int64_t dst_sb = (nd == 1) ? per : nd * per;  // batch stride of dst in elements
        // x = cur[:, 0:even:2], y = cur[:, 1:even:2]  (sum_product.py:692-694)
        int64_t xs[4] = {cs[0], 2 * cs[1], cs[2], cs[3]};
        int64_t ys[4] = {cs[0], 2 * cs[1], cs[2], cs[3]};
        // products go to dst[b, 0:half]; an odd leftover cur[b, d-1] goes to dst[b, half] (sum_product.py:696-698)
        st = launch_semiring_matmul(semiring, cur, xs, cur + cs[1], ys, dst, nullptr, B, half, S, S, S, stream, dst_sb, shift_ws);
        if (st != FB2_OK) return st;
        if (nd != half) {
            st = launch_copy_matrices(cur + (d - 1) * cs[1], cs[0], cs[2], cs[3], dst + half * per, dst_sb, B, S, stream);
            if (st != FB2_OK) return st;
        }
        cur = dst;
        cs[0] = nd * per; cs[1] = per; cs[2] = S; cs[3] = 1;
        d = nd;
        which ^= 1;
    }
    return FB2_OK;
}
