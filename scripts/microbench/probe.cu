// Probes behind the design choices of the dataflow kernels (nvcc -gencode arch=compute_100a,code=sm_100a -o probe probe.cu):
//  (1) how many clusters of 2/4/8/16 CTAs (256 threads, one CTA per SM) are co-resident on this GPU;
//  (2) issue rate of mma.sync.m16n8k8 tf32 per SM with 8 and 16 warps (cycles per instruction per SM sub-partition).
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256, 1) dummy_cluster(float* out) {
    extern __shared__ float sm[];
    if (out) out[blockIdx.x] = sm[threadIdx.x];
}

template <int NACC>
__global__ void mma_rate(float* out, int iters, long long* cycles) {
    float d[NACC][4];
    for (int q = 0; q < NACC; ++q)
        for (int c = 0; c < 4; ++c) d[q][c] = 0.f;
    unsigned a[4] = {threadIdx.x, threadIdx.x + 1, threadIdx.x + 2, threadIdx.x + 3}, b0 = threadIdx.x * 3, b1 = threadIdx.x * 5;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int q = 0; q < NACC; ++q)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[q][0]), "+f"(d[q][1]), "+f"(d[q][2]), "+f"(d[q][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
    const long long t1 = clock64();
    float s = 0.f;
    for (int q = 0; q < NACC; ++q)
        for (int c = 0; c < 4; ++c) s += d[q][c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
    cudaFuncSetAttribute(dummy_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaFuncSetAttribute(dummy_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int cs : {1, 2, 4, 8, 16}) {
        for (int smem : {0, 100 * 1024}) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(cs * 32);
            cfg.blockDim = dim3(256);
            cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = cs;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            int n = -1;
            cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy_cluster, &cfg);
            printf("cluster size %2d smem %6d: max active clusters %d (%d CTAs) %s\n", cs, smem, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    }
    float* out;
    long long* cyc;
    cudaMalloc(&out, 1 << 20);
    cudaMalloc(&cyc, 1024 * 8);
    for (int warps : {4, 8, 16}) {
        const int iters = 2000;
        mma_rate<6><<<148, warps * 32>>>(out, iters, cyc);
        mma_rate<6><<<148, warps * 32>>>(out, iters, cyc);
        cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        const double per_smsp = (double)h[0] / (iters * 6.0 * warps / 4.0);
        printf("mma.sync m16n8k8 tf32: %2d warps/SM, 6 independent accumulators: %.2f cycles per instruction per sub-partition (%lld cycles)\n", warps, per_smsp, h[0]);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
