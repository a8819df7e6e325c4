// One-way latency of a store -> remote-SM load hand-over through L2 on B200, for the instruction variants the
// dataflow kernel could use.  Block 0 and block `peer` bounce a counter; time per hop = elapsed / (2*iters).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pingpong pingpong.cu   (the binary is not tracked)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int ST, int LD>
__global__ void pingpong(unsigned long long* a, unsigned long long* b, int iters, int peer, long long* out) {
    if (threadIdx.x != 0) return;
    const bool first = blockIdx.x == 0;
    if (!first && (int)blockIdx.x != peer) return;
    unsigned long long* mine = first ? a : b;   // I write here
    unsigned long long* theirs = first ? b : a; // I poll here
    long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) {
        if (first) {
            if (ST == 0) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(mine), "l"((unsigned long long)i) : "memory");
            if (ST == 1) atomicExch(mine, (unsigned long long)i);
            if (ST == 2) asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(mine), "l"((unsigned long long)i) : "memory");
            if (ST == 3) *(volatile unsigned long long*)mine = i;
            if (ST == 4) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(mine), "l"((unsigned long long)i) : "memory"); __threadfence(); }
        }
        unsigned long long v;
        do {
            if (LD == 0) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(theirs) : "memory");
            if (LD == 1) v = atomicAdd(theirs, 0ull);
            if (LD == 2) asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(theirs) : "memory");
            if (LD == 3) v = *(volatile unsigned long long*)theirs;
        } while (v < (unsigned long long)i);
        if (!first) {
            if (ST == 0) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(mine), "l"((unsigned long long)i) : "memory");
            if (ST == 1) atomicExch(mine, (unsigned long long)i);
            if (ST == 2) asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(mine), "l"((unsigned long long)i) : "memory");
            if (ST == 3) *(volatile unsigned long long*)mine = i;
            if (ST == 4) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(mine), "l"((unsigned long long)i) : "memory"); __threadfence(); }
        }
    }
    if (first) *out = clock64() - t0;
}

template <int ST, int LD>
void run(const char* name, unsigned long long* buf, long long* out, int peer) {
    const int iters = 20000;
    cudaMemset(buf, 0, 4096);
    pingpong<ST, LD><<<148, 32>>>(buf, buf + 256, iters, peer, out);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("%-44s peer %3d : %7.1f cycles per hop (%s)\n", name, peer, (double)h / (2.0 * iters), cudaGetErrorString(e));
}

int main() {
    unsigned long long* buf;
    long long* out;
    cudaMalloc(&buf, 4096);
    cudaMalloc(&out, 8);
    for (int peer : {1, 2, 37, 74, 100, 147}) {
        run<0, 0>("st.relaxed.gpu   -> ld.relaxed.gpu", buf, out, peer);
    }
    const int peer = 100;
    run<1, 0>("atom.exch        -> ld.relaxed.gpu", buf, out, peer);
    run<2, 0>("st.release.gpu   -> ld.relaxed.gpu", buf, out, peer);
    run<3, 3>("st.volatile      -> ld.volatile", buf, out, peer);
    run<4, 0>("st.relaxed+fence -> ld.relaxed.gpu", buf, out, peer);
    run<0, 1>("st.relaxed.gpu   -> atom.add 0", buf, out, peer);
    run<1, 1>("atom.exch        -> atom.add 0", buf, out, peer);
    run<0, 2>("st.relaxed.gpu   -> ld.acquire.gpu", buf, out, peer);
    return 0;
}
