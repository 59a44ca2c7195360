// Micro-benchmark: FP64 DFMA dependent-issue latency and per-SM throughput on the current GPU.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a fp64_lat.cu -o fp64_lat
#include <cstdio>
#include <cuda_runtime.h>
template <int CHAINS>
__global__ void chain_kernel(double* out, long long* cyc, int iters, double a, double b) {
    double x[CHAINS];
    for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x * 1e-3 + c;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) x[c] = fma(x[c], a, b);
    }
    long long t1 = clock64();
    double s = 0; for (int c = 0; c < CHAINS; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int CHAINS> void run(int warps_per_sm, int nsm) {
    double* out; long long* cyc;
    int threads = 32 * warps_per_sm;
    cudaMalloc(&out, sizeof(double) * threads * nsm); cudaMalloc(&cyc, sizeof(long long) * nsm);
    int iters = 2000;
    chain_kernel<CHAINS><<<nsm, threads>>>(out, cyc, iters, 0.999, 1e-3);
    chain_kernel<CHAINS><<<nsm, threads>>>(out, cyc, iters, 0.999, 1e-3);
    long long h; cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double per = (double)h / (iters * 8.0);
    printf("chains %d warps/SM %2d: %.2f clk per chain step (=> %.2f clk / DFMA / warp, %.2f DFMA-warp-instr/clk/SM)\n", CHAINS, warps_per_sm,
           per, per / CHAINS, warps_per_sm * CHAINS / per);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    run<1>(1, nsm); run<2>(1, nsm); run<4>(1, nsm); run<8>(1, nsm);
    run<1>(4, nsm); run<2>(4, nsm); run<4>(4, nsm);
    run<1>(8, nsm); run<2>(8, nsm); run<4>(8, nsm); run<8>(8, nsm);
    run<4>(16, nsm); run<4>(32, nsm);
    return 0;
}
