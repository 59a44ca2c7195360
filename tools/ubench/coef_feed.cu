// Micro-benchmark: how fast can warp-uniform FP64 coefficients be fed to DFMAs -- broadcast LDS.128 from shared memory (what
// geometry_kernel does) against the constant bank (uniform data path)?  Access pattern of K1's 3-D mode sums: run-time loop
// over m, unrolled loop over k and the rows; every loaded pair (E, O) feeds two FMAs with per-thread cos / sin(k phi).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a coef_feed.cu -o coef_feed
#include <cstdio>
#include <cuda_runtime.h>
constexpr int M = 12, NT = 6, ROWS = 9, ROWP = 2 * ROWS;       // NCSX-like: 12 x 7 x 18 doubles = 12 KB
__constant__ double c_tab[M * (NT + 1) * ROWP];

template <int SRC>      // 0: shared memory, 1: constant bank
__global__ void __launch_bounds__(128) feed_kernel(const double* __restrict__ g_tab, double* __restrict__ out, int reps, int mcount) {
    extern __shared__ __align__(16) double s_tab[];
    for (int i = threadIdx.x; i < M * (NT + 1) * ROWP; i += blockDim.x) s_tab[i] = g_tab[i];
    __syncthreads();
    double cn[NT + 1], sn[NT + 1];
    for (int k = 0; k <= NT; ++k) { cn[k] = 1.0 / (1 + k + threadIdx.x); sn[k] = 0.5 / (2 + k + threadIdx.x); }
    double A[ROWS], B[ROWS];
    for (int j = 0; j < ROWS; ++j) { A[j] = 0.0; B[j] = 0.0; }
    for (int r = 0; r < reps; ++r)
        for (int m = 0; m < mcount; ++m) {
            const double* row = (SRC == 0 ? s_tab : c_tab) + m * (NT + 1) * ROWP;
#pragma unroll
            for (int k = 1; k <= NT; ++k)
#pragma unroll
                for (int j = 0; j < ROWS; ++j) {
                    const double2 eo = *reinterpret_cast<const double2*>(row + k * ROWP + 2 * j);
                    A[j] = fma(eo.x, cn[k], A[j]);
                    B[j] = fma(eo.y, sn[k], B[j]);
                }
        }
    double s = 0.0;
    for (int j = 0; j < ROWS; ++j) s += A[j] + B[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int SRC> void run(const double* d_tab, double* d_out, int nsm, int ctas_per_sm) {
    const int reps = 400;
    const size_t smem = sizeof(double) * M * (NT + 1) * ROWP;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    feed_kernel<SRC><<<nsm * ctas_per_sm, 128, smem>>>(d_tab, d_out, 10, M);
    cudaEventRecord(e0);
    feed_kernel<SRC><<<nsm * ctas_per_sm, 128, smem>>>(d_tab, d_out, reps, M);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fma = (double)nsm * ctas_per_sm * 128 * reps * M * NT * ROWS * 2;
    printf("%s, %d CTAs/SM (%d warps/SM): %.3f ms, %.2f TFLOP/s\n", SRC == 0 ? "shared (LDS.128 broadcast)" : "constant bank", ctas_per_sm,
           4 * ctas_per_sm, ms, 2.0 * fma / (ms * 1e-3) / 1e12);
}
int main() {
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    const int n = M * (NT + 1) * ROWP;
    double h[n]; for (int i = 0; i < n; ++i) h[i] = 1e-3 * (i % 97);
    double *d_tab, *d_out; cudaMalloc(&d_tab, sizeof(h)); cudaMalloc(&d_out, sizeof(double) * nsm * 8 * 128);
    cudaMemcpy(d_tab, h, sizeof(h), cudaMemcpyHostToDevice); cudaMemcpyToSymbol(c_tab, h, sizeof(h));
    for (int c : {2, 3, 4}) { run<0>(d_tab, d_out, nsm, c); run<1>(d_tab, d_out, nsm, c); }
    return 0;
}
