// dfma_bench.cu — FP64 pipe microbenchmark for B200 (sm_100a): latency of a dependent DFMA chain, and DFMA throughput per SM as a
// function of the independent chains per warp (ILP) and the warps per SM.  BASELINE.md §2 asks for these numbers before any FP64-pipe
// fraction is quoted.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_bench scripts/dfma_bench.cu ; run: ./dfma_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP> __global__ void dfma_kernel(double *out, long long *cyc, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-9 + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 16; r++)
#pragma unroll
            for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int ILP> void run(int warps_per_sm, int sms, double clock_ghz) {
    const int iters = 2048;
    double *out; long long *cyc;
    cudaMalloc(&out, sizeof(double) * sms * warps_per_sm * 32);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    dfma_kernel<ILP><<<sms, warps_per_sm * 32>>>(out, cyc, iters, 0.999999, 1e-7);
    cudaEventRecord(e0);
    dfma_kernel<ILP><<<sms, warps_per_sm * 32>>>(out, cyc, iters, 0.999999, 1e-7);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[1024]; cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < sms; i++) mean += h[i]; mean /= sms;
    const double n_dep = (double)iters * 16;                     // DFMAs per chain
    const double total = n_dep * ILP * 32.0 * warps_per_sm * sms; // thread-level DFMAs
    printf("ILP %2d warps/SM %2d : %.2f cycles per dependent DFMA step, %.1f DFMA/clk/SM, %.2f TFLOP/s (events: %.3f ms)\n", ILP, warps_per_sm,
           mean / n_dep, n_dep * ILP * 32.0 * warps_per_sm / mean, 2.0 * total / (ms * 1e-3) / 1e12, ms);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("%s, %d SMs, clock %.0f MHz\n", p.name, sms, p.clockRate / 1e3);
    run<1>(1, sms, 0); run<2>(1, sms, 0); run<4>(1, sms, 0); run<8>(1, sms, 0);
    run<1>(4, sms, 0); run<2>(4, sms, 0); run<4>(4, sms, 0); run<8>(4, sms, 0);
    run<1>(8, sms, 0); run<4>(8, sms, 0); run<1>(16, sms, 0); run<4>(16, sms, 0); run<8>(16, sms, 0); run<1>(32, sms, 0); run<4>(32, sms, 0); run<4>(64, sms, 0);
    return 0;
}
