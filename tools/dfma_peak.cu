// DFMA peak of the device (the FP64 roof SURVEY.md §8(d) asks for; MEASURED_PEAKS.json holds no FP64 figure):
// every thread runs 8 independent fused multiply-add chains, 2 flops per DFMA, timed with CUDA events.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/dfma_peak tools/dfma_peak.cu && /tmp/dfma_peak
#include <cuda_runtime.h>
#include <cstdio>

__global__ void __launch_bounds__(256) dfma_kernel(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 4096;
    double *out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    dfma_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
    cudaDeviceSynchronize();
    double best = 0, sum = 0;
    const int reps = 10;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0);
        dfma_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double tf = 2.0 * 64.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
        sum += tf;
    }
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"what\": \"DFMA peak, 8 independent chains per thread, %d x %d threads, 2 flops per DFMA\", \"device\": \"%s\", \"sms\": %d, "
           "\"fp64_tflops_best\": %.3f, \"fp64_tflops_mean\": %.3f, \"dfma_per_sm_per_clk_at_max_clock\": %.2f, \"max_clock_mhz\": %d}\n",
           blocks, threads, p.name, p.multiProcessorCount, best, sum / reps,
           best * 1e12 / 2.0 / p.multiProcessorCount / (clk * 1e3), clk / 1000);
    return 0;
}
