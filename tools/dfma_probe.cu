// dfma_probe.cu -- independent FP64 FMA-pipe probe (no dependency on libfock_b200): the denominator of the ERI roofline.
// Every thread runs 8 independent DFMA chains from registers; 148 x 8 blocks of 256 threads; the best of 5 timed launches.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/dfma_probe.cu -o build/dfma_probe && build/dfma_probe
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_chain(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 4096;
    double *out;
    cudaMalloc(&out, (size_t)blocks * threads * sizeof(double));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    dfma_chain<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
    cudaDeviceSynchronize();
    double best = 0;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        dfma_chain<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double tf = 2.0 * 8 * 16 * (double)iters * blocks * threads / (ms * 1e-3) * 1e-12;
        if (tf > best) best = tf;
    }
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"device\": \"%s\", \"sms\": %d, \"sm_clock_attr_mhz\": %.0f, \"dfma_tflops\": %.2f, \"nominal_64_fma_per_sm_clk_tflops\": %.2f}\n", p.name,
           p.multiProcessorCount, clk * 1e-3, best, 2.0 * 64 * p.multiProcessorCount * clk * 1e3 * 1e-12);
    return 0;
}
