// FP64 peak of the device, measured: BASELINE.md asks for a measured DFMA peak before any kernel is called compute-bound.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu && tools/fp64_peak
// Two numbers: DFMA (fused, 2 flop per instruction — the vendor's peak) and separate DMUL + DADD (what the engine can use:
// it is compiled with -fmad=false because every product and sum must be rounded on its own, DESIGN.md "f64 contract").
#include <cstdio>
#include <cuda_runtime.h>

template <bool FUSED>
__global__ void __launch_bounds__(256) k_flops(double* out, double a, double b, int iters) {
    double x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = threadIdx.x * 1e-3 + k;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (FUSED) x[k] = __fma_rn(x[k], a, b);
            else x[k] = __dadd_rn(__dmul_rn(x[k], a), b);
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <bool FUSED>
static double run(int sms) {
    const int blocks = sms * 8, iters = 1 << 15;
    double* out;
    cudaMalloc(&out, sizeof(double) * blocks * 256);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_flops<FUSED><<<blocks, 256>>>(out, 1.0000001, 1e-9, 1024);  // warm-up
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        k_flops<FUSED><<<blocks, 256>>>(out, 1.0000001, 1e-9, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaFree(out);
    return 2.0 * 8.0 * iters * (double)blocks * 256.0 / (best * 1e-3) / 1e12;  // TFLOP/s
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const double fused = run<true>(p.multiProcessorCount), split = run<false>(p.multiProcessorCount);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"fp64_dfma_tflops\": %.2f, \"fp64_dmul_dadd_tflops\": %.2f, \"how\": \"8 independent chains per thread, 8 blocks of 256 threads per SM, best of 5, CUDA events\"}\n",
           p.name, p.multiProcessorCount, fused, split);
    return 0;
}
