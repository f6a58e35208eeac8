// fp64_peak.cu -- measures the FP64 CUDA-core peaks used as roofline denominators:
// (a) DFMA, (b) un-fused DADD+DMUL (what the parity-constrained kernels can issue),
// (c) DADD only.  Prints one JSON line.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <stdio.h>

template <int MODE>
__global__ void __launch_bounds__(256) k_peak(double *out, double x, double y, int iters) {
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = threadIdx.x * 1e-3 + k;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (MODE == 0) a[k] = __fma_rn(a[k], x, y);
            if (MODE == 1) a[k] = __dadd_rn(__dmul_rn(a[k], x), y);
            if (MODE == 2) a[k] = __dadd_rn(a[k], y);
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static double run(int sms, int iters) {
    const int grid = sms * 8, threads = 256;
    double *out;
    cudaMalloc(&out, sizeof(double) * grid * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_peak<MODE><<<grid, threads>>>(out, 1.0000001, 1e-9, iters);  // warm-up
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        k_peak<MODE><<<grid, threads>>>(out, 1.0000001, 1e-9, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaFree(out);
    const double ops_per_iter = MODE == 2 ? 8.0 : 16.0;  // flops per thread-iteration
    return ops_per_iter * iters * (double)grid * threads / (best * 1e-3) / 1e12;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    if (sms == 0) {
        printf("{\"error\": \"no device\"}\n");
        return 1;
    }
    const int iters = 1 << 15;
    double fma = run<0>(sms, iters), muladd = run<1>(sms, iters), add = run<2>(sms, iters);
    printf("{\"sm_count\": %d, \"dfma_tflops\": %.3f, \"dmul_dadd_unfused_tflops\": %.3f, "
           "\"dadd_only_tflops\": %.3f}\n", sms, fma, muladd, add);
    return 0;
}
