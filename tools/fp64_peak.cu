// FP64 peak microbenchmarks for the roofline denominators of the FP64 contractions (Hankel, Legendre, Procrustes GEMMs):
//   dfma : register-only chains of fma.rn.f64        (FP64 pipe)
//   dmma : register-only chains of mma.sync.m8n8k4.f64 (DMMA, the only FP64 tensor shape of sm_100a)
// Build:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/_bin/fp64_peak tools/fp64_peak.cu
// Run (GPU box): tools/_bin/fp64_peak > gpurun_out/fp64_peak.json      (one JSON line per kernel, best of 5 timed launches)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

template <int CHAINS>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
    double acc[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c] = threadIdx.x * 1e-3 + c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) acc[c] = fma(acc[c], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += acc[c];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;      // never true: keeps the chains alive
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int CHAINS>
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters, double a, double b) {
    double d0[CHAINS], d1[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { d0[c] = threadIdx.x * 1e-3 + c; d1[c] = c; }
    const double av = a + threadIdx.x * 1e-9, bv = b;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) dmma884(d0[c], d1[c], av, bv);
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += d0[c] + d1[c];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double best_ms(F launch) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; ++w) launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    double* out;
    CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 256));
    const int iters = 1 << 14;
    for (int ctas_per_sm : {2, 4, 8}) {
        const int grid = sms * ctas_per_sm;
        {
            constexpr int CH = 8;
            const double ms = best_ms([&] { dfma_kernel<CH><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); });
            const double flops = 2.0 * CH * (double)iters * 256.0 * grid;
            printf("{\"kernel\": \"dfma\", \"chains\": %d, \"ctas_per_sm\": %d, \"sms\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", CH, ctas_per_sm, sms, ms, flops / ms * 1e-9);
        }
        {
            constexpr int CH = 8;
            const double ms = best_ms([&] { dmma_kernel<CH><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); });
            const double flops = 2.0 * 8 * 8 * 4 * CH * (double)iters * 8.0 * grid;      // 8 warps per CTA, 512 flops per DMMA
            printf("{\"kernel\": \"dmma_m8n8k4\", \"chains\": %d, \"ctas_per_sm\": %d, \"sms\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", CH, ctas_per_sm, sms, ms, flops / ms * 1e-9);
        }
    }
    CK(cudaGetLastError());
    CK(cudaFree(out));
    return 0;
}
