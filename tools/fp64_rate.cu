// Micro-benchmark: fp64 FMA issue rate per SM on the GPU it runs on (independent chains, varying warps per CTA).
// Used to decide what bounds the Schur / Cholesky kernels.
#include <cstdio>
#include <cuda_runtime.h>
template <int kChains>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double x[kChains];
#pragma unroll
    for (int c = 0; c < kChains; ++c) x[c] = threadIdx.x + c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < kChains; ++c) x[c] = fma(x[c], a, b);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < kChains; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void ffma_kernel(float* out, int iters, float a, float b) {
    float x[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) x[c] = threadIdx.x + c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < 8; ++c) x[c] = fmaf(x[c], a, b);
    }
    float s = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double* out; cudaMalloc(&out, sizeof(double) * sms * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    printf("%s, %d SMs, %d kHz nominal\n", p.name, sms, khz);
    for (int warps : {1, 2, 4, 8, 16, 32}) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            dfma_kernel<8><<<sms, warps * 32>>>(out, iters, 1.0000001, 1e-9);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fma_per_sm = double(iters) * 8 * warps * 32;
        printf("dfma chains=8 warps/SM=%2d: %.3f ms  %.1f FMA lanes/SM/clk(@%.0f MHz)  %.2f TFLOP/s\n", warps, ms,
               fma_per_sm / (ms * 1e-3) / (khz * 1e3), khz / 1e3, 2 * fma_per_sm * sms / (ms * 1e-3) / 1e12);
    }
    for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0); dfma_kernel<1><<<sms, 32>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1); }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("dfma dependent chain, 1 warp: %.1f cycles per FMA (latency)\n", ms * 1e-3 * khz * 1e3 / iters);
    for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0); ffma_kernel<<<sms, 1024>>>((float*)out, iters, 1.0000001f, 1e-9f); cudaEventRecord(e1); cudaEventSynchronize(e1); }
    cudaEventElapsedTime(&ms, e0, e1);
    printf("ffma 32 warps: %.1f FMA lanes/SM/clk\n", double(iters) * 8 * 1024 / (ms * 1e-3) / (khz * 1e3));
    return 0;
}
