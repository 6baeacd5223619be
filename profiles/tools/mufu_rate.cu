// Microbenchmark: MUFU (ex2 / rcp / lg2) and FFMA issue rates per SM on this GPU.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mufu_rate mufu_rate.cu && ./mufu_rate
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(float* out, int iters) {
    float a[8];
    for (int j = 0; j < 8; ++j) a[j] = 1.0f + threadIdx.x * 1e-3f + j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[j]));
            if (OP == 1) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[j]));
            if (OP == 2) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a[j]));
            if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[j]));
        }
    }
    float s = 0;
    for (int j = 0; j < 8; ++j) s += a[j];
    if (s == 123.456f) out[0] = s;
}
template <int OP>
void run(const char* name) {
    float* d; cudaMalloc(&d, 4);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int iters = 20000;
    k<OP><<<sms, 1024>>>(d, 10);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<sms, 1024>>>(d, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)1024 * 8 * iters;          // per SM
    printf("%s: %.3f ms  -> %.2f lane-ops/clk/SM at max clock %d MHz\n", name, ms, ops / (ms * 1e-3 * clk * 1e3), clk / 1000);
}
int main() { run<0>("ex2"); run<1>("rcp"); run<2>("lg2"); run<3>("ffma"); return 0; }
