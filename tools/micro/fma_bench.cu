// fma_bench.cu -- fp32 FMA issue-rate probe for sm_100a: scalar FFMA vs packed FFMA2 vs a mix (which pipes run what).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/fma_bench tools/micro/fma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ void ffma2(u64& d, u64 a, u64 b) { asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b)); }
__device__ __forceinline__ void ffma1(float& d, float a, float b) { asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(d) : "f"(a), "f"(b)); }

template <int MODE>   // 0: 32 scalar FFMA / iter   1: 16 FFMA2 / iter   2: 16 FFMA2 + 16 FFMA interleaved / iter  3: 16 FFMA2 + 8 FFMA
__global__ void __launch_bounds__(1024, 1) probe(float* out, int iters, float a0, float b0) {
    float acc[32];
    u64 acc2[16];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = (float)i;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc2[i] = (u64)i;
    float a = a0 + threadIdx.x * 1e-9f, b = b0, c = a0 * 0.5f, d = b0 * 0.25f;
    u64 pa, pb;
    asm("mov.b64 %0, {%1, %2};" : "=l"(pa) : "f"(a), "f"(c));
    asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b), "f"(d));
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) ffma1(acc[i], (i & 1) ? a : c, (i & 2) ? b : d);
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) ffma2(acc2[i], pa, pb);
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) { ffma2(acc2[i], pa, pb); ffma1(acc[i], a, b); }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) { ffma2(acc2[i], pa, pb); if (i & 1) ffma1(acc[i], a, b); }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += acc[i];
#pragma unroll
    for (int i = 0; i < 16; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc2[i])); s += lo + hi; }
    if (s == 123.456f) out[0] = s;
}

template <int MODE>
void run(const char* name, int threads, double fma_per_thread_iter) {
    float* out; cudaMalloc(&out, 4);
    int dev = 0, sms = 0, khz = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<sms, threads>>>(out, iters, 1.0001f, 0.9999f);
    cudaEventRecord(e0);
    probe<MODE><<<sms, threads>>>(out, iters, 1.0001f, 0.9999f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma = fma_per_thread_iter * iters * (double)threads * sms;
    printf("%-28s threads/SM %4d : %7.3f ms  %7.2f TFLOP/s  %6.1f FMA/clk/SM (at %.3f GHz nominal)\n", name, threads, ms, 2 * fma / ms * 1e-9,
           fma / (ms * 1e-3) / sms / (khz * 1e3), khz * 1e-6);
    cudaFree(out);
}
int main() {
    for (int th : {128, 256, 512, 1024}) {
        run<0>("scalar FFMA x32", th, 32);
        run<1>("FFMA2 x16", th, 32);
        run<2>("FFMA2 x16 + FFMA x16", th, 48);
        run<3>("FFMA2 x16 + FFMA x8", th, 40);
    }
    return 0;
}
