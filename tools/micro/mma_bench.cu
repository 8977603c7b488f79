// mma_bench.cu -- issue rate of the warp-level (legacy) tensor-core path on sm_100a: mma.sync.m16n8k8 tf32 and m16n8k4? bf16 m16n8k16.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/mma_bench tools/micro/mma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int MODE, int NACC>
__global__ void __launch_bounds__(1024, 1) probe(float* out, int iters, unsigned seed) {
    float c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f; }
    unsigned a[4] = {seed, seed + 1, seed + 2, seed + 3}, b[2] = {seed + 4, seed + 5};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) { if (MODE == 0) mma_tf32(c[i], a, b); else mma_bf16(c[i], a, b); }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456f) out[0] = s;
}
template <int MODE, int NACC>
void run(const char* name, int threads, double mac_per_mma) {
    float* out; cudaMalloc(&out, 4);
    int sms = 0, khz = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0); cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE, NACC><<<sms, threads>>>(out, iters, 0u);
    cudaEventRecord(e0);
    probe<MODE, NACC><<<sms, threads>>>(out, iters, 0u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double mma = (double)NACC * iters * (threads / 32) * sms;
    double cyc = ms * 1e-3 * khz * 1e3;
    printf("%-22s acc %2d warps/SM %2d : %7.3f ms  %8.1f TFLOP/s  %7.1f MAC/clk/SM  %5.2f clk per mma per SMSP\n", name, NACC, threads / 32, ms,
           2 * mma * mac_per_mma / ms * 1e-9, mma * mac_per_mma / cyc / sms, cyc / (mma / sms / 4));
    cudaFree(out);
}
int main() {
    for (int th : {128, 256, 512}) {
        run<0, 4>("mma.sync tf32 m16n8k8", th, 1024);
        run<0, 12>("mma.sync tf32 m16n8k8", th, 1024);
        run<1, 12>("mma.sync bf16 m16n8k16", th, 2048);
    }
    return 0;
}
