// read_bench.cu -- achievable HBM READ bandwidth on this GPU for (a) a flat stream, (b) the attention access pattern
// (rows of 8 KB, one per (t,b), consumed by one warp each in 512-byte pieces, eight loads in flight per lane).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/read_bench tools/micro/read_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
template <int UNROLL>
__global__ void flat_read(const float* __restrict__ x, long long n4, float* out) {
    float s = 0.f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < n4; i += UNROLL * stride) {
        float4 v[UNROLL];
#pragma unroll
        for (int k = 0; k < UNROLL; ++k) v[k] = ldg_stream4(x + 4 * (i + k * stride));
#pragma unroll
        for (int k = 0; k < UNROLL; ++k) s += v[k].x + v[k].y + v[k].z + v[k].w;
    }
    if (s == 123.456f) out[0] = s;
}
// rows of `rowf` floats; warp w of the grid takes rows w, w+W, ...
__global__ void row_read(const float* __restrict__ x, int rows, int rowf, float* out) {
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long W = ((long long)gridDim.x * blockDim.x) >> 5;
    float s = 0.f;
    const int n4 = rowf / 4;
    for (long long r = warp; r < rows; r += W) {
        const float* row = x + r * rowf;
        for (int i0 = lane; i0 < n4; i0 += 256) {
            float4 v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) if (i0 + 32 * k < n4) v[k] = ldg_stream4(row + 4 * (i0 + 32 * k));
#pragma unroll
            for (int k = 0; k < 8; ++k) if (i0 + 32 * k < n4) s += v[k].x + v[k].y + v[k].z + v[k].w;
        }
    }
    if (s == 123.456f) out[0] = s;
}
int main() {
    const long long bytes = 1024LL << 20;       // 1 GiB >> L2
    float *x, *out; cudaMalloc(&x, bytes); cudaMalloc(&out, 4); cudaMemset(x, 0, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    auto timeit = [&](const char* name, auto launch, double b) {
        launch(); cudaEventRecord(e0); for (int i = 0; i < 5; ++i) launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
        printf("%-46s %8.3f ms  %7.1f GB/s\n", name, ms, b / ms * 1e-6);
    };
    for (int ctas : {2, 4, 8}) {
        char nm[96];
        snprintf(nm, 96, "flat float4 x8 in flight, %d CTAs/SM x 256", ctas);
        timeit(nm, [&] { flat_read<8><<<sms * ctas, 256>>>(x, bytes / 16, out); }, (double)bytes);
        snprintf(nm, 96, "flat float4 x4 in flight, %d CTAs/SM x 256", ctas);
        timeit(nm, [&] { flat_read<4><<<sms * ctas, 256>>>(x, bytes / 16, out); }, (double)bytes);
    }
    // attention pattern: 102.4 MB = 12800 rows of 8000 B, a different 102 MB window each launch to defeat L2
    const int rows = 12800, rowf = 2000;
    int rep = 0;
    for (int ctas : {4, 6, 8}) {
        char nm[96];
        snprintf(nm, 96, "12800 rows x 8000 B, warp per row, %d CTAs/SM", ctas);
        timeit(nm, [&] { row_read<<<sms * ctas, 256>>>(x + (long long)(rep++ % 9) * rows * rowf, rows, rowf, out); }, (double)rows * rowf * 4);
    }
    return 0;
}
