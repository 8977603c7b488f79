// umma_bench.cu -- cycles per tcgen05.mma (kind::tf32, M=128, K=8) as a function of N, for A from shared memory (SS)
// and A from tensor memory (TS): the cost model behind the skinny-product kernels.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/umma_bench tools/micro/umma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
template <bool TS>
__global__ void __launch_bounds__(128, 1) probe(int N, int iters, int nacc, long long* out, int ncommit = 0) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ __align__(8) uint64_t dummy[2];
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<float*>(smem + (base - smem_u32(smem)))[i] = 1.0f;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&dummy[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&dummy[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t da = desc_kmajor(base), db = desc_kmajor(base + 16384);
        const uint32_t a_tmem = tmem + 448;          // columns 448..511: never written, contents irrelevant
        const long long t0 = clock64();
        const uint32_t d0 = tmem, d1 = tmem + (nacc > 1 ? N : 0), d2 = tmem + (nacc > 2 ? 2 * N : 0);
#define ISSUE(D, KK)                                                                                                        \
    if (TS) {                                                                                                               \
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" \
                     ::"r"(D), "r"(a_tmem + 8u * KK), "l"(db + (uint64_t)(2 * KK)), "r"(idesc), "r"(1u) : "memory");      \
    } else {                                                                                                                \
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"   \
                     ::"r"(D), "l"(da + (uint64_t)(2 * KK)), "l"(db + (uint64_t)(2 * KK)), "r"(idesc), "r"(1u) : "memory"); \
    }
        for (int i = 0; i < iters; i += 12) {
            ISSUE(d0, 0) ISSUE(d1, 1) ISSUE(d2, 2) ISSUE(d0, 3) ISSUE(d1, 0) ISSUE(d2, 1)
            ISSUE(d0, 2) ISSUE(d1, 3) ISSUE(d2, 0) ISSUE(d0, 1) ISSUE(d1, 2) ISSUE(d2, 3)
            for (int c = 0; c < ncommit; ++c)      // the real kernels commit to 1-2 mbarriers after every 8-12 MMAs
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&dummy[c])) : "memory");
        }
        const long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        } while (!ok);
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// the k-block body of the N-stacked TS kernel: [fence] + 4 x (N=BN lo*raw, N=2BN hi*[raw|lo]) + ncommit commits
__global__ void __launch_bounds__(128, 1) kblock_probe(int BN, int kblocks, int ncommit, int fence, long long* out, int st_traffic = 0) {
    __shared__ volatile int stop_flag;
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ __align__(8) uint64_t dummy[2];
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&dummy[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&dummy[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) stop_flag = 0;
    __syncthreads();
    if (st_traffic && threadIdx.x >= 32) {          // warps 1-3: tcgen05.st into columns 256..383 of their lane quadrants, like the split warps
        const uint32_t q = threadIdx.x >> 5;
        uint32_t v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = threadIdx.x + i;
        int it = 0;
        while (!stop_flag) {
            const uint32_t ta = tmem + ((q * 32u) << 16) + 256u + (uint32_t)((it & 7) * 16);
            asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(ta), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
            if ((++it & (st_traffic - 1)) == 0) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        const uint32_t ib = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 4) << 24);
        const uint32_t id1 = ib | ((uint32_t)(BN >> 3) << 17), id2 = ib | ((uint32_t)((2 * BN) >> 3) << 17);
        const uint64_t db = desc_kmajor(base + 16384);
        const uint32_t a_hi = tmem + 448, a_lo = tmem + 480;
        const long long t0 = clock64();
        for (int kb = 0; kb < kblocks; ++kb) {
            if (fence) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                             ::"r"(tmem + 6 * BN), "r"(a_lo + 8u * kk), "l"(db + (uint64_t)(2 * kk)), "r"(id1), "r"(1u) : "memory");
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                             ::"r"(tmem + (uint32_t)(kk % 3) * 2 * BN), "r"(a_hi + 8u * kk), "l"(db + (uint64_t)(2 * kk)), "r"(id2), "r"(1u) : "memory");
            }
            for (int c = 0; c < ncommit; ++c)
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&dummy[c])) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        } while (!ok);
        if (blockIdx.x == 0) out[1] = clock64() - t0;
        stop_flag = 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}
int main() {
    long long* out; cudaMallocManaged(&out, 16);
    {
        const int sm2 = 16384 + 32768 + 1024;
        cudaFuncSetAttribute(kblock_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, sm2);
        for (int stt : {2, 16}) {
            kblock_probe<<<148, 128, sm2>>>(32, 400, 1, 1, out, stt); cudaDeviceSynchronize();
            printf("TS k-block body, 1 commit, with concurrent tcgen05.st traffic from 3 warps (wait every %2d): %7.1f clk per k-block  %s\n", stt, (double)out[1] / 400,
                   cudaGetErrorString(cudaGetLastError()));
        }
        for (int nc = 0; nc <= 2; ++nc)
            for (int fe = 0; fe <= 1; ++fe) {
                kblock_probe<<<148, 128, sm2>>>(32, 400, nc, fe, out); cudaDeviceSynchronize();
                printf("TS k-block body (BN=32: 4 x [N32 + N64]), %d commit(s), fence %d: %7.1f clk per k-block  %s\n", nc, fe, (double)out[1] / 400,
                       cudaGetErrorString(cudaGetLastError()));
            }
    }
    const int smem = 16384 + 32768 + 1024, iters = 4800;
    cudaFuncSetAttribute(probe<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(probe<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int ctas : {1, 148}) {
        for (int N : {32, 64, 128, 256}) {
            for (int nacc : {1, 3}) {
                if (nacc * N > 448) continue;
                probe<false><<<ctas, 128, smem>>>(N, iters, nacc, out); cudaDeviceSynchronize();
                const double ss = (double)out[1] / iters, ssi = (double)out[0] / iters;
                probe<true><<<ctas, 128, smem>>>(N, iters, nacc, out); cudaDeviceSynchronize();
                const double ts = (double)out[1] / iters, tsi = (double)out[0] / iters;
                if (N == 32 && nacc == 3) {
                    for (int nc = 1; nc <= 2; ++nc) {
                        probe<true><<<ctas, 128, smem>>>(N, iters, nacc, out, nc); cudaDeviceSynchronize();
                        printf("CTAs %3d  N 32 TS, %d commit(s) per 12 MMAs: %6.1f clk/mma\n", ctas, nc, (double)out[1] / iters);
                    }
                }
                printf("CTAs %3d  M128 N%3d K8 tf32, %d accumulator(s):  SS %6.1f clk/mma (issue %5.1f)   TS %6.1f clk/mma (issue %5.1f)   %s\n", ctas, N, nacc, ss, ssi, ts, tsi,
                       cudaGetErrorString(cudaGetLastError()));
            }
        }
    }
    return 0;
}
