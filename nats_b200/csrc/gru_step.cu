// gru_step.cu -- one recurrent GRU step in ONE launch: the skinny product h_{t-1}.[U|Ux] on the tensor cores (tcgen05
// 3xTF32, weights streamed by TMA as MN-major tiles), a deterministic split-K fix-up through L2, and the gate
// arithmetic of nats.py:336-356 / 505-518 (forward) or its reverse (backward) as the epilogue.  Replaces, per step,
// a split-K GEMM launch + a gate launch and the round trip of the split-K slabs.
//
// forward CTA (tile of 32 hidden units, K-slice, direction):
//   A operand (128 UMMA rows) = three 32-column blocks of [U|Ux] : reset-gate, update-gate, candidate columns of the
//   tile's units (TMA boxes at column offsets g*D + j0; the fourth row block is unused), B operand = h_{t-1} [batch, K].
//   TMEM lanes 0-31 / 32-63 / 64-95 then hold the three pre-activations of unit j0+lane for every batch column.
//   The last-arriving CTA of a tile (atomic ticket) sums the K-slices in fixed order, the three gate blocks meet in
//   shared memory and 128 epilogue threads finish h_t, the saved gates and the running masked sum for ctx_mean.
// backward CTA (tile of 128 hidden units, K-slice of 3D, direction): A = rows of [U|Ux] (K-major), B = dG_{t+1};
//   the fix-up CTA turns d h_t into dG_t / dGx_t / the elementwise carry (mirror of gru_gates_bwd_kernel).
#include "gemm.cuh"
#include "ops.cuh"
#include "tc_common.cuh"

namespace nats {

namespace {

using namespace tc;

constexpr int kThreads = 384;
constexpr int kSplitThreads = 256;
constexpr int kBlockK = 32;
constexpr uint32_t kATile = 128 * 128;          // bytes of one 128-row operand tile (raw or lo)

struct GruFwdDir {
    const float* xproj;                          // [B,3D] input projection incl. biases (row stride 3D)
    const float* h_prev; int ld_hprev;
    const float* mask;                           // [B] or NULL
    float* h_out; int ld_hout;
    float* r; float* u; float* c; float* p;      // [B,D] save slots or NULL
    float* ctxsum; int ld_ctxsum;                // or NULL
    float* slab;                                 // split-K exchange [tile][split][4][BN][32]
    int* counters;                               // [tile]
};
struct alignas(64) GruFwdParams {
    CUtensorMap mapW[2];                         // [U|Ux] as [K, 3D], MN-major boxes (32 cols x 32 k)
    CUtensorMap mapH[2];                         // h_{t-1} as [B, K], K-major box (32 k x BN rows)
    GruFwdDir d[2];
    int B, D, K, nsplit, kchunk;
};

struct GruBwdDir {
    const float* dh_a; int ld_a;                 // d cost / d h_t arriving from above (d context slice)
    const float* dh_b; int ld_b;                 // elementwise carry from step t+1 (NULL at the last step)
    const float* mean_grad; int ld_mean; const float* coef;   // ctx-mean path (or NULL)
    const float* r; const float* u; const float* c; const float* p;
    const float* h_prev; int ld_hprev;           // NULL = zeros
    const float* mask;
    float* dG; float* dGx; float* dh_elem;
    float* slab; int* counters;
};
struct alignas(64) GruBwdParams {
    CUtensorMap mapU[2];                         // [U|Ux] as [D, 3D], K-major box (32 k x 128 rows)
    CUtensorMap mapG[2];                         // dG_{t+1} as [B, 3D], K-major box (32 k x BN rows)
    GruBwdDir d[2];
    int B, D, K, nsplit, kchunk;
};

template <int BN>
__device__ __forceinline__ void load_acc(uint32_t tmem_d, int q, float (&acc)[BN]) {
#pragma unroll
    for (int c0 = 0; c0 < BN; c0 += 16) {
        uint32_t t0[16], t1[16], t2[16], t3[16];
        const uint32_t ta = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
        tmem_ld16(ta, t0);
        tmem_ld16(ta + BN, t1);
        tmem_ld16(ta + 2 * BN, t2);
        tmem_ld16(ta + 3 * BN, t3);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int t = 0; t < 16; ++t)
            acc[c0 + t] = ((__uint_as_float(t0[t]) + __uint_as_float(t1[t])) + __uint_as_float(t2[t])) + __uint_as_float(t3[t]);
    }
}

// split-K fix-up: every CTA of a tile parks its partial accumulators in L2; the last one to arrive (atomic ticket)
// re-reads all of them in split order (deterministic sum).  Returns true for that last CTA.
template <int BN>
__device__ __forceinline__ bool splitk_fixup(float (&acc)[BN], float* slab, int* counter, int tile, int nsplit, int split,
                                             int q, int lane, int t, int* s_flag) {
    if (nsplit <= 1) return true;
    float* mine = slab + ((size_t)(tile * nsplit + split) * 4 + q) * BN * 32;
#pragma unroll
    for (int col = 0; col < BN; ++col) __stcg(mine + col * 32 + lane, acc[col]);
    __threadfence();
    named_bar_sync(1, 128);
    if (t == 0) {
        const int prev = atomicAdd(counter, 1);
        const int last = (prev == nsplit - 1) ? 1 : 0;
        if (last) *counter = 0;                     // self-cleaning for the next step
        *s_flag = last;
    }
    named_bar_sync(1, 128);
    if (*s_flag == 0) return false;
    __threadfence();
#pragma unroll
    for (int col = 0; col < BN; ++col) acc[col] = 0.f;
    for (int s = 0; s < nsplit; ++s) {
        const float* sl = slab + ((size_t)(tile * nsplit + s) * 4 + q) * BN * 32;
#pragma unroll
        for (int col = 0; col < BN; ++col) acc[col] += __ldcg(sl + col * 32 + lane);
    }
    return true;
}

// ------------------------------------------------------------------------------------------------ forward
template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreads, 1) gru_fwd_step_kernel(const __grid_constant__ GruFwdParams P) {
    constexpr uint32_t kBBytes = BN * 128;
    constexpr uint32_t kStageBytes = 2 * kATile + 2 * kBBytes;
    constexpr uint32_t kTmemCols = 4 * BN;

    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t tma_full[STAGES];
    __shared__ __align__(8) uint64_t mma_full[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ __align__(8) uint64_t accum_bar;
    __shared__ uint32_t tmem_base_slot;
    __shared__ int s_flag;

    const int dir = blockIdx.z, tile = blockIdx.x, split = blockIdx.y;
    const GruFwdDir& dd = P.d[dir];
    const int j0 = tile * 32;
    const int kbeg = split * P.kchunk;
    const int kend = min(P.K, kbeg + P.kchunk);
    const int nkb = (kend > kbeg) ? (kend - kbeg + kBlockK - 1) / kBlockK : 0;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem_base = (smem_u32(smem) + 1023u) & ~1023u;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&tma_full[s], 1);
            mbar_init(&mma_full[s], kSplitThreads);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_slot;

    if (warp < 8) {
        // residual pass over the three loaded weight blocks and the state tile
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % STAGES;
            mbar_wait(&tma_full[s], (uint32_t)((kb / STAGES) & 1));
            const uint32_t st = smem_base + (uint32_t)s * kStageBytes;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const uint32_t off = (uint32_t)(tid + i * kSplitThreads) * 16u;       // 768 chunks = 3 x 4096 B
                float4 v;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(st + off));
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(st + kATile + off), "f"(resid(v.x)),
                             "f"(resid(v.y)), "f"(resid(v.z)), "f"(resid(v.w))
                             : "memory");
            }
#pragma unroll
            for (int i = 0; i < ((int)(kBBytes / 16) + kSplitThreads - 1) / kSplitThreads; ++i) {
                const uint32_t qq = (uint32_t)(tid + i * kSplitThreads);
                if (qq < kBBytes / 16) {
                    const uint32_t off = qq * 16u;
                    float4 v;
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(st + 2 * kATile + off));
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(st + 2 * kATile + kBBytes + off),
                                 "f"(resid(v.x)), "f"(resid(v.y)), "f"(resid(v.z)), "f"(resid(v.w))
                                 : "memory");
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(&mma_full[s]);
        }
    } else {
        if (warp == 8 && lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(&empty_bar[s], (uint32_t)(((kb / STAGES) & 1) ^ 1));
                const uint32_t st = smem_base + (uint32_t)s * kStageBytes;
                const int k0 = kbeg + kb * kBlockK;
                mbar_expect_tx(&tma_full[s], 3 * 4096 + kBBytes);
#pragma unroll
                for (int gte = 0; gte < 3; ++gte)
                    tma_load_3d(st + gte * 4096, &P.mapW[dir], &tma_full[s], gte * P.D + j0, k0, 0);
                tma_load_3d(st + 2 * kATile, &P.mapH[dir], &tma_full[s], k0, 0, 0);
            }
        } else if (warp == 9 && lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(BN >> 3) << 17) | ((128u >> 4) << 24);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(&mma_full[s], (uint32_t)((kb / STAGES) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = smem_base + (uint32_t)s * kStageBytes;
                const uint64_t a_raw = desc_mnmajor(st), a_lo = desc_mnmajor(st + kATile);
                const uint64_t b_raw = desc_kmajor(st + 2 * kATile), b_lo = desc_kmajor(st + 2 * kATile + kBBytes);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const uint64_t adv_a = (uint64_t)((kk * 1024) >> 4), adv_b = (uint64_t)((kk * 32) >> 4);
                    const int gstep = kb * 4 + kk;
                    umma_tf32(tmem_d + 3u * BN, a_lo + adv_a, b_raw + adv_b, idesc, gstep != 0 ? 1u : 0u);
                    umma_tf32(tmem_d + 3u * BN, a_raw + adv_a, b_lo + adv_b, idesc, 1u);
                    umma_tf32(tmem_d + (uint32_t)(gstep % 3) * BN, a_raw + adv_a, b_raw + adv_b, idesc, gstep >= 3 ? 1u : 0u);
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(&accum_bar);
        }
        __syncwarp();
        // ------------------------------------------------ epilogue: fix-up + gates (nats.py:341-354)
        const int q = warp - 8, t = tid - 256;
        float acc[BN];
        if (nkb > 0) {
            mbar_wait(&accum_bar, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            load_acc<BN>(tmem_d, q, acc);
        } else {
#pragma unroll
            for (int col = 0; col < BN; ++col) acc[col] = 0.f;
        }
        const bool last = splitk_fixup<BN>(acc, dd.slab, dd.counters + tile, tile, P.nsplit, split, q, lane, t, &s_flag);
        if (last) {
            float* X = reinterpret_cast<float*>(smem + (smem_base - smem_u32(smem)));     // [4][BN][32], stage memory is free now
#pragma unroll
            for (int col = 0; col < BN; ++col) X[(q * BN + col) * 32 + lane] = acc[col];
            named_bar_sync(1, 128);
            const int u = t & 31, bq = t >> 5;
            const int j = j0 + u;
            const int D = P.D;
            if (j < D) {
                for (int b = bq; b < P.B; b += 4) {
                    const float* x = dd.xproj + (long long)b * 3 * D;
                    const float gr = X[(0 * BN + b) * 32 + u] + x[j];
                    const float gu = X[(1 * BN + b) * 32 + u] + x[D + j];
                    const float pp = X[(2 * BN + b) * 32 + u];
                    const float xc = x[2 * D + j];
                    const float r = sigmoidf_(gr), uu = sigmoidf_(gu);
                    const float c = tanhf(pp * r + xc);
                    const float hp = dd.h_prev[(long long)b * dd.ld_hprev + j];
                    const float hn = uu * hp + (1.f - uu) * c;
                    const float m = dd.mask ? dd.mask[b] : 1.f;
                    const float h = m * hn + (1.f - m) * hp;
                    dd.h_out[(long long)b * dd.ld_hout + j] = h;
                    if (dd.r) {
                        const long long idx = (long long)b * D + j;
                        dd.r[idx] = r; dd.u[idx] = uu; dd.c[idx] = c; dd.p[idx] = pp;
                    }
                    if (dd.ctxsum) dd.ctxsum[(long long)b * dd.ld_ctxsum + j] += m * h;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 8) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(kTmemCols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ backward
template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreads, 1) gru_bwd_step_kernel(const __grid_constant__ GruBwdParams P) {
    constexpr uint32_t kBBytes = BN * 128;
    constexpr uint32_t kStageBytes = 2 * kATile + 2 * kBBytes;
    constexpr uint32_t kTmemCols = 4 * BN;

    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t tma_full[STAGES];
    __shared__ __align__(8) uint64_t mma_full[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ __align__(8) uint64_t accum_bar;
    __shared__ uint32_t tmem_base_slot;
    __shared__ int s_flag;

    const int dir = blockIdx.z, tile = blockIdx.x, split = blockIdx.y;
    const GruBwdDir& dd = P.d[dir];
    const int j0 = tile * 128;
    const int kbeg = split * P.kchunk;
    const int kend = min(P.K, kbeg + P.kchunk);
    const int nkb = (kend > kbeg) ? (kend - kbeg + kBlockK - 1) / kBlockK : 0;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem_base = (smem_u32(smem) + 1023u) & ~1023u;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&tma_full[s], 1);
            mbar_init(&mma_full[s], kSplitThreads);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_slot;

    if (warp < 8) {
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % STAGES;
            mbar_wait(&tma_full[s], (uint32_t)((kb / STAGES) & 1));
            const uint32_t st = smem_base + (uint32_t)s * kStageBytes;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t off = (uint32_t)(tid + i * kSplitThreads) * 16u;
                float4 v;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(st + off));
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(st + kATile + off), "f"(resid(v.x)),
                             "f"(resid(v.y)), "f"(resid(v.z)), "f"(resid(v.w))
                             : "memory");
            }
#pragma unroll
            for (int i = 0; i < ((int)(kBBytes / 16) + kSplitThreads - 1) / kSplitThreads; ++i) {
                const uint32_t qq = (uint32_t)(tid + i * kSplitThreads);
                if (qq < kBBytes / 16) {
                    const uint32_t off = qq * 16u;
                    float4 v;
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(st + 2 * kATile + off));
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(st + 2 * kATile + kBBytes + off),
                                 "f"(resid(v.x)), "f"(resid(v.y)), "f"(resid(v.z)), "f"(resid(v.w))
                                 : "memory");
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(&mma_full[s]);
        }
    } else {
        if (warp == 8 && lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(&empty_bar[s], (uint32_t)(((kb / STAGES) & 1) ^ 1));
                const uint32_t st = smem_base + (uint32_t)s * kStageBytes;
                const int k0 = kbeg + kb * kBlockK;
                mbar_expect_tx(&tma_full[s], kATile + kBBytes);
                tma_load_3d(st, &P.mapU[dir], &tma_full[s], k0, j0, 0);
                tma_load_3d(st + 2 * kATile, &P.mapG[dir], &tma_full[s], k0, 0, 0);
            }
        } else if (warp == 9 && lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((128u >> 4) << 24);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(&mma_full[s], (uint32_t)((kb / STAGES) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = smem_base + (uint32_t)s * kStageBytes;
                const uint64_t a_raw = desc_kmajor(st), a_lo = desc_kmajor(st + kATile);
                const uint64_t b_raw = desc_kmajor(st + 2 * kATile), b_lo = desc_kmajor(st + 2 * kATile + kBBytes);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const uint64_t adv = (uint64_t)((kk * 32) >> 4);
                    const int gstep = kb * 4 + kk;
                    umma_tf32(tmem_d + 3u * BN, a_lo + adv, b_raw + adv, idesc, gstep != 0 ? 1u : 0u);
                    umma_tf32(tmem_d + 3u * BN, a_raw + adv, b_lo + adv, idesc, 1u);
                    umma_tf32(tmem_d + (uint32_t)(gstep % 3) * BN, a_raw + adv, b_raw + adv, idesc, gstep >= 3 ? 1u : 0u);
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(&accum_bar);
        }
        __syncwarp();
        // ------------------------------------------------ epilogue: fix-up + reverse of the gate arithmetic
        const int q = warp - 8, t = tid - 256;
        float acc[BN];
        if (nkb > 0) {
            mbar_wait(&accum_bar, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            load_acc<BN>(tmem_d, q, acc);
        } else {
#pragma unroll
            for (int col = 0; col < BN; ++col) acc[col] = 0.f;
        }
        const bool last = splitk_fixup<BN>(acc, dd.slab, dd.counters + tile, tile, P.nsplit, split, q, lane, t, &s_flag);
        if (last) {
            const int D = P.D;
            const int j = j0 + q * 32 + lane;             // this thread's hidden unit; acc[b] = (dG_{t+1}.[U|Ux]^T)[b, j]
            if (j < D) {
#pragma unroll
                for (int b = 0; b < BN; ++b) {          // fully unrolled: acc[] stays in registers
                    if (b >= P.B) continue;
                    const long long idx = (long long)b * D + j;
                    const float m = dd.mask ? dd.mask[b] : 1.f;
                    float dh = acc[b];
                    if (dd.dh_a) dh += dd.dh_a[(long long)b * dd.ld_a + j];
                    if (dd.dh_b) dh += dd.dh_b[(long long)b * dd.ld_b + j];
                    if (dd.mean_grad) dh += m * dd.coef[b] * dd.mean_grad[(long long)b * dd.ld_mean + j];
                    const float r = dd.r[idx], uu = dd.u[idx], c = dd.c[idx], p = dd.p[idx];
                    const float hp = dd.h_prev ? dd.h_prev[(long long)b * dd.ld_hprev + j] : 0.f;
                    const float dhn = m * dh;
                    const float du = dhn * (hp - c);
                    const float dc = dhn * (1.f - uu);
                    const float dpc = dc * (1.f - c * c);
                    const float dp = dpc * r;
                    const float dr = dpc * p;
                    const float dgr = dr * r * (1.f - r);
                    const float dgu = du * uu * (1.f - uu);
                    const long long row3 = (long long)b * 3 * D;
                    dd.dG[row3 + j] = dgr; dd.dG[row3 + D + j] = dgu; dd.dG[row3 + 2 * D + j] = dp;
                    dd.dGx[row3 + j] = dgr; dd.dGx[row3 + D + j] = dgu; dd.dGx[row3 + 2 * D + j] = dpc;
                    dd.dh_elem[idx] = (1.f - m) * dh + dhn * uu;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 8) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(kTmemCols) : "memory");
    }
}

template <int BN, int STAGES>
constexpr size_t step_smem() { return (size_t)STAGES * (2 * kATile + 2 * BN * 128) + 1024; }

template <int BN, int STAGES>
int set_step_attrs() {
    NATS_CUDA_OK(cudaFuncSetAttribute(gru_fwd_step_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)step_smem<BN, STAGES>()));
    NATS_CUDA_OK(cudaFuncSetAttribute(gru_bwd_step_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)step_smem<BN, STAGES>()));
    return 0;
}

inline int pick_bn(int B) { return B <= 32 ? 32 : (B <= 64 ? 64 : 128); }

inline int pick_nsplit(const nats_ctx* ctx, int tiles, int ndir, int K) {
    int s = ctx->num_sms / (tiles * ndir > 0 ? tiles * ndir : 1);
    s = min(s, K / 64);
    s = min(s, 16);
    return max(s, 1);
}

}  // namespace

int gru_step_setup() {
    NATS_TRY((set_step_attrs<32, 5>()));
    NATS_TRY((set_step_attrs<64, 4>()));
    NATS_TRY((set_step_attrs<128, 3>()));
    return 0;
}

static int g_fused_steps = 0;      // NATS_FUSED_STEP=1 enables the fused recurrent-step kernels (slower than the
                                   // split GEMM + gate launches on B200 as of round 1: see DESIGN.md)
void gru_step_enable(int on) { g_fused_steps = on; }
bool gru_step_eligible(int B, int D) {
    return g_fused_steps && gemm_get_tensor_cores() >= 2 && tma_available() && B >= 1 && B <= 128 && D >= 32 &&
           (D % 4) == 0;
}

long long gru_step_slab_floats(int B, int D) {
    const int bn = pick_bn(B);
    const long long tiles = (D + 31) / 32;        // forward tiling is the finer one
    return 2LL * tiles * 16 * 4 * bn * 32;
}
long long gru_step_counter_ints(int D) { return 2LL * ((D + 31) / 32) + 64; }

int gru_step_fwd(const nats_ctx* ctx, cudaStream_t st, const GruStepFwd* dirs, int ndir, int B, int D,
                 float* slab, int* counters) {
    NATS_REQUIRE(ndir >= 1 && ndir <= 2 && gru_step_eligible(B, D), "gru_step_fwd shape");
    GruFwdParams P;
    memset(&P, 0, sizeof(P));
    const int BN = pick_bn(B);
    const int tiles = cdiv(D, 32);
    const int nsplit = pick_nsplit(ctx, tiles, ndir, D);
    P.B = B; P.D = D; P.K = D; P.nsplit = nsplit;
    P.kchunk = ((cdiv(D, nsplit) + 31) / 32) * 32;
    const long long slab_per_dir = (long long)tiles * nsplit * 4 * BN * 32;
    for (int i = 0; i < ndir; ++i) {
        const GruStepFwd& s = dirs[i];
        NATS_REQUIRE((reinterpret_cast<uintptr_t>(s.Ucat) & 15) == 0 && (reinterpret_cast<uintptr_t>(s.h_prev) & 15) == 0 &&
                         (s.ld_hprev & 3) == 0,
                     "gru_step_fwd alignment");
        NATS_TRY(tma_map_3d(s.Ucat, 3LL * D, D, 3LL * D, 1, 0, 32, true, &P.mapW[i]));
        NATS_TRY(tma_map_3d(s.h_prev, D, B, s.ld_hprev, 1, 0, BN, false, &P.mapH[i]));
        GruFwdDir& d = P.d[i];
        d.xproj = s.xproj; d.h_prev = s.h_prev; d.ld_hprev = s.ld_hprev; d.mask = s.mask;
        d.h_out = s.h_out; d.ld_hout = s.ld_hout; d.r = s.r; d.u = s.u; d.c = s.c; d.p = s.p;
        d.ctxsum = s.ctxsum; d.ld_ctxsum = s.ld_ctxsum;
        d.slab = slab + i * slab_per_dir;
        d.counters = counters + i * tiles;
    }
    dim3 grid(tiles, nsplit, ndir);
    ProfScope ps(st, K_GRU_STEP_FWD, 2.0 * ndir * B * 3.0 * D * D, 4.0 * ndir * (3.0 * D * D + 8.0 * B * D));
    if (BN == 32) gru_fwd_step_kernel<32, 5><<<grid, kThreads, step_smem<32, 5>(), st>>>(P);
    else if (BN == 64) gru_fwd_step_kernel<64, 4><<<grid, kThreads, step_smem<64, 4>(), st>>>(P);
    else gru_fwd_step_kernel<128, 3><<<grid, kThreads, step_smem<128, 3>(), st>>>(P);
    NATS_LAUNCH_OK();
    return 0;
}

int gru_step_bwd(const nats_ctx* ctx, cudaStream_t st, const GruStepBwd* dirs, int ndir, int B, int D,
                 float* slab, int* counters) {
    NATS_REQUIRE(ndir >= 1 && ndir <= 2 && gru_step_eligible(B, D), "gru_step_bwd shape");
    GruBwdParams P;
    memset(&P, 0, sizeof(P));
    const int BN = pick_bn(B);
    const int tiles = cdiv(D, 128);
    const int K = 3 * D;
    const int nsplit = pick_nsplit(ctx, tiles, ndir, K);
    P.B = B; P.D = D; P.K = K; P.nsplit = nsplit;
    P.kchunk = ((cdiv(K, nsplit) + 31) / 32) * 32;
    const long long slab_per_dir = (long long)tiles * nsplit * 4 * BN * 32;
    for (int i = 0; i < ndir; ++i) {
        const GruStepBwd& s = dirs[i];
        NATS_REQUIRE((reinterpret_cast<uintptr_t>(s.Ucat) & 15) == 0 && (reinterpret_cast<uintptr_t>(s.dG_next) & 15) == 0,
                     "gru_step_bwd alignment");
        NATS_TRY(tma_map_3d(s.Ucat, K, D, K, 1, 0, 128, false, &P.mapU[i]));
        NATS_TRY(tma_map_3d(s.dG_next, K, B, K, 1, 0, BN, false, &P.mapG[i]));
        GruBwdDir& d = P.d[i];
        d.dh_a = s.g.dh_a; d.ld_a = s.g.ld_a; d.dh_b = s.g.dh_b; d.ld_b = s.g.ld_b;
        d.mean_grad = s.g.mean_grad; d.ld_mean = s.g.ld_mean; d.coef = s.g.coef;
        d.r = s.g.r; d.u = s.g.u; d.c = s.g.c; d.p = s.g.p;
        d.h_prev = s.g.h_prev; d.ld_hprev = s.g.ld_hprev; d.mask = s.g.mask;
        d.dG = s.g.dG; d.dGx = s.g.dGx; d.dh_elem = s.g.dh_elem;
        d.slab = slab + i * slab_per_dir;
        d.counters = counters + i * tiles;
    }
    dim3 grid(tiles, nsplit, ndir);
    ProfScope ps(st, K_GRU_STEP_BWD, 2.0 * ndir * B * 3.0 * D * D, 4.0 * ndir * (3.0 * D * D + 16.0 * B * D));
    if (BN == 32) gru_bwd_step_kernel<32, 5><<<grid, kThreads, step_smem<32, 5>(), st>>>(P);
    else if (BN == 64) gru_bwd_step_kernel<64, 4><<<grid, kThreads, step_smem<64, 4>(), st>>>(P);
    else gru_bwd_step_kernel<128, 3><<<grid, kThreads, step_smem<128, 3>(), st>>>(P);
    NATS_LAUNCH_OK();
    return 0;
}

}  // namespace nats
