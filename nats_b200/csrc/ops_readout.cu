// ops_readout.cu -- softmax / cross-entropy rows of the readout (nats.py:763-770, 861-864).
#include "ops.cuh"

#include <cooperative_groups.h>

namespace nats {

namespace {

constexpr int kRowThreads = 512;

__global__ void __launch_bounds__(kRowThreads) nll_rows_kernel(const float* __restrict__ logits, int V,
                                                               const int64_t* __restrict__ y,
                                                               const float* __restrict__ ymask,
                                                               float* __restrict__ lse, float* __restrict__ rowcost) {
    __shared__ float red[32];
    const int r = blockIdx.x;
    const float* row = logits + (long long)r * V;
    float mx = -INFINITY;
    for (int v = threadIdx.x; v < V; v += kRowThreads) mx = fmaxf(mx, row[v]);
    mx = block_max(mx, red);
    float s = 0.f;
    for (int v = threadIdx.x; v < V; v += kRowThreads) s += expf(row[v] - mx);
    s = block_sum(s, red);
    if (threadIdx.x == 0) {
        const float l = mx + logf(s);
        lse[r] = l;
        const long long id = y[r];
        const float tgt = (id >= 0 && id < V) ? row[id] : 0.f;
        rowcost[r] = (l - tgt) * (ymask ? ymask[r] : 1.f);      // nats.py:768-770
    }
}

__global__ void dlogits_kernel(float* __restrict__ logits, int rows, int V, const int64_t* __restrict__ y,
                               const float* __restrict__ ymask, const float* __restrict__ lse, float scale) {
    // rows on grid.x (2^31 - 1 blocks), vocabulary chunks on grid.y: rows = Ty*B exceeds the 65535 limit of grid.y for
    // large batches (e.g. batch 160 at maxlen 500)
    const int r = blockIdx.x;
    const float w = (ymask ? ymask[r] : 1.f) * scale;
    const float l = lse[r];
    const long long id = y[r];
    float* row = logits + (long long)r * V;
    for (int v = blockIdx.y * blockDim.x + threadIdx.x; v < V; v += gridDim.y * blockDim.x) {
        float p = expf(row[v] - l);
        if (v == id) p -= 1.f;
        row[v] = p * w;
    }
}

// counter-based uniform in [0,1): splitmix64 of (seed, step, row)
__device__ __forceinline__ float uniform01(uint64_t seed, uint64_t step, uint64_t row) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (step * 0x100000001B3ull + row + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (float)(z >> 40) * (1.0f / 16777216.0f);
}

__global__ void __launch_bounds__(kRowThreads) softmax_sample_kernel(const float* __restrict__ logits, int V,
                                                                     float* __restrict__ probs,
                                                                     int64_t* __restrict__ sample, uint64_t seed,
                                                                     uint64_t step) {
    __shared__ float red[32];
    __shared__ float chunk_sum[kRowThreads];
    __shared__ int s_pick;
    const int r = blockIdx.x, tid = threadIdx.x;
    const float* row = logits + (long long)r * V;
    float* prow = probs + (long long)r * V;
    float mx = -INFINITY;
    for (int v = tid; v < V; v += kRowThreads) mx = fmaxf(mx, row[v]);
    mx = block_max(mx, red);
    float s = 0.f;
    for (int v = tid; v < V; v += kRowThreads) s += expf(row[v] - mx);
    s = block_sum(s, red);
    const float inv = 1.f / s;
    // contiguous chunk per thread so that the inverse-CDF walk is in vocabulary order
    const int per = (V + kRowThreads - 1) / kRowThreads;
    const int v0 = tid * per, v1 = min(V, v0 + per);
    float cs = 0.f;
    for (int v = v0; v < v1; ++v) {
        const float p = expf(row[v] - mx) * inv;                 // nats.py:861
        prow[v] = p;
        cs += p;
    }
    chunk_sum[tid] = cs;
    if (tid == 0) s_pick = -1;
    __syncthreads();
    if (sample) {
        const float u = uniform01(seed, step, (uint64_t)r);
        if (tid == 0) {
            float c = 0.f;
            int pick = kRowThreads - 1;
            for (int i = 0; i < kRowThreads; ++i) {
                if (u < c + chunk_sum[i]) { pick = i; break; }
                c += chunk_sum[i];
            }
            s_pick = pick;
            red[0] = c;
        }
        __syncthreads();
        if (tid == s_pick) {                                      // nats.py:864 (multinomial -> argmax)
            float c = red[0];
            int choice = (v1 > v0) ? (v1 - 1) : (V - 1);
            for (int v = v0; v < v1; ++v) {
                c += prow[v];
                if (u < c) { choice = v; break; }
            }
            sample[r] = choice;
        }
    }
}

// Beam-search shapes (a handful of rows, no sampling, |V| <= 32768): one CTA per row leaves 138 SMs idle and walks the
// row three times.  A cluster of 8 CTAs per row keeps its eighth of the row in registers (one read), and the row maximum
// and the normaliser are combined through distributed shared memory in a fixed order (same value in every CTA).
constexpr int kSmCluster = 8, kSmThreads = 512, kSmPer = 8;

__global__ void __cluster_dims__(kSmCluster, 1, 1) __launch_bounds__(kSmThreads)
    softmax_cluster_kernel(const float* __restrict__ logits, int V, float* __restrict__ probs) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ float red[32];
    __shared__ float s_part[2];
    const int seg = blockIdx.x, tid = threadIdx.x;
    const float* row = logits + (long long)blockIdx.y * V;
    float* prow = probs + (long long)blockIdx.y * V;
    const int seglen = (V + kSmCluster - 1) / kSmCluster;
    const int base = seg * seglen, end = min(V, base + seglen);
    float x[kSmPer];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < kSmPer; ++k) {
        const int i = base + k * kSmThreads + tid;
        x[k] = (i < end) ? row[i] : -INFINITY;
        mx = fmaxf(mx, x[k]);
    }
    mx = block_max(mx, red);
    if (tid == 0) s_part[0] = mx;
    cluster.sync();
    float gmx = -INFINITY;
#pragma unroll
    for (int q = 0; q < kSmCluster; ++q) gmx = fmaxf(gmx, *cluster.map_shared_rank(&s_part[0], q));
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kSmPer; ++k) {
        const int i = base + k * kSmThreads + tid;
        x[k] = (i < end) ? expf(x[k] - gmx) : 0.f;
        s += x[k];
    }
    s = block_sum(s, red);
    if (tid == 0) s_part[1] = s;
    cluster.sync();
    float gs = 0.f;
#pragma unroll
    for (int q = 0; q < kSmCluster; ++q) gs += *cluster.map_shared_rank(&s_part[1], q);
    const float inv = 1.f / gs;
#pragma unroll
    for (int k = 0; k < kSmPer; ++k) {
        const int i = base + k * kSmThreads + tid;
        if (i < end) prow[i] = x[k] * inv;                               // nats.py:861
    }
    cluster.sync();                                                      // remote reads done before any CTA exits
}

}  // namespace

int nll_rows(cudaStream_t st, const float* logits, int rows, int V, const int64_t* y, const float* ymask, float* lse,
             float* rowcost) {
    if (rows == 0) return 0;
    ProfScope ps(st, K_NLL, 0.0, 8.0 * rows * V);
    nll_rows_kernel<<<rows, kRowThreads, 0, st>>>(logits, V, y, ymask, lse, rowcost);
    NATS_LAUNCH_OK();
    return 0;
}
int dlogits_inplace(cudaStream_t st, float* logits, int rows, int V, const int64_t* y, const float* ymask,
                    const float* lse, float scale) {
    if (rows == 0) return 0;
    int gy = cdiv(V, 256 * 4);
    if (gy < 1) gy = 1;
    if (gy > 65535) gy = 65535;
    dim3 grid(rows, gy);
    ProfScope ps(st, K_DLOGITS, 0.0, 8.0 * rows * V);
    dlogits_kernel<<<grid, 256, 0, st>>>(logits, rows, V, y, ymask, lse, scale);
    NATS_LAUNCH_OK();
    return 0;
}
int softmax_sample_rows(cudaStream_t st, const float* logits, int rows, int V, float* probs, int64_t* sample,
                        uint64_t seed, uint64_t step) {
    if (rows == 0) return 0;
    ProfScope ps(st, K_SOFTMAX_SAMPLE, 0.0, 16.0 * rows * V);
    static const int force_simple = [] { const char* e = getenv("NATS_SOFTMAX_SIMPLE"); return e && atoi(e) != 0; }();
    if (!force_simple && sample == nullptr && rows <= 148 && V <= kSmCluster * kSmThreads * kSmPer) {
        softmax_cluster_kernel<<<dim3(kSmCluster, rows), kSmThreads, 0, st>>>(logits, V, probs);
        NATS_LAUNCH_OK();
        return 0;
    }
    softmax_sample_kernel<<<rows, kRowThreads, 0, st>>>(logits, V, probs, sample, seed, step);
    NATS_LAUNCH_OK();
    return 0;
}

}  // namespace nats
