// ops_beam.cu -- the distraction re-ranking terms of beam search (nats.py:982-995) and the history
// bookkeeping copies (nats.py:1015-1023) on device.
#include "ops.cuh"

namespace nats {

namespace {

// one CTA per (history step s, hypothesis i): KL(alpha_s || alpha_now), cos-dist(ctx_s, ctx_now), cos-dist(h_s, h_now)
__global__ void __launch_bounds__(256) beam_pair_kernel(const float* __restrict__ hist_alpha,
                                                        const float* __restrict__ hist_ctx,
                                                        const float* __restrict__ hist_state, int len_cap, int hist_len,
                                                        int Tx, int C, int D, const float* __restrict__ cur_alpha,
                                                        const float* __restrict__ cur_ctx,
                                                        const float* __restrict__ cur_state, float* __restrict__ scratch) {
    __shared__ float red[32];
    const int s = blockIdx.x, i = blockIdx.y, tid = threadIdx.x;
    const long long h = (long long)i * len_cap + s;
    float* out = scratch + ((long long)i * hist_len + s) * 3;
    {   // scipy.stats.entropy(pk, qk): both normalised to sum 1, sum pk*log(pk/qk) with 0*log(0) = 0  (nats.py:990)
        const float* p = hist_alpha + h * Tx;
        const float* q = cur_alpha + (long long)i * Tx;
        float sp = 0.f, sq = 0.f;
        for (int t = tid; t < Tx; t += 256) { sp += p[t]; sq += q[t]; }
        sp = block_sum(sp, red);
        sq = block_sum(sq, red);
        float kl = 0.f;
        for (int t = tid; t < Tx; t += 256) {
            const float pn = p[t] / sp, qn = q[t] / sq;
            if (pn > 0.f) kl += pn * logf(pn / qn);
        }
        kl = block_sum(kl, red);
        if (tid == 0) out[0] = kl;
    }
    {   // scipy.spatial.distance.cosine (nats.py:991)
        const float* u = hist_ctx + h * C;
        const float* v = cur_ctx + (long long)i * C;
        float uv = 0.f, uu = 0.f, vv = 0.f;
        for (int t = tid; t < C; t += 256) { const float a = u[t], b = v[t]; uv = fmaf(a, b, uv); uu = fmaf(a, a, uu); vv = fmaf(b, b, vv); }
        uv = block_sum(uv, red); uu = block_sum(uu, red); vv = block_sum(vv, red);
        if (tid == 0) out[1] = 1.f - uv / (sqrtf(uu) * sqrtf(vv));
    }
    {   // nats.py:992
        const float* u = hist_state + h * D;
        const float* v = cur_state + (long long)i * D;
        float uv = 0.f, uu = 0.f, vv = 0.f;
        for (int t = tid; t < D; t += 256) { const float a = u[t], b = v[t]; uv = fmaf(a, b, uv); uu = fmaf(a, a, uu); vv = fmaf(b, b, vv); }
        uv = block_sum(uv, red); uu = block_sum(uu, red); vv = block_sum(vv, red);
        if (tid == 0) out[2] = 1.f - uv / (sqrtf(uu) * sqrtf(vv));
    }
}

__global__ void beam_minmax_kernel(const float* __restrict__ scratch, int hist_len, int live_k, float kl, float cf,
                                   float sf, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= live_k) return;
    float mn = INFINITY, mc = -INFINITY, ms = -INFINITY;
    for (int s = 0; s < hist_len; ++s) {
        const float* v = scratch + ((long long)i * hist_len + s) * 3;
        mn = fminf(mn, v[0]); mc = fmaxf(mc, v[1]); ms = fmaxf(ms, v[2]);
    }
    out[i] = -kl * mn;                 // nats.py:993
    out[live_k + i] = cf * mc;         // nats.py:994
    out[2 * live_k + i] = sf * ms;     // nats.py:995
}

__global__ void beam_reorder_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                    const float* __restrict__ cur, const int32_t* __restrict__ parent, int len_cap,
                                    int hist_len, int dim) {
    const int j = blockIdx.x, s = blockIdx.y;
    const int par = parent[j];
    const float* from = (s < hist_len) ? src + ((long long)par * len_cap + s) * dim : cur + (long long)par * dim;
    float* to = dst + ((long long)j * len_cap + s) * dim;
    for (int t = threadIdx.x; t < dim; t += blockDim.x) to[t] = from[t];
}

}  // namespace

int beam_distraction_scores(cudaStream_t st, const float* hist_alpha, const float* hist_ctx, const float* hist_state,
                            int len_cap, int hist_len, int live_k, int Tx, int C, int D, const float* cur_alpha,
                            const float* cur_ctx, const float* cur_state, float kl, float cf, float sf, float* scratch,
                            float* out) {
    NATS_REQUIRE(hist_len >= 1 && live_k >= 1 && hist_len <= len_cap, "beam history shape");
    dim3 grid(hist_len, live_k);
    ProfScope ps(st, K_BEAM);
    beam_pair_kernel<<<grid, 256, 0, st>>>(hist_alpha, hist_ctx, hist_state, len_cap, hist_len, Tx, C, D, cur_alpha,
                                           cur_ctx, cur_state, scratch);
    NATS_LAUNCH_OK();
    beam_minmax_kernel<<<cdiv(live_k, 64), 64, 0, st>>>(scratch, hist_len, live_k, kl, cf, sf, out);
    NATS_LAUNCH_OK();
    return 0;
}

int beam_reorder_append(cudaStream_t st, const float* src, float* dst, const float* cur, const int32_t* parent,
                        int n_new, int len_cap, int hist_len, int dim) {
    NATS_REQUIRE(n_new >= 1 && hist_len < len_cap, "beam reorder shape");
    dim3 grid(n_new, hist_len + 1);
    ProfScope ps(st, K_BEAM);
    beam_reorder_kernel<<<grid, 256, 0, st>>>(src, dst, cur, parent, len_cap, hist_len, dim);
    NATS_LAUNCH_OK();
    return 0;
}


namespace {

// One CTA per hypothesis row.  Round r selects the largest probability that comes strictly after the previous pick in
// the total order (value descending, index ascending): no marking, no copy of the row, any vocabulary size; the row
// (120 KB at |V| = 30000) is re-read from L1/L2 K times.
__global__ void __launch_bounds__(256) beam_topk_kernel(const float* __restrict__ probs, int V, int K, int mask_unk,
                                                        float* __restrict__ out_p, int32_t* __restrict__ out_idx) {
    __shared__ float s_v[8];
    __shared__ int s_i[8];
    __shared__ float s_pv;
    __shared__ int s_pi;
    const float* row = probs + (long long)blockIdx.x * V;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float pv = INFINITY;
    int pi = -1;
    for (int r = 0; r < K; ++r) {
        float bv = -1.f;
        int bi = 0x7fffffff;
        for (int i = tid; i < V; i += 256) {
            float v = __ldg(row + i);
            if (mask_unk && i == 1) v = 1e-20f;                        // nats.py:975: next_p[:,1] = 1e-20
            if (!(v == v)) v = 0.f;                                      // NaN never wins
            const bool after = (v < pv) || (v == pv && i > pi);
            if (after && (v > bv || (v == bv && i < bi))) { bv = v; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { s_v[warp] = bv; s_i[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            float v = s_v[0];
            int i = s_i[0];
#pragma unroll
            for (int w = 1; w < 8; ++w)
                if (s_v[w] > v || (s_v[w] == v && s_i[w] < i)) { v = s_v[w]; i = s_i[w]; }
            if (i == 0x7fffffff) { v = 0.f; i = -1; }                    // fewer than K candidates
            out_p[(long long)blockIdx.x * K + r] = v;
            out_idx[(long long)blockIdx.x * K + r] = i;
            s_pv = v; s_pi = i;
        }
        __syncthreads();
        pv = s_pv; pi = s_pi;
        if (pi < 0) {                                                    // exhausted: fill the rest
            if (tid == 0)
                for (int q = r + 1; q < K; ++q) { out_p[(long long)blockIdx.x * K + q] = 0.f; out_idx[(long long)blockIdx.x * K + q] = -1; }
            break;
        }
    }
}

}  // namespace

int beam_topk(cudaStream_t st, const float* probs, int n, int V, int K, int mask_unk, float* out_p, int32_t* out_idx) {
    NATS_REQUIRE(n >= 1 && V >= 1 && K >= 1, "beam_topk shape");
    beam_topk_kernel<<<n, 256, 0, st>>>(probs, V, K, mask_unk, out_p, out_idx);
    NATS_LAUNCH_OK();
    return 0;
}

}  // namespace nats
