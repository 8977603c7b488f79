// enc_persistent.cu -- the bidirectional GRU encoder recurrence (nats.py:336-372, both directions) as ONE persistent,
// weight-stationary kernel per pass.
//
// Why: a recurrent step is a [B x D] x [D x 3D] product with B = 32: launched per step it costs ~16 us (split-K tensor-core
// product + gate kernel), most of it fixed overhead and re-streaming the 12 MB weight matrix from L2.  Here every CTA
// owns `upc` hidden units of one direction for the whole sequence:
//   * its 3*upc columns of [U|Ux] (forward) / its upc rows of [U|Ux] (backward) are loaded ONCE into shared memory
//     (168 KB at D = 1000, upc = 14: 72 CTAs per direction, 144 of the 148 SMs);
//   * per step it streams the previous state h_{t-1} (forward) / the gate derivatives dG_{t+1} (backward) from L2 in
//     128-deep chunks (register-staged, transposed + XOR-swizzled into shared memory, next chunk in flight during compute),
//     accumulates an 8(batch) x 4(column) register tile per thread with exact fp32 FMAs, K split over 8 thread groups;
//   * since a CTA sees the FULL K extent of its units, the GRU gate arithmetic (and its reverse) is finished in place --
//     no split-K slabs, no second kernel -- and h_t / dG_t are published for the other CTAs;
//   * one arrive/spin barrier per direction and step (monotonic counter in L2) orders the steps.
// The grid must be co-resident (2*P CTAs <= #SMs, one CTA per SM by shared-memory footprint); a spinning CTA traps
// after ~2 s instead of hanging the GPU.
#include "ops.cuh"

namespace nats {

namespace {

constexpr int kKC = 128;        // chunk depth
constexpr int kTB = 8;          // register tile: batch rows
constexpr int kTC = 8;          // register tile: columns  (8x8: 4 LDS.128 per 64 FMA -- the shared-memory port keeps up)
constexpr int kMaxThreads = 384;

struct EncPFwd {
    const float* Ucat[2];       // [D,3D]
    const float* xproj[2];      // [Tx*n,3D] by source position
    const float* mask;          // [Tx,n] or NULL
    float* cc;                  // [Tx,n,2D]
    float* r[2]; float* u[2]; float* c[2]; float* p[2];   // [Tx*n,D] by position, or NULL
    float* ctxsum;              // [n,2D]
    unsigned* bar;              // [2] zero-initialised
    long long* dbg;             // optional [8] phase cycle counters of CTA (0,0)
    int Tx, n, D, upc, P, BP, NS;   // BP = batch padded to a power of two >= 8; NS = slabs of the K-split reduction
};

struct EncPBwd {
    const float* Ucat[2];
    const float* dcc;           // [Tx,n,2D]  d cost / d context
    const float* mean_grad;     // [n,2D] or NULL
    const float* coef;          // [n]
    const float* mask;          // [Tx,n]
    const float* cc;            // [Tx,n,2D]  saved states
    const float* r[2]; const float* u[2]; const float* c[2]; const float* p[2];
    float* dG[2]; float* dGx[2];   // [Tx*n,3D] by position
    unsigned* bar;
    int Tx, n, D, upc, P, BP, NS;
};

__device__ __forceinline__ float4 ldcg4(const float* p) {
    float4 r;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float ldcg1(const float* p) {
    float r;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

// all CTAs of one direction: arrive, then wait until `target` arrivals (monotonic counter)
__device__ __forceinline__ void dir_barrier(unsigned* ctr, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        const long long t0 = clock64();
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
            if (v >= target) break;
            __nanosleep(40);
            if (clock64() - t0 > 4000000000LL) __trap();        // ~2 s: never hang the device
        } while (true);
        __threadfence();
    }
    __syncthreads();
}

// Stage one [rows x kKC] chunk of a row-major [rows, ld] global matrix into shared memory TRANSPOSED as hT[kk][row],
// float4 columns XOR-swizzled with (kk >> 2) & 7 (conflict-free transposing stores, 16-byte aligned reads).
template <int NV>
__device__ __forceinline__ void chunk_load(const float* __restrict__ src, long long ld, int rows, int K, int k0, int BP,
                                           int tid, int nthreads, float4 (&v)[NV]) {
    const int per_row = kKC / 4;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int f = tid + i * nthreads;
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        if (f < BP * per_row) {
            const int b = f / per_row, k4 = f - b * per_row;
            const int k = k0 + 4 * k4;
            if (b < rows && k < K) {
                const float* p = src + (long long)b * ld + k;
                if (k + 3 < K) val = ldcg4(p);
                else {
                    val.x = ldcg1(p);
                    if (k + 1 < K) val.y = ldcg1(p + 1);
                    if (k + 2 < K) val.z = ldcg1(p + 2);
                }
            }
        }
        v[i] = val;
    }
}
template <int NV>
__device__ __forceinline__ void chunk_store(float* __restrict__ hT, int BP, int tid, int nthreads, const float4 (&v)[NV]) {
    const int per_row = kKC / 4;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int f = tid + i * nthreads;
        if (f < BP * per_row) {
            const int b = f / per_row, k4 = f - b * per_row;
            const int col = 4 * (((b >> 2) ^ (k4 & ((BP >> 2) - 1) & 7))) + (b & 3);
            hT[(4 * k4 + 0) * BP + col] = v[i].x;
            hT[(4 * k4 + 1) * BP + col] = v[i].y;
            hT[(4 * k4 + 2) * BP + col] = v[i].z;
            hT[(4 * k4 + 3) * BP + col] = v[i].w;
        }
    }
}

// acc[8][8] += hT[kk][bg*8 .. +7] (x) W[k][cg*8 .. +7]   for this thread's k-slice (kk = ks, ks+KS, ...) of the chunk
__device__ __forceinline__ void chunk_fma(const float* __restrict__ hT, const float* __restrict__ Wt, int ncolp, int BP,
                                          int k0, int K, int ks, int KS, int bg, int cg, float (&acc)[kTB][kTC]) {
    const int kmax = min(kKC, K - k0);
    const int swm = ((BP >> 2) - 1) & 7;
#pragma unroll 2
    for (int kk = ks; kk < kmax; kk += KS) {
        const int sw = (kk >> 2) & swm;
        const float4 h0 = *reinterpret_cast<const float4*>(hT + kk * BP + 4 * ((2 * bg) ^ sw));
        const float4 h1 = *reinterpret_cast<const float4*>(hT + kk * BP + 4 * ((2 * bg + 1) ^ sw));
        const float* wr = Wt + (size_t)(k0 + kk) * ncolp + 8 * cg;
        const float4 w0 = *reinterpret_cast<const float4*>(wr);
        const float4 w1 = *reinterpret_cast<const float4*>(wr + 4);
        const float hv[kTB] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
        const float wv[kTC] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int i = 0; i < kTB; ++i)
#pragma unroll
            for (int j = 0; j < kTC; ++j) acc[i][j] = fmaf(hv[i], wv[j], acc[i][j]);
    }
}

// K-split reduction: group ks adds its tile into slab (ks % kNS) in round (ks / kNS); rounds are separated by barriers.
__device__ __forceinline__ void reduce_to_slabs(float* __restrict__ slabs, int BP, int ncolp, int ks, int KS, int NS, int bg,
                                                int cg, bool worker, const float (&acc)[kTB][kTC]) {
    const int rounds = (KS + NS - 1) / NS;
    for (int r = 0; r < rounds; ++r) {
        if (worker && ks / NS == r) {
            float* sl = slabs + (size_t)(ks % NS) * BP * ncolp;
#pragma unroll
            for (int i = 0; i < kTB; ++i) {
                float4* p0 = reinterpret_cast<float4*>(sl + (size_t)(bg * kTB + i) * ncolp + cg * kTC);
                float4 a0 = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
                float4 a1 = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
                if (r > 0) {
                    const float4 o0 = p0[0], o1 = p0[1];
                    a0.x += o0.x; a0.y += o0.y; a0.z += o0.z; a0.w += o0.w;
                    a1.x += o1.x; a1.y += o1.y; a1.z += o1.z; a1.w += o1.w;
                }
                p0[0] = a0; p0[1] = a1;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ forward
template <int NV>
__global__ void __launch_bounds__(kMaxThreads, 1) enc_persist_fwd_kernel(const __grid_constant__ EncPFwd a) {
    extern __shared__ __align__(16) float sm[];
    const int dir = blockIdx.y, D = a.D, n = a.n, C = 2 * D, D3 = 3 * D, upc = a.upc;
    const int j0 = blockIdx.x * upc;
    const int nu = min(upc, D - j0);
    const int ncol = 3 * upc, ncolp = (ncol + 7) & ~7;
    const int BP = a.BP;
    const int NBG = BP / kTB, NCG = ncolp / kTC;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int KS = nthreads / (NBG * NCG);
    float* Wt = sm;                                   // [D][ncolp]   column = gate*upc + unit
    float* buf = Wt + (size_t)D * ncolp;              // chunk ring [2][kKC][BP]  /  reduction slabs [NS][BP][ncolp]
    const int ring = kKC * BP;

    {   // weights of this CTA's units, once
        const float* U = a.Ucat[dir];
        for (int idx = tid; idx < D * ncolp; idx += nthreads) {
            const int k = idx / ncolp, col = idx - k * ncolp;
            const int g = col / upc, uu = col - g * upc;
            Wt[idx] = (col < ncol && uu < nu) ? __ldg(U + (long long)k * D3 + g * D + j0 + uu) : 0.f;
        }
    }
    const int cg = tid % NCG, bg = (tid / NCG) % NBG, ks = tid / (NCG * NBG);
    const bool worker = ks < KS;
    // this thread's (batch, unit) pairs in the gate phase: the same ones at every step -> carries live in registers
    constexpr int kPairs = 2;
    float h_carry[kPairs], cs_carry[kPairs];
#pragma unroll
    for (int q = 0; q < kPairs; ++q) { h_carry[q] = 0.f; cs_carry[q] = 0.f; }
    __syncthreads();

    long long t_comp = 0, t_gate = 0, t_bar = 0, t_first = 0;
    for (int s = 0; s < a.Tx; ++s) {
        const int pos = dir == 0 ? s : a.Tx - 1 - s;
        const int prev = dir == 0 ? pos - 1 : pos + 1;
        const long long c0 = clock64();
        long long c1 = c0;
        // gate inputs of this step (independent of h_{t-1}): requested now, consumed after the product
        float gx[kPairs][3], gm[kPairs];
#pragma unroll
        for (int q = 0; q < kPairs; ++q) {
            const int idx = tid + q * nthreads;
            gx[q][0] = gx[q][1] = gx[q][2] = 0.f; gm[q] = 1.f;
            if (idx < n * nu) {
                const int b = idx / nu, uu = idx - b * nu;
                const float* x = a.xproj[dir] + ((long long)pos * n + b) * D3 + j0 + uu;
                gx[q][0] = __ldg(x); gx[q][1] = __ldg(x + D); gx[q][2] = __ldg(x + 2 * D);
                if (a.mask) gm[q] = __ldg(a.mask + (long long)pos * n + b);
            }
        }
        float acc[kTB][kTC];
#pragma unroll
        for (int i = 0; i < kTB; ++i)
#pragma unroll
            for (int j = 0; j < kTC; ++j) acc[i][j] = 0.f;

        if (s > 0) {
            const float* hsrc = a.cc + (long long)prev * n * C + dir * D;     // rows b, stride C
            const int nch = (D + kKC - 1) / kKC;
            float4 regs[NV];
            chunk_load<NV>(hsrc, C, n, D, 0, BP, tid, nthreads, regs);
            chunk_store<NV>(buf, BP, tid, nthreads, regs);
            __syncthreads();
            c1 = clock64();
            for (int ch = 0; ch < nch; ++ch) {
                const float* cur = buf + (ch & 1) * ring;
                if (ch + 1 < nch) chunk_load<NV>(hsrc, C, n, D, (ch + 1) * kKC, BP, tid, nthreads, regs);
                if (worker) chunk_fma(cur, Wt, ncolp, BP, ch * kKC, D, ks, KS, bg, cg, acc);
                if (ch + 1 < nch) chunk_store<NV>(buf + ((ch + 1) & 1) * ring, BP, tid, nthreads, regs);
                __syncthreads();
            }
            reduce_to_slabs(buf, BP, ncolp, ks, KS, a.NS, bg, cg, worker, acc);
        }
        const long long c2 = clock64();
        // gates for the (b, unit) pairs of this thread (nats.py:341-354)
#pragma unroll
        for (int q = 0; q < kPairs; ++q) {
            const int idx = tid + q * nthreads;
            if (idx < n * nu) {
                const int b = idx / nu, uu = idx - b * nu;
                const int j = j0 + uu;
                float gr = gx[q][0], gu = gx[q][1], pp = 0.f;
                if (s > 0) {
                    const int nsl = KS < a.NS ? KS : a.NS;
                    for (int sl = 0; sl < nsl; ++sl) {
                        const float* pr = buf + ((size_t)sl * BP + b) * ncolp;
                        gr += pr[uu]; gu += pr[upc + uu]; pp += pr[2 * upc + uu];
                    }
                }
                const float r = sigmoidf_(gr), uz = sigmoidf_(gu);
                const float c = tanhf(pp * r + gx[q][2]);
                const float hp = h_carry[q];
                const float hn = uz * hp + (1.f - uz) * c;
                const float m = gm[q];
                const float h = m * hn + (1.f - m) * hp;
                h_carry[q] = h;
                cs_carry[q] += m * h;
                a.cc[((long long)pos * n + b) * C + dir * D + j] = h;
                if (a.r[dir]) {
                    const long long o = ((long long)pos * n + b) * D + j;
                    a.r[dir][o] = r; a.u[dir][o] = uz; a.c[dir][o] = c; a.p[dir][o] = pp;
                }
            }
        }
        const long long c3 = clock64();
        if (s + 1 < a.Tx) dir_barrier(a.bar + dir, (unsigned)a.P * (unsigned)(s + 1));
        const long long c4 = clock64();
        t_first += c1 - c0; t_comp += c2 - c1; t_gate += c3 - c2; t_bar += c4 - c3;
    }
#pragma unroll
    for (int q = 0; q < kPairs; ++q) {
        const int idx = tid + q * nthreads;
        if (idx < n * nu) {
            const int b = idx / nu, uu = idx - b * nu;
            a.ctxsum[(long long)b * C + dir * D + j0 + uu] += cs_carry[q];
        }
    }
    if (a.dbg && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) {
        a.dbg[0] = t_first; a.dbg[1] = t_comp; a.dbg[2] = t_gate; a.dbg[3] = t_bar;
    }
}

// ------------------------------------------------------------------------------------------------ backward
// CTA owns units j0..j0+upc-1: out[b][u] = sum_k dG_{t+1}[b,k] * Ucat[j0+u, k]   (K = 3D), then the gate backward of
// step t for its units; the elementwise carry of d h stays in shared memory across steps.
template <int NV>
__global__ void __launch_bounds__(kMaxThreads, 1) enc_persist_bwd_kernel(const __grid_constant__ EncPBwd a) {
    extern __shared__ __align__(16) float sm[];
    const int dir = blockIdx.y, D = a.D, n = a.n, C = 2 * D, D3 = 3 * D, upc = a.upc;
    const int j0 = blockIdx.x * upc;
    const int nu = min(upc, D - j0);
    const int ncolp = (upc + 7) & ~7;
    const int BP = a.BP;
    const int NBG = BP / kTB, NCG = ncolp / kTC;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int KS = nthreads / (NBG * NCG);
    float* Wt = sm;                                   // [3D][ncolp]   Wt[k][u] = Ucat[j0+u, k]
    float* buf = Wt + (size_t)D3 * ncolp;             // chunk ring [2][kKC][BP]  /  reduction slabs [NS][BP][ncolp]
    const int ring = kKC * BP;

    {
        const float* U = a.Ucat[dir];
        for (int idx = tid; idx < D3 * ncolp; idx += nthreads) {
            const int uu = idx / D3, k = idx - uu * D3;      // coalesced along k
            Wt[(size_t)k * ncolp + uu] = (uu < nu) ? __ldg(U + (long long)(j0 + uu) * D3 + k) : 0.f;
        }
    }
    const int cg = tid % NCG, bg = (tid / NCG) % NBG, ks = tid / (NCG * NBG);
    const bool worker = ks < KS;
    constexpr int kPairs = 2;
    float carry[kPairs];                              // elementwise part of d h_{t-1} of this thread's (b, unit) pairs
#pragma unroll
    for (int q = 0; q < kPairs; ++q) carry[q] = 0.f;
    __syncthreads();

    for (int s = a.Tx - 1; s >= 0; --s) {
        const int pos = dir == 0 ? s : a.Tx - 1 - s;
        const int prev = dir == 0 ? pos - 1 : pos + 1;            // position of h_{t-1}
        const int nextp = dir == 0 ? pos + 1 : pos - 1;           // position processed by step s+1
        // gate-backward inputs of this step: requested before the product, consumed after it
        float g_dcc[kPairs], g_m[kPairs], g_r[kPairs], g_u[kPairs], g_c[kPairs], g_p[kPairs], g_hp[kPairs], g_mg[kPairs];
#pragma unroll
        for (int q = 0; q < kPairs; ++q) {
            const int idx = tid + q * nthreads;
            g_dcc[q] = g_r[q] = g_u[q] = g_c[q] = g_p[q] = g_hp[q] = g_mg[q] = 0.f; g_m[q] = 1.f;
            if (idx < n * nu) {
                const int b = idx / nu, uu = idx - b * nu;
                const int j = j0 + uu;
                const long long o = ((long long)pos * n + b) * D + j;
                g_dcc[q] = __ldg(a.dcc + ((long long)pos * n + b) * C + dir * D + j);
                g_m[q] = __ldg(a.mask + (long long)pos * n + b);
                g_r[q] = __ldg(a.r[dir] + o); g_u[q] = __ldg(a.u[dir] + o); g_c[q] = __ldg(a.c[dir] + o); g_p[q] = __ldg(a.p[dir] + o);
                g_hp[q] = s > 0 ? __ldg(a.cc + ((long long)prev * n + b) * C + dir * D + j) : 0.f;
                if (a.mean_grad) g_mg[q] = __ldg(a.coef + b) * __ldg(a.mean_grad + (long long)b * C + dir * D + j);
            }
        }
        float acc[kTB][kTC];
#pragma unroll
        for (int i = 0; i < kTB; ++i)
#pragma unroll
            for (int j = 0; j < kTC; ++j) acc[i][j] = 0.f;

        if (s < a.Tx - 1) {
            const float* gsrc = a.dG[dir] + (long long)nextp * n * D3;        // [n, 3D]
            const int nch = (D3 + kKC - 1) / kKC;
            float4 regs[NV];
            chunk_load<NV>(gsrc, D3, n, D3, 0, BP, tid, nthreads, regs);
            chunk_store<NV>(buf, BP, tid, nthreads, regs);
            __syncthreads();
            for (int ch = 0; ch < nch; ++ch) {
                const float* cur = buf + (ch & 1) * ring;
                if (ch + 1 < nch) chunk_load<NV>(gsrc, D3, n, D3, (ch + 1) * kKC, BP, tid, nthreads, regs);
                if (worker) chunk_fma(cur, Wt, ncolp, BP, ch * kKC, D3, ks, KS, bg, cg, acc);
                if (ch + 1 < nch) chunk_store<NV>(buf + ((ch + 1) & 1) * ring, BP, tid, nthreads, regs);
                __syncthreads();
            }
            reduce_to_slabs(buf, BP, ncolp, ks, KS, a.NS, bg, cg, worker, acc);
        }
#pragma unroll
        for (int q = 0; q < kPairs; ++q) {
            const int idx = tid + q * nthreads;
            if (idx < n * nu) {
                const int b = idx / nu, uu = idx - b * nu;
                const int j = j0 + uu;
                const float m = g_m[q];
                float dh = g_dcc[q] + carry[q] + m * g_mg[q];
                if (s < a.Tx - 1) {
                    const int nsl = KS < a.NS ? KS : a.NS;
                    for (int sl = 0; sl < nsl; ++sl) dh += buf[((size_t)sl * BP + b) * ncolp + uu];
                }
                const float r = g_r[q], uz = g_u[q], c = g_c[q], p = g_p[q], hp = g_hp[q];
                const float dhn = m * dh;
                const float du = dhn * (hp - c);
                const float dc = dhn * (1.f - uz);
                const float dpc = dc * (1.f - c * c);
                const float dp = dpc * r;
                const float dr = dpc * p;
                const float dgr = dr * r * (1.f - r);
                const float dgu = du * uz * (1.f - uz);
                const long long row3 = ((long long)pos * n + b) * D3;
                a.dG[dir][row3 + j] = dgr; a.dG[dir][row3 + D + j] = dgu; a.dG[dir][row3 + 2 * D + j] = dp;
                a.dGx[dir][row3 + j] = dgr; a.dGx[dir][row3 + D + j] = dgu; a.dGx[dir][row3 + 2 * D + j] = dpc;
                carry[q] = (1.f - m) * dh + dhn * uz;
            }
        }
        if (s > 0) dir_barrier(a.bar + dir, (unsigned)a.P * (unsigned)(a.Tx - s));
    }
}

struct PersistPlan {
    int upc_f, P_f, threads_fwd, nv_fwd, ns_fwd; size_t smem_fwd;
    int upc_b, P_b, threads_bwd, nv_bwd, ns_bwd; size_t smem_bwd;
    int BP;
    bool ok;
};

PersistPlan plan(const nats_ctx* ctx, int n, int D) {
    PersistPlan pl;
    memset(&pl, 0, sizeof(pl));
    if (n < 1 || n > 64 || D < 4 || (D % 4) != 0) return pl;
    const int per_dir = ctx->num_sms / 2;
    if (per_dir < 1) return pl;
    int BP = 8;
    while (BP < n) BP *= 2;                           // power of two: the chunk swizzle XORs float4 column indices
    pl.BP = BP;
    const int NBG = BP / kTB;
    const size_t ring = 2u * kKC * BP;
    auto shape = [&](int ncolp, int ks_cap, int* threads, int* ns, int* nv) -> bool {
        const int NCG = ncolp / kTC, grp = NBG * NCG;
        int ks = kMaxThreads / grp;
        if (ks > ks_cap) ks = ks_cap;
        if (ks < 1) return false;
        int th = ((ks * grp + 31) / 32) * 32;
        if (th > kMaxThreads) th = (kMaxThreads / 32) * 32;
        if (th < 128) th = 128;
        *threads = th;
        int s = (int)(ring / ((size_t)BP * ncolp));   // reduction slabs live in the idle chunk ring
        const int KS = th / grp;
        if (s > KS) s = KS;
        if (s < 1) return false;
        *ns = s;
        *nv = (BP * (kKC / 4) + th - 1) / th;
        return true;
    };
    {   // forward: 3*upc columns per CTA
        pl.upc_f = (D + per_dir - 1) / per_dir;
        pl.P_f = (D + pl.upc_f - 1) / pl.upc_f;
        const int ncolp = (3 * pl.upc_f + 7) & ~7;
        if (!shape(ncolp, 16, &pl.threads_fwd, &pl.ns_fwd, &pl.nv_fwd)) return pl;
        const size_t slabs = (size_t)pl.ns_fwd * BP * ncolp;
        pl.smem_fwd = ((size_t)D * ncolp + (ring > slabs ? ring : slabs)) * sizeof(float);
        if ((long long)n * pl.upc_f > 2LL * pl.threads_fwd) return pl;
    }
    {   // backward: upc columns per CTA (multiple of 8: no padded columns)
        pl.upc_b = (((D + per_dir - 1) / per_dir) + 7) & ~7;
        pl.P_b = (D + pl.upc_b - 1) / pl.upc_b;
        const int ncolp = pl.upc_b;
        if (!shape(ncolp, 32, &pl.threads_bwd, &pl.ns_bwd, &pl.nv_bwd)) return pl;
        const size_t slabs = (size_t)pl.ns_bwd * BP * ncolp;
        pl.smem_bwd = ((size_t)3 * D * ncolp + (ring > slabs ? ring : slabs)) * sizeof(float);
        if ((long long)n * pl.upc_b > 2LL * pl.threads_bwd) return pl;
    }
    const size_t lim = (size_t)ctx->max_smem_optin - 1024;
    pl.ok = pl.smem_fwd <= lim && pl.smem_bwd <= lim && 2 * pl.P_f <= ctx->num_sms && 2 * pl.P_b <= ctx->num_sms &&
            pl.nv_fwd <= 8 && pl.nv_bwd <= 8;
    return pl;
}

static int g_persist = 0;

}  // namespace

void enc_persistent_enable(int on) { g_persist = on; }

int enc_persistent_setup(const nats_ctx* ctx) {
    const int lim = ctx->max_smem_optin - 1024;
    NATS_CUDA_OK(cudaFuncSetAttribute(enc_persist_fwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    NATS_CUDA_OK(cudaFuncSetAttribute(enc_persist_fwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    NATS_CUDA_OK(cudaFuncSetAttribute(enc_persist_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    NATS_CUDA_OK(cudaFuncSetAttribute(enc_persist_bwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    return 0;
}

bool enc_persistent_eligible(const nats_ctx* ctx, int n, int D) { return g_persist && plan(ctx, n, D).ok; }

int enc_persistent_fwd(const nats_ctx* ctx, cudaStream_t st, const EncPersistFwdArgs& g) {
    const PersistPlan pl = plan(ctx, g.n, g.D);
    NATS_REQUIRE(pl.ok, "persistent encoder not applicable");
    NATS_REQUIRE((reinterpret_cast<uintptr_t>(g.cc) & 15) == 0, "cc alignment");
    EncPFwd a;
    memset(&a, 0, sizeof(a));
    for (int d = 0; d < 2; ++d) {
        a.Ucat[d] = g.Ucat[d]; a.xproj[d] = g.xproj[d];
        a.r[d] = g.r[d]; a.u[d] = g.u[d]; a.c[d] = g.c[d]; a.p[d] = g.p[d];
    }
    a.mask = g.mask; a.cc = g.cc; a.ctxsum = g.ctxsum; a.bar = g.bar;
    a.dbg = reinterpret_cast<long long*>(g.bar + 16);      // scratch ints after the two counters (debug phase timers)
    a.Tx = g.Tx; a.n = g.n; a.D = g.D; a.upc = pl.upc_f; a.P = pl.P_f; a.BP = pl.BP; a.NS = pl.ns_fwd;
    NATS_CUDA_OK(memset_async(st, g.bar, 0, 2 * sizeof(unsigned)));
    ProfScope ps(st, K_ENC_PERSIST_FWD, 2.0 * 2 * g.Tx * (double)g.n * 3.0 * g.D * g.D, 4.0 * 2 * 3.0 * g.D * g.D);
    dim3 grid(pl.P_f, 2);
    if (pl.nv_fwd <= 4) enc_persist_fwd_kernel<4><<<grid, pl.threads_fwd, pl.smem_fwd, st>>>(a);
    else enc_persist_fwd_kernel<8><<<grid, pl.threads_fwd, pl.smem_fwd, st>>>(a);
    NATS_LAUNCH_OK();
    return 0;
}

int enc_persistent_bwd(const nats_ctx* ctx, cudaStream_t st, const EncPersistBwdArgs& g) {
    const PersistPlan pl = plan(ctx, g.n, g.D);
    NATS_REQUIRE(pl.ok, "persistent encoder not applicable");
    EncPBwd a;
    memset(&a, 0, sizeof(a));
    for (int d = 0; d < 2; ++d) {
        a.Ucat[d] = g.Ucat[d];
        a.r[d] = g.r[d]; a.u[d] = g.u[d]; a.c[d] = g.c[d]; a.p[d] = g.p[d];
        a.dG[d] = g.dG[d]; a.dGx[d] = g.dGx[d];
    }
    a.dcc = g.dcc; a.mean_grad = g.mean_grad; a.coef = g.coef; a.mask = g.mask; a.cc = g.cc; a.bar = g.bar;
    a.Tx = g.Tx; a.n = g.n; a.D = g.D; a.upc = pl.upc_b; a.P = pl.P_b; a.BP = pl.BP; a.NS = pl.ns_bwd;
    NATS_CUDA_OK(memset_async(st, g.bar, 0, 2 * sizeof(unsigned)));
    ProfScope ps(st, K_ENC_PERSIST_BWD, 2.0 * 2 * g.Tx * (double)g.n * 3.0 * g.D * g.D, 4.0 * 2 * 3.0 * g.D * g.D);
    dim3 grid(pl.P_b, 2);
    if (pl.nv_bwd <= 4) enc_persist_bwd_kernel<4><<<grid, pl.threads_bwd, pl.smem_bwd, st>>>(a);
    else enc_persist_bwd_kernel<8><<<grid, pl.threads_bwd, pl.smem_bwd, st>>>(a);
    NATS_LAUNCH_OK();
    return 0;
}

}  // namespace nats
