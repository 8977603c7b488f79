// enc_persistent.cu -- the bidirectional GRU encoder recurrence (nats.py:336-372, both directions) as ONE persistent,
// weight-stationary kernel per pass.
//
// Why: a recurrent step is a [B x D] x [D x 3D] product with B = 32: launched per step it costs ~16 us (split-K tensor-core
// product + gate kernel), most of it fixed overhead and re-streaming the 12 MB weight matrix from L2.  Here every CTA
// owns `upc` hidden units of one direction for the whole sequence:
//   * its 3*upc columns of [U|Ux] (forward) / its upc rows of [U|Ux] (backward) are loaded ONCE into shared memory
//     (192 KB at D = 1000: 72 + 72 CTAs forward, 63 + 63 backward, one per SM);
//   * the K extent is cut into 8-deep slices dealt round-robin to the 16 warps.  A warp stages ITS slice of h_{t-1}
//     (forward) / dG_{t+1} (backward) from L2 into a private 1 KB shared buffer (two 16-byte loads per lane, the next
//     slice in flight during the arithmetic, only __syncwarp in the loop) and accumulates a 32 x ncol partial product
//     with packed fp32 FMAs (FFMA2, exact fp32): lanes = 4 batch groups x 8 column groups, 8x6 register tile (forward)
//     or 2 slices x 4 x 4, 8x4 tile (backward); every shared load is conflict-free or a broadcast, addresses are
//     immediates of two base registers;
//   * the 16 partials are added in a FIXED order through shared-memory slabs (deterministic), and since the CTA sees
//     the full K extent of its units the GRU gate arithmetic (and its reverse) is finished in place -- no split-K slabs
//     in global memory, no second kernel -- and h_t / dG_t are published for the other CTAs;
//   * one arrive/spin barrier per direction and step (monotonic counter in L2) orders the steps.
// The grid must be co-resident (2*P CTAs <= #SMs, one CTA per SM by shared-memory footprint); a spinning CTA traps
// after ~2 s instead of hanging the GPU.
#include "ops.cuh"

namespace nats {

namespace {

constexpr int kWarps = 16;
constexpr int kThreads = kWarps * 32;
constexpr int kChunk = 8 * kWarps;      // k rows consumed per round of all warps (8 per warp)
constexpr int kBP = 32;                 // batch rows of the CTA tile (n <= 32)
constexpr int kHS = 36;                 // row stride of a staging buffer: 32 batch columns + 4 (the two slices of a warp hit disjoint banks)
constexpr int kStageWarp = 8 * kHS;     // floats per warp-private staging buffer
constexpr int kStageFloats = kWarps * kStageWarp;

struct EncPFwd {
    const float* Ucat[2];       // [D,3D]
    const float* xproj[2];      // [Tx*n,3D] by source position
    const float* mask;          // [Tx,n] or NULL
    float* cc;                  // [Tx,n,2D]
    float* r[2]; float* u[2]; float* c[2]; float* p[2];   // [Tx*n,D] by position, or NULL
    float* ctxsum;              // [n,2D]
    unsigned* bar;              // [2] zero-initialised
    long long* dbg;             // optional [8] phase cycle counters of CTA (0,0)
    int Tx, n, D, upc, P, NS;   // NS = slabs of the cross-warp reduction
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
    long long* dbg;
    int Tx, n, D, upc, P, NS;
};

__device__ __forceinline__ float4 ldcg4(const float* p) {
    float4 r;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
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
            if (clock64() - t0 > 4000000000LL) __trap();        // ~2 s: never hang the device
        } while (true);
        __threadfence();
    }
    __syncthreads();
}

// Packed fp32 FMA (sm_100 FFMA2): two exact fp32 FMAs per lane and instruction.
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void ffma2(u64& d, u64 a, u64 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b)); }

// 8 x TC register tile held as packed pairs: for row pair i and column pair j
//   X[i][j] = (acc[2i][2j],   acc[2i+1][2j+1])   += (h[2i], h[2i+1]) * (w[2j],   w[2j+1])
//   Y[i][j] = (acc[2i][2j+1], acc[2i+1][2j])     += (h[2i], h[2i+1]) * (w[2j+1], w[2j])
// so the h pairs and the w pairs come straight out of the vector shared loads; the swapped w pair is an operand
// modifier of FFMA2 (no instruction).
template <int TC>
struct Frag {
    u64 X[4][TC / 2], Y[4][TC / 2];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < TC / 2; ++j) { X[i][j] = 0ull; Y[i][j] = 0ull; }
    }
    __device__ __forceinline__ void get(float (&acc)[8][TC]) const {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < TC / 2; ++j) {
                unpack2(X[i][j], acc[2 * i][2 * j], acc[2 * i + 1][2 * j + 1]);
                unpack2(Y[i][j], acc[2 * i][2 * j + 1], acc[2 * i + 1][2 * j]);
            }
    }
};

// Partial product of this warp's K slices:  f += src[0..nrows)[k] (x) Wt[k][cols of this lane]   over k in the warp's slices.
//   src: row-major [nrows, ld] in global memory (written by other CTAs in the previous step: read through L2)
//   Wt : shared, [K][NCOL], NCOL = (8/SL)*TC
//   hb : this warp's private staging buffer [8][kHS]
// lane layout: slice sl = lane / (32/SL) (takes rows SL*t + sl of the staged 8), then bg (4 groups of 8 batch rows)
// x cg (8/SL groups of TC columns).
template <int TC, int SL, int PF>
__device__ __forceinline__ void kslice_product(const float* __restrict__ src, long long ld, int nrows, int K,
                                               const float* __restrict__ Wt, float* __restrict__ hb, int warp, int lane,
                                               int rot, Frag<TC>& f) {
    constexpr int NCG = 8 / SL, NCOL = NCG * TC, KT = 8 / SL, LPS = 32 / SL;
    const int sl = lane / LPS, l2 = lane % LPS, bg = l2 / NCG, cg = l2 % NCG;
    const float* hrd = hb + sl * kHS + bg * 8;
    const float* wrd = Wt + (size_t)(8 * warp + sl) * NCOL + cg * TC;
    const float* grow = src + (long long)lane * ld + 8 * warp;      // lane = batch row of the staged slice
    const bool rowok = lane < nrows;
    const int nch = (K + kChunk - 1) / kChunk;
    // PF slices in flight per warp (registers): the L2 round trip under load is > 1000 cycles
    float4 v[PF][2];
#pragma unroll
    for (int p = 0; p < PF; ++p) {
        v[p][0] = make_float4(0.f, 0.f, 0.f, 0.f); v[p][1] = v[p][0];
        const int cc = (p + rot) % nch;      // CTAs walk the K extent from different starting chunks: no L2 hot spot
        if (rowok && p < nch && cc * kChunk + 8 * warp < K) { v[p][0] = ldcg4(grow + cc * kChunk); v[p][1] = ldcg4(grow + cc * kChunk + 4); }
    }
    for (int c0 = 0; c0 < nch; c0 += PF) {
#pragma unroll
        for (int p = 0; p < PF; ++p) {
            const int c = c0 + p;
            if (c >= nch) break;
            const int cc = (c + rot) % nch;
            const bool act = cc * kChunk + 8 * warp < K;            // warp-uniform (K % 8 == 0)
            __syncwarp();
            hb[0 * kHS + lane] = v[p][0].x; hb[1 * kHS + lane] = v[p][0].y; hb[2 * kHS + lane] = v[p][0].z; hb[3 * kHS + lane] = v[p][0].w;
            hb[4 * kHS + lane] = v[p][1].x; hb[5 * kHS + lane] = v[p][1].y; hb[6 * kHS + lane] = v[p][1].z; hb[7 * kHS + lane] = v[p][1].w;
            __syncwarp();
            const int cn = (c + PF + rot) % nch;
            if (rowok && c + PF < nch && cn * kChunk + 8 * warp < K) {
                v[p][0] = ldcg4(grow + cn * kChunk);
                v[p][1] = ldcg4(grow + cn * kChunk + 4);
            }
            if (act) {
                const float* w = wrd + (size_t)cc * kChunk * NCOL;
#pragma unroll
                for (int t = 0; t < KT; ++t) {
                    const float4 h0 = *reinterpret_cast<const float4*>(hrd + SL * t * kHS);
                    const float4 h1 = *reinterpret_cast<const float4*>(hrd + SL * t * kHS + 4);
                    const u64 hp[4] = {pack2(h0.x, h0.y), pack2(h0.z, h0.w), pack2(h1.x, h1.y), pack2(h1.z, h1.w)};
                    float wv[TC];
                    if (TC == 4) {
                        const float4 q = *reinterpret_cast<const float4*>(w + SL * t * NCOL);
                        wv[0] = q.x; wv[1] = q.y; wv[2] = q.z; wv[3] = q.w;
                    } else {
#pragma unroll
                        for (int j = 0; j < TC / 2; ++j) {
                            const float2 q = *reinterpret_cast<const float2*>(w + SL * t * NCOL + 2 * j);
                            wv[2 * j] = q.x; wv[2 * j + 1] = q.y;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < TC / 2; ++j) {
                        const u64 wp = pack2(wv[2 * j], wv[2 * j + 1]);
                        const u64 ws = pack2(wv[2 * j + 1], wv[2 * j]);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            ffma2(f.X[i][j], hp[i], wp);
                            ffma2(f.Y[i][j], hp[i], ws);
                        }
                    }
                }
            }
        }
    }
}

// Slabs are kept in FRAGMENT order: element (row b, column col) of the 32 x NCOL tile lives at
//   ((i * LPS + bg*NCG + cg) * TC + j),  b = 8*bg + i,  col = cg*TC + j          (conflict-free vector stores)
template <int TC, int SL>
__device__ __forceinline__ int slab_index(int b, int col) {
    constexpr int NCG = 8 / SL, LPS = 32 / SL;
    const int bg = b >> 3, i = b & 7, cg = col / TC, j = col - cg * TC;
    return (i * LPS + bg * NCG + cg) * TC + j;
}

// Cross-warp reduction in a fixed order: warp w adds its partial into slab (w % NS) in round (w / NS).
template <int TC, int SL>
__device__ __forceinline__ void reduce_to_slabs(float* __restrict__ slabs, int NS, int warp, int lane, const Frag<TC>& f) {
    constexpr int LPS = 32 / SL, SLAB = 8 * LPS * TC;
    float acc[8][TC];
    f.get(acc);
    if (SL == 2) {          // the two slices of a warp hold the same tile positions in lanes l and l + 16
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < TC; ++j) acc[i][j] += __shfl_xor_sync(0xffffffffu, acc[i][j], 16);
    }
    const int rounds = (kWarps + NS - 1) / NS;
    for (int r = 0; r < rounds; ++r) {
        if (warp / NS == r && lane < LPS) {
            float* sl = slabs + (size_t)(warp % NS) * SLAB;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float* p = sl + (i * LPS + lane) * TC;
                if (TC == 4) {
                    float4 a = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
                    if (r > 0) {
                        const float4 o = *reinterpret_cast<const float4*>(p);
                        a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
                    }
                    *reinterpret_cast<float4*>(p) = a;
                } else {
#pragma unroll
                    for (int j = 0; j < TC / 2; ++j) {
                        float2 a = make_float2(acc[i][2 * j], acc[i][2 * j + 1]);
                        if (r > 0) {
                            const float2 o = *reinterpret_cast<const float2*>(p + 2 * j);
                            a.x += o.x; a.y += o.y;
                        }
                        *reinterpret_cast<float2*>(p + 2 * j) = a;
                    }
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ forward
constexpr int kFwdTC = 6, kFwdSL = 1, kFwdCols = (8 / kFwdSL) * kFwdTC;        // 48 columns = 3 gates x <= 16 units
constexpr int kFwdSlab = 8 * (32 / kFwdSL) * kFwdTC;
constexpr int kFwdPF = 2;

__global__ void __launch_bounds__(kThreads, 1) enc_persist_fwd_kernel(const __grid_constant__ EncPFwd a) {
    extern __shared__ __align__(16) float sm[];
    const int dir = blockIdx.y, D = a.D, n = a.n, C = 2 * D, D3 = 3 * D, upc = a.upc;
    const int j0 = blockIdx.x * upc;
    const int nu = min(upc, D - j0);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* Wt = sm;                                   // [D][48]   column = gate*upc + unit
    float* buf = Wt + (size_t)D * kFwdCols;           // staging buffers [16][8][kHS]  /  reduction slabs [NS][fragment order]

    {   // weights of this CTA's units, once
        const float* U = a.Ucat[dir];
        for (int idx = tid; idx < D * kFwdCols; idx += kThreads) {
            const int k = idx / kFwdCols, col = idx - k * kFwdCols;
            const int g = col / upc, uu = col - g * upc;
            Wt[idx] = (g < 3 && uu < nu) ? __ldg(U + (long long)k * D3 + g * D + j0 + uu) : 0.f;
        }
    }
    // this thread's (batch, unit) pair in the gate phase: the same one at every step -> carries live in registers
    const bool gate_thread = tid < n * nu;
    const int gb = gate_thread ? tid / nu : 0, guu = gate_thread ? tid - gb * nu : 0;
    const int gj = j0 + guu;
    float h_carry = 0.f, cs_carry = 0.f;
    const int nsl = min(kWarps, a.NS);
    __syncthreads();

    long long t_comp = 0, t_red = 0, t_gate = 0, t_bar = 0;
    for (int s = 0; s < a.Tx; ++s) {
        const int pos = dir == 0 ? s : a.Tx - 1 - s;
        const int prev = dir == 0 ? pos - 1 : pos + 1;
        const long long c0 = clock64();
        // gate inputs of this step (independent of h_{t-1}): requested now, consumed after the product
        float gx0 = 0.f, gx1 = 0.f, gx2 = 0.f, gm = 1.f;
        if (gate_thread) {
            const float* x = a.xproj[dir] + ((long long)pos * n + gb) * D3 + gj;
            gx0 = __ldg(x); gx1 = __ldg(x + D); gx2 = __ldg(x + 2 * D);
            if (a.mask) gm = __ldg(a.mask + (long long)pos * n + gb);
        }
        long long c1 = c0;
        if (s > 0) {
            Frag<kFwdTC> f;
            f.clear();
            kslice_product<kFwdTC, kFwdSL, kFwdPF>(a.cc + (long long)prev * n * C + dir * D, C, n, D, Wt, buf + warp * kStageWarp, warp, lane, (int)blockIdx.x, f);
            __syncthreads();                          // every warp is done with its staging buffer: the slabs alias them
            c1 = clock64();
            reduce_to_slabs<kFwdTC, kFwdSL>(buf, a.NS, warp, lane, f);
        }
        const long long c2 = clock64();
        // gates for the (b, unit) pair of this thread (nats.py:341-354)
        if (gate_thread) {
            float gr = gx0, gu = gx1, pp = 0.f;
            if (s > 0) {
                const int i0 = slab_index<kFwdTC, kFwdSL>(gb, guu), i1 = slab_index<kFwdTC, kFwdSL>(gb, upc + guu),
                          i2 = slab_index<kFwdTC, kFwdSL>(gb, 2 * upc + guu);
                for (int sl = 0; sl < nsl; ++sl) {
                    const float* pr = buf + (size_t)sl * kFwdSlab;
                    gr += pr[i0]; gu += pr[i1]; pp += pr[i2];
                }
            }
            const float r = sigmoidf_(gr), uz = sigmoidf_(gu);
            const float c = tanhf(pp * r + gx2);
            const float hp = h_carry;
            const float hn = uz * hp + (1.f - uz) * c;
            const float h = gm * hn + (1.f - gm) * hp;
            h_carry = h;
            cs_carry += gm * h;
            a.cc[((long long)pos * n + gb) * C + dir * D + gj] = h;
            if (a.r[dir]) {
                const long long o = ((long long)pos * n + gb) * D + gj;
                a.r[dir][o] = r; a.u[dir][o] = uz; a.c[dir][o] = c; a.p[dir][o] = pp;
            }
        }
        const long long c3 = clock64();
        if (s + 1 < a.Tx) dir_barrier(a.bar + dir, (unsigned)a.P * (unsigned)(s + 1));
        const long long c4 = clock64();
        t_comp += c1 - c0; t_red += c2 - c1; t_gate += c3 - c2; t_bar += c4 - c3;
    }
    if (gate_thread) a.ctxsum[(long long)gb * C + dir * D + gj] += cs_carry;
    if (a.dbg && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) {
        a.dbg[0] = t_comp; a.dbg[1] = t_red; a.dbg[2] = t_gate; a.dbg[3] = t_bar;
    }
}

// ------------------------------------------------------------------------------------------------ backward
// CTA owns units j0..j0+upc-1 (upc <= 16): out[b][u] = sum_k dG_{t+1}[b,k] * Ucat[j0+u, k]   (K = 3D), then the gate
// backward of step t for its units; the elementwise carry of d h stays in a register across steps.
constexpr int kBwdTC = 4, kBwdSL = 2, kBwdCols = (8 / kBwdSL) * kBwdTC;        // 16 columns
constexpr int kBwdSlab = 8 * (32 / kBwdSL) * kBwdTC;
constexpr int kBwdPF = 4;

__global__ void __launch_bounds__(kThreads, 1) enc_persist_bwd_kernel(const __grid_constant__ EncPBwd a) {
    extern __shared__ __align__(16) float sm[];
    const int dir = blockIdx.y, D = a.D, n = a.n, C = 2 * D, D3 = 3 * D, upc = a.upc;
    const int j0 = blockIdx.x * upc;
    const int nu = min(upc, D - j0);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* Wt = sm;                                   // [3D][16]   Wt[k][u] = Ucat[j0+u, k]
    float* buf = Wt + (size_t)D3 * kBwdCols;

    {
        const float* U = a.Ucat[dir];
        for (int idx = tid; idx < D3 * kBwdCols; idx += kThreads) {
            const int uu = idx / D3, k = idx - uu * D3;      // coalesced along k
            Wt[(size_t)k * kBwdCols + uu] = (uu < nu) ? __ldg(U + (long long)(j0 + uu) * D3 + k) : 0.f;
        }
    }
    const bool gate_thread = tid < n * nu;
    const int gb = gate_thread ? tid / nu : 0, guu = gate_thread ? tid - gb * nu : 0;
    const int gj = j0 + guu;
    const int gi = slab_index<kBwdTC, kBwdSL>(gb, guu);
    const int nsl = min(kWarps, a.NS);
    float carry = 0.f;                                // elementwise part of d h_{t-1}
    const float mg = (gate_thread && a.mean_grad) ? __ldg(a.coef + gb) * __ldg(a.mean_grad + (long long)gb * C + dir * D + gj) : 0.f;
    __syncthreads();

    long long t_comp = 0, t_red = 0, t_gate = 0, t_bar = 0;
    for (int s = a.Tx - 1; s >= 0; --s) {
        const int pos = dir == 0 ? s : a.Tx - 1 - s;
        const int prev = dir == 0 ? pos - 1 : pos + 1;            // position of h_{t-1}
        const int nextp = dir == 0 ? pos + 1 : pos - 1;           // position processed by step s+1
        const long long c0 = clock64();
        // gate-backward inputs of this step: requested before the product, consumed after it
        float g_dcc = 0.f, g_m = 1.f, g_r = 0.f, g_u = 0.f, g_c = 0.f, g_p = 0.f, g_hp = 0.f;
        if (gate_thread) {
            const long long o = ((long long)pos * n + gb) * D + gj;
            g_dcc = __ldg(a.dcc + ((long long)pos * n + gb) * C + dir * D + gj);
            g_m = __ldg(a.mask + (long long)pos * n + gb);
            g_r = __ldg(a.r[dir] + o); g_u = __ldg(a.u[dir] + o); g_c = __ldg(a.c[dir] + o); g_p = __ldg(a.p[dir] + o);
            g_hp = s > 0 ? __ldg(a.cc + ((long long)prev * n + gb) * C + dir * D + gj) : 0.f;
        }
        long long c1 = c0;
        if (s < a.Tx - 1) {
            Frag<kBwdTC> f;
            f.clear();
            kslice_product<kBwdTC, kBwdSL, kBwdPF>(a.dG[dir] + (long long)nextp * n * D3, D3, n, D3, Wt, buf + warp * kStageWarp, warp, lane, (int)blockIdx.x, f);
            __syncthreads();
            c1 = clock64();
            reduce_to_slabs<kBwdTC, kBwdSL>(buf, a.NS, warp, lane, f);
        }
        const long long c2 = clock64();
        if (gate_thread) {
            const float m = g_m;
            float dh = g_dcc + carry + m * mg;
            if (s < a.Tx - 1)
                for (int sl = 0; sl < nsl; ++sl) dh += buf[(size_t)sl * kBwdSlab + gi];
            const float r = g_r, uz = g_u, c = g_c, p = g_p, hp = g_hp;
            const float dhn = m * dh;
            const float du = dhn * (hp - c);
            const float dc = dhn * (1.f - uz);
            const float dpc = dc * (1.f - c * c);
            const float dp = dpc * r;
            const float dr = dpc * p;
            const float dgr = dr * r * (1.f - r);
            const float dgu = du * uz * (1.f - uz);
            const long long row3 = ((long long)pos * n + gb) * D3;
            a.dG[dir][row3 + gj] = dgr; a.dG[dir][row3 + D + gj] = dgu; a.dG[dir][row3 + 2 * D + gj] = dp;
            a.dGx[dir][row3 + gj] = dgr; a.dGx[dir][row3 + D + gj] = dgu; a.dGx[dir][row3 + 2 * D + gj] = dpc;
            carry = (1.f - m) * dh + dhn * uz;
        }
        const long long c3 = clock64();
        if (s > 0) dir_barrier(a.bar + dir, (unsigned)a.P * (unsigned)(a.Tx - s));
        const long long c4 = clock64();
        t_comp += c1 - c0; t_red += c2 - c1; t_gate += c3 - c2; t_bar += c4 - c3;
    }
    if (a.dbg && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) {
        a.dbg[4] = t_comp; a.dbg[5] = t_red; a.dbg[6] = t_gate; a.dbg[7] = t_bar;
    }
}

struct PersistPlan {
    int upc_f, P_f, ns_fwd; size_t smem_fwd;
    int upc_b, P_b, ns_bwd; size_t smem_bwd;
    bool ok;
};

PersistPlan plan(const nats_ctx* ctx, int n, int D) {
    PersistPlan pl;
    memset(&pl, 0, sizeof(pl));
    if (n < 1 || n > kBP || D < 8 || (D % 8) != 0) return pl;
    const int per_dir = ctx->num_sms / 2;
    if (per_dir < 1) return pl;
    const size_t lim = (size_t)ctx->max_smem_optin - 1024;
    {   // forward: 3*upc <= 48 columns per CTA
        pl.upc_f = (D + per_dir - 1) / per_dir;
        if (pl.upc_f < 8) pl.upc_f = D < 8 ? D : 8;
        if (3 * pl.upc_f > kFwdCols) return pl;
        pl.P_f = (D + pl.upc_f - 1) / pl.upc_f;
        const size_t w = (size_t)D * kFwdCols;
        if (w * sizeof(float) + kStageFloats * sizeof(float) > lim) return pl;
        size_t room = lim / sizeof(float) - w;
        if (room > 2u * kStageFloats) room = 2u * kStageFloats;           // 32 KB is plenty
        pl.ns_fwd = (int)(room / kFwdSlab);
        if (pl.ns_fwd > kWarps) pl.ns_fwd = kWarps;
        if (pl.ns_fwd < 1) return pl;
        const size_t slabs = (size_t)pl.ns_fwd * kFwdSlab;
        pl.smem_fwd = (w + (slabs > (size_t)kStageFloats ? slabs : (size_t)kStageFloats)) * sizeof(float);
    }
    {   // backward: upc <= 16 columns per CTA
        pl.upc_b = kBwdCols;
        pl.P_b = (D + pl.upc_b - 1) / pl.upc_b;
        const size_t w = (size_t)3 * D * kBwdCols;
        if (w * sizeof(float) + kStageFloats * sizeof(float) > lim) return pl;
        size_t room = lim / sizeof(float) - w;
        if (room > 2u * kStageFloats) room = 2u * kStageFloats;
        pl.ns_bwd = (int)(room / kBwdSlab);
        if (pl.ns_bwd > kWarps) pl.ns_bwd = kWarps;
        if (pl.ns_bwd < 1) return pl;
        const size_t slabs = (size_t)pl.ns_bwd * kBwdSlab;
        pl.smem_bwd = (w + (slabs > (size_t)kStageFloats ? slabs : (size_t)kStageFloats)) * sizeof(float);
    }
    pl.ok = pl.smem_fwd <= lim && pl.smem_bwd <= lim && 2 * pl.P_f <= ctx->num_sms && 2 * pl.P_b <= ctx->num_sms &&
            n * pl.upc_f <= kThreads && n * pl.upc_b <= kThreads;
    return pl;
}

static int g_persist = 0;

}  // namespace

void enc_persistent_enable(int on) { g_persist = on; }

int enc_persistent_setup(const nats_ctx* ctx) {
    const int lim = ctx->max_smem_optin - 1024;
    NATS_CUDA_OK(cudaFuncSetAttribute(enc_persist_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    NATS_CUDA_OK(cudaFuncSetAttribute(enc_persist_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    return 0;
}

// g_persist: 0 off, 1 both passes, 2 forward only, 3 backward only
bool enc_persistent_eligible(const nats_ctx* ctx, int n, int D, int pass) {
    if (!(g_persist == 1 || (g_persist == 2 && pass == 0) || (g_persist == 3 && pass == 1))) return false;
    return plan(ctx, n, D).ok;
}

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
    a.Tx = g.Tx; a.n = g.n; a.D = g.D; a.upc = pl.upc_f; a.P = pl.P_f; a.NS = pl.ns_fwd;
    NATS_CUDA_OK(memset_async(st, g.bar, 0, 2 * sizeof(unsigned)));
    ProfScope ps(st, K_ENC_PERSIST_FWD, 2.0 * 2 * g.Tx * (double)g.n * 3.0 * g.D * g.D, 4.0 * 2 * 3.0 * g.D * g.D);
    dim3 grid(pl.P_f, 2);
    enc_persist_fwd_kernel<<<grid, kThreads, pl.smem_fwd, st>>>(a);
    NATS_LAUNCH_OK();
    return 0;
}

int enc_persistent_bwd(const nats_ctx* ctx, cudaStream_t st, const EncPersistBwdArgs& g) {
    const PersistPlan pl = plan(ctx, g.n, g.D);
    NATS_REQUIRE(pl.ok, "persistent encoder not applicable");
    NATS_REQUIRE((reinterpret_cast<uintptr_t>(g.dG[0]) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.dG[1]) & 15) == 0, "dG alignment");
    EncPBwd a;
    memset(&a, 0, sizeof(a));
    for (int d = 0; d < 2; ++d) {
        a.Ucat[d] = g.Ucat[d];
        a.r[d] = g.r[d]; a.u[d] = g.u[d]; a.c[d] = g.c[d]; a.p[d] = g.p[d];
        a.dG[d] = g.dG[d]; a.dGx[d] = g.dGx[d];
    }
    a.dcc = g.dcc; a.mean_grad = g.mean_grad; a.coef = g.coef; a.mask = g.mask; a.cc = g.cc; a.bar = g.bar;
    a.dbg = reinterpret_cast<long long*>(g.bar + 16);
    a.Tx = g.Tx; a.n = g.n; a.D = g.D; a.upc = pl.upc_b; a.P = pl.P_b; a.NS = pl.ns_bwd;
    NATS_CUDA_OK(memset_async(st, g.bar, 0, 2 * sizeof(unsigned)));
    ProfScope ps(st, K_ENC_PERSIST_BWD, 2.0 * 2 * g.Tx * (double)g.n * 3.0 * g.D * g.D, 4.0 * 2 * 3.0 * g.D * g.D);
    dim3 grid(pl.P_b, 2);
    enc_persist_bwd_kernel<<<grid, kThreads, pl.smem_bwd, st>>>(a);
    NATS_LAUNCH_OK();
    return 0;
}

}  // namespace nats
