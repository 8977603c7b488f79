// gates.cuh -- the GRU gate arithmetic of one (sample, unit) element (nats.py:336-356, 505-518, 551-565) for the
// stand-alone gate kernels (ops_elem.cu).
#pragma once
#include "ops.cuh"

namespace nats {

// Sum of split-K slabs in ascending order with the loads of four slabs in flight at once (a runtime-trip-count loop of
// load+add pairs would serialise one L2 round trip per slab: the in-order issue stalls on each add).
__device__ __forceinline__ void load4(const float* __restrict__ p, long long stride, int s0, int n, float (&v)[4]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = (s0 + k < n) ? p[(long long)(s0 + k) * stride] : 0.f;
}
__device__ __forceinline__ float add4(float acc, const float (&v)[4], int s0, int n) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (s0 + k < n) acc += v[k];
    return acc;
}

template <int MODE>
__device__ __forceinline__ void gru_gate_fwd_elem(const GateFwd& a, int b, int j, int D, int idx) {
    const long long row3 = (long long)b * 3 * D;
    // every load of this element is issued before the first dependent add: one L2 round trip instead of one per slab
    const float hp = a.h_prev ? a.h_prev[(long long)b * a.ld_hprev + j] : 0.f;
    const float m = a.mask ? a.mask[b] : 1.f;
    const float cs_old = a.ctxsum ? a.ctxsum[(long long)b * a.ld_ctxsum + j] : 0.f;   // read with the other inputs, not at the tail
    float x0 = 0.f, x1 = 0.f, x2 = 0.f;
    if (MODE == 0) {
        const float* x = a.xproj + row3;
        x0 = x[j]; x1 = x[D + j]; x2 = x[2 * D + j];
    } else {
        x0 = __ldg(a.bias + j); x1 = __ldg(a.bias + D + j); x2 = __ldg(a.bias + 2 * D + j);
    }
    float gr = 0.f, gu = 0.f, pp = 0.f;
    const int nsplit = a.nsplit;
    const float* ps = a.part + row3 + j;
    for (int s0 = 0; s0 < nsplit; s0 += 4) {
        float v0[4], v1[4], v2[4];
        load4(ps, a.part_stride, s0, nsplit, v0);
        load4(ps + D, a.part_stride, s0, nsplit, v1);
        load4(ps + 2 * D, a.part_stride, s0, nsplit, v2);
        gr = add4(gr, v0, s0, nsplit); gu = add4(gu, v1, s0, nsplit); pp = add4(pp, v2, s0, nsplit);
    }
    float xc;
    if (MODE == 0) {
        gr += x0; gu += x1; xc = x2;
    } else {
        float qr = 0.f, qu = 0.f, qc = 0.f;
        const int nsplit2 = a.nsplit2;
        const float* qs = a.part2 + row3 + j;
        for (int s0 = 0; s0 < nsplit2; s0 += 4) {
            float v0[4], v1[4], v2[4];
            load4(qs, a.part2_stride, s0, nsplit2, v0);
            load4(qs + D, a.part2_stride, s0, nsplit2, v1);
            load4(qs + 2 * D, a.part2_stride, s0, nsplit2, v2);
            qr = add4(qr, v0, s0, nsplit2); qu = add4(qu, v1, s0, nsplit2); qc = add4(qc, v2, s0, nsplit2);
        }
        gr += x0 + qr;
        gu += x1 + qu;
        pp += x2;
        xc = qc;
    }
    const float r = sigmoidf_(gr), u = sigmoidf_(gu);
    const float c = tanhf(pp * r + xc);
    const float hn = u * hp + (1.f - u) * c;
    const float h = m * hn + (1.f - m) * hp;
    a.h_out[(long long)b * a.ld_hout + j] = h;
    if (a.r) {
        a.r[idx] = r; a.u[idx] = u; a.c[idx] = c; a.p[idx] = pp;
    }
    if (a.ctxsum) a.ctxsum[(long long)b * a.ld_ctxsum + j] = cs_old + m * h;
}

}  // namespace nats
