// enc_tc_test.cu -- standalone check + timing of the persistent tcgen05 encoder kernels (enc_tc.cu) against a double
// precision CPU recurrence.  build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o tools/micro/enc_tc_test tools/micro/enc_tc_test.cu -lcuda
// run: enc_tc_test [D] [n] [Tx_check] [Tx_time]
#include <cstdio>
#include <cstdlib>
#include <cstdarg>
#include <cmath>
#include <vector>
#include <random>
#include "../../nats_b200/csrc/enc_tc.cu"

namespace nats {
static char g_err[1024];
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap); }
bool prof_enabled() { return false; }
void prof_begin(cudaStream_t, int, double, double) {}
void prof_end(cudaStream_t) {}
int pdl_enabled() { return 0; }
}
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

using namespace nats;

static void cpu_fwd(int Tx, int n, int D, const std::vector<float>& U, const std::vector<float>& xp, const std::vector<float>& mask,
                    int dir, std::vector<double>& cc, std::vector<double>& sv) {
    std::vector<double> h(n * D, 0.0), hn(n * D);
    const int D3 = 3 * D;
    std::vector<double> pre(D3);
    for (int s = 0; s < Tx; ++s) {
        const int pos = dir == 0 ? s : Tx - 1 - s;
        for (int b = 0; b < n; ++b) {
            for (int j = 0; j < D3; ++j) pre[j] = 0.0;
            for (int k = 0; k < D; ++k) {
                const double hv = h[b * D + k];
                if (hv == 0.0) continue;
                const float* w = &U[(size_t)k * D3];
                for (int j = 0; j < D3; ++j) pre[j] += hv * (double)w[j];
            }
            const float* x = &xp[((size_t)pos * n + b) * D3];
            const double m = mask[pos * n + b];
            for (int j = 0; j < D; ++j) {
                const double r = 1.0 / (1.0 + exp(-(pre[j] + x[j]))), u = 1.0 / (1.0 + exp(-(pre[D + j] + x[D + j])));
                const double c = tanh(pre[2 * D + j] * r + x[2 * D + j]);
                const double hv = u * h[b * D + j] + (1.0 - u) * c;
                hn[b * D + j] = m * hv + (1.0 - m) * h[b * D + j];
                const size_t o = ((size_t)pos * n + b) * D + j;
                sv[o * 4] = r; sv[o * 4 + 1] = u; sv[o * 4 + 2] = c; sv[o * 4 + 3] = pre[2 * D + j];
                cc[((size_t)pos * n + b) * 2 * D + dir * D + j] = hn[b * D + j];
            }
        }
        h = hn;
    }
}

// reverse mode of cpu_fwd for one direction: dG / dGx per position (double)
static void cpu_bwd(int Tx, int n, int D, const std::vector<float>& U, const std::vector<float>& mask, int dir,
                    const std::vector<double>& cc, const std::vector<double>& sv, const std::vector<float>& dcc,
                    const std::vector<float>& mean, const std::vector<float>& coef, std::vector<double>& dG, std::vector<double>& dGx) {
    const int D3 = 3 * D, C = 2 * D;
    std::vector<double> carry((size_t)n * D, 0.0), prod((size_t)n * D, 0.0);
    for (int q = 0; q < Tx; ++q) {
        const int pos = dir == 0 ? Tx - 1 - q : q;
        const int prev = dir == 0 ? pos - 1 : pos + 1, next = dir == 0 ? pos + 1 : pos - 1;
        if (q > 0) {
            for (int b = 0; b < n; ++b)
                for (int d = 0; d < D; ++d) {
                    double acc = 0.0;
                    const double* g = &dG[((size_t)next * n + b) * D3];
                    const float* w = &U[(size_t)d * D3];
                    for (int k = 0; k < D3; ++k) acc += g[k] * (double)w[k];
                    prod[(size_t)b * D + d] = acc;
                }
        }
        for (int b = 0; b < n; ++b)
            for (int d = 0; d < D; ++d) {
                const size_t o = ((size_t)pos * n + b) * D + d;
                const double m = mask[pos * n + b];
                const double r = sv[o * 4], u = sv[o * 4 + 1], c = sv[o * 4 + 2], pp = sv[o * 4 + 3];
                const double hp = (q < Tx - 1) ? cc[((size_t)prev * n + b) * C + dir * D + d] : 0.0;
                double dh = dcc[((size_t)pos * n + b) * C + dir * D + d];
                if (q > 0) dh += carry[(size_t)b * D + d] + prod[(size_t)b * D + d];
                dh += m * (double)coef[b] * (double)mean[(size_t)b * C + dir * D + d];
                const double dhn = m * dh, du = dhn * (hp - c), dc = dhn * (1.0 - u), dpc = dc * (1.0 - c * c);
                const double dp = dpc * r, dr = dpc * pp, dgr = dr * r * (1.0 - r), dgu = du * u * (1.0 - u);
                double* g = &dG[((size_t)pos * n + b) * D3];
                double* gx = &dGx[((size_t)pos * n + b) * D3];
                g[d] = dgr; g[D + d] = dgu; g[2 * D + d] = dp;
                gx[d] = dgr; gx[D + d] = dgu; gx[2 * D + d] = dpc;
                carry[(size_t)b * D + d] = (1.0 - m) * dh + dhn * u;
            }
    }
}

static double maxerr(const std::vector<float>& a, const std::vector<double>& b, double* scale) {
    double e = 0.0, s = 0.0;
    for (size_t i = 0; i < a.size(); ++i) {
        const double dd = fabs((double)a[i] - b[i]);
        if (!(dd <= e)) e = dd;
        if (fabs(b[i]) > s) s = fabs(b[i]);
    }
    if (scale) *scale = s;
    return e;
}

int main(int argc, char** argv) {
    const int D = argc > 1 ? atoi(argv[1]) : 1000, n = argc > 2 ? atoi(argv[2]) : 32;
    const int Tc = argc > 3 ? atoi(argv[3]) : 12, Tt = argc > 4 ? atoi(argv[4]) : 400;
    nats_ctx ctx; memset(&ctx, 0, sizeof(ctx));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    ctx.num_sms = prop.multiProcessorCount; ctx.max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    printf("device %s, %d SMs, smem optin %d\n", prop.name, ctx.num_sms, ctx.max_smem_optin);
    if (enc_tc_setup(&ctx)) { printf("setup failed: %s\n", g_err); return 1; }
    const bool el0 = enc_tc_eligible(&ctx, n, D, 0), el1 = enc_tc_eligible(&ctx, n, D, 1);
    for (int pass = 0; pass < 2; ++pass) {
        const TcPlan p = plan(&ctx, n, D, pass);
        printf("pass %d: ok %d BN %d NT %d S %d dpc %d dps %d Kc %d nkbA %d NS %d smem %zu (static %zu) resident/SM %d\n", pass, (int)p.ok, p.BN, p.NT, p.S,
               p.dpc, p.dps, p.Kc, p.nkbA, p.NS, p.smem, g_static_smem, g_resident);
    }
    if (!el0 || !el1) { printf("not eligible (%d %d)\n", (int)el0, (int)el1); return 2; }
    const int D3 = 3 * D, C = 2 * D, Tmax = Tt > Tc ? Tt : Tc;
    std::mt19937 rng(1234);
    std::normal_distribution<float> nd(0.f, 1.f);
    std::vector<float> U[2], xp[2], mask((size_t)Tmax * n), hdcc((size_t)Tmax * n * C), hmean((size_t)n * C), hcoef(n);
    const float wscale = 1.0f / sqrtf((float)D);
    for (int d = 0; d < 2; ++d) {
        U[d].resize((size_t)D * D3); for (auto& v : U[d]) v = nd(rng) * wscale;
        xp[d].resize((size_t)Tmax * n * D3); for (auto& v : xp[d]) v = nd(rng);
    }
    for (auto& v : hdcc) v = nd(rng);
    for (auto& v : hmean) v = nd(rng);
    for (auto& v : hcoef) v = 0.05f + 0.01f * fabsf(nd(rng));
    for (int t = 0; t < Tmax; ++t) for (int b = 0; b < n; ++b) mask[t * n + b] = (t < Tc - (b % 5)) || t >= Tc ? 1.f : 0.f;
    float *dU[2], *dxp[2], *dmask, *dcc, *dr[2], *du[2], *dc[2], *dp[2], *dctx, *dscr, *ddcc, *dmean, *dcoef, *dGd[2], *dGxd[2];
    unsigned* dbar; unsigned long long* ddbg;
    for (int d = 0; d < 2; ++d) {
        CK(cudaMalloc(&dU[d], U[d].size() * 4)); CK(cudaMemcpy(dU[d], U[d].data(), U[d].size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMalloc(&dxp[d], xp[d].size() * 4)); CK(cudaMemcpy(dxp[d], xp[d].data(), xp[d].size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMalloc(&dr[d], (size_t)Tmax * n * D * 4)); CK(cudaMalloc(&du[d], (size_t)Tmax * n * D * 4));
        CK(cudaMalloc(&dc[d], (size_t)Tmax * n * D * 4)); CK(cudaMalloc(&dp[d], (size_t)Tmax * n * D * 4));
        CK(cudaMalloc(&dGd[d], (size_t)Tmax * n * D3 * 4)); CK(cudaMalloc(&dGxd[d], (size_t)Tmax * n * D3 * 4));
        CK(cudaMemset(dGd[d], 0xff, (size_t)Tmax * n * D3 * 4));
    }
    CK(cudaMalloc(&dmask, mask.size() * 4)); CK(cudaMemcpy(dmask, mask.data(), mask.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&ddcc, hdcc.size() * 4)); CK(cudaMemcpy(ddcc, hdcc.data(), hdcc.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dmean, hmean.size() * 4)); CK(cudaMemcpy(dmean, hmean.data(), hmean.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dcoef, hcoef.size() * 4)); CK(cudaMemcpy(dcoef, hcoef.data(), hcoef.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dcc, (size_t)Tmax * n * C * 4)); CK(cudaMemset(dcc, 0xff, (size_t)Tmax * n * C * 4));   // NaN fill: catches reads of unwritten rows
    CK(cudaMalloc(&dctx, (size_t)n * C * 4));
    const long long scr = enc_tc_scratch_floats(n, D), nbar = enc_tc_counter_ints();
    CK(cudaMalloc(&dscr, scr * 4)); CK(cudaMemset(dscr, 0xff, scr * 4));
    CK(cudaMalloc(&dbar, nbar * 4)); CK(cudaMalloc(&ddbg, 128 * 8)); CK(cudaMemset(ddbg, 0, 128 * 8));
    EncTcFwdArgs a; memset(&a, 0, sizeof(a));
    for (int d = 0; d < 2; ++d) { a.Ucat[d] = dU[d]; a.xproj[d] = dxp[d]; a.r[d] = dr[d]; a.u[d] = du[d]; a.c[d] = dc[d]; a.p[d] = dp[d]; }
    a.mask = dmask; a.cc = dcc; a.ctxsum = dctx; a.bar = dbar; a.bar_ints = nbar; a.scratch = dscr; a.scratch_floats = scr; a.n = n; a.D = D;
    EncTcBwdArgs bw; memset(&bw, 0, sizeof(bw));
    for (int d = 0; d < 2; ++d) { bw.Ucat[d] = dU[d]; bw.r[d] = dr[d]; bw.u[d] = du[d]; bw.c[d] = dc[d]; bw.p[d] = dp[d]; bw.dG[d] = dGd[d]; bw.dGx[d] = dGxd[d]; }
    bw.dcc = ddcc; bw.mean_grad = dmean; bw.coef = dcoef; bw.mask = dmask; bw.cc = dcc; bw.bar = dbar; bw.bar_ints = nbar;
    bw.scratch = dscr; bw.scratch_floats = scr; bw.n = n; bw.D = D;
    // ---- correctness at Tx = Tc
    a.Tx = Tc; a.dbg = nullptr; bw.Tx = Tc; bw.dbg = nullptr;
    if (enc_tc_fwd(&ctx, 0, a)) { printf("launch failed: %s\n", g_err); return 1; }
    CK(cudaDeviceSynchronize());
    std::vector<float> hcc((size_t)Tc * n * C), hs[4], hctx((size_t)n * C);
    CK(cudaMemcpy(hcc.data(), dcc, hcc.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hctx.data(), dctx, hctx.size() * 4, cudaMemcpyDeviceToHost));
    std::vector<double> rcc((size_t)Tc * n * C, 0.0);
    std::vector<double> svd[2];
    double worst = 0.0;
    for (int d = 0; d < 2; ++d) {
        svd[d].resize((size_t)Tc * n * D * 4);
        cpu_fwd(Tc, n, D, U[d], xp[d], mask, d, rcc, svd[d]);
        float* dev[4] = {dr[d], du[d], dc[d], dp[d]};
        const char* nm[4] = {"r", "u", "c", "p"};
        for (int q = 0; q < 4; ++q) {
            hs[q].resize((size_t)Tc * n * D);
            CK(cudaMemcpy(hs[q].data(), dev[q], hs[q].size() * 4, cudaMemcpyDeviceToHost));
            double e = 0.0;
            for (size_t i = 0; i < hs[q].size(); ++i) { const double dd = fabs((double)hs[q][i] - svd[d][i * 4 + q]); if (!(dd <= e)) e = dd; }
            printf("dir %d saved %s: max abs err %.3e\n", d, nm[q], e);
            if (!(e <= worst)) worst = e;
        }
    }
    double e = maxerr(hcc, rcc, nullptr), ectx = 0.0;
    for (int b = 0; b < n; ++b) for (int j = 0; j < C; ++j) {
        double sref = 0.0;
        for (int t = 0; t < Tc; ++t) sref += mask[t * n + b] * rcc[((size_t)t * n + b) * C + j];
        const double dd = fabs((double)hctx[b * C + j] - sref); if (!(dd <= ectx)) ectx = dd;
    }
    printf("cc: max abs err %.3e   ctxsum: max abs err %.3e\n", e, ectx);
    bool pass = e < 2e-5 && worst < 2e-5 && ectx < 2e-4;
    printf("CHECK forward %s\n", pass ? "PASS" : "FAIL");
    // ---- backward on the device's own forward results
    if (enc_tc_bwd(&ctx, 0, bw)) { printf("launch failed: %s\n", g_err); return 1; }
    CK(cudaDeviceSynchronize());
    bool passb = true;
    for (int d = 0; d < 2; ++d) {
        std::vector<double> rG((size_t)Tc * n * D3, 0.0), rGx((size_t)Tc * n * D3, 0.0);
        cpu_bwd(Tc, n, D, U[d], mask, d, rcc, svd[d], hdcc, hmean, hcoef, rG, rGx);
        std::vector<float> hG((size_t)Tc * n * D3), hGx((size_t)Tc * n * D3);
        CK(cudaMemcpy(hG.data(), dGd[d], hG.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hGx.data(), dGxd[d], hGx.size() * 4, cudaMemcpyDeviceToHost));
        double s1, s2;
        const double e1 = maxerr(hG, rG, &s1), e2 = maxerr(hGx, rGx, &s2);
        printf("dir %d dG: max abs err %.3e (max |ref| %.3e)   dGx: %.3e (%.3e)\n", d, e1, s1, e2, s2);
        if (!(e1 < 1e-5 * (1.0 + s1)) || !(e2 < 1e-5 * (1.0 + s2))) passb = false;
    }
    printf("CHECK backward %s\n", passb ? "PASS" : "FAIL");
    // ---- timing at Tx = Tt
    a.Tx = Tt; a.dbg = ddbg; bw.Tx = Tt; bw.dbg = ddbg + 64;

    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int it = 0; it < 3; ++it) {
        CK(cudaEventRecord(e0));
        if (enc_tc_fwd(&ctx, 0, a)) { printf("launch failed: %s\n", g_err); return 1; }
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("forward  Tx=%d: %.3f ms  = %.2f us/step\n", Tt, ms, ms * 1000.0 / Tt);
    }
    for (int it = 0; it < 3; ++it) {
        CK(cudaEventRecord(e0));
        if (enc_tc_bwd(&ctx, 0, bw)) { printf("launch failed: %s\n", g_err); return 1; }
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("backward Tx=%d: %.3f ms  = %.2f us/step\n", Tt, ms, ms * 1000.0 / Tt);
    }
    unsigned long long hd[128]; CK(cudaMemcpy(hd, ddbg, 128 * 8, cudaMemcpyDeviceToHost));
    for (int k = 0; k < 2; ++k) {
        const unsigned long long* h = hd + 64 * k;
        const double mhz = (double)(h[32 + 5] - h[32 + 0]) / (double)(h[5] - h[0]) * 1000.0;
        printf("%s step 8, CTA 0, SM clock %.0f MHz; ns after the direction flag was seen:\n", k ? "bwd" : "fwd", mhz);
        printf("   first k-block issued %lld | last MMA issued %lld | accum ready %lld | tmem read + partial stores %lld | partials complete %lld | gates+stores %lld | arrive %lld\n",
               (long long)(h[1] - h[0]), (long long)(h[2] - h[0]), (long long)(h[3] - h[0]), (long long)(h[6] - h[0]),
               (long long)(h[4] - h[0]), (long long)(h[9] - h[0]), (long long)(h[5] - h[0]));
    }
    return pass && passb ? 0 : 3;
}
