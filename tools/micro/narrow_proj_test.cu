// narrow_proj_test.cu -- standalone check + timing of the few-row kernels of ops_readout.cu (narrow_proj, cluster softmax).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o tools/micro/narrow_proj_test tools/micro/narrow_proj_test.cu
#include <cstdio>
#include <cstdlib>
#include <cstdarg>
#include <cmath>
#include <vector>
#include "../../nats_b200/csrc/ops_readout.cu"

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

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 10, D = 1000, W = 100, C = 2000;
    std::vector<float> hx0(n * D), hx1(n * W), hx2(n * C), hW0((size_t)D * W), hW1((size_t)W * W), hW2((size_t)C * W), hb(3 * W);
    unsigned s = 12345;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffff) / 65536.0f - 0.5f; };
    for (auto* v : {&hx0, &hx1, &hx2, &hW0, &hW1, &hW2, &hb}) for (auto& e : *v) e = 0.1f * rnd();
    float *x0, *x1, *x2, *W0, *W1, *W2, *b, *out, *flush;
    CK(cudaMalloc(&x0, hx0.size() * 4)); CK(cudaMalloc(&x1, hx1.size() * 4)); CK(cudaMalloc(&x2, hx2.size() * 4));
    CK(cudaMalloc(&W0, hW0.size() * 4)); CK(cudaMalloc(&W1, hW1.size() * 4)); CK(cudaMalloc(&W2, hW2.size() * 4));
    CK(cudaMalloc(&b, hb.size() * 4)); CK(cudaMalloc(&out, n * W * 4)); CK(cudaMalloc(&flush, 256u << 20));
    CK(cudaMemcpy(x0, hx0.data(), hx0.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(x1, hx1.data(), hx1.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(x2, hx2.data(), hx2.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(W0, hW0.data(), hW0.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(W1, hW1.data(), hW1.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(W2, hW2.data(), hW2.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(b, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice));
    NarrowProj a;
    memset(&a, 0, sizeof(a));
    a.x[0] = x0; a.ldx[0] = D; a.W[0] = W0; a.ldw[0] = W; a.bias[0] = b; a.K[0] = D;
    a.x[1] = x1; a.ldx[1] = W; a.W[1] = W1; a.ldw[1] = W; a.bias[1] = b + W; a.K[1] = W;
    a.x[2] = x2; a.ldx[2] = C; a.W[2] = W2; a.ldw[2] = W; a.bias[2] = b + 2 * W; a.K[2] = C;
    a.n = n; a.N = W; a.out = out; a.ldo = W; a.act_tanh = 1;
    nats_ctx cx; memset(&cx, 0, sizeof(cx)); { int v; cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, 0); cx.max_smem_optin = v; }
    if (narrow_proj_setup(&cx) != 0) { printf("setup failed: %s\n", g_err); return 1; }
    if (!narrow_proj_eligible(a)) { printf("not eligible\n"); return 1; }
    if (narrow_proj(0, a) != 0) { printf("launch failed: %s\n", g_err); return 1; }
    CK(cudaDeviceSynchronize());
    std::vector<float> ho(n * W);
    CK(cudaMemcpy(ho.data(), out, ho.size() * 4, cudaMemcpyDeviceToHost));
    double worst = 0;
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < W; ++c) {
            double v = hb[c] + hb[W + c] + hb[2 * W + c];
            for (int k = 0; k < D; ++k) v += (double)hx0[r * D + k] * hW0[(size_t)k * W + c];
            for (int k = 0; k < W; ++k) v += (double)hx1[r * W + k] * hW1[(size_t)k * W + c];
            for (int k = 0; k < C; ++k) v += (double)hx2[r * C + k] * hW2[(size_t)k * W + c];
            worst = fmax(worst, fabs(tanh(v) - ho[r * W + c]));
        }
    printf("narrow_proj n=%d: max abs err %.3e\n", n, worst);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int mode = 0; mode < 2; ++mode) {       // 0: warm L2, 1: L2 flushed before every launch
        float tot = 0;
        const int reps = 20;
        for (int i = 0; i < reps; ++i) {
            if (mode) CK(cudaMemsetAsync(flush, i, 256u << 20, 0));
            CK(cudaEventRecord(e0, 0));
            narrow_proj(0, a);
            CK(cudaEventRecord(e1, 0));
            CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); tot += ms;
        }
        printf("  %s: %.2f us per launch (event-bracketed single launches)\n", mode ? "cold L2" : "warm L2", tot / reps * 1e3);
    }
    return 0;
}
