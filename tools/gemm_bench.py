"""Micro-benchmark of the library's GEMM engine through nats_debug_gemm (eager launches, CUDA-event timed).
   python tools/gemm_bench.py path M N K ta tb [splitk] [iters]"""
import ctypes, sys, numpy as np, torch
sys.path.insert(0, '.')
from nats_b200 import nats, _lib
eng = nats.get_engine()
def bench(path, M, N, K, ta, tb, splitk=1, iters=200, nbuf=1):
    lda = M if ta else K; ldb = K if tb else N
    As = [torch.randn((K if ta else M, lda), device='cuda') for _ in range(nbuf)]
    Bs = [torch.randn((N if tb else K, ldb), device='cuda') for _ in range(nbuf)]
    C = torch.zeros((max(splitk, 1), M, N), device='cuda')
    def run(i):
        A, B = As[i % nbuf], Bs[i % nbuf]
        rc = eng.lib.nats_debug_gemm(eng.ctx, eng.stream(), path, ta, tb, M, N, K, ctypes.c_void_p(A.data_ptr()), lda,
                                     ctypes.c_void_p(B.data_ptr()), ldb, ctypes.c_void_p(C.data_ptr()), N,
                                     ctypes.c_void_p(0), 0, splitk, 1, 0, 0, 0)
        _lib.check(rc)
    for i in range(10): run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): run(i)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    return us
if __name__ == '__main__':
    if len(sys.argv) > 1:
        a = [int(v) for v in sys.argv[1:]]
        print(a, '%.2f us' % bench(*a))
    else:
        for (name, M, N, K, ta, tb, sk) in [('enc fwd h.U  (1 dir)', 32, 3000, 1000, 0, 0, 6), ('enc fwd h.U  (1 dir)', 32, 3000, 1000, 0, 0, 3),
                                            ('enc bwd dG.Ut (1 dir)', 32, 1000, 3000, 0, 1, 18), ('enc bwd dG.Ut (1 dir)', 32, 1000, 3000, 0, 1, 9),
                                            ('dec ctx.W1', 32, 3000, 2000, 0, 0, 6), ('dU weight grad', 1000, 3000, 12768, 1, 0, 1),
                                            ('in-proj', 12800, 3000, 100, 0, 0, 1), ('logits', 960, 30000, 100, 0, 0, 1)]:
            for path in (0, 1, 2):
                us = bench(path, M, N, K, ta, tb, sk)
                fl = 2.0 * M * N * K
                print('%-24s path %d splitk %2d  %9.2f us  %7.1f TFLOP/s  weights %.1f GB/s' % (name, path, sk, us, fl / us / 1e6, 4.0 * N * K / us / 1e3))


def bench_graph(path, M, N, K, ta, tb, splitk, chain=100, reps=20):
    """the same launch `chain` times inside one CUDA graph (dependent chain on one stream)"""
    lda = M if ta else K; ldb = K if tb else N
    A = torch.randn((K if ta else M, lda), device='cuda'); B = torch.randn((N if tb else K, ldb), device='cuda')
    C = torch.zeros((max(splitk, 1), M, N), device='cuda')
    def run():
        rc = eng.lib.nats_debug_gemm(eng.ctx, eng.stream(), path, ta, tb, M, N, K, ctypes.c_void_p(A.data_ptr()), lda,
                                     ctypes.c_void_p(B.data_ptr()), ldb, ctypes.c_void_p(C.data_ptr()), N,
                                     ctypes.c_void_p(0), 0, splitk, 1, 0, 0, 0)
        _lib.check(rc)
    for _ in range(3): run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(chain): run()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * chain)
