import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tools')
import gemm_bench as G
for (name, M, N, K, ta, tb, sk) in [('enc fwd h.U (1 dir)', 32, 3000, 1000, 0, 0, 6), ('enc fwd h.U (1 dir)', 32, 3000, 1000, 0, 0, 3),
                                    ('enc fwd h.U (1 dir)', 32, 3000, 1000, 0, 0, 12), ('enc bwd', 32, 1000, 3000, 0, 1, 18), ('tiny', 32, 128, 32, 0, 0, 1)]:
    for path in (0, 2):
        print('%-22s path %d splitk %2d  eager %7.2f us   in-graph %7.2f us' % (name, path, sk, G.bench(path, M, N, K, ta, tb, sk), G.bench_graph(path, M, N, K, ta, tb, sk)))
