import ctypes, sys, numpy as np, torch
sys.path.insert(0, '.')
from nats_b200 import nats, _lib
eng = nats.get_engine()
def run(path, M, N, K, ta=0, tb=1, pattern='rand'):
    rng = np.random.RandomState(1)
    A = rng.randn(K if ta else M, M if ta else K).astype('float32')
    B = rng.randn(N if tb else K, K if tb else N).astype('float32')
    if pattern == 'ones':
        A[:] = 1; B[:] = 1
    if pattern == 'rowid':      # A(i,k) = i, B = 1/K  -> C[i,j] = i
        a = np.arange(M, dtype='float32')[:, None] * np.ones((1, K), 'float32'); A = a.T.copy() if ta else a
        B[:] = 1.0 / K
    if pattern == 'colid':
        A[:] = 1.0 / K
        b = np.arange(N, dtype='float32')[:, None] * np.ones((1, K), 'float32'); B = b.copy() if tb else b.T.copy()
    if pattern == 'kid':        # A(i,k)=1 if k==i%K
        A[:] = 0; 
        for i in range(M):
            if ta: A[i % K, i] = 1
            else: A[i, i % K] = 1
        b = (np.arange(N)[:, None] * 1000 + np.arange(K)[None, :]).astype('float32'); B = b.copy() if tb else b.T.copy()
    Ad, Bd = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    Cd = torch.zeros((M, N), device='cuda')
    rc = eng.lib.nats_debug_gemm(eng.ctx, eng.stream(), path, ta, tb, M, N, K, ctypes.c_void_p(Ad.data_ptr()), A.shape[1],
                                 ctypes.c_void_p(Bd.data_ptr()), B.shape[1], ctypes.c_void_p(Cd.data_ptr()), N,
                                 ctypes.c_void_p(0), 0, 1, 1, 0, 0, 0)
    _lib.check(rc); torch.cuda.synchronize()
    a = A.T if ta else A; b = B.T if tb else B
    return Cd.cpu().numpy(), a.astype('float64') @ b.astype('float64')
np.set_printoptions(linewidth=200, precision=4, suppress=True)
import itertools
PATH = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for pat, (ta, tb) in itertools.product(['ones', 'rowid', 'colid', 'kid', 'rand'], [(0, 1), (0, 0), (1, 0), (1, 1)]):
    got, ref = run(PATH, 128, 128, 32, ta, tb, pat)
    print('ta,tb', ta, tb, end=' ')
    print(pat, 'maxerr', np.abs(got - ref).max(), 'ref absmax', np.abs(ref).max())
    if np.abs(got - ref).max() > 1e-3:
        print(' got[:4,:8]\n', got[:4, :8], '\n ref[:4,:8]\n', ref[:4, :8])
        print(' got[64:66,:8]', got[64:66, :8]); print(' got[:4,64:72]', got[:4, 64:72])
        bad = np.argwhere(np.abs(got - ref) > 1e-3); print(' #bad', len(bad), 'first', bad[:5].tolist(), 'last', bad[-3:].tolist())
for (M, N, K) in [(128, 128, 64), (128, 128, 96), (256, 128, 32), (128, 256, 32), (128, 32, 32), (32, 256, 64)]:
    for ta, tb in [(0, 1), (0, 0), (1, 0), (1, 1)]:
        got, ref = run(PATH, M, N, K, ta, tb, 'rand'); print((M, N, K), ta, tb, 'maxerr', np.abs(got - ref).max())
print('--- error vs K (relative to sqrt(K)), signed bias')
for K in [32, 128, 512, 1024, 4096, 12800]:
    for path in (0, PATH):
        got, ref = run(path, 128, 128, K, 0, 1, 'rand')
        d = got - ref
        print('K', K, 'path', path, 'max|err|/sqrtK %.2e' % (np.abs(d).max() / np.sqrt(K)), 'mean(err*sign(ref))/sqrtK %.2e' % ((d * np.sign(ref)).mean() / np.sqrt(K)))
