import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from recnn_b200 import _lib
from scripts.diag_tc import run, normal, positive
print("== accumulation probe (all-positive): signed mean rel err")
for K in (64, 256, 1290, 4096):
    for kind in ("tc", "simt"):
        got, want = run(256, 256, K, 0, 0, positive, kind=kind)
        r = (got - want) / want
        print("K", K, kind, "mean %.3e rms %.3e max %.3e" % (r.mean(), np.sqrt((r ** 2).mean()), np.abs(r).max()))
print("== normal data: rms err / sqrt(K)")
for K in (64, 256, 1290, 4096):
    for kind in ("tc", "simt"):
        got, want = run(256, 256, K, 0, 0, normal, kind=kind)
        d = got - want
        print("K", K, kind, "rms/sqrtK %.3e max/sqrtK %.3e" % (np.sqrt((d ** 2).mean()) / np.sqrt(K), np.abs(d).max() / np.sqrt(K)))
L = _lib.lib(); DEV = "cuda:0"
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(64 * 1024 * 1024, device=DEV)
def timeit(fn, iters=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
for (M, N, K, amn, bmn) in [(4096, 256, 1290, 0, 0), (4096, 256, 256, 0, 0), (16384, 256, 1290, 0, 0), (4096, 256, 256, 0, 1), (256, 1290, 4096, 1, 1)]:
    def mk(rows, cols):
        ld = (cols + 3) // 4 * 4
        return torch.randn(rows, ld, device=DEV), ld
    A, lda = mk(K, M) if amn else mk(M, K)
    B, ldb = mk(K, N) if bmn else mk(N, K)
    ldc = (N + 3) // 4 * 4
    C = torch.empty(M, ldc, device=DEV)
    for tile in (64, 128):
        ms = timeit(lambda: _lib.check(L.recnn_gemm_tf32x3(M, N, K, A.data_ptr(), lda, amn, B.data_ptr(), ldb, bmn, C.data_ptr(), ldc, tile, st)))
        print("tc   M%d N%d K%d amn%d bmn%d tile_n %3d: %.3f ms  %.1f TFLOP/s (x3: %.1f)" % (M, N, K, amn, bmn, tile, ms, 2.0 * M * N * K / ms / 1e9, 6.0 * M * N * K / ms / 1e9))
