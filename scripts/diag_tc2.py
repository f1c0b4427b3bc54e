"""MN-major layout experiment + timing of the K-major kernel at the layer-1 shape."""
import sys, os, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np, torch
    from recnn_b200 import _lib
    from scripts.diag_tc import run, normal
    for (M, N, K) in [(128, 256, 16), (256, 128, 200)]:
        for a_mn, b_mn in [(0, 1), (1, 1), (1, 0)]:
            got, want = run(M, N, K, a_mn, b_mn, normal)
            print("variant", os.environ.get("RECNN_TC_MN_VARIANT"), M, N, K, "a_mn", a_mn, "b_mn", b_mn,
                  "rel %.3g" % (np.abs(got - want).max() / np.abs(want).max()))
    sys.exit(0)
for v in "0123":
    env = dict(os.environ, RECNN_TC_MN_VARIANT=v)
    r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True, timeout=200)
    print("\n".join(l for l in r.stdout.splitlines() if l.startswith("variant")))
    if r.returncode != 0:
        print("variant", v, "child failed:", r.stderr[-400:])
import numpy as np, torch
from recnn_b200 import _lib
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
for (M, N, K) in [(4096, 256, 1290), (4096, 256, 256), (16384, 256, 1290)]:
    ld = (K + 3) // 4 * 4
    A = torch.randn(M, ld, device=DEV); B = torch.randn(N, ld, device=DEV); C = torch.empty(M, N, device=DEV)
    for tile in (64, 128):
        ms = timeit(lambda: _lib.check(L.recnn_gemm_tf32x3(M, N, K, A.data_ptr(), ld, 0, B.data_ptr(), ld, 0, C.data_ptr(), N, tile, st)))
        print("tc   M%d N%d K%d tile_n %3d: %.3f ms  %.1f TFLOP/s (x3 passes: %.1f)" % (M, N, K, tile, ms, 2.0 * M * N * K / ms / 1e9, 6.0 * M * N * K / ms / 1e9))
    ms = timeit(lambda: _lib.check(L.recnn_gemm_fp32(M, N, K, A.data_ptr(), ld, 0, B.data_ptr(), ld, 0, C.data_ptr(), N, st)))
    print("simt M%d N%d K%d: %.3f ms  %.1f TFLOP/s" % (M, N, K, ms, 2.0 * M * N * K / ms / 1e9))
