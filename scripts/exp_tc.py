import sys, os, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import numpy as np, torch
    from recnn_b200 import _lib
    L = _lib.lib(); DEV = "cuda:0"; st = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(64 * 1024 * 1024, device=DEV)
    for (M, N, K, tile) in [(4096, 256, 1290, 128), (4096, 256, 1290, 64), (16384, 256, 1290, 128)]:
        ld = (K + 3) // 4 * 4
        A = torch.randn(M, ld, device=DEV); B = torch.randn(N, ld, device=DEV); C = torch.empty(M, N, device=DEV)
        f = lambda: _lib.check(L.recnn_gemm_tf32x3(M, N, K, A.data_ptr(), ld, 0, B.data_ptr(), ld, 0, C.data_ptr(), N, tile, st))
        for _ in range(3): f()
        ts = []
        for _ in range(10):
            flush.zero_(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        print("dbg=%s M%d tile%d: %.1f us" % (os.environ.get("RECNN_TC_DBG", "0"), M, tile, 1000 * float(np.median(ts))))
    sys.exit(0)
for dbg in ("0", "1", "2", "3", "4", "8", "12"):
    r = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, RECNN_TC_DBG=dbg), capture_output=True, text=True, timeout=120)
    print(r.stdout.strip() or r.stderr[-300:])
