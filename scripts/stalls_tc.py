"""Where the tcgen05 GEMM's warps wait (dbg bit 16 = stall accounting, clock64 cycles per CTA)."""
import sys, os, subprocess, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import numpy as np, torch
    from recnn_b200 import _lib
    L = _lib.lib(); DEV = "cuda:0"
    h = ctypes.CDLL(_lib.lib_path())
    h.recnn_debug_set_trace.argtypes = [ctypes.c_void_p]
    st = torch.cuda.current_stream().cuda_stream
    names = ["wk.full", "wk.a_free", "wk.drain", "wk.loop", "mma.split", "mma.acc_empty", "tma.empty", "wk.drain_ld"]
    for (M, N, K, tile, amn, bmn) in [(4096, 256, 1290, 128, 0, 0), (4096, 256, 1290, 64, 0, 0), (4096, 256, 256, 64, 0, 0)]:
        if amn:
            A = torch.randn(K, M, device=DEV); B = torch.randn(K, (N + 3) // 4 * 4, device=DEV); lda, ldb = M, B.shape[1]
        else:
            ld = (K + 3) // 4 * 4
            A = torch.randn(M, ld, device=DEV); B = torch.randn(N, ld, device=DEV); lda = ldb = ld
        ldc = (N + 3) // 4 * 4; C = torch.empty(M, ldc, device=DEV)
        ncta = ((M + 127) // 128) * ((N + tile - 1) // tile)
        tr = torch.zeros(ncta * 16, dtype=torch.int64, device=DEV)
        for it in range(4):
            tr.zero_(); torch.cuda.synchronize()
            h.recnn_debug_set_trace(tr.data_ptr())
            _lib.check(L.recnn_gemm_tf32x3(M, N, K, A.data_ptr(), lda, amn, B.data_ptr(), ldb, bmn, C.data_ptr(), ldc, tile, st))
            torch.cuda.synchronize()
            h.recnn_debug_set_trace(None)
        t = tr.cpu().numpy().astype(np.float64)
        stamps, prof = t[:ncta * 8].reshape(ncta, 8), t[ncta * 8:].reshape(ncta, 8)
        span = (stamps[:, 7].max() - stamps[:, 0].min()) / 1e3
        med = np.median(prof, axis=0)
        print("dbg=%-3s M%d N%d K%d t%d mn%d%d span %.1f us | " % (os.environ.get("RECNN_TC_DBG"), M, N, K, tile, amn, bmn, span) +
              "  ".join("%s %.0f" % (n, v) for n, v in zip(names, med)))
    sys.exit(0)
# RECNN_B200_WORKERS16=1 in the environment profiles the 16-worker kernel at tile 64
for dbg in (("16",) if os.environ.get("STALLS_QUICK") else ("16", str(16 + (4 << 8)), str(16 + (8 << 8)))):
    r = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, RECNN_TC_DBG=dbg), capture_output=True, text=True, timeout=120)
    print(r.stdout.strip() or r.stderr[-600:])
