"""Kernel-internal (globaltimer) span of the layer-1 GEMM under the Problem.dbg ablation bits.

bit 1: no hi/lo split   bit 2: no MMA issue   bit 4: no hi store   bit 8: no proxy fence
"""
import sys, os, subprocess, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import numpy as np, torch
    from recnn_b200 import _lib
    L = _lib.lib(); DEV = "cuda:0"
    h = ctypes.CDLL(_lib.lib_path())
    h.recnn_debug_set_trace.argtypes = [ctypes.c_void_p]
    st = torch.cuda.current_stream().cuda_stream
    out = []
    for (M, N, K, tile, amn, bmn) in [(4096, 256, 1290, 128, 0, 0), (4096, 256, 1290, 64, 0, 0), (4096, 256, 256, 128, 0, 0)]:
        ld = (K + 3) // 4 * 4
        A = torch.randn(M, ld, device=DEV); B = torch.randn(N, ld, device=DEV); C = torch.empty(M, N, device=DEV)
        ncta = ((M + 127) // 128) * ((N + tile - 1) // tile)
        tr = torch.zeros(ncta * 8, dtype=torch.int64, device=DEV)
        spans = []
        for it in range(6):
            tr.zero_(); torch.cuda.synchronize()
            h.recnn_debug_set_trace(tr.data_ptr())
            _lib.check(L.recnn_gemm_tf32x3(M, N, K, A.data_ptr(), ld, amn, B.data_ptr(), ld, bmn, C.data_ptr(), N, tile, st))
            torch.cuda.synchronize()
            h.recnn_debug_set_trace(None)
            t = tr.cpu().numpy().reshape(ncta, 8).astype(np.float64)
            spans.append(((t[:, 7].max() - t[:, 0].min()) / 1e3, np.median(t[:, 7] - t[:, 0]) / 1e3,
                          np.median(t[:, 5] - t[:, 1]) / 1e3))
        s = sorted(spans)[len(spans) // 2]
        out.append("K%d/t%d: span %.1f cta %.1f main %.1f" % (K, tile, s[0], s[1], s[2]))
    print("dbg=%-2s  %s" % (os.environ.get("RECNN_TC_DBG", "0"), "   ".join(out)))
    sys.exit(0)
for dbg in ("0",):
    r = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, RECNN_TC_DBG=dbg), capture_output=True, text=True, timeout=120)
    print(r.stdout.strip() or r.stderr[-300:])
