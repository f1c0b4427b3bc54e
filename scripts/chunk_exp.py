"""Accumulation-chunk length experiment: speed (globaltimer span) and accuracy for CH overrides (dbg bits 8..11)."""
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
    for (M, N, K, tile) in [(4096, 256, 1290, 128), (4096, 256, 1290, 64), (4096, 256, 256, 128)]:
        ld = (K + 3) // 4 * 4
        A = torch.randn(M, ld, device=DEV); B = torch.randn(N, ld, device=DEV); C = torch.empty(M, N, device=DEV)
        ncta = ((M + 127) // 128) * ((N + tile - 1) // tile)
        tr = torch.zeros(ncta * 16, dtype=torch.int64, device=DEV)
        spans = []
        for it in range(6):
            tr.zero_(); torch.cuda.synchronize()
            h.recnn_debug_set_trace(tr.data_ptr())
            _lib.check(L.recnn_gemm_tf32x3(M, N, K, A.data_ptr(), ld, 0, B.data_ptr(), ld, 0, C.data_ptr(), N, tile, st))
            torch.cuda.synchronize()
            h.recnn_debug_set_trace(None)
            t = tr.cpu().numpy()[:ncta * 8].reshape(ncta, 8).astype(np.float64)
            spans.append((t[:, 7].max() - t[:, 0].min()) / 1e3)
        out.append("K%d/t%d %.1f us" % (K, tile, sorted(spans)[3]))
    rng = np.random.default_rng(11)
    for K in (1290, 4096):
        Kp = (K + 3) // 4 * 4
        a = np.zeros((256, Kp), np.float32); b = np.zeros((256, Kp), np.float32)
        a[:, :K] = rng.uniform(0.5, 1.0, (256, K)); b[:, :K] = rng.uniform(0.5, 1.0, (256, K))
        a_d, b_d = torch.from_numpy(a).to(DEV), torch.from_numpy(b).to(DEV); c_d = torch.empty(256, 256, device=DEV)
        _lib.check(L.recnn_gemm_tf32x3(256, 256, K, a_d.data_ptr(), Kp, 0, b_d.data_ptr(), Kp, 0, c_d.data_ptr(), 256, 0, st))
        want = a.astype(np.float64) @ b.astype(np.float64).T
        rel = (c_d.cpu().numpy() - want) / want
        out.append("pos K%d bias %.2e max %.2e" % (K, rel.mean(), np.abs(rel).max()))
        a = np.zeros((256, Kp), np.float32); b = np.zeros((256, Kp), np.float32)
        a[:, :K] = rng.standard_normal((256, K)); b[:, :K] = rng.standard_normal((256, K))
        a_d, b_d = torch.from_numpy(a).to(DEV), torch.from_numpy(b).to(DEV)
        _lib.check(L.recnn_gemm_tf32x3(256, 256, K, a_d.data_ptr(), Kp, 0, b_d.data_ptr(), Kp, 0, c_d.data_ptr(), 256, 0, st))
        want = a.astype(np.float64) @ b.astype(np.float64).T
        err = c_d.cpu().numpy() - want
        ref32 = (torch.from_numpy(a) @ torch.from_numpy(b).T).numpy() - want
        out.append("nrm K%d rms %.2e (cpu fp32 %.2e)" % (K, np.sqrt((err ** 2).mean()), np.sqrt((ref32 ** 2).mean())))
    print("CH=%s | %s" % ((int(os.environ.get("RECNN_TC_DBG", "0")) >> 8) or "dflt", " | ".join(out)))
    sys.exit(0)
for ch in (0, 4, 8, 15):
    r = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, RECNN_TC_DBG=str(ch << 8)), capture_output=True, text=True, timeout=120)
    print(r.stdout.strip() or r.stderr[-600:])
