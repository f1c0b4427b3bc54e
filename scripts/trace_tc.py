import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from recnn_b200 import _lib
L = _lib.lib(); DEV = "cuda:0"
h = ctypes.CDLL(_lib.lib_path())
h.recnn_debug_set_trace.argtypes = [ctypes.c_void_p]
st = torch.cuda.current_stream().cuda_stream
names = ["entry", "prologue", "first_split", "last_mma_issued", "last_split", "drained", "epilogue", "exit"]
for (M, N, K, tile) in [(4096, 256, 256, 128), (4096, 256, 1290, 128), (4096, 128, 256, 64)]:
    ld = (K + 3) // 4 * 4
    A = torch.randn(M, ld, device=DEV); B = torch.randn(N, ld, device=DEV); C = torch.empty(M, N, device=DEV)
    ncta = ((M + 127) // 128) * ((N + tile - 1) // tile)
    tr = torch.zeros(ncta * 8, dtype=torch.int64, device=DEV)
    for it in range(3):
        tr.zero_(); torch.cuda.synchronize()
        h.recnn_debug_set_trace(tr.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.recnn_gemm_tf32x3(M, N, K, A.data_ptr(), ld, 0, B.data_ptr(), ld, 0, C.data_ptr(), N, tile, st))
        e1.record(); torch.cuda.synchronize()
        h.recnn_debug_set_trace(None)
    t = tr.cpu().numpy().reshape(ncta, 8).astype(np.float64)
    t0 = t[:, 0].min()
    rel = (t - t0) / 1000.0
    print("M%d N%d K%d tile%d ctas %d  event time %.1f us" % (M, N, K, tile, ncta, e0.elapsed_time(e1) * 1000))
    for i, nm in enumerate(names):
        print("   %-16s mean %7.2f us  min %7.2f  max %7.2f" % (nm, rel[:, i].mean(), rel[:, i].min(), rel[:, i].max()))
