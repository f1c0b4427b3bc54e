"""A/B timing of the tcgen05 GEMM variants (in-kernel B split = default, pre-split B planes, 16 workers, two
cross-term accumulators, LEAN issue loop), each alone on the GPU, on the shapes of the update step.  CUDA events, L2 flushed between launches.  Prints one JSON line.

    python scripts/ab_gemm_variants.py > gpurun_out/ab_gemm.json
"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recnn_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
L = _lib.lib()
pre = L.recnn_debug_gemm_tf32x3_presplit
pre.restype = C.c_int
pre.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_void_p,
                C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
st = torch.cuda.current_stream().cuda_stream


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts)), float(np.min(ts))


def pad4(n):
    return (n + 3) // 4 * 4


out = []
for name, M, N, K, b_mn in [("L1 fwd actor  [4096x1290]x[1290x256]", 4096, 256, 1290, 0),
                            ("L2 fwd        [4096x256]x[256x256]", 4096, 256, 256, 0),
                            ("L3 fwd actor  [4096x256]x[256x128]", 4096, 128, 256, 0),
                            ("dX through W2 [4096x256]x[256x256] (B MN-major)", 4096, 256, 256, 1)]:
    lda, ldc = pad4(K), pad4(N)
    a = torch.randn(M, lda, device=dev)
    b = torch.randn(K, pad4(N), device=dev) if b_mn else torch.randn(N, pad4(K), device=dev)
    ldb = b.shape[1]
    c0 = torch.empty(M, ldc, device=dev)
    c1 = torch.empty(M, ldc, device=dev)
    hi, lo = torch.empty_like(b), torch.empty_like(b)
    for tile in (64, 128):
        if tile == 128 and N < 128:
            continue
        _lib.check(pre(M, N, K, a.data_ptr(), lda, b.data_ptr(), ldb, b_mn, c1.data_ptr(), ldc, tile, hi.data_ptr(),
                       lo.data_ptr(), st))          # fills the planes
        classic = timeit(lambda: _lib.check(L.recnn_gemm_tf32x3(M, N, K, a.data_ptr(), lda, 0, b.data_ptr(), ldb, b_mn,
                                                                c0.data_ptr(), ldc, tile, st)))
        presplit = timeit(lambda: _lib.check(pre(M, N, K, a.data_ptr(), lda, None, ldb, b_mn, c1.data_ptr(), ldc, tile,
                                                 hi.data_ptr(), lo.data_ptr(), st)))
        same = bool(torch.equal(c0[:, :N], c1[:, :N]))
        rec = {"gemm": name, "tile_n": tile, "classic_us_median": classic[0], "classic_us_min": classic[1],
               "presplit_us_median": presplit[0], "presplit_us_min": presplit[1], "bit_identical": same,
               "tf32_tflops_presplit": 3 * 2.0 * M * N * K / (presplit[0] * 1e-6) / 1e12}
        if tile == 64:                   # four split groups (16 worker warps) instead of two
            c2 = torch.empty(M, ldc, device=dev)
            prev = _lib.set_option("workers16", 1)
            try:
                w16 = timeit(lambda: _lib.check(L.recnn_gemm_tf32x3(M, N, K, a.data_ptr(), lda, 0, b.data_ptr(), ldb,
                                                                    b_mn, c2.data_ptr(), ldc, tile, st)))
            finally:
                _lib.set_option("workers16", prev)
            rec.update({"workers16_us_median": w16[0], "workers16_us_min": w16[1],
                        "workers16_bit_identical": bool(torch.equal(c0[:, :N], c2[:, :N])),
                        "tf32_tflops_workers16": 3 * 2.0 * M * N * K / (w16[0] * 1e-6) / 1e12})
        if tile == 64:                   # two cross-term accumulators (three-way accumulator rotation), 8 and 16 workers
            for w16 in (0,):
                c3 = torch.empty(M, ldc, device=dev)
                p0, p1 = _lib.set_option("lo2", 1), _lib.set_option("workers16", w16)
                try:
                    t = timeit(lambda: _lib.check(L.recnn_gemm_tf32x3(M, N, K, a.data_ptr(), lda, 0, b.data_ptr(), ldb,
                                                                      b_mn, c3.data_ptr(), ldc, tile, st)))
                finally:
                    _lib.set_option("lo2", p0)
                    _lib.set_option("workers16", p1)
                key = "lo2_w16" if w16 else "lo2"
                rec.update({key + "_us_median": t[0], key + "_us_min": t[1],
                            key + "_max_abs_diff": float((c3[:, :N] - c0[:, :N]).abs().max()),
                            "tf32_tflops_" + key: 3 * 2.0 * M * N * K / (t[0] * 1e-6) / 1e12})
        # LEAN kernels (experiment hooks compiled out, running counters in the MMA warp), 8 and (tile 64) 16 workers
        for w16 in ((0, 1) if tile == 64 else (0,)):
            c4 = torch.empty(M, ldc, device=dev)
            p0, p1 = _lib.set_option("lean", 1), _lib.set_option("workers16", w16)
            try:
                t = timeit(lambda: _lib.check(L.recnn_gemm_tf32x3(M, N, K, a.data_ptr(), lda, 0, b.data_ptr(), ldb,
                                                                  b_mn, c4.data_ptr(), ldc, tile, st)))
            except Exception as exc:                       # a trap in an unvalidated kernel must not lose the other rows
                rec["lean_error"] = str(exc)[:200]
                break
            finally:
                _lib.set_option("lean", p0)
                _lib.set_option("workers16", p1)
            key = "lean_w16" if w16 else "lean"
            rec.update({key + "_us_median": t[0], key + "_us_min": t[1],
                        key + "_bit_identical": bool(torch.equal(c0[:, :N], c4[:, :N])),
                        "tf32_tflops_" + key: 3 * 2.0 * M * N * K / (t[0] * 1e-6) / 1e12})
        out.append(rec)
print(json.dumps({"ab_presplit": out}))
