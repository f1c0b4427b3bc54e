"""Diagnostics for the tcgen05 GEMM: per-layout correctness table and accumulation-rounding probe."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from recnn_b200 import _lib
DEV = "cuda:0"
L = _lib.lib()

def pad4(n): return (n + 3) // 4 * 4

def run(M, N, K, a_mn, b_mn, gen, tile_n=0, kind="tc"):
    rng = np.random.default_rng(0)
    def mk(rows, cols):
        h = gen(rng, (rows, cols)).astype(np.float32)
        d = torch.zeros(rows, pad4(cols), device=DEV); d[:, :cols] = torch.from_numpy(h).to(DEV)
        return h, d, pad4(cols)
    a_h, a_d, lda = mk(K, M) if a_mn else mk(M, K)
    b_h, b_d, ldb = mk(K, N) if b_mn else mk(N, K)
    ldc = pad4(N); c = torch.full((M, ldc), float("nan"), device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    if kind == "tc":
        _lib.check(L.recnn_gemm_tf32x3(M, N, K, a_d.data_ptr(), lda, int(a_mn), b_d.data_ptr(), ldb, int(b_mn), c.data_ptr(), ldc, tile_n, st))
    else:
        _lib.check(L.recnn_gemm_fp32(M, N, K, a_d.data_ptr(), lda, int(a_mn), b_d.data_ptr(), ldb, int(b_mn), c.data_ptr(), ldc, st))
    torch.cuda.synchronize()
    A = (a_h.T if a_mn else a_h).astype(np.float64); B = (b_h.T if b_mn else b_h).astype(np.float64)
    return c[:, :N].cpu().numpy().astype(np.float64), A @ B.T

normal = lambda rng, s: rng.standard_normal(s)
positive = lambda rng, s: rng.uniform(0.5, 1.0, s)
print("== layout table (normal data): max|err|/max|want|")
for (M, N, K) in [(128, 256, 16), (128, 256, 64), (128, 64, 48), (256, 128, 200), (300, 256, 1290)]:
    for a_mn, b_mn in [(0, 0), (0, 1), (1, 1), (1, 0)]:
        try:
            got, want = run(M, N, K, a_mn, b_mn, normal)
            e = np.abs(got - want).max() / np.abs(want).max()
            bad_rows = np.unique(np.where(np.abs(got - want) > 1e-3 * np.abs(want).max())[0])
            bad_cols = np.unique(np.where(np.abs(got - want) > 1e-3 * np.abs(want).max())[1])
            print(M, N, K, "a_mn", a_mn, "b_mn", b_mn, "rel %.3g" % e, "nan", int(np.isnan(got).sum()),
                  "bad rows", len(bad_rows), list(bad_rows[:6]), "bad cols", len(bad_cols), list(bad_cols[:6]))
        except Exception as ex:
            print(M, N, K, a_mn, b_mn, "EXC", str(ex)[:200])
print("== accumulation rounding probe (all-positive operands, K-major): signed mean rel err, rms rel err")
for K in (16, 64, 256, 1024, 4096):
    for kind in ("tc", "simt"):
        got, want = run(256, 256, K, 0, 0, positive, kind=kind)
        r = (got - want) / want
        print("K", K, kind, "mean %.3e rms %.3e max %.3e" % (r.mean(), np.sqrt((r ** 2).mean()), np.abs(r).max()))
print("== normal data, K-major: rms err / sqrt(K)")
for K in (16, 64, 256, 1024, 4096):
    for kind in ("tc", "simt"):
        got, want = run(256, 256, K, 0, 0, normal, kind=kind)
        d = got - want
        print("K", K, kind, "rms/sqrtK %.3e max/sqrtK %.3e mean %.3e" % (np.sqrt((d ** 2).mean()) / np.sqrt(K), np.abs(d).max() / np.sqrt(K), d.mean()))
