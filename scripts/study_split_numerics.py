"""CPU study (numpy, no GPU): operand-splitting schemes for fp32-grade GEMMs on reduced-precision tensor cores.

  tf32x3 : x = hi + lo, hi = rna_tf32(x), lo = rna_tf32(x - hi); products hi*hi + lo*hi + hi*lo      (3 MMAs of K=8)
  fp16x2 : per-row power-of-two scale s, xs = x*s; h0 = fp16(xs), h1 = fp16(xs - h0); h0*h0 + h0*h1 + h1*h0
           (3 MMAs of K=16 at twice the tf32 rate => half the tensor time and half the MMA instructions)
  bf16x3 : x = b0 + b1 + b2 (bf16 each), 6 products                                                   (no gain, for reference)

Products and sums are exact (float64) here: this isolates the OPERAND representation error, which is what differs
between the schemes; the accumulation error (tensor-core fp32 accumulate, chunked) is common to all of them.
Prints max|err| / (sqrt(K) * scale) and the error relative to the largest output, next to plain fp32 matmul.
"""
import numpy as np


def rna_tf32(x):
    b = x.astype(np.float32).view(np.uint32)
    return ((b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def split_tf32(x):
    hi = rna_tf32(x)
    lo = rna_tf32((x - hi).astype(np.float32))
    return hi.astype(np.float64), lo.astype(np.float64)


def split_fp16(x):
    """rows scaled so that the row maximum lands in [2^13, 2^14) (fp16 max is 65504)"""
    m = np.max(np.abs(x), axis=1, keepdims=True)
    e = np.where(m > 0, np.floor(np.log2(np.maximum(m, 1e-300))), 0.0)
    s = np.exp2(13.0 - e)
    xs = (x.astype(np.float64) * s).astype(np.float32)
    h0 = xs.astype(np.float16)
    h1 = (xs - h0.astype(np.float32)).astype(np.float16)
    return h0.astype(np.float64), h1.astype(np.float64), s


def split_bf16(x):
    def bf(v):
        b = v.astype(np.float32).view(np.uint32)
        return ((b + np.uint32(0x8000)) & np.uint32(0xFFFF0000)).view(np.float32)
    b0 = bf(x)
    b1 = bf((x - b0).astype(np.float32))
    b2 = bf((x - b0 - b1).astype(np.float32))
    return [v.astype(np.float64) for v in (b0, b1, b2)]


def study(name, A, B):
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    K = A.shape[1]
    scale = np.abs(ref).max()
    out = {"fp32 matmul": (A.astype(np.float32) @ B.astype(np.float32).T).astype(np.float64)}
    ah, al = split_tf32(A)
    bh, bl = split_tf32(B)
    out["tf32x3"] = ah @ bh.T + al @ bh.T + ah @ bl.T
    a0, a1, sa = split_fp16(A)
    b0, b1, sb = split_fp16(B)
    out["fp16x2 (row scaled)"] = (a0 @ b0.T + a0 @ b1.T + a1 @ b0.T) / sa / sb.T
    x0, x1, x2 = split_bf16(A)
    y0, y1, y2 = split_bf16(B)
    out["bf16x3 (6 products)"] = x0 @ y0.T + x0 @ y1.T + x1 @ y0.T + x1 @ y1.T + x0 @ y2.T + x2 @ y0.T
    print("%s  [M=%d N=%d K=%d]" % (name, A.shape[0], B.shape[0], K))
    for k, v in out.items():
        err = np.abs(v - ref)
        rowrel = (err / (np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64).T + 1e-300)).max()
        print("   %-22s max|err|/max|C| %.2e   max|err| / (|A||B|^T) %.2e" % (k, err.max() / scale, rowrel))


rng = np.random.default_rng(0)
M, N, K = 512, 256, 1290
act = np.maximum(rng.standard_normal((M, K)), 0) * 2.0 * rng.integers(0, 2, (M, K))      # relu + dropout activations
w = rng.uniform(-0.028, 0.028, (N, K))
study("forward: dropout(relu(z)) x weights", act.astype(np.float32), w.astype(np.float32))
state = rng.standard_normal((M, K)) * np.where(rng.random((M, K)) < 0.01, 50.0, 1.0)     # embeddings with outliers
study("layer 1: embeddings with 50x outliers x weights", state.astype(np.float32), w.astype(np.float32))
g = rng.standard_normal((M, 256)) * np.exp(rng.uniform(-14, -2, (M, 1)) * np.log(10) / 2.3)  # rows spanning decades
w2 = rng.uniform(-0.06, 0.06, (256, 256))
study("backward: gradient rows spanning 5 decades x W2^T", g.astype(np.float32), w2.T.copy().astype(np.float32))
gz = (rng.standard_normal((4096, 256)) * 1e-6).astype(np.float32)                           # dW = dZ^T X over 4096 rows
x = np.maximum(rng.standard_normal((4096, 256)), 0).astype(np.float32)
study("weight gradient: dZ^T (1e-6) x activations, K = 4096 rows", gz.T.copy(), x.T.copy())
