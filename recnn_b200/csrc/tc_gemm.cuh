// tcgen05 (5th-gen tensor core) GEMM with fp32-grade accuracy: error-compensated
// 3xTF32, accumulators in TMEM, operands staged by TMA, sm_100a only.
//
//   C[m,n] = sum_k A(m,k) * B(n,k)            (+ fused epilogue)
//
// Why 3xTF32: the parity bar is the reference's fp32 CPU result to 1e-5
// (BASELINE.json north_star); one TF32 pass has a 10-bit mantissa (~1e-3).  Each
// fp32 operand x is split into  hi = x with the low 13 mantissa bits cleared
// (what kind::tf32 reads from a raw fp32 word -- verified on device by
// tests/test_gpu_tc.py::test_tf32_operand_truncation) and  lo = rna_tf32(x - hi)
// (exact subtraction, then round-to-nearest).  Three MMAs per k-slice
//   lo_a*hi_b + hi_a*lo_b + hi_a*hi_b
// accumulate in fp32 in TMEM; the dropped lo*lo term is < 2^-20 relative.
//
// Pipeline (one 128 x BN output tile per CTA, 192 threads):
//   warp 0      TMA producer: raw fp32 tiles of A and B -> smem stage s        (full[s])
//   warps 2-5   splitter: read raw tile, write the `lo` tile next to it         (split[s])
//               (the raw tile itself is the `hi` operand; nothing is rewritten)
//   warp 1      MMA issuer (one lane): 3 x BK/8 tcgen05.mma per stage, then
//               tcgen05.commit -> empty[s]; after the last k-block -> accum_full
//   warps 2-5   epilogue: tcgen05.ld the accumulator (warp w owns TMEM lanes
//               32*(w%4)..), apply bias/relu/dropout/..., store.
// Operand layouts in smem are the canonical UMMA layouts written by TMA with
// hardware swizzle, so the splitter is swizzle-agnostic (same offset in the
// `lo` buffer) and the same smem descriptors serve hi and lo.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "gemm_simt.cuh"   // Epilogue struct + EpiKind

namespace recnn {
namespace tc {

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must not hang the GPU (a wedged box costs a whole lease), so
// after ~2 s of polling the kernel traps and the launch surfaces as a CUDA error instead.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) {
      printf("recnn_b200: mbarrier wait timed out (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z,
             threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on `bar` once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float tf32_lo(float x) {
  // x = hi + lo exactly, hi = x with the low 13 mantissa bits cleared; return rna_tf32(lo)
  const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  const float lo = x - hi;
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(lo));
  return __uint_as_float(r);
}

// ------------------------------------------------------------------ descriptors
// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout).
__host__ __device__ constexpr uint64_t smem_desc_base(uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  return (uint64_t(lbo_bytes >> 4) << 16) | (uint64_t(sbo_bytes >> 4) << 32) | (uint64_t(1) << 46) |
         (uint64_t(layout_type) << 61);
}
// Instruction descriptor for kind::tf32, fp32 accumulate (cute::UMMA::InstrDescriptor).
__host__ __device__ constexpr uint32_t instr_desc_tf32(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(a_mn) << 15) | (uint32_t(b_mn) << 16) |
         (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

struct Problem {
  int M, N, K0, K1;      // K = K0 (+ K1 from a second A tensor: virtual concat along K, K-major A only)
  int k_chunk;           // k extent per split (multiple of BK), split index = blockIdx.z
  int b_k1_offset;       // B's k coordinate where the K1 segment starts (== K0 for a dense weight)
  int n_out_offset;      // column offset added when storing (C window)
};

template <int BN_, int BK_, int STAGES_, bool A_MN_, bool B_MN_>
struct Cfg {
  static constexpr int BM = 128, BN = BN_, BK = BK_, STAGES = STAGES_;
  static constexpr bool A_MN = A_MN_, B_MN = B_MN_;
  static constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = 2 * (A_BYTES + B_BYTES);      // raw + lo
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int THREADS = 192;
  // K-major operand: rows of BK*4 bytes, swizzle span == row; MN-major: 128-byte rows of 32 elements
  static constexpr int K_SWZ = BK * 4;                              // 64 or 128
  static_assert(BK == 16 || BK == 32, "BK");
  static_assert(BN == 64 || BN == 128 || BN == 256, "BN");
  static_assert(SMEM_BYTES <= 227 * 1024, "smem");
};

// ------------------------------------------------------------------ the kernel
template <class C, int EPI>
__global__ void __launch_bounds__(C::THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
               const __grid_constant__ CUtensorMap map_b, Problem p, Epilogue epi) {
  constexpr int BM = C::BM, BN = C::BN, BK = C::BK, STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  auto stage_ptr = [&](int s, int which) {   // 0 rawA, 1 rawB, 2 loA, 3 loB
    uint8_t* base = smem + (size_t)s * C::STAGE_BYTES;
    switch (which) {
      case 0: return base;
      case 1: return base + C::A_BYTES;
      case 2: return base + C::A_BYTES + C::B_BYTES;
      default: return base + 2 * C::A_BYTES + C::B_BYTES;
    }
  };
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * C::STAGE_BYTES);
  uint64_t* full = bars;                    // [STAGES] TMA -> splitter
  uint64_t* split = bars + STAGES;          // [STAGES] splitter -> MMA
  uint64_t* empty = bars + 2 * STAGES;      // [STAGES] MMA -> TMA
  uint64_t* accum_full = bars + 3 * STAGES; // MMA -> epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM, z = blockIdx.z;
  const int K = p.K0 + p.K1;
  // k-blocks: segment 0 then segment 1, each padded up to a multiple of BK (TMA zero-fills the tail)
  const int nkb0 = (p.K0 + BK - 1) / BK, nkb1 = (p.K1 + BK - 1) / BK;
  const int kb_per_split = p.k_chunk / BK;
  const int kb_begin = z * kb_per_split;
  const int kb_end = min(nkb0 + nkb1, kb_begin + kb_per_split);
  const int num_kb = max(kb_end - kb_begin, 0);
  (void)K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a0);
    tma_prefetch_desc(&map_b);
    if (p.K1 > 0) tma_prefetch_desc(&map_a1);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&split[s], 4);      // one arrive per splitter warp
      mbar_init(&empty[s], 1);
    }
    mbar_init(accum_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {                  // whole warp: TMEM allocation (BN fp32 columns)
    tmem_alloc(tmem_slot, BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], C::A_BYTES + C::B_BYTES);
        const int kb = kb_begin + i;
        const bool seg1 = kb >= nkb0;
        const int ka = seg1 ? (kb - nkb0) * BK : kb * BK;                   // k coordinate inside A's segment
        const int kbcol = seg1 ? p.b_k1_offset + (kb - nkb0) * BK : kb * BK;  // k coordinate in B
        const CUtensorMap* ma = seg1 ? &map_a1 : &map_a0;
        uint8_t* dst_a = stage_ptr(s, 0);
        uint8_t* dst_b = stage_ptr(s, 1);
        if (!C::A_MN) {
          tma_load_2d(dst_a, ma, &full[s], ka, m0);                         // box {BK, 128}
        } else {
#pragma unroll
          for (int c = 0; c < BM / 32; ++c)                                 // box {32, BK} per 32-wide M chunk
            tma_load_2d(dst_a + c * (BK * 128), ma, &full[s], m0 + 32 * c, ka);
        }
        if (!C::B_MN) {
          tma_load_2d(dst_b, &map_b, &full[s], kbcol, n0);                  // box {BK, BN}
        } else {
#pragma unroll
          for (int c = 0; c < BN / 32; ++c)
            tma_load_2d(dst_b + c * (BK * 128), &map_b, &full[s], n0 + 32 * c, kbcol);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = instr_desc_tf32(BM, BN, C::A_MN, C::B_MN);
      // K-major: LBO unused (1), SBO = 8 rows * swizzle span.  MN-major: LBO = chunk stride, SBO = 1024.
      // K-major: rows of K_SWZ bytes, LBO unused (1), SBO = 8 rows.  MN-major fp32/tf32 operands must
      // use the 128B_BASE32B layout (cute: "for mn-major tf32 operands, SW128_32B is the only available
      // smem layout"): 128-byte rows of 32 MN elements, swizzle period 4 k-rows => SBO = 512 B between
      // 4-row groups, LBO = pitch between 32-element MN chunks (BK rows * 128 B).
      constexpr uint64_t a_base = C::A_MN ? smem_desc_base(BK * 128, 512, 1)
                                          : smem_desc_base(16, 8 * C::K_SWZ, C::K_SWZ == 128 ? 2 : 4);
      constexpr uint64_t b_base = C::B_MN ? smem_desc_base(BK * 128, 512, 1)
                                          : smem_desc_base(16, 8 * C::K_SWZ, C::K_SWZ == 128 ? 2 : 4);
      constexpr uint32_t a_kstep = C::A_MN ? 1024 : 32;     // bytes to advance per 8-wide k-slice
      constexpr uint32_t b_kstep = C::B_MN ? 1024 : 32;
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&split[s], ph);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(stage_ptr(s, 0)), b_hi = smem_u32(stage_ptr(s, 1));
        const uint32_t a_lo = smem_u32(stage_ptr(s, 2)), b_lo = smem_u32(stage_ptr(s, 3));
#pragma unroll
        for (int k = 0; k < BK / 8; ++k) {
          const uint64_t da_hi = a_base | uint64_t(((a_hi + k * a_kstep) & 0x3FFFF) >> 4);
          const uint64_t da_lo = a_base | uint64_t(((a_lo + k * a_kstep) & 0x3FFFF) >> 4);
          const uint64_t db_hi = b_base | uint64_t(((b_hi + k * b_kstep) & 0x3FFFF) >> 4);
          const uint64_t db_lo = b_base | uint64_t(((b_lo + k * b_kstep) & 0x3FFFF) >> 4);
          mma_tf32(tmem_base, da_lo, db_hi, idesc, (i | k) != 0);   // small terms first
          mma_tf32(tmem_base, da_hi, db_lo, idesc, 1);
          mma_tf32(tmem_base, da_hi, db_hi, idesc, 1);
        }
        mma_commit(&empty[s]);          // frees the stage once these MMAs have read it
      }
      mma_commit(accum_full);
    }
  } else {
    // ===================================================== splitter (warps 2..5), then epilogue
    const int t = threadIdx.x - 64;     // 0..127
    constexpr int VEC_PER_STAGE = (C::A_BYTES + C::B_BYTES) / 16;
    for (int i = 0; i < num_kb; ++i) {
      const int s = i % STAGES;
      const uint32_t ph = (i / STAGES) & 1;
      mbar_wait(&full[s], ph);
      const float4* raw = reinterpret_cast<const float4*>(stage_ptr(s, 0));      // rawA|rawB contiguous
      float4* lo = reinterpret_cast<float4*>(stage_ptr(s, 2));                   // loA|loB contiguous
#pragma unroll 4
      for (int v = t; v < VEC_PER_STAGE; v += 128) {
        const float4 x = raw[v];
        lo[v] = make_float4(tf32_lo(x.x), tf32_lo(x.y), tf32_lo(x.z), tf32_lo(x.w));
      }
      fence_proxy_async();              // generic-proxy writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&split[s]);
    }
    // ---- epilogue: this warp may touch TMEM lanes 32*(warp%4) .. +31
    mbar_wait(accum_full, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int m = m0 + 32 * q + lane;
    const uint32_t lane_base = tmem_base + (uint32_t(32 * q) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      if (num_kb > 0) {
        tmem_ld32(lane_base + c0, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      if (m < p.M) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int n = n0 + c0 + j;
          if (n < p.N) epi_store<EPI>(epi, p.M, p.N, m, n + p.n_out_offset, __uint_as_float(r[j]), z);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, BN);
}

// ------------------------------------------------------------------ host side
// cuTensorMapEncodeTiled is fetched from the driver at run time (the library does not
// link libcuda, so it also loads on a box without a GPU).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn();

// 2-D fp32 tensor [rows, cols] with row pitch ld (floats, multiple of 4); box = {box_cols, box_rows}.
// swizzle_bytes: 64 / 128 = the classic 16-byte-unit swizzles; 1032 = 128B span with 32-byte units
// (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, the partner of the UMMA 128B_BASE32B layout).
int make_tmap(CUtensorMap* out, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_cols,
              int box_rows, int swizzle_bytes);

// A global operand.  rows/cols describe the tensor as stored (only used for B; A's extents come
// from the Problem): K-major B is [N_total, K_total], MN-major B is [K_total, N_total].
struct Operand {
  const float* ptr;
  long long ld;
  long long rows, cols;
};

// Launch one GEMM.  Returns the effective number of k-splits (> 0) or a negative RECNN_E_* code.
template <bool A_MN, bool B_MN, int EPI>
int launch(const Operand& A0, const Operand& A1, const Operand& B, const Problem& p, int splits, int bn,
           const Epilogue& epi, cudaStream_t st);

}  // namespace tc
}  // namespace recnn
