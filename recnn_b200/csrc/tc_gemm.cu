// Host side of the tcgen05 3xTF32 GEMM: tensor-map construction, launch
// configuration, and the generic C-ABI entry points recnn_gemm_tf32x3 / recnn_gemm_fp32.
#include "tc_gemm.cuh"

#include <stdlib.h>
#include <string.h>

#include <mutex>

namespace recnn {
namespace tc {

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap(CUtensorMap* out, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_cols,
              int box_rows, int swizzle_bytes) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return RECNN_E_CUDA;
  }
  RECNN_REQUIRE(reinterpret_cast<uintptr_t>(base) % 16 == 0, "TMA base must be 16-byte aligned");
  RECNN_REQUIRE(ld % 4 == 0 && ld >= cols, "TMA row pitch must be a multiple of 4 floats");
  RECNN_REQUIRE(box_cols * 4 <= (swizzle_bytes == 1032 ? 128 : swizzle_bytes) && box_rows <= 256, "TMA box");
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = swizzle_bytes == 1032 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                               : swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                               : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                     : CU_TENSOR_MAP_SWIZZLE_32B;
  const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld box=%dx%d", (int)r, (long long)rows,
              (long long)cols, (long long)ld, box_cols, box_rows);
    return RECNN_E_CUDA;
  }
  return RECNN_OK;
}

int split_plan(int K_total_blocks, int splits_req, int* k_chunk, int bk) {
  if (splits_req < 1) splits_req = 1;
  const int kb_per = (int)ceil_div(K_total_blocks, splits_req);
  *k_chunk = kb_per * bk;
  return (int)ceil_div(K_total_blocks, kb_per);
}

// Step-timeline tracing (debugging aid): when a span buffer is installed, launch number i of this
// process writes {first CTA entry, last CTA exit} to span[2*i ..] (shape/config via recnn_debug_span_meta); the slot index is baked into the kernel parameters, so CUDA-graph replays refresh it.
static unsigned long long* g_span = nullptr;
static int g_span_cap = 0, g_span_next = 0;
static long long g_span_meta[4096][4];
extern "C" RECNN_API void recnn_debug_set_span(unsigned long long* buf, int capacity) {
  g_span = buf; g_span_cap = capacity < 4096 ? capacity : 4096; g_span_next = 0;
}
extern "C" RECNN_API int recnn_debug_span_count(void) { return g_span_next; }
extern "C" RECNN_API void recnn_debug_span_meta(int i, long long* out4) {
  for (int j = 0; j < 4; ++j) out4[j] = g_span_meta[i][j];
}

template <class C, int EPI>
static int launch_cfg(const Operand& A0, const Operand& A1, const Operand& B, const Problem& p_in, int splits,
                      const Epilogue& epi, cudaStream_t st) {
  Problem p = p_in;
  static bool attr_set = false;
  if (!attr_set) {
    RECNN_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<C, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          C::SMEM_BYTES));
    attr_set = true;
  }
  const int nkb = (int)(ceil_div(p.K0, C::BK) + ceil_div(p.K1, C::BK));
  splits = split_plan(nkb, splits, &p.k_chunk, C::BK);
  CUtensorMap ma0, ma1, mb, mb_lo;
  // K-major operand: tensor [rows = M|N, cols = K], box {BK, tile rows}; MN-major: tensor [rows = K, cols = M|N], box {32, BK}
  if (!C::A_MN) {
    RECNN_PROPAGATE(make_tmap(&ma0, A0.ptr, p.M, p.K0, A0.ld, C::BK, C::BM, C::K_SWZ));
    if (p.K1 > 0) RECNN_PROPAGATE(make_tmap(&ma1, A1.ptr, p.M, p.K1, A1.ld, C::BK, C::BM, C::K_SWZ));
    else ma1 = ma0;
  } else {
    RECNN_REQUIRE(p.K1 == 0, "MN-major A cannot be a K-concat");
    RECNN_PROPAGATE(make_tmap(&ma0, A0.ptr, p.K0, p.M, A0.ld, 32, C::BK, 1032));
    ma1 = ma0;
  }
  if (!C::B_MN) RECNN_PROPAGATE(make_tmap(&mb, B.ptr, B.rows, B.cols, B.ld, C::BK, C::BN, C::K_SWZ));
  else RECNN_PROPAGATE(make_tmap(&mb, B.ptr, B.rows, B.cols, B.ld, 32, C::BK, 1032));
  if (C::B_PRE) {
    RECNN_REQUIRE(B.lo != nullptr, "pre-split B needs its lo plane");
    if (!C::B_MN) RECNN_PROPAGATE(make_tmap(&mb_lo, B.lo, B.rows, B.cols, B.ld, C::BK, C::BN, C::K_SWZ));
    else RECNN_PROPAGATE(make_tmap(&mb_lo, B.lo, B.rows, B.cols, B.ld, 32, C::BK, 1032));
  } else {
    mb_lo = mb;
  }
  dim3 grid((unsigned)ceil_div(p.N, C::BN), (unsigned)ceil_div(p.M, C::BM), (unsigned)splits);
  if (g_span && g_span_next < g_span_cap) {
    g_span_meta[g_span_next][0] = C::BN | (C::A_MN << 12) | (C::B_MN << 13) | (C::B_PRE << 14) | ((C::WORKERS == 16) << 15) | (EPI << 16) | (C::LO2 << 20);
    g_span_meta[g_span_next][1] = p.M;
    g_span_meta[g_span_next][2] = p.N;
    g_span_meta[g_span_next][3] = (long long)(p.K0 + p.K1) | ((long long)splits << 32);
    p.span = g_span + 2 * g_span_next++;
  } else {
    p.span = nullptr;
  }
  static const bool debug = getenv("RECNN_B200_DEBUG") != nullptr;
  if (debug)
    fprintf(stderr, "[tc_gemm] BN=%d A_MN=%d B_MN=%d EPI=%d M=%d N=%d K0=%d K1=%d k_chunk=%d bk1=%d nout=%d bn_off=%d "
            "grid=%u,%u,%u a0=%p ld=%lld a1=%p ld=%lld b=%p ld=%lld rows=%lld cols=%lld out=%p ldo=%lld\n",
            C::BN, (int)C::A_MN, (int)C::B_MN, EPI, p.M, p.N, p.K0, p.K1, p.k_chunk, p.b_k1_offset, p.n_out_offset,
            p.b_n_offset, grid.x, grid.y, grid.z, (const void*)A0.ptr, A0.ld, (const void*)A1.ptr, A1.ld,
            (const void*)B.ptr, B.ld, B.rows, B.cols, (void*)epi.out, epi.ldo);
  if (C::LEAN && option(OPT_PDL) != 0) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = dim3(C::THREADS, 1, 1);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    RECNN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, tc_gemm_kernel<C, EPI>, ma0, ma1, mb, mb_lo, p, epi));
  } else {
    tc_gemm_kernel<C, EPI><<<grid, C::THREADS, C::SMEM_BYTES, st>>>(ma0, ma1, mb, mb_lo, p, epi);
  }
  RECNN_CHECK_LAUNCH("tc_gemm_kernel");
  if (debug) {
    const cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      set_error("tc_gemm_kernel execution failed: %s", cudaGetErrorString(e));
      fprintf(stderr, "[tc_gemm] FAILED: %s\n", cudaGetErrorString(e));
      return RECNN_E_CUDA;
    }
  }
  return splits;
}

template <bool A_MN, bool B_MN, int EPI>
int launch(const Operand& A0, const Operand& A1, const Operand& B, const Problem& p, int splits, int bn,
           const Epilogue& epi, cudaStream_t st) {
  // stages chosen to fill ~190 KB of shared memory
  // BK = 32 (128-byte TMA rows), A split into tensor memory: stage = 16 KB (raw A) + 2 * BN/8 KB (B hi | B lo)
  // Pre-split B (weights as hi/lo planes) exists for the forward and input-gradient GEMMs: A K-major,
  // every epilogue but the split-K partial.
  if constexpr (!A_MN && EPI != EPI_PARTIAL) {
    if (B.lo != nullptr) {
      if (option(OPT_LEAN) != 0) {             // unvalidated combinations for the round-2 A/B (see Cfg::LEAN)
        if (bn >= 128) return launch_cfg<Cfg<128, 32, 4, A_MN, B_MN, true, 8, false, true>, EPI>(A0, A1, B, p, splits, epi, st);
        if (option(OPT_WORKERS16) != 0)
          return launch_cfg<Cfg<64, 32, 6, A_MN, B_MN, true, 16, false, true>, EPI>(A0, A1, B, p, splits, epi, st);
        return launch_cfg<Cfg<64, 32, 6, A_MN, B_MN, true, 8, false, true>, EPI>(A0, A1, B, p, splits, epi, st);
      }
      if (bn >= 128) return launch_cfg<Cfg<128, 32, 4, A_MN, B_MN, true>, EPI>(A0, A1, B, p, splits, epi, st);
      return launch_cfg<Cfg<64, 32, 6, A_MN, B_MN, true>, EPI>(A0, A1, B, p, splits, epi, st);
    }
  }
  RECNN_REQUIRE(B.lo == nullptr, "pre-split B is not available for this GEMM form");
  const bool lean = option(OPT_LEAN) != 0;
  if (bn < 128) {
    const bool w16 = option(OPT_WORKERS16) != 0, lo2 = option(OPT_LO2) != 0;
    if (lo2) return launch_cfg<Cfg<64, 32, 6, A_MN, B_MN, false, 8, true>, EPI>(A0, A1, B, p, splits, epi, st);
    if (w16 && lean) return launch_cfg<Cfg<64, 32, 6, A_MN, B_MN, false, 16, false, true>, EPI>(A0, A1, B, p, splits, epi, st);
    if (w16) return launch_cfg<Cfg<64, 32, 6, A_MN, B_MN, false, 16>, EPI>(A0, A1, B, p, splits, epi, st);
    if (lean) return launch_cfg<Cfg<64, 32, 6, A_MN, B_MN, false, 8, false, true>, EPI>(A0, A1, B, p, splits, epi, st);
  } else if (lean) {
    return launch_cfg<Cfg<128, 32, 4, A_MN, B_MN, false, 8, false, true>, EPI>(A0, A1, B, p, splits, epi, st);
  }
  if (bn >= 128) return launch_cfg<Cfg<128, 32, 4, A_MN, B_MN>, EPI>(A0, A1, B, p, splits, epi, st);
  return launch_cfg<Cfg<64, 32, 6, A_MN, B_MN>, EPI>(A0, A1, B, p, splits, epi, st);
}

// ---- TF32 hi/lo planes of whole arrays (weights) -------------------------------------------------
struct SplitJobs {
  SplitJob j[8];
};
__global__ void __launch_bounds__(256) split_planes_kernel(SplitJobs jobs) {
  const SplitJob job = jobs.j[blockIdx.y];
  const long long n4 = job.count >> 2;
  const float4* __restrict__ src = reinterpret_cast<const float4*>(job.src);
  float4* __restrict__ hi = reinterpret_cast<float4*>(job.hi);
  float4* __restrict__ lo = reinterpret_cast<float4*>(job.lo);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 h, l;
    tf32_split4(src[i], h, l);
    hi[i] = h;
    lo[i] = l;
  }
}

int launch_split_planes(const SplitJob* jobs, int n_jobs, cudaStream_t st) {
  RECNN_REQUIRE(jobs && n_jobs >= 0 && n_jobs <= 8, "at most 8 arrays per launch");
  if (n_jobs == 0) return RECNN_OK;
  SplitJobs js;
  long long longest = 0;
  for (int i = 0; i < n_jobs; ++i) {
    const SplitJob& j = jobs[i];
    RECNN_REQUIRE(j.src && j.hi && j.lo && j.count > 0 && j.count % 4 == 0, "split job");
    RECNN_REQUIRE(((reinterpret_cast<uintptr_t>(j.src) | reinterpret_cast<uintptr_t>(j.hi) |
                    reinterpret_cast<uintptr_t>(j.lo)) & 15) == 0, "split planes must be 16-byte aligned");
    js.j[i] = j;
    if (j.count > longest) longest = j.count;
  }
  for (int i = n_jobs; i < 8; ++i) js.j[i] = SplitJob{nullptr, nullptr, nullptr, 0};
  // 1.7 MB per arena: two float4 per thread, so the launch is a single wave however many arenas it covers
  const long long blocks = ceil_div(longest >> 2, 256 * 2);
  dim3 grid((unsigned)(blocks < 1 ? 1 : blocks), (unsigned)n_jobs);
  split_planes_kernel<<<grid, 256, 0, st>>>(js);
  RECNN_CHECK_LAUNCH("split_planes_kernel");
  return RECNN_OK;
}

// explicit instantiations used by step.cu / the generic entry points
#define RECNN_TC_INST(A_MN, B_MN, EPI) \
  template int launch<A_MN, B_MN, EPI>(const Operand&, const Operand&, const Operand&, const Problem&, int, int, \
                                       const Epilogue&, cudaStream_t);
RECNN_TC_INST(false, false, EPI_HIDDEN)
RECNN_TC_INST(false, false, EPI_LINEAR)
RECNN_TC_INST(false, false, EPI_STORE)
RECNN_TC_INST(false, true, EPI_STORE)
RECNN_TC_INST(false, true, EPI_GATE)
RECNN_TC_INST(true, true, EPI_STORE)
RECNN_TC_INST(true, true, EPI_PARTIAL)
RECNN_TC_INST(false, false, EPI_PARTIAL)
RECNN_TC_INST(true, false, EPI_STORE)

}  // namespace tc
}  // namespace recnn

using namespace recnn;

static unsigned long long* g_trace = nullptr;
// debugging aid: device buffer (8 u64 per CTA) that recnn_gemm_tf32x3 fills with %globaltimer stamps
extern "C" RECNN_API void recnn_debug_set_trace(unsigned long long* buf) { g_trace = buf; }

// C[M,N] (row pitch ldc) = A . B^T in 3xTF32 on the tensor cores.
//   a_mn = 0: A is [M,K] row-major (pitch lda)   a_mn = 1: A is [K,M] row-major
//   b_mn = 0: B is [N,K] row-major (pitch ldb)   b_mn = 1: B is [K,N] row-major
// All pitches and base addresses must be multiples of 4 floats / 16 bytes (TMA).
extern "C" int recnn_gemm_tf32x3(int M, int N, int K, const float* A, int64_t lda, int a_mn, const float* B,
                                 int64_t ldb, int b_mn, float* C, int64_t ldc, int tile_n, void* stream) {
  RECNN_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0, "null pointer / sizes");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Epilogue e;
  memset(&e, 0, sizeof(e));
  e.out = C;
  e.ldo = ldc;
  tc::Operand a0 = {A, lda, 0, 0}, a1 = {nullptr, 0, 0, 0};
  tc::Operand b = {B, ldb, b_mn ? K : N, b_mn ? N : K};
  tc::Problem p = {M, N, K, 0, 0, K, 0, 0, 0, getenv("RECNN_TC_DBG") ? atoi(getenv("RECNN_TC_DBG")) : 0, g_trace, nullptr};
  if (tile_n <= 0) tile_n = N > 64 ? 128 : 64;
  int r;
  if (!a_mn && !b_mn) r = tc::launch<false, false, EPI_STORE>(a0, a1, b, p, 1, tile_n, e, st);
  else if (!a_mn && b_mn) r = tc::launch<false, true, EPI_STORE>(a0, a1, b, p, 1, tile_n, e, st);
  else if (a_mn && b_mn) r = tc::launch<true, true, EPI_STORE>(a0, a1, b, p, 1, tile_n, e, st);
  else r = tc::launch<true, false, EPI_STORE>(a0, a1, b, p, 1, tile_n, e, st);
  return r < 0 ? r : RECNN_OK;
}

// Test hook (not part of the product ABI): the same GEMM with B first split into hi/lo planes (caller
// scratch, same geometry as B) and consumed by the B_PRE kernels.  A must be K-major.  Results must be
// bit-identical to recnn_gemm_tf32x3.
extern "C" RECNN_API int recnn_debug_gemm_tf32x3_presplit(int M, int N, int K, const float* A, int64_t lda,
                                                          const float* B, int64_t ldb, int b_mn, float* C,
                                                          int64_t ldc, int tile_n, float* b_hi, float* b_lo,
                                                          void* stream) {
  RECNN_REQUIRE(A && C && b_hi && b_lo && M > 0 && N > 0 && K > 0, "null pointer / sizes");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long b_rows = b_mn ? K : N;
  if (B) {                                       // B == NULL: the planes are already filled (timing runs)
    const tc::SplitJob job = {B, b_hi, b_lo, b_rows * ldb};
    RECNN_PROPAGATE(tc::launch_split_planes(&job, 1, st));
  }
  Epilogue e;
  memset(&e, 0, sizeof(e));
  e.out = C;
  e.ldo = ldc;
  tc::Operand a0 = {A, lda, 0, 0}, a1 = {nullptr, 0, 0, 0};
  tc::Operand b = {b_hi, ldb, b_mn ? K : N, b_mn ? N : K, b_lo};
  tc::Problem p = {M, N, K, 0, 0, K, 0, 0, 0, 0, nullptr, nullptr};
  if (tile_n <= 0) tile_n = N > 64 ? 128 : 64;
  const int r = b_mn ? tc::launch<false, true, EPI_STORE>(a0, a1, b, p, 1, tile_n, e, st)
                     : tc::launch<false, false, EPI_STORE>(a0, a1, b, p, 1, tile_n, e, st);
  return r < 0 ? r : RECNN_OK;
}

// Same contract on the fp32 CUDA cores (exact fp32 FMA chain) -- the arbitrary-shape path.
extern "C" int recnn_gemm_fp32(int M, int N, int K, const float* A, int64_t lda, int a_mn, const float* B,
                               int64_t ldb, int b_mn, float* C, int64_t ldc, void* stream) {
  RECNN_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0, "null pointer / sizes");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Epilogue e;
  memset(&e, 0, sizeof(e));
  e.out = C;
  e.ldo = ldc;
  const MatView a = mat(A, lda), b = mat(B, ldb);
  if (!a_mn && !b_mn) return launch_gemm_simt<true, true, EPI_STORE>(a, b, M, N, K, 1, e, st);
  if (!a_mn && b_mn) return launch_gemm_simt<true, false, EPI_STORE>(a, b, M, N, K, 1, e, st);
  if (a_mn && b_mn) return launch_gemm_simt<false, false, EPI_STORE>(a, b, M, N, K, 1, e, st);
  return launch_gemm_simt<false, true, EPI_STORE>(a, b, M, N, K, 1, e, st);
}
