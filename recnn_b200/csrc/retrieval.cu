// Nearest-item retrieval over the embedding table: the serving step that follows Actor.forward in the
// reference (examples/streamlit_demo.py:189-215: faiss IndexFlatL2 / IndexFlatIP / cosine over the item
// matrix; recnn/data/db_con.py:45-56: MilvusConnection.search(search_vecs, topk)).
//
//   scores[q, j] = <query_q, item_j>            one [Q, D] x [D, n_items] contraction on the tensor cores
//                                               (tcgen05 3xTF32, the update step's own GEMM kernel), in slabs
//                                               of query rows so that a slab of scores stays inside the 126 MB L2
//   key[q, j]    = |item_j|^2 - 2 scores        L2   (+ |query_q|^2 at the end: the squared distance faiss/Milvus report)
//                = -scores                      IP   (larger inner product = better)
//                = -scores / |item_j|           COS  (/ |query_q| at the end)
//   top-k smallest keys per query, ties broken towards the smaller item id (== a stable argsort of the keys).
//
// Top-k: every CTA owns one (query, column range); each thread keeps the best k of its strided share in a sorted
// register list (an element is compared with the list's worst first, so the insertion runs ~k ln(n/k) times);
// the CTA then merges its 256 sorted lists by k rounds of a block arg-min over the list heads; a second tiny
// kernel merges the column ranges of a query the same way.  Everything is exact (no approximate search).
#include <float.h>
#include <string.h>

#include "common.cuh"
#include "gemm_simt.cuh"
#include "tc_gemm.cuh"

namespace recnn {

constexpr int kTopkThreads = 256;

// per item: |t|^2 (L2) or 1/|t| (COS)
__global__ void __launch_bounds__(256)
item_norms_kernel(const float* __restrict__ table, long long n_items, int dim, int metric, float* __restrict__ out) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n_items) return;
  const float* t = table + row * dim;
  float s = 0.f;
  for (int d = lane; d < dim; d += 32) s = fmaf(t[d], t[d], s);
  s = warp_sum(s);
  if (lane == 0) out[row] = metric == RECNN_METRIC_COS ? 1.0f / fmaxf(sqrtf(s), 1e-30f) : s;
}

struct Cand {
  float key;
  int id;
};
__device__ __forceinline__ bool better(float ka, int ia, float kb, int ib) { return ka < kb || (ka == kb && ia < ib); }

// block arg-min over one candidate per thread; returns the winner's (key, id, owner thread) to every thread
__device__ __forceinline__ void block_argmin(float key, int id, float* s_key, int* s_id, int* s_owner, float& wk,
                                             int& wi, int& wo) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int owner = threadIdx.x;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float k2 = __shfl_xor_sync(0xffffffffu, key, o);
    const int i2 = __shfl_xor_sync(0xffffffffu, id, o);
    const int o2 = __shfl_xor_sync(0xffffffffu, owner, o);
    if (better(k2, i2, key, id)) { key = k2; id = i2; owner = o2; }
  }
  __syncthreads();
  if (lane == 0) { s_key[warp] = key; s_id[warp] = id; s_owner[warp] = owner; }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    key = lane < nw ? s_key[lane] : FLT_MAX;
    id = lane < nw ? s_id[lane] : 0x7fffffff;
    owner = lane < nw ? s_owner[lane] : -1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float k2 = __shfl_xor_sync(0xffffffffu, key, o);
      const int i2 = __shfl_xor_sync(0xffffffffu, id, o);
      const int o2 = __shfl_xor_sync(0xffffffffu, owner, o);
      if (better(k2, i2, key, id)) { key = k2; id = i2; owner = o2; }
    }
    if (lane == 0) { s_key[0] = key; s_id[0] = id; s_owner[0] = owner; }
  }
  __syncthreads();
  wk = s_key[0]; wi = s_id[0]; wo = s_owner[0];
}

// grid (splits, n_queries).  scores [n_queries, ld]; writes k candidates per (query, split), ascending.
template <int KMAX>
__global__ void __launch_bounds__(kTopkThreads)
topk_partial_kernel(const float* __restrict__ scores, long long ld, long long n_items, const float* __restrict__ norms,
                    int metric, int k, Cand* __restrict__ part) {
  __shared__ float s_key[32];
  __shared__ int s_id[32], s_owner[32];
  const int q = blockIdx.y, sp = blockIdx.x, splits = gridDim.x;
  const long long per = (n_items + splits - 1) / splits;
  const long long lo = sp * per, hi = min(n_items, lo + per);
  const float* row = scores + (long long)q * ld;
  float keys[KMAX];
  int ids[KMAX];
#pragma unroll
  for (int i = 0; i < KMAX; ++i) { keys[i] = FLT_MAX; ids[i] = 0x7fffffff; }
  for (long long j = lo + threadIdx.x; j < hi; j += blockDim.x) {
    const float s = __ldcs(row + j);
    float key = metric == RECNN_METRIC_L2 ? fmaf(-2.0f, s, __ldg(norms + j))
              : metric == RECNN_METRIC_COS ? -s * __ldg(norms + j) : -s;
    if (!(key == key)) key = FLT_MAX;                 // NaN scores rank last
    int id = (int)j;
    if (better(key, id, keys[KMAX - 1], ids[KMAX - 1])) {
      // sorted insertion by a chain of compare-exchanges (fully unrolled: the list stays in registers)
#pragma unroll
      for (int i = 0; i < KMAX; ++i) {
        if (better(key, id, keys[i], ids[i])) {
          const float tk = keys[i]; const int ti = ids[i];
          keys[i] = key; ids[i] = id;
          key = tk; id = ti;
        }
      }
    }
  }
  // merge the 256 sorted lists: k rounds of arg-min over the heads
  int head = 0;
  Cand* out = part + ((long long)q * splits + sp) * k;
  for (int r = 0; r < k; ++r) {
    float hk = FLT_MAX; int hid = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < KMAX; ++i) if (i == head) { hk = keys[i]; hid = ids[i]; }
    float wk; int wi, wo;
    block_argmin(hk, hid, s_key, s_id, s_owner, wk, wi, wo);
    if ((int)threadIdx.x == wo) ++head;
    if (threadIdx.x == 0) { out[r].key = wk; out[r].id = wi; }
  }
}

// one warp per query: merge `splits` ascending lists of k candidates; finish the metric
__global__ void __launch_bounds__(128)
topk_merge_kernel(const Cand* __restrict__ part, int splits, int k, long long n_queries, int metric,
                  const float* __restrict__ queries, int dim, int64_t* __restrict__ ids_out, float* __restrict__ dist_out) {
  const long long q = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= n_queries) return;
  float qn = 0.f;
  for (int d = lane; d < dim; d += 32) { const float v = queries[q * dim + d]; qn = fmaf(v, v, qn); }
  qn = warp_sum(qn);
  const Cand* mine = part + q * (long long)splits * k;
  // lane l walks lists l, l+32, ...: keeps one head per owned list in a small loop (splits <= 32 in practice)
  int head = 0;                                      // lane's position in list `lane` (splits <= 32)
  for (int r = 0; r < k; ++r) {
    float key = FLT_MAX; int id = 0x7fffffff;
    if (lane < splits && head < k) { key = mine[(long long)lane * k + head].key; id = mine[(long long)lane * k + head].id; }
    int owner = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float k2 = __shfl_xor_sync(0xffffffffu, key, o);
      const int i2 = __shfl_xor_sync(0xffffffffu, id, o);
      const int o2 = __shfl_xor_sync(0xffffffffu, owner, o);
      if (better(k2, i2, key, id)) { key = k2; id = i2; owner = o2; }
    }
    if (lane == owner) ++head;
    if (lane == 0) {
      float d;
      if (metric == RECNN_METRIC_L2) d = fmaxf(key + qn, 0.f);                 // squared L2 distance
      else if (metric == RECNN_METRIC_COS) d = -key / fmaxf(sqrtf(qn), 1e-30f);  // cosine similarity
      else d = -key;                                                           // inner product
      ids_out[q * k + r] = id == 0x7fffffff ? -1 : (int64_t)id;
      dist_out[q * k + r] = d;
    }
  }
}

static int topk_splits(int64_t n_queries, int64_t n_items) {
  int64_t s = ceil_div(2 * kNumSMs, n_queries);
  const int64_t max_by_items = ceil_div(n_items, 4 * kTopkThreads);
  if (s > max_by_items) s = max_by_items;
  if (s > 32) s = 32;
  return (int)(s < 1 ? 1 : s);
}
static int64_t slab_rows(int64_t n_items) {
  // a slab of scores should stay L2-resident between the GEMM that writes it and the top-k pass that reads it
  const int64_t ld = round_up(n_items, 4);
  int64_t r = (96ll << 20) / (ld * 4);
  r = r / 128 * 128;
  return r < 128 ? 128 : (r > 4096 ? 4096 : r);
}

}  // namespace recnn

using namespace recnn;

extern "C" int recnn_item_norms(const float* table, int64_t n_items, int32_t dim, int32_t metric, float* out,
                                void* stream) {
  RECNN_REQUIRE(table && out && n_items > 0 && dim > 0, "table/out");
  RECNN_REQUIRE(metric == RECNN_METRIC_L2 || metric == RECNN_METRIC_COS, "norms exist for L2 and COS");
  const int64_t blocks = ceil_div(n_items, 8);
  item_norms_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(table, n_items, dim, metric, out);
  RECNN_CHECK_LAUNCH("item_norms_kernel");
  return RECNN_OK;
}

extern "C" int64_t recnn_retrieve_workspace_bytes(int64_t n_queries, int64_t n_items, int32_t k) {
  if (n_queries <= 0 || n_items <= 0 || k <= 0) return 0;
  const int64_t rows = n_queries < slab_rows(n_items) ? round_up(n_queries, 1) : slab_rows(n_items);
  const int64_t scores = round_up(rows * round_up(n_items, 4) * 4, 256);
  const int64_t part = round_up(rows * 32 * (int64_t)k * (int64_t)sizeof(Cand), 256);
  return scores + part;
}

extern "C" int recnn_retrieve_topk(const float* queries, int64_t n_queries, int32_t dim, const float* table,
                                   int64_t n_items, const float* norms, int32_t metric, int32_t k, int64_t* ids_out,
                                   float* dist_out, void* workspace, int64_t workspace_bytes, void* stream) {
  RECNN_REQUIRE(queries && table && ids_out && dist_out && workspace, "null pointer");
  RECNN_REQUIRE(n_queries >= 0 && n_items > 0 && dim > 0, "sizes");
  RECNN_REQUIRE(metric == RECNN_METRIC_L2 || metric == RECNN_METRIC_IP || metric == RECNN_METRIC_COS, "metric");
  RECNN_REQUIRE(metric == RECNN_METRIC_IP || norms != nullptr, "L2 / COS need recnn_item_norms");
  RECNN_REQUIRE(k >= 1 && k <= 64 && k <= n_items, "1 <= k <= min(64, n_items)");
  RECNN_REQUIRE(n_items < (1ll << 31), "n_items must fit int32");
  if (n_queries == 0) return RECNN_OK;
  RECNN_REQUIRE(workspace_bytes >= recnn_retrieve_workspace_bytes(n_queries, n_items, k), "workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t ld = round_up(n_items, 4);
  const int64_t slab = n_queries < slab_rows(n_items) ? n_queries : slab_rows(n_items);
  float* scores = static_cast<float*>(workspace);
  Cand* part = reinterpret_cast<Cand*>(static_cast<char*>(workspace) + round_up(slab * ld * 4, 256));
  const bool tc_ok = dim % 4 == 0 && (reinterpret_cast<uintptr_t>(queries) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(table) & 15) == 0;
  for (int64_t q0 = 0; q0 < n_queries; q0 += slab) {
    const int64_t nq = n_queries - q0 < slab ? n_queries - q0 : slab;
    const float* Q = queries + q0 * dim;
    Epilogue e;
    memset(&e, 0, sizeof(e));
    e.out = scores;
    e.ldo = ld;
    if (tc_ok) {
      tc::Operand a0 = {Q, dim, 0, 0}, a1 = {nullptr, 0, 0, 0}, b = {table, dim, n_items, dim};
      tc::Problem p;
      memset(&p, 0, sizeof(p));
      p.M = (int)nq; p.N = (int)n_items; p.K0 = dim; p.b_k1_offset = dim;
      const int r = tc::launch<false, false, EPI_STORE>(a0, a1, b, p, 1, 128, e, st);
      if (r < 0) return r;
    } else {
      RECNN_PROPAGATE((launch_gemm_simt<true, true, EPI_STORE>(mat(Q, dim), mat(table, dim), (int)nq, (int)n_items, dim,
                                                                1, e, st)));
    }
    const int splits = topk_splits(nq, n_items);
    dim3 grid((unsigned)splits, (unsigned)nq);
    if (k <= 16)
      topk_partial_kernel<16><<<grid, kTopkThreads, 0, st>>>(scores, ld, n_items, norms, metric, k, part);
    else
      topk_partial_kernel<64><<<grid, kTopkThreads, 0, st>>>(scores, ld, n_items, norms, metric, k, part);
    RECNN_CHECK_LAUNCH("topk_partial_kernel");
    topk_merge_kernel<<<(unsigned)ceil_div(nq, 4), 128, 0, st>>>(part, splits, k, nq, metric, Q, dim, ids_out + q0 * k,
                                                                  dist_out + q0 * k);
    RECNN_CHECK_LAUNCH("topk_merge_kernel");
  }
  return RECNN_OK;
}
