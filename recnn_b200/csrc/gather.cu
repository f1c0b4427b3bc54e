// Frame gather: embedding rows -> state / next_state / action / reward.
// Restates recnn/data/utils.py:51-71 (batch_tensor_embeddings) as one HBM-bound
// copy kernel.  Bit-exact (no arithmetic on the payload).
//
// Layout: one warp per sample row.  The warp reads the row's F+1 item ids with
// one coalesced load, then for each slot j streams the D-float table row with
// 16-byte loads (table rows are D*4-byte aligned) and writes it to
//   state[n, j*D ...]        if j <  F
//   next_state[n, (j-1)*D..] if j >= 1
//   action[n, ...]           if j == F
// Rows of state/next_state are (F*D+F)*4 bytes long, which is only 8-byte
// aligned (1290*4 = 5160 = 8 mod 16), so stores are 8-byte vectors.
// Slots are processed four at a time: all four table reads are issued before
// the first store so each lane keeps 4 independent 16-byte loads in flight.
//
// Algorithmic bytes per row (D=128, F=10), SURVEY.md 8d: read 5,764 B
// (11 table rows + 11 ids + 11 ratings), write 10,840 B  => 16,604 B/row.
#include "common.cuh"

namespace recnn {

// ST16: the destination pitches are multiples of 4 floats and the bases 16-byte aligned (the step's padded images,
// pitch 1292), so a lane's 4 floats go out as ONE 16-byte store instead of two 8-byte ones.
template <bool VEC, bool ST16 = false>
__global__ void __launch_bounds__(256)
frame_gather_kernel(const float* __restrict__ table, long long n_items, int dim,
                    const long long* __restrict__ items, const float* __restrict__ ratings,
                    long long n_rows, int frame, long long s_ld, long long a_ld,
                    float* __restrict__ state, float* __restrict__ next_state,
                    float* __restrict__ action, float* __restrict__ reward, int* __restrict__ oob) {
  const int lane = threadIdx.x & 31;
  const long long warps_per_grid = (long long)gridDim.x * (blockDim.x >> 5);
  const int f1 = frame + 1;
  const long long s_dim = s_ld;       // row pitch of state / next_state (>= frame*dim + frame)

  for (long long n = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); n < n_rows;
       n += warps_per_grid) {
    // ---- ratings tail + reward (tiny) --------------------------------------
    for (int t = lane; t < f1; t += 32) {
      const float r = ratings[n * f1 + t];
      if (t < frame && state) state[n * s_dim + (long long)frame * dim + t] = r;
      if (t >= 1 && next_state) next_state[n * s_dim + (long long)frame * dim + (t - 1)] = r;
      if (t == frame && reward) reward[n] = r;
    }
    // ---- item ids: one coalesced read, broadcast by shuffle ------------------
    for (int j0 = 0; j0 < f1; j0 += 32) {
      long long my_id = 0;
      if (j0 + lane < f1) {
        my_id = items[n * f1 + j0 + lane];
        if (my_id < 0 || my_id >= n_items) {
          if (oob) atomicOr(oob, 1);
          my_id = 0;
        }
      }
      const int cnt = min(32, f1 - j0);
      for (int jj = 0; jj < cnt; jj += 4) {
        long long id[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) id[u] = __shfl_sync(0xffffffffu, my_id, min(jj + u, cnt - 1));
        if (VEC) {
          for (int c = lane * 4; c < dim; c += 128) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (jj + u < cnt) v[u] = __ldg(reinterpret_cast<const float4*>(table + id[u] * dim + c));
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (jj + u >= cnt) break;
              const int j = j0 + jj + u;
              const float2 lo = make_float2(v[u].x, v[u].y), hi = make_float2(v[u].z, v[u].w);
              if (j < frame && state) {
                float* dst = state + n * s_dim + (long long)j * dim + c;
                if constexpr (ST16) *reinterpret_cast<float4*>(dst) = v[u];
                else { float2* d = reinterpret_cast<float2*>(dst); d[0] = lo; d[1] = hi; }
              }
              if (j >= 1 && next_state) {
                float* dst = next_state + n * s_dim + (long long)(j - 1) * dim + c;
                if constexpr (ST16) *reinterpret_cast<float4*>(dst) = v[u];
                else { float2* d = reinterpret_cast<float2*>(dst); d[0] = lo; d[1] = hi; }
              }
              if (j == frame && action) {
                float* dst = action + n * a_ld + c;     // lead-padded action rows start 8 bytes into a 16-byte unit
                float2* d = reinterpret_cast<float2*>(dst); d[0] = lo; d[1] = hi;
              }
            }
          }
        } else {
          for (int c = lane; c < dim; c += 32) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (jj + u >= cnt) break;
              const int j = j0 + jj + u;
              const float v = __ldg(table + id[u] * dim + c);
              if (j < frame && state) state[n * s_dim + (long long)j * dim + c] = v;
              if (j >= 1 && next_state) next_state[n * s_dim + (long long)(j - 1) * dim + c] = v;
              if (j == frame && action) action[n * a_ld + c] = v;
            }
          }
        }
      }
    }
  }
}

__global__ void done_from_sizes_kernel(const long long* __restrict__ sizes, long long n_users, int frame,
                                       float* __restrict__ done, long long n_rows) {
  // done[cumsum(sizes - frame) - 1] = 1.  n_users is small (25 in the reference's
  // default batch): a single thread walks the prefix sum; the zero-fill was a memset.
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    long long acc = 0;
    for (long long u = 0; u < n_users; ++u) {
      acc += sizes[u] - frame;
      long long pos = acc - 1;
      if (pos < 0) pos += n_rows;              // torch negative-index semantics
      if (pos >= 0 && pos < n_rows) done[pos] = 1.0f;
    }
  }
}

}  // namespace recnn

using namespace recnn;

namespace recnn {
int launch_frame_gather(const float* table, int64_t n_items, int dim, const int64_t* items, const float* ratings,
                        int64_t n_rows, int frame, int64_t s_ld, int64_t a_ld, float* state, float* next_state,
                        float* action, float* reward, int* oob_flag, cudaStream_t st) {
  RECNN_REQUIRE(table && items && ratings, "table/items/ratings must be non-null");
  RECNN_REQUIRE(n_items > 0 && dim > 0 && frame > 0 && n_rows >= 0, "sizes must be positive");
  if (n_rows == 0) return RECNN_OK;
  const long long s_dim = s_ld;
  RECNN_REQUIRE(s_ld >= (int64_t)frame * dim + frame, "state pitch too small");
  const bool aligned = (reinterpret_cast<uintptr_t>(table) % 16 == 0) &&
                       (!state || reinterpret_cast<uintptr_t>(state) % 8 == 0) &&
                       (!next_state || reinterpret_cast<uintptr_t>(next_state) % 8 == 0) &&
                       (!action || reinterpret_cast<uintptr_t>(action) % 8 == 0);
  const bool vec = aligned && (dim % 4 == 0) && (s_dim % 2 == 0) && (a_ld % 2 == 0);
  const int warps_per_block = 8;
  const int64_t blocks = ceil_div(n_rows, warps_per_block);
  const int grid = (int)(blocks < (int64_t)kNumSMs * 8 ? blocks : (int64_t)kNumSMs * 8);
  // 16-byte stores whenever the destination pitches allow it (the step's state images have a 16-byte-multiple
  // pitch; the public API's dense 1290-float rows are only 8-byte aligned)
  const bool st16 = vec && s_ld % 4 == 0 && (((long long)frame * dim) % 4 == 0) &&
                    (!state || reinterpret_cast<uintptr_t>(state) % 16 == 0) &&
                    (!next_state || reinterpret_cast<uintptr_t>(next_state) % 16 == 0);
  if (st16)
    frame_gather_kernel<true, true><<<grid, 256, 0, st>>>(table, n_items, dim, (const long long*)items, ratings,
                                                          n_rows, frame, s_ld, a_ld, state, next_state, action, reward, oob_flag);
  else if (vec)
    frame_gather_kernel<true><<<grid, 256, 0, st>>>(table, n_items, dim, (const long long*)items, ratings,
                                                    n_rows, frame, s_ld, a_ld, state, next_state, action, reward, oob_flag);
  else
    frame_gather_kernel<false><<<grid, 256, 0, st>>>(table, n_items, dim, (const long long*)items, ratings,
                                                     n_rows, frame, s_ld, a_ld, state, next_state, action, reward, oob_flag);
  RECNN_CHECK_LAUNCH("frame_gather_kernel");
  return RECNN_OK;
}
}  // namespace recnn

extern "C" int recnn_frame_gather(const float* table, int64_t n_items, int dim, const int64_t* items,
                                  const float* ratings, int64_t n_rows, int frame, float* state,
                                  float* next_state, float* action, float* reward, int* oob_flag,
                                  void* stream) {
  return recnn::launch_frame_gather(table, n_items, dim, items, ratings, n_rows, frame,
                                    (int64_t)frame * dim + frame, dim, state, next_state, action, reward, oob_flag,
                                    static_cast<cudaStream_t>(stream));
}

extern "C" int recnn_done_from_sizes(const int64_t* sizes, int64_t n_users, int frame, float* done,
                                     int64_t n_rows, void* stream) {
  RECNN_REQUIRE(sizes && done, "sizes/done must be non-null");
  RECNN_REQUIRE(n_users >= 0 && n_rows >= 0 && frame > 0, "sizes must be non-negative");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_rows == 0) return RECNN_OK;
  RECNN_CHECK_CUDA(cudaMemsetAsync(done, 0, sizeof(float) * n_rows, st));
  done_from_sizes_kernel<<<1, 32, 0, st>>>((const long long*)sizes, n_users, frame, done, n_rows);
  RECNN_CHECK_LAUNCH("done_from_sizes_kernel");
  return RECNN_OK;
}
