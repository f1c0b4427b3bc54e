// Frame gather: embedding rows -> state / next_state / action / reward.
// Restates recnn/data/utils.py:51-71 (batch_tensor_embeddings) as one HBM-bound
// copy kernel.  Bit-exact (no arithmetic on the payload).
//
// Layout: one warp per (sample row, group of four slots).  The warp reads the unit's item ids,
// then for each slot j streams the D-float table row with
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
  // Work unit = (row, group of four slots): a warp keeps exactly one batch of four table rows in flight per unit.
  // Units rather than whole rows because 4096 rows over the ~3,500 resident warps of the machine is 1.15 rows per
  // warp -- i.e. TWO rounds for everybody; with three units per row it is 3.5 -> four rounds of a third the length.
  const int groups = (f1 + 3) >> 2;
  const long long n_units = n_rows * groups;

  for (long long unit = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); unit < n_units;
       unit += warps_per_grid) {
    const long long n = unit / groups;
    const int j0 = (int)(unit - n * groups) * 4;             // first slot of this unit
    if (j0 == 0) {
      // ---- ratings tail + reward (tiny): the row's first unit ---------------
      for (int t = lane; t < f1; t += 32) {
        const float r = ratings[n * f1 + t];
        if (t < frame && state) state[n * s_dim + (long long)frame * dim + t] = r;
        if (t >= 1 && next_state) next_state[n * s_dim + (long long)frame * dim + (t - 1)] = r;
        if (t == frame && reward) reward[n] = r;
      }
    }
    // ---- item ids of the unit's (up to) four slots: lanes 0..3 read, everybody gets them by shuffle --------
    {
      const int cnt = min(4, f1 - j0);
      long long my_id = 0;
      if (lane < cnt) {
        my_id = items[n * f1 + j0 + lane];
        if (my_id < 0 || my_id >= n_items) {
          if (oob) atomicOr(oob, 1);
          my_id = 0;
        }
      }
      {
        const int jj = 0;
        long long id[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) id[u] = __shfl_sync(0xffffffffu, my_id, min(u, cnt - 1));
        if (VEC && ST16) {
          for (int c = lane * 4; c < dim; c += 128) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (jj + u < cnt) v[u] = __ldg(reinterpret_cast<const float4*>(table + id[u] * dim + c));
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (jj + u >= cnt) break;
              const int j = j0 + jj + u;
              if (j < frame && state) *reinterpret_cast<float4*>(state + n * s_dim + (long long)j * dim + c) = v[u];
              if (j >= 1 && next_state)
                *reinterpret_cast<float4*>(next_state + n * s_dim + (long long)(j - 1) * dim + c) = v[u];
              if (j == frame && action) {   // lead-padded action rows start 8 bytes into a 16-byte unit
                float2* d = reinterpret_cast<float2*>(action + n * a_ld + c);
                d[0] = make_float2(v[u].x, v[u].y);
                d[1] = make_float2(v[u].z, v[u].w);
              }
            }
          }
        } else if (VEC) {
          // destinations are only 8-byte aligned (dense 1290-float rows): lane l moves the 8 bytes at column 2l of
          // every 64-column span, so each store instruction writes 256 CONTIGUOUS bytes (full 32-byte sectors);
          // 16-byte loads with 8-byte stores would leave every sector half written per instruction
          for (int c = lane * 2; c < dim; c += 128) {
            float2 v[4][2];
            const bool second = c + 64 < dim;
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (jj + u < cnt) {
                v[u][0] = __ldg(reinterpret_cast<const float2*>(table + id[u] * dim + c));
                if (second) v[u][1] = __ldg(reinterpret_cast<const float2*>(table + id[u] * dim + c + 64));
              }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (jj + u >= cnt) break;
              const int j = j0 + jj + u;
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                if (h == 1 && !second) break;
                const int cc = c + 64 * h;
                if (j < frame && state) *reinterpret_cast<float2*>(state + n * s_dim + (long long)j * dim + cc) = v[u][h];
                if (j >= 1 && next_state)
                  *reinterpret_cast<float2*>(next_state + n * s_dim + (long long)(j - 1) * dim + cc) = v[u][h];
                if (j == frame && action) *reinterpret_cast<float2*>(action + n * a_ld + cc) = v[u][h];
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
  const int64_t units = n_rows * ((frame + 1 + 3) / 4);            // (row, four slots) work units, one per warp
  const int64_t blocks = ceil_div(units, warps_per_block);
  const int grid = (int)(blocks < (int64_t)kNumSMs * 16 ? blocks : (int64_t)kNumSMs * 16);
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
