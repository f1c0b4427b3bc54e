// Peer-memory all-reduce of the gradient arenas (data-parallel update step, one process per GPU).
//
// The reference is single-process; BASELINE north_star asks for the minibatch to be sharded over the
// GPUs of one box with an all-reduce of the Actor/Critic gradients over NVLink.  Calling NCCL between
// the phases of the step costs three host-launched collectives and cuts the step's CUDA graph in three;
// instead every rank maps its peers' staging buffers (cudaIpc, NVLink/NVSwitch peer access) and ONE
// kernel per gradient arena does
//     copy my gradient -> my staging slot | signal peers | wait for peers | sum all ranks' slots in
//     rank order -> my gradient (in place)  [+ the L1 norm of the summed gradient for the reference's
//     clip_grad_norm_(params, -1, 1) quirk]
// It is an ordinary kernel on the step's stream, so the whole data-parallel step is captured in one
// CUDA graph exactly like the single-GPU step.  The sum is taken in rank order 0..W-1 on every rank,
// so all replicas hold bit-identical gradients (and therefore weights) after every step.
//
// Protocol (epoch e = number of collectives issued so far on this communicator + 1; all on one stream):
//   * staging is double buffered by e & 1.  A rank may overwrite slot e & 1 only when every peer has
//     finished reading the data of epoch e-2; a peer signals epoch e-1 only after its epoch e-2 kernel
//     has completed (stream order), and this rank waited for all epoch e-1 signals before finishing e-1.
//   * signal: after all CTAs copied (fence + counter), the last CTA stores e into flags[my_rank] in
//     EVERY peer's memory (remote store), so the waiting side polls its own HBM.
//   * waits are bounded (~20 s of %globaltimer): a lost peer makes the kernel trap instead of hanging the GPU.
#include <string.h>

#include "common.cuh"
#include "pointwise.cuh"

namespace recnn {

constexpr int kMaxRanks = 8;
constexpr int kCommThreads = 512;

struct CommDev {                       // lives at the head of every rank's shared allocation
  unsigned flags[kMaxRanks];           // flags[src] = last epoch src has published  (written by peers)
  unsigned epoch;                      // collectives completed by this rank           (local)
  unsigned arrive;                     // CTAs that finished copying (wraps to 0)        (local)
  unsigned done;                       // CTAs that finished reducing (wraps to 0)       (local)
  unsigned pad[32 - kMaxRanks - 3];
};
static_assert(sizeof(CommDev) == 128, "CommDev");

struct CommPeers {                     // kernel parameter
  CommDev* ctrl[kMaxRanks];
  float* stage[kMaxRanks];             // 2 slots of `capacity` floats each
  long long capacity;
  int rank, world;
};

}  // namespace recnn

struct recnn_comm {
  recnn::CommPeers peers;
  void* local_base;
  void* opened[recnn::kMaxRanks];
  cudaIpcMemHandle_t handle;
  bool connected;
};

namespace recnn {

__device__ __forceinline__ unsigned long long comm_gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// grid <= number of SMs (all CTAs must be co-resident: they wait for each other's peers)
__global__ void __launch_bounds__(kCommThreads, 1)
allreduce_kernel(CommPeers c, float* __restrict__ buf, long long n, float max_norm, float* coef, float* l1_out,
                 float* block_partials) {
  __shared__ float red[32];
  __shared__ unsigned s_epoch;
  CommDev* me = c.ctrl[c.rank];
  if (threadIdx.x == 0) s_epoch = *((volatile unsigned*)&me->epoch) + 1;
  __syncthreads();
  const unsigned e = s_epoch;
  const long long slot = (long long)(e & 1u) * c.capacity;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  const bool vec = (n & 3) == 0 && ((reinterpret_cast<uintptr_t>(buf) & 15) == 0);

  // ---- A: publish my contribution
  float* mine = c.stage[c.rank] + slot;
  if (vec) {
    const float4* src = reinterpret_cast<const float4*>(buf);
    float4* dst = reinterpret_cast<float4*>(mine);
    for (long long i = tid; i < n / 4; i += nth) dst[i] = src[i];
  } else {
    for (long long i = tid; i < n; i += nth) mine[i] = buf[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicInc(&me->arrive, gridDim.x - 1);
    if (t == gridDim.x - 1) {                       // every CTA's slice is visible system-wide
      __threadfence_system();
      for (int p = 0; p < c.world; ++p) st_release_sys(&c.ctrl[p]->flags[c.rank], e);
    }
  }
  // ---- B: wait for every rank's contribution
  if (threadIdx.x < c.world) {
    const unsigned long long t0 = comm_gtimer();
    // epochs only grow; compare as a signed distance so that a 32-bit wrap is harmless
    while ((int)(ld_acquire_sys(&me->flags[threadIdx.x]) - e) < 0) {
      if (comm_gtimer() - t0 > 20000000000ull) {
        printf("recnn_b200 allreduce: rank %d timed out waiting for rank %d (epoch %u)\n", c.rank, (int)threadIdx.x, e);
        __trap();
      }
    }
  }
  __syncthreads();
  // ---- C: sum in rank order (bit-identical on every rank), in place
  float l1 = 0.f;
  if (vec) {
    float4* out = reinterpret_cast<float4*>(buf);
    for (long long i = tid; i < n / 4; i += nth) {
      float4 s = __ldcg(reinterpret_cast<const float4*>(c.stage[0] + slot) + i);
      for (int p = 1; p < c.world; ++p) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(c.stage[p] + slot) + i);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
      out[i] = s;
      l1 += fabsf(s.x) + fabsf(s.y) + fabsf(s.z) + fabsf(s.w);
    }
  } else {
    for (long long i = tid; i < n; i += nth) {
      float s = __ldcg(c.stage[0] + slot + i);
      for (int p = 1; p < c.world; ++p) s += __ldcg(c.stage[p] + slot + i);
      buf[i] = s;
      l1 += fabsf(s);
    }
  }
  if (coef) {
    l1 = block_sum(l1, red);
    if (threadIdx.x == 0) block_partials[blockIdx.x] = l1;
  }
  // ---- D: the last CTA closes the epoch (and finishes the norm)
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicInc(&me->done, gridDim.x - 1) == gridDim.x - 1;
  __syncthreads();
  if (is_last) {
    if (coef) {
      float t = 0.f;
      for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) t += block_partials[b];
      t = block_sum(t, red);
      if (threadIdx.x == 0) {
        *coef = fminf(max_norm / (t + 1e-6f), 1.0f);     // clip_grad_norm_: max_norm / (total_norm + 1e-6), <= 1
        if (l1_out) *l1_out = t;
      }
    }
    if (threadIdx.x == 0) *((volatile unsigned*)&me->epoch) = e;
  }
}

int launch_comm_allreduce(const recnn_comm* comm, float* buf, int64_t n, float max_norm, float* coef, float* l1_out,
                          float* block_partials, cudaStream_t st) {
  RECNN_REQUIRE(comm != nullptr && comm->connected, "communicator is not connected");
  RECNN_REQUIRE(n > 0 && n <= comm->peers.capacity, "all-reduce larger than the communicator's staging capacity");
  const int64_t per = (int64_t)kCommThreads * 4 * 2;              // two float4 per thread
  int64_t blocks = ceil_div(n, per);
  const int grid = (int)(blocks < 1 ? 1 : (blocks > kNumSMs ? kNumSMs : blocks));
  allreduce_kernel<<<grid, kCommThreads, 0, st>>>(comm->peers, buf, n, max_norm, coef, l1_out, block_partials);
  RECNN_CHECK_LAUNCH("allreduce_kernel");
  return RECNN_OK;
}

}  // namespace recnn

using namespace recnn;

extern "C" int recnn_comm_create(int32_t rank, int32_t world, int64_t capacity_floats, recnn_comm** out) {
  RECNN_REQUIRE(out != nullptr, "out");
  RECNN_REQUIRE(world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world, "rank/world (at most 8 ranks)");
  RECNN_REQUIRE(capacity_floats > 0, "capacity");
  capacity_floats = round_up(capacity_floats, 4);
  recnn_comm* c = new recnn_comm();
  c->connected = false;
  for (int i = 0; i < kMaxRanks; ++i) { c->opened[i] = nullptr; c->peers.ctrl[i] = nullptr; c->peers.stage[i] = nullptr; }
  const size_t bytes = sizeof(CommDev) + 2 * sizeof(float) * (size_t)capacity_floats;
  cudaError_t e = cudaMalloc(&c->local_base, bytes);
  if (e != cudaSuccess) {
    delete c;
    set_error("cudaMalloc of the all-reduce staging buffer (%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    return RECNN_E_CUDA;
  }
  cudaMemset(c->local_base, 0, bytes);
  c->peers.capacity = capacity_floats;
  c->peers.rank = rank;
  c->peers.world = world;
  e = cudaIpcGetMemHandle(&c->handle, c->local_base);
  if (e != cudaSuccess) {
    cudaFree(c->local_base);
    delete c;
    set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    return RECNN_E_CUDA;
  }
  cudaDeviceSynchronize();
  *out = c;
  return RECNN_OK;
}

extern "C" int32_t recnn_comm_handle_bytes(void) { return (int32_t)sizeof(cudaIpcMemHandle_t); }

extern "C" int recnn_comm_local_handle(const recnn_comm* c, void* out) {
  RECNN_REQUIRE(c && out, "args");
  memcpy(out, &c->handle, sizeof(cudaIpcMemHandle_t));
  return RECNN_OK;
}

// all_handles: world consecutive handles in rank order (this rank's own entry is ignored)
extern "C" int recnn_comm_connect(recnn_comm* c, const void* all_handles) {
  RECNN_REQUIRE(c && all_handles, "args");
  RECNN_REQUIRE(!c->connected, "already connected");
  const char* hs = static_cast<const char*>(all_handles);
  for (int p = 0; p < c->peers.world; ++p) {
    void* base = c->local_base;
    if (p != c->peers.rank) {
      cudaIpcMemHandle_t h;
      memcpy(&h, hs + (size_t)p * sizeof(h), sizeof(h));
      cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        set_error("cudaIpcOpenMemHandle(rank %d) failed: %s", p, cudaGetErrorString(e));
        for (int q = 0; q < p; ++q)
          if (c->opened[q]) { cudaIpcCloseMemHandle(c->opened[q]); c->opened[q] = nullptr; }
        (void)cudaGetLastError();
        return RECNN_E_CUDA;
      }
      c->opened[p] = base;
    }
    c->peers.ctrl[p] = static_cast<CommDev*>(base);
    c->peers.stage[p] = reinterpret_cast<float*>(static_cast<char*>(base) + sizeof(CommDev));
  }
  c->connected = true;
  return RECNN_OK;
}

extern "C" int recnn_comm_allreduce(const recnn_comm* c, float* buf, int64_t n, void* stream) {
  return launch_comm_allreduce(c, buf, n, 0.f, nullptr, nullptr, nullptr, static_cast<cudaStream_t>(stream));
}

extern "C" int recnn_comm_destroy(recnn_comm* c) {
  if (!c) return RECNN_OK;
  cudaDeviceSynchronize();
  for (int p = 0; p < kMaxRanks; ++p)
    if (c->opened[p]) cudaIpcCloseMemHandle(c->opened[p]);
  if (c->local_base) cudaFree(c->local_base);
  delete c;
  return RECNN_OK;
}
