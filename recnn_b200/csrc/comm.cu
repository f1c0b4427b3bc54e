// Peer-memory all-reduce of the gradient arenas (data-parallel update step, one process per GPU).
//
// The reference is single-process; BASELINE north_star asks for the minibatch to be sharded over the
// GPUs of one box with an all-reduce of the Actor/Critic gradients over NVLink.  Calling NCCL between
// the phases of the step costs three host-launched collectives and cuts the step's CUDA graph in three;
// instead every rank maps its peers' staging buffers (cudaIpc, NVLink/NVSwitch peer access) and ONE
// kernel per gradient arena does a TWO-SHOT all-reduce with remote STORES only (posted writes: nothing
// ever waits on an NVLink read round trip), fused with everything that used to sit around it:
//
//   A  reduce-scatter, push side: rank r owns slice r of the arena.  Every rank stores slice s of its local
//      gradient into rank s's contribution buffer [my rank] (+ its few "aux" floats: loss partial sums and
//      the global row count it assumed), fences, and raises flag A on every peer.
//   B  owner side: wait for all W flags A, sum the W contributions of my slice IN RANK ORDER (the same
//      bits on every rank), store the reduced slice into every peer's result buffer (+ this CTA's partial
//      L1 norm of the slice), fence, raise flag B on every peer.
//   C  wait for all W flags B.  Every rank now holds the whole reduced gradient: sum the aux floats in rank
//      order (-> global loss means), form the reference's clip_grad_norm_(params, -1, 1) coefficient from the
//      L1 partials (fixed order), write the gradient back to the caller's arena (scaled by the coefficient when
//      there is one) and -- when the caller passes a built-in optimizer -- apply SGD/Adam in the same pass.
//
// Bytes on NVLink per rank and arena: 2 (W-1)/W x 1.72 MB (3.0 MB at W = 8) instead of the (W-1) x 1.72 MB of
// remote READS (12 MB at W = 8) of the one-shot version of round 1; two flag hops instead of one.
// It is an ordinary kernel on the step's stream, so the whole data-parallel step is captured in one
// CUDA graph exactly like the single-GPU step.  All sums are taken in rank order 0..W-1 on every rank,
// so all replicas hold bit-identical gradients (and therefore weights) after every step.
//
// Protocol (epoch e = number of collectives issued so far on this communicator + 1; all on one stream):
//   * contribution / result / aux buffers are double buffered by e & 1.  A rank can only be one epoch ahead
//     of any peer (it needs every peer's flags of epoch e to finish e), so when it writes buffers (e+2) & 1
//     = e & 1 during epoch e+2 every peer has completed its epoch-e kernel and no longer reads them.
//   * flags hold epochs and are compared as signed distances (a 32-bit wrap is harmless).
//   * waits are bounded (~20 s of %globaltimer): a lost peer makes the kernel trap instead of hanging the GPU.
//   * ranks that disagree on n_rows_global (uneven shards without batch["n_rows_global"]) raise *err_flag.
#include <string.h>

#include <type_traits>

#include "common.cuh"
#include "pointwise.cuh"

namespace recnn {

constexpr int kMaxRanks = 8;
constexpr int kCommThreads = 512;
constexpr int kMaxAux = 8;

struct CommDev {                       // lives at the head of every rank's shared allocation
  unsigned flag_a[kMaxRanks];          // flag_a[src] = last epoch whose contributions src has published here
  unsigned flag_b[kMaxRanks];          // flag_b[src] = last epoch whose reduced slice src has published here
  unsigned epoch;                      // collectives completed by this rank           (local)
  unsigned arrive_a, arrive_b, done;   // CTA counters (wrap to 0)                     (local)
  unsigned pad[12];
  float aux[2][kMaxRanks][kMaxAux + 1];           // [e & 1][src]: aux floats, then the row count src assumed
  unsigned long long aux_ll[2][kMaxRanks][kMaxAux + 1];   // the same for scalar-only exchanges: {value, epoch} words
  float l1[2][kMaxRanks][kNumSMs];                // [e & 1][src][cta]: partial L1 norms of src's reduced slice
};

struct CommPeers {                     // kernel parameter
  CommDev* ctrl[kMaxRanks];
  float* contrib[kMaxRanks];           // [2][W][slice_cap] floats: contributions to THAT rank's slice
  float* result[kMaxRanks];            // [2][capacity] floats: the reduced gradient, assembled by the owners
  long long capacity, slice_cap;
  int rank, world;
};

struct CommOpt {                       // optional fused optimizer (kind == RECNN_OPT_EXTERNAL: none)
  int kind;
  OptConsts k;
  double beta1, beta2, lr, wd, n_sma_threshold;
  int k_look;
  float *p, *m, *v, *slow;
  int* t;
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
__device__ __forceinline__ void st_relaxed_sys(unsigned* p, unsigned v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Raise flags[rank] = e on every peer.  Called by the whole CTA that arrived last (all CTAs' payload stores are fenced
// and counted by then): ONE system-scope fence, then W relaxed flag stores issued by W different threads in parallel.
// (r2m8 measured what W sequential st.release.sys from one thread cost: each release drains the thread's outstanding
// remote stores again -- 31 us per collective at W = 4 and 47 us at W = 8 for a 16-byte payload.)
__device__ __forceinline__ void raise_flags(CommPeers& c, bool phase_b, unsigned e) {
  if ((int)threadIdx.x < c.world) {
    __threadfence_system();
    CommDev* peer = c.ctrl[threadIdx.x];
    st_relaxed_sys(phase_b ? &peer->flag_b[c.rank] : &peer->flag_a[c.rank], e);
  }
}
// "LL" words of the scalar-only exchange: {value, epoch} travel in ONE 8-byte store, so the receiver polls the
// payload itself and no fence or flag is needed (NCCL's low-latency protocol, for 36 bytes per peer)
__device__ __forceinline__ void st_ll(unsigned long long* p, float v, unsigned e) {
  const unsigned long long w = ((unsigned long long)e << 32) | (unsigned long long)__float_as_uint(v);
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned long long ld_ll(const unsigned long long* p) {
  unsigned long long w;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
  return w;
}
// threads [0, world) poll one peer's flag each; bounded
__device__ __forceinline__ void wait_flags(const unsigned* flags, unsigned e, int world, int rank, const char* what) {
  if ((int)threadIdx.x < world) {
    const unsigned long long t0 = comm_gtimer();
    while ((int)(ld_acquire_sys(&flags[threadIdx.x]) - e) < 0) {
      if (comm_gtimer() - t0 > 20000000000ull) {
        printf("recnn_b200 allreduce: rank %d timed out waiting for %s of rank %d (epoch %u)\n", rank, what,
               (int)threadIdx.x, e);
        __trap();
      }
    }
  }
  __syncthreads();
}

// grid <= number of SMs (all CTAs must be co-resident: they wait for each other's peers).
// VEC = 4: n % 4 == 0 and buf 16-byte aligned (the arenas); VEC = 1: anything.
template <int VEC>
__global__ void __launch_bounds__(kCommThreads, 1)
allreduce_kernel(CommPeers c, float* __restrict__ buf, long long n, float max_norm, float* coef_out, float* l1_out,
                 const float* aux_in, float* aux_out, int n_aux, float check_val, int* err_flag, CommOpt opt,
                 GradSource src) {
  __shared__ float red[32];
  __shared__ unsigned s_epoch;
  __shared__ bool s_last, s_sig;
  __shared__ float s_coef;
  __shared__ OptStep s_st;
  CommDev* me = c.ctrl[c.rank];
  if (threadIdx.x == 0) s_epoch = *((volatile unsigned*)&me->epoch) + 1;
  __syncthreads();
  const unsigned e = s_epoch;
  const int par = (int)(e & 1u);
  const int W = c.world;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  using V = typename std::conditional<VEC == 4, float4, float>::type;
  const long long units = n / VEC;                               // VEC == 4 => n % 4 == 0
  const long long slice = (units + W - 1) / W;                   // units per owner (the last slice may be short)

  if (n == 0) {
    // ---- scalar-only exchange (one CTA): LL words to every peer, poll my own copies, sum in rank order.  One NVLink
    // write latency end to end; no fences, no flags.
    const int nw = n_aux + 1;
    if ((int)threadIdx.x < W) {
      unsigned long long* dst = c.ctrl[threadIdx.x]->aux_ll[par][c.rank];
      for (int j = 0; j < n_aux; ++j) st_ll(dst + j, aux_in[j], e);
      st_ll(dst + n_aux, check_val, e);
      const unsigned long long* mine = me->aux_ll[par][threadIdx.x];
      const unsigned long long t0 = comm_gtimer();
      for (int j = 0; j < nw; ++j) {
        while ((unsigned)(ld_ll(mine + j) >> 32) != e) {
          if (comm_gtimer() - t0 > 20000000000ull) {
            printf("recnn_b200 scalar exchange: rank %d timed out waiting for rank %d (epoch %u)\n", c.rank, (int)threadIdx.x, e);
            __trap();
          }
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      bool bad = false;
      for (int p = 0; p < W; ++p) bad = bad || __uint_as_float((unsigned)ld_ll(&me->aux_ll[par][p][n_aux])) != check_val;
      if (bad && err_flag) *err_flag = 1;
      for (int j = 0; j < n_aux; ++j) {
        float t = 0.f;
        for (int p = 0; p < W; ++p) t += __uint_as_float((unsigned)ld_ll(&me->aux_ll[par][p][j]));
        aux_out[j] = t;
      }
      *((volatile unsigned*)&me->epoch) = e;
    }
    return;
  }
  // ---- A: push slice s of my gradient to owner s (remote stores), aux floats to everyone
  for (long long i = tid; i < units; i += nth) {
    const int s = (int)(i / slice);
    V* dst = reinterpret_cast<V*>(c.contrib[s] + ((long long)par * W + c.rank) * c.slice_cap);
    V g;
    if (src.n_layers) {              // the local gradient is still in split-K partials: reduce them on the way out
      if constexpr (VEC == 4) g = grad4_at(src, buf, (unsigned)(4 * i));
      else g = grad_at(src, buf, i);
    } else {
      g = reinterpret_cast<const V*>(buf)[i];
    }
    dst[i - (long long)s * slice] = g;
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < W) {
    float* dst = c.ctrl[threadIdx.x]->aux[par][c.rank];
    for (int j = 0; j < n_aux; ++j) dst[j] = aux_in[j];
    dst[kMaxAux] = check_val;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_sig = atomicInc(&me->arrive_a, gridDim.x - 1) == gridDim.x - 1;
  __syncthreads();
  if (s_sig) raise_flags(c, false, e);       // every CTA's stores are visible system-wide
  // ---- B: owner: wait for all contributions, reduce my slice in rank order, push it to everyone
  wait_flags(me->flag_a, e, W, c.rank, "the contribution");
  {
    const long long lo = (long long)c.rank * slice;
    const long long cnt = units - lo < slice ? (units - lo > 0 ? units - lo : 0) : slice;
    const float* mine = c.contrib[c.rank] + (long long)par * W * c.slice_cap;
    float l1 = 0.f;
    for (long long i = tid; i < cnt; i += nth) {
      V s = __ldcg(reinterpret_cast<const V*>(mine) + i);
      for (int p = 1; p < W; ++p) {
        const V v = __ldcg(reinterpret_cast<const V*>(mine + (long long)p * c.slice_cap) + i);
        if constexpr (VEC == 4) { s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
        else s += v;
      }
      if constexpr (VEC == 4) l1 += fabsf(s.x) + fabsf(s.y) + fabsf(s.z) + fabsf(s.w);
      else l1 += fabsf(s);
      for (int p = 0; p < W; ++p)
        reinterpret_cast<V*>(c.result[p] + (long long)par * c.capacity)[lo + i] = s;
    }
    if (coef_out) {
      l1 = block_sum(l1, red);
      if ((int)threadIdx.x < W) c.ctrl[threadIdx.x]->l1[par][c.rank][blockIdx.x] = l1;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_sig = atomicInc(&me->arrive_b, gridDim.x - 1) == gridDim.x - 1;
  __syncthreads();
  if (s_sig) raise_flags(c, true, e);
  // ---- C: everyone: the reduced gradient is complete here
  wait_flags(me->flag_b, e, W, c.rank, "the reduced slice");
  if (threadIdx.x < 32) {
    float coef = 1.0f;
    if (coef_out) {
      // ||g||_1 from the owners' per-CTA partials: lane l takes entries l, l+32, ... of the (rank-major, CTA-minor)
      // list, then a shuffle tree -- the same order on every rank and CTA, so every replica gets the same bits
      const int total = W * (int)gridDim.x;
      float t = 0.f;
      for (int k = (int)threadIdx.x; k < total; k += 32) t += __ldcg(&me->l1[par][k / (int)gridDim.x][k % (int)gridDim.x]);
      t = warp_sum(t);
      coef = fminf(max_norm / (t + 1e-6f), 1.0f);     // clip_grad_norm_: max_norm / (total_norm + 1e-6), <= 1
      if (blockIdx.x == 0 && threadIdx.x == 0) {
        *coef_out = coef;
        if (l1_out) *l1_out = t;
      }
    }
    if (threadIdx.x == 0) {
      s_coef = coef;
      if (opt.kind != RECNN_OPT_EXTERNAL)
        s_st = opt_step_scalars(opt.kind, opt.beta1, opt.beta2, opt.lr, opt.wd, opt.n_sma_threshold, opt.k_look, *opt.t + 1);
      if (blockIdx.x == 0) {
        bool bad = false;
        for (int p = 0; p < W; ++p) bad = bad || __ldcg(&me->aux[par][p][kMaxAux]) != check_val;
        if (bad && err_flag) *err_flag = 1;
        for (int j = 0; j < n_aux; ++j) {
          float t = 0.f;
          for (int p = 0; p < W; ++p) t += __ldcg(&me->aux[par][p][j]);
          aux_out[j] = t;
        }
      }
    }
  }
  __syncthreads();
  {
    const float coef = s_coef;
    const bool scale = coef_out != nullptr;
    const int t_next = opt.kind != RECNN_OPT_EXTERNAL ? *opt.t + 1 : 0;
    const OptStep ost = s_st;
    const float* res = c.result[c.rank] + (long long)par * c.capacity;
    for (long long i = tid; i < units; i += nth) {
      V g = __ldcg(reinterpret_cast<const V*>(res) + i);
      if (scale) {
        if constexpr (VEC == 4) { g.x = __fmul_rn(g.x, coef); g.y = __fmul_rn(g.y, coef); g.z = __fmul_rn(g.z, coef); g.w = __fmul_rn(g.w, coef); }
        else g = __fmul_rn(g, coef);
      }
      reinterpret_cast<V*>(buf)[i] = g;              // the caller-visible .grad (scaled, as the reference leaves it)
      if (opt.kind != RECNN_OPT_EXTERNAL) {
        if constexpr (VEC == 4) {
          opt_apply(opt.kind, opt.k, ost, t_next, opt.p, opt.m, opt.v, opt.slow, 4 * i + 0, g.x);
          opt_apply(opt.kind, opt.k, ost, t_next, opt.p, opt.m, opt.v, opt.slow, 4 * i + 1, g.y);
          opt_apply(opt.kind, opt.k, ost, t_next, opt.p, opt.m, opt.v, opt.slow, 4 * i + 2, g.z);
          opt_apply(opt.kind, opt.k, ost, t_next, opt.p, opt.m, opt.v, opt.slow, 4 * i + 3, g.w);
        } else {
          opt_apply(opt.kind, opt.k, ost, t_next, opt.p, opt.m, opt.v, opt.slow, i, g);
        }
      }
    }
  }
  // ---- D: the last CTA closes the epoch (and advances the optimizer's step count)
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicInc(&me->done, gridDim.x - 1) == gridDim.x - 1;
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    if (opt.kind != RECNN_OPT_EXTERNAL) *opt.t = *opt.t + 1;
    *((volatile unsigned*)&me->epoch) = e;
  }
}

int launch_comm_allreduce(const recnn_comm* comm, float* buf, int64_t n, const CommReduce& r, cudaStream_t st) {
  RECNN_REQUIRE(comm != nullptr && comm->connected, "communicator is not connected");
  RECNN_REQUIRE(n >= 0 && n <= comm->peers.capacity, "all-reduce larger than the communicator's staging capacity");
  RECNN_REQUIRE(n == 0 || buf != nullptr, "buf");
  RECNN_REQUIRE(r.n_aux >= 0 && r.n_aux <= kMaxAux && (r.n_aux == 0 || (r.aux_in && r.aux_out)), "aux floats");
  CommOpt opt;
  memset(&opt, 0, sizeof(opt));
  opt.kind = RECNN_OPT_EXTERNAL;
  if (r.optim && r.optim->kind != RECNN_OPT_EXTERNAL) {
    RECNN_REQUIRE(r.net && r.net->params && r.net->opt_t, "fused optimizer needs params and the step counter");
    RECNN_REQUIRE(r.optim->kind == RECNN_OPT_SGD || r.optim->kind == RECNN_OPT_ADAM || r.optim->kind == RECNN_OPT_RANGER,
                  "built-in optimizer kind must be SGD, ADAM or RANGER");
    if (r.optim->kind == RECNN_OPT_ADAM) RECNN_REQUIRE(r.net->opt_m && r.net->opt_v, "Adam needs exp_avg / exp_avg_sq arenas");
    if (r.optim->kind == RECNN_OPT_RANGER) RECNN_REQUIRE(r.net->opt_m && r.net->opt_v && r.net->opt_slow, "Ranger needs exp_avg / exp_avg_sq / slow arenas");
    if (r.optim->kind == RECNN_OPT_SGD && r.optim->momentum != 0.f) RECNN_REQUIRE(r.net->opt_m, "SGD momentum needs a buffer arena");
    opt.kind = r.optim->kind;
    opt.k = opt_consts(*r.optim);
    opt.beta1 = r.optim->beta1; opt.beta2 = r.optim->beta2; opt.lr = r.optim->lr;
    opt.wd = r.optim->weight_decay; opt.n_sma_threshold = r.optim->n_sma_threshold; opt.k_look = r.optim->k;
    opt.p = r.net->params; opt.m = r.net->opt_m; opt.v = r.net->opt_v; opt.slow = r.net->opt_slow; opt.t = r.net->opt_t;
  }
  GradSource gsrc;
  memset(&gsrc, 0, sizeof(gsrc));
  if (r.src) gsrc = *r.src;
  const int64_t per = (int64_t)kCommThreads * 4 * 2;              // two float4 per thread
  int64_t blocks = ceil_div(n, per);
  const int grid = (int)(blocks < 1 ? 1 : (blocks > kNumSMs ? kNumSMs : blocks));
  const bool vec = (n & 3) == 0 && ((reinterpret_cast<uintptr_t>(buf) & 15) == 0) &&
                   (opt.kind == RECNN_OPT_EXTERNAL || (reinterpret_cast<uintptr_t>(opt.p) & 15) == 0);
  if (vec)
    allreduce_kernel<4><<<grid, kCommThreads, 0, st>>>(comm->peers, buf, n, r.max_norm, r.coef, r.l1_out, r.aux_in,
                                                        r.aux_out, r.n_aux, r.check_val, r.err_flag, opt, gsrc);
  else
    allreduce_kernel<1><<<grid, kCommThreads, 0, st>>>(comm->peers, buf, n, r.max_norm, r.coef, r.l1_out, r.aux_in,
                                                        r.aux_out, r.n_aux, r.check_val, r.err_flag, opt, gsrc);
  RECNN_CHECK_LAUNCH("allreduce_kernel");
  return RECNN_OK;
}

}  // namespace recnn

using namespace recnn;

// staging layout of one rank: CommDev | contrib [2][W][slice_cap] | result [2][capacity]   (floats)
static size_t comm_ctrl_bytes() { return (size_t)round_up((int64_t)sizeof(CommDev), 256); }
static void comm_carve(recnn_comm* c, int p, void* base) {
  char* b = static_cast<char*>(base);
  c->peers.ctrl[p] = reinterpret_cast<CommDev*>(b);
  c->peers.contrib[p] = reinterpret_cast<float*>(b + comm_ctrl_bytes());
  c->peers.result[p] = c->peers.contrib[p] + 2ll * c->peers.world * c->peers.slice_cap;
}

extern "C" int recnn_comm_create(int32_t rank, int32_t world, int64_t capacity_floats, recnn_comm** out) {
  RECNN_REQUIRE(out != nullptr, "out");
  RECNN_REQUIRE(world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world, "rank/world (at most 8 ranks)");
  RECNN_REQUIRE(capacity_floats > 0, "capacity");
  capacity_floats = round_up(capacity_floats, 4);
  recnn_comm* c = new recnn_comm();
  c->connected = false;
  for (int i = 0; i < kMaxRanks; ++i) {
    c->opened[i] = nullptr; c->peers.ctrl[i] = nullptr; c->peers.contrib[i] = nullptr; c->peers.result[i] = nullptr;
  }
  c->peers.capacity = capacity_floats;
  c->peers.slice_cap = round_up(ceil_div(capacity_floats, world), 4);
  c->peers.rank = rank;
  c->peers.world = world;
  const size_t bytes = comm_ctrl_bytes() +
                       sizeof(float) * (size_t)(2 * world * c->peers.slice_cap + 2 * capacity_floats);
  cudaError_t e = cudaMalloc(&c->local_base, bytes);
  if (e != cudaSuccess) {
    delete c;
    set_error("cudaMalloc of the all-reduce staging buffer (%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    return RECNN_E_CUDA;
  }
  cudaMemset(c->local_base, 0, bytes);
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
    comm_carve(c, p, base);
  }
  c->connected = true;
  return RECNN_OK;
}

extern "C" int recnn_comm_allreduce(const recnn_comm* c, float* buf, int64_t n, void* stream) {
  return launch_comm_allreduce(c, buf, n, CommReduce(), static_cast<cudaStream_t>(stream));
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
