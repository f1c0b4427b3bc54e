// Orchestration of the DDPG / TD3 update step on one GPU (one rank's shard).
// Restates recnn/nn/update/ddpg.py:8-104, misc.py:10-55 and td3.py:8-150 as a
// fixed sequence of kernel launches on the caller's stream: no allocation, no
// synchronisation, no host read-back -- so the whole call is CUDA-graph
// capturable.  Step-dependent scalars (Adam's t, the Philox step) live on the
// device for the same reason.
#include "common.cuh"
#include "gemm_simt.cuh"
#include "pointwise.cuh"

namespace recnn {

// ---------------------------------------------------------------- workspace
struct Workspace {
  float* S;        // [N,S]  frame form only
  float* S2;       // [N,S]
  float* ACT;      // [N,A]
  float* REW;      // [N]
  float* hb[6];    // [N,H] activation / gradient buffers (lifetimes in DESIGN.md)
  float* ab[3];    // [N,A] next_action / gen_action / d gen_action
  float* y;        // [N] TD target
  float* qtmp;     // [N]
  float* dq;       // [N]
  float* partial;  // split-K partials of the largest weight gradient
  float* block_partials;  // [1024]
  float* scalars;         // [8]: 0 = clip coef
  unsigned* tickets;      // [8]
  int64_t bytes;
};

static int dw_splits(int C, int K1, int64_t n_rows) {
  // enough (m,n,z) tiles for ~1.5 waves of 148 SMs, chunks of >= 256 rows
  const int64_t tiles = ceil_div(C, 128) * ceil_div(K1, 128);
  int64_t s = ceil_div(220, tiles);
  const int64_t max_s = n_rows / 256 > 0 ? n_rows / 256 : 1;
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  return (int)s;
}

static int64_t partial_floats(const recnn_dims& d, int64_t n_rows) {
  const int in_c = d.state_dim + d.action_dim;
  int64_t best = 0;
  const int shapes[5][2] = {{d.hidden, in_c + 1}, {d.hidden, d.state_dim + 1}, {d.hidden, d.hidden + 1},
                            {d.action_dim, d.hidden + 1}, {1, d.hidden + 1}};
  for (auto& s : shapes) {
    const int64_t f = (int64_t)dw_splits(s[0], s[1], n_rows) * s[0] * s[1];
    if (f > best) best = f;
  }
  return best;
}

static Workspace carve(const recnn_dims& d, int64_t n, void* base) {
  Workspace w;
  char* p = static_cast<char*>(base);
  int64_t off = 0;
  auto take = [&](int64_t floats) {
    float* r = base ? reinterpret_cast<float*>(p + off) : nullptr;
    off += round_up(floats * 4, 256);
    return r;
  };
  w.S = take(n * d.state_dim);
  w.S2 = take(n * d.state_dim);
  w.ACT = take(n * d.action_dim);
  w.REW = take(n);
  for (auto& b : w.hb) b = take(n * d.hidden);
  for (auto& b : w.ab) b = take(n * d.action_dim);
  w.y = take(n);
  w.qtmp = take(n);
  w.dq = take(n);
  w.partial = take(partial_floats(d, n));
  w.block_partials = take(1024);
  w.scalars = take(8);
  w.tickets = reinterpret_cast<unsigned*>(take(8));
  w.bytes = off;
  return w;
}

// ---------------------------------------------------------------- building blocks
struct Rng {
  const uint8_t* const* masks;
  unsigned long long seed;
  const long long* step;
};

static Epilogue base_epi() {
  Epilogue e;
  e.out = nullptr; e.ldo = 0; e.bias = nullptr; e.mask = nullptr; e.train = 0; e.seed = 0;
  e.rng_step = nullptr; e.stream_id = 0; e.h = nullptr; e.ldh = 0; e.gate_scale = 1.f;
  e.apply_tanh = 0; e.noise = nullptr; e.noise_clip = 0.f; e.noise_std = 0.f; e.add_noise = 0;
  return e;
}

// h = dropout(relu(X W^T + b))   X may be the virtual concat [x0 | x1]
static int hidden_layer(const MatView& X, int K, const float* W, const float* b, int H, int64_t n,
                        bool train, const uint8_t* mask, const Rng& rng, unsigned stream_id, float* out,
                        cudaStream_t st) {
  Epilogue e = base_epi();
  e.out = out; e.ldo = H; e.bias = b;
  e.train = train ? 1 : 0; e.mask = mask; e.seed = rng.seed; e.rng_step = rng.step; e.stream_id = stream_id;
  return launch_gemm_simt<true, true, EPI_HIDDEN>(X, mat(W, K), (int)n, H, K, 1, e, st);
}

struct NoiseSpec {
  int add; const float* noise; float clip, std; unsigned long long seed; const long long* step; unsigned stream_id;
};

static int linear_out(const float* X, int K, const float* W, const float* b, int out_dim, int64_t n,
                      int apply_tanh, const NoiseSpec* nz, float* out, cudaStream_t st) {
  Epilogue e = base_epi();
  e.out = out; e.ldo = out_dim; e.bias = b; e.apply_tanh = apply_tanh;
  if (nz && nz->add) {
    e.add_noise = 1; e.noise = nz->noise; e.noise_clip = nz->clip; e.noise_std = nz->std;
    e.seed = nz->seed; e.rng_step = nz->step; e.stream_id = nz->stream_id;
  }
  return launch_gemm_simt<true, true, EPI_LINEAR>(mat(X, K), mat(W, K), (int)n, out_dim, K, 1, e, st);
}

// dX = (dZ W) * gate(h)     dZ [n,C], W [C,K] (row-major, optionally a column window), out [n,K]
static int backprop_hidden(const float* dZ, int C, const float* W, long long ldw, int K, int64_t n,
                           const float* h, float gate_scale, float* out, cudaStream_t st) {
  Epilogue e = base_epi();
  e.out = out; e.ldo = K; e.h = h; e.ldh = K; e.gate_scale = gate_scale;
  if (h) return launch_gemm_simt<true, false, EPI_GATE>(mat(dZ, C), mat(W, ldw), (int)n, K, C, 1, e, st);
  return launch_gemm_simt<true, false, EPI_STORE>(mat(dZ, C), mat(W, ldw), (int)n, K, C, 1, e, st);
}

// dW[c,k] = sum_n dZ[n,c] X[n,k];  db[c] = sum_n dZ[n,c]   (X may be a concat view)
static int weight_grad(const float* dZ, int C, MatView X, int K, int64_t n, float* dW, float* db,
                       const Workspace& ws, cudaStream_t st) {
  X.ones_at = K;                        // virtual bias column
  const int K1 = K + 1;
  const int splits_req = dw_splits(C, K1, n);
  Epilogue e = base_epi();
  e.out = ws.partial; e.ldo = K1;
  // the launcher rounds the chunk; recompute the effective split count the same way
  const int k_chunk = (int)round_up(ceil_div(n, splits_req), 16);
  const int splits = (int)ceil_div(n, k_chunk);
  RECNN_PROPAGATE((launch_gemm_simt<false, false, EPI_PARTIAL>(mat(dZ, C), X, C, K1, (int)n, splits_req, e, st)));
  return launch_reduce_partials(ws.partial, splits, C, K1, dW, db, st);
}

struct Ctx {
  const recnn_step_args* a;
  recnn_dims d;
  NetLayout la, lc;
  Workspace ws;
  const float *S, *S2, *ACT, *REW, *DONE;
  int64_t n;
  cudaStream_t st;
  Rng rng;
  bool train;
  float gate;
};

// Critic hidden layers on (s, act):  h1 -> out1, h2 -> out2
static int critic_hidden(const Ctx& c, const float* params, const float* s, const float* act, bool train,
                         int mask_base, float* out1, float* out2) {
  const int S = c.d.state_dim, A = c.d.action_dim, H = c.d.hidden;
  const MatView X = mat_cat(s, S, S, act, A);
  const uint8_t* m1 = (train && c.rng.masks) ? c.rng.masks[mask_base] : nullptr;
  const uint8_t* m2 = (train && c.rng.masks) ? c.rng.masks[mask_base + 1] : nullptr;
  RECNN_PROPAGATE(hidden_layer(X, S + A, params + c.lc.w1, params + c.lc.b1, H, c.n, train, m1, c.rng,
                               mask_base, out1, c.st));
  return hidden_layer(mat(out1, H), H, params + c.lc.w2, params + c.lc.b2, H, c.n, train, m2, c.rng,
                      mask_base + 1, out2, c.st);
}

static int actor_hidden(const Ctx& c, const float* params, const float* s, bool train, int mask_base,
                        float* out1, float* out2) {
  const int S = c.d.state_dim, H = c.d.hidden;
  const uint8_t* m1 = (train && c.rng.masks) ? c.rng.masks[mask_base] : nullptr;
  const uint8_t* m2 = (train && c.rng.masks) ? c.rng.masks[mask_base + 1] : nullptr;
  RECNN_PROPAGATE(hidden_layer(mat(s, S), S, params + c.la.w1, params + c.la.b1, H, c.n, train, m1, c.rng,
                               mask_base, out1, c.st));
  return hidden_layer(mat(out1, H), H, params + c.la.w2, params + c.la.b2, H, c.n, train, m2, c.rng,
                      mask_base + 1, out2, c.st);
}

static HeadArgs head_args(const Ctx& c, const float* params, const float* h2, int mode) {
  HeadArgs h;
  h.h2 = h2; h.w3 = params + c.lc.w3; h.b3 = params + c.lc.b3;
  h.n_rows = c.n; h.n_rows_global = c.a->n_rows_global; h.hidden = c.d.hidden; h.mode = mode;
  h.reward = c.REW; h.done = c.DONE;
  h.gamma = c.a->gamma; h.min_value = c.a->min_value; h.max_value = c.a->max_value;
  h.y = c.ws.y; h.tmp = c.ws.qtmp; h.out = nullptr; h.dq = c.ws.dq; h.loss = nullptr;
  h.block_partials = c.ws.block_partials; h.ticket = c.ws.tickets;
  return h;
}

// ---------------------------------------------------------------- phases
static int phase_value_grad(Ctx& c) {
  const recnn_step_args& a = *c.a;
  const int S = c.d.state_dim, A = c.d.action_dim, H = c.d.hidden;
  const bool td3 = a.algo == RECNN_ALGO_TD3;
  const int n_critics = td3 ? 2 : 1;
  float *X0 = c.ws.hb[0], *X1 = c.ws.hb[1], *c1 = c.ws.hb[2], *c2 = c.ws.hb[3], *dz2 = c.ws.hb[4],
        *dz1 = c.ws.hb[5];
  float* a2 = c.ws.ab[0];

  // target policy on next_state, eval mode (misc.py:28 / td3.py:73) (+ clipped noise, td3.py:74-78)
  RECNN_PROPAGATE(actor_hidden(c, a.target_policy.params, c.S2, false, 0, X0, X1));
  NoiseSpec nz = {td3 ? 1 : 0, a.noise, a.noise_clip, a.noise_std, a.seed, (const long long*)a.rng_step, 15u};
  RECNN_PROPAGATE(linear_out(X1, H, a.target_policy.params + c.la.w3, a.target_policy.params + c.la.b3, A, c.n,
                             0, &nz, a2, c.st));
  if (a.next_action_out)
    RECNN_CHECK_CUDA(cudaMemcpyAsync(a.next_action_out, a2, sizeof(float) * c.n * A, cudaMemcpyDeviceToDevice, c.st));

  // target critic(s) -> TD target y (misc.py:29-35 / td3.py:80-86)
  for (int i = 0; i < n_critics; ++i) {
    RECNN_PROPAGATE(critic_hidden(c, a.target_value[i].params, c.S2, a2, false, 0, X0, X1));
    HeadArgs h = head_args(c, a.target_value[i].params, X1,
                           td3 ? (i == 0 ? HEAD_TARGET_TD3_A : HEAD_TARGET_TD3_B) : HEAD_TARGET_DDPG);
    RECNN_PROPAGATE(launch_critic_head(h, c.st));
  }

  // online critic(s): value, loss, backward (misc.py:37-43 / td3.py:88-101)
  for (int i = 0; i < n_critics; ++i) {
    const float* P = a.value[i].params;
    RECNN_PROPAGATE(critic_hidden(c, P, c.S, c.ACT, c.train, 2 * i, c1, c2));
    HeadArgs h = head_args(c, P, c2, HEAD_VALUE);
    h.loss = a.losses + i;
    RECNN_PROPAGATE(launch_critic_head(h, c.st));
    if (!a.learn) continue;
    float* G = a.value[i].grads;
    RECNN_REQUIRE(G != nullptr, "value net needs a grad arena when learn=1");
    // layer 3: dW3 = dq^T h2, db3 = sum dq ; dz2 = (dq w3) * gate(h2)
    RECNN_PROPAGATE(weight_grad(c.ws.dq, 1, mat(c2, H), H, c.n, G + c.lc.w3, G + c.lc.b3, c.ws, c.st));
    RECNN_PROPAGATE(launch_critic_head_bwd(c.ws.dq, 0.f, P + c.lc.w3, c2, c.gate, dz2, c.n, H, c.st));
    RECNN_PROPAGATE(weight_grad(dz2, H, mat(c1, H), H, c.n, G + c.lc.w2, G + c.lc.b2, c.ws, c.st));
    RECNN_PROPAGATE(backprop_hidden(dz2, H, P + c.lc.w2, H, H, c.n, c1, c.gate, dz1, c.st));
    RECNN_PROPAGATE(weight_grad(dz1, H, mat_cat(c.S, S, S, c.ACT, A), S + A, c.n, G + c.lc.w1, G + c.lc.b1,
                                c.ws, c.st));
  }
  return RECNN_OK;
}

static int phase_value_opt(Ctx& c) {
  const recnn_step_args& a = *c.a;
  if (!a.learn || a.value_optim.kind == RECNN_OPT_EXTERNAL) return RECNN_OK;
  const int n_critics = a.algo == RECNN_ALGO_TD3 ? 2 : 1;
  for (int i = 0; i < n_critics; ++i)
    RECNN_PROPAGATE(launch_optimizer(a.value_optim, a.value[i], c.lc.count, nullptr, c.st));
  return RECNN_OK;
}

static int phase_policy_loss(Ctx& c) {
  const recnn_step_args& a = *c.a;
  const int H = c.d.hidden, A = c.d.action_dim;
  const bool td3 = a.algo == RECNN_ALGO_TD3;
  const int pm = td3 ? 4 : 2, vm = td3 ? 6 : 4;     // mask slots (header: call order)
  float *p1 = c.ws.hb[0], *p2 = c.ws.hb[1], *v1 = c.ws.hb[2], *v2 = c.ws.hb[3];
  float* gen = c.ws.ab[1];
  // gen_action = policy_net(state); policy_loss = -value_net(state, gen_action)  (ddpg.py:78-79, td3.py:116-118)
  RECNN_PROPAGATE(actor_hidden(c, a.policy.params, c.S, c.train, pm, p1, p2));
  RECNN_PROPAGATE(linear_out(p2, H, a.policy.params + c.la.w3, a.policy.params + c.la.b3, A, c.n, 0, nullptr,
                             gen, c.st));
  if (a.gen_action_out)
    RECNN_CHECK_CUDA(cudaMemcpyAsync(a.gen_action_out, gen, sizeof(float) * c.n * A, cudaMemcpyDeviceToDevice, c.st));
  RECNN_PROPAGATE(critic_hidden(c, a.value[0].params, c.S, gen, c.train, vm, v1, v2));
  HeadArgs h = head_args(c, a.value[0].params, v2, HEAD_POLICY);
  h.loss = a.losses + 2;
  return launch_critic_head(h, c.st);
}

static int phase_policy_grad(Ctx& c) {
  const recnn_step_args& a = *c.a;
  if (!a.do_policy_step) return RECNN_OK;
  const int S = c.d.state_dim, A = c.d.action_dim, H = c.d.hidden;
  float *p1 = c.ws.hb[0], *p2 = c.ws.hb[1], *v1 = c.ws.hb[2], *v2 = c.ws.hb[3], *dv2 = c.ws.hb[4],
        *dv1 = c.ws.hb[5];
  float* dgen = c.ws.ab[2];
  const float* Pc = a.value[0].params;
  const float* Pa = a.policy.params;
  float* G = a.policy.grads;
  RECNN_REQUIRE(G != nullptr, "policy net needs a grad arena on a policy step");
  const float dq = -1.0f / (float)a.n_rows_global;            // d(-mean q)/dq
  // through the critic, input-gradient only, and only the action slice of layer 1
  RECNN_PROPAGATE(launch_critic_head_bwd(nullptr, dq, Pc + c.lc.w3, v2, c.gate, dv2, c.n, H, c.st));
  RECNN_PROPAGATE(backprop_hidden(dv2, H, Pc + c.lc.w2, H, H, c.n, v1, c.gate, dv1, c.st));
  RECNN_PROPAGATE(backprop_hidden(dv1, H, Pc + c.lc.w1 + S, S + A, A, c.n, nullptr, 1.f, dgen, c.st));
  // actor backward
  RECNN_PROPAGATE(weight_grad(dgen, A, mat(p2, H), H, c.n, G + c.la.w3, G + c.la.b3, c.ws, c.st));
  float* dp2 = dv2;   // dv2/dv1 are dead once dgen exists
  float* dp1 = dv1;
  RECNN_PROPAGATE(backprop_hidden(dgen, A, Pa + c.la.w3, H, H, c.n, p2, c.gate, dp2, c.st));
  RECNN_PROPAGATE(weight_grad(dp2, H, mat(p1, H), H, c.n, G + c.la.w2, G + c.la.b2, c.ws, c.st));
  RECNN_PROPAGATE(backprop_hidden(dp2, H, Pa + c.la.w2, H, H, c.n, p1, c.gate, dp1, c.st));
  RECNN_PROPAGATE(weight_grad(dp1, H, mat(c.S, S), S, c.n, G + c.la.w1, G + c.la.b1, c.ws, c.st));
  return RECNN_OK;
}

static int phase_policy_opt(Ctx& c) {
  const recnn_step_args& a = *c.a;
  if (!a.do_policy_step) return RECNN_OK;
  float* coef = c.ws.scalars;
  // clip_grad_norm_(policy params, max_norm=-1, norm_type=1)   (ddpg.py:92, td3.py:133)
  RECNN_PROPAGATE(launch_l1_clip_coef(a.policy.grads, c.la.count, -1.0f, coef, a.losses + 3,
                                      c.ws.block_partials, c.ws.tickets + 1, c.st));
  if (a.policy_optim.kind == RECNN_OPT_EXTERNAL)
    return launch_scale_inplace(a.policy.grads, c.la.count, coef, c.st);
  return launch_optimizer(a.policy_optim, a.policy, c.la.count, coef, c.st);
}

static int phase_soft_update(Ctx& c) {
  const recnn_step_args& a = *c.a;
  if (!a.do_policy_step) return RECNN_OK;
  const int n_critics = a.algo == RECNN_ALGO_TD3 ? 2 : 1;
  for (int i = 0; i < n_critics; ++i)        // ddpg.py:95-97 / td3.py:136-141
    RECNN_PROPAGATE(launch_polyak(a.target_value[i].params, a.value[i].params, c.lc.count, a.soft_tau, c.st));
  if (a.algo == RECNN_ALGO_DDPG)              // ddpg.py:98-100; TD3 never updates its target policy
    RECNN_PROPAGATE(launch_polyak(a.target_policy.params, a.policy.params, c.la.count, a.soft_tau, c.st));
  return RECNN_OK;
}

static int run_step(const recnn_step_args* a, int algo, void* stream) {
  RECNN_REQUIRE(a != nullptr, "args");
  RECNN_REQUIRE(a->algo == algo, "args->algo does not match the entry point");
  RECNN_REQUIRE(a->n_rows > 0 && a->n_rows_global >= a->n_rows, "n_rows");
  RECNN_REQUIRE(a->dims.state_dim > 0 && a->dims.action_dim > 0 && a->dims.hidden > 0, "dims");
  RECNN_REQUIRE(a->n_rows < (1ll << 31) / (a->dims.state_dim + a->dims.action_dim + 1), "n_rows too large for int32 tile indexing");
  RECNN_REQUIRE(a->losses && a->workspace, "losses/workspace");
  const bool frames = a->table != nullptr;
  if (frames) {
    RECNN_REQUIRE(a->items && a->ratings && a->n_items > 0 && a->frame > 0 && a->emb_dim > 0, "frame-form batch");
    RECNN_REQUIRE(a->dims.state_dim == a->frame * a->emb_dim + a->frame && a->dims.action_dim == a->emb_dim,
                  "state_dim/action_dim do not match frame*dim+frame / dim");
  } else {
    RECNN_REQUIRE(a->state && a->next_state && a->action && a->reward, "dense batch needs state/next_state/action/reward");
  }
  RECNN_REQUIRE(a->done != nullptr, "done");
  const int n_critics = algo == RECNN_ALGO_TD3 ? 2 : 1;
  RECNN_REQUIRE(a->policy.params && a->target_policy.params, "policy nets");
  for (int i = 0; i < n_critics; ++i)
    RECNN_REQUIRE(a->value[i].params && a->target_value[i].params, "value nets");
  if (a->dropout && !a->masks[0]) RECNN_REQUIRE(a->rng_step != nullptr, "perf-mode dropout needs rng_step");
  if (algo == RECNN_ALGO_TD3 && !a->noise) RECNN_REQUIRE(a->rng_step != nullptr, "perf-mode noise needs rng_step");

  Ctx c;
  c.a = a;
  c.d = a->dims;
  c.la = actor_layout(c.d);
  c.lc = critic_layout(c.d);
  c.n = a->n_rows;
  c.st = static_cast<cudaStream_t>(stream);
  c.ws = carve(c.d, c.n, a->workspace);
  if (c.ws.bytes > a->workspace_bytes) {
    set_error("workspace too small: need %lld bytes, got %lld", (long long)c.ws.bytes, (long long)a->workspace_bytes);
    return RECNN_E_WORKSPACE;
  }
  c.rng.masks = a->masks[0] ? a->masks : nullptr;
  c.rng.seed = a->seed;
  c.rng.step = (const long long*)a->rng_step;
  c.train = a->dropout != 0;
  c.gate = c.train ? 2.0f : 1.0f;
  c.DONE = a->done;
  // tickets of the deterministic two-level reductions start at zero
  RECNN_CHECK_CUDA(cudaMemsetAsync(c.ws.tickets, 0, 8 * sizeof(unsigned), c.st));
  if (frames) {
    if (a->phases & RECNN_PH_GATHER)
    RECNN_PROPAGATE(recnn_frame_gather(a->table, a->n_items, a->emb_dim, a->items, a->ratings, c.n, a->frame,
                                       c.ws.S, c.ws.S2, c.ws.ACT, c.ws.REW, nullptr, stream));
    c.S = c.ws.S; c.S2 = c.ws.S2; c.ACT = c.ws.ACT;
    c.REW = a->reward ? a->reward : c.ws.REW;
  } else {
    c.S = a->state; c.S2 = a->next_state; c.ACT = a->action; c.REW = a->reward;
  }
  if (a->phases & RECNN_PH_VALUE_GRAD) RECNN_PROPAGATE(phase_value_grad(c));
  if (a->phases & RECNN_PH_VALUE_OPT) RECNN_PROPAGATE(phase_value_opt(c));
  if (a->phases & RECNN_PH_POLICY_LOSS) RECNN_PROPAGATE(phase_policy_loss(c));
  if (a->phases & RECNN_PH_POLICY_GRAD) RECNN_PROPAGATE(phase_policy_grad(c));
  if (a->phases & RECNN_PH_POLICY_OPT) RECNN_PROPAGATE(phase_policy_opt(c));
  if (a->phases & RECNN_PH_SOFT_UPDATE) RECNN_PROPAGATE(phase_soft_update(c));
  return RECNN_OK;
}

}  // namespace recnn

using namespace recnn;

extern "C" int64_t recnn_step_workspace_bytes(const recnn_dims* d, int64_t n_rows, int32_t algo) {
  (void)algo;
  if (!d || n_rows <= 0) return 0;
  return carve(*d, n_rows, nullptr).bytes;
}

extern "C" int recnn_ddpg_step(const recnn_step_args* args, void* stream) {
  return run_step(args, RECNN_ALGO_DDPG, stream);
}

extern "C" int recnn_td3_step(const recnn_step_args* args, void* stream) {
  return run_step(args, RECNN_ALGO_TD3, stream);
}

extern "C" int64_t recnn_actor_param_count(const recnn_dims* d) { return d ? actor_layout(*d).count : 0; }
extern "C" int64_t recnn_critic_param_count(const recnn_dims* d) { return d ? critic_layout(*d).count : 0; }

extern "C" int recnn_actor_forward(const recnn_dims* d, const float* params, const float* state, int64_t n_rows,
                                   const uint8_t* mask1, const uint8_t* mask2, int apply_tanh, float* action_out,
                                   float* scratch, void* stream) {
  RECNN_REQUIRE(d && params && state && action_out && scratch, "null pointer");
  RECNN_REQUIRE((mask1 == nullptr) == (mask2 == nullptr), "give both masks or neither");
  if (n_rows <= 0) return RECNN_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const NetLayout l = actor_layout(*d);
  const int H = d->hidden;
  float* h1 = scratch;
  float* h2 = scratch + n_rows * H;
  const uint8_t* masks[2] = {mask1, mask2};
  Rng rng = {masks, 0, nullptr};
  const bool train = mask1 != nullptr;
  RECNN_PROPAGATE(hidden_layer(mat(state, d->state_dim), d->state_dim, params + l.w1, params + l.b1, H, n_rows,
                               train, mask1, rng, 0, h1, st));
  RECNN_PROPAGATE(hidden_layer(mat(h1, H), H, params + l.w2, params + l.b2, H, n_rows, train, mask2, rng, 1, h2, st));
  return linear_out(h2, H, params + l.w3, params + l.b3, d->action_dim, n_rows, apply_tanh, nullptr, action_out, st);
}

extern "C" int recnn_critic_forward(const recnn_dims* d, const float* params, const float* state,
                                    const float* action, int64_t n_rows, const uint8_t* mask1,
                                    const uint8_t* mask2, float* value_out, float* scratch, void* stream) {
  RECNN_REQUIRE(d && params && state && action && value_out && scratch, "null pointer");
  RECNN_REQUIRE((mask1 == nullptr) == (mask2 == nullptr), "give both masks or neither");
  if (n_rows <= 0) return RECNN_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const NetLayout l = critic_layout(*d);
  const int H = d->hidden, S = d->state_dim, A = d->action_dim;
  float* h1 = scratch;
  float* h2 = scratch + n_rows * H;
  const uint8_t* masks[2] = {mask1, mask2};
  Rng rng = {masks, 0, nullptr};
  const bool train = mask1 != nullptr;
  RECNN_PROPAGATE(hidden_layer(mat_cat(state, S, S, action, A), S + A, params + l.w1, params + l.b1, H, n_rows,
                               train, mask1, rng, 0, h1, st));
  RECNN_PROPAGATE(hidden_layer(mat(h1, H), H, params + l.w2, params + l.b2, H, n_rows, train, mask2, rng, 1, h2, st));
  HeadArgs h;
  h.h2 = h2; h.w3 = params + l.w3; h.b3 = params + l.b3; h.n_rows = n_rows; h.n_rows_global = n_rows;
  h.hidden = H; h.mode = HEAD_PLAIN; h.reward = nullptr; h.done = nullptr; h.gamma = 0; h.min_value = 0;
  h.max_value = 0; h.y = nullptr; h.tmp = nullptr; h.out = value_out; h.dq = nullptr; h.loss = nullptr;
  h.block_partials = nullptr; h.ticket = nullptr;
  return launch_critic_head(h, st);
}

extern "C" int recnn_linear_forward(const float* x, int64_t n_rows, int in_dim, const float* weight,
                                    const float* bias, int out_dim, int relu, float* out, void* stream) {
  RECNN_REQUIRE(x && weight && bias && out, "null pointer");
  RECNN_REQUIRE(in_dim > 0 && out_dim > 0 && n_rows >= 0, "sizes");
  if (n_rows == 0) return RECNN_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (relu) {
    Rng rng = {nullptr, 0, nullptr};
    return hidden_layer(mat(x, in_dim), in_dim, weight, bias, out_dim, n_rows, false, nullptr, rng, 0, out, st);
  }
  return linear_out(x, in_dim, weight, bias, out_dim, n_rows, 0, nullptr, out, st);
}
