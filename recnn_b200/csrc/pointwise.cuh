// Declarations of the non-GEMM kernels' host launchers (pointwise.cu).
#pragma once
#include "common.cuh"

namespace recnn {

enum HeadMode {
  HEAD_PLAIN = 0,        // out[n] = q
  HEAD_TARGET_DDPG = 1,  // y[n] = clamp(r + (1-d)*gamma*q, min, max)            misc.py:30-35
  HEAD_TARGET_TD3_A = 2, // tmp[n] = q                                            td3.py:81
  HEAD_TARGET_TD3_B = 3, // y[n] = r + (1-d)*gamma*min(tmp[n], q)  (no clamp)     td3.py:82-86
  HEAD_VALUE = 4,        // diff=q-y; dq[n]=2*diff/Ng; loss += diff^2/Ng          misc.py:37-39
  HEAD_POLICY = 5        // loss += -q/Ng                                          ddpg.py:79,87
};

struct HeadArgs {
  const float* h2;      // [N, H]
  const float* w3;      // [H]
  const float* b3;      // [1]
  int64_t n_rows;
  int64_t n_rows_global;
  int hidden;
  int mode;
  const float* reward;  // [N]
  const float* done;    // [N]
  float gamma, min_value, max_value;
  float* y;             // TD target in/out
  float* tmp;           // TD3 first critic
  float* out;           // HEAD_PLAIN / q copy (may be null)
  float* dq;            // HEAD_VALUE
  float* loss;          // device scalar to write (VALUE / POLICY)
  float* block_partials;  // >= gridDim floats
  unsigned* ticket;       // zero-initialised, self-resetting
};
int launch_critic_head(const HeadArgs& a, cudaStream_t st);

// Everything between the critic's last hidden layer and its backward GEMMs in ONE kernel (was five launches
// on the step's critical path): q = h2 w3 + b3; TD target y from the target critic's hidden layer (DDPG,
// misc.py:28-35) or a precomputed y (TD3's min of two targets); MSE loss; dq = 2 (q - y) / N_global;
// dz2 = dq w3 * gate(h2); dW3 = dq^T h2 and db3 = sum dq (deterministic two-level reductions).
struct ValueHeadArgs {
  const float *h2, *w3, *b3;         // online critic: hidden2 [N, H], head weights
  const float *th2, *tw3, *tb3;      // target critic (fused TD target) or null (y is an input)
  const float *reward, *done;
  float gamma, min_value, max_value;
  float* y;                          // [N] TD target: written when th2 != null, read otherwise
  int64_t n_rows, n_rows_global;
  int hidden;
  int learn;                         // 0: loss only
  float gate_scale;
  float* dz2;                        // [N, H]
  float* gw3;                        // [H] gradient of the head weights
  float* gb3;                        // [1]
  float* loss;
  float* block_partials;             // >= 148 * (H + 2) floats
  unsigned* ticket;
};
bool value_head_fusable(int hidden);
int launch_value_head_fused(const ValueHeadArgs& a, cudaStream_t st);

// dz2[n,c] = dq_n * w3[c] * (h2[n,c] > 0 ? gate_scale : 0); dq_n = dq ? dq[n] : dq_const
int launch_critic_head_bwd(const float* dq, float dq_const, const float* w3, const float* h2,
                           float gate_scale, float* dz2, int64_t n_rows, int hidden, cudaStream_t st);

// grad arena <- sum over split-K partials; part is [splits][C][K1] where column K1-1 is the
// bias gradient: w_dst[c*ldw+k] for k<K1-1 (ldw = arena row pitch), b_dst[c] for k==K1-1.
int launch_reduce_partials(const float* part, int splits, int C, int K1, float* w_dst, long long ldw,
                           float* b_dst, cudaStream_t st);
int launch_head_grad_partials(const float* dq, const float* h2, int64_t n_rows, int H, int64_t rows_per_split,
                              int splits, float* part, cudaStream_t st);
int launch_colsum_partials(const float* dz, int64_t n_rows, int C, int64_t rows_per_split, int splits,
                           float* part, int K1, cudaStream_t st);

// *coef = min(max_norm / (||g||_1 + 1e-6), 1)  (clip_grad_norm_(.., -1, 1): ddpg.py:92);  *l1_out = ||g||_1
int launch_l1_clip_coef(const float* grads, int64_t count, float max_norm, float* coef, float* l1_out,
                        float* block_partials, unsigned* ticket, cudaStream_t st);

// comm.cu: buf <- sum over the communicator's ranks (in place, rank order).  With coef != nullptr also the
// L1 norm of the summed gradient and its clip coefficient, as launch_l1_clip_coef would give them.
int launch_comm_allreduce(const recnn_comm* comm, float* buf, int64_t n, float max_norm, float* coef, float* l1_out,
                          float* block_partials, cudaStream_t st);

// zero the `lead` leading and the trailing pad columns of up to three [n_rows, ld] buffers whose data columns are
// [lead, lead + cols) (the lead-padded action images of the step); one launch instead of three full memsets
int launch_zero_pad_columns(float* b0, float* b1, float* b2, int64_t n_rows, int ld, int lead, int cols, cudaStream_t st);
int launch_scale_inplace(float* x, int64_t count, const float* scale, cudaStream_t st);

// ticket: zero-initialised self-resetting counter; when given, the kernel itself increments *net.opt_t
int launch_optimizer(const recnn_optim& o, const recnn_net& net, int64_t count, const float* grad_scale,
                     cudaStream_t st, unsigned* ticket = nullptr);

int launch_bump64(long long* t, cudaStream_t st);
int launch_polyak(float* target, const float* net, int64_t count, double tau, cudaStream_t st);

}  // namespace recnn
