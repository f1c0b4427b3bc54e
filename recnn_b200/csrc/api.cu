// Error plumbing + version of the C ABI (include/recnn_b200.h).
#include <stdarg.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace recnn {
static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

}  // namespace recnn

extern "C" RECNN_API int64_t recnn_b200_launch_count(void) { return recnn::g_launches.load(); }
extern "C" int recnn_b200_abi_version(void) { return RECNN_B200_ABI_VERSION; }
extern "C" const char* recnn_b200_last_error(void) { return recnn::g_error; }

// Layout probes for foreign-function bindings (ctypes mirrors are checked against these).
extern "C" RECNN_API int64_t recnn_sizeof_step_args(void) { return (int64_t)sizeof(recnn_step_args); }
extern "C" RECNN_API int64_t recnn_offsetof_step_args(int field) {
  switch (field) {
    case 0: return offsetof(recnn_step_args, dims);
    case 1: return offsetof(recnn_step_args, n_rows);
    case 2: return offsetof(recnn_step_args, table);
    case 3: return offsetof(recnn_step_args, policy);
    case 4: return offsetof(recnn_step_args, policy_optim);
    case 5: return offsetof(recnn_step_args, gamma);
    case 6: return offsetof(recnn_step_args, soft_tau);
    case 7: return offsetof(recnn_step_args, masks);
    case 8: return offsetof(recnn_step_args, seed);
    case 9: return offsetof(recnn_step_args, losses);
    case 10: return offsetof(recnn_step_args, workspace_bytes);
    case 11: return offsetof(recnn_step_args, comm);
    default: return -1;
  }
}

// ---- CUDA-graph helpers (include/recnn_b200.h) ------------------------------------------------------------------
extern "C" int recnn_graph_begin(void* stream) {
  RECNN_CHECK_CUDA(cudaStreamBeginCapture(static_cast<cudaStream_t>(stream), cudaStreamCaptureModeThreadLocal));
  return RECNN_OK;
}

extern "C" int recnn_graph_end(void* stream, void** graph_out) {
  RECNN_REQUIRE(graph_out != nullptr, "graph_out");
  *graph_out = nullptr;
  cudaGraph_t g = nullptr;
  RECNN_CHECK_CUDA(cudaStreamEndCapture(static_cast<cudaStream_t>(stream), &g));
  cudaGraphExec_t exec = nullptr;
  const cudaError_t e = cudaGraphInstantiateWithFlags(&exec, g, cudaGraphInstantiateFlagUseNodePriority);
  cudaGraphDestroy(g);
  RECNN_CHECK_CUDA(e);
  *graph_out = exec;
  return RECNN_OK;
}

extern "C" int recnn_graph_launch(void* graph, void* stream) {
  RECNN_REQUIRE(graph != nullptr, "graph");
  RECNN_CHECK_CUDA(cudaGraphLaunch(static_cast<cudaGraphExec_t>(graph), static_cast<cudaStream_t>(stream)));
  return RECNN_OK;
}

extern "C" int recnn_graph_destroy(void* graph) {
  if (graph) RECNN_CHECK_CUDA(cudaGraphExecDestroy(static_cast<cudaGraphExec_t>(graph)));
  return RECNN_OK;
}
