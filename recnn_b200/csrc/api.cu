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
