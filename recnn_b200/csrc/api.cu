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

// Runtime switches.  Defaults are the measured-best variants (profiles/README.md); the environment
// variables are read once, recnn_debug_set_option() overrides them at any time (used by the A/B
// legs of bench.py and by tests that check the variants against each other).
static std::atomic<int> g_options[OPT_COUNT];
static std::atomic<bool> g_options_init{false};
static void init_options() {
  if (g_options_init.load(std::memory_order_acquire)) return;
  const struct { Option o; const char* env; int def; } table[] = {
      {OPT_GATHER_VARIANT, "RECNN_B200_GATHER", 0},      // 0: one warp per row   1: balanced (row, slot) units
                                                         // 2: as 0 with 16-byte stores where the pitches allow (in-step)
      {OPT_PRESPLIT, "RECNN_B200_PRESPLIT", 0},          // 1: weights pre-split into TF32 hi/lo planes
      {OPT_WORKERS16, "RECNN_B200_WORKERS16", 0},        // 1: 64-wide GEMM tiles run 16 worker warps (4 groups)
                                                         // 2: ... and every GEMM uses 64-wide tiles
      {OPT_LO2, "RECNN_B200_LO2", 0},                    // 1: 64-wide tiles keep two cross-term accumulators
      {OPT_BN64, "RECNN_B200_BN64", 0},                  // 1: every GEMM of the step uses 64-wide tiles
      {OPT_LEAN, "RECNN_B200_LEAN", 0},                  // 1: GEMM kernels without experiment hooks, running counters
                                                         //    in the MMA warp (unvalidated on hardware: round 2)
      {OPT_TAIL, "RECNN_B200_TAIL", 0},                  // 1: dZ column sums run on the side stream beside the dW GEMMs
      {OPT_DWSPLIT, "RECNN_B200_DWSPLIT", 0},            // 1: split-K of the weight gradients sized for ONE wave of CTAs
      {OPT_PADZERO, "RECNN_B200_PADZERO", 0},            // 1: zero only the pad columns of the action images (one kernel)
      {OPT_PDL, "RECNN_B200_PDL", 0},                    // 1: LEAN GEMMs are launched with programmatic stream
                                                         //    serialization (prologue overlaps the predecessor's tail)
  };
  for (const auto& t : table) {
    const char* e = getenv(t.env);
    g_options[t.o].store(e && *e ? atoi(e) : t.def, std::memory_order_relaxed);
  }
  g_options_init.store(true, std::memory_order_release);
}
int option(Option o) {
  init_options();
  return g_options[o].load(std::memory_order_relaxed);
}
}  // namespace recnn

// name in {"gather_variant", "presplit", "workers16"}; returns the previous value, or -1 for an unknown name
extern "C" RECNN_API int recnn_debug_set_option(const char* name, int value) {
  recnn::init_options();
  if (!name) return -1;
  int idx = -1;
  if (strcmp(name, "gather_variant") == 0) idx = recnn::OPT_GATHER_VARIANT;
  else if (strcmp(name, "presplit") == 0) idx = recnn::OPT_PRESPLIT;
  else if (strcmp(name, "workers16") == 0) idx = recnn::OPT_WORKERS16;
  else if (strcmp(name, "lo2") == 0) idx = recnn::OPT_LO2;
  else if (strcmp(name, "bn64") == 0) idx = recnn::OPT_BN64;
  else if (strcmp(name, "lean") == 0) idx = recnn::OPT_LEAN;
  else if (strcmp(name, "pdl") == 0) idx = recnn::OPT_PDL;
  else if (strcmp(name, "tail") == 0) idx = recnn::OPT_TAIL;
  else if (strcmp(name, "dwsplit") == 0) idx = recnn::OPT_DWSPLIT;
  else if (strcmp(name, "padzero") == 0) idx = recnn::OPT_PADZERO;
  if (idx < 0) return -1;
  return recnn::g_options[idx].exchange(value);
}

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
