// Per-device caches and experiment switches of libcnf_b200 (see cnf_common.h).  Everything here is filled once
// (per device / per process) under a mutex and read-only afterwards.
#include <cuda_runtime.h>

#include <cstdlib>
#include <mutex>

#include "cnf_common.h"

namespace {
constexpr int kMaxDev = 64, kMaxFuncs = 256;
struct FuncSlot { const void* func; int bytes; };
struct DevState {
  bool ready = false;
  CnfDevInfo info{};
  FuncSlot funcs[kMaxFuncs];
  int n_funcs = 0;
};
DevState g_dev[kMaxDev];
std::mutex g_mu;

const char* const kSwitchNames[CNF_SW_COUNT] = {
    "CNF_NO_ZEROCOPY", "CNF_DEEP_APPLY", "CNF_DEEP_TRAIN", "CNF_FORCE_LEAN", "CNF_FP32R", "CNF_FP32_NO_WL",
    "CNF_FP32_NT", "CNF_FP32_WS", "CNF_NO_LEAN_TRAIN", "CNF_SPLIT_GENERIC", "CNF_SPLIT_SEQ", "CNF_SPLIT_TRAIN",
    "CNF_TC_EPI", "CNF_TC_GENERIC", "CNF_METRICS_STAGES", "CNF_FP32R_TRAIN"};
struct Switches {
  bool live;
  const char* v[CNF_SW_COUNT];
  Switches() {
    live = getenv("CNF_LIVE_ENV") != nullptr;
    for (int i = 0; i < CNF_SW_COUNT; ++i) v[i] = getenv(kSwitchNames[i]);
  }
};
}  // namespace

const char* cnf_switch(CnfSwitch s) {
  static const Switches sw;       // thread-safe one-time initialisation
  if (sw.live) return getenv(kSwitchNames[s]);
  return sw.v[s];
}

static int dev_state(DevState** out) {
  int dev = 0;
  CNF_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDev) { cnf_set_error("device id %d out of range", dev); return CNF_E_CUDA; }
  DevState& st = g_dev[dev];
  if (!st.ready) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (!st.ready) {
      int v = 0, s = 0;
      CNF_CHECK_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
      CNF_CHECK_CUDA(cudaDeviceGetAttribute(&s, cudaDevAttrMultiProcessorCount, dev));
      st.info.id = dev; st.info.sms = s; st.info.max_smem = v;
      st.ready = true;
    }
  }
  *out = &st;
  return CNF_OK;
}

int cnf_dev_info(CnfDevInfo* out) {
  DevState* st = nullptr;
  int rc = dev_state(&st);
  if (rc) return rc;
  *out = st->info;
  return CNF_OK;
}

int cnf_func_smem(const void* func, int bytes) {
  DevState* st = nullptr;
  int rc = dev_state(&st);
  if (rc) return rc;
  std::lock_guard<std::mutex> lock(g_mu);
  int slot = -1;
  for (int i = 0; i < st->n_funcs; ++i)
    if (st->funcs[i].func == func) { slot = i; break; }
  if (slot >= 0 && st->funcs[slot].bytes >= bytes) return CNF_OK;
  CNF_CHECK_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  if (slot < 0 && st->n_funcs < kMaxFuncs) slot = st->n_funcs++;
  if (slot >= 0) { st->funcs[slot].func = func; st->funcs[slot].bytes = bytes; }
  return CNF_OK;
}
