// Shared host/device definitions for libcnf_b200 (not part of the public ABI).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/cnf.h"

#define CNF_CH 16          // hidden-unit chunk held in registers by the fp32 kernels
#define CNF_GRAD_ROWS 296      // 2 x 148 SMs: rows of the gradient partial buffer the per-CTA kernels use
#define CNF_GRAD_ROWS_MAX 1184 // rows allocated: the register-resident training kernel owns one row per WARP (4 x 296)
#define CNF_GRAD_BUDGET_FLOATS (64ll << 20)

// Everything the kernels need to know about the model, derived from cnf_flow_desc.
// Passed by value as a kernel parameter.
struct CnfDims {
  int K, L, m, d0, d1, d0p;
  int nets;     // bit0: s-net, bit1: t-net
  int n_nets;   // popcount(nets)
  int H[CNF_MAX_HIDDEN];   // true hidden widths
  int Hp[CNF_MAX_HIDDEN];  // padded to CNF_CH
  int Hmax;                // max padded width (0 if m == 0)
  int w_off[CNF_MAX_HIDDEN + 1];  // float offsets inside one net block
  int b_off[CNF_MAX_HIDDEN + 1];
  int net_stride, layer_stride, n_packed;
  int tab_pi, tab_cond, tab_trans, n_tables;  // offsets inside the int32 tables
  int n_flat;
  int grad_rows;       // rows used by kernels whose CTAs own a row each (<= CNF_GRAD_ROWS)
  int grad_rows_max;   // rows of the partial buffer (plan info n_grad_rows; <= CNF_GRAD_ROWS_MAX, budget-limited)
};

#ifndef __CUDACC__
#define CNF_HD
#else
#define CNF_HD __host__ __device__
#endif
// What the fused tail of a flow kernel needs (kernel parameter, by value).  mode: CNF_METRICS_LOGITS
// (statistics of softmax(z)) or CNF_METRICS_CALIBRATED (softmax(log(softmax(z)+1e-7) - log_priors)).
struct CnfTail {
  int mode, center, bins;
  const int64_t* y;           // labels or NULL
  const double* log_priors;   // [K] (calibrated mode)
  const double* edges;        // [bins+1] or NULL (no ECE bins)
  double* acc;                // [3*bins+3] or NULL (no statistics)
  double* probs_out;          // [N,K] calibrated probabilities or NULL
};

// shared-memory bytes of the tail's per-CTA state: conf sums, log priors, float edges, counts, correct
CNF_HD inline int cnf_tail_smem_bytes(int bins, int K) {
  return ((bins * 8 + K * 8 + (bins + 2) * 4 + bins * 8) + 15) / 16 * 16;
}

void cnf_set_error(const char* fmt, ...);

// ---- per-device state (cnf_device.cu) -------------------------------------------------------------------
// The library keeps no mutable state besides these caches, all keyed by CUDA device id and filled once under a
// mutex: device limits, the dynamic-shared-memory attribute already granted to each kernel, and the internal
// stream pool of cnf_flow_apply_host.  Any number of devices per process; calls act on the CURRENT device.
struct CnfDevInfo { int id, sms, max_smem; };
int cnf_dev_info(CnfDevInfo* out);
// cudaFuncSetAttribute(func, MaxDynamicSharedMemorySize, bytes) once per (device, kernel): repeated launches with
// a size already granted cost one table lookup.
int cnf_func_smem(const void* func, int bytes);
template <typename Kern>
static inline int cnf_kernel_smem(Kern k, size_t bytes) { return cnf_func_smem(reinterpret_cast<const void*>(k), (int)bytes); }

// Experiment switches (environment variables, see NOTES.md): looked up ONCE per process and cached -- the launch
// path never calls getenv -- unless CNF_LIVE_ENV is set (tests and micro-benchmarks flip switches between calls).
enum CnfSwitch {
  CNF_SW_NO_ZEROCOPY, CNF_SW_DEEP_APPLY, CNF_SW_DEEP_TRAIN, CNF_SW_FORCE_LEAN, CNF_SW_FP32R, CNF_SW_FP32_NO_WL,
  CNF_SW_FP32_NT, CNF_SW_FP32_WS, CNF_SW_NO_LEAN_TRAIN, CNF_SW_SPLIT_GENERIC, CNF_SW_SPLIT_SEQ, CNF_SW_SPLIT_TRAIN,
  CNF_SW_TC_EPI, CNF_SW_TC_GENERIC, CNF_SW_METRICS_STAGES, CNF_SW_FP32R_TRAIN, CNF_SW_COUNT
};
const char* cnf_switch(CnfSwitch s);   // value of the switch or nullptr
int cnf_make_dims(const cnf_flow_desc* desc, CnfDims* out);
long long cnf_tc_blob_bytes(const cnf_flow_desc* desc, const CnfDims& d);

static inline int cnf_round_up(int x, int m) { return (x + m - 1) / m * m; }

#define CNF_CHECK_CUDA(expr)                                                        \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      cnf_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return CNF_E_CUDA;                                                            \
    }                                                                               \
  } while (0)

#ifdef __CUDACC__
// torch.optim.Adam's update of one entry (no weight decay), every rounding spelled out so that each kernel that
// applies it -- cnf_adam_step, cnf_adam_step_dev and the one-launch tails -- produces the same bits
// (calibrators.py:259, 295: exp_avg.lerp_(grad, 1-b1); exp_avg_sq.mul_(b2).addcmul_(grad, grad, 1-b2);
//  p -= (lr/bc1) * exp_avg / (sqrt(exp_avg_sq)/sqrt(bc2) + eps)).
__device__ __forceinline__ float cnf_adam_entry(float p, float g, float& m, float& v, float lr_over_bc1,
                                                float inv_sqrt_bc2, float b1, float b2, float eps) {
  m = __fmaf_rn(__fsub_rn(g, m), 1.f - b1, m);
  v = __fmaf_rn(__fmul_rn(1.f - b2, g), g, __fmul_rn(v, b2));
  const float denom = __fmaf_rn(sqrtf(v), inv_sqrt_bc2, eps);
  return __fmaf_rn(-lr_over_bc1, __fdiv_rn(m, denom), p);
}
#endif
