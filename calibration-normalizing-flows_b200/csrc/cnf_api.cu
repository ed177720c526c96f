// extern "C" entry points of the flow itself: precision dispatch between the fp32 CUDA-core
// kernels (cnf_flow_fp32.cu) and the bf16 tcgen05 kernel (cnf_flow_tc.cu).  See include/cnf.h.
#include <cuda_runtime.h>

#include <cstdlib>
#include <mutex>

#include "cnf_common.h"

int cnf_fp32_apply(const cnf_flow_desc* desc, const float* packed, const int32_t* tables, const float* x, float* z,
                   float* logdet, float* zs, int64_t N, int inverse, cudaStream_t st);
int cnf_fp32_train(const cnf_flow_desc* desc, const float* packed, const int32_t* tables, const float* x,
                   const int64_t* y, const float* gz, const float* gld, float* gx, float* partials, double* loss_acc,
                   int64_t N, float eps, float gamma, float inv_n, int head, int64_t* rows_used, cudaStream_t st);
int cnf_tc_apply(const cnf_flow_desc* desc, const void* packed_tc, const int32_t* tables, const float* x, float* z,
                 float* logdet, int64_t N, int inverse, cudaStream_t st);



int cnf_fp32_predict(const cnf_flow_desc* desc, const float* packed, const int32_t* tables, const float* x, float* z,
                     float* logdet, int64_t N, const CnfTail& ta, cudaStream_t st);
int cnf_tc_predict(const cnf_flow_desc* desc, const void* packed_tc, const int32_t* tables, const float* x, float* z,
                   float* logdet, int64_t N, const CnfTail& ta, cudaStream_t st);

static int check_prec(const cnf_flow_desc* desc) {
  if (!desc) { cnf_set_error("null descriptor"); return CNF_E_ARG; }
  if (desc->precision != CNF_PREC_FP32 && desc->precision != CNF_PREC_BF16_TC) {
    cnf_set_error("unknown precision %d", desc->precision);
    return CNF_E_ARG;
  }
  return CNF_OK;
}

extern "C" int cnf_flow_forward(const cnf_flow_desc* desc, const void* packed, const int32_t* tables, const float* x,
                                float* z, float* logdet, float* zs, int64_t N, void* stream) {
  int rc = check_prec(desc);
  if (rc) return rc;
  if (desc->precision == CNF_PREC_BF16_TC) {
    if (zs) { cnf_set_error("intermediate outputs (zs) are only produced by the fp32 path"); return CNF_E_UNSUPPORTED; }
    return cnf_tc_apply(desc, packed, tables, x, z, logdet, N, 0, (cudaStream_t)stream);
  }
  return cnf_fp32_apply(desc, (const float*)packed, tables, x, z, logdet, zs, N, 0, (cudaStream_t)stream);
}

extern "C" int cnf_flow_inverse(const cnf_flow_desc* desc, const void* packed, const int32_t* tables, const float* z,
                                float* x, float* logdet, float* xs, int64_t N, void* stream) {
  int rc = check_prec(desc);
  if (rc) return rc;
  if (desc->precision == CNF_PREC_BF16_TC) {
    if (xs) { cnf_set_error("intermediate outputs (xs) are only produced by the fp32 path"); return CNF_E_UNSUPPORTED; }
    return cnf_tc_apply(desc, packed, tables, z, x, logdet, N, 1, (cudaStream_t)stream);
  }
  return cnf_fp32_apply(desc, (const float*)packed, tables, z, x, logdet, xs, N, 1, (cudaStream_t)stream);
}

extern "C" int cnf_flow_predict(const cnf_flow_desc* desc, const void* packed, const int32_t* tables, const float* x,
                                int64_t N, int32_t center, int32_t mode, const double* log_priors, float* z,
                                float* logdet, double* probs_out, const int64_t* y, int32_t bins, const double* edges,
                                double* acc, void* stream) {
  int rc = check_prec(desc);
  if (rc) return rc;
  if (N == 0) return CNF_OK;
  if (!packed || !tables || !x || N < 0) { cnf_set_error("cnf_flow_predict: null pointer / negative N"); return CNF_E_ARG; }
  if (mode != CNF_METRICS_LOGITS && mode != CNF_METRICS_CALIBRATED) { cnf_set_error("cnf_flow_predict: bad mode %d", mode); return CNF_E_ARG; }
  if (mode == CNF_METRICS_CALIBRATED && !log_priors) { cnf_set_error("cnf_flow_predict: calibrated mode needs log_priors"); return CNF_E_ARG; }
  if (probs_out && mode != CNF_METRICS_CALIBRATED) { cnf_set_error("cnf_flow_predict: probs_out needs the calibrated mode"); return CNF_E_ARG; }
  if (acc && (!y || !edges || bins < 1 || bins > 1024)) { cnf_set_error("cnf_flow_predict: statistics need labels, edges and 1 <= bins <= 1024"); return CNF_E_ARG; }
  if (!z && !probs_out && !acc) { cnf_set_error("cnf_flow_predict: no output requested"); return CNF_E_ARG; }
  if (center && desc->K > 128) { cnf_set_error("cnf_flow_predict: fused centring covers K <= 128"); return CNF_E_UNSUPPORTED; }
  CnfTail ta;
  ta.mode = mode; ta.center = center ? 1 : 0; ta.bins = acc ? bins : 1;
  ta.y = acc ? y : nullptr; ta.log_priors = log_priors; ta.edges = acc ? edges : nullptr; ta.acc = acc; ta.probs_out = probs_out;
  if (desc->precision == CNF_PREC_BF16_TC) return cnf_tc_predict(desc, packed, tables, x, z, logdet, N, ta, (cudaStream_t)stream);
  return cnf_fp32_predict(desc, (const float*)packed, tables, x, z, logdet, N, ta, (cudaStream_t)stream);
}

extern "C" int cnf_nll_train_step(const cnf_flow_desc* desc, const void* packed, const int32_t* tables, const float* x,
                                  const int64_t* y, int64_t N, float eps, float gamma, float inv_n_total,
                                  float* grad_partials, double* loss_acc, void* stream) {
  int rc = check_prec(desc);
  if (rc) return rc;
  if (desc->precision != CNF_PREC_FP32) { cnf_set_error("training runs on the fp32 path"); return CNF_E_UNSUPPORTED; }
  return cnf_fp32_train(desc, (const float*)packed, tables, x, y, nullptr, nullptr, nullptr, grad_partials, loss_acc, N,
                        eps, gamma, inv_n_total, CNF_HEAD_NLL, nullptr, (cudaStream_t)stream);
}

extern "C" int cnf_nll_train_step_rows(const cnf_flow_desc* desc, const void* packed, const int32_t* tables,
                                       const float* x, const int64_t* y, int64_t N, float eps, float gamma,
                                       float inv_n_total, float* grad_partials, double* loss_acc, int64_t* rows_used,
                                       void* stream) {
  int rc = check_prec(desc);
  if (rc) return rc;
  if (desc->precision != CNF_PREC_FP32) { cnf_set_error("training runs on the fp32 path"); return CNF_E_UNSUPPORTED; }
  if (!rows_used) { cnf_set_error("cnf_nll_train_step_rows: null rows_used"); return CNF_E_ARG; }
  return cnf_fp32_train(desc, (const float*)packed, tables, x, y, nullptr, nullptr, nullptr, grad_partials, loss_acc, N,
                        eps, gamma, inv_n_total, CNF_HEAD_NLL, rows_used, (cudaStream_t)stream);
}

extern "C" int cnf_flow_backward(const cnf_flow_desc* desc, const void* packed, const int32_t* tables, const float* x,
                                 const float* g_z, const float* g_logdet, float* g_x, float* grad_partials, int64_t N,
                                 void* stream) {
  int rc = check_prec(desc);
  if (rc) return rc;
  if (desc->precision != CNF_PREC_FP32) { cnf_set_error("backward runs on the fp32 path"); return CNF_E_UNSUPPORTED; }
  return cnf_fp32_train(desc, (const float*)packed, tables, x, nullptr, g_z, g_logdet, g_x, grad_partials, nullptr, N,
                        0.f, 0.f, 0.f, CNF_HEAD_EXTERNAL, nullptr, (cudaStream_t)stream);
}

extern "C" int cnf_flow_backward_rows(const cnf_flow_desc* desc, const void* packed, const int32_t* tables,
                                      const float* x, const float* g_z, const float* g_logdet, float* g_x,
                                      float* grad_partials, int64_t N, int64_t* rows_used, void* stream) {
  int rc = check_prec(desc);
  if (rc) return rc;
  if (desc->precision != CNF_PREC_FP32) { cnf_set_error("backward runs on the fp32 path"); return CNF_E_UNSUPPORTED; }
  if (!rows_used) { cnf_set_error("cnf_flow_backward_rows: null rows_used"); return CNF_E_ARG; }
  return cnf_fp32_train(desc, (const float*)packed, tables, x, nullptr, g_z, g_logdet, g_x, grad_partials, nullptr, N,
                        0.f, 0.f, 0.f, CNF_HEAD_EXTERNAL, rows_used, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------
// Host-buffer entry point: the whole N-sample pass with the H2D copy of x and the D2H copies of
// z / log-det inside, chunked over a small pool of internal streams so that the copy-in, the
// flow kernel and the copy-out of neighbouring chunks overlap (PCIe is full duplex).  The chunk
// loop runs here, not in Python: at ~250k samples per chunk the per-chunk host cost must stay
// in the microseconds.  Asynchronous like every other call: completion is ordered on `stream`.
// ------------------------------------------------------------------------------------------
namespace {
constexpr int kHostSlots = 4;
constexpr int kMaxDev = 64;
struct HostPool {
  bool ready = false;
  cudaStream_t s[kHostSlots];
  cudaEvent_t done[kHostSlots];
  cudaEvent_t start;
};
HostPool g_pool[kMaxDev];   // one per device, created on first use and kept for the life of the process
std::mutex g_pool_mu;

int pool_get(HostPool** out) {
  int dev = 0;
  CNF_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDev) { cnf_set_error("device id %d out of range", dev); return CNF_E_CUDA; }
  HostPool& p = g_pool[dev];
  std::lock_guard<std::mutex> lock(g_pool_mu);
  if (!p.ready) {
    for (int i = 0; i < kHostSlots; ++i) {
      CNF_CHECK_CUDA(cudaStreamCreateWithFlags(&p.s[i], cudaStreamNonBlocking));
      CNF_CHECK_CUDA(cudaEventCreateWithFlags(&p.done[i], cudaEventDisableTiming));
    }
    CNF_CHECK_CUDA(cudaEventCreateWithFlags(&p.start, cudaEventDisableTiming));
    p.ready = true;
  }
  *out = &p;
  return CNF_OK;
}
}  // namespace

extern "C" int cnf_flow_apply_host(const cnf_flow_desc* desc, const void* packed, const int32_t* tables,
                                   const float* x_host, float* z_host, float* logdet_host, int64_t N, int32_t inverse,
                                   void* workspace, int64_t workspace_bytes, int64_t chunk, void* stream) {
  int rc = check_prec(desc);
  if (rc) return rc;
  if (N == 0) return CNF_OK;
  if (!packed || !tables || !x_host || !z_host || !logdet_host || !workspace || N < 0 || chunk < 1) {
    cnf_set_error("cnf_flow_apply_host: bad argument");
    return CNF_E_ARG;
  }
  const int K = desc->K;
  const int64_t per_slot = chunk * (2 * (int64_t)K + 1) * (int64_t)sizeof(float);
  int slots = (int)(workspace_bytes / per_slot);
  if (slots > kHostSlots) slots = kHostSlots;
  if (slots < 1) { cnf_set_error("cnf_flow_apply_host: workspace smaller than one chunk (%lld bytes needed)", (long long)per_slot); return CNF_E_ARG; }
  cudaStream_t user = (cudaStream_t)stream;
  // Zero-copy fast path: when all three host buffers are pinned (device-addressable under UVA) and
  // the tensor-core kernel runs, launch it ONCE on the host pointers.  Its tile prefetch
  // (cp.async, one tile ahead per slot, ~2 MB in flight) hides the PCIe latency and both
  // directions stream concurrently, without per-chunk DMA set-up costs: 1.18 ms vs 1.58 ms per
  // 10^6 K=10 samples on B200 (46 GB/s per direction when both are busy).
  if (!cnf_switch(CNF_SW_NO_ZEROCOPY)) {
    cudaPointerAttributes ax, az, al;
    const bool ok = cudaPointerGetAttributes(&ax, x_host) == cudaSuccess && ax.type == cudaMemoryTypeHost &&
                    cudaPointerGetAttributes(&az, z_host) == cudaSuccess && az.type == cudaMemoryTypeHost &&
                    cudaPointerGetAttributes(&al, logdet_host) == cudaSuccess && al.type == cudaMemoryTypeHost;
    cudaGetLastError();   // a pageable pointer makes cudaPointerGetAttributes report an error on old drivers
    // Measured on one B200 (round 2, 10^7 samples per call): the copy-engine pipeline below with 2^19-row chunks carries
    // 1.00 G samples/s, the zero-copy launch 0.90; the fp32 register kernel on host pointers only 0.58 (its 8-byte row
    // reads make small PCIe requests).  Zero-copy stays for the tensor-core kernel on calls of up to 2^21 rows, where
    // the pipeline cannot fill (1.18 vs 1.58 ms per 10^6 rows).
    if (ok && desc->precision == CNF_PREC_BF16_TC && N <= (1 << 21))
      return cnf_tc_apply(desc, packed, tables, (const float*)ax.devicePointer, (float*)az.devicePointer,
                          (float*)al.devicePointer, N, inverse, user);
  }
  HostPool* poolp = nullptr;
  if ((rc = pool_get(&poolp))) return rc;
  HostPool& pool = *poolp;
  CNF_CHECK_CUDA(cudaEventRecord(pool.start, user));          // weights were packed on the caller's stream
  for (int i = 0; i < slots; ++i) CNF_CHECK_CUDA(cudaStreamWaitEvent(pool.s[i], pool.start, 0));
  int64_t lo = 0;
  for (int64_t c = 0; lo < N; ++c, lo += chunk) {
    const int64_t n = (N - lo < chunk) ? N - lo : chunk;
    const int i = (int)(c % slots);
    cudaStream_t st = pool.s[i];
    float* xin = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + (size_t)i * per_slot);
    float* zout = xin + chunk * K;
    float* ld = zout + chunk * K;
    CNF_CHECK_CUDA(cudaMemcpyAsync(xin, x_host + lo * K, (size_t)n * K * sizeof(float), cudaMemcpyHostToDevice, st));
    if (desc->precision == CNF_PREC_BF16_TC) rc = cnf_tc_apply(desc, packed, tables, xin, zout, ld, n, inverse, st);
    else rc = cnf_fp32_apply(desc, (const float*)packed, tables, xin, zout, ld, nullptr, n, inverse, st);
    if (rc) return rc;
    CNF_CHECK_CUDA(cudaMemcpyAsync(z_host + lo * K, zout, (size_t)n * K * sizeof(float), cudaMemcpyDeviceToHost, st));
    CNF_CHECK_CUDA(cudaMemcpyAsync(logdet_host + lo, ld, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  for (int i = 0; i < slots; ++i) {
    CNF_CHECK_CUDA(cudaEventRecord(pool.done[i], pool.s[i]));
    CNF_CHECK_CUDA(cudaStreamWaitEvent(user, pool.done[i], 0));
  }
  return CNF_OK;
}
