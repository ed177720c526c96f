// ECE / NLL / accuracy sufficient statistics in one pass over [N,K] (sm_100a), and the
// Calibrator.predict tail.  HBM-bound: each CTA streams a contiguous tile of rows with
// coalesced 128-bit loads into (skewed) shared memory, one thread then owns one row.
//
// Reference arithmetic restated here:
//   expected_calibration_error   utils/metrics.py:35-73  (right-closed bins (lo, hi],
//                                edges i*width compared in the array's dtype)
//   neg_log_likelihood           utils/metrics.py:6-15   (-log(p_y + 1e-7))
//   accuracy                     utils/metrics.py:76-80  (first arg-max, as np.argmax)
//   predict_post / predict       calibrators.py:350-353, 40-44
//                                softmax(log(softmax(z) + 1e-7) - log_priors): the inner
//                                softmax and log run in float32 (scipy on float32 logits,
//                                NumPy weak-scalar promotion), the outer softmax in float64.
#include <cuda_runtime.h>

#include "cnf_common.h"

namespace {

__device__ __forceinline__ int skew(int a) { return a + (a >> 5); }

template <typename T> __device__ __forceinline__ T t_log(T v);
template <> __device__ __forceinline__ float t_log<float>(float v) { return logf(v); }
template <> __device__ __forceinline__ double t_log<double>(double v) { return log(v); }

struct BinCache {
  int bin; unsigned cnt, correct; double sconf;
};

__device__ __forceinline__ void flush(BinCache& c, unsigned* s_cnt, unsigned* s_cor, double* s_conf) {
  if (c.bin >= 0 && c.cnt) {
    atomicAdd(s_cnt + c.bin, c.cnt);
    atomicAdd(s_cor + c.bin, c.correct);
    atomicAdd(s_conf + c.bin, c.sconf);
  }
  c.cnt = 0; c.correct = 0; c.sconf = 0.0;
}

// T = element type of the input rows; mode as CNF_METRICS_*.
template <typename T, bool STAGED>
__global__ void metrics_kernel(const T* __restrict__ in, const int64_t* __restrict__ y, int64_t N, int K, int bins,
                               int mode, const double* __restrict__ log_priors, const double* __restrict__ edges,
                               double* __restrict__ acc, double* __restrict__ probs_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int NT = blockDim.x, tid = threadIdx.x;
  // layout: [edges T (bins+1)] [s_conf double bins] [s_cnt u32 bins] [s_cor u32 bins] [lp double K] [tile]
  double* s_conf = reinterpret_cast<double*>(smem_raw);
  double* s_lp = s_conf + bins;
  T* s_edges = reinterpret_cast<T*>(s_lp + K);
  unsigned* s_cnt = reinterpret_cast<unsigned*>(s_edges + (bins + 2));
  unsigned* s_cor = s_cnt + bins;
  size_t off = reinterpret_cast<unsigned char*>(s_cor + bins) - smem_raw;
  off = (off + 15) / 16 * 16;
  T* tile = reinterpret_cast<T*>(smem_raw + off);
  __shared__ double red[32];

  for (int i = tid; i < bins; i += NT) { s_conf[i] = 0.0; s_cnt[i] = 0u; s_cor[i] = 0u; }
  if (edges != nullptr)
    for (int i = tid; i <= bins; i += NT) s_edges[i] = (T)edges[i];
  if (log_priors != nullptr)
    for (int i = tid; i < K; i += NT) s_lp[i] = log_priors[i];
  __syncthreads();

  BinCache cache; cache.bin = -1; cache.cnt = 0; cache.correct = 0; cache.sconf = 0.0;
  double a_nll = 0.0, a_correct = 0.0, a_n = 0.0;
  const int64_t ntiles = (N + NT - 1) / NT;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t base = t * NT;
    if (STAGED) {
      const int64_t avail = (N - base) * (int64_t)K;
      const int total = (int)(avail < (int64_t)NT * K ? avail : (int64_t)NT * K);
      const T* gp = in + base * K;
      __syncthreads();
      for (int e = tid; e < total; e += NT) tile[skew(e)] = gp[e];
      __syncthreads();
    }
    const int64_t n = base + tid;
    if (n < N) {
      const T* grow = in + n * K;
      auto get = [&](int j) -> T { return STAGED ? tile[skew(tid * K + j)] : grow[j]; };
      int yy = (y != nullptr) ? (int)y[n] : -1;
      T conf, py;
      int pred = 0;
      if (mode == CNF_METRICS_PROBS) {
        conf = get(0);
        for (int j = 1; j < K; ++j) { T v = get(j); if (v > conf) { conf = v; pred = j; } }
        py = (yy >= 0 && yy < K) ? get(yy) : (T)0;
        a_nll -= (double)t_log<T>(py + (T)1e-7);
      } else {
        // float32 softmax of the logits (scipy.special.softmax on float32)
        float mx = (float)get(0);
        for (int j = 1; j < K; ++j) { float v = (float)get(j); if (v > mx) { mx = v; pred = j; } }
        float se = 0.f;
        for (int j = 0; j < K; ++j) se += expf((float)get(j) - mx);
        if (mode == CNF_METRICS_LOGITS) {
          // arg-max over the float32 probabilities: ties resolve to the first index
          float best = -1.f; pred = 0;
          for (int j = 0; j < K; ++j) { float p = expf((float)get(j) - mx) / se; if (p > best) { best = p; pred = j; } }
          conf = (T)best;
          float pyf = (yy >= 0 && yy < K) ? expf((float)get(yy) - mx) / se : 0.f;
          py = (T)pyf;
          a_nll -= (double)logf(pyf + 1e-7f);
        } else {
          // u_j = (double)log(p_j + 1e-7f) - log_prior_j ; float64 softmax of u
          double umax = -INFINITY;
          for (int j = 0; j < K; ++j) {
            float p = expf((float)get(j) - mx) / se;
            double u = (double)logf(p + 1e-7f) - s_lp[j];
            if (u > umax) umax = u;
          }
          double sd = 0.0;
          for (int j = 0; j < K; ++j) {
            float p = expf((float)get(j) - mx) / se;
            double u = (double)logf(p + 1e-7f) - s_lp[j];
            sd += exp(u - umax);
          }
          double best = -1.0, pyd = 0.0; pred = 0;
          for (int j = 0; j < K; ++j) {
            float p = expf((float)get(j) - mx) / se;
            double u = (double)logf(p + 1e-7f) - s_lp[j];
            double q = exp(u - umax) / sd;
            if (probs_out != nullptr) probs_out[n * K + j] = q;
            if (q > best) { best = q; pred = j; }
            if (j == yy) pyd = q;
          }
          conf = (T)best; py = (T)pyd;
          // calibrated probabilities are float64 in the reference: bin and score in double
          a_nll -= log(pyd + 1e-7);
          if (edges != nullptr) {
            // binning below is done in T; for float inputs re-do it in double here
            double c = best;
            int j = (int)ceil(c * bins) - 1;
            j = j < 0 ? 0 : (j > bins - 1 ? bins - 1 : j);
            while (j > 0 && !(edges[j] < c)) --j;
            while (j < bins - 1 && !(c <= edges[j + 1])) ++j;
            const bool inbin = (edges[j] < c) && (c <= edges[j + 1]);
            const unsigned ok = (pred == yy) ? 1u : 0u;
            if (inbin) {
              if (j != cache.bin) { flush(cache, s_cnt, s_cor, s_conf); cache.bin = j; }
              cache.cnt += 1; cache.correct += ok; cache.sconf += c;
            }
            a_correct += ok; a_n += 1.0;
            continue;
          }
        }
      }
      const unsigned ok = (pred == yy) ? 1u : 0u;
      a_correct += ok; a_n += 1.0;
      if (edges != nullptr) {
        const T c = conf;
        int j = (int)ceil((double)c * bins) - 1;
        j = j < 0 ? 0 : (j > bins - 1 ? bins - 1 : j);
        while (j > 0 && !(s_edges[j] < c)) --j;
        while (j < bins - 1 && !(c <= s_edges[j + 1])) ++j;
        const bool inbin = (s_edges[j] < c) && (c <= s_edges[j + 1]);
        if (inbin) {
          if (j != cache.bin) { flush(cache, s_cnt, s_cor, s_conf); cache.bin = j; }
          cache.cnt += 1; cache.correct += ok; cache.sconf += (double)c;
        }
      }
    }
  }
  flush(cache, s_cnt, s_cor, s_conf);
  // block reduction of the three scalars
  double v3[3] = {a_nll, a_correct, a_n};
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    double v = v3[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    if (tid == 0 && acc != nullptr) {
      double tsum = 0.0;
      for (int w = 0; w < (NT + 31) / 32; ++w) tsum += red[w];
      atomicAdd(acc + 3 * bins + q, tsum);
    }
  }
  __syncthreads();
  if (acc != nullptr)
    for (int i = tid; i < bins; i += NT) {
      if (s_cnt[i]) {
        atomicAdd(acc + i, (double)s_cnt[i]);
        atomicAdd(acc + bins + i, s_conf[i]);
        atomicAdd(acc + 2 * bins + i, (double)s_cor[i]);
      }
    }
}

template <typename T>
int launch_metrics(const T* in, const int64_t* y, int64_t N, int K, int bins, int mode, const double* lp,
                   const double* edges, double* acc, double* probs_out, cudaStream_t st) {
  int dev = 0, max_smem = 0, sms = 0;
  CNF_CHECK_CUDA(cudaGetDevice(&dev));
  CNF_CHECK_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  CNF_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t fixed = (size_t)bins * 8 + (size_t)K * 8 + (size_t)(bins + 2) * sizeof(T) + (size_t)bins * 8 + 32;
  int nt = 256;
  bool staged = false;
  for (; nt >= 64; nt >>= 1) {
    const size_t tile = ((size_t)nt * K + (size_t)nt * K / 32 + 8) * sizeof(T);
    if (fixed + tile <= 100 * 1024) { staged = true; break; }   // keep >= 2 CTAs per SM
  }
  if (!staged) nt = 128;
  const size_t tile = staged ? ((size_t)nt * K + (size_t)nt * K / 32 + 8) * sizeof(T) : 0;
  const size_t smem = fixed + tile;
  const int64_t ntiles = (N + nt - 1) / nt;
  const int64_t cap = (int64_t)sms * 4;
  const int grid = (int)(ntiles < cap ? ntiles : cap);
  if (staged) {
    CNF_CHECK_CUDA(cudaFuncSetAttribute(metrics_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    metrics_kernel<T, true><<<grid, nt, smem, st>>>(in, y, N, K, bins, mode, lp, edges, acc, probs_out);
  } else {
    CNF_CHECK_CUDA(cudaFuncSetAttribute(metrics_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    metrics_kernel<T, false><<<grid, nt, smem, st>>>(in, y, N, K, bins, mode, lp, edges, acc, probs_out);
  }
  CNF_CHECK_CUDA(cudaGetLastError());
  return CNF_OK;
}

}  // namespace

extern "C" int cnf_metrics(const void* in, int32_t is_f64, const int64_t* y, int64_t N, int32_t K, int32_t bins,
                           int32_t mode, const double* log_priors, const double* edges, double* acc, void* stream) {
  if (!in || !y || !acc || !edges || N < 0 || K < 1 || bins < 1 || bins > 4096) {
    cnf_set_error("cnf_metrics: bad argument"); return CNF_E_ARG;
  }
  if (mode < 0 || mode > 2) { cnf_set_error("cnf_metrics: bad mode %d", mode); return CNF_E_ARG; }
  if (mode == CNF_METRICS_CALIBRATED && !log_priors) { cnf_set_error("cnf_metrics: calibrated mode needs log_priors"); return CNF_E_ARG; }
  if (mode != CNF_METRICS_PROBS && is_f64) { cnf_set_error("cnf_metrics: logits must be float32"); return CNF_E_ARG; }
  if (N == 0) return CNF_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (is_f64) return launch_metrics<double>((const double*)in, y, N, K, bins, mode, log_priors, edges, acc, nullptr, st);
  return launch_metrics<float>((const float*)in, y, N, K, bins, mode, log_priors, edges, acc, nullptr, st);
}

extern "C" int cnf_calibrated_probs(const float* z, int64_t N, int32_t K, const double* log_priors,
                                    double* probs_out, void* stream) {
  if (!z || !log_priors || !probs_out || N < 0 || K < 1) { cnf_set_error("cnf_calibrated_probs: bad argument"); return CNF_E_ARG; }
  if (N == 0) return CNF_OK;
  cudaStream_t st = (cudaStream_t)stream;
  // no labels: metrics_kernel treats y == nullptr as "label -1" and acc == nullptr as "no statistics"
  return launch_metrics<float>(z, nullptr, N, K, 1, CNF_METRICS_CALIBRATED,
                               log_priors, nullptr, nullptr, probs_out, st);
}
