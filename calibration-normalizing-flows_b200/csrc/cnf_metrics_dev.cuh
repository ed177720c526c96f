// Device-side pieces of the ECE / NLL / accuracy statistics shared by the streaming metrics kernel
// (cnf_metrics.cu) and the fused tails of the flow kernels (cnf_flow_fp32.cu, cnf_flow_tc.cu): per-row
// statistics, the per-lane run cache and the fixed-point shared-memory histogram.  Not part of the ABI.
//
// Reference arithmetic restated here: utils/metrics.py:35-73 (ECE, right-closed bins), :6-15 (NLL),
// :76-80 (accuracy); calibrators.py:40-44, 350-353 (Calibrator.predict tail); calibrators.py:17, 42
// (row-mean centring, numpy float32 pairwise summation order).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "cnf_common.h"

namespace {

template <typename T> __device__ __forceinline__ T t_log(T v);
template <> __device__ __forceinline__ float t_log<float>(float v) { return logf(v); }
template <> __device__ __forceinline__ double t_log<double>(double v) { return log(v); }

struct BinCache {
  int bin; unsigned cnt, correct; double sconf;
};

// Confidence sums are kept in shared memory as 2^-40 fixed point, 64 bits split into two 32-bit
// words: only 32-bit integer atomics are native on shared memory (double and 64-bit atomicAdd
// compile to CAS loops that spin under the contention of a popular bin).  The low-word add returns
// the old value, so each add knows whether it wrapped and carries into the high word itself; the
// pair therefore always sums to the exact 64-bit total.  Order-independent; a CTA sums < 2^24.
#define CNF_FX_SCALE 1099511627776.0 /* 2^40 */
__device__ __forceinline__ void flush(BinCache& c, unsigned* s_cnt, unsigned* s_cor, double* s_conf) {
  if (c.bin >= 0 && c.cnt) {
    atomicAdd(s_cnt + c.bin, c.cnt);
    atomicAdd(s_cor + c.bin, c.correct);
    const unsigned long long fx = (unsigned long long)__double2ll_rn(c.sconf * CNF_FX_SCALE);
    unsigned* w = reinterpret_cast<unsigned*>(s_conf) + 2 * c.bin;
    const unsigned lo = (unsigned)fx, hi = (unsigned)(fx >> 32);
    const unsigned old = atomicAdd(w, lo);
    const unsigned carry = (old + lo < old) ? 1u : 0u;
    if (hi + carry) atomicAdd(w + 1, hi + carry);
  }
  c.cnt = 0; c.correct = 0; c.sconf = 0.0;
}

// Statistics of one row.  get(jj) returns the element stored at position jj of the row and
// col(jj) its class index (identity unless the row is read in a lane-rotated order to dodge
// shared-memory bank conflicts); ties resolve to the smallest class index, as np.argmax does.
template <typename T, int mode, bool ROT, int KT, typename Get, typename Col>
__device__ __forceinline__ void row_stats(Get get, Col col, int yy, int64_t n, int Krt, int bins,
                                          const double* s_lp, const T* s_edges, const double* __restrict__ edges,
                                          double* __restrict__ probs_out, int& out_bin, unsigned& out_ok,
                                          double& out_conf, double& a_nll, unsigned& a_correct, unsigned& a_n) {
  const int K = KT > 0 ? KT : Krt;      // compile-time row width for the specialised instantiations
  out_bin = -1;
  T conf, py = (T)0;
  int pred = 0;
  if (mode == CNF_METRICS_PROBS) {
    conf = get(0); pred = col(0);
    if (ROT) {
      if (pred == yy) py = conf;
#pragma unroll
      for (int jj = 1; jj < K; ++jj) {
        const T v = get(jj);
        const int c = col(jj);
        if (v > conf || (v == conf && c < pred)) { conf = v; pred = c; }
        if (c == yy) py = v;
      }
    } else {   // elements arrive in class order: strict '>' keeps the first maximum, as np.argmax
#pragma unroll
      for (int jj = 1; jj < K; ++jj) {
        const T v = get(jj);
        if (v > conf) { conf = v; pred = jj; }
      }
      py = (yy >= 0 && yy < K) ? get(yy) : (T)0;
    }
    a_nll -= (double)t_log<T>(py + (T)1e-7);
  } else {
    // float32 softmax of the logits (scipy.special.softmax on float32)
    float mx = (float)get(0);
#pragma unroll
    for (int jj = 1; jj < K; ++jj) mx = fmaxf(mx, (float)get(jj));
    float se = 0.f;
#pragma unroll
    for (int jj = 0; jj < K; ++jj) se += expf((float)get(jj) - mx);
    if (mode == CNF_METRICS_LOGITS) {
      float best = -1.f, pyf = 0.f;
#pragma unroll
      for (int jj = 0; jj < K; ++jj) {
        const float pj = expf((float)get(jj) - mx) / se;
        const int c = col(jj);
        if (pj > best || (pj == best && c < pred)) { best = pj; pred = c; }
        if (c == yy) pyf = pj;
      }
      conf = (T)best; py = (T)pyf;
      a_nll -= (double)logf(pyf + 1e-7f);
    } else {
      // Calibrator.predict tail (calibrators.py:44): q = softmax(log(p + 1e-7) - log_priors) in float64 on float32 p.
      // exp(log(p_j + 1e-7) - lp_j) = (p_j + 1e-7) / prior_j, so q_j = w_j / sum_k w_k with
      // w_j = (double)(p_j + 1e-7f) * (1 / prior_j): the same quantity without the log / exp round trip (the
      // reference's float32 log contributes <= 1e-6 relative rounding noise to q, which this form does not carry).
      // s_lp[] holds 1 / prior_j = exp(-log_prior_j), formed once per CTA.
      double wsum = 0.0, best = -1.0, wy = 0.0;
#pragma unroll
      for (int jj = 0; jj < K; ++jj) {
        const float pj = expf((float)get(jj) - mx) / se;
        const int c = col(jj);
        const double w = (double)(pj + 1e-7f) * s_lp[c];
        wsum += w;
        if (w > best || (w == best && c < pred)) { best = w; pred = c; }
        if (c == yy) wy = w;
      }
      if (probs_out != nullptr) {
#pragma unroll
        for (int jj = 0; jj < K; ++jj) {
          const float pj = expf((float)get(jj) - mx) / se;
          const int c = col(jj);
          probs_out[n * K + c] = (double)(pj + 1e-7f) * s_lp[c] / wsum;
        }
      }
      best /= wsum;
      const double pyd = wy / wsum;
      a_nll -= log(pyd + 1e-7);
      const unsigned ok = (pred == yy) ? 1u : 0u;
      a_correct += ok; a_n += 1u;
      if (edges != nullptr) {   // calibrated probabilities are float64 in the reference: bin in double
        const double c = best;
        int j = (int)ceil(c * bins) - 1;
        j = j < 0 ? 0 : (j > bins - 1 ? bins - 1 : j);
        while (j > 0 && !(edges[j] < c)) --j;
        while (j < bins - 1 && !(c <= edges[j + 1])) ++j;
        if ((edges[j] < c) && (c <= edges[j + 1])) { out_bin = j; out_ok = ok; out_conf = c; }
      }
      return;
    }
  }
  const unsigned ok = (pred == yy) ? 1u : 0u;
  a_correct += ok; a_n += 1u;
  if (edges != nullptr) {
    const T c = conf;
    int j = __float2int_ru((float)c * (float)bins) - 1;
    j = j < 0 ? 0 : (j > bins - 1 ? bins - 1 : j);
    T e_lo = s_edges[j], e_hi = s_edges[j + 1];
    if (!((e_lo < c) && (c <= e_hi))) {          // rare: the product rounded across an edge
      while (j > 0 && !(s_edges[j] < c)) --j;
      while (j < bins - 1 && !(c <= s_edges[j + 1])) ++j;
      e_lo = s_edges[j]; e_hi = s_edges[j + 1];
    }
    if ((e_lo < c) && (c <= e_hi)) { out_bin = j; out_ok = ok; out_conf = (double)c; }
  }
}

// Every lane keeps a run cache (bin, count, correct, sum conf) and touches the shared histogram
// only when its bin changes; with a confident classifier most consecutive samples share a bin.
// (A warp-aggregated variant -- ballot/popc/shuffle per distinct bin -- measured 2x slower.)
__device__ __forceinline__ void warp_accumulate(int bin, unsigned ok, double conf, int bins, BinCache& cache,
                                                unsigned* s_cnt, unsigned* s_cor, double* s_conf, int lane) {
  (void)bins; (void)lane;
  if (bin >= 0) {
    if (bin != cache.bin) { flush(cache, s_cnt, s_cor, s_conf); cache.bin = bin; }
    cache.cnt += 1; cache.correct += ok; cache.sconf += conf;
  }
}


// Block-level tail of a statistics pass: flush the run cache, fold the per-thread NLL / correct / row
// counters and add the CTA's shared histogram to the global accumulator acc[3*bins+3].  Every thread of
// the CTA must call it (threads without rows pass zeros and a cache with bin == -1).  red: 32 doubles
// of shared memory.
__device__ __forceinline__ void stats_finish_block(double a_nll, unsigned a_correct_u, unsigned a_n_u, BinCache& cache,
                                                   unsigned* s_cnt, unsigned* s_cor, double* s_conf, int bins,
                                                   double* __restrict__ acc, double* red, int tid, int NT) {
  flush(cache, s_cnt, s_cor, s_conf);
  // per-thread row counts stay below 2^32 (a thread sees at most N / grid-threads rows); exact in double
  double v3[3] = {a_nll, (double)a_correct_u, (double)a_n_u};
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
        atomicAdd(acc + bins + i, (double)reinterpret_cast<const unsigned long long*>(s_conf)[i] * (1.0 / CNF_FX_SCALE));
        atomicAdd(acc + 2 * bins + i, (double)s_cor[i]);
      }
    }
}

// Row mean in the summation order of numpy's float32 add.reduce (pairwise sum: a plain loop below 8
// elements, eight interleaved partial sums up to 128, calibrators.py:17 / :42 -> np.mean(axis=1)),
// divided in float32.  get(j) returns element j of the row.  K <= 128.
template <typename Get>
__device__ __forceinline__ float numpy_row_mean(Get get, int K) {
  float res;
  if (K < 8) {
    res = 0.f;
    for (int i = 0; i < K; ++i) res = __fadd_rn(res, get(i));
  } else {
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = get(j);
    int i = 8;
    for (; i < K - (K % 8); i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], get(i + j));
    }
    res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                    __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < K; ++i) res = __fadd_rn(res, get(i));
  }
  return __fdiv_rn(res, (float)K);
}

struct TailSmem {
  double* s_conf; double* s_lp; float* s_edges; unsigned* s_cnt; unsigned* s_cor;
};
__device__ __forceinline__ TailSmem tail_carve(unsigned char* base, int bins, int K) {
  TailSmem m;
  m.s_conf = reinterpret_cast<double*>(base);
  m.s_lp = m.s_conf + bins;
  m.s_edges = reinterpret_cast<float*>(m.s_lp + K);
  m.s_cnt = reinterpret_cast<unsigned*>(m.s_edges + (bins + 2));
  m.s_cor = m.s_cnt + bins;
  return m;
}
__device__ __forceinline__ void tail_init(const TailSmem& m, const CnfTail& ta, int K, int tid, int NT) {
  for (int i = tid; i < ta.bins; i += NT) { m.s_conf[i] = 0.0; m.s_cnt[i] = 0u; m.s_cor[i] = 0u; }
  if (ta.edges != nullptr)
    for (int i = tid; i <= ta.bins; i += NT) m.s_edges[i] = (float)ta.edges[i];
  if (ta.log_priors != nullptr)
    for (int i = tid; i < K; i += NT) m.s_lp[i] = exp(-ta.log_priors[i]);     // 1 / prior (see row_stats)
}

// Tail of one finished sample: get(j) = calibrated logit j (logical class order).
template <int MODE, typename Get>
__device__ __forceinline__ void tail_row(Get get, int64_t n, int K, const CnfTail& ta, const TailSmem& m, BinCache& cache,
                                         double& a_nll, unsigned& a_correct, unsigned& a_n) {
  const int yy = ta.y != nullptr ? (int)ta.y[n] : -1;
  int r_bin = -1;
  unsigned r_ok = 0u;
  double r_conf = 0.0;
  auto col = [&](int jj) -> int { return jj; };
  row_stats<float, MODE, false, 0>(get, col, yy, n, K, ta.bins, m.s_lp, m.s_edges, ta.acc != nullptr ? ta.edges : nullptr,
                                   ta.probs_out, r_bin, r_ok, r_conf, a_nll, a_correct, a_n);
  if (ta.acc != nullptr) warp_accumulate(r_bin, r_ok, r_conf, ta.bins, cache, m.s_cnt, m.s_cor, m.s_conf, 0);
}

}  // namespace
