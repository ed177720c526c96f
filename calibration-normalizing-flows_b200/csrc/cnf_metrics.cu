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

#include <cstdlib>

#include "cnf_common.h"
#include "cnf_tc_ptx.cuh"   // mbarrier + cp.async.bulk wrappers
#include "cnf_metrics_dev.cuh"

namespace {

__device__ __forceinline__ int skew(int a) { return a + (a >> 5); }

struct MetricsSmem {
  double* s_conf; double* s_lp; unsigned* s_cnt; unsigned* s_cor; unsigned char* ring;
};

template <typename T>
__device__ __forceinline__ T* carve(unsigned char* smem_raw, int bins, int K, MetricsSmem& m) {
  m.s_conf = reinterpret_cast<double*>(smem_raw);
  m.s_lp = m.s_conf + bins;
  T* s_edges = reinterpret_cast<T*>(m.s_lp + K);
  m.s_cnt = reinterpret_cast<unsigned*>(s_edges + (bins + 2));
  m.s_cor = m.s_cnt + bins;
  size_t off = reinterpret_cast<unsigned char*>(m.s_cor + bins) - smem_raw;
  off = (off + 15) / 16 * 16;
  m.ring = smem_raw + off;
  return s_edges;
}

__device__ __forceinline__ void finish_block(double a_nll, unsigned a_correct_u, unsigned a_n_u, BinCache& cache,
                                             const MetricsSmem& m, int bins, double* __restrict__ acc, double* red,
                                             int tid, int NT) {
  flush(cache, m.s_cnt, m.s_cor, m.s_conf);
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
      if (m.s_cnt[i]) {
        atomicAdd(acc + i, (double)m.s_cnt[i]);
        atomicAdd(acc + bins + i, (double)reinterpret_cast<const unsigned long long*>(m.s_conf)[i] * (1.0 / CNF_FX_SCALE));
        atomicAdd(acc + 2 * bins + i, (double)m.s_cor[i]);
      }
    }
}

// Streaming kernel: every warp owns a ring of `stages` shared-memory tiles of 32 rows, filled with
// 16-byte cp.async copies (fully coalesced, no block-level barrier); lane i then reads row i.
// PRIV: every lane owns a private histogram slice in shared memory (hist[bin][lane], bank == lane, plain
// read-modify-write, no atomics, no run cache): count and correct packed 16+16 bits, confidence sum in
// the accumulation type below.  Used when bins <= 32 and a lane sees < 65536 rows; otherwise the
// run-cache + shared-atomics path.  With real classifier outputs consecutive rows rarely share a bin,
// so the run cache flushed (4 shared atomics + a fixed-point conversion) on almost every row.
template <typename T, int MODE> struct ConfAcc { typedef float type; };
template <int MODE> struct ConfAcc<double, MODE> { typedef double type; };
template <> struct ConfAcc<float, CNF_METRICS_CALIBRATED> { typedef double type; };

template <typename T, int MODE, int KT, bool PRIV>
__global__ void __launch_bounds__(256, (MODE == CNF_METRICS_CALIBRATED || sizeof(T) == 8) ? 2 : 4)
metrics_stream_kernel(const T* __restrict__ in, const int64_t* __restrict__ y, int64_t N, int Krt,
                                      int bins, const double* __restrict__ log_priors,
                                      const double* __restrict__ edges, double* __restrict__ acc,
                                      double* __restrict__ probs_out, int stages, int rot) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double red[32];
  const int K = KT > 0 ? KT : Krt;
  const int NT = blockDim.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = NT >> 5;
  MetricsSmem m;
  T* s_edges = carve<T>(smem_raw, bins, K, m);
  for (int i = tid; i < bins; i += NT) { m.s_conf[i] = 0.0; m.s_cnt[i] = 0u; m.s_cor[i] = 0u; }   // 0.0 == 0ull
  if (edges != nullptr)
    for (int i = tid; i <= bins; i += NT) s_edges[i] = (T)edges[i];
  if (log_priors != nullptr)
    for (int i = tid; i < K; i += NT) m.s_lp[i] = exp(-log_priors[i]);     // 1 / prior (see row_stats)
  __syncthreads();

  const int tile_elems = 32 * K;
  const int tile_chunks = tile_elems * (int)sizeof(T) / 16;
  T* ring = reinterpret_cast<T*>(m.ring) + (size_t)warp * stages * tile_elems;
  typedef typename ConfAcc<T, MODE>::type CA;
  // private histograms sit behind the rings: [warp][bin][lane] confidence sums, then packed counts
  CA* h_cf = reinterpret_cast<CA*>(m.ring + (((size_t)nwarps * stages * tile_elems * sizeof(T) + 15) / 16) * 16);
  unsigned* h_pk = reinterpret_cast<unsigned*>(h_cf + (size_t)nwarps * bins * 32);
  if (PRIV) {
    for (int i = tid; i < nwarps * bins * 32; i += NT) { h_cf[i] = (CA)0; h_pk[i] = 0u; }
    __syncthreads();
    h_cf += (size_t)warp * bins * 32 + lane;
    h_pk += (size_t)warp * bins * 32 + lane;
  }
  unsigned n_rows = 0u, n_ok = 0u;
  const int64_t ntiles = (N + 31) / 32;
  const int64_t gw = (int64_t)blockIdx.x * nwarps + warp, GW = (int64_t)gridDim.x * nwarps;
  // One TMA bulk copy per tile (lane 0: expect_tx + cp.async.bulk, completion on the stage's mbarrier)
  // instead of 16-byte cp.async copies issued by every lane: ~15 fewer instructions per 32 rows.
  __shared__ uint64_t tile_bar[8][4];
  if (lane == 0)
    for (int s = 0; s < stages; ++s) mbar_init(&tile_bar[warp][s], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const unsigned tile_bytes = (unsigned)(tile_elems * sizeof(T));
  auto issue = [&](int64_t tile, int stage) {
    if (lane == 0 && tile < ntiles && (tile + 1) * 32 <= N) {
      mbar_expect_tx(&tile_bar[warp][stage], tile_bytes);
      bulk_copy_g2s(ring + (size_t)stage * tile_elems, in + tile * tile_elems, tile_bytes, &tile_bar[warp][stage]);
    }
  };
  for (int s = 0; s < stages - 1; ++s) issue(gw + (int64_t)s * GW, s);
  unsigned phase_bits = 0u;      // bit s: parity of the next completion of stage s

  BinCache cache; cache.bin = -1; cache.cnt = 0; cache.correct = 0; cache.sconf = 0.0;
  double a_nll = 0.0;
  unsigned a_correct = 0u, a_n = 0u;
  int stage = 0;
  for (int64_t tile = gw, i = 0; tile < ntiles; tile += GW, ++i) {
    int pre = stage + stages - 1;
    if (pre >= stages) pre -= stages;
    issue(tile + (int64_t)(stages - 1) * GW, pre);
    const int64_t n = tile * 32 + lane;
    const int yy = (y != nullptr && n < N) ? (int)y[n] : -1;
    if ((tile + 1) * 32 <= N) {           // a full tile was copied into this stage: wait for its bytes
      mbar_wait(&tile_bar[warp][stage], (phase_bits >> stage) & 1u);
      phase_bits ^= 1u << stage;
    }
    int r_bin = -1;
    unsigned r_ok = 0u;
    double r_conf = 0.0;
    if (n < N) {
      const bool full = (tile + 1) * 32 <= N;
      const T* srow = ring + (size_t)stage * tile_elems + lane * K;
      const T* grow = in + n * K;
      if (full) {
        if (rot) {
          auto col = [&](int jj) -> int { int c = jj + lane; return c >= K ? c - K * (c / K) : c; };
          auto get = [&](int jj) -> T { return srow[col(jj)]; };
          row_stats<T, MODE, true, KT>(get, col, yy, n, K, bins, m.s_lp, s_edges, edges, probs_out, r_bin, r_ok, r_conf, a_nll,
                       a_correct, a_n);
        } else {
          auto col = [&](int jj) -> int { return jj; };
          auto get = [&](int jj) -> T { return srow[jj]; };
          row_stats<T, MODE, false, KT>(get, col, yy, n, K, bins, m.s_lp, s_edges, edges, probs_out, r_bin, r_ok, r_conf, a_nll,
                       a_correct, a_n);
        }
      } else {
        auto col = [&](int jj) -> int { return jj; };
        auto get = [&](int jj) -> T { return grow[jj]; };
        row_stats<T, MODE, false, KT>(get, col, yy, n, K, bins, m.s_lp, s_edges, edges, probs_out, r_bin, r_ok, r_conf, a_nll,
                     a_correct, a_n);
      }
    }
    if (PRIV) {
      if (r_bin >= 0) {
        h_pk[r_bin * 32] += 1u + (r_ok << 16);
        h_cf[r_bin * 32] += (CA)r_conf;
      }
    } else {
      warp_accumulate(r_bin, r_ok, r_conf, bins, cache, m.s_cnt, m.s_cor, m.s_conf, lane);
    }
    __syncwarp();
    if (++stage == stages) stage = 0;
  }
  (void)n_rows; (void)n_ok;
  finish_block(a_nll, a_correct, a_n, cache, m, bins, acc, red, tid, NT);
  if (PRIV && acc != nullptr) {
    // lane-private slices -> global: one warp-reduction per (bin), warps and lanes summed in fixed order
    h_cf -= (size_t)warp * bins * 32 + lane;
    h_pk -= (size_t)warp * bins * 32 + lane;
    for (int b = warp; b < bins; b += nwarps) {
      double cf = 0.0;
      unsigned cn = 0u, co = 0u;
      for (int w = 0; w < nwarps; ++w) {
        const unsigned pk = h_pk[((size_t)w * bins + b) * 32 + lane];
        cn += pk & 0xffffu; co += pk >> 16;
        cf += (double)h_cf[((size_t)w * bins + b) * 32 + lane];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        cn += __shfl_xor_sync(0xffffffffu, cn, o);
        co += __shfl_xor_sync(0xffffffffu, co, o);
        cf += __shfl_xor_sync(0xffffffffu, cf, o);
      }
      if (lane == 0 && cn) {
        atomicAdd(acc + b, (double)cn);
        atomicAdd(acc + bins + b, cf);
        atomicAdd(acc + 2 * bins + b, (double)co);
      }
    }
  }
}

// Fallback for rows too wide to stage: one thread per row straight from global memory.
template <typename T, int MODE>
__global__ void metrics_direct_kernel(const T* __restrict__ in, const int64_t* __restrict__ y, int64_t N, int K,
                                      int bins, const double* __restrict__ log_priors,
                                      const double* __restrict__ edges, double* __restrict__ acc,
                                      double* __restrict__ probs_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double red[32];
  const int NT = blockDim.x, tid = threadIdx.x;
  MetricsSmem m;
  T* s_edges = carve<T>(smem_raw, bins, K, m);
  for (int i = tid; i < bins; i += NT) { m.s_conf[i] = 0.0; m.s_cnt[i] = 0u; m.s_cor[i] = 0u; }   // 0.0 == 0ull
  if (edges != nullptr)
    for (int i = tid; i <= bins; i += NT) s_edges[i] = (T)edges[i];
  if (log_priors != nullptr)
    for (int i = tid; i < K; i += NT) m.s_lp[i] = exp(-log_priors[i]);     // 1 / prior (see row_stats)
  __syncthreads();
  BinCache cache; cache.bin = -1; cache.cnt = 0; cache.correct = 0; cache.sconf = 0.0;
  double a_nll = 0.0;
  unsigned a_correct = 0u, a_n = 0u;
  for (int64_t n0 = (int64_t)blockIdx.x * NT; n0 < N; n0 += (int64_t)gridDim.x * NT) {
    const int64_t n = n0 + tid;
    int r_bin = -1;
    unsigned r_ok = 0u;
    double r_conf = 0.0;
    if (n < N) {
      const T* grow = in + n * K;
      const int yy = (y != nullptr) ? (int)y[n] : -1;
      auto col = [&](int jj) -> int { return jj; };
      auto get = [&](int jj) -> T { return grow[jj]; };
      row_stats<T, MODE, false, 0>(get, col, yy, n, K, bins, m.s_lp, s_edges, edges, probs_out, r_bin, r_ok, r_conf, a_nll,
                   a_correct, a_n);
    }
    warp_accumulate(r_bin, r_ok, r_conf, bins, cache, m.s_cnt, m.s_cor, m.s_conf, tid & 31);
  }
  finish_block(a_nll, a_correct, a_n, cache, m, bins, acc, red, tid, NT);
}

template <typename T>
int launch_metrics(const T* in, const int64_t* y, int64_t N, int K, int bins, int mode, const double* lp,
                   const double* edges, double* acc, double* probs_out, cudaStream_t st) {
  int dev = 0, sms = 0;
  CNF_CHECK_CUDA(cudaGetDevice(&dev));
  CNF_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t fixed = (size_t)bins * 8 + (size_t)K * 8 + (size_t)(bins + 2) * sizeof(T) + (size_t)bins * 8 + 32;
  const size_t tile = (size_t)32 * K * sizeof(T);
  // widest ring that keeps the CTA under ~56 KB (4 CTAs per SM); needs 16 B-aligned rows of tiles
  int nwarps = 0, stages = 0;
  const bool aligned = ((uintptr_t)in % 16 == 0);
  const bool wide_acc = sizeof(T) == 8 || mode == CNF_METRICS_CALIBRATED;
  // lane-private histograms pay off for the 8-byte inputs (4.20 -> 4.53 TB/s on B200); for float32 rows the
  // extra shared memory costs a ring stage and the shared-atomics path stays ahead (4.26 vs 4.01 TB/s)
  const size_t hist_per_warp = (bins <= 32 && wide_acc) ? (size_t)bins * 32 * 12 : 0;
  const int wopts[3] = {8, 4, 2}, sopts[3] = {4, 3, 2};
  const size_t budget = wide_acc ? 100 * 1024 : 56 * 1024;
  for (int wi = 0; wi < 3 && !nwarps && aligned; ++wi)
    for (int si = 0; si < 3; ++si)
      if (fixed + 16 + (tile * sopts[si] + hist_per_warp) * wopts[wi] <= budget) { nwarps = wopts[wi]; stages = sopts[si]; break; }
  if (const char* v = cnf_switch(CNF_SW_METRICS_STAGES)) { const int sv = atoi(v); if (sv >= 2 && sv <= 4 && nwarps) stages = sv; }
  if (nwarps) {
    const size_t smem = fixed + tile * nwarps * stages + (hist_per_warp ? 16 + hist_per_warp * nwarps : 0);
    const int64_t nt = (N + 31) / 32;
    const int per_sm = (int)(200 * 1024 / (smem + 1024)) > 8 ? 8 : (int)(200 * 1024 / (smem + 1024));
    int64_t grid = (nt + nwarps - 1) / nwarps;
    const int64_t cap = (int64_t)sms * (per_sm < 1 ? 1 : per_sm);
    if (grid > cap) grid = cap;
    const int words = K * (int)sizeof(T) / 4;
    const int rot = (words % 8 == 0) ? 1 : 0;     // row stride would hit >= 8-way bank conflicts
    const int64_t tiles_per_warp = (nt + grid * nwarps - 1) / (grid * nwarps);
    const bool priv = hist_per_warp > 0 && tiles_per_warp < 65536;     // 16-bit private counts
#define LAUNCH_STREAM_KP(M, KT, P)                                                                           \
  do {                                                                                                       \
    { const int rc2 = cnf_kernel_smem(metrics_stream_kernel<T, M, KT, P>, smem); if (rc2) return rc2; }     \
    metrics_stream_kernel<T, M, KT, P><<<(int)grid, nwarps * 32, smem, st>>>(in, y, N, K, bins, lp, edges, acc, \
                                                                               probs_out, stages, rot);      \
  } while (0)
    // K = 10 (CIFAR-10-shaped logits, configs C2/C3/C5) has its own instantiation with the row loops unrolled
#define LAUNCH_STREAM(M)                                                                                     \
  do {                                                                                                       \
    if (K == 10 && priv) LAUNCH_STREAM_KP(M, 10, true);                                                      \
    else if (K == 10) LAUNCH_STREAM_KP(M, 10, false);                                                        \
    else if (priv) LAUNCH_STREAM_KP(M, 0, true);                                                             \
    else LAUNCH_STREAM_KP(M, 0, false);                                                                      \
  } while (0)
    if (mode == CNF_METRICS_PROBS) LAUNCH_STREAM(CNF_METRICS_PROBS);
    else if (mode == CNF_METRICS_LOGITS) LAUNCH_STREAM(CNF_METRICS_LOGITS);
    else LAUNCH_STREAM(CNF_METRICS_CALIBRATED);
#undef LAUNCH_STREAM_KP
#undef LAUNCH_STREAM
  } else {
    const int nt = 128;
    const int64_t cap = (int64_t)sms * 8;
    const int64_t want = (N + nt - 1) / nt;
    const int grid = (int)(want < cap ? want : cap);
#define LAUNCH_DIRECT(M)                                                                                     \
  do {                                                                                                       \
    { const int rc2 = cnf_kernel_smem(metrics_direct_kernel<T, M>, fixed); if (rc2) return rc2; }           \
    metrics_direct_kernel<T, M><<<grid, nt, fixed, st>>>(in, y, N, K, bins, lp, edges, acc, probs_out);        \
  } while (0)
    if (mode == CNF_METRICS_PROBS) LAUNCH_DIRECT(CNF_METRICS_PROBS);
    else if (mode == CNF_METRICS_LOGITS) LAUNCH_DIRECT(CNF_METRICS_LOGITS);
    else LAUNCH_DIRECT(CNF_METRICS_CALIBRATED);
#undef LAUNCH_DIRECT
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
