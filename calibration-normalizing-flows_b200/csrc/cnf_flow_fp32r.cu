// fp32 register-resident forward / inverse kernel for K = 10 logits (BASELINE configs C2 / C3 / C5: CIFAR-10
// shaped), one hidden layer, both conditioner nets, no random_flip.  Same arithmetic as flow_apply_kernel
// (cnf_flow_fp32.cu) -- the 1e-5 parity path -- laid out for the FMA pipe:
//
//   * a thread owns SPT samples and keeps their 10 logits in REGISTERS for the whole stack (two arrays of five:
//     physical slots 0..4 and 5..9; the per-layer flip of flows/flows.py:112 only swaps which array conditions
//     and which is transformed, so no data moves and no index is dynamic);
//   * all layers' weights stay in shared memory, re-arranged per hidden unit as three float4
//     [W1(5) | b1 | W2(5) | 0]: one hidden unit costs 3 broadcast LDS.128 against 10*SPT FMAs (the generic kernel
//     pays a weight load per 4 FMAs of ONE sample and keeps the tile in shared memory);
//   * the hidden vector is never materialised (stream over hidden units: a = b1 + W1.u, r = relu(a), out += W2*r);
//   * rows are read and written straight from global memory as float2 (40-byte rows; a warp covers 1280
//     contiguous bytes): the pass moves 84 B per 15 360 FMAs, I/O staging would buy nothing.
//
// Reference arithmetic: coupling forward / inverse flows/flows.py:101-126, conditioner MLP flows/utils.py:26-31,
// fused tail calibrators.py:17, 40-44, 350-353 + utils/metrics.py (see cnf_metrics_dev.cuh).
#include <cuda_runtime.h>

#include "cnf_common.h"
#include "cnf_metrics_dev.cuh"
#include "cnf_fp32r_dev.cuh"   // unit_fma / net_eval / layer_eval / weight staging / the training kernel (shared with cnf_flow_fp32rk.cu)

namespace {

template <int R_NT, int SPT, int U, int MINB, int TAIL, bool M2 = false>
__global__ void __launch_bounds__(R_NT, MINB)
flow_reg10_kernel(CnfDims d, const float* __restrict__ packed, const int* __restrict__ tables,
                  const float* __restrict__ xin, float* __restrict__ zout, float* __restrict__ logdet, int64_t N,
                  int inverse, CnfTail ta) {
  extern __shared__ __align__(16) float smem[];
  __shared__ double tail_red[32];
  const int tid = threadIdx.x;
  const int Hp = M2 ? d.H[1] : d.H[0], L = d.L;             // units streamed per net (true width: padding units are zeros)
  const int lay4 = 2 * (3 * Hp + (M2 ? 2 * RD : 0));        // float4 per layer
  // shared memory: [tail state][per layer: 2 nets x ([5 x 2 float4]) Hp x 3 float4][per layer: 2 x 8 floats of b2][per layer: 10 ints]
  const int tail_floats = TAIL ? cnf_tail_smem_bytes(ta.bins, RK) / 4 : 0;
  float4* ws = reinterpret_cast<float4*>(smem + tail_floats);
  float* b2s = reinterpret_cast<float*>(ws + (size_t)L * lay4);
  int* maps = reinterpret_cast<int*>(b2s + L * 16);       // per layer: packed input index of conditioning slot e, packed output index of transformed slot e
  TailSmem tsm;
  BinCache cache; cache.bin = -1; cache.cnt = 0; cache.correct = 0; cache.sconf = 0.0;
  double a_nll = 0.0;
  unsigned a_correct = 0u, a_n = 0u;
  if (TAIL) {
    tsm = tail_carve(reinterpret_cast<unsigned char*>(smem), ta.bins, RK);
    tail_init(tsm, ta, RK, tid, R_NT);
  }
  stage_weights_reg10<R_NT, M2>(d, packed, tables, reinterpret_cast<float*>(ws), b2s, maps, tid, Hp);

  const int TS = R_NT * SPT;
  const int64_t ntiles = (N + TS - 1) / TS;
  const bool rev_io = (L & 1) != 0;          // pi_L is the reversal for odd L, the identity for even L
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t base = tile * TS;
    float lo[SPT][RD], hi[SPT][RD], ld[SPT];
    // ---- rows -> registers (logical column j sits in physical slot pi(j); forward: pi_0 = id, inverse: pi_L) ----
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
      const int64_t n = base + tid + k * R_NT;
      float v[RK];
      if (n < N) {
        const float2* p = reinterpret_cast<const float2*>(xin + n * RK);
#pragma unroll
        for (int j = 0; j < RD; ++j) { const float2 t2 = __ldg(p + j); v[2 * j] = t2.x; v[2 * j + 1] = t2.y; }
      } else {
#pragma unroll
        for (int j = 0; j < RK; ++j) v[j] = 0.f;
      }
      if (TAIL && ta.center) {               // calibrators.py:17 / :42 in numpy's float32 summation order (K = 10)
        float r8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r8[j] = v[j];
        float res = __fadd_rn(__fadd_rn(__fadd_rn(r8[0], r8[1]), __fadd_rn(r8[2], r8[3])),
                              __fadd_rn(__fadd_rn(r8[4], r8[5]), __fadd_rn(r8[6], r8[7])));
        res = __fadd_rn(res, v[8]);
        res = __fadd_rn(res, v[9]);
        const float mean = __fdiv_rn(res, 10.f);
#pragma unroll
        for (int j = 0; j < RK; ++j) v[j] = __fsub_rn(v[j], mean);
      }
      const bool rev = inverse && rev_io;
#pragma unroll
      for (int j = 0; j < RD; ++j) {
        lo[k][j] = rev ? v[RK - 1 - j] : v[j];
        hi[k][j] = rev ? v[RD - 1 - j] : v[RD + j];
      }
      ld[k] = 0.f;
    }
    // ---- the stack -------------------------------------------------------------------------------------------
    for (int li = 0; li < L; ++li) {
      const int l = inverse ? L - 1 - li : li;
      const float4* w = ws + (size_t)l * lay4;
      const float* b2 = b2s + l * 16;
      if (l & 1) layer_eval<SPT, U, M2>(w, b2, Hp, inverse, lo, hi, ld);
      else       layer_eval<SPT, U, M2>(w, b2, Hp, inverse, hi, lo, ld);
    }
    // ---- registers -> rows (+ fused tail) -----------------------------------------------------------------------
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
      const int64_t n = base + tid + k * R_NT;
      if (n >= N) continue;
      const bool rev = !inverse && rev_io;
      float v[RK];
#pragma unroll
      for (int j = 0; j < RD; ++j) {
        v[j] = rev ? hi[k][RD - 1 - j] : lo[k][j];
        v[RD + j] = rev ? lo[k][RD - 1 - j] : hi[k][j];
      }
      if (!TAIL || logdet != nullptr) logdet[n] = ld[k];
      if (!TAIL || zout != nullptr) {
        float2* p = reinterpret_cast<float2*>(zout + n * RK);
#pragma unroll
        for (int j = 0; j < RD; ++j) p[j] = make_float2(v[2 * j], v[2 * j + 1]);
      }
      if (TAIL) {
        const int yy = ta.y != nullptr ? (int)ta.y[n] : -1;
        int r_bin = -1;
        unsigned r_ok = 0u;
        double r_conf = 0.0;
        auto col = [&](int jj) -> int { return jj; };
        auto get = [&](int jj) -> float { return v[jj]; };     // KT = 10: every loop is unrolled, indices are static
        row_stats<float, TAIL == 0 ? CNF_METRICS_LOGITS : TAIL, false, RK>(
            get, col, yy, n, RK, ta.bins, tsm.s_lp, tsm.s_edges, ta.acc != nullptr ? ta.edges : nullptr, ta.probs_out,
            r_bin, r_ok, r_conf, a_nll, a_correct, a_n);
        if (ta.acc != nullptr) warp_accumulate(r_bin, r_ok, r_conf, ta.bins, cache, tsm.s_cnt, tsm.s_cor, tsm.s_conf, 0);
      }
    }
  }
  if (TAIL) stats_finish_block(a_nll, a_correct, a_n, cache, tsm.s_cnt, tsm.s_cor, tsm.s_conf, ta.bins, ta.acc, tail_red, tid, R_NT);
}

template <int R_NT, int SPT, int U, int MINB, int TAIL, bool M2 = false>
int launch_reg10(const CnfDims& d, const float* packed, const int32_t* tables, const float* x, float* z, float* logdet,
                 int64_t N, int inverse, const CnfTail& ta, size_t smem, int sms, cudaStream_t st) {
  { const int rc = cnf_kernel_smem(flow_reg10_kernel<R_NT, SPT, U, MINB, TAIL, M2>, smem); if (rc) return rc; }
  int per_sm = 0;
  CNF_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, flow_reg10_kernel<R_NT, SPT, U, MINB, TAIL, M2>, R_NT, smem));
  if (per_sm < 1) per_sm = 1;
  const int64_t ntiles = (N + R_NT * SPT - 1) / (R_NT * SPT);
  const int64_t cap = (int64_t)sms * per_sm;
  const int grid = (int)(ntiles < cap ? ntiles : cap);
  flow_reg10_kernel<R_NT, SPT, U, MINB, TAIL, M2><<<grid, R_NT, smem, st>>>(d, packed, tables, x, z, logdet, N, inverse, ta);
  CNF_CHECK_CUDA(cudaGetLastError());
  return CNF_OK;
}

}  // namespace

// Shapes this kernel serves: K = 10, one hidden layer, both nets, standard flips (no random_flip), 8-byte aligned rows,
// weights of all layers within the shared-memory budget; forward / inverse / fused predict (two_ok) also two hidden
// layers with at most five units in the first -- the reference's default conditioner, hidden_size=[5, 5].
bool cnf_fp32r_supported(const cnf_flow_desc* desc, const CnfDims& d, const float* x, const float* z, int tail_bins,
                         int max_smem, size_t* smem_out, bool two_ok) {
  if (d.K != RK || d.nets != 3 || desc->perm != nullptr) return false;
  if (!(d.m == 1 || (two_ok && d.m == 2 && d.H[0] <= RD))) return false;
  if (((uintptr_t)x | (uintptr_t)z) % 8 != 0) return false;
  const int unit_floats = d.m == 2 ? 2 * (d.H[1] * 12 + 8 * RD) : 2 * d.Hp[0] * 12;
  const size_t smem = (size_t)d.L * (unit_floats + 16) * sizeof(float) + (size_t)((d.L * 10 + 3) / 4) * 4 * sizeof(int) +
                      (tail_bins > 0 ? cnf_tail_smem_bytes(tail_bins, RK) : 0);
  if ((long long)smem > max_smem - 1024) return false;
  *smem_out = smem;
  return true;
}

// Shapes the register-resident TRAINING kernel serves: K <= 10 (K = 10 here, K = 2 .. 9 in cnf_flow_fp32rk.cu), one
// hidden layer or two with at most five units in the first, any of the scale / shift nets, standard flips.
bool cnf_fp32r_train_supported(const cnf_flow_desc* desc, const CnfDims& d, const float* x, int max_smem, size_t* smem_out) {
  if (d.K < 2 || d.K > RK || d.n_nets < 1 || desc->perm != nullptr) return false;
  if (!(d.m == 1 || (d.m == 2 && d.H[0] <= RD))) return false;
  if (d.K == RK && (uintptr_t)x % 8 != 0) return false;
  const int unit_floats = d.m == 2 ? 2 * (d.H[1] * 12 + 8 * RD) : 2 * d.Hp[0] * 12;
  const size_t smem = (size_t)d.L * (unit_floats + 16) * sizeof(float) + (size_t)((d.L * 10 + 3) / 4) * 4 * sizeof(int);
  if ((long long)smem > max_smem - 1024) return false;
  *smem_out = smem;
  return true;
}

int cnf_fp32rk_train(const CnfDims& d, const float* packed, const int32_t* tables, const float* x, const int64_t* y,
                     float* partials, double* loss_acc, int64_t N, float eps, float gamma, float inv_n, size_t smem_fwd,
                     int sms, int max_smem, int64_t* rows_out, const float* gz_ext, const float* gld_ext, cudaStream_t st);

int cnf_fp32r_apply(const CnfDims& d, const float* packed, const int32_t* tables, const float* x, float* z, float* logdet,
                    int64_t N, int inverse, const CnfTail* tail, size_t smem, int sms, int variant, cudaStream_t st) {
  const CnfTail ta = tail ? *tail : CnfTail();
  if (d.m == 2) {          // first hidden layer materialised (the reference's default [5, 5] conditioner)
    // These nets are tiny (150 FMAs per sample and layer at [5, 5]): resident warps matter more than weight-load reuse.
    // Measured at K = 10, L = 6, [5, 5], 10^7 samples (default_shape_speed.py): 8 samples per thread 11.4 G samples/s,
    // 4 per thread 15.5 G, 2 per thread 15.7 G (the generic tile kernel: 2.3 G).
    // Small batches (these shapes leave the generic tile kernels from 1,024 samples on): 128-sample tiles (one sample
    // per thread) while they do not fill the SMs, then 256- and 512-sample tiles.
    const int spt = variant == 5 ? 1 : (variant == 4 ? 2 : 4);
#define R2(SPT, MB, TL) return launch_reg10<128, SPT, 2, MB, TL, true>(d, packed, tables, x, z, logdet, N, inverse, ta, smem, sms, st)
    if (!tail) { if (spt == 1) R2(1, 8, 0); if (spt == 2) R2(2, 6, 0); R2(4, 4, 0); }
    if (ta.mode == CNF_METRICS_LOGITS) { if (spt == 1) R2(1, 8, CNF_METRICS_LOGITS); if (spt == 2) R2(2, 6, CNF_METRICS_LOGITS); R2(4, 4, CNF_METRICS_LOGITS); }
    if (spt == 1) R2(1, 8, CNF_METRICS_CALIBRATED);
    if (spt == 2) R2(2, 6, CNF_METRICS_CALIBRATED);
    R2(4, 4, CNF_METRICS_CALIBRATED);
#undef R2
  }
  if (!tail) {
#define RV(NT, SPT, U, MB) return launch_reg10<NT, SPT, U, MB, 0>(d, packed, tables, x, z, logdet, N, inverse, ta, smem, sms, st)
    // Measured on B200 at the C2 shape, 10^7 samples (profiles/microbench/fp32r_speed.py): 128 threads x 8 samples
    // 1.59 G samples/s; x 6 (3 CTAs/SM) 1.52; x 4 1.44; 256 x 8 (1 CTA/SM) 1.57; 128 x 12 (1 CTA/SM, 4 warps) 1.59;
    // 64 x 8 1.19; generic flow_apply_kernel 0.79.  Samples per thread (weight loads per FMA) matter, occupancy does not.
    switch (variant) {      // (threads per CTA, samples per thread, unroll, min CTAs per SM): experiment switch
      case 1: RV(128, 4, 2, 3);
      case 2: RV(128, 6, 2, 3);
      case 3: RV(256, 8, 2, 1);
      case 4: RV(128, 2, 2, 3);
      case 5: RV(128, 1, 2, 3);
      case 6: RV(64, 2, 2, 3);
      case 7: RV(64, 1, 2, 3);
      default: RV(128, 8, 2, 2);
    }
#undef RV
  }
#define RT(SPT, MB)                                                                                                       \
  do {                                                                                                                    \
    if (ta.mode == CNF_METRICS_LOGITS) return launch_reg10<128, SPT, 2, MB, CNF_METRICS_LOGITS>(d, packed, tables, x, z, logdet, N, inverse, ta, smem, sms, st); \
    return launch_reg10<128, SPT, 2, MB, CNF_METRICS_CALIBRATED>(d, packed, tables, x, z, logdet, N, inverse, ta, smem, sms, st); \
  } while (0)
  if (variant == 4) RT(2, 3);       // the same tile sizes as the plain forward: 256 / 512 / 1024 samples
  if (variant == 1) RT(4, 3);
  RT(8, 2);
#undef RT
}

// Fused NLL training step on the register-resident kernel; rows_used = partial rows written (one per warp).
int cnf_fp32r_train(const CnfDims& d, const float* packed, const int32_t* tables, const float* x, const int64_t* y,
                    float* partials, double* loss_acc, int64_t N, float eps, float gamma, float inv_n, size_t smem_fwd,
                    int sms, int max_smem, int variant, int64_t* rows_out, const float* gz_ext, const float* gld_ext,
                    cudaStream_t st) {
  if (d.K != RK) return cnf_fp32rk_train(d, packed, tables, x, y, partials, loss_acc, N, eps, gamma, inv_n, smem_fwd, sms, max_smem, rows_out, gz_ext, gld_ext, st);
#define TVX(NT, SPT, U, MB, M2)                                                                                     \
  do {                                                                                                               \
    const size_t smem = smem_fwd + (size_t)3 * SPT * RD * NT * sizeof(float);                                       \
    if ((long long)smem > max_smem - 1024) break;      /* this variant's plan does not fit: try the next smaller one */ \
    int rc = cnf_kernel_smem(train_reg10_kernel<NT, SPT, U, MB, M2>, smem);                                              \
    if (rc) return rc;                                                                                               \
    int per_sm = 0;                                                                                                  \
    CNF_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, train_reg10_kernel<NT, SPT, U, MB, M2>, NT, smem)); \
    if (per_sm < 1) per_sm = 1;                                                                                      \
    const int64_t ntiles = (N + NT * SPT - 1) / (NT * SPT);                                                          \
    int64_t cap = (int64_t)sms * per_sm;                                                                             \
    if (partials && cap > d.grad_rows_max / (NT / 32)) cap = d.grad_rows_max / (NT / 32);                            \
    if (cap < 1) { cnf_set_error("partial buffer too small for the register-resident training kernel"); return CNF_E_SMEM; } \
    const int grid = (int)(ntiles < cap ? ntiles : cap);                                                             \
    const int64_t rows = (int64_t)grid * (NT / 32);                                                                  \
    if (rows_out) *rows_out = rows;                                                                                  \
    if (partials) CNF_CHECK_CUDA(cudaMemsetAsync(partials, 0, (size_t)(rows_out ? rows : d.grad_rows_max) * d.n_packed * sizeof(float), st)); \
    train_reg10_kernel<NT, SPT, U, MB, M2><<<grid, NT, smem, st>>>(d, packed, tables, x, y, partials, loss_acc, N, eps, gamma, inv_n, gz_ext, gld_ext); \
    CNF_CHECK_CUDA(cudaGetLastError());                                                                              \
    return CNF_OK;                                                                                                   \
  } while (0)
#define TV(NT, SPT, U, MB) TVX(NT, SPT, U, MB, false)
  if (d.m == 2) {          // two hidden layers, the first of at most five units (the reference's default [5, 5])
    // these shapes come here from 1,024 samples on (their alternative is the generic one-thread-per-sample tile kernel):
    // the largest tile that still gives every SM one, 128 x 1 at calibration-set sizes
    if (variant == 5) TVX(128, 4, 2, 2, true);
    if (variant == 2) TVX(128, 8, 2, 1, true);
    if (variant == 0) {
      if (N < 128LL * 2 * sms) TVX(128, 1, 2, 4, true);
      if (N < 128LL * 4 * sms) TVX(128, 2, 2, 4, true);
      if (N < 256LL * 4 * sms) TVX(128, 4, 2, 2, true);
    }
    TVX(256, 4, 2, 1, true);
    TVX(128, 4, 2, 2, true);
    cnf_set_error("register-resident training kernel: the weights of %d layers do not fit shared memory", d.L);
    return CNF_E_SMEM;
  }
  // Measured on B200 at the C2 shape, 4 Mi samples (profiles/microbench/fp32r_train_speed.py): 256 threads x 8 samples
  // (one CTA per SM, the whole register file) 266.8 M samples/s; 256 x 6 241; 128 x 8 214; 128 x 4 (2 CTAs/SM) 206;
  // the 32-sample-tile split kernel 164.6.  Samples per thread amortise the weight loads and the butterfly.
  switch (variant) {      // experiment switch (CNF_FP32R_TRAIN); a plan that does not fit shared memory falls through
    case 1: TV(128, 6, 2, 1);
    case 2: TV(128, 8, 2, 1);
    case 4: TV(256, 6, 2, 1);
    case 5: TV(128, 4, 2, 2);
    default: break;        // (tried: 128 x 2, 256 x 4, 256 x 2 for 65k..160k samples -- never ahead of the tile kernel there)
  }
  {
    // 2048-sample tiles (256 x 8) unless 1536-sample tiles (256 x 6) need fewer or cheaper waves over the SMs
    // (one wave of either takes ~1.23 / ~1.04 ms at the C2 shape: 400,000 samples are 2 waves of either)
    const long long w8 = (N + 2048LL * sms - 1) / (2048LL * sms), w6 = (N + 1536LL * sms - 1) / (1536LL * sms);
    if (variant == 0 && w6 * 104 < w8 * 123) TV(256, 6, 2, 1);
  }
  TV(256, 8, 2, 1);
  TV(128, 8, 2, 1);
  TV(128, 4, 2, 2);
  cnf_set_error("register-resident training kernel: the weights of %d layers do not fit shared memory", d.L);
  return CNF_E_SMEM;
#undef TV
#undef TVX
}
