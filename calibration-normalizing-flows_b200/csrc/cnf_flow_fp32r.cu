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

namespace {

constexpr int RK = 10, RD = 5;          // classes, half split

// one conditioner net for this thread's SPT samples: out[k][q] = b2[q] + sum_h W2[q][h] * relu(b1[h] + sum_j W1[h][j] c[k][j])
template <int SPT, int U>
__device__ __forceinline__ void net_eval(const float4* __restrict__ w, const float* __restrict__ b2, int Hp,
                                         const float (&c)[SPT][RD], float (&out)[SPT][RD]) {
#pragma unroll
  for (int k = 0; k < SPT; ++k)
#pragma unroll
    for (int q = 0; q < RD; ++q) out[k][q] = b2[q];
#pragma unroll U
  for (int h = 0; h < Hp; ++h) {
    const float4 v0 = w[3 * h], v1 = w[3 * h + 1], v2 = w[3 * h + 2];
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
      float a = fmaf(v0.x, c[k][0], v1.y);
      a = fmaf(v0.y, c[k][1], a);
      a = fmaf(v0.z, c[k][2], a);
      a = fmaf(v0.w, c[k][3], a);
      a = fmaf(v1.x, c[k][4], a);
      const float r = fmaxf(a, 0.f);
      out[k][0] = fmaf(v1.z, r, out[k][0]);
      out[k][1] = fmaf(v1.w, r, out[k][1]);
      out[k][2] = fmaf(v2.x, r, out[k][2]);
      out[k][3] = fmaf(v2.y, r, out[k][3]);
      out[k][4] = fmaf(v2.z, r, out[k][4]);
    }
  }
}

// one coupling layer: c conditions, t is transformed in place (flows/flows.py:105-109 / :119-125).  One net's outputs
// are consumed before the other net runs (forward: scale first, t*e^s then + shift; inverse: shift first, (t - shift)
// then * e^-s), so only one [SPT][5] output block is live at a time.
template <int SPT, int U>
__device__ __forceinline__ void layer_eval(const float4* __restrict__ w, const float* __restrict__ b2, int Hp, int inverse,
                                           const float (&c)[SPT][RD], float (&t)[SPT][RD], float (&ld)[SPT]) {
  float o[SPT][RD];
  if (!inverse) {
    net_eval<SPT, U>(w, b2, Hp, c, o);                       // s
#pragma unroll
    for (int k = 0; k < SPT; ++k)
#pragma unroll
      for (int q = 0; q < RD; ++q) { t[k][q] *= expf(o[k][q]); ld[k] += o[k][q]; }
    net_eval<SPT, U>(w + 3 * Hp, b2 + 8, Hp, c, o);          // shift
#pragma unroll
    for (int k = 0; k < SPT; ++k)
#pragma unroll
      for (int q = 0; q < RD; ++q) t[k][q] += o[k][q];
  } else {
    net_eval<SPT, U>(w + 3 * Hp, b2 + 8, Hp, c, o);          // shift
#pragma unroll
    for (int k = 0; k < SPT; ++k)
#pragma unroll
      for (int q = 0; q < RD; ++q) t[k][q] -= o[k][q];
    net_eval<SPT, U>(w, b2, Hp, c, o);                       // s
#pragma unroll
    for (int k = 0; k < SPT; ++k)
#pragma unroll
      for (int q = 0; q < RD; ++q) { t[k][q] *= expf(-o[k][q]); ld[k] -= o[k][q]; }
  }
}

template <int R_NT, int SPT, int U, int MINB, int TAIL>
__global__ void __launch_bounds__(R_NT, MINB)
flow_reg10_kernel(CnfDims d, const float* __restrict__ packed, const int* __restrict__ tables,
                  const float* __restrict__ xin, float* __restrict__ zout, float* __restrict__ logdet, int64_t N,
                  int inverse, CnfTail ta) {
  extern __shared__ __align__(16) float smem[];
  __shared__ double tail_red[32];
  const int tid = threadIdx.x;
  const int Hp = d.Hp[0], L = d.L;
  // shared memory: [tail state][per layer: 2 nets x Hp x 3 float4][per layer: 2 x 8 floats of b2]
  const int tail_floats = TAIL ? cnf_tail_smem_bytes(ta.bins, RK) / 4 : 0;
  float4* ws = reinterpret_cast<float4*>(smem + tail_floats);
  float* b2s = reinterpret_cast<float*>(ws + (size_t)L * 2 * Hp * 3);
  TailSmem tsm;
  BinCache cache; cache.bin = -1; cache.cnt = 0; cache.correct = 0; cache.sconf = 0.0;
  double a_nll = 0.0;
  unsigned a_correct = 0u, a_n = 0u;
  if (TAIL) {
    tsm = tail_carve(reinterpret_cast<unsigned char*>(smem), ta.bins, RK);
    tail_init(tsm, ta, RK, tid, R_NT);
  }
  // ---- stage the weights: packed [d1][Hp] / [Hp] / [d0][Hp] / [d0p] per net -> per hidden unit 12 floats, inputs and
  //      outputs placed by their physical slot inside the conditioning / transformed half (tables) ------------------
  {
    float* wf = reinterpret_cast<float*>(ws);
    const int per_layer = 2 * Hp * 12;
    for (int i = tid; i < L * per_layer; i += R_NT) {
      const int l = i / per_layer, r = i - l * per_layer;
      const int net = r / (Hp * 12), rr = r - net * (Hp * 12);
      const int h = rr / 12, e = rr - h * 12;
      const float* Wn = packed + (size_t)l * d.layer_stride + (size_t)net * d.net_stride;
      const int* cond = tables + d.tab_cond + l * RD;
      const int* trans = tables + d.tab_trans + l * RD;
      const int cbase = (l & 1) ? 0 : RD, tbase = (l & 1) ? RD : 0;   // even layers condition on slots 5..9
      float v = 0.f;
      if (e < 5) {                      // W1 column of the conditioning input that lives in slot cbase + e
        for (int j = 0; j < RD; ++j) if (cond[j] - cbase == e) v = Wn[d.w_off[0] + j * Hp + h];
      } else if (e == 5) {
        v = Wn[d.b_off[0] + h];
      } else if (e < 11) {              // W2 row of the output that lands in slot tbase + (e - 6)
        for (int q = 0; q < RD; ++q) if (trans[q] - tbase == e - 6) v = Wn[d.w_off[1] + q * Hp + h];
      }
      wf[i] = v;
    }
    for (int i = tid; i < L * 16; i += R_NT) {
      const int l = i >> 4, net = (i >> 3) & 1, e = i & 7;
      const float* Wn = packed + (size_t)l * d.layer_stride + (size_t)net * d.net_stride;
      const int* trans = tables + d.tab_trans + l * RD;
      const int tbase = (l & 1) ? RD : 0;
      float v = 0.f;
      for (int q = 0; q < RD; ++q) if (trans[q] - tbase == e) v = Wn[d.b_off[1] + q];
      b2s[i] = v;
    }
  }
  __syncthreads();

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
      const float4* w = ws + (size_t)l * 2 * Hp * 3;
      const float* b2 = b2s + l * 16;
      if (l & 1) layer_eval<SPT, U>(w, b2, Hp, inverse, lo, hi, ld);
      else       layer_eval<SPT, U>(w, b2, Hp, inverse, hi, lo, ld);
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

template <int R_NT, int SPT, int U, int MINB, int TAIL>
int launch_reg10(const CnfDims& d, const float* packed, const int32_t* tables, const float* x, float* z, float* logdet,
                 int64_t N, int inverse, const CnfTail& ta, size_t smem, int sms, cudaStream_t st) {
  { const int rc = cnf_kernel_smem(flow_reg10_kernel<R_NT, SPT, U, MINB, TAIL>, smem); if (rc) return rc; }
  int per_sm = 0;
  CNF_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, flow_reg10_kernel<R_NT, SPT, U, MINB, TAIL>, R_NT, smem));
  if (per_sm < 1) per_sm = 1;
  const int64_t ntiles = (N + R_NT * SPT - 1) / (R_NT * SPT);
  const int64_t cap = (int64_t)sms * per_sm;
  const int grid = (int)(ntiles < cap ? ntiles : cap);
  flow_reg10_kernel<R_NT, SPT, U, MINB, TAIL><<<grid, R_NT, smem, st>>>(d, packed, tables, x, z, logdet, N, inverse, ta);
  CNF_CHECK_CUDA(cudaGetLastError());
  return CNF_OK;
}

}  // namespace

// Shapes this kernel serves: K = 10, one hidden layer, both nets, standard flips (no random_flip), 8-byte aligned rows,
// weights of all layers within the shared-memory budget.
bool cnf_fp32r_supported(const cnf_flow_desc* desc, const CnfDims& d, const float* x, const float* z, int tail_bins,
                         int max_smem, size_t* smem_out) {
  if (d.K != RK || d.m != 1 || d.nets != 3 || desc->perm != nullptr) return false;
  if (((uintptr_t)x | (uintptr_t)z) % 8 != 0) return false;
  const size_t smem = (size_t)d.L * (2 * d.Hp[0] * 12 + 16) * sizeof(float) + (tail_bins > 0 ? cnf_tail_smem_bytes(tail_bins, RK) : 0);
  if ((long long)smem > max_smem - 1024) return false;
  *smem_out = smem;
  return true;
}

int cnf_fp32r_apply(const CnfDims& d, const float* packed, const int32_t* tables, const float* x, float* z, float* logdet,
                    int64_t N, int inverse, const CnfTail* tail, size_t smem, int sms, int variant, cudaStream_t st) {
  const CnfTail ta = tail ? *tail : CnfTail();
  if (!tail) {
#define RV(NT, SPT, U, MB) return launch_reg10<NT, SPT, U, MB, 0>(d, packed, tables, x, z, logdet, N, inverse, ta, smem, sms, st)
    // Measured on B200 at the C2 shape, 10^7 samples (profiles/microbench/fp32r_speed.py): 128 threads x 8 samples
    // 1.59 G samples/s; x 6 (3 CTAs/SM) 1.52; x 4 1.44; 256 x 8 (1 CTA/SM) 1.57; 128 x 12 (1 CTA/SM, 4 warps) 1.59;
    // 64 x 8 1.19; generic flow_apply_kernel 0.79.  Samples per thread (weight loads per FMA) matter, occupancy does not.
    switch (variant) {      // (threads per CTA, samples per thread, unroll, min CTAs per SM): experiment switch
      case 1: RV(128, 4, 2, 3);
      case 2: RV(128, 6, 2, 3);
      case 3: RV(256, 8, 2, 1);
      default: RV(128, 8, 2, 2);
    }
#undef RV
  }
  if (ta.mode == CNF_METRICS_LOGITS) return launch_reg10<128, 8, 2, 2, CNF_METRICS_LOGITS>(d, packed, tables, x, z, logdet, N, inverse, ta, smem, sms, st);
  return launch_reg10<128, 8, 2, 2, CNF_METRICS_CALIBRATED>(d, packed, tables, x, z, logdet, N, inverse, ta, smem, sms, st);
}
