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

// one hidden unit for one sample: out += W2[:, h] * relu(b1[h] + W1[h, :] . in)
__device__ __forceinline__ void unit_fma(const float4& v0, const float4& v1, const float4& v2, const float (&in)[RD],
                                         float (&out)[RD]) {
  float a = fmaf(v0.x, in[0], v1.y);
  a = fmaf(v0.y, in[1], a);
  a = fmaf(v0.z, in[2], a);
  a = fmaf(v0.w, in[3], a);
  a = fmaf(v1.x, in[4], a);
  const float r = fmaxf(a, 0.f);
  out[0] = fmaf(v1.z, r, out[0]);
  out[1] = fmaf(v1.w, r, out[1]);
  out[2] = fmaf(v2.x, r, out[2]);
  out[3] = fmaf(v2.y, r, out[3]);
  out[4] = fmaf(v2.z, r, out[4]);
}

// one conditioner net for this thread's SPT samples: out[k][q] = b2[q] + sum_h W2[q][h] * relu(b1[h] + sum_j W1[h][j] c[k][j])
// M2 (two hidden layers, the first of at most five units -- the reference's default conditioner hidden_size=[5, 5],
// flows/flows.py:69): the first hidden layer is materialised in registers (five units per sample, records
// [W1(5) | b1 | 0 0] in front of the net's unit records) and takes the place of c in the stream over the second one.
template <int SPT, int U, bool M2 = false>
__device__ __forceinline__ void net_eval(const float4* __restrict__ w, const float* __restrict__ b2, int Hn,
                                         const float (&c)[SPT][RD], float (&out)[SPT][RD]) {
#pragma unroll
  for (int k = 0; k < SPT; ++k)
#pragma unroll
    for (int q = 0; q < RD; ++q) out[k][q] = b2[q];
  float h1[M2 ? SPT : 1][RD];
  if (M2) {
#pragma unroll
    for (int i = 0; i < RD; ++i) {
      const float4 v0 = w[2 * i], v1 = w[2 * i + 1];
#pragma unroll
      for (int k = 0; k < SPT; ++k) {
        float a = fmaf(v0.x, c[k][0], v1.y);
        a = fmaf(v0.y, c[k][1], a);
        a = fmaf(v0.z, c[k][2], a);
        a = fmaf(v0.w, c[k][3], a);
        a = fmaf(v1.x, c[k][4], a);
        h1[M2 ? k : 0][i] = fmaxf(a, 0.f);
      }
    }
    w += 2 * RD;
  }
#pragma unroll U
  for (int h = 0; h < Hn; ++h) {
    const float4 v0 = w[3 * h], v1 = w[3 * h + 1], v2 = w[3 * h + 2];
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
      if constexpr (M2) unit_fma(v0, v1, v2, h1[k], out[k]);
      else              unit_fma(v0, v1, v2, c[k], out[k]);
    }
  }
}

// one coupling layer: c conditions, t is transformed in place (flows/flows.py:105-109 / :119-125).  One net's outputs
// are consumed before the other net runs (forward: scale first, t*e^s then + shift; inverse: shift first, (t - shift)
// then * e^-s), so only one [SPT][5] output block is live at a time.
template <int SPT, int U, bool M2 = false>
__device__ __forceinline__ void layer_eval(const float4* __restrict__ w, const float* __restrict__ b2, int Hp, int inverse,
                                           const float (&c)[SPT][RD], float (&t)[SPT][RD], float (&ld)[SPT]) {
  float o[SPT][RD];
  const int ns4 = 3 * Hp + (M2 ? 2 * RD : 0);               // float4 per net: [first-layer records] unit records
  if (!inverse) {
    net_eval<SPT, U, M2>(w, b2, Hp, c, o);                   // s
#pragma unroll
    for (int k = 0; k < SPT; ++k)
#pragma unroll
      for (int q = 0; q < RD; ++q) { t[k][q] *= expf(o[k][q]); ld[k] += o[k][q]; }
    net_eval<SPT, U, M2>(w + ns4, b2 + 8, Hp, c, o);         // shift
#pragma unroll
    for (int k = 0; k < SPT; ++k)
#pragma unroll
      for (int q = 0; q < RD; ++q) t[k][q] += o[k][q];
  } else {
    net_eval<SPT, U, M2>(w + ns4, b2 + 8, Hp, c, o);         // shift
#pragma unroll
    for (int k = 0; k < SPT; ++k)
#pragma unroll
      for (int q = 0; q < RD; ++q) t[k][q] -= o[k][q];
    net_eval<SPT, U, M2>(w, b2, Hp, c, o);                   // s
#pragma unroll
    for (int k = 0; k < SPT; ++k)
#pragma unroll
      for (int q = 0; q < RD; ++q) { t[k][q] *= expf(-o[k][q]); ld[k] -= o[k][q]; }
  }
}

// Stage the weights of all layers: packed [d1][Hp] / [Hp] / [d0][Hp] / [d0p] per net -> per hidden unit 12 floats
// [W1 of slot 0..4 | b1 | W2 of slot 0..4 | 0], inputs and outputs placed by their PHYSICAL slot inside the conditioning /
// transformed half (even layers condition on slots 5..9, odd ones on 0..4).  maps[l][0..4] = packed input index of
// conditioning slot e, maps[l][5..9] = packed output index of transformed slot e (from the index tables; formed first
// so that the copy below reads the packed blob with h fastest, i.e. coalesced, and never searches).
template <int R_NT, bool M2 = false>
__device__ __forceinline__ void stage_weights_reg10(const CnfDims& d, const float* __restrict__ packed,
                                                    const int* __restrict__ tables, float* wf, float* b2s, int* maps, int tid, int Hn) {
  // M2: per net 5 first-layer records of 8 floats [W1 of slot 0..4 | b1 | 0 0], then one 12-float record per unit of
  // the SECOND hidden layer (true width, not padded) [Wm from first-layer unit 0..4 | bm | W3 of slot 0..4 | 0]
  // Hn = hidden units streamed per net: the true width for the forward kernel, the padded one for the training kernel
  const int Hp = Hn, L = d.L;
  const int pre = M2 ? 8 * RD : 0;                         // floats in front of a net's unit records
  const int last = M2 ? 2 : 1;                             // index of the last Linear
  for (int i = tid; i < L * 10; i += R_NT) {
    const int l = i / 10, e = i - l * 10;
    const int* cond = tables + d.tab_cond + l * RD;
    const int* trans = tables + d.tab_trans + l * RD;
    const int cbase = (l & 1) ? 0 : RD, tbase = (l & 1) ? RD : 0;
    int v = 0;
    if (e < 5) { for (int j = 0; j < RD; ++j) if (cond[j] - cbase == e) v = j; }
    else       { for (int q = 0; q < RD; ++q) if (trans[q] - tbase == e - 5) v = q; }
    maps[i] = v;
  }
  __syncthreads();
  const int per_net = 12 * Hp;
  const int Hrow = M2 ? d.Hp[1] : d.Hp[0];                 // row length of the packed matrices the unit records read
  for (int i = tid; i < L * 2 * per_net; i += R_NT) {
    const int ln = i / per_net, r = i - ln * per_net;      // ln = l * 2 + net
    const int e = r / Hp, h = r - e * Hp;                  // h fastest: coalesced reads of the packed rows
    const int l = ln >> 1, net = ln & 1;
    const float* Wn = packed + (size_t)l * d.layer_stride + (size_t)net * d.net_stride;
    const int* mp = maps + l * 10;
    float v = 0.f;
    if (M2) {
      if (e < 5) v = e < d.H[0] ? __ldg(Wn + d.w_off[1] + e * Hrow + h) : 0.f;      // middle Linear [in][out]
      else if (e == 5) v = __ldg(Wn + d.b_off[1] + h);
      else if (e < 11) v = __ldg(Wn + d.w_off[2] + mp[e - 1] * Hrow + h);
    } else {
      if (e < 5) v = __ldg(Wn + d.w_off[0] + mp[e] * Hrow + h);
      else if (e == 5) v = __ldg(Wn + d.b_off[0] + h);
      else if (e < 11) v = __ldg(Wn + d.w_off[1] + mp[e - 1] * Hrow + h);
    }
    wf[(size_t)ln * (pre + per_net) + pre + (size_t)h * 12 + e] = v;
  }
  if constexpr (M2) {
    for (int i = tid; i < L * 2 * pre; i += R_NT) {
      const int ln = i / pre, r = i - ln * pre;
      const int e = r / RD, u = r - e * RD;                // u = first-layer unit
      const int l = ln >> 1, net = ln & 1;
      const float* Wn = packed + (size_t)l * d.layer_stride + (size_t)net * d.net_stride;
      const int* mp = maps + l * 10;
      float v = 0.f;
      if (u < d.H[0]) {
        if (e < 5) v = __ldg(Wn + d.w_off[0] + mp[e] * d.Hp[0] + u);
        else if (e == 5) v = __ldg(Wn + d.b_off[0] + u);
      }
      wf[(size_t)ln * (pre + per_net) + (size_t)u * 8 + e] = v;
    }
  }
  for (int i = tid; i < L * 16; i += R_NT) {
    const int l = i >> 4, net = (i >> 3) & 1, e = i & 7;
    const float* Wn = packed + (size_t)l * d.layer_stride + (size_t)net * d.net_stride;
    b2s[i] = e < RD ? __ldg(Wn + d.b_off[last] + maps[l * 10 + 5 + e]) : 0.f;
  }
  __syncthreads();
}

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

// ---------------------------------------------------------------------------------------------------------
// Training: the fused NLL step (forward, loss head, backward, weight gradients) of calibrators.py:287-293 with the
// same register-resident scheme, for batches that fill the GPU.  No tape and no barrier: after the forward pass
// and the loss head a thread walks the layers backwards; a layer's input is recovered from its output by the
// inverse (x = (y - t) e^-s with s, t recomputed from the unchanged conditioning half -- the recompute the
// backward pass needs anyway), then a second walk over the hidden units forms, per unit,
//   a = b1 + W1.c, r = relu(a), gh = [a > 0] W2^T g_out,  g_c += W1^T gh            (15 FMAs per sample)
//   dW2[:, h] += r g_out,  dW1[h, :] += gh c,  db1[h] += gh                          (11 FMAs per sample)
// summed first over the thread's SPT samples in registers, then over the warp's 32 lanes by a transposing
// butterfly (16 values -> 16 shuffles), and added by red.global into the WARP's own row of the partial buffer
// (rows are private to a warp and tiles follow each other on it: the sums are order-deterministic).
// ---------------------------------------------------------------------------------------------------------

// v[0..16) hold one value per index on every lane; on return v[0] of lane L is the warp-wide sum of value
// reduce16_index(L).  8 + 4 + 2 + 1 + 1 = 16 shuffles instead of 16 x 5.
__device__ __forceinline__ int reduce16_index(int lane) { return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1); }
__device__ __forceinline__ float warp_reduce16(float (&v)[16], int lane) {
#pragma unroll
  for (int w = 8, bit = 16; w >= 1; w >>= 1, bit >>= 1) {
    const bool up = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < w) {
        const float send = up ? v[i] : v[i + w];
        const float keep = up ? v[i + w] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
      }
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// Backward of one conditioner net for this thread's SPT samples (c: conditioning values, go: gradient on the net's
// five outputs).  Adds W1^T gh into gc and the weight gradients into the warp's partial row Gn (packed layout of one
// net; goff = this lane's entry offset for reduce16_index(lane), -1 for none, see train kernel).
template <int SPT, int U>
__device__ __forceinline__ void net_backward(const float4* __restrict__ w, int Hp, const float (&c)[SPT][RD],
                                             const float (&go)[SPT][RD], float (&gc)[SPT][RD], float* __restrict__ Gn,
                                             int goff, int lane) {
#pragma unroll U
  for (int h = 0; h < Hp; ++h) {
    const float4 v0 = w[3 * h], v1 = w[3 * h + 1], v2 = w[3 * h + 2];
    float wg[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) wg[i] = 0.f;
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
      float a = fmaf(v0.x, c[k][0], v1.y);
      a = fmaf(v0.y, c[k][1], a);
      a = fmaf(v0.z, c[k][2], a);
      a = fmaf(v0.w, c[k][3], a);
      a = fmaf(v1.x, c[k][4], a);
      const float r = fmaxf(a, 0.f);
      float gh = v1.z * go[k][0];
      gh = fmaf(v1.w, go[k][1], gh);
      gh = fmaf(v2.x, go[k][2], gh);
      gh = fmaf(v2.y, go[k][3], gh);
      gh = fmaf(v2.z, go[k][4], gh);
      gh = a > 0.f ? gh : 0.f;
      gc[k][0] = fmaf(v0.x, gh, gc[k][0]);
      gc[k][1] = fmaf(v0.y, gh, gc[k][1]);
      gc[k][2] = fmaf(v0.z, gh, gc[k][2]);
      gc[k][3] = fmaf(v0.w, gh, gc[k][3]);
      gc[k][4] = fmaf(v1.x, gh, gc[k][4]);
#pragma unroll
      for (int e = 0; e < RD; ++e) {
        wg[e] = fmaf(r, go[k][e], wg[e]);           // dW2[slot e][h]
        wg[RD + e] = fmaf(gh, c[k][e], wg[RD + e]); // dW1[h][slot e]
      }
      wg[10] += gh;                                 // db1[h]
    }
    const float tot = warp_reduce16(wg, lane);
    if (goff >= 0) atomicAdd(Gn + goff + h, tot);
  }
}

// The same for a net with two hidden layers, the first of at most five units (hidden_size=[5, 5]): the first layer is
// recomputed into registers, the stream over the second layer's units also accumulates the gradient on the first
// layer's activations, and a last pass over the five first-layer units forms W1^T ga, dW1 and db1.
// goff: entry offset of this lane for the second-layer / last-Linear values (0..4 dW3 of output slot e, 5..9 dWm from
// first-layer unit e, 10 dbm); goff1: for the first-layer values (0..4 dW1 of input slot e, 5 db1); -1 for none.
template <int SPT, int U>
__device__ __forceinline__ void net_backward_m2(const float4* __restrict__ w, int Hn, int H1, const float (&c)[SPT][RD],
                                                const float (&go)[SPT][RD], float (&gc)[SPT][RD], float* __restrict__ Gn,
                                                int goff, int goff1, int lane) {
  float h1[SPT][RD], gh1[SPT][RD];
#pragma unroll
  for (int i = 0; i < RD; ++i) {
    const float4 v0 = w[2 * i], v1 = w[2 * i + 1];
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
      float a = fmaf(v0.x, c[k][0], v1.y);
      a = fmaf(v0.y, c[k][1], a);
      a = fmaf(v0.z, c[k][2], a);
      a = fmaf(v0.w, c[k][3], a);
      a = fmaf(v1.x, c[k][4], a);
      h1[k][i] = fmaxf(a, 0.f);
      gh1[k][i] = 0.f;
    }
  }
  const float4* wu = w + 2 * RD;
#pragma unroll U
  for (int h = 0; h < Hn; ++h) {
    const float4 v0 = wu[3 * h], v1 = wu[3 * h + 1], v2 = wu[3 * h + 2];
    float wg[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) wg[i] = 0.f;
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
      float a = fmaf(v0.x, h1[k][0], v1.y);
      a = fmaf(v0.y, h1[k][1], a);
      a = fmaf(v0.z, h1[k][2], a);
      a = fmaf(v0.w, h1[k][3], a);
      a = fmaf(v1.x, h1[k][4], a);
      const float r = fmaxf(a, 0.f);
      float gh = v1.z * go[k][0];
      gh = fmaf(v1.w, go[k][1], gh);
      gh = fmaf(v2.x, go[k][2], gh);
      gh = fmaf(v2.y, go[k][3], gh);
      gh = fmaf(v2.z, go[k][4], gh);
      gh = a > 0.f ? gh : 0.f;
      gh1[k][0] = fmaf(v0.x, gh, gh1[k][0]);
      gh1[k][1] = fmaf(v0.y, gh, gh1[k][1]);
      gh1[k][2] = fmaf(v0.z, gh, gh1[k][2]);
      gh1[k][3] = fmaf(v0.w, gh, gh1[k][3]);
      gh1[k][4] = fmaf(v1.x, gh, gh1[k][4]);
#pragma unroll
      for (int e = 0; e < RD; ++e) {
        wg[e] = fmaf(r, go[k][e], wg[e]);              // dW3[slot e][h]
        wg[RD + e] = fmaf(gh, h1[k][e], wg[RD + e]);   // dWm[first-layer unit e][h]
      }
      wg[10] += gh;                                    // dbm[h]
    }
    const float tot = warp_reduce16(wg, lane);
    if (goff >= 0) atomicAdd(Gn + goff + h, tot);
  }
#pragma unroll
  for (int i = 0; i < RD; ++i) {
    const float4 v0 = w[2 * i], v1 = w[2 * i + 1];
    float wg[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) wg[j] = 0.f;
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
      const float ga = h1[k][i] > 0.f ? gh1[k][i] : 0.f;
      gc[k][0] = fmaf(v0.x, ga, gc[k][0]);
      gc[k][1] = fmaf(v0.y, ga, gc[k][1]);
      gc[k][2] = fmaf(v0.z, ga, gc[k][2]);
      gc[k][3] = fmaf(v0.w, ga, gc[k][3]);
      gc[k][4] = fmaf(v1.x, ga, gc[k][4]);
#pragma unroll
      for (int e = 0; e < RD; ++e) wg[e] = fmaf(ga, c[k][e], wg[e]);   // dW1[unit i][slot e]
      wg[RD] += ga;                                                      // db1[unit i]
    }
    const float tot = warp_reduce16(wg, lane);
    if (goff1 >= 0 && i < H1) atomicAdd(Gn + goff1 + i, tot);
  }
}

template <int R_NT, int SPT, int U, int MINB, bool M2 = false>
__global__ void __launch_bounds__(R_NT, MINB)
train_reg10_kernel(CnfDims d, const float* __restrict__ packed, const int* __restrict__ tables,
                   const float* __restrict__ xin, const int64_t* __restrict__ labels, float* __restrict__ partials,
                   double* __restrict__ loss_acc, int64_t N, float eps, float gamma, float inv_n) {
  extern __shared__ __align__(16) float smem[];
  __shared__ double red[4][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Hp = M2 ? d.H[1] : d.Hp[0], L = d.L;            // units streamed per net
  const int ns4 = 3 * Hp + (M2 ? 2 * RD : 0), lay4 = 2 * ns4;   // float4 per net / per layer
  const int Hrow = M2 ? d.Hp[1] : d.Hp[0];                  // row length of the packed matrices the unit gradients land in
  const int last = M2 ? 2 : 1;                              // index of the last Linear
  float4* ws = reinterpret_cast<float4*>(smem);
  float* b2s = reinterpret_cast<float*>(ws + (size_t)L * lay4);
  int* maps = reinterpret_cast<int*>(b2s + L * 16);       // per layer: [5] packed input index of slot e, [5] packed output index
  float* park = reinterpret_cast<float*>(maps + ((L * 10 + 3) / 4) * 4);   // [3][SPT][5][R_NT] per-thread parking slots
  const bool do_bwd = partials != nullptr;
  stage_weights_reg10<R_NT, M2>(d, packed, tables, reinterpret_cast<float*>(ws), b2s, maps, tid, Hp);

  const int TS = R_NT * SPT;
  const int64_t ntiles = (N + TS - 1) / TS;
  const bool rev_io = (L & 1) != 0;
  const int vidx = reduce16_index(lane);                  // which of the 16 reduced values this lane ends up holding
  const bool writer = (lane & 1) == 0;
  float* Grow = do_bwd ? partials + ((size_t)blockIdx.x * (R_NT / 32) + warp) * d.n_packed : nullptr;
  double a_loss = 0.0, a_ce = 0.0, a_ld = 0.0, a_bad = 0.0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t base = tile * TS;
    float lo[SPT][RD], hi[SPT][RD], ld[SPT];
    bool valid[SPT];
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
      const int64_t n = base + tid + k * R_NT;
      valid[k] = n < N;
      float v[RK];
      if (valid[k]) {
        const float2* p = reinterpret_cast<const float2*>(xin + n * RK);
#pragma unroll
        for (int j = 0; j < RD; ++j) { const float2 t2 = __ldg(p + j); v[2 * j] = t2.x; v[2 * j + 1] = t2.y; }
      } else {
#pragma unroll
        for (int j = 0; j < RK; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < RD; ++j) { lo[k][j] = v[j]; hi[k][j] = v[RD + j]; }
      ld[k] = 0.f;
    }
    // ---- forward ------------------------------------------------------------------------------------------------
    for (int l = 0; l < L; ++l) {
      const float4* w = ws + (size_t)l * lay4;
      const float* b2 = b2s + l * 16;
      if (l & 1) layer_eval<SPT, U, M2>(w, b2, Hp, 0, lo, hi, ld);
      else       layer_eval<SPT, U, M2>(w, b2, Hp, 0, hi, lo, ld);
    }
    // ---- loss head (calibrators.py:288-291; eps == 0: CrossEntropyLoss, run_experiment3D.py:107) ------------------
    float glo[SPT][RD], ghi[SPT][RD];
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
      const int64_t n = base + tid + k * R_NT;
      float zv[RK];     // logical order: z[j] = a[pi_L(j)]
#pragma unroll
      for (int j = 0; j < RD; ++j) {
        zv[j] = rev_io ? hi[k][RD - 1 - j] : lo[k][j];
        zv[RD + j] = rev_io ? lo[k][RD - 1 - j] : hi[k][j];
      }
      int yy = valid[k] ? (int)labels[n] : 0;
      yy = min(max(yy, 0), RK - 1);
      float mx = zv[0];
#pragma unroll
      for (int j = 1; j < RK; ++j) mx = fmaxf(mx, zv[j]);
      float se = 0.f, zy = 0.f;
      float pj[RK];
#pragma unroll
      for (int j = 0; j < RK; ++j) { pj[j] = expf(zv[j] - mx); se += pj[j]; zy = (j == yy) ? zv[j] : zy; }
      const float inv_se = 1.f / se;
      const float py = expf(zy - mx) * inv_se;
      float ce, coef;
      if (eps == 0.f) { ce = (zy - mx) - logf(se); coef = 1.f; }
      else            { ce = logf(py + eps); coef = py / (py + eps); }
      if (valid[k]) {
        const float tot = ce + gamma * ld[k];
        a_loss += (double)tot; a_ce += (double)ce; a_ld += (double)ld[k];
        if (!isfinite(tot)) a_bad += 1.0;
      }
      const float sc = valid[k] ? -inv_n * coef : 0.f;
      float gz[RK];
#pragma unroll
      for (int j = 0; j < RK; ++j) gz[j] = sc * ((j == yy ? 1.f : 0.f) - pj[j] * inv_se);
#pragma unroll
      for (int j = 0; j < RD; ++j) {          // back to physical slots
        glo[k][j] = rev_io ? gz[RK - 1 - j] : gz[j];
        ghi[k][j] = rev_io ? gz[RD - 1 - j] : gz[RD + j];
      }
    }
    if (!do_bwd) continue;
    // ---- backward -------------------------------------------------------------------------------------------------
    for (int l = L - 1; l >= 0; --l) {
      const float4* w = ws + (size_t)l * lay4;
      const float* b2 = b2s + l * 16;
      const int* mp = maps + l * 10;
      float* Gl = Grow + (size_t)l * d.layer_stride;
      // this lane's entry in a net's packed gradient block for the value it holds after the butterfly:
      // values 0..4 = dW(last Linear) of output slot e, 5..9 = dW(the Linear in front of the streamed units) of its
      // input e (a conditioning slot, or with M2 a first-layer unit), 10 = that Linear's bias
      int goff = -1, goff1 = -1;
      if (writer) {
        if (vidx < 5) goff = d.w_off[last] + mp[5 + vidx] * Hrow;
        else if (vidx < 10) goff = M2 ? (vidx - 5 < d.H[0] ? d.w_off[1] + (vidx - 5) * Hrow : -1) : d.w_off[0] + mp[vidx - 5] * Hrow;
        else if (vidx == 10) goff = d.b_off[last - 1];
        if (M2) {                                   // first-layer values: 0..4 dW1 of input slot e, 5 db1
          if (vidx < 5) goff1 = d.w_off[0] + mp[vidx] * d.Hp[0];
          else if (vidx == 5) goff1 = d.b_off[0];
        }
      }
      const int boff = (writer && vidx < 5) ? d.b_off[last] + mp[5 + vidx] : -1;      // db(last) of output slot vidx
      // Register pressure: of the four [SPT][5] blocks of a layer (conditioning values c, transformed values t and the
      // gradients on both) only two or three are used inside each hidden-unit loop; the others wait in a per-thread
      // shared-memory slot (park[slot][value][tid]: conflict-free), which is what lets SPT reach 8.
      auto park_put = [&](int slot, const float (&a)[SPT][RD]) {
#pragma unroll
        for (int k = 0; k < SPT; ++k)
#pragma unroll
          for (int q = 0; q < RD; ++q) park[((slot * SPT + k) * RD + q) * R_NT + tid] = a[k][q];
      };
      auto park_get = [&](int slot, float (&a)[SPT][RD]) {
#pragma unroll
        for (int k = 0; k < SPT; ++k)
#pragma unroll
          for (int q = 0; q < RD; ++q) a[k][q] = park[((slot * SPT + k) * RD + q) * R_NT + tid];
      };
      auto run = [&](float (&c)[SPT][RD], float (&t)[SPT][RD], float (&gcnd)[SPT][RD], float (&gt)[SPT][RD]) {
        // recompute s and shift from the conditioning half; step the transformed half back to the layer input
        float o[SPT][RD];
        park_put(0, gcnd);
        park_put(1, gt);
        net_eval<SPT, U, M2>(w + ns4, b2 + 8, Hp, c, o);                 // shift
#pragma unroll
        for (int k = 0; k < SPT; ++k)
#pragma unroll
          for (int q = 0; q < RD; ++q) t[k][q] -= o[k][q];              // y - shift  (= x e^s)
        park_put(2, t);
        net_eval<SPT, U, M2>(w, b2, Hp, c, o);                           // s
        park_get(2, t);
        park_get(1, gt);
        float gs[SPT][RD];
#pragma unroll
        for (int k = 0; k < SPT; ++k)
#pragma unroll
          for (int q = 0; q < RD; ++q) {
            const float gy = gt[k][q];
            gs[k][q] = fmaf(gy, t[k][q], valid[k] ? -gamma * inv_n : 0.f);   // g_s = g_y x e^s + g_ld
            t[k][q] *= expf(-o[k][q]);                                   // x
            o[k][q] = gy * expf(o[k][q]);                                // g_x of the transformed half
          }
        {
          float wb[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) wb[i] = 0.f;
#pragma unroll
          for (int k = 0; k < SPT; ++k)
#pragma unroll
            for (int q = 0; q < RD; ++q) { wb[q] += gt[k][q]; wb[8 + q] += gs[k][q]; }
          // values 0..4: db2 of the shift net, 8..12: db2 of the scale net
          const float tot = warp_reduce16(wb, lane);
          if (writer && vidx < 5) atomicAdd(Gl + d.net_stride + d.b_off[last] + mp[5 + vidx], tot);
          if (writer && vidx >= 8 && vidx < 13) atomicAdd(Gl + d.b_off[last] + mp[5 + vidx - 8], tot);
        }
        park_put(1, t);          // x: final for this layer
        park_put(2, o);          // g_x of the transformed half: becomes gt below
        park_get(0, gcnd);
        park_put(0, gs);
        // shift net first (its output gradient is g_y itself), then the scale net
        if constexpr (M2) net_backward_m2<SPT, U>(w + ns4, Hp, d.H[0], c, gt, gcnd, Gl + d.net_stride, goff, goff1, lane);
        else              net_backward<SPT, U>(w + ns4, Hp, c, gt, gcnd, Gl + d.net_stride, goff, lane);
        park_get(0, gt);         // gs
        if constexpr (M2) net_backward_m2<SPT, U>(w, Hp, d.H[0], c, gt, gcnd, Gl, goff, goff1, lane);
        else              net_backward<SPT, U>(w, Hp, c, gt, gcnd, Gl, goff, lane);
        park_get(1, t);
        park_get(2, gt);
      };
      (void)boff; (void)goff1;
      if (l & 1) run(lo, hi, glo, ghi);
      else       run(hi, lo, ghi, glo);
    }
  }
  // ---- loss sums: lanes -> warp -> CTA -> global (float64) ---------------------------------------------------------
  if (loss_acc != nullptr) {
    double v4[4] = {a_loss, a_ce, a_ld, a_bad};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v4[q] += __shfl_xor_sync(0xffffffffu, v4[q], o);
      if (lane == 0) red[q][warp] = v4[q];
    }
    __syncthreads();
    if (tid < 4) {
      double s = 0.0;
      for (int w = 0; w < R_NT / 32; ++w) s += red[tid][w];
      atomicAdd(loss_acc + tid, s);
    }
  }
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
                    int sms, int max_smem, int variant, int64_t* rows_out, cudaStream_t st) {
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
    train_reg10_kernel<NT, SPT, U, MB, M2><<<grid, NT, smem, st>>>(d, packed, tables, x, y, partials, loss_acc, N, eps, gamma, inv_n); \
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
