// Device code shared by the register-resident fp32 kernels (cnf_flow_fp32r.cu: K = 10, every variant;
// cnf_flow_fp32rk.cu: the training kernel for K = 2 .. 9).  Included inside each file's translation unit; everything
// sits in an anonymous namespace.
//
// A thread keeps its samples' logits in two register arrays of five, lo and hi.  K = 10: physical slots 0..4 and
// 5..9.  Any K <= 10 (KK template parameter, d0 = K/2 transformed and d1 = K - d0 conditioning dims per layer,
// flows/flows.py:81-86): lo holds physical slots [0, d0), hi the upper d0 slots [d1, K); for odd K the middle slot
// d0 -- a conditioning dim of EVERY layer, never transformed -- is kept in both, at position d0.  Even
// layers condition on hi and transform lo, odd layers the other way round; array positions past what K uses carry
// zeros and zero weights, so their "outputs" are s = t = 0 exactly and x e^0 + 0 leaves them alone.
#pragma once
#include <cuda_runtime.h>

#include "cnf_common.h"

namespace {

constexpr int RK = 10, RD = 5;          // classes of the forward kernels, register-array length

// physical slot -> position in the lo / hi array (-1: not held there)
template <int KK> struct RegMap {
  static constexpr int D0 = KK / 2, D1 = KK - KK / 2;
  static constexpr bool ODD = (KK & 1) != 0;
  static __host__ __device__ constexpr int lo_pos(int p) { return p < D0 ? p : ((ODD && p == D0) ? D0 : -1); }
  static __host__ __device__ constexpr int hi_pos(int p) { return p >= D1 ? p - D1 : ((ODD && p == D0) ? D0 : -1); }
};

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
                                           const float (&c)[SPT][RD], float (&t)[SPT][RD], float (&ld)[SPT], int nets = 3) {
  float o[SPT][RD];
  const int ns4 = 3 * Hp + (M2 ? 2 * RD : 0);               // float4 per net: [first-layer records] unit records
  if (!inverse) {
    if (nets & 1) {
      net_eval<SPT, U, M2>(w, b2, Hp, c, o);                 // s
#pragma unroll
      for (int k = 0; k < SPT; ++k)
#pragma unroll
        for (int q = 0; q < RD; ++q) { t[k][q] *= expf(o[k][q]); ld[k] += o[k][q]; }
    }
    if (nets & 2) {
      net_eval<SPT, U, M2>(w + ns4, b2 + 8, Hp, c, o);       // shift
#pragma unroll
      for (int k = 0; k < SPT; ++k)
#pragma unroll
        for (int q = 0; q < RD; ++q) t[k][q] += o[k][q];
    }
  } else {
    if (nets & 2) {
      net_eval<SPT, U, M2>(w + ns4, b2 + 8, Hp, c, o);       // shift
#pragma unroll
      for (int k = 0; k < SPT; ++k)
#pragma unroll
        for (int q = 0; q < RD; ++q) t[k][q] -= o[k][q];
    }
    if (nets & 1) {
      net_eval<SPT, U, M2>(w, b2, Hp, c, o);                 // s
#pragma unroll
      for (int k = 0; k < SPT; ++k)
#pragma unroll
        for (int q = 0; q < RD; ++q) { t[k][q] *= expf(-o[k][q]); ld[k] -= o[k][q]; }
    }
  }
}

// Stage the weights of all layers: packed [d1][Hp] / [Hp] / [d0][Hp] / [d0p] per net -> per hidden unit 12 floats
// [W1 of position 0..4 | b1 | W2 of position 0..4 | 0], inputs and outputs placed by their POSITION in the conditioning /
// transformed register array (RegMap; even layers condition on hi, odd ones on lo).  maps[l][0..4] = packed input index
// of conditioning position e, maps[l][5..9] = packed output index of transformed position e, -1 where K does not use
// the position (formed first so that the copy below reads the packed blob with h fastest, i.e. coalesced, and never
// searches).  A net the flow does not have (NICE: no scale net) is staged as zeros and never evaluated.
template <int R_NT, bool M2 = false, int KK = RK>
__device__ __forceinline__ void stage_weights_reg10(const CnfDims& d, const float* __restrict__ packed,
                                                    const int* __restrict__ tables, float* wf, float* b2s, int* maps, int tid, int Hn) {
  // M2: per net 5 first-layer records of 8 floats [W1 of position 0..4 | b1 | 0 0], then one 12-float record per unit of
  // the SECOND hidden layer (true width, not padded) [Wm from first-layer unit 0..4 | bm | W3 of position 0..4 | 0]
  // Hn = hidden units streamed per net: the true width for the forward kernel, the padded one for the training kernel
  using MP = RegMap<KK>;
  const int Hp = Hn, L = d.L;
  const int pre = M2 ? 8 * RD : 0;                         // floats in front of a net's unit records
  const int last = M2 ? 2 : 1;                             // index of the last Linear
  for (int i = tid; i < L * 10; i += R_NT) {
    const int l = i / 10, e = i - l * 10;
    const int* cond = tables + d.tab_cond + l * MP::D1;
    const int* trans = tables + d.tab_trans + l * MP::D0;
    int v = -1;
    if (e < 5) { for (int j = 0; j < MP::D1; ++j) if (((l & 1) ? MP::lo_pos(cond[j]) : MP::hi_pos(cond[j])) == e) v = j; }
    else       { for (int q = 0; q < MP::D0; ++q) if (((l & 1) ? MP::hi_pos(trans[q]) : MP::lo_pos(trans[q])) == e - 5) v = q; }
    maps[i] = v;
  }
  __syncthreads();
  const int per_net = 12 * Hp;
  const int Hrow = M2 ? d.Hp[1] : d.Hp[0];                 // row length of the packed matrices the unit records read
  for (int i = tid; i < L * 2 * per_net; i += R_NT) {
    const int ln = i / per_net, r = i - ln * per_net;      // ln = l * 2 + net
    const int e = r / Hp, h = r - e * Hp;                  // h fastest: coalesced reads of the packed rows
    const int l = ln >> 1, net = ln & 1;
    const float* Wn = packed + (size_t)l * d.layer_stride + (size_t)((net == 1 && (d.nets & 1)) ? 1 : 0) * d.net_stride;
    const int* mp = maps + l * 10;
    float v = 0.f;
    if (d.nets & (1 << net)) {
      if (M2) {
        if (e < 5) v = e < d.H[0] ? __ldg(Wn + d.w_off[1] + e * Hrow + h) : 0.f;      // middle Linear [in][out]
        else if (e == 5) v = __ldg(Wn + d.b_off[1] + h);
        else if (e < 11) v = mp[e - 1] >= 0 ? __ldg(Wn + d.w_off[2] + mp[e - 1] * Hrow + h) : 0.f;
      } else {
        if (e < 5) v = mp[e] >= 0 ? __ldg(Wn + d.w_off[0] + mp[e] * Hrow + h) : 0.f;
        else if (e == 5) v = __ldg(Wn + d.b_off[0] + h);
        else if (e < 11) v = mp[e - 1] >= 0 ? __ldg(Wn + d.w_off[1] + mp[e - 1] * Hrow + h) : 0.f;
      }
    }
    wf[(size_t)ln * (pre + per_net) + pre + (size_t)h * 12 + e] = v;
  }
  if constexpr (M2) {
    for (int i = tid; i < L * 2 * pre; i += R_NT) {
      const int ln = i / pre, r = i - ln * pre;
      const int e = r / RD, u = r - e * RD;                // u = first-layer unit
      const int l = ln >> 1, net = ln & 1;
      const float* Wn = packed + (size_t)l * d.layer_stride + (size_t)((net == 1 && (d.nets & 1)) ? 1 : 0) * d.net_stride;
      const int* mp = maps + l * 10;
      float v = 0.f;
      if (u < d.H[0] && (d.nets & (1 << net))) {
        if (e < 5) v = mp[e] >= 0 ? __ldg(Wn + d.w_off[0] + mp[e] * d.Hp[0] + u) : 0.f;
        else if (e == 5) v = __ldg(Wn + d.b_off[0] + u);
      }
      wf[(size_t)ln * (pre + per_net) + (size_t)u * 8 + e] = v;
    }
  }
  for (int i = tid; i < L * 16; i += R_NT) {
    const int l = i >> 4, net = (i >> 3) & 1, e = i & 7;
    const float* Wn = packed + (size_t)l * d.layer_stride + (size_t)((net == 1 && (d.nets & 1)) ? 1 : 0) * d.net_stride;
    const int q = e < RD ? maps[l * 10 + 5 + e] : -1;
    b2s[i] = (q >= 0 && (d.nets & (1 << net))) ? __ldg(Wn + d.b_off[last] + q) : 0.f;
  }
  __syncthreads();
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
// The same over 8 values (K <= 5: a hidden unit's weight gradients are at most 2 + 5 + 1 values): 4 + 2 + 1 + 1 + 1 = 9
// shuffles; on return v[0] of lane L is the warp-wide sum of value reduce8_index(L), the same on the 4 lanes of a group.
__device__ __forceinline__ int reduce8_index(int lane) { return ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1); }
__device__ __forceinline__ float warp_reduce8(float (&v)[8], int lane) {
#pragma unroll
  for (int w = 4, bit = 16; w >= 1; w >>= 1, bit >>= 1) {
    const bool up = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i < w) {
        const float send = up ? v[i] : v[i + w];
        const float keep = up ? v[i + w] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
      }
    }
  }
  const float t = v[0] + __shfl_xor_sync(0xffffffffu, v[0], 2);
  return t + __shfl_xor_sync(0xffffffffu, t, 1);
}
// Where a hidden unit's weight-gradient values sit in the reduced vector.  K >= 6: 16 values, [0,5) last-Linear row
// entries of output position e, [5,10) the Linear in front of the streamed units (input e), 10 its bias.  K <= 5 (at
// most two transformed and three conditioning dims): 8 values, [0,2) outputs, [2,7) inputs, 7 the bias.
template <int KK> struct GradLayout {
  static constexpr bool SMALL = KK <= 5;
  static constexpr int NV = SMALL ? 8 : 16, NO = SMALL ? 2 : RD, P_O = 0, P_I = SMALL ? 2 : RD, P_B = SMALL ? 7 : 10;
  static constexpr int PB_S = SMALL ? 4 : 8;     // bias butterfly: shift-net sums at [0, NO), scale-net sums at [PB_S, PB_S + NO)
  static __device__ __forceinline__ int index(int lane) { return SMALL ? reduce8_index(lane) : reduce16_index(lane); }
  static __device__ __forceinline__ bool writer(int lane) { return SMALL ? (lane & 3) == 0 : (lane & 1) == 0; }
};
template <int NV> struct WarpReduce;
template <> struct WarpReduce<16> { static __device__ __forceinline__ float run(float (&v)[16], int lane); };
template <> struct WarpReduce<8> { static __device__ __forceinline__ float run(float (&v)[8], int lane) { return warp_reduce8(v, lane); } };
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
__device__ __forceinline__ float WarpReduce<16>::run(float (&v)[16], int lane) { return warp_reduce16(v, lane); }

// Backward of one conditioner net for this thread's SPT samples (c: conditioning values, go: gradient on the net's
// five outputs).  Adds W1^T gh into gc and the weight gradients into the warp's partial row Gn (packed layout of one
// net; goff = this lane's entry offset for reduce16_index(lane), -1 for none, see train kernel).
template <int SPT, int U, int KK = RK>
__device__ __forceinline__ void net_backward(const float4* __restrict__ w, int Hp, const float (&c)[SPT][RD],
                                             const float (&go)[SPT][RD], float (&gc)[SPT][RD], float* __restrict__ Gn,
                                             int goff, int lane) {
#pragma unroll U
  for (int h = 0; h < Hp; ++h) {
    const float4 v0 = w[3 * h], v1 = w[3 * h + 1], v2 = w[3 * h + 2];
    using GL = GradLayout<KK>;
    float wg[GL::NV];
#pragma unroll
    for (int i = 0; i < GL::NV; ++i) wg[i] = 0.f;
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
      for (int e = 0; e < GL::NO; ++e) wg[GL::P_O + e] = fmaf(r, go[k][e], wg[GL::P_O + e]);      // dW2[slot e][h]
#pragma unroll
      for (int e = 0; e < RD; ++e) wg[GL::P_I + e] = fmaf(gh, c[k][e], wg[GL::P_I + e]);          // dW1[h][slot e]
      wg[GL::P_B] += gh;                                                                          // db1[h]
    }
    const float tot = WarpReduce<GL::NV>::run(wg, lane);
    if (goff >= 0) atomicAdd(Gn + goff + h, tot);
  }
}

// The same for a net with two hidden layers, the first of at most five units (hidden_size=[5, 5]): the first layer is
// recomputed into registers, the stream over the second layer's units also accumulates the gradient on the first
// layer's activations, and a last pass over the five first-layer units forms W1^T ga, dW1 and db1.
// goff: entry offset of this lane for the second-layer / last-Linear values (0..4 dW3 of output slot e, 5..9 dWm from
// first-layer unit e, 10 dbm); goff1: for the first-layer values (0..4 dW1 of input slot e, 5 db1); -1 for none.
template <int SPT, int U, int KK = RK>
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
    using GL = GradLayout<KK>;
    float wg[GL::NV];
#pragma unroll
    for (int i = 0; i < GL::NV; ++i) wg[i] = 0.f;
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
      for (int e = 0; e < GL::NO; ++e) wg[GL::P_O + e] = fmaf(r, go[k][e], wg[GL::P_O + e]);      // dW3[slot e][h]
#pragma unroll
      for (int e = 0; e < RD; ++e) wg[GL::P_I + e] = fmaf(gh, h1[k][e], wg[GL::P_I + e]);         // dWm[first-layer unit e][h]
      wg[GL::P_B] += gh;                                                                          // dbm[h]
    }
    const float tot = WarpReduce<GL::NV>::run(wg, lane);
    if (goff >= 0) atomicAdd(Gn + goff + h, tot);
  }
#pragma unroll
  for (int i = 0; i < RD; ++i) {
    if (i >= H1) break;       // units the first hidden layer does not have (zero records: they would add exact zeros)
    const float4 v0 = w[2 * i], v1 = w[2 * i + 1];
    float wg[8];      // six values per first-layer unit: the 8-value butterfly whatever K is (goff1 follows reduce8_index)
#pragma unroll
    for (int j = 0; j < 8; ++j) wg[j] = 0.f;
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
    const float tot = warp_reduce8(wg, lane);
    if (goff1 >= 0) atomicAdd(Gn + goff1 + i, tot);
  }
}

template <int R_NT, int SPT, int U, int MINB, bool M2 = false, int KK = RK>
__global__ void __launch_bounds__(R_NT, MINB)
train_reg10_kernel(CnfDims d, const float* __restrict__ packed, const int* __restrict__ tables,
                   const float* __restrict__ xin, const int64_t* __restrict__ labels, float* __restrict__ partials,
                   double* __restrict__ loss_acc, int64_t N, float eps, float gamma, float inv_n,
                   const float* __restrict__ gz_ext = nullptr, const float* __restrict__ gld_ext = nullptr) {
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
  using MP = RegMap<KK>;
  const int nets = d.nets;
  const size_t off_t = (nets & 1) ? (size_t)d.net_stride : 0;      // the shift net's block inside a layer's packed gradients
  stage_weights_reg10<R_NT, M2, KK>(d, packed, tables, reinterpret_cast<float*>(ws), b2s, maps, tid, Hp);

  const int TS = R_NT * SPT;
  const int64_t ntiles = (N + TS - 1) / TS;
  const bool rev_io = (L & 1) != 0;
  using GL = GradLayout<KK>;
  const int vidx = GL::index(lane);                       // which of the reduced values this lane ends up holding
  const bool writer = GL::writer(lane);
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
      float v[KK];
      if (valid[k]) {
        if constexpr (KK == RK) {
          const float2* p = reinterpret_cast<const float2*>(xin + n * RK);
#pragma unroll
          for (int j = 0; j < RD; ++j) { const float2 t2 = __ldg(p + j); v[2 * j] = t2.x; v[2 * j + 1] = t2.y; }
        } else {
#pragma unroll
          for (int j = 0; j < KK; ++j) v[j] = __ldg(xin + n * KK + j);
        }
      } else {
#pragma unroll
        for (int j = 0; j < KK; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < RD; ++j) {          // physical slot -> register position (RegMap); unused positions carry zeros
        lo[k][j] = j < MP::D0 ? v[j] : ((MP::ODD && j == MP::D0) ? v[MP::D0] : 0.f);
        hi[k][j] = j < MP::D0 ? v[MP::D1 + j] : ((MP::ODD && j == MP::D0) ? v[MP::D0] : 0.f);
      }
      ld[k] = 0.f;
    }
    // ---- forward ------------------------------------------------------------------------------------------------
    for (int l = 0; l < L; ++l) {
      const float4* w = ws + (size_t)l * lay4;
      const float* b2 = b2s + l * 16;
      if (l & 1) layer_eval<SPT, U, M2>(w, b2, Hp, 0, lo, hi, ld, nets);
      else       layer_eval<SPT, U, M2>(w, b2, Hp, 0, hi, lo, ld, nets);
    }
    // ---- loss head (calibrators.py:288-291; eps == 0: CrossEntropyLoss, run_experiment3D.py:107) ------------------
    // (gz_ext != nullptr: external head -- the upstream gradients dL/dz [N, K] and dL/dlog_det [N] come from the
    //  caller, the autograd backward of the drop-in Flow, run_experiment3D.py:133-134; no loss is formed)
    float glo[SPT][RD], ghi[SPT][RD], gld[SPT];
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
      const int64_t n = base + tid + k * R_NT;
      if (gz_ext != nullptr) {
        float gz[KK];
#pragma unroll
        for (int j = 0; j < KK; ++j) gz[j] = valid[k] ? __ldg(gz_ext + n * KK + j) : 0.f;
        gld[k] = valid[k] ? __ldg(gld_ext + n) : 0.f;
#pragma unroll
        for (int j = 0; j < RD; ++j) {          // logical -> physical slot -> register position, as below
          glo[k][j] = j < MP::D0 ? (rev_io ? gz[KK - 1 - j] : gz[j])
                                 : ((MP::ODD && j == MP::D0) ? (rev_io ? gz[KK - 1 - MP::D0] : gz[MP::D0]) : 0.f);
          ghi[k][j] = j < MP::D0 ? (rev_io ? gz[KK - 1 - (MP::D1 + j)] : gz[MP::D1 + j]) : 0.f;
        }
        continue;
      }
      gld[k] = valid[k] ? -gamma * inv_n : 0.f;
      float av[KK];     // by physical slot
#pragma unroll
      for (int p = 0; p < KK; ++p) av[p] = p < MP::D1 ? lo[k][p < MP::D0 ? p : MP::D0] : hi[k][p - MP::D1];
      float zv[KK];     // logical order: z[j] = a[pi_L(j)], pi_L = the reversal for odd L
#pragma unroll
      for (int j = 0; j < KK; ++j) zv[j] = rev_io ? av[KK - 1 - j] : av[j];
      int yy = valid[k] ? (int)labels[n] : 0;
      yy = min(max(yy, 0), KK - 1);
      float mx = zv[0];
#pragma unroll
      for (int j = 1; j < KK; ++j) mx = fmaxf(mx, zv[j]);
      float se = 0.f, zy = 0.f;
      float pj[KK];
#pragma unroll
      for (int j = 0; j < KK; ++j) { pj[j] = expf(zv[j] - mx); se += pj[j]; zy = (j == yy) ? zv[j] : zy; }
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
      float gz[KK];
#pragma unroll
      for (int j = 0; j < KK; ++j) gz[j] = sc * ((j == yy ? 1.f : 0.f) - pj[j] * inv_se);
      float ga[KK];     // back to physical slots
#pragma unroll
      for (int p = 0; p < KK; ++p) ga[p] = rev_io ? gz[KK - 1 - p] : gz[p];
#pragma unroll
      for (int j = 0; j < RD; ++j) {          // (odd K: the middle slot's gradient rides in lo; it is never transformed, so
        glo[k][j] = j < MP::D0 ? ga[j] : ((MP::ODD && j == MP::D0) ? ga[MP::D0] : 0.f);   //  no weight gradient reads it)
        ghi[k][j] = j < MP::D0 ? ga[MP::D1 + j] : 0.f;
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
      // (GradLayout; K >= 6:) values 0..4 = dW(last Linear) of output slot e, 5..9 = dW(the Linear in front of the streamed units) of its
      // input e (a conditioning slot, or with M2 a first-layer unit), 10 = that Linear's bias
      int goff = -1, goff1 = -1;
      if (writer) {         // (positions K does not use have no packed entry: -1)
        if (vidx < GL::P_O + GL::NO) goff = mp[5 + vidx - GL::P_O] >= 0 ? d.w_off[last] + mp[5 + vidx - GL::P_O] * Hrow : -1;
        else if (vidx >= GL::P_I && vidx < GL::P_I + RD)
          goff = M2 ? (vidx - GL::P_I < d.H[0] ? d.w_off[1] + (vidx - GL::P_I) * Hrow : -1)
                    : (mp[vidx - GL::P_I] >= 0 ? d.w_off[0] + mp[vidx - GL::P_I] * Hrow : -1);
        else if (vidx == GL::P_B) goff = d.b_off[last - 1];
      }
      if (M2 && (lane & 3) == 0) {                  // first-layer values (8-value butterfly): 0..4 dW1 of input slot e, 5 db1
        const int v8 = reduce8_index(lane);
        if (v8 < 5) goff1 = mp[v8] >= 0 ? d.w_off[0] + mp[v8] * d.Hp[0] : -1;
        else if (v8 == 5) goff1 = d.b_off[0];
      }
      const int boff = (writer && vidx < GL::NO && mp[5 + vidx] >= 0) ? d.b_off[last] + mp[5 + vidx] : -1;      // db(last) of output slot vidx
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
        if (nets & 2) {
          net_eval<SPT, U, M2>(w + ns4, b2 + 8, Hp, c, o);               // shift
#pragma unroll
          for (int k = 0; k < SPT; ++k)
#pragma unroll
            for (int q = 0; q < RD; ++q) t[k][q] -= o[k][q];            // y - shift  (= x e^s)
        }
        park_put(2, t);
        if (nets & 1) {
          net_eval<SPT, U, M2>(w, b2, Hp, c, o);                         // s
        } else {
#pragma unroll
          for (int k = 0; k < SPT; ++k)
#pragma unroll
            for (int q = 0; q < RD; ++q) o[k][q] = 0.f;                  // NICE: s = 0 (e^0 = 1 exactly below)
        }
        park_get(2, t);
        park_get(1, gt);
        float gs[SPT][RD];
#pragma unroll
        for (int k = 0; k < SPT; ++k)
#pragma unroll
          for (int q = 0; q < RD; ++q) {
            const float gy = gt[k][q];
            gs[k][q] = fmaf(gy, t[k][q], gld[k]);                            // g_s = g_y x e^s + g_ld
            t[k][q] *= expf(-o[k][q]);                                   // x
            o[k][q] = gy * expf(o[k][q]);                                // g_x of the transformed half
          }
        {
          float wb[GL::NV];
#pragma unroll
          for (int i = 0; i < GL::NV; ++i) wb[i] = 0.f;
#pragma unroll
          for (int k = 0; k < SPT; ++k)
#pragma unroll
            for (int q = 0; q < GL::NO; ++q) { wb[q] += gt[k][q]; wb[GL::PB_S + q] += gs[k][q]; }
          // values [0, NO): db2 of the shift net, [PB_S, PB_S + NO): db2 of the scale net
          const float tot = WarpReduce<GL::NV>::run(wb, lane);
          if (boff >= 0 && (nets & 2)) atomicAdd(Gl + off_t + boff, tot);
          if (writer && vidx >= GL::PB_S && vidx < GL::PB_S + GL::NO && (nets & 1) && mp[5 + vidx - GL::PB_S] >= 0)
            atomicAdd(Gl + d.b_off[last] + mp[5 + vidx - GL::PB_S], tot);
        }
        park_put(1, t);          // x: final for this layer
        park_put(2, o);          // g_x of the transformed half: becomes gt below
        park_get(0, gcnd);
        park_put(0, gs);
        // shift net first (its output gradient is g_y itself), then the scale net
        if (nets & 2) {
          if constexpr (M2) net_backward_m2<SPT, U, KK>(w + ns4, Hp, d.H[0], c, gt, gcnd, Gl + off_t, goff, goff1, lane);
          else              net_backward<SPT, U, KK>(w + ns4, Hp, c, gt, gcnd, Gl + off_t, goff, lane);
        }
        park_get(0, gt);         // gs
        if (nets & 1) {
          if constexpr (M2) net_backward_m2<SPT, U, KK>(w, Hp, d.H[0], c, gt, gcnd, Gl, goff, goff1, lane);
          else              net_backward<SPT, U, KK>(w, Hp, c, gt, gcnd, Gl, goff, lane);
        }
        park_get(1, t);
        park_get(2, gt);
      };
      (void)goff1;
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

}  // namespace
