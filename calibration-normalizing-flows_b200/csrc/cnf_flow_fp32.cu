// fp32 CUDA-core kernels of the coupling-flow stack (sm_100a): the 1e-5 parity path for
// every shape, and the only path for shapes the tensor-core kernel does not cover.
//
// One fused kernel per flow stack.  A CTA owns a tile of TS = NT*SPT samples; one thread
// owns SPT samples for the whole stack (all L coupling layers), so there is no barrier
// between layers.  The tile lives in shared memory feature-major, act[slot][sample]
// (row stride TSP = TS+4: conflict-free for per-sample scalar access and for the 128-bit
// row reads of the weight-gradient pass).  Packed weights of all layers are staged once
// per CTA in shared memory when they fit (read with broadcast LDS.128), otherwise read
// through L1 with __ldg.  Flips/permutations are index maps (tables), data never moves.
//
// Reference arithmetic restated here:
//   coupling forward  z = x_b + (1-m)(x*exp(s)+t), ld = sum (1-m)s     flows/flows.py:101-112
//   coupling inverse  x = x_b + (1-m)(z-t)*exp(-s), ld = -sum (1-m)s   flows/flows.py:114-126
//   conditioner MLP   Linear-ReLU-...-Linear                            flows/utils.py:26-31
//   NLL head          -mean(log(softmax(z)[y]+eps) + gamma*ld)          calibrators.py:288-291
//                     (eps=0: CrossEntropyLoss)                         run_experiment3D.py:107
#include <cuda_runtime.h>

#include <cstdlib>

#include "cnf_common.h"
#include "cnf_metrics_dev.cuh"

namespace {

constexpr int CH = CNF_CH;

template <bool WS>
__device__ __forceinline__ float4 ldw4(const float* p) {
  if (WS) return *reinterpret_cast<const float4*>(p);
  return __ldg(reinterpret_cast<const float4*>(p));
}
template <bool WS>
__device__ __forceinline__ float ldw1(const float* p) {
  if (WS) return *p;
  return __ldg(p);
}

// Per-CTA shared-memory carve-up (in floats from the dynamic smem base).
struct Smem {
  int w;       // packed weights (WS only)
  int tab;     // int tables
  int act;     // [K][TSP]
  int outs;    // [2][d0][TSP]   s and t outputs
  int hid;     // [hid_rows][TSP]
  // backward only
  int gact;    // [K][TSP]
  int tape;    // [L][d0][TSP]
  int gout;    // [2][d0][TSP]
  int gbuf;    // [2][Hmax][TSP] (m >= 2)
  int tail;    // fused statistics tail (forward kernel only): cnf_tail_smem_bytes / 4 floats
  int total;   // floats
};

__host__ __device__ inline int hid_rows(const CnfDims& d, bool store_last) {
  int r = 0;
  for (int j = 0; j < d.m; ++j)
    if (j < d.m - 1 || store_last) r += d.Hp[j];
  return r;
}

// ws: 0 = weights stay in global memory, 1 = all layers staged once, 2 = one layer staged at a time
__host__ __device__ inline Smem make_smem(const CnfDims& d, int TSP, int ws, bool backward, int tail_bins = 0) {
  Smem s;
  int off = 0;
  s.tail = off; off += tail_bins > 0 ? cnf_tail_smem_bytes(tail_bins, d.K) / 4 : 0;   // first: 16-byte aligned doubles
  s.w = off; off += ws == 1 ? d.n_packed : (ws == 2 ? (int)d.layer_stride : 0);
  s.tab = off; off += (d.n_tables + 3) / 4 * 4;
  s.act = off; off += d.K * TSP;
  s.outs = off; off += 2 * d.d0 * TSP;
  s.hid = off; off += hid_rows(d, backward) * TSP;
  s.gact = off; off += backward ? d.K * TSP : 0;
  s.tape = off; off += backward ? d.L * d.d0 * TSP : 0;
  s.gout = off; off += backward ? 2 * d.d0 * TSP : 0;
  s.gbuf = off; off += (backward && d.m >= 2) ? 2 * d.Hmax * TSP : 0;
  s.total = off;
  return s;
}

// h[k][r] = b[r0+r] + sum_i W[i*ldw + r0 + r] * src[row(i)][sample k]
template <int SPT, bool WS>
__device__ __forceinline__ void chunk_from_inputs(float (&h)[SPT][CH], const float* W, int ldw, const float* b,
                                                  int r0, int n_in, const float* src, const int* idx, int TSP,
                                                  int tid, int NT) {
  if (b != nullptr) {
#pragma unroll
    for (int r4 = 0; r4 < CH / 4; ++r4) {
      float4 bv = ldw4<WS>(b + r0 + 4 * r4);
#pragma unroll
      for (int k = 0; k < SPT; ++k) {
        h[k][4 * r4 + 0] = bv.x; h[k][4 * r4 + 1] = bv.y; h[k][4 * r4 + 2] = bv.z; h[k][4 * r4 + 3] = bv.w;
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < SPT; ++k)
#pragma unroll
      for (int r = 0; r < CH; ++r) h[k][r] = 0.f;
  }
#pragma unroll 4
  for (int i = 0; i < n_in; ++i) {
    const int row = idx ? idx[i] : i;
    float xin[SPT];
#pragma unroll
    for (int k = 0; k < SPT; ++k) xin[k] = src[row * TSP + tid + k * NT];
    const float* wrow = W + (size_t)i * ldw + r0;
#pragma unroll
    for (int r4 = 0; r4 < CH / 4; ++r4) {
      float4 wv = ldw4<WS>(wrow + 4 * r4);
#pragma unroll
      for (int k = 0; k < SPT; ++k) {
        h[k][4 * r4 + 0] = fmaf(wv.x, xin[k], h[k][4 * r4 + 0]);
        h[k][4 * r4 + 1] = fmaf(wv.y, xin[k], h[k][4 * r4 + 1]);
        h[k][4 * r4 + 2] = fmaf(wv.z, xin[k], h[k][4 * r4 + 2]);
        h[k][4 * r4 + 3] = fmaf(wv.w, xin[k], h[k][4 * r4 + 3]);
      }
    }
  }
}

// for each of n_out rows o: acc = sum_r W[o*ldw + r0 + r] * h[k][r];
//   dst[row(o)][sample k] = (init ? (b ? b[o] : 0) : dst) + acc
template <int SPT, bool WS>
__device__ __forceinline__ void chunk_to_outputs(const float (&h)[SPT][CH], const float* W, int ldw, const float* b,
                                                 int r0, int n_out, float* dst, const int* idx, bool init,
                                                 bool accumulate_always, int TSP, int tid, int NT) {
#pragma unroll 2
  for (int o = 0; o < n_out; ++o) {
    const float* wrow = W + (size_t)o * ldw + r0;
    float acc[SPT];
#pragma unroll
    for (int k = 0; k < SPT; ++k) acc[k] = 0.f;
#pragma unroll
    for (int r4 = 0; r4 < CH / 4; ++r4) {
      float4 wv = ldw4<WS>(wrow + 4 * r4);
#pragma unroll
      for (int k = 0; k < SPT; ++k) {
        acc[k] = fmaf(wv.x, h[k][4 * r4 + 0], acc[k]);
        acc[k] = fmaf(wv.y, h[k][4 * r4 + 1], acc[k]);
        acc[k] = fmaf(wv.z, h[k][4 * r4 + 2], acc[k]);
        acc[k] = fmaf(wv.w, h[k][4 * r4 + 3], acc[k]);
      }
    }
    const int row = idx ? idx[o] : o;
    float base0 = 0.f;
    if (init && !accumulate_always && b != nullptr) base0 = ldw1<WS>(b + o);
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
      float* p = dst + row * TSP + tid + k * NT;
      float base = (init && !accumulate_always) ? base0 : *p;
      *p = base + acc[k];
    }
  }
}

// Conditioner MLP for this thread's SPT samples.  Inputs: act rows cond[0..d1).
// Outputs (if want_out): out[q][sample], q < d0.  Hidden post-ReLU activations of layers
// 0..m-2 (and m-1 when store_last) are written to hid (layer j at row offset sum Hp[<j]).
template <int SPT, bool WS>
__device__ __forceinline__ void net_forward(const CnfDims& d, const float* Wn, const float* act, const int* cond,
                                            float* out, float* hid, bool store_last, bool want_out, int TSP,
                                            int tid, int NT) {
  float h[SPT][CH];
  if (d.m == 0) {
    if (!want_out) return;
    for (int r0 = 0; r0 < d.d0p; r0 += CH) {
      chunk_from_inputs<SPT, WS>(h, Wn + d.w_off[0], d.d0p, Wn + d.b_off[0], r0, d.d1, act, cond, TSP, tid, NT);
#pragma unroll
      for (int r = 0; r < CH; ++r)
        if (r0 + r < d.d0) {
#pragma unroll
          for (int k = 0; k < SPT; ++k) out[(r0 + r) * TSP + tid + k * NT] = h[k][r];
        }
    }
    return;
  }
  int hoff = 0, hoff_prev = 0;
  for (int j = 0; j < d.m; ++j) {
    const bool last = (j == d.m - 1);
    const int n_in = (j == 0) ? d.d1 : d.Hp[j - 1];
    const float* src = (j == 0) ? act : hid + hoff_prev * TSP;
    const int* idx = (j == 0) ? cond : nullptr;
    for (int r0 = 0; r0 < d.Hp[j]; r0 += CH) {
      chunk_from_inputs<SPT, WS>(h, Wn + d.w_off[j], d.Hp[j], Wn + d.b_off[j], r0, n_in, src, idx, TSP, tid, NT);
#pragma unroll
      for (int k = 0; k < SPT; ++k)
#pragma unroll
        for (int r = 0; r < CH; ++r) h[k][r] = fmaxf(h[k][r], 0.f);
      if (!last || store_last) {
#pragma unroll
        for (int r = 0; r < CH; ++r)
#pragma unroll
          for (int k = 0; k < SPT; ++k) hid[(hoff + r0 + r) * TSP + tid + k * NT] = h[k][r];
      }
      if (last && want_out)
        chunk_to_outputs<SPT, WS>(h, Wn + d.w_off[d.m], d.Hp[j], Wn + d.b_off[d.m], r0, d.d0, out, nullptr,
                                  r0 == 0, false, TSP, tid, NT);
    }
    hoff_prev = hoff;
    hoff += d.Hp[j];
  }
}

// cooperative: copy n floats global -> shared
__device__ __forceinline__ void coop_copy(float* dst, const float* src, int n, int tid, int NT) {
  const int n4 = n / 4;
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(dst);
  for (int i = tid; i < n4; i += NT) d4[i] = __ldg(s4 + i);
  for (int i = n4 * 4 + tid; i < n; i += NT) dst[i] = __ldg(src + i);
}

// cooperative tile load: global [TS][K] (row-major) -> act[slot(f)][sample]
__device__ __forceinline__ void load_tile(float* act, const float* g, int64_t base, int64_t N, int K, int TS,
                                          int TSP, const int* slot_of, int tid, int NT) {
  const int total = TS * K;
  const float* gp = g + base * K;
  const int64_t avail = (N - base) * (int64_t)K;
  for (int e = tid; e < total; e += NT) {
    const int s = e / K, f = e - s * K;
    const float v = (e < avail) ? __ldg(gp + e) : 0.f;
    act[(slot_of ? slot_of[f] : f) * TSP + s] = v;
  }
}
__device__ __forceinline__ void store_tile(const float* act, float* g, int64_t base, int64_t N, int K, int TS,
                                           int TSP, const int* slot_of, int tid, int NT) {
  const int total = TS * K;
  float* gp = g + base * K;
  const int64_t avail = (N - base) * (int64_t)K;
  for (int e = tid; e < total; e += NT) {
    const int s = e / K, f = e - s * K;
    if (e < avail) gp[e] = act[(slot_of ? slot_of[f] : f) * TSP + s];
  }
}

// --------------------------------------------------------------------------------------
// forward / inverse
// --------------------------------------------------------------------------------------
// TAIL (0 / CNF_METRICS_LOGITS / CNF_METRICS_CALIBRATED): the fused Calibrator.predict tail and ECE / NLL /
// accuracy statistics taken from the finished tile while it is still in shared memory (calibrators.py:40-44,
// 350-353; utils/metrics.py:35-73, 6-15, 76-80), and the row-mean centring of the raw logits as a prologue
// (calibrators.py:17, 42).  zout / logdet may then be NULL: the pass reads 4K (+8) bytes per sample and writes
// only what was asked for.
template <int SPT, bool WS, int TAIL>
__global__ void __launch_bounds__(256, 2) flow_apply_kernel(CnfDims d, const float* __restrict__ packed, const int* __restrict__ tables,
                                  const float* __restrict__ xin, float* __restrict__ zout,
                                  float* __restrict__ logdet, float* __restrict__ zs, int64_t N, int inverse,
                                  CnfTail ta) {
  extern __shared__ __align__(16) float smem[];
  __shared__ double tail_red[32];
  const int NT = blockDim.x, tid = threadIdx.x;
  const int TS = NT * SPT, TSP = TS + 4;
  const Smem sm = make_smem(d, TSP, WS, false, TAIL ? ta.bins : 0);
  TailSmem tsm;
  BinCache cache; cache.bin = -1; cache.cnt = 0; cache.correct = 0; cache.sconf = 0.0;
  double a_nll = 0.0;
  unsigned a_correct = 0u, a_n = 0u;
  if (TAIL) {
    tsm = tail_carve(reinterpret_cast<unsigned char*>(smem + sm.tail), ta.bins, d.K);
    tail_init(tsm, ta, d.K, tid, NT);
  }
  int* tab = reinterpret_cast<int*>(smem + sm.tab);
  float* act = smem + sm.act;
  float* outs_s = smem + sm.outs;
  float* outs_t = outs_s + d.d0 * TSP;
  float* hid = smem + sm.hid;
  const float* W = WS ? smem + sm.w : packed;
  if (WS) coop_copy(smem + sm.w, packed, d.n_packed, tid, NT);
  for (int i = tid; i < d.n_tables; i += NT) tab[i] = tables[i];
  __syncthreads();
  const int* pi_last = tab + d.tab_pi + d.L * d.K;
  const int64_t ntiles = (N + TS - 1) / TS;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t base = tile * TS;
    load_tile(act, xin, base, N, d.K, TS, TSP, inverse ? pi_last : nullptr, tid, NT);
    __syncthreads();
    if (TAIL && ta.center) {   // forward only (host-checked): act row f holds logical column f
#pragma unroll
      for (int k = 0; k < SPT; ++k) {
        float* col = act + tid + k * NT;
        const float mean = numpy_row_mean([&](int j) -> float { return col[j * TSP]; }, d.K);
        for (int j = 0; j < d.K; ++j) col[j * TSP] = __fsub_rn(col[j * TSP], mean);
      }
    }
    float ld[SPT];
#pragma unroll
    for (int k = 0; k < SPT; ++k) ld[k] = 0.f;
    for (int li = 0; li < d.L; ++li) {
      const int l = inverse ? d.L - 1 - li : li;
      const int* cond = tab + d.tab_cond + l * d.d1;
      const int* trans = tab + d.tab_trans + l * d.d0;
      const float* Wl = W + (size_t)l * d.layer_stride;
      int slot = 0;
      if (d.nets & 1) { net_forward<SPT, WS>(d, Wl, act, cond, outs_s, hid, false, true, TSP, tid, NT); ++slot; }
      if (d.nets & 2)
        net_forward<SPT, WS>(d, Wl + (size_t)slot * d.net_stride, act, cond, outs_t, hid, false, true, TSP, tid, NT);
      for (int q = 0; q < d.d0; ++q) {
        const int p = trans[q];
#pragma unroll
        for (int k = 0; k < SPT; ++k) {
          const int s = tid + k * NT;
          const float xv = act[p * TSP + s];
          const float sv = (d.nets & 1) ? outs_s[q * TSP + s] : 0.f;
          const float tv = (d.nets & 2) ? outs_t[q * TSP + s] : 0.f;
          float yv;
          if (!inverse) { yv = xv * expf(sv) + tv; ld[k] += sv; }
          else          { yv = (xv - tv) * expf(-sv); ld[k] -= sv; }
          act[p * TSP + s] = yv;
        }
      }
      if (zs != nullptr) {
        // forward: zs[l] is the output of layer l in its logical order pi_{l+1};
        // inverse: xs[li] is the input of layer l in logical order pi_l.
        const int* pi = tab + d.tab_pi + (inverse ? l : l + 1) * d.K;
        float* dst = zs + (size_t)li * N * d.K;
#pragma unroll
        for (int k = 0; k < SPT; ++k) {
          const int s = tid + k * NT;
          const int64_t n = base + s;
          if (n < N)
            for (int j = 0; j < d.K; ++j) dst[n * d.K + j] = act[pi[j] * TSP + s];
        }
      }
    }
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
      const int64_t n = base + tid + k * NT;
      if (n < N && (!TAIL || logdet != nullptr)) logdet[n] = ld[k];
      if (TAIL && n < N) {   // the finished sample, logical class order (forward: z[j] = act[pi_L(j)])
        const float* col = act + tid + k * NT;
        auto get = [&](int jj) -> float { return col[(inverse ? jj : pi_last[jj]) * TSP]; };
        tail_row<TAIL == 0 ? CNF_METRICS_LOGITS : TAIL>(get, n, d.K, ta, tsm, cache, a_nll, a_correct, a_n);
      }
    }
    __syncthreads();
    if (!TAIL || zout != nullptr) store_tile(act, zout, base, N, d.K, TS, TSP, inverse ? nullptr : pi_last, tid, NT);
    __syncthreads();
  }
  if (TAIL) stats_finish_block(a_nll, a_correct, a_n, cache, tsm.s_cnt, tsm.s_cor, tsm.s_conf, ta.bins, ta.acc, tail_red, tid, NT);
}

// --------------------------------------------------------------------------------------
// training: forward + loss head + backward, one pass
// --------------------------------------------------------------------------------------
// G[a*ldg + b] += sum_s A[rowA(a)][s] * B[b][s]   (a < na, b < nb), atomics on global.
// One work item = one b and a block of 8 a's.  With fewer items than threads the sample loop of an item is
// split over nseg adjacent lanes (interleaved float4s, so a group still reads contiguous shared memory) and
// the partial sums are folded with shuffles before the atomics.
template <int NSEG>   // lanes per item, compile-time for the common cases (0: run-time nseg_rt)
__device__ __forceinline__ void wgrad_outer_seg(float* G, int ldg, const float* A, const int* idxA, int na,
                                                const float* B, int nb, float* Gbias_b, int TS, int TSP, int tid,
                                                int NT, int nseg_rt) {
  constexpr int AB = 8;
  const int nablk = (na + AB - 1) / AB;
  const int items = nb * nablk;
  const int ts4 = TS >> 2;
  const int nseg = NSEG ? NSEG : nseg_rt;
  const int seg = tid & (nseg - 1);
  const int groups = NT / nseg;
  for (int it0 = 0; it0 < items; it0 += groups) {      // block-uniform trip count: every lane reaches the shuffles
    const int it_raw = it0 + tid / nseg;
    const bool on = it_raw < items;
    const int it = on ? it_raw : items - 1;
    const int b = it % nb, ablk = it / nb;
    const int a0 = ablk * AB;
    const int an = min(AB, na - a0);
    float acc[AB];
#pragma unroll
    for (int a = 0; a < AB; ++a) acc[a] = 0.f;
    float bsum = 0.f;
    int rowA[AB];
#pragma unroll
    for (int a = 0; a < AB; ++a) {
      const int aa = a0 + min(a, an - 1);
      rowA[a] = (idxA ? idxA[aa] : aa) * TSP;
    }
    const float* Brow = B + b * TSP;
    for (int s4 = seg; s4 < ts4; s4 += nseg) {
      const int s = s4 << 2;
      const float4 bv = *reinterpret_cast<const float4*>(Brow + s);
      bsum += (bv.x + bv.y) + (bv.z + bv.w);
#pragma unroll
      for (int a = 0; a < AB; ++a) {
        const float4 av = *reinterpret_cast<const float4*>(A + rowA[a] + s);
        acc[a] = fmaf(av.x, bv.x, acc[a]);
        acc[a] = fmaf(av.y, bv.y, acc[a]);
        acc[a] = fmaf(av.z, bv.z, acc[a]);
        acc[a] = fmaf(av.w, bv.w, acc[a]);
      }
    }
    for (int o = nseg >> 1; o > 0; o >>= 1) {
#pragma unroll
      for (int a = 0; a < AB; ++a) acc[a] += __shfl_xor_sync(0xffffffffu, acc[a], o);
      bsum += __shfl_xor_sync(0xffffffffu, bsum, o);
    }
    if (on && seg == 0) {
#pragma unroll
      for (int a = 0; a < AB; ++a)
        if (a < an) atomicAdd(G + (size_t)(a0 + a) * ldg + b, acc[a]);
      if (Gbias_b != nullptr && ablk == 0) atomicAdd(Gbias_b + b, bsum);
    }
  }
}

__device__ __forceinline__ void wgrad_outer(float* G, int ldg, const float* A, const int* idxA, int na,
                                            const float* B, int nb, float* Gbias_b, int TS, int TSP, int tid,
                                            int NT) {
  const int items = nb * ((na + 7) / 8);
  const int ts4 = TS >> 2;
  int nseg = 1;
  while (nseg < 32 && items * nseg * 2 <= NT && (ts4 % (nseg * 2)) == 0) nseg <<= 1;
  if (nseg == 1) wgrad_outer_seg<1>(G, ldg, A, idxA, na, B, nb, Gbias_b, TS, TSP, tid, NT, 1);
  else if (nseg == 2) wgrad_outer_seg<2>(G, ldg, A, idxA, na, B, nb, Gbias_b, TS, TSP, tid, NT, 2);
  else wgrad_outer_seg<0>(G, ldg, A, idxA, na, B, nb, Gbias_b, TS, TSP, tid, NT, nseg);
}

// Gb[a] += sum_s A[a][s]
__device__ __forceinline__ void wgrad_rowsum(float* Gb, const float* A, int na, int TS, int TSP, int tid, int NT) {
  for (int a = tid; a < na; a += NT) {
    float acc = 0.f;
    const float* row = A + a * TSP;
    for (int s = 0; s < TS; s += 4) {
      const float4 v = *reinterpret_cast<const float4*>(row + s);
      acc += (v.x + v.y) + (v.z + v.w);
    }
    atomicAdd(Gb + a, acc);
  }
}

// Backward of one conditioner net for the tile.  On entry: hid holds every hidden layer's
// post-ReLU activations (from net_forward with store_last), gout[q][s] the gradient wrt the
// net's d0 outputs.  Adds input gradients to gact rows cond[], weight gradients to Gn
// (global, this net's block of the CTA's partial row).  Contains __syncthreads().
template <int SPT, bool WS>
__device__ __forceinline__ void net_backward(const CnfDims& d, const float* Wn, float* Gn, const float* act,
                                             const int* cond, const float* gout, float* hid, float* gbuf,
                                             float* gact, int TS, int TSP, int tid, int NT) {
  float g[SPT][CH];
  if (d.m == 0) {
    // thread: gact[cond[c]] += sum_q W[c][q] gout[q]
    for (int r0 = 0; r0 < d.d0p; r0 += CH) {
#pragma unroll
      for (int r = 0; r < CH; ++r)
#pragma unroll
        for (int k = 0; k < SPT; ++k)
          g[k][r] = (r0 + r < d.d0) ? gout[(r0 + r) * TSP + tid + k * NT] : 0.f;
      chunk_to_outputs<SPT, WS>(g, Wn + d.w_off[0], d.d0p, nullptr, r0, d.d1, gact, cond, false, true, TSP, tid, NT);
    }
    __syncthreads();
    wgrad_outer(Gn + d.w_off[0], d.d0p, act, cond, d.d1, gout, d.d0, Gn + d.b_off[0], TS, TSP, tid, NT);
    __syncthreads();
    return;
  }
  int hoff[CNF_MAX_HIDDEN];
  {
    int o = 0;
    for (int j = 0; j < d.m; ++j) { hoff[j] = o; o += d.Hp[j]; }
  }
  // last linear: dW[q][r] = sum_s gout[q][s] hid_{m-1}[r][s]; db[q] = sum_s gout[q][s]
  __syncthreads();
  wgrad_outer(Gn + d.w_off[d.m], d.Hp[d.m - 1], gout, nullptr, d.d0, hid + hoff[d.m - 1] * TSP, d.Hp[d.m - 1],
              nullptr, TS, TSP, tid, NT);
  wgrad_rowsum(Gn + d.b_off[d.m], gout, d.d0, TS, TSP, tid, NT);
  __syncthreads();
  for (int j = d.m - 1; j >= 0; --j) {
    float* hj = hid + hoff[j] * TSP;
    float* gin_cur = gbuf + ((j & 1) ? d.Hmax * TSP : 0);         // gradient wrt hid_j (pre-mask), j < m-1
    float* gin_prev = gbuf + (((j - 1) & 1) ? d.Hmax * TSP : 0);  // accumulates gradient wrt hid_{j-1}
    for (int r0 = 0; r0 < d.Hp[j]; r0 += CH) {
      if (j == d.m - 1) {
        chunk_from_inputs<SPT, WS>(g, Wn + d.w_off[d.m], d.Hp[j], nullptr, r0, d.d0, gout, nullptr, TSP, tid, NT);
      } else {
#pragma unroll
        for (int r = 0; r < CH; ++r)
#pragma unroll
          for (int k = 0; k < SPT; ++k) g[k][r] = gin_cur[(r0 + r) * TSP + tid + k * NT];
      }
      // ReLU mask, then overwrite the activation in place with the pre-activation gradient
#pragma unroll
      for (int r = 0; r < CH; ++r)
#pragma unroll
        for (int k = 0; k < SPT; ++k) {
          float* hp = hj + (r0 + r) * TSP + tid + k * NT;
          g[k][r] = (*hp > 0.f) ? g[k][r] : 0.f;
          *hp = g[k][r];
        }
      if (j == 0)
        chunk_to_outputs<SPT, WS>(g, Wn + d.w_off[0], d.Hp[0], nullptr, r0, d.d1, gact, cond, false, true, TSP, tid, NT);
      else
        chunk_to_outputs<SPT, WS>(g, Wn + d.w_off[j], d.Hp[j], nullptr, r0, d.Hp[j - 1], gin_prev, nullptr, r0 == 0,
                                  false, TSP, tid, NT);
    }
    __syncthreads();
    // dW_j[in][out] = sum_s in_j[in][s] gpre_j[out][s]; db_j[out] = sum_s gpre_j[out][s]
    if (j == 0)
      wgrad_outer(Gn + d.w_off[0], d.Hp[0], act, cond, d.d1, hj, d.Hp[0], Gn + d.b_off[0], TS, TSP, tid, NT);
    else
      wgrad_outer(Gn + d.w_off[j], d.Hp[j], hid + hoff[j - 1] * TSP, nullptr, d.Hp[j - 1], hj, d.Hp[j],
                  Gn + d.b_off[j], TS, TSP, tid, NT);
    __syncthreads();
  }
}

__device__ __forceinline__ double block_sum(double v, double* red, int tid, int NT) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (tid == 0)
    for (int w = 0; w < (NT + 31) / 32; ++w) t += red[w];
  return t;
}

template <int SPT, bool WS>
__global__ void flow_train_kernel(CnfDims d, const float* __restrict__ packed, const int* __restrict__ tables,
                                  const float* __restrict__ xin, const int64_t* __restrict__ labels,
                                  const float* __restrict__ gz_ext, const float* __restrict__ gld_ext,
                                  float* __restrict__ gx_out, float* __restrict__ partials,
                                  double* __restrict__ loss_acc, int64_t N, float eps, float gamma, float inv_n,
                                  int head, int wl) {
  extern __shared__ __align__(16) float smem[];
  __shared__ double red[32];
  const int NT = blockDim.x, tid = threadIdx.x;
  const int TS = NT * SPT, TSP = TS + 4;
  const bool do_bwd = (partials != nullptr);
  // wl (with WS): only the current coupling layer's weights are in shared memory, restaged at every layer
  // boundary; the room this frees goes into a wider tile (twice the resident warps at the C2 shape)
  const Smem sm = make_smem(d, TSP, WS ? (wl ? 2 : 1) : 0, true);
  int* tab = reinterpret_cast<int*>(smem + sm.tab);
  float* act = smem + sm.act;
  float* outs_s = smem + sm.outs;
  float* outs_t = outs_s + d.d0 * TSP;
  float* hid = smem + sm.hid;
  float* gact = smem + sm.gact;
  float* tape = smem + sm.tape;
  float* gout_s = smem + sm.gout;
  float* gout_t = gout_s + d.d0 * TSP;
  float* gbuf = smem + sm.gbuf;
  const float* W = WS ? smem + sm.w : packed;
  float* Grow = do_bwd ? partials + (size_t)(blockIdx.x % d.grad_rows) * d.n_packed : nullptr;
  if (WS && !wl) coop_copy(smem + sm.w, packed, d.n_packed, tid, NT);
  for (int i = tid; i < d.n_tables; i += NT) tab[i] = tables[i];
  __syncthreads();
  const int* pi_last = tab + d.tab_pi + d.L * d.K;
  const int64_t ntiles = (N + TS - 1) / TS;
  auto stage_layer = [&](int l) {
    __syncthreads();   // every thread is done with the layer staged before
    coop_copy(smem + sm.w, packed + (size_t)l * d.layer_stride, (int)d.layer_stride, tid, NT);
    __syncthreads();
  };
  double a_loss = 0.0, a_ce = 0.0, a_ld = 0.0, a_bad = 0.0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t base = tile * TS;
    load_tile(act, xin, base, N, d.K, TS, TSP, nullptr, tid, NT);
    if (head == CNF_HEAD_EXTERNAL) load_tile(gact, gz_ext, base, N, d.K, TS, TSP, pi_last, tid, NT);
    __syncthreads();
    float ld[SPT];
#pragma unroll
    for (int k = 0; k < SPT; ++k) ld[k] = 0.f;
    // ---- forward, recording the pre-layer value of every transformed slot -----------
    for (int l = 0; l < d.L; ++l) {
      const int* cond = tab + d.tab_cond + l * d.d1;
      const int* trans = tab + d.tab_trans + l * d.d0;
      if (WS && wl) stage_layer(l);
      const float* Wl = (WS && wl) ? W : W + (size_t)l * d.layer_stride;
      int slot = 0;
      if (d.nets & 1) { net_forward<SPT, WS>(d, Wl, act, cond, outs_s, hid, false, true, TSP, tid, NT); ++slot; }
      if (d.nets & 2)
        net_forward<SPT, WS>(d, Wl + (size_t)slot * d.net_stride, act, cond, outs_t, hid, false, true, TSP, tid, NT);
      for (int q = 0; q < d.d0; ++q) {
        const int p = trans[q];
#pragma unroll
        for (int k = 0; k < SPT; ++k) {
          const int s = tid + k * NT;
          const float xv = act[p * TSP + s];
          const float sv = (d.nets & 1) ? outs_s[q * TSP + s] : 0.f;
          const float tv = (d.nets & 2) ? outs_t[q * TSP + s] : 0.f;
          tape[(l * d.d0 + q) * TSP + s] = xv;
          act[p * TSP + s] = xv * expf(sv) + tv;
          ld[k] += sv;
        }
      }
    }
    // ---- loss head ------------------------------------------------------------------
    float gld[SPT];
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
      const int s = tid + k * NT;
      const int64_t n = base + s;
      const bool valid = n < N;
      gld[k] = 0.f;
      if (head == CNF_HEAD_NLL) {
        float mx = -INFINITY;
        for (int j = 0; j < d.K; ++j) mx = fmaxf(mx, act[pi_last[j] * TSP + s]);
        float se = 0.f;
        for (int j = 0; j < d.K; ++j) se += expf(act[pi_last[j] * TSP + s] - mx);
        int yy = valid ? (int)labels[n] : 0;
        yy = min(max(yy, 0), d.K - 1);  // out-of-range labels are clamped, never read out of bounds
        const float zy = act[pi_last[yy] * TSP + s];
        const float inv_se = 1.f / se;
        const float py = expf(zy - mx) * inv_se;
        float ce, coef;
        if (eps == 0.f) { ce = (zy - mx) - logf(se); coef = 1.f; }
        else            { ce = logf(py + eps); coef = py / (py + eps); }
        if (valid) {
          const float tot = ce + gamma * ld[k];
          a_loss += (double)tot; a_ce += (double)ce; a_ld += (double)ld[k];
          if (!isfinite(tot)) a_bad += 1.0;
        }
        if (do_bwd) {
          const float sc = valid ? -inv_n * coef : 0.f;
          for (int j = 0; j < d.K; ++j) {
            const int p = pi_last[j];
            const float pj = expf(act[p * TSP + s] - mx) * inv_se;
            gact[p * TSP + s] = sc * ((j == yy ? 1.f : 0.f) - pj);
          }
          gld[k] = valid ? -gamma * inv_n : 0.f;
        }
      } else {
        gld[k] = valid ? gld_ext[n] : 0.f;
      }
    }
    // ---- backward -------------------------------------------------------------------
    if (do_bwd) {
      for (int l = d.L - 1; l >= 0; --l) {
        const int* cond = tab + d.tab_cond + l * d.d1;
        const int* trans = tab + d.tab_trans + l * d.d0;
        if (WS && wl && l != d.L - 1) stage_layer(l);   // the forward pass left layer L-1 staged
        const float* Wl = (WS && wl) ? W : W + (size_t)l * d.layer_stride;
        float* Gl = Grow + (size_t)l * d.layer_stride;
        int slot = 0;
        if (d.nets & 1) {
          net_forward<SPT, WS>(d, Wl, act, cond, outs_s, hid, true, true, TSP, tid, NT);
          for (int q = 0; q < d.d0; ++q) {
            const int p = trans[q];
#pragma unroll
            for (int k = 0; k < SPT; ++k) {
              const int s = tid + k * NT;
              const float gy = gact[p * TSP + s];
              const float es = expf(outs_s[q * TSP + s]);
              const float xv = tape[(l * d.d0 + q) * TSP + s];
              gout_s[q * TSP + s] = gy * xv * es + gld[k];
              gout_t[q * TSP + s] = gy;
              gact[p * TSP + s] = gy * es;
            }
          }
          net_backward<SPT, WS>(d, Wl, Gl, act, cond, gout_s, hid, gbuf, gact, TS, TSP, tid, NT);
          ++slot;
        } else if (d.nets & 2) {
          for (int q = 0; q < d.d0; ++q) {
            const int p = trans[q];
#pragma unroll
            for (int k = 0; k < SPT; ++k) gout_t[q * TSP + tid + k * NT] = gact[p * TSP + tid + k * NT];
          }
        }
        if (d.nets & 2) {
          net_forward<SPT, WS>(d, Wl + (size_t)slot * d.net_stride, act, cond, outs_t, hid, true, false, TSP, tid, NT);
          net_backward<SPT, WS>(d, Wl + (size_t)slot * d.net_stride, Gl + (size_t)slot * d.net_stride, act, cond,
                                gout_t, hid, gbuf, gact, TS, TSP, tid, NT);
        }
        // step the tile state back to the input of layer l
        for (int q = 0; q < d.d0; ++q) {
          const int p = trans[q];
#pragma unroll
          for (int k = 0; k < SPT; ++k) act[p * TSP + tid + k * NT] = tape[(l * d.d0 + q) * TSP + tid + k * NT];
        }
      }
      if (gx_out != nullptr) {
        __syncthreads();
        store_tile(gact, gx_out, base, N, d.K, TS, TSP, nullptr, tid, NT);
      }
    }
    __syncthreads();
  }
  if (loss_acc != nullptr && head == CNF_HEAD_NLL) {
    double t0 = block_sum(a_loss, red, tid, NT);
    double t1 = block_sum(a_ce, red, tid, NT);
    double t2 = block_sum(a_ld, red, tid, NT);
    double t3 = block_sum(a_bad, red, tid, NT);
    if (tid == 0) {
      atomicAdd(loss_acc + 0, t0);
      atomicAdd(loss_acc + 1, t1);
      atomicAdd(loss_acc + 2, t2);
      atomicAdd(loss_acc + 3, t3);
    }
  }
}

// --------------------------------------------------------------------------------------
// "Lean" training kernel for single-hidden-layer nets whose full tile plan does not fit shared memory
// (wide K and/or wide hidden layers, e.g. K = 100, H = 512): flow_train_kernel keeps a tape of every
// layer's pre-layer values (L*d0*TS floats) and a whole net's hidden activations (Hp*TS floats), which
// at that shape leaves room for 32-sample tiles only -- one warp per SM.  Here
//   * there is no tape: going backwards, a layer's input is recovered by inverting the layer,
//     x = (y - t) * exp(-s), with s and t recomputed from the (unchanged) conditioning logits;
//   * the hidden layer is processed in chunks of 16 units: recompute h -> 16 x TS shared-memory slab ->
//     cooperative weight-gradient sums -> per-thread pre-activation gradient -> slab -> sums;
//   * each chunk's weights (first-Linear columns, bias, last-Linear rows: (d1+1+d0) x 16 floats) are
//     staged cooperatively in shared memory, so the per-thread FMA loops read them with broadcast
//     LDS.128 instead of going to L2 through a nearly empty L1.
// Same arithmetic per element as flow_train_kernel; the recovered layer inputs differ from taped ones by
// fp32 rounding of the inversion (~1e-7 relative).
// --------------------------------------------------------------------------------------
struct LeanSmem { int tab, act, gact, outs_s, outs_t, gout_s, hc, wb, b1, total; };

__host__ __device__ inline LeanSmem make_lean(const CnfDims& d, int TSP) {
  LeanSmem s;
  int off = 0;
  s.tab = off; off += (d.n_tables + 3) / 4 * 4;
  s.act = off; off += d.K * TSP;
  s.gact = off; off += d.K * TSP;
  s.outs_s = off; off += d.d0 * TSP;
  s.outs_t = off; off += d.d0 * TSP;      // doubles as the t-net's output gradient in the backward pass
  s.gout_s = off; off += d.d0 * TSP;
  s.hc = off; off += CH * TSP;
  s.wb = off; off += 2 * (d.d1 + 1 + d.d0) * CH;      // two chunk buffers: the next chunk lands while this one is used
  s.b1 = off; off += (d.d0 + 3) / 4 * 4;
  s.total = off;
  return s;
}

// wb <- [W0[c][r0..r0+16) for c < d1 | b0[r0..) | W1[q][r0..) for q < d0] of one net (m == 1)
__device__ __forceinline__ void stage_chunk(float* wb, const CnfDims& d, const float* __restrict__ Wn, int r0, int tid,
                                            int NT) {
  const int Hp = d.Hp[0];
  const int n4 = (d.d1 + 1 + d.d0) * (CH / 4);
  for (int i = tid; i < n4; i += NT) {
    const int row = i / (CH / 4), c4 = i - row * (CH / 4);
    const float* src = row < d.d1 ? Wn + d.w_off[0] + (size_t)row * Hp + r0
                       : row == d.d1 ? Wn + d.b_off[0] + r0
                                     : Wn + d.w_off[1] + (size_t)(row - d.d1 - 1) * Hp + r0;
    reinterpret_cast<float4*>(wb)[i] = __ldg(reinterpret_cast<const float4*>(src) + c4);
  }
}

// The same copy split in two so that the global loads of the NEXT chunk fly while the current chunk is being
// computed: fetch into registers early, store to shared memory after the barrier that frees the buffer.
constexpr int LEAN_PF = 8;       // float4 per thread; covers (d1+1+d0)*4 <= 8*NT, i.e. K < 2*NT
__device__ __forceinline__ const float* chunk_src(const CnfDims& d, const float* __restrict__ Wn, int r0, int row) {
  const int Hp = d.Hp[0];
  return row < d.d1 ? Wn + d.w_off[0] + (size_t)row * Hp + r0
         : row == d.d1 ? Wn + d.b_off[0] + r0
                       : Wn + d.w_off[1] + (size_t)(row - d.d1 - 1) * Hp + r0;
}
__device__ __forceinline__ void fetch_chunk(float4 (&pf)[LEAN_PF], const CnfDims& d, const float* __restrict__ Wn,
                                            int r0, int tid, int NT) {
  const int n4 = (d.d1 + 1 + d.d0) * (CH / 4);
#pragma unroll
  for (int j = 0; j < LEAN_PF; ++j) {
    const int i = tid + j * NT;
    if (i < n4) {
      const int row = i / (CH / 4), c4 = i - row * (CH / 4);
      pf[j] = __ldg(reinterpret_cast<const float4*>(chunk_src(d, Wn, r0, row)) + c4);
    }
  }
}
__device__ __forceinline__ void store_chunk(float* wb, const float4 (&pf)[LEAN_PF], const CnfDims& d, int tid, int NT) {
  const int n4 = (d.d1 + 1 + d.d0) * (CH / 4);
#pragma unroll
  for (int j = 0; j < LEAN_PF; ++j) {
    const int i = tid + j * NT;
    if (i < n4) reinterpret_cast<float4*>(wb)[i] = pf[j];
  }
}

// cp.async variant: no registers held across the chunk's compute (the training kernel has none to spare)
__device__ __forceinline__ void stage_chunk_async(float* wb, const CnfDims& d, const float* __restrict__ Wn, int r0,
                                                  int tid, int NT) {
  const int n4 = (d.d1 + 1 + d.d0) * (CH / 4);
  for (int i = tid; i < n4; i += NT) {
    const int row = i / (CH / 4), c4 = i - row * (CH / 4);
    const unsigned dst = (unsigned)__cvta_generic_to_shared(reinterpret_cast<float4*>(wb) + i);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(reinterpret_cast<const float4*>(chunk_src(d, Wn, r0, row)) + c4) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void stage_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// out[q][s] = net(act[cond])[q] for this thread's samples; block-cooperative (contains __syncthreads)
template <int SPT, bool PRE>
__device__ __forceinline__ void net_forward_lean(const CnfDims& d, const float* Wn, const float* act, const int* cond,
                                                 float* out, float* wb, float* b1s, int TSP, int tid, int NT) {
  const int Hp = d.Hp[0];
  __syncthreads();
  for (int q = tid; q < d.d0; q += NT) b1s[q] = __ldg(Wn + d.b_off[1] + q);
  // PRE: register prefetch of the next chunk (forward kernel only: in the training kernel the 32 extra
  // registers push the backward into spills, 126 vs 76 ms per step)
  const bool pre = PRE && (d.d1 + 1 + d.d0) * (CH / 4) <= LEAN_PF * NT;
  float4 pf[LEAN_PF];
  if (pre) fetch_chunk(pf, d, Wn, 0, tid, NT);
  const int wbsz = (d.d1 + 1 + d.d0) * CH;
  float* wcur = wb;
  if (!PRE) stage_chunk_async(wb, d, Wn, 0, tid, NT);
  for (int r0 = 0; r0 < Hp; r0 += CH) {
    if (PRE) {
      if (r0) __syncthreads();          // every thread is done with the previous chunk's weights
      if (pre) {
        store_chunk(wb, pf, d, tid, NT);
        if (r0 + CH < Hp) fetch_chunk(pf, d, Wn, r0 + CH, tid, NT);     // in flight during this chunk's FMAs
      } else {
        stage_chunk(wb, d, Wn, r0, tid, NT);
      }
      __syncthreads();
    } else {                            // two buffers: chunk r0 has landed, chunk r0+16 starts flying
      stage_wait();
      __syncthreads();
      if (r0 + CH < Hp) stage_chunk_async(wcur == wb ? wb + wbsz : wb, d, Wn, r0 + CH, tid, NT);
    }
    float h[SPT][CH];
    chunk_from_inputs<SPT, true>(h, wcur, CH, wcur + d.d1 * CH, 0, d.d1, act, cond, TSP, tid, NT);
#pragma unroll
    for (int k = 0; k < SPT; ++k)
#pragma unroll
      for (int r = 0; r < CH; ++r) h[k][r] = fmaxf(h[k][r], 0.f);
    chunk_to_outputs<SPT, true>(h, wcur + (d.d1 + 1) * CH, CH, b1s, 0, d.d0, out, nullptr, r0 == 0, false, TSP, tid, NT);
    if (!PRE) wcur = (wcur == wb) ? wb + wbsz : wb;
  }
}

// Backward of one net for the tile, chunk by chunk.  gout[q][s]: gradient on the net's outputs.
template <int SPT>
__device__ __forceinline__ void net_backward_lean(const CnfDims& d, const float* Wn, float* Gn, const float* act,
                                                  const int* cond, const float* gout, float* hc, float* wb,
                                                  float* gact, int TS, int TSP, int tid, int NT) {
  const int Hp = d.Hp[0];
  __syncthreads();                                    // gout is complete
  wgrad_rowsum(Gn + d.b_off[1], gout, d.d0, TS, TSP, tid, NT);
  const int wbsz = (d.d1 + 1 + d.d0) * CH;
  float* wcur = wb;
  stage_chunk_async(wb, d, Wn, 0, tid, NT);
  for (int r0 = 0; r0 < Hp; r0 += CH) {
    stage_wait();
    __syncthreads();                                  // chunk r0 has landed; previous chunk's slab and weights are free
    if (r0 + CH < Hp) stage_chunk_async(wcur == wb ? wb + wbsz : wb, d, Wn, r0 + CH, tid, NT);
    float* const wbc = wcur;
    wcur = (wcur == wb) ? wb + wbsz : wb;
    float h[SPT][CH], g[SPT][CH];
    chunk_from_inputs<SPT, true>(h, wbc, CH, wbc + d.d1 * CH, 0, d.d1, act, cond, TSP, tid, NT);
#pragma unroll
    for (int k = 0; k < SPT; ++k)
#pragma unroll
      for (int r = 0; r < CH; ++r) {
        h[k][r] = fmaxf(h[k][r], 0.f);
        hc[r * TSP + tid + k * NT] = h[k][r];
      }
    __syncthreads();
    // dW1[q][r0+r] += sum_s gout[q][s] h[r][s]
    wgrad_outer(Gn + d.w_off[1] + r0, Hp, gout, nullptr, d.d0, hc, CH, nullptr, TS, TSP, tid, NT);
    // g[r] = sum_q W1[q][r0+r] gout[q], masked by the ReLU
    chunk_from_inputs<SPT, true>(g, wbc + (d.d1 + 1) * CH, CH, nullptr, 0, d.d0, gout, nullptr, TSP, tid, NT);
    __syncthreads();                                  // the weight-gradient pass has read the slab
#pragma unroll
    for (int k = 0; k < SPT; ++k)
#pragma unroll
      for (int r = 0; r < CH; ++r) {
        g[k][r] = (h[k][r] > 0.f) ? g[k][r] : 0.f;
        hc[r * TSP + tid + k * NT] = g[k][r];
      }
    // gact[cond[c]] += sum_r W0[c][r0+r] g[r]   (per thread, own samples)
    chunk_to_outputs<SPT, true>(g, wbc, CH, nullptr, 0, d.d1, gact, cond, false, true, TSP, tid, NT);
    __syncthreads();
    // dW0[c][r0+r] += sum_s act[cond[c]][s] g[r][s]; db0[r0+r] += sum_s g[r][s]
    wgrad_outer(Gn + d.w_off[0] + r0, Hp, act, cond, d.d1, hc, CH, Gn + d.b_off[0] + r0, TS, TSP, tid, NT);
  }
  __syncthreads();
}

template <int SPT>
__global__ void flow_train_lean_kernel(CnfDims d, const float* __restrict__ packed, const int* __restrict__ tables,
                                       const float* __restrict__ xin, const int64_t* __restrict__ labels,
                                       const float* __restrict__ gz_ext, const float* __restrict__ gld_ext,
                                       float* __restrict__ gx_out, float* __restrict__ partials,
                                       double* __restrict__ loss_acc, int64_t N, float eps, float gamma, float inv_n,
                                       int head) {
  extern __shared__ __align__(16) float smem[];
  __shared__ double red[32];
  const int NT = blockDim.x, tid = threadIdx.x;
  const int TS = NT * SPT, TSP = TS + 4;
  const bool do_bwd = (partials != nullptr);
  const LeanSmem sm = make_lean(d, TSP);
  int* tab = reinterpret_cast<int*>(smem + sm.tab);
  float* act = smem + sm.act;
  float* gact = smem + sm.gact;
  float* outs_s = smem + sm.outs_s;
  float* outs_t = smem + sm.outs_t;
  float* gout_s = smem + sm.gout_s;
  float* gout_t = outs_t;
  float* hc = smem + sm.hc;
  float* wb = smem + sm.wb;
  float* b1s = smem + sm.b1;
  float* Grow = do_bwd ? partials + (size_t)(blockIdx.x % d.grad_rows) * d.n_packed : nullptr;
  for (int i = tid; i < d.n_tables; i += NT) tab[i] = tables[i];
  __syncthreads();
  const int* pi_last = tab + d.tab_pi + d.L * d.K;
  const int64_t ntiles = (N + TS - 1) / TS;
  double a_loss = 0.0, a_ce = 0.0, a_ld = 0.0, a_bad = 0.0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t base = tile * TS;
    __syncthreads();
    load_tile(act, xin, base, N, d.K, TS, TSP, nullptr, tid, NT);
    if (head == CNF_HEAD_EXTERNAL) load_tile(gact, gz_ext, base, N, d.K, TS, TSP, pi_last, tid, NT);
    __syncthreads();
    float ld[SPT];
#pragma unroll
    for (int k = 0; k < SPT; ++k) ld[k] = 0.f;
    // ---- forward ----------------------------------------------------------------------
    for (int l = 0; l < d.L; ++l) {
      const int* cond = tab + d.tab_cond + l * d.d1;
      const int* trans = tab + d.tab_trans + l * d.d0;
      const float* Wl = packed + (size_t)l * d.layer_stride;
      int slot = 0;
      if (d.nets & 1) { net_forward_lean<SPT, false>(d, Wl, act, cond, outs_s, wb, b1s, TSP, tid, NT); ++slot; }
      if (d.nets & 2) net_forward_lean<SPT, false>(d, Wl + (size_t)slot * d.net_stride, act, cond, outs_t, wb, b1s, TSP, tid, NT);
      for (int q = 0; q < d.d0; ++q) {
        const int p = trans[q];
#pragma unroll
        for (int k = 0; k < SPT; ++k) {
          const int s = tid + k * NT;
          const float sv = (d.nets & 1) ? outs_s[q * TSP + s] : 0.f;
          const float tv = (d.nets & 2) ? outs_t[q * TSP + s] : 0.f;
          act[p * TSP + s] = act[p * TSP + s] * expf(sv) + tv;
          ld[k] += sv;
        }
      }
    }
    // ---- loss head (same as flow_train_kernel) ---------------------------------------------
    float gld[SPT];
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
      const int s = tid + k * NT;
      const int64_t n = base + s;
      const bool valid = n < N;
      gld[k] = 0.f;
      if (head == CNF_HEAD_NLL) {
        float mx = -INFINITY;
        for (int j = 0; j < d.K; ++j) mx = fmaxf(mx, act[pi_last[j] * TSP + s]);
        float se = 0.f;
        for (int j = 0; j < d.K; ++j) se += expf(act[pi_last[j] * TSP + s] - mx);
        int yy = valid ? (int)labels[n] : 0;
        yy = min(max(yy, 0), d.K - 1);
        const float zy = act[pi_last[yy] * TSP + s];
        const float inv_se = 1.f / se;
        const float py = expf(zy - mx) * inv_se;
        float ce, coef;
        if (eps == 0.f) { ce = (zy - mx) - logf(se); coef = 1.f; }
        else            { ce = logf(py + eps); coef = py / (py + eps); }
        if (valid) {
          const float tot = ce + gamma * ld[k];
          a_loss += (double)tot; a_ce += (double)ce; a_ld += (double)ld[k];
          if (!isfinite(tot)) a_bad += 1.0;
        }
        if (do_bwd) {
          const float sc = valid ? -inv_n * coef : 0.f;
          for (int j = 0; j < d.K; ++j) {
            const int p = pi_last[j];
            const float pj = expf(act[p * TSP + s] - mx) * inv_se;
            gact[p * TSP + s] = sc * ((j == yy ? 1.f : 0.f) - pj);
          }
          gld[k] = valid ? -gamma * inv_n : 0.f;
        }
      } else {
        gld[k] = valid ? gld_ext[n] : 0.f;
      }
    }
    // ---- backward, recovering each layer's input by inverting the layer -------------------------
    if (do_bwd) {
      for (int l = d.L - 1; l >= 0; --l) {
        const int* cond = tab + d.tab_cond + l * d.d1;
        const int* trans = tab + d.tab_trans + l * d.d0;
        const float* Wl = packed + (size_t)l * d.layer_stride;
        float* Gl = Grow + (size_t)l * d.layer_stride;
        int slot = 0;
        if (d.nets & 1) { net_forward_lean<SPT, false>(d, Wl, act, cond, outs_s, wb, b1s, TSP, tid, NT); ++slot; }
        if (d.nets & 2) net_forward_lean<SPT, false>(d, Wl + (size_t)slot * d.net_stride, act, cond, outs_t, wb, b1s, TSP, tid, NT);
        for (int q = 0; q < d.d0; ++q) {
          const int p = trans[q];
#pragma unroll
          for (int k = 0; k < SPT; ++k) {
            const int s = tid + k * NT;
            const float sv = (d.nets & 1) ? outs_s[q * TSP + s] : 0.f;
            const float tv = (d.nets & 2) ? outs_t[q * TSP + s] : 0.f;
            const float es = expf(sv);
            const float xv = (act[p * TSP + s] - tv) * expf(-sv);
            const float gy = gact[p * TSP + s];
            act[p * TSP + s] = xv;                       // the input of layer l
            gout_s[q * TSP + s] = gy * xv * es + gld[k];
            gout_t[q * TSP + s] = gy;                    // overwrites t (already consumed)
            gact[p * TSP + s] = gy * es;
          }
        }
        slot = 0;
        if (d.nets & 1) {
          net_backward_lean<SPT>(d, Wl, Gl, act, cond, gout_s, hc, wb, gact, TS, TSP, tid, NT);
          ++slot;
        }
        if (d.nets & 2)
          net_backward_lean<SPT>(d, Wl + (size_t)slot * d.net_stride, Gl + (size_t)slot * d.net_stride, act, cond, gout_t,
                                 hc, wb, gact, TS, TSP, tid, NT);
      }
      if (gx_out != nullptr) {
        __syncthreads();
        store_tile(gact, gx_out, base, N, d.K, TS, TSP, nullptr, tid, NT);
      }
    }
  }
  if (loss_acc != nullptr && head == CNF_HEAD_NLL) {
    double t0 = block_sum(a_loss, red, tid, NT);
    double t1 = block_sum(a_ce, red, tid, NT);
    double t2 = block_sum(a_ld, red, tid, NT);
    double t3 = block_sum(a_bad, red, tid, NT);
    if (tid == 0) {
      atomicAdd(loss_acc + 0, t0);
      atomicAdd(loss_acc + 1, t1);
      atomicAdd(loss_acc + 2, t2);
      atomicAdd(loss_acc + 3, t3);
    }
  }
}

// Forward / inverse counterpart of the lean training kernel: same staged 16-unit weight chunks, no
// hidden-activation storage.  Used for single-hidden-layer nets whose weights do not fit shared memory.
template <int SPT>
__global__ void flow_apply_lean_kernel(CnfDims d, const float* __restrict__ packed, const int* __restrict__ tables,
                                       const float* __restrict__ xin, float* __restrict__ zout,
                                       float* __restrict__ logdet, float* __restrict__ zs, int64_t N, int inverse) {
  extern __shared__ __align__(16) float smem[];
  const int NT = blockDim.x, tid = threadIdx.x;
  const int TS = NT * SPT, TSP = TS + 4;
  int off = 0;
  int* tab = reinterpret_cast<int*>(smem + off); off += (d.n_tables + 3) / 4 * 4;
  float* act = smem + off; off += d.K * TSP;
  float* outs_s = smem + off; off += d.d0 * TSP;
  float* outs_t = smem + off; off += d.d0 * TSP;
  float* wb = smem + off; off += (d.d1 + 1 + d.d0) * CH;
  float* b1s = smem + off;
  for (int i = tid; i < d.n_tables; i += NT) tab[i] = tables[i];
  __syncthreads();
  const int* pi_last = tab + d.tab_pi + d.L * d.K;
  const int64_t ntiles = (N + TS - 1) / TS;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t base = tile * TS;
    __syncthreads();
    load_tile(act, xin, base, N, d.K, TS, TSP, inverse ? pi_last : nullptr, tid, NT);
    __syncthreads();
    float ld[SPT];
#pragma unroll
    for (int k = 0; k < SPT; ++k) ld[k] = 0.f;
    for (int li = 0; li < d.L; ++li) {
      const int l = inverse ? d.L - 1 - li : li;
      const int* cond = tab + d.tab_cond + l * d.d1;
      const int* trans = tab + d.tab_trans + l * d.d0;
      const float* Wl = packed + (size_t)l * d.layer_stride;
      int slot = 0;
      if (d.nets & 1) { net_forward_lean<SPT, true>(d, Wl, act, cond, outs_s, wb, b1s, TSP, tid, NT); ++slot; }
      if (d.nets & 2) net_forward_lean<SPT, true>(d, Wl + (size_t)slot * d.net_stride, act, cond, outs_t, wb, b1s, TSP, tid, NT);
      for (int q = 0; q < d.d0; ++q) {
        const int p = trans[q];
#pragma unroll
        for (int k = 0; k < SPT; ++k) {
          const int s = tid + k * NT;
          const float xv = act[p * TSP + s];
          const float sv = (d.nets & 1) ? outs_s[q * TSP + s] : 0.f;
          const float tv = (d.nets & 2) ? outs_t[q * TSP + s] : 0.f;
          float yv;
          if (!inverse) { yv = xv * expf(sv) + tv; ld[k] += sv; }
          else          { yv = (xv - tv) * expf(-sv); ld[k] -= sv; }
          act[p * TSP + s] = yv;
        }
      }
      if (zs != nullptr) {
        const int* pi = tab + d.tab_pi + (inverse ? l : l + 1) * d.K;
        float* dst = zs + (size_t)li * N * d.K;
#pragma unroll
        for (int k = 0; k < SPT; ++k) {
          const int s = tid + k * NT;
          const int64_t n = base + s;
          if (n < N)
            for (int j = 0; j < d.K; ++j) dst[n * d.K + j] = act[pi[j] * TSP + s];
        }
      }
    }
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
      const int64_t n = base + tid + k * NT;
      if (n < N) logdet[n] = ld[k];
    }
    __syncthreads();
    store_tile(act, zout, base, N, d.K, TS, TSP, inverse ? nullptr : pi_last, tid, NT);
  }
}

// --------------------------------------------------------------------------------------
// Small-batch training kernel (one hidden layer, Hp <= 256).
// The reference trains full-batch on calibration sets of a few thousand samples
// (calibrators.py:267-295); with one thread per sample those leave most of the GPU idle and a step
// costs one thread's serial walk through the whole stack (~300 us).  Here a CTA owns a 32-sample
// tile (lane = sample) and its warps split the hidden layer: warp w owns hidden units
// [16w, 16w+16) of both conditioners of every coupling layer.
//   forward:   h = relu(W0[:, chunk]^T x + b0) in registers, partial outputs W1[:, chunk] h to shared
//              memory, barrier, the warps sum the partials per output and apply the coupling;
//   backward:  output gradients from the taped x and s, the hidden units recomputed (h stays in registers),
//              then g = relu'(h) * W1[:, chunk]^T gout,
//              partial input gradients to shared memory, and the chunk's weight gradients with lanes =
//              matrix entries (h and g pass through a per-warp slab; an entry is one 32-sample dot product);
//              every (layer, net, chunk) block of the CTA's private partial row has exactly one owner
//              warp, so the first tile stores and later tiles use reductions without contention.
// One coupling layer's weights are staged per step with cp.async into a double buffer.
// Same arithmetic per element as flow_train_kernel; sums over samples and hidden units associate
// differently (fp32 rounding only).
// --------------------------------------------------------------------------------------
constexpr int SPL_TS = 32;

struct SplitSmem { int tab, act, gact, tape, stape, part, gout, gld, ldp, slab, w, total; };

// NW: warps of the CTA, NWn: warps per conditioner (= 16-unit chunks of the hidden layer)
__host__ __device__ inline SplitSmem make_split(const CnfDims& d, int NW, int NWn) {
  SplitSmem s;
  const int dpart = d.d0 > d.d1 ? d.d0 : d.d1;
  int off = 0;
  s.tab = off; off += (d.n_tables + 3) / 4 * 4;
  s.act = off; off += d.K * SPL_TS;
  s.gact = off; off += d.K * SPL_TS;
  s.tape = off; off += d.L * d.d0 * SPL_TS;         // pre-layer values of the transformed logits
  s.stape = off; off += d.L * d.d0 * SPL_TS;        // scale-net outputs of the forward pass
  s.part = off; off += 2 * NWn * dpart * SPL_TS;    // forward: [net][chunk][q]; backward: [warp][c]
  s.gout = off; off += 2 * d.d0 * SPL_TS;
  s.gld = off; off += SPL_TS;
  s.ldp = off; off += NW * SPL_TS;
  s.slab = off; off += NW * 2 * CH * (SPL_TS + 4);   // per warp: h and g of its chunk, [16][36] each
  s.w = off; off += 2 * d.layer_stride;
  s.total = off;
  return s;
}

__device__ __forceinline__ void stage_layer_async(float* dst, const float* __restrict__ src, int n, int tid, int NT) {
  const int n4 = n >> 2;     // layer_stride is a multiple of 4 floats
  for (int i = tid; i < n4; i += NT) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(reinterpret_cast<float4*>(dst) + i);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a), "l"(reinterpret_cast<const float4*>(src) + i) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// h[r] = relu(b0[r0+r] + sum_i W0[i][r0+r] act[cond[i]][lane])
__device__ __forceinline__ void split_hidden(float (&h)[1][CH], const CnfDims& d, int d1, const float* Wn, int r0,
                                             const float* act, const int* cond, int lane) {
  chunk_from_inputs<1, true>(h, Wn + d.w_off[0], d.Hp[0], Wn + d.b_off[0], r0, d1, act, cond, SPL_TS, lane, 0);
#pragma unroll
  for (int r = 0; r < CH; ++r) h[0][r] = fmaxf(h[0][r], 0.f);
}

// dst[q][lane] = sum_r W1[q][r0+r] h[r]   (this warp's share of the net's d0 outputs)
__device__ __forceinline__ void split_partial_out(float* dst, const float (&h)[1][CH], const CnfDims& d, int d0,
                                                  const float* Wn, int r0, int lane) {
  const float* W1 = Wn + d.w_off[1] + r0;
#pragma unroll 5
  for (int q = 0; q < d0; ++q) {
    const float* wrow = W1 + (size_t)q * d.Hp[0];
    float acc = 0.f;
#pragma unroll
    for (int r4 = 0; r4 < CH / 4; ++r4) {
      const float4 wv = *reinterpret_cast<const float4*>(wrow + 4 * r4);
      acc = fmaf(wv.x, h[0][4 * r4 + 0], acc);
      acc = fmaf(wv.y, h[0][4 * r4 + 1], acc);
      acc = fmaf(wv.z, h[0][4 * r4 + 2], acc);
      acc = fmaf(wv.w, h[0][4 * r4 + 3], acc);
    }
    dst[q * SPL_TS + lane] = acc;
  }
}

constexpr int SPL_SLP = SPL_TS + 4;   // row stride of the per-warp slabs: 128-bit row reads by 16 lanes, 2-way at most

__device__ __forceinline__ float row_sum32(const float* a) {
  float acc = 0.f;
#pragma unroll
  for (int s4 = 0; s4 < SPL_TS / 4; ++s4) {
    const float4 av = *reinterpret_cast<const float4*>(a + 4 * s4);
    acc += (av.x + av.y) + (av.z + av.w);
  }
  return acc;
}
__device__ __forceinline__ void grad_put(float* p, float v, bool first) {
  if (first) *p = v; else atomicAdd(p, v);
}

// Weight gradients of one chunk, one matrix per half-warp at the same time; lane -> hidden unit r, AB rows per pass:
//   lanes  0..15: last Linear   dW1[q][r0+r] = sum_s gout[q][s] h[r][s]
//   lanes 16..31: first Linear  dW0[i][r0+r] = sum_s x_i[s] g[r][s],  db0[r0+r] = sum_s g[r][s]
// (the row operand is one broadcast per half-warp, the slab row is loaded once per AB rows)
template <int AB>
__device__ __forceinline__ void split_wgrad(const CnfDims& d, int d0, int d1, float* Gn, int r0, const float* gout, const float* act,
                                            const int* cond, const float* sl_h, const float* sl_g, int lane,
                                            bool first) {
  const int Hp = d.Hp[0];
  const int half = lane >> 4, r = lane & 15;
  const float* brow = (half ? sl_g : sl_h) + r * SPL_SLP;
  const float* abase = half ? act : gout;
  const int nrows = half ? d1 : d0;
  const int nmax = d0 > d1 ? d0 : d1;
  float* Gm = Gn + (half ? d.w_off[0] : d.w_off[1]) + r0 + r;
  for (int a0 = 0; a0 < nmax; a0 += AB) {
    int aoff[AB];
    float acc[AB];
#pragma unroll
    for (int a = 0; a < AB; ++a) {
      const int idx = min(a0 + a, nrows - 1);
      aoff[a] = (half ? cond[idx] : idx) * SPL_TS;
      acc[a] = 0.f;
    }
    float bs = 0.f;
#pragma unroll
    for (int s4 = 0; s4 < SPL_TS / 4; ++s4) {
      const float4 bv = *reinterpret_cast<const float4*>(brow + 4 * s4);
      bs += (bv.x + bv.y) + (bv.z + bv.w);
#pragma unroll
      for (int a = 0; a < AB; ++a) {
        const float4 av = *reinterpret_cast<const float4*>(abase + aoff[a] + 4 * s4);
        acc[a] = fmaf(av.x, bv.x, acc[a]);
        acc[a] = fmaf(av.y, bv.y, acc[a]);
        acc[a] = fmaf(av.z, bv.z, acc[a]);
        acc[a] = fmaf(av.w, bv.w, acc[a]);
      }
    }
#pragma unroll
    for (int a = 0; a < AB; ++a)
      if (a0 + a < nrows) grad_put(Gm + (size_t)(a0 + a) * Hp, acc[a], first);
    if (half && a0 == 0) grad_put(Gn + d.b_off[0] + r0 + r, bs, first);
  }
}

// Backward of this warp's chunk of one net.  Per lane (= sample): g = relu'(h) * W1[:, chunk]^T gout and the
// chunk's share of the input gradient, added into pg[c][lane] (set when init).  Then the chunk's weight
// gradients as 32-sample dot products of shared-memory rows: h and g go through the warp's private slab
// (slab: [2][16][SPL_SLP]).
__device__ __forceinline__ void split_net_backward(const CnfDims& d, int d0, int d1, const float* Wn, float* Gn, int r0,
                                                   const float (&h)[1][CH], const float* gout, const float* act,
                                                   const int* cond, float* pg, float* slab, bool init,
                                                   bool bias_owner, int lane, bool first) {
  const int Hp = d.Hp[0];
  float* sl_h = slab;
  float* sl_g = slab + CH * SPL_SLP;
  float g[CH];
#pragma unroll
  for (int r = 0; r < CH; ++r) g[r] = 0.f;
#pragma unroll 5
  for (int q = 0; q < d0; ++q) {
    const float go = gout[q * SPL_TS + lane];
    const float* wrow = Wn + d.w_off[1] + (size_t)q * Hp + r0;
#pragma unroll
    for (int r4 = 0; r4 < CH / 4; ++r4) {
      const float4 wv = *reinterpret_cast<const float4*>(wrow + 4 * r4);
      g[4 * r4 + 0] = fmaf(wv.x, go, g[4 * r4 + 0]);
      g[4 * r4 + 1] = fmaf(wv.y, go, g[4 * r4 + 1]);
      g[4 * r4 + 2] = fmaf(wv.z, go, g[4 * r4 + 2]);
      g[4 * r4 + 3] = fmaf(wv.w, go, g[4 * r4 + 3]);
    }
  }
  __syncwarp();                                       // the slab's previous readers are done
#pragma unroll
  for (int r = 0; r < CH; ++r) {
    g[r] = (h[0][r] > 0.f) ? g[r] : 0.f;
    sl_h[r * SPL_SLP + lane] = h[0][r];
    sl_g[r * SPL_SLP + lane] = g[r];
  }
#pragma unroll 5
  for (int i = 0; i < d1; ++i) {
    const float* wrow = Wn + d.w_off[0] + (size_t)i * Hp + r0;
    float acc = 0.f;
#pragma unroll
    for (int r4 = 0; r4 < CH / 4; ++r4) {
      const float4 wv = *reinterpret_cast<const float4*>(wrow + 4 * r4);
      acc = fmaf(wv.x, g[4 * r4 + 0], acc);
      acc = fmaf(wv.y, g[4 * r4 + 1], acc);
      acc = fmaf(wv.z, g[4 * r4 + 2], acc);
      acc = fmaf(wv.w, g[4 * r4 + 3], acc);
    }
    float* pp = pg + i * SPL_TS + lane;
    *pp = init ? acc : *pp + acc;
  }
  __syncwarp();                                       // slab complete
  // weight gradients of the chunk; K = 10 (d0 = d1 = 5) gets an exact 5-row pass
  if (d0 <= 5 && d1 <= 5) split_wgrad<5>(d, d0, d1, Gn, r0, gout, act, cond, sl_h, sl_g, lane, first);
  else split_wgrad<8>(d, d0, d1, Gn, r0, gout, act, cond, sl_h, sl_g, lane, first);
  // db1[q] = sum_s gout[q][s], by the owner of chunk 0
  if (bias_owner)
    for (int q = lane; q < d0; q += 32) grad_put(Gn + d.b_off[1] + q, row_sum32(gout + q * SPL_TS), first);
}

// Loss head of a 32-sample tile, lane = sample (act / gact rows of SPL_TS floats): accumulates the NLL sums,
// writes d loss / d z into gact (NLL head) and returns d loss / d logdet of the lane's sample.
__device__ __forceinline__ float tile_head32(const CnfDims& d, const float* act, float* gact, const int* pi_last,
                                             const int64_t* __restrict__ labels, const float* __restrict__ gld_ext,
                                             int64_t base, int64_t N, int lane, float ld, float eps, float gamma,
                                             float inv_n, int head, bool do_bwd, double& a_loss, double& a_ce,
                                             double& a_ld, double& a_bad) {
  constexpr int TS = SPL_TS;
  const int64_t n = base + lane;
  const bool valid = n < N;
  float gld = 0.f;
  if (head == CNF_HEAD_NLL) {
    float mx = -INFINITY;
    for (int j = 0; j < d.K; ++j) mx = fmaxf(mx, act[pi_last[j] * TS + lane]);
    float se = 0.f;
    for (int j = 0; j < d.K; ++j) se += expf(act[pi_last[j] * TS + lane] - mx);
    int yy = valid ? (int)labels[n] : 0;
    yy = min(max(yy, 0), d.K - 1);  // out-of-range labels are clamped, never read out of bounds
    const float zy = act[pi_last[yy] * TS + lane];
    const float inv_se = 1.f / se;
    const float py = expf(zy - mx) * inv_se;
    float ce, coef;
    if (eps == 0.f) { ce = (zy - mx) - logf(se); coef = 1.f; }
    else            { ce = logf(py + eps); coef = py / (py + eps); }
    if (valid) {
      const float tot = ce + gamma * ld;
      a_loss += (double)tot; a_ce += (double)ce; a_ld += (double)ld;
      if (!isfinite(tot)) a_bad += 1.0;
    }
    if (do_bwd) {
      const float sc = valid ? -inv_n * coef : 0.f;
      for (int j = 0; j < d.K; ++j) {
        const int p = pi_last[j];
        const float pj = expf(act[p * TS + lane] - mx) * inv_se;
        gact[p * TS + lane] = sc * ((j == yy ? 1.f : 0.f) - pj);
      }
      gld = valid ? -gamma * inv_n : 0.f;
    }
  } else {
    gld = valid ? gld_ext[n] : 0.f;
  }
  return gld;
}

// DC: 5 = the coupling split of K = 10 (d0 = d1 = 5; BASELINE configs C2/C3/C5) known at compile time, 0 = any
template <int DC>
__global__ void __launch_bounds__(512)
flow_train_split_kernel(CnfDims d, const float* __restrict__ packed, const int* __restrict__ tables,
                        const float* __restrict__ xin, const int64_t* __restrict__ labels,
                        const float* __restrict__ gz_ext, const float* __restrict__ gld_ext,
                        float* __restrict__ gx_out, float* __restrict__ partials, double* __restrict__ loss_acc,
                        int64_t N, float eps, float gamma, float inv_n, int head) {
  extern __shared__ __align__(16) float smem[];
  const int NT = blockDim.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, NW = NT >> 5;
  constexpr int TS = SPL_TS;
  const bool do_bwd = (partials != nullptr);
  const int d0 = DC ? DC : d.d0, d1 = DC ? DC : d.d1;
  const bool has_s = d.nets & 1, has_t = d.nets & 2;
  // with both conditioners and room for twice the warps, one half of the CTA runs the s-net and the other
  // the t-net concurrently (the host launches 2 x chunks warps); otherwise every warp runs both in turn
  const int nch = d.Hp[0] / CH;
  const bool par = has_s && has_t && NW == 2 * nch;
  const int NWn = par ? NW / 2 : NW;
  const int cw = par ? warp % NWn : warp;             // this warp's chunk
  const bool do_s = has_s && (!par || warp < NWn), do_t = has_t && (!par || warp >= NWn);
  const SplitSmem sm = make_split(d, NW, NWn);
  const int dpart = d0 > d1 ? d0 : d1;
  int* tab = reinterpret_cast<int*>(smem + sm.tab);
  float* act = smem + sm.act;
  float* gact = smem + sm.gact;
  float* tape = smem + sm.tape;
  float* stape = smem + sm.stape;
  float* part = smem + sm.part;
  float* gout_s = smem + sm.gout;
  float* gout_t = gout_s + d0 * TS;
  float* gld_sm = smem + sm.gld;
  float* ldp = smem + sm.ldp;
  float* slab = smem + sm.slab + (size_t)warp * 2 * CH * SPL_SLP;
  float* wbuf = smem + sm.w;
  const int r0 = cw * CH;                             // this warp's hidden units
  const int t_slot = has_s ? 1 : 0;                   // the t-net's block inside a layer
  float* Grow = do_bwd ? partials + (size_t)blockIdx.x * d.n_packed : nullptr;   // grid <= grad_rows: a private row
  for (int i = tid; i < d.n_tables; i += NT) tab[i] = tables[i];
  const int* pi_last = tab + d.tab_pi + d.L * d.K;
  const int64_t ntiles = (N + TS - 1) / TS;
  double a_loss = 0.0, a_ce = 0.0, a_ld = 0.0, a_bad = 0.0;
  bool first = true;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t base = tile * TS;
    __syncthreads();                                  // tables visible; the previous tile is fully consumed
    stage_layer_async(wbuf, packed, d.layer_stride, tid, NT);
    load_tile(act, xin, base, N, d.K, TS, TS, nullptr, tid, NT);
    if (head == CNF_HEAD_EXTERNAL) load_tile(gact, gz_ext, base, N, d.K, TS, TS, pi_last, tid, NT);
    float ld_part = 0.f;
    // ---- forward ----------------------------------------------------------------------
    for (int l = 0; l < d.L; ++l) {
      stage_wait();
      __syncthreads();                                // layer l's weights and the tile state are visible
      if (l + 1 < d.L)
        stage_layer_async(wbuf + ((l + 1) & 1) * d.layer_stride, packed + (size_t)(l + 1) * d.layer_stride,
                          d.layer_stride, tid, NT);
      const float* Wl = wbuf + (l & 1) * d.layer_stride;
      const int* cond = tab + d.tab_cond + l * d1;
      const int* trans = tab + d.tab_trans + l * d0;
      float h[1][CH];
      if (do_s) {
        split_hidden(h, d, d1, Wl, r0, act, cond, lane);
        split_partial_out(part + (size_t)cw * dpart * TS, h, d, d0, Wl, r0, lane);
      }
      if (do_t) {
        const float* Wt = Wl + (size_t)t_slot * d.net_stride;
        split_hidden(h, d, d1, Wt, r0, act, cond, lane);
        split_partial_out(part + (size_t)(NWn + cw) * dpart * TS, h, d, d0, Wt, r0, lane);
      }
      __syncthreads();
      for (int q = warp; q < d0; q += NW) {
        float sv = 0.f, tv = 0.f;
        if (has_s) {
          sv = Wl[d.b_off[1] + q];
          for (int w = 0; w < NWn; ++w) sv += part[((size_t)w * dpart + q) * TS + lane];
        }
        if (has_t) {
          tv = Wl[(size_t)t_slot * d.net_stride + d.b_off[1] + q];
          for (int w = 0; w < NWn; ++w) tv += part[((size_t)(NWn + w) * dpart + q) * TS + lane];
        }
        const int p = trans[q];
        const float xv = act[p * TS + lane];
        tape[(l * d0 + q) * TS + lane] = xv;
        stape[(l * d0 + q) * TS + lane] = sv;
        act[p * TS + lane] = xv * expf(sv) + tv;
        ld_part += sv;
      }
    }
    ldp[warp * TS + lane] = ld_part;
    __syncthreads();
    // ---- loss head (warp 0, lane = sample) ---------------------------------------------
    if (warp == 0) {
      float ld = 0.f;
      for (int w = 0; w < NW; ++w) ld += ldp[w * TS + lane];
      gld_sm[lane] = tile_head32(d, act, gact, pi_last, labels, gld_ext, base, N, lane, ld, eps, gamma, inv_n, head,
                                 do_bwd, a_loss, a_ce, a_ld, a_bad);
    }
    // ---- backward ---------------------------------------------------------------------
    if (do_bwd) {
      for (int l = d.L - 1; l >= 0; --l) {
        if (l != d.L - 1) stage_wait();
        __syncthreads();                              // gact / gout / act of the step before are visible
        if (l > 0)
          stage_layer_async(wbuf + ((l - 1) & 1) * d.layer_stride, packed + (size_t)(l - 1) * d.layer_stride,
                            d.layer_stride, tid, NT);
        const float* Wl = wbuf + (l & 1) * d.layer_stride;
        float* Gl = Grow + (size_t)l * d.layer_stride;
        const int* cond = tab + d.tab_cond + l * d1;
        const int* trans = tab + d.tab_trans + l * d0;
        const float* Wt = Wl + (size_t)t_slot * d.net_stride;
        float hs[1][CH], ht[1][CH];
        // output gradients of both nets from the taped x and s (the transformed logits are not inputs of this
        // layer's nets, so stepping them back here does not disturb the recompute below)
        const float gld = gld_sm[lane];
        for (int q = warp; q < d0; q += NW) {
          const int p = trans[q];
          const float gy = gact[p * TS + lane];
          const float xv = tape[(l * d0 + q) * TS + lane];
          if (has_s) {
            const float es = expf(stape[(l * d0 + q) * TS + lane]);
            gout_s[q * TS + lane] = gy * xv * es + gld;
            gact[p * TS + lane] = gy * es;
          }
          gout_t[q * TS + lane] = gy;
          act[p * TS + lane] = xv;                     // the tile state steps back to the input of layer l
        }
        if (do_s) split_hidden(hs, d, d1, Wl, r0, act, cond, lane);
        if (do_t && par) split_hidden(ht, d, d1, Wt, r0, act, cond, lane);   // its own warps: ahead of the barrier
        __syncthreads();
        float* pg = part + (size_t)warp * dpart * TS;
        if (do_s) split_net_backward(d, d0, d1, Wl, Gl, r0, hs, gout_s, act, cond, pg, slab, true, cw == 0, lane, first);
        if (do_t && !par) split_hidden(ht, d, d1, Wt, r0, act, cond, lane);  // after the s-net: the two never live together
        if (do_t)
          split_net_backward(d, d0, d1, Wt, Gl + (size_t)t_slot * d.net_stride, r0, ht, gout_t, act, cond, pg, slab,
                             !do_s, cw == 0, lane, first);
        __syncthreads();
        for (int c = warp; c < d1; c += NW) {
          float acc = 0.f;
          for (int w = 0; w < NW; ++w) acc += part[((size_t)w * dpart + c) * TS + lane];
          gact[cond[c] * TS + lane] += acc;
        }
      }
      if (gx_out != nullptr) {
        __syncthreads();
        store_tile(gact, gx_out, base, N, d.K, TS, TS, nullptr, tid, NT);
      }
      first = false;
    }
  }
  if (loss_acc != nullptr && head == CNF_HEAD_NLL && warp == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a_loss += __shfl_xor_sync(0xffffffffu, a_loss, o);
      a_ce += __shfl_xor_sync(0xffffffffu, a_ce, o);
      a_ld += __shfl_xor_sync(0xffffffffu, a_ld, o);
      a_bad += __shfl_xor_sync(0xffffffffu, a_bad, o);
    }
    if (lane == 0) {
      atomicAdd(loss_acc + 0, a_loss);
      atomicAdd(loss_acc + 1, a_ce);
      atomicAdd(loss_acc + 2, a_ld);
      atomicAdd(loss_acc + 3, a_bad);
    }
  }
}

// --------------------------------------------------------------------------------------
// The same 32-sample-tile scheme for conditioners with two or more hidden layers (each <= 256 units):
// warp w owns hidden units [16w, 16w+16) of EVERY hidden layer.  A layer's activations go to shared memory
// (hbuf) so that the next layer's chunks can read all of them, with a barrier per layer; going backwards
// the pre-activation gradient of layer j is exchanged the same way (gx, two buffers in turn) and warp w
// forms its chunk of layer j-1's gradient from it.  Weight gradients: rows of the layer's input (shared
// memory) against the warp's 16 x 32 slab of gradients, 8 rows per pass, the two half-warps on alternate
// row blocks.  Weights are read from global memory (uniform 128-bit loads, L1/L2 resident: a middle
// Linear alone is 64 KB at 128 x 128).  The one-thread-per-sample kernel stores every hidden activation of
// a tile plus two gradient buffers per sample and only fits 32-64 samples per SM for such nets.
// --------------------------------------------------------------------------------------
struct DeepSmem { int tab, act, gact, tape, part, gout, gld, ldp, slab, hbuf, gx, total; };

__host__ __device__ inline DeepSmem make_deep(const CnfDims& d, int NW) {
  DeepSmem s;
  const int dpart = d.d0 > d.d1 ? d.d0 : d.d1;
  int hsum = 0;
  for (int j = 0; j < d.m; ++j) hsum += d.Hp[j];
  int off = 0;
  s.tab = off; off += (d.n_tables + 3) / 4 * 4;
  s.act = off; off += d.K * SPL_TS;
  s.gact = off; off += d.K * SPL_TS;
  s.tape = off; off += d.L * d.d0 * SPL_TS;
  s.part = off; off += NW * dpart * SPL_TS;         // one net's partial sums at a time
  s.gout = off; off += 2 * d.d0 * SPL_TS;
  s.gld = off; off += SPL_TS;
  s.ldp = off; off += NW * SPL_TS;
  s.slab = off; off += NW * 2 * CH * (SPL_TS + 4);
  s.hbuf = off; off += hsum * SPL_TS;
  s.gx = off; off += 2 * d.Hmax * SPL_TS;
  s.total = off;
  return s;
}

// G[i*ldg + r] (+)= sum_s rows[row(i)][s] * slab[r][s], i < nrows, r < 16; Gb[r] (+)= sum_s slab[r][s].
// Lanes 0..15 take the even 8-row blocks, lanes 16..31 the odd ones.
__device__ __forceinline__ void deep_wgrad(float* G, int ldg, const float* rows, const int* idx, int nrows,
                                           const float* slab, float* Gb, int lane, bool first) {
  const int half = lane >> 4, r = lane & 15;
  const float* brow = slab + r * SPL_SLP;
  for (int a0 = half * 8; a0 < nrows; a0 += 16) {
    int aoff[8];
    float acc[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      const int i = min(a0 + a, nrows - 1);
      aoff[a] = (idx ? idx[i] : i) * SPL_TS;
      acc[a] = 0.f;
    }
    float bs = 0.f;
#pragma unroll
    for (int s4 = 0; s4 < SPL_TS / 4; ++s4) {
      const float4 bv = *reinterpret_cast<const float4*>(brow + 4 * s4);
      bs += (bv.x + bv.y) + (bv.z + bv.w);
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        const float4 av = *reinterpret_cast<const float4*>(rows + aoff[a] + 4 * s4);
        acc[a] = fmaf(av.x, bv.x, acc[a]);
        acc[a] = fmaf(av.y, bv.y, acc[a]);
        acc[a] = fmaf(av.z, bv.z, acc[a]);
        acc[a] = fmaf(av.w, bv.w, acc[a]);
      }
    }
#pragma unroll
    for (int a = 0; a < 8; ++a)
      if (a0 + a < nrows) grad_put(G + (size_t)(a0 + a) * ldg + r, acc[a], first);
    if (Gb != nullptr && a0 == 0) grad_put(Gb + r, bs, first);
  }
}

// Forward of one conditioner on a 32-sample tile, warp w owning units [16w, 16w+16) of every hidden layer:
// each layer's activations go to hbuf (a block barrier after each layer); h = this warp's chunk of the last
// hidden layer; with pdst, this warp's share of the d0 outputs goes to pdst[warp][q][lane].  Block-cooperative.
__device__ __forceinline__ void deep_net_forward(const CnfDims& d, const int* hoff, const int* nch, float* hbuf,
                                                 const float* act, const float* Wn, const int* cond, float* pdst,
                                                 int dpart, int warp, int lane, float (&h)[1][CH]) {
  constexpr int TS = SPL_TS;
  const int m = d.m, r0 = warp * CH, HL = d.Hp[m - 1];
  for (int j = 0; j < m; ++j) {
    if (warp < nch[j]) {
      chunk_from_inputs<1, false>(h, Wn + d.w_off[j], d.Hp[j], Wn + d.b_off[j], r0, j ? d.Hp[j - 1] : d.d1,
                                  j ? hbuf + hoff[j - 1] * TS : act, j ? nullptr : cond, TS, lane, 0);
#pragma unroll
      for (int r = 0; r < CH; ++r) {
        h[0][r] = fmaxf(h[0][r], 0.f);
        hbuf[(hoff[j] + r0 + r) * TS + lane] = h[0][r];
      }
    }
    __syncthreads();
  }
  if (pdst != nullptr && warp < nch[m - 1]) {
    for (int q = 0; q < d.d0; ++q) {
      const float* wrow = Wn + d.w_off[m] + (size_t)q * HL + r0;
      float acc = 0.f;
#pragma unroll
      for (int r4 = 0; r4 < CH / 4; ++r4) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(wrow) + r4);
        acc = fmaf(wv.x, h[0][4 * r4 + 0], acc);
        acc = fmaf(wv.y, h[0][4 * r4 + 1], acc);
        acc = fmaf(wv.z, h[0][4 * r4 + 2], acc);
        acc = fmaf(wv.w, h[0][4 * r4 + 3], acc);
      }
      pdst[((size_t)warp * dpart + q) * TS + lane] = acc;
    }
  }
}

// Forward / inverse (+ log-det, optional per-layer outputs) on the same tiles: for nets with two or more hidden
// layers at any batch size and for single-hidden-layer nets on small batches, where one thread per sample
// leaves the GPU idle.  Same arithmetic per element as flow_apply_kernel.
struct DeepFwdSmem { int tab, act, part, outs, ldp, hbuf, total; };
__host__ __device__ inline DeepFwdSmem make_deep_fwd(const CnfDims& d, int NW) {
  DeepFwdSmem s;
  int hsum = 0;
  for (int j = 0; j < d.m; ++j) hsum += d.Hp[j];
  int off = 0;
  s.tab = off; off += (d.n_tables + 3) / 4 * 4;
  s.act = off; off += d.K * SPL_TS;
  s.part = off; off += NW * d.d0 * SPL_TS;
  s.outs = off; off += d.d0 * SPL_TS;
  s.ldp = off; off += NW * SPL_TS;
  s.hbuf = off; off += hsum * SPL_TS;
  s.total = off;
  return s;
}

__global__ void __launch_bounds__(512)
flow_apply_deep_kernel(CnfDims d, const float* __restrict__ packed, const int* __restrict__ tables,
                       const float* __restrict__ xin, float* __restrict__ zout, float* __restrict__ logdet,
                       float* __restrict__ zs, int64_t N, int inverse) {
  extern __shared__ __align__(16) float smem[];
  const int NT = blockDim.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, NW = NT >> 5;
  constexpr int TS = SPL_TS;
  const int m = d.m, d0 = d.d0, d1 = d.d1;
  const bool has_s = d.nets & 1, has_t = d.nets & 2;
  const DeepFwdSmem sm = make_deep_fwd(d, NW);
  int* tab = reinterpret_cast<int*>(smem + sm.tab);
  float* act = smem + sm.act;
  float* part = smem + sm.part;
  float* outs_s = smem + sm.outs;
  float* ldp = smem + sm.ldp;
  float* hbuf = smem + sm.hbuf;
  int hoff[CNF_MAX_HIDDEN], nch[CNF_MAX_HIDDEN];
  {
    int o = 0;
    for (int j = 0; j < m; ++j) { hoff[j] = o; o += d.Hp[j]; nch[j] = d.Hp[j] / CH; }
  }
  const int t_slot = has_s ? 1 : 0;
  for (int i = tid; i < d.n_tables; i += NT) tab[i] = tables[i];
  const int* pi_last = tab + d.tab_pi + d.L * d.K;
  const int64_t ntiles = (N + TS - 1) / TS;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t base = tile * TS;
    __syncthreads();                                   // tables visible; the previous tile is stored
    load_tile(act, xin, base, N, d.K, TS, TS, inverse ? pi_last : nullptr, tid, NT);
    __syncthreads();
    float ld_part = 0.f;
    float h[1][CH];
    for (int li = 0; li < d.L; ++li) {
      const int l = inverse ? d.L - 1 - li : li;
      const float* Wl = packed + (size_t)l * d.layer_stride;
      const int* cond = tab + d.tab_cond + l * d1;
      const int* trans = tab + d.tab_trans + l * d0;
      if (has_s) {
        deep_net_forward(d, hoff, nch, hbuf, act, Wl, cond, part, d0, warp, lane, h);
        __syncthreads();
        for (int q = warp; q < d0; q += NW) {
          float sv = __ldg(Wl + d.b_off[m] + q);
          for (int w = 0; w < nch[m - 1]; ++w) sv += part[((size_t)w * d0 + q) * TS + lane];
          outs_s[q * TS + lane] = sv;
        }
      }
      if (has_t) deep_net_forward(d, hoff, nch, hbuf, act, Wl + (size_t)t_slot * d.net_stride, cond, part, d0, warp, lane, h);
      __syncthreads();
      for (int q = warp; q < d0; q += NW) {
        float sv = 0.f, tv = 0.f;
        if (has_s) sv = outs_s[q * TS + lane];         // written by this very thread
        if (has_t) {
          tv = __ldg(Wl + (size_t)t_slot * d.net_stride + d.b_off[m] + q);
          for (int w = 0; w < nch[m - 1]; ++w) tv += part[((size_t)w * d0 + q) * TS + lane];
        }
        const int p = trans[q];
        const float xv = act[p * TS + lane];
        if (!inverse) { act[p * TS + lane] = xv * expf(sv) + tv; ld_part += sv; }
        else          { act[p * TS + lane] = (xv - tv) * expf(-sv); ld_part -= sv; }
      }
      __syncthreads();
      if (zs != nullptr) {
        // forward: zs[l] is the output of layer l in its logical order pi_{l+1};
        // inverse: xs[li] is the input of layer l in logical order pi_l.
        store_tile(act, zs + (size_t)li * N * d.K, base, N, d.K, TS, TS, tab + d.tab_pi + (inverse ? l : l + 1) * d.K,
                   tid, NT);
      }
    }
    ldp[warp * TS + lane] = ld_part;
    __syncthreads();
    if (warp == 0) {
      float ld = 0.f;
      for (int w = 0; w < NW; ++w) ld += ldp[w * TS + lane];
      if (base + lane < N) logdet[base + lane] = ld;
    }
    store_tile(act, zout, base, N, d.K, TS, TS, inverse ? nullptr : pi_last, tid, NT);
  }
}

__global__ void __launch_bounds__(512)
flow_train_deep_kernel(CnfDims d, const float* __restrict__ packed, const int* __restrict__ tables,
                       const float* __restrict__ xin, const int64_t* __restrict__ labels,
                       const float* __restrict__ gz_ext, const float* __restrict__ gld_ext,
                       float* __restrict__ gx_out, float* __restrict__ partials, double* __restrict__ loss_acc,
                       int64_t N, float eps, float gamma, float inv_n, int head) {
  extern __shared__ __align__(16) float smem[];
  const int NT = blockDim.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, NW = NT >> 5;
  constexpr int TS = SPL_TS;
  const bool do_bwd = (partials != nullptr);
  const int m = d.m, d0 = d.d0, d1 = d.d1;
  const bool has_s = d.nets & 1, has_t = d.nets & 2;
  const DeepSmem sm = make_deep(d, NW);
  const int dpart = d0 > d1 ? d0 : d1;
  int* tab = reinterpret_cast<int*>(smem + sm.tab);
  float* act = smem + sm.act;
  float* gact = smem + sm.gact;
  float* tape = smem + sm.tape;
  float* part = smem + sm.part;
  float* gout_s = smem + sm.gout;
  float* gout_t = gout_s + d0 * TS;
  float* gld_sm = smem + sm.gld;
  float* ldp = smem + sm.ldp;
  float* sl_h = smem + sm.slab + (size_t)warp * 2 * CH * SPL_SLP;
  float* sl_g = sl_h + CH * SPL_SLP;
  float* hbuf = smem + sm.hbuf;
  float* gxb = smem + sm.gx;
  int hoff[CNF_MAX_HIDDEN], nch[CNF_MAX_HIDDEN];
  {
    int o = 0;
    for (int j = 0; j < m; ++j) { hoff[j] = o; o += d.Hp[j]; nch[j] = d.Hp[j] / CH; }
  }
  const int r0 = warp * CH;
  const int HL = d.Hp[m - 1];                          // width of the last hidden layer
  const int t_slot = has_s ? 1 : 0;
  float* Grow = do_bwd ? partials + (size_t)blockIdx.x * d.n_packed : nullptr;   // grid <= grad_rows: a private row
  for (int i = tid; i < d.n_tables; i += NT) tab[i] = tables[i];
  const int* pi_last = tab + d.tab_pi + d.L * d.K;
  const int64_t ntiles = (N + TS - 1) / TS;
  double a_loss = 0.0, a_ce = 0.0, a_ld = 0.0, a_bad = 0.0;
  bool first = true;

  auto net_fwd = [&](const float* Wn, const int* cond, float* pdst, float (&h)[1][CH]) {
    deep_net_forward(d, hoff, nch, hbuf, act, Wn, cond, pdst, dpart, warp, lane, h);
  };

  // Backward of one conditioner whose activations are in hbuf: weight gradients into Gn, this warp's share of
  // the input gradient into part[warp][c][lane] (set when init, else added).
  auto net_bwd = [&](const float* Wn, float* Gn, const float* gout, const int* cond, bool init) {
    float g[CH];
    int cur = 0;
    if (warp < nch[m - 1]) {
#pragma unroll
      for (int r = 0; r < CH; ++r) g[r] = 0.f;
      for (int q = 0; q < d0; ++q) {
        const float go = gout[q * TS + lane];
        const float* wrow = Wn + d.w_off[m] + (size_t)q * HL + r0;
#pragma unroll
        for (int r4 = 0; r4 < CH / 4; ++r4) {
          const float4 wv = __ldg(reinterpret_cast<const float4*>(wrow) + r4);
          g[4 * r4 + 0] = fmaf(wv.x, go, g[4 * r4 + 0]);
          g[4 * r4 + 1] = fmaf(wv.y, go, g[4 * r4 + 1]);
          g[4 * r4 + 2] = fmaf(wv.z, go, g[4 * r4 + 2]);
          g[4 * r4 + 3] = fmaf(wv.w, go, g[4 * r4 + 3]);
        }
      }
      __syncwarp();
#pragma unroll
      for (int r = 0; r < CH; ++r) {
        const float hv = hbuf[(hoff[m - 1] + r0 + r) * TS + lane];
        g[r] = (hv > 0.f) ? g[r] : 0.f;
        sl_h[r * SPL_SLP + lane] = hv;
        sl_g[r * SPL_SLP + lane] = g[r];
      }
      __syncwarp();
      // last Linear: dW[q][r0+r] = sum_s gout[q][s] h[r][s]
      deep_wgrad(Gn + d.w_off[m] + r0, HL, gout, nullptr, d0, sl_h, nullptr, lane, first);
    }
    if (warp == 0)
      for (int q = lane; q < d0; q += 32) grad_put(Gn + d.b_off[m] + q, row_sum32(gout + q * TS), first);
    for (int j = m - 1; j >= 0; --j) {
      // warps below nch[j] hold the masked gradient of layer j's pre-activations in g and in sl_g
      if (warp < nch[j]) {
        deep_wgrad(Gn + d.w_off[j] + r0, d.Hp[j], j ? hbuf + hoff[j - 1] * TS : act, j ? nullptr : cond,
                   j ? d.Hp[j - 1] : d1, sl_g, Gn + d.b_off[j] + r0, lane, first);
        if (j == 0) {
          for (int c = 0; c < d1; ++c) {
            const float* wrow = Wn + d.w_off[0] + (size_t)c * d.Hp[0] + r0;
            float acc = 0.f;
#pragma unroll
            for (int r4 = 0; r4 < CH / 4; ++r4) {
              const float4 wv = __ldg(reinterpret_cast<const float4*>(wrow) + r4);
              acc = fmaf(wv.x, g[4 * r4 + 0], acc);
              acc = fmaf(wv.y, g[4 * r4 + 1], acc);
              acc = fmaf(wv.z, g[4 * r4 + 2], acc);
              acc = fmaf(wv.w, g[4 * r4 + 3], acc);
            }
            float* pp = part + ((size_t)warp * dpart + c) * TS + lane;
            *pp = init ? acc : *pp + acc;
          }
        } else {
#pragma unroll
          for (int r = 0; r < CH; ++r) gxb[(cur * d.Hmax + r0 + r) * TS + lane] = g[r];
        }
      }
      if (j == 0) break;
      __syncthreads();                                 // layer j's gradient is complete in gx[cur]
      if (warp < nch[j - 1]) {
        // g_{j-1}[r0+rp] = relu'(h_{j-1}) * sum_r W_j[r0+rp][r] g_j[r]
#pragma unroll
        for (int r = 0; r < CH; ++r) g[r] = 0.f;
        const float* Wj = Wn + d.w_off[j] + (size_t)r0 * d.Hp[j];
        const float* gsrc = gxb + (size_t)cur * d.Hmax * TS + lane;
        for (int r4 = 0; r4 < d.Hp[j] / 4; ++r4) {
          const float g0 = gsrc[(4 * r4 + 0) * TS], g1 = gsrc[(4 * r4 + 1) * TS];
          const float g2 = gsrc[(4 * r4 + 2) * TS], g3 = gsrc[(4 * r4 + 3) * TS];
#pragma unroll
          for (int rp = 0; rp < CH; ++rp) {
            const float4 wv = __ldg(reinterpret_cast<const float4*>(Wj + (size_t)rp * d.Hp[j]) + r4);
            g[rp] = fmaf(wv.x, g0, g[rp]);
            g[rp] = fmaf(wv.y, g1, g[rp]);
            g[rp] = fmaf(wv.z, g2, g[rp]);
            g[rp] = fmaf(wv.w, g3, g[rp]);
          }
        }
        __syncwarp();                                  // the slab's readers of the layer above are done
#pragma unroll
        for (int r = 0; r < CH; ++r) {
          const float hv = hbuf[(hoff[j - 1] + r0 + r) * TS + lane];
          g[r] = (hv > 0.f) ? g[r] : 0.f;
          sl_g[r * SPL_SLP + lane] = g[r];
        }
        __syncwarp();
      }
      cur ^= 1;
    }
  };

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t base = tile * TS;
    __syncthreads();
    load_tile(act, xin, base, N, d.K, TS, TS, nullptr, tid, NT);
    if (head == CNF_HEAD_EXTERNAL) load_tile(gact, gz_ext, base, N, d.K, TS, TS, pi_last, tid, NT);
    __syncthreads();
    float ld_part = 0.f;
    float h[1][CH];
    // ---- forward ----------------------------------------------------------------------
    for (int l = 0; l < d.L; ++l) {
      const float* Wl = packed + (size_t)l * d.layer_stride;
      const int* cond = tab + d.tab_cond + l * d1;
      const int* trans = tab + d.tab_trans + l * d0;
      if (has_s) {
        net_fwd(Wl, cond, part, h);
        __syncthreads();
        // s outputs summed right away (kept in gout_s, free in the forward pass) so that the t-net can reuse
        // the partial buffer; its first writes come after the barriers inside net_fwd
        for (int q = warp; q < d0; q += NW) {
          float sv = __ldg(Wl + d.b_off[m] + q);
          for (int w = 0; w < nch[m - 1]; ++w) sv += part[((size_t)w * dpart + q) * TS + lane];
          gout_s[q * TS + lane] = sv;
        }
      }
      if (has_t) net_fwd(Wl + (size_t)t_slot * d.net_stride, cond, part, h);
      __syncthreads();
      for (int q = warp; q < d0; q += NW) {
        float sv = 0.f, tv = 0.f;
        if (has_s) sv = gout_s[q * TS + lane];         // written by this very thread
        if (has_t) {
          tv = __ldg(Wl + (size_t)t_slot * d.net_stride + d.b_off[m] + q);
          for (int w = 0; w < nch[m - 1]; ++w) tv += part[((size_t)w * dpart + q) * TS + lane];
        }
        const int p = trans[q];
        const float xv = act[p * TS + lane];
        tape[(l * d0 + q) * TS + lane] = xv;
        act[p * TS + lane] = xv * expf(sv) + tv;
        ld_part += sv;
      }
      __syncthreads();
    }
    ldp[warp * TS + lane] = ld_part;
    __syncthreads();
    if (warp == 0) {
      float ld = 0.f;
      for (int w = 0; w < NW; ++w) ld += ldp[w * TS + lane];
      gld_sm[lane] = tile_head32(d, act, gact, pi_last, labels, gld_ext, base, N, lane, ld, eps, gamma, inv_n, head,
                                 do_bwd, a_loss, a_ce, a_ld, a_bad);
    }
    // ---- backward ---------------------------------------------------------------------
    if (do_bwd) {
      for (int l = d.L - 1; l >= 0; --l) {
        __syncthreads();
        const float* Wl = packed + (size_t)l * d.layer_stride;
        float* Gl = Grow + (size_t)l * d.layer_stride;
        const int* cond = tab + d.tab_cond + l * d1;
        const int* trans = tab + d.tab_trans + l * d0;
        const float* Wt = Wl + (size_t)t_slot * d.net_stride;
        if (has_s) net_fwd(Wl, cond, part, h);
        __syncthreads();
        const float gld = gld_sm[lane];
        for (int q = warp; q < d0; q += NW) {
          const int p = trans[q];
          const float gy = gact[p * TS + lane];
          const float xv = tape[(l * d0 + q) * TS + lane];
          if (has_s) {
            float sv = __ldg(Wl + d.b_off[m] + q);
            for (int w = 0; w < nch[m - 1]; ++w) sv += part[((size_t)w * dpart + q) * TS + lane];
            const float es = expf(sv);
            gout_s[q * TS + lane] = gy * xv * es + gld;
            gact[p * TS + lane] = gy * es;
          }
          gout_t[q * TS + lane] = gy;
          act[p * TS + lane] = xv;                     // the tile state steps back to the input of layer l
        }
        __syncthreads();
        if (has_s) net_bwd(Wl, Gl, gout_s, cond, true);
        if (has_t) {
          __syncthreads();                             // the s-net's activations in hbuf are no longer needed
          net_fwd(Wt, cond, nullptr, h);
          net_bwd(Wt, Gl + (size_t)t_slot * d.net_stride, gout_t, cond, !has_s);
        }
        __syncthreads();
        for (int c = warp; c < d1; c += NW) {
          float acc = 0.f;
          for (int w = 0; w < nch[0]; ++w) acc += part[((size_t)w * dpart + c) * TS + lane];
          gact[cond[c] * TS + lane] += acc;
        }
      }
      if (gx_out != nullptr) {
        __syncthreads();
        store_tile(gact, gx_out, base, N, d.K, TS, TS, nullptr, tid, NT);
      }
      first = false;
    }
  }
  if (loss_acc != nullptr && head == CNF_HEAD_NLL && warp == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a_loss += __shfl_xor_sync(0xffffffffu, a_loss, o);
      a_ce += __shfl_xor_sync(0xffffffffu, a_ce, o);
      a_ld += __shfl_xor_sync(0xffffffffu, a_ld, o);
      a_bad += __shfl_xor_sync(0xffffffffu, a_bad, o);
    }
    if (lane == 0) {
      atomicAdd(loss_acc + 0, a_loss);
      atomicAdd(loss_acc + 1, a_ce);
      atomicAdd(loss_acc + 2, a_ld);
      atomicAdd(loss_acc + 3, a_bad);
    }
  }
}

// --------------------------------------------------------------------------------------
// launch plumbing
// --------------------------------------------------------------------------------------
struct LaunchCfg { int spt, nt; bool ws; int wl; size_t smem; };

// limits of the current device (cached per device in cnf_device.cu); thread-local so that concurrent host threads
// driving different devices do not see each other's values
thread_local int g_max_smem = -1, g_num_sms = -1;

int device_limits() {
  CnfDevInfo di;
  int rc = cnf_dev_info(&di);
  if (rc) return rc;
  g_max_smem = di.max_smem; g_num_sms = di.sms;
  return CNF_OK;
}

// Pick the widest tile whose shared-memory plan fits, preferring weights in shared memory.
int choose_cfg(const CnfDims& d, bool backward, LaunchCfg* out, int tail_bins = 0) {
  const int budget = g_max_smem - (backward ? 1024 : 256);
  const int spts[2] = {2, 1};
  // 256-thread CTAs share one staged copy of the weights between twice as many warps: 16 instead of 8
  // resident warps per SM at the C2 shape, 1.83 -> 1.34 ms per 2^20 samples on B200
  int nt_max = 256;
  if (const char* v = cnf_switch(CNF_SW_FP32_NT)) { const int n = atoi(v); if (n == 256 || n == 128) nt_max = n; }
  int ws_first = 1;
  if (const char* v = cnf_switch(CNF_SW_FP32_WS)) ws_first = atoi(v) ? 1 : 0;   // experiment switch: 0 = never stage the weights
  out->wl = 0;
  // training: a full-width tile with one layer's weights staged at a time beats a narrower tile with all of them
  if (backward && ws_first && !cnf_switch(CNF_SW_FP32_NO_WL)) {
    for (int mode = 1; mode <= 2; ++mode) {
      const int TSP = nt_max + 4;
      const size_t bytes = (size_t)make_smem(d, TSP, mode, true).total * 4;
      if ((long long)bytes <= budget) {
        out->spt = 1; out->nt = nt_max; out->ws = true; out->wl = mode == 2; out->smem = bytes;
        return CNF_OK;
      }
    }
  }
  for (int ws = ws_first; ws >= 0; --ws)
    for (int nt = nt_max; nt >= 32; nt >>= 1)
      for (int si = 0; si < 2; ++si) {
        const int spt = spts[si];
        if (nt < 128 && spt > 1) continue;
        const int TSP = nt * spt + 4;
        const Smem sm = make_smem(d, TSP, ws, backward, tail_bins);
        const size_t bytes = (size_t)sm.total * 4;
        if ((long long)bytes <= budget) {
          out->spt = spt; out->nt = nt; out->ws = ws != 0; out->smem = bytes;
          return CNF_OK;
        }
      }
  cnf_set_error("model does not fit shared memory (K=%d, Hmax=%d, L=%d)", d.K, d.Hmax, d.L);
  return CNF_E_SMEM;
}

template <typename Kern>
int set_smem(Kern k, size_t bytes) { return cnf_kernel_smem(k, bytes); }

}  // namespace

// register-resident kernel for K = 10 (cnf_flow_fp32r.cu)
bool cnf_fp32r_supported(const cnf_flow_desc* desc, const CnfDims& d, const float* x, const float* z, int tail_bins,
                         int max_smem, size_t* smem_out, bool two_ok);
int cnf_fp32r_apply(const CnfDims& d, const float* packed, const int32_t* tables, const float* x, float* z, float* logdet,
                    int64_t N, int inverse, const CnfTail* tail, size_t smem, int sms, int variant, cudaStream_t st);
bool cnf_fp32r_train_supported(const cnf_flow_desc* desc, const CnfDims& d, const float* x, int max_smem, size_t* smem_out);
int cnf_fp32r_train(const CnfDims& d, const float* packed, const int32_t* tables, const float* x, const int64_t* y,
                    float* partials, double* loss_acc, int64_t N, float eps, float gamma, float inv_n, size_t smem_fwd,
                    int sms, int max_smem, int variant, int64_t* rows_out, const float* gz_ext, const float* gld_ext,
                    cudaStream_t st);

int cnf_fp32_apply(const cnf_flow_desc* desc, const float* packed, const int32_t* tables, const float* x, float* z,
                   float* logdet, float* zs, int64_t N, int inverse, cudaStream_t st) {
  CnfDims d;
  int rc = cnf_make_dims(desc, &d);
  if (rc) return rc;
  if (N == 0) return CNF_OK;
  if (!packed || !tables || !x || !z || !logdet || N < 0) { cnf_set_error("null pointer / negative N"); return CNF_E_ARG; }
  if ((rc = device_limits())) return rc;
  // K = 10, one hidden layer, both nets, standard flips: the register-resident kernel from 65,536 samples, with 2, 4 or
  // 8 samples per thread (256- / 512- / 1024-sample tiles) so that the tiles fill the SMs.  Measured on B200
  // (profiles/microbench/fp32r_crossover.py, us per call, generic kernel / best register variant): N = 65,536: 141 / 108
  // (x2); 131,072: 197 / 153 (x4); 262,144: 386 / 232 (x8); 1 Mi: 1331 / 789; 4 Mi: 5303 / 2669.
  // (CNF_FP32R: "off" disables, a digit forces a (threads, samples per thread) variant -- experiments)
  {
    const char* sw = cnf_switch(CNF_SW_FP32R);
    size_t smem_r = 0;
    // (two tiny hidden layers -- the reference's default [5, 5] -- from 1,024 samples: their alternative is the generic
    //  tile kernel, 5x slower at every size measured)
    const bool tiny2 = d.m == 2;
    if (!zs && N >= (tiny2 ? 1024 : 65536) && !(sw && sw[0] == 'o') && cnf_fp32r_supported(desc, d, x, z, 0, g_max_smem, &smem_r, true)) {
      const int variant = sw ? atoi(sw) : (tiny2 ? (N >= 113000 ? 1 : (N >= 256LL * g_num_sms ? 4 : 5))
                                                 : (N >= 240000 ? 0 : (N >= 113000 ? 1 : 4)));
      return cnf_fp32r_apply(d, packed, tables, x, z, logdet, N, inverse, nullptr, smem_r, g_num_sms, variant, st);
    }
  }
  // 32-sample tiles with the hidden layers split over the warps: nets with two or more hidden layers (the widest
  // of at least 64 units), and single-hidden-layer nets on small batches ("0" disables, "1" forces: experiments)
  {
    const char* sw = cnf_switch(CNF_SW_DEEP_APPLY);
    bool fits = d.m >= 1;
    for (int j = 0; j < d.m; ++j) fits = fits && d.Hp[j] <= 256;
    const bool want = sw ? atoi(sw) != 0 : (d.Hmax >= 64 && (d.m >= 2 || N <= 32768));
    if (fits && want) {
      const int nw = d.Hmax / CH;
      const size_t bytes = (size_t)make_deep_fwd(d, nw).total * sizeof(float);
      if ((long long)bytes <= g_max_smem - 1024) {
        const int64_t nts = (N + SPL_TS - 1) / SPL_TS;
        int per_sm = (int)(g_max_smem / (bytes + 2048));
        const int by_threads = 2048 / (nw * 32), by_regs = 65536 / (64 * nw * 32);   // the kernel needs 64 registers (grid cap only)
        per_sm = per_sm < by_threads ? per_sm : by_threads;
        per_sm = per_sm < by_regs ? per_sm : by_regs;
        if (per_sm < 1) per_sm = 1;
        const int64_t cap = (int64_t)g_num_sms * per_sm;
        const int gridd = (int)(nts < cap ? nts : cap);
        if ((rc = set_smem(flow_apply_deep_kernel, bytes))) return rc;
        flow_apply_deep_kernel<<<gridd, nw * 32, bytes, st>>>(d, packed, tables, x, z, logdet, zs, N, inverse);
        CNF_CHECK_CUDA(cudaGetLastError());
        return CNF_OK;
      }
    }
  }
  LaunchCfg c;
  c.nt = 0; c.spt = 1; c.ws = false; c.wl = 0; c.smem = 0;
  rc = choose_cfg(d, false, &c);
  // single-hidden-layer nets whose weights do not fit shared memory: staged-chunk kernel instead of L1 reads
  if (d.m == 1 && (rc != CNF_OK || !c.ws) && !cnf_switch(CNF_SW_NO_LEAN_TRAIN)) {
    for (int nt = 256; nt >= 64; nt >>= 1) {
      const int TSP = nt + 4;
      const size_t bytes = ((size_t)(d.n_tables + 3) / 4 * 4 + (size_t)d.K * TSP + 2 * (size_t)d.d0 * TSP +
                            (size_t)(d.d1 + 1 + d.d0) * CNF_CH + (size_t)(d.d0 + 3) / 4 * 4) * sizeof(float);
      if ((long long)bytes > (g_max_smem - 1024) / 2 && nt > 64) continue;      // two CTAs per SM when possible
      if ((long long)bytes > g_max_smem - 1024) continue;
      const int64_t ntl = (N + nt - 1) / nt;
      int per_sm = (int)(g_max_smem / (bytes + 1024));
      per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);
      const int64_t capl = (int64_t)g_num_sms * per_sm;
      const int gridl = (int)(ntl < capl ? ntl : capl);
      if ((rc = set_smem(flow_apply_lean_kernel<1>, bytes))) return rc;
      flow_apply_lean_kernel<1><<<gridl, nt, bytes, st>>>(d, packed, tables, x, z, logdet, zs, N, inverse);
      CNF_CHECK_CUDA(cudaGetLastError());
      return CNF_OK;
    }
  }
  if (rc) return rc;
  const int64_t ntiles = (N + c.nt * c.spt - 1) / (c.nt * c.spt);
  int ctas_per_sm = (int)(g_max_smem / (c.smem + 1024));
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  if (ctas_per_sm > 8) ctas_per_sm = 8;
  const int grid = (int)(ntiles < (int64_t)g_num_sms * ctas_per_sm ? ntiles : (int64_t)g_num_sms * ctas_per_sm);
#define LAUNCH_APPLY(SPT, WS)                                                                         \
  do {                                                                                                \
    if ((rc = set_smem(flow_apply_kernel<SPT, WS, 0>, c.smem))) return rc;                            \
    flow_apply_kernel<SPT, WS, 0><<<grid, c.nt, c.smem, st>>>(d, packed, tables, x, z, logdet, zs, N, inverse, CnfTail()); \
  } while (0)
  if (c.spt == 2 && c.ws) LAUNCH_APPLY(2, true);
  else if (c.spt == 2) LAUNCH_APPLY(2, false);
  else if (c.ws) LAUNCH_APPLY(1, true);
  else LAUNCH_APPLY(1, false);
#undef LAUNCH_APPLY
  CNF_CHECK_CUDA(cudaGetLastError());
  return CNF_OK;
}

// Fused pass of cnf_flow_predict on the fp32 path: [centring] -> flow -> [calibrated probabilities | statistics]
// in one launch of flow_apply_kernel<.., TAIL> (every shape whose tile plan fits shared memory).
int cnf_fp32_predict(const cnf_flow_desc* desc, const float* packed, const int32_t* tables, const float* x, float* z,
                     float* logdet, int64_t N, const CnfTail& ta, cudaStream_t st) {
  CnfDims d;
  int rc = cnf_make_dims(desc, &d);
  if (rc) return rc;
  if (N == 0) return CNF_OK;
  if ((rc = device_limits())) return rc;
  {
    size_t smem_r = 0;
    const char* sw = cnf_switch(CNF_SW_FP32R);
    const bool tiny2 = d.m == 2;
    if (N >= (sw ? 1 : (tiny2 ? 1024 : 65536)) && !(sw && sw[0] == 'o') && cnf_fp32r_supported(desc, d, x, z, ta.bins, g_max_smem, &smem_r, true))
      return cnf_fp32r_apply(d, packed, tables, x, z, logdet, N, 0, &ta, smem_r, g_num_sms,
                             tiny2 ? (N >= 113000 ? 1 : (N >= 256LL * g_num_sms ? 4 : 5)) : (N >= 240000 ? 0 : (N >= 113000 ? 1 : 4)), st);
  }
  LaunchCfg c;
  c.nt = 0; c.spt = 1; c.ws = false; c.wl = 0; c.smem = 0;
  if ((rc = choose_cfg(d, false, &c, ta.bins))) return rc;
  const int64_t ntiles = (N + c.nt * c.spt - 1) / (c.nt * c.spt);
  int ctas_per_sm = (int)(g_max_smem / (c.smem + 1024));
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  if (ctas_per_sm > 8) ctas_per_sm = 8;
  const int grid = (int)(ntiles < (int64_t)g_num_sms * ctas_per_sm ? ntiles : (int64_t)g_num_sms * ctas_per_sm);
#define LAUNCH_TAIL(SPT, WS, M)                                                                        \
  do {                                                                                                 \
    if ((rc = set_smem(flow_apply_kernel<SPT, WS, M>, c.smem))) return rc;                             \
    flow_apply_kernel<SPT, WS, M><<<grid, c.nt, c.smem, st>>>(d, packed, tables, x, z, logdet, nullptr, N, 0, ta); \
  } while (0)
#define LAUNCH_TAIL_M(SPT, WS)                                                                         \
  do {                                                                                                 \
    if (ta.mode == CNF_METRICS_LOGITS) LAUNCH_TAIL(SPT, WS, CNF_METRICS_LOGITS);                       \
    else LAUNCH_TAIL(SPT, WS, CNF_METRICS_CALIBRATED);                                                 \
  } while (0)
  if (c.spt == 2 && c.ws) LAUNCH_TAIL_M(2, true);
  else if (c.spt == 2) LAUNCH_TAIL_M(2, false);
  else if (c.ws) LAUNCH_TAIL_M(1, true);
  else LAUNCH_TAIL_M(1, false);
#undef LAUNCH_TAIL_M
#undef LAUNCH_TAIL
  CNF_CHECK_CUDA(cudaGetLastError());
  return CNF_OK;
}

int cnf_fp32_train(const cnf_flow_desc* desc, const float* packed, const int32_t* tables, const float* x,
                   const int64_t* y, const float* gz, const float* gld, float* gx, float* partials, double* loss_acc,
                   int64_t N, float eps, float gamma, float inv_n, int head, int64_t* rows_used, cudaStream_t st) {
  CnfDims d;
  int rc = cnf_make_dims(desc, &d);
  if (rc) return rc;
  if (!packed || !tables || N < 0) { cnf_set_error("null pointer / negative N"); return CNF_E_ARG; }
  // Rows of the partial buffer a launch of `grid` CTAs writes (CTA b -> row b % grad_rows).  With rows_used
  // (HOST out) only those rows are zeroed and reported, so that small batches do not pay for clearing and
  // reducing all grad_rows rows; without it every row is cleared (cnf_grad_reduce reads them all).
  auto clear_rows = [&](int64_t grid) -> int {
    int64_t rows = d.grad_rows_max;
    if (rows_used) { rows = grid < d.grad_rows_max ? grid : d.grad_rows_max; *rows_used = rows; }
    if (partials && rows > 0) CNF_CHECK_CUDA(cudaMemsetAsync(partials, 0, (size_t)rows * d.n_packed * sizeof(float), st));
    return CNF_OK;
  };
  // an empty batch may come with null data pointers (an empty torch tensor has none)
  if (head == CNF_HEAD_NLL && ((N > 0 && !y) || !loss_acc)) { cnf_set_error("NLL head needs labels and loss_acc"); return CNF_E_ARG; }
  if (head == CNF_HEAD_EXTERNAL && ((N > 0 && (!gz || !gld)) || !partials)) { cnf_set_error("external head needs g_z, g_logdet, partials"); return CNF_E_ARG; }
  if ((rc = device_limits())) return rc;
  if (N == 0) return clear_rows(0);
  if (!x) { cnf_set_error("null x"); return CNF_E_ARG; }
  // K = 10, one hidden layer, both nets, NLL head, a batch that fills the GPU: the register-resident kernel.  Measured
  // crossover (fp32r_crossover.py, us per fused pass, 32-sample-tile kernel / register kernel): N = 131,072: 819 / 953;
  // 200,000: 1232 / 994; 262,144: 1612 / 1216; 1 Mi: 6326 / 4594; 4 Mi: 25462 / 15689 (one wave of its 2048-sample
  // tiles takes ~0.95 ms whatever it holds, the tile kernel's time grows with N: they cross at ~155,000 samples).
  // (CNF_FP32R_TRAIN: "off" disables, a digit forces a variant at any N >= 1,024 -- experiments)
  // (also the external head -- the autograd backward of the drop-in Flow, run_experiment3D.py:133-134 -- when the
  //  caller wants no input gradient: K = 3, 10 x [5, 5], N = 1,500: 275 -> ~70 us)
  if ((head == CNF_HEAD_NLL || (head == CNF_HEAD_EXTERNAL && gx == nullptr)) &&
      N >= (d.m == 2 ? 1024 : (cnf_switch(CNF_SW_FP32R_TRAIN) ? 1024 : 160000))) {
    const char* sw = cnf_switch(CNF_SW_FP32R_TRAIN);
    size_t smem_r = 0;
    if (!(sw && sw[0] == 'o') && cnf_fp32r_train_supported(desc, d, x, g_max_smem - 1024, &smem_r)) {
      rc = cnf_fp32r_train(d, packed, tables, x, y, partials, loss_acc, N, eps, gamma, inv_n, smem_r, g_num_sms,
                           g_max_smem, sw ? atoi(sw) : 0, rows_used, head == CNF_HEAD_EXTERNAL ? gz : nullptr,
                           head == CNF_HEAD_EXTERNAL ? gld : nullptr, st);
      if (rc != CNF_E_SMEM) return rc;      // (too many layers for its shared-memory plan: the tile kernels below)
    }
  }
  // single-hidden-layer nets up to 256 hidden units: 32-sample tiles with the hidden layer split over the warps
  // (faster than one thread per sample at every batch size measured, 3x at N <= 10,000)
  {
    const char* sw = cnf_switch(CNF_SW_SPLIT_TRAIN);          // "0" disables (experiments)
    const bool want = sw ? atoi(sw) != 0 : true;
    if (want && d.m == 1 && d.Hp[0] <= 256) {
      const int nch = d.Hp[0] / CH;
      // both nets side by side when the warps fit and every tile gets an SM of its own (the 512-thread CTA is alone on its SM)
      const bool side_by_side = d.n_nets == 2 && nch <= 8 && (N + SPL_TS - 1) / SPL_TS <= g_num_sms && !cnf_switch(CNF_SW_SPLIT_SEQ);
      const int nw = side_by_side ? 2 * nch : nch;
      const size_t bytes = (size_t)make_split(d, nw, nch).total * sizeof(float);
      if ((long long)bytes <= g_max_smem - 1024) {
        const int64_t nts = (N + SPL_TS - 1) / SPL_TS;
        int per_sm = (int)(g_max_smem / (bytes + 2048));
        const int by_threads = 2048 / (nw * 32), by_regs = 65536 / (128 * nw * 32);
        per_sm = per_sm < by_threads ? per_sm : by_threads;
        per_sm = per_sm < by_regs ? per_sm : by_regs;
        if (per_sm < 1) per_sm = 1;
        int64_t cap = (int64_t)g_num_sms * per_sm;
        if (cap > d.grad_rows) cap = d.grad_rows;          // every CTA owns one row of the partial buffer
        const int grids = (int)(nts < cap ? nts : cap);
        const bool dc5 = d.d0 == 5 && d.d1 == 5 && !cnf_switch(CNF_SW_SPLIT_GENERIC);
        if ((rc = dc5 ? set_smem(flow_train_split_kernel<5>, bytes) : set_smem(flow_train_split_kernel<0>, bytes))) return rc;
        if ((rc = clear_rows(grids))) return rc;
        if (dc5)
          flow_train_split_kernel<5><<<grids, nw * 32, bytes, st>>>(d, packed, tables, x, y, gz, gld, gx, partials, loss_acc,
                                                                   N, eps, gamma, inv_n, head);
        else
          flow_train_split_kernel<0><<<grids, nw * 32, bytes, st>>>(d, packed, tables, x, y, gz, gld, gx, partials, loss_acc,
                                                                   N, eps, gamma, inv_n, head);
        CNF_CHECK_CUDA(cudaGetLastError());
        return CNF_OK;
      }
    }
  }
  // hidden layers of up to 256 units, the widest of at least 64: the same tiles with the activations exchanged through
  // shared memory (narrower nets are quicker with one thread per sample; "1" forces, "0" disables: experiments)
  {
    const char* sw = cnf_switch(CNF_SW_DEEP_TRAIN);
    // (also single-hidden-layer nets that did not fit the plan above with its staged weights, on small batches:
    //  K=30 / hidden 256 at N=10,000: 2.3 -> 0.8 ms per step)
    bool fits = d.m >= 2 || (d.m == 1 && N <= 65536);
    for (int j = 0; j < d.m; ++j) fits = fits && d.Hp[j] <= 256;
    const bool want = sw ? atoi(sw) != 0 : d.Hmax >= 64;
    if (fits && want) {
      const int nw = d.Hmax / CH;
      const size_t bytes = (size_t)make_deep(d, nw).total * sizeof(float);
      if ((long long)bytes <= g_max_smem - 1024) {
        const int64_t nts = (N + SPL_TS - 1) / SPL_TS;
        int per_sm = (int)(g_max_smem / (bytes + 2048));
        const int by_threads = 2048 / (nw * 32), by_regs = 65536 / (128 * nw * 32);
        per_sm = per_sm < by_threads ? per_sm : by_threads;
        per_sm = per_sm < by_regs ? per_sm : by_regs;
        if (per_sm < 1) per_sm = 1;
        int64_t cap = (int64_t)g_num_sms * per_sm;
        if (cap > d.grad_rows) cap = d.grad_rows;
        const int gridd = (int)(nts < cap ? nts : cap);
        if ((rc = set_smem(flow_train_deep_kernel, bytes))) return rc;
        if ((rc = clear_rows(gridd))) return rc;
        flow_train_deep_kernel<<<gridd, nw * 32, bytes, st>>>(d, packed, tables, x, y, gz, gld, gx, partials, loss_acc, N,
                                                             eps, gamma, inv_n, head);
        CNF_CHECK_CUDA(cudaGetLastError());
        return CNF_OK;
      }
    }
  }
  LaunchCfg c;
  c.nt = 0; c.spt = 1; c.ws = false; c.wl = 0; c.smem = 0;
  rc = choose_cfg(d, true, &c);
  // single-hidden-layer nets whose full plan only fits narrow tiles (or does not fit at all): lean kernel
  const char* force_lean = cnf_switch(CNF_SW_FORCE_LEAN);      // experiment switch: "<nt>" forces the lean kernel with that tile
  if (d.m == 1 && (rc != CNF_OK || c.nt * c.spt < 128 || force_lean) && !cnf_switch(CNF_SW_NO_LEAN_TRAIN)) {
    for (int nt = force_lean ? atoi(force_lean) : 256; nt >= 64; nt >>= 1) {
      const size_t bytes = (size_t)make_lean(d, nt + 4).total * sizeof(float);
      if ((long long)bytes > g_max_smem - 1024) continue;
      const int64_t ntl = (N + nt - 1) / nt;
      int per_sm = (int)(g_max_smem / (bytes + 2048));
      const int max_threads = 2048 / nt;
      per_sm = per_sm < 1 ? 1 : (per_sm > max_threads ? max_threads : per_sm);
      if (per_sm > 8) per_sm = 8;
      const int64_t capl = (int64_t)g_num_sms * per_sm;
      const int gridl = (int)(ntl < capl ? ntl : capl);
      if ((rc = set_smem(flow_train_lean_kernel<1>, bytes))) return rc;
      if ((rc = clear_rows(gridl))) return rc;
      flow_train_lean_kernel<1><<<gridl, nt, bytes, st>>>(d, packed, tables, x, y, gz, gld, gx, partials, loss_acc, N, eps,
                                                          gamma, inv_n, head);
      CNF_CHECK_CUDA(cudaGetLastError());
      return CNF_OK;
    }
  }
  if (rc) return rc;
  const int64_t ntiles = (N + c.nt * c.spt - 1) / (c.nt * c.spt);
  int ctas_per_sm = (int)(g_max_smem / (c.smem + 2048));
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  if (ctas_per_sm > 2) ctas_per_sm = 2;
  const int64_t cap = (int64_t)g_num_sms * ctas_per_sm;
  const int grid = (int)(ntiles < cap ? ntiles : cap);
  if ((rc = clear_rows(grid))) return rc;
#define LAUNCH_TRAIN(SPT, WS)                                                                          \
  do {                                                                                                 \
    if ((rc = set_smem(flow_train_kernel<SPT, WS>, c.smem))) return rc;                                \
    flow_train_kernel<SPT, WS><<<grid, c.nt, c.smem, st>>>(d, packed, tables, x, y, gz, gld, gx, partials, \
                                                           loss_acc, N, eps, gamma, inv_n, head, c.wl); \
  } while (0)
  if (c.spt == 2 && c.ws) LAUNCH_TRAIN(2, true);
  else if (c.spt == 2) LAUNCH_TRAIN(2, false);
  else if (c.ws) LAUNCH_TRAIN(1, true);
  else LAUNCH_TRAIN(1, false);
#undef LAUNCH_TRAIN
  CNF_CHECK_CUDA(cudaGetLastError());
  return CNF_OK;
}
