// bf16 tensor-core kernel of the coupling-flow stack for sm_100a: tcgen05.mma with TMEM
// accumulators, one persistent CTA per SM, forward and inverse, fp32 coupling arithmetic.
//
// Shape coverage (cnf_tc_supported): one hidden layer per conditioner, K <= 16 classes
// (d1+1 <= 16, d0 <= 8), N1 = n_nets*pad16(H) <= 256.  Everything else runs on the fp32 kernel.
//
// Per coupling layer and per tile of 128 samples (TMEM lane = sample row), per conditioner net (s, then t):
//   GEMM1  D1[128 x Hp] = A1[128 x 16] . B1[net]^T     tcgen05.mma SS, bf16 in, fp32 out (TMEM)
//          A1 row = (u_0..u_{d1-1}, 1, 0..): the conditioning logits plus a constant-one column
//          that folds the first-layer bias; B1 = first Linear of the s-net and of the t-net one
//          after the other (flows/utils.py:26-31, flows/flows.py:105).
//   EPI1   h = relu(D1) -> bf16, written back to TMEM in place (A2 aliases the low half of D1),
//          released to the issuer 64 hidden units at a time.
//   GEMM2  D2[128 x 16] += A2[128 x Hp] . B2[net]^T     tcgen05.mma TS (A from TMEM), Hp/16 k-steps
//          B2 = block-diagonal last Linears: columns 0..7 <- first present net, 8..15 <- second;
//          the other net's GEMM1 is issued right behind (the in-order tensor pipe resolves the WAR on D1).
// then
//   EPI2   s,t = D2 + b2 (fp32); y = x*exp(s)+t, ld += sum s   (flows/flows.py:107-109); inverse
//          x = (y-t)*exp(-s), ld -= sum s (flows/flows.py:121-125).
// A CTA keeps THREE tiles in flight (TMEM slots of 160 columns: D1 128 + D2 16; one epilogue warpgroup and
// one MMA-issuer warp each) so the tensor pipe works on one tile while CUDA cores run the other tiles'
// epilogues; all packed weights of all layers stay resident in shared memory in the canonical
// no-swizzle K-major UMMA layout.  TAPE (training): the epilogue also writes, per layer, the pre-layer
// values of the transformed logits and the scale-net outputs for cnf_flow_tcb.cu.  SH: compile-time
// shape of BASELINE configs C2/C3/C5 (see the kernel).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "cnf_common.h"
#include "cnf_tc_dims.h"
#include "cnf_tc_ptx.cuh"
#include "cnf_metrics_dev.cuh"

int cnf_pack_bf16(const float* flat, const int32_t* gather, void* packed, int n, cudaStream_t st);
// wide variant (cnf_flow_tcw.cu): used when the resident-weight kernel does not cover the shape
bool cnf_tcw_supported(const CnfDims& d);
long long cnf_tcw_blob_bytes(const CnfDims& d);
long long cnf_tcw_gather_len(const CnfDims& d);
int cnf_tcw_plan_build(const CnfDims& d, int32_t* g);
int cnf_tcw_pack(const CnfDims& d, const float* flat, const int32_t* gather_tc, void* packed_tc, cudaStream_t st);
int cnf_tcw_apply(const CnfDims& d, const void* packed_tc, const int32_t* tables, const float* x, float* z,
                  float* logdet, int64_t N, int inverse, cudaStream_t st);

namespace {

// Timing-only experiment switches (wrong results): bit 0 = EPI1 does no work, bit 1 = no EPI2 math, bit 2 = no tile
// load / store, bit 3 = no GEMM2 MMAs.
#ifndef CNF_TC_EXP
#define CNF_TC_EXP 0
#endif
#ifndef CNF_TC_SLOTS
#define CNF_TC_SLOTS 3
#endif
constexpr int TC_SLOTS = CNF_TC_SLOTS;             // tiles in flight per CTA
constexpr int TC_THREADS = 128 + 128 * TC_SLOTS;  // warps 0..2 MMA issuers, warp 3 TMEM allocator, then one epilogue warpgroup per slot
constexpr int SLOT_COLS = 160;                    // TMEM columns per slot: D1 (<=128) + D2 (16), padded

constexpr int A1_BYTES = TILE_M * 32;   // 128 rows x 16 bf16
constexpr int LBO1 = 128, SBO1 = 256;   // A1 / B1: k-halves adjacent, 8-row groups 256 B apart
constexpr int LBO2 = 256, SBO2 = 128;   // B2 per k-step: two 8-row groups adjacent, k-halves 256 B apart


}  // namespace

bool cnf_tc_dims(const CnfDims& d, TcDims* t) {
  if (d.m != 1 || d.n_nets < 1) return false;
  if (d.d1 + 1 > 16 || d.d0 > 8) return false;
  const int N1 = d.n_nets * d.Hp[0];
  if (d.Hp[0] < 16 || d.Hp[0] > 128) return false;   // one net's hidden units fill at most 128 TMEM columns
  t->K = d.K; t->L = d.L; t->d0 = d.d0; t->d1 = d.d1; t->Hp = d.Hp[0]; t->nets = d.nets; t->n_nets = d.n_nets;
  t->N1 = N1;
  t->b_layer_bytes = N1 * 16 * 2;
  t->b1_off = 0;
  t->b2_off = d.L * t->b_layer_bytes;
  t->bias_off = 2 * d.L * t->b_layer_bytes;
  t->n_bf16 = t->bias_off / 2;
  t->n_f32 = d.L * 16;
  t->blob_bytes = t->bias_off + t->n_f32 * 4;
  t->tab_pi = d.tab_pi; t->tab_cond = d.tab_cond; t->tab_trans = d.tab_trans; t->n_tables = d.n_tables;
  int off = (t->blob_bytes + 127) / 128 * 128;
  t->sm_tab = off; off += (d.n_tables * 4 + 127) / 128 * 128;
  t->sm_slot = off;
  {
    const int tile_bytes = (d.K * TILE_M * 4 + 127) / 128 * 128;
    t->sm_act = A1_BYTES;                       // offsets inside one slot
    t->sm_raw_in = t->sm_act + tile_bytes;      // cp.async landing zone of the NEXT tile (row-major)
    t->sm_raw_out = t->sm_raw_in + tile_bytes;  // row-major staging of the finished tile
    t->sm_slot_stride = t->sm_raw_out + tile_bytes;
  }
  off += TC_SLOTS * t->sm_slot_stride;
  t->sm_bar = off; off += 128;
  t->sm_tail = off;          // fused statistics tail (cnf_flow_predict): the launcher adds its bytes to sm_total
  t->sm_total = off;
  return t->sm_total <= 227 * 1024;
}

namespace {

// ------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------
// EPI selects how EPI1 turns fp32 hidden units into the bf16 A operand of GEMM2:
//   0: cvt.rn.relu.bf16x2.f32 (round to nearest; F2FP runs on the quarter-rate XU pipe)
//   1: byte-permute truncation + packed bf16 max(.,0); the mean shrink of truncation
//      (E[ulp loss] ~ 0.72 * 2^-9 relative) is compensated on the fp32 GEMM2 output.
template <int EPI>
__device__ __forceinline__ uint32_t pack_hidden(uint32_t lo, uint32_t hi) {
  if (EPI == 0) return pack_relu_bf16(__uint_as_float(lo), __uint_as_float(hi));
  const uint32_t pk = __byte_perm(lo, hi, 0x7632);
  __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(&pk);
  const uint32_t zero = 0u;
  v = __hmax2(v, *reinterpret_cast<const __nv_bfloat162*>(&zero));
  return *reinterpret_cast<const uint32_t*>(&v);
}

// SH: 1 = the shape constants of BASELINE configs C2/C3/C5 (K = 10: d0 = d1 = 5, 128 hidden units, both nets)
// are compile-time, which removes the guards and index arithmetic of the generic loops (about a quarter
// of the epilogue's instructions); 0 = every shape the kernel covers, read from TcDims.
// TAIL (0 / CNF_METRICS_LOGITS / CNF_METRICS_CALIBRATED): fused Calibrator.predict tail / ECE-NLL-accuracy
// statistics on the finished tile and row-mean centring of the raw logits as a prologue (cnf_flow_predict).
template <int EPI, bool TAPE, int SH, int TAIL>
__global__ void __launch_bounds__(TC_THREADS, 1)
flow_tc_kernel(TcDims p, const uint8_t* __restrict__ blob, const int* __restrict__ tables,
               const float* __restrict__ xin, float* __restrict__ zout, float* __restrict__ logdet, int64_t N,
               int inverse, int io16, float* __restrict__ tape, CnfTail ta) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ double tail_red[32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  TailSmem tsm;
  BinCache cache; cache.bin = -1; cache.cnt = 0; cache.correct = 0; cache.sconf = 0.0;
  double a_nll = 0.0;
  unsigned a_correct = 0u, a_n = 0u;
  if (TAIL) {
    tsm = tail_carve(smem + p.sm_tail, ta.bins, p.K);
    tail_init(tsm, ta, p.K, tid, TC_THREADS);     // visible to the epilogue warps after the set-up barrier below
  }
  int* tab = reinterpret_cast<int*>(smem + p.sm_tab);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.sm_bar);
  // per slot s: a1_ready[s] (128 arrivals, once per layer)      epilogue -> MMA: A1 row block in smem
  //             d1_ready[s] (commit, once per net phase)        MMA -> epilogue: this net's D1 is complete
  //             a2_ready[2s+g] (128 arrivals, once per phase)   epilogue -> MMA: 64 more hidden columns in TMEM
  //             d2_ready[s] (commit, once per layer)            MMA -> epilogue: D2 complete
  // A waiter is never more than one phase behind any of them (see the hand-off order below).
  uint64_t* a1_ready = bars;
  uint64_t* d1_ready = bars + TC_SLOTS;
  uint64_t* d2_ready = bars + 2 * TC_SLOTS;
  uint64_t* a2_ready = bars + 3 * TC_SLOTS;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 5 * TC_SLOTS);

  // ---- one-time setup ---------------------------------------------------------------------
  {
    const uint4* src = reinterpret_cast<const uint4*>(blob);
    uint4* dst = reinterpret_cast<uint4*>(smem);
    for (int i = tid; i < p.blob_bytes / 16; i += TC_THREADS) dst[i] = __ldg(src + i);
    for (int i = tid; i < p.n_tables; i += TC_THREADS) tab[i] = tables[i];
  }
  if (tid == 0) {
    for (int s = 0; s < TC_SLOTS; ++s) {
      mbar_init(a1_ready + s, 128);
      mbar_init(d1_ready + s, 1);
      mbar_init(d2_ready + s, 1);
      mbar_init(a2_ready + 2 * s, 128);
      mbar_init(a2_ready + 2 * s + 1, 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 3) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_async_smem();          // the weight image was written through the generic proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int64_t ntiles = (N + TILE_M - 1) / TILE_M;
  const int G = gridDim.x;
  const int Hp = SH ? 128 : p.Hp;      // hidden units (TMEM columns of D1) per net phase
  const int D0 = SH ? 5 : p.d0, D1 = SH ? 5 : p.d1, KK = SH ? 10 : p.K;
  const int n_ph = SH ? 2 : p.n_nets;           // net phases per layer: s then t (or the single present net)
  const int n_grp = (Hp + 63) / 64;    // EPI1 releases the hidden units to GEMM2 in groups of 64 columns
  const int d2_col = 128;              // D2 sits above D1 inside the slot

  if (warp < TC_SLOTS) {
    // ================================ MMA issuers: warp s drives slot s ========================
    if (lane == 0) {
      const int s = warp;
      const int64_t first = blockIdx.x + (int64_t)s * G;
      const int64_t total = first < ntiles ? ((ntiles - first + TC_SLOTS * (int64_t)G - 1) / (TC_SLOTS * (int64_t)G)) * p.L : 0;
      const uint32_t idesc1 = make_idesc(Hp), idesc2 = make_idesc(16);
      const uint32_t smem_base = smem_u32(smem);
      const uint32_t tm = tmem_base + s * SLOT_COLS;
      const uint64_t ad = make_desc(smem_base + p.sm_slot + s * p.sm_slot_stride, LBO1, SBO1);
      int li = 0;
      uint32_t cnt = 0;                // net phases issued so far on this slot
      for (int64_t it = 0; it < total; ++it) {
        const int l = inverse ? p.L - 1 - li : li;
        // descriptors of this layer's B1 / B2 images; the start-address field (16-byte units) is the only
        // part that changes between nets and k-steps, so they are advanced by plain adds
        const uint64_t db1 = make_desc(smem_base + p.b1_off + l * p.b_layer_bytes, LBO1, SBO1);
        const uint64_t db2 = make_desc(smem_base + p.b2_off + l * p.b_layer_bytes, LBO2, SBO2);
        mbar_wait_backoff(a1_ready + s, (uint32_t)(it & 1));
        tc_fence_after();
#pragma unroll
        for (int ph = 0; ph < (SH ? 2 : 2); ++ph) {
          if (ph >= n_ph) break;
          // GEMM1 of this net: D1[128 x Hp] = A1 . B1[ph]^T   (rows ph*Hp.. of the B1 image)
          mma_ss(tm, ad, db1 + (uint64_t)(ph * (Hp / 8) * (SBO1 / 16)), idesc1, 0u);
          tc_commit(d1_ready + s);
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            if (g >= n_grp) break;
            mbar_wait_backoff(a2_ready + 2 * s + g, cnt & 1);
            tc_fence_after();
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {     // GEMM2 k-steps of this net accumulate into the shared 16-column D2
              const int j = 4 * g + jj;
              if (j < Hp / 16 && !(CNF_TC_EXP & 8))
                mma_ts(tm + d2_col, tm + j * 8, db2 + (uint64_t)((ph * (Hp / 16) + j) * 32), idesc2,
                       (ph > 0 || j > 0) ? 1u : 0u);
            }
          }
          ++cnt;
        }
        tc_commit(d2_ready + s);
        if (++li == p.L) li = 0;
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ================================ epilogue warpgroups =====================================
    const int slot = (warp - 4) >> 2;
    const int t = tid - 128 * (1 + slot);             // sample row inside the tile == TMEM lane
    uint8_t* a1 = smem + p.sm_slot + slot * p.sm_slot_stride;
    float* act = reinterpret_cast<float*>(a1 + p.sm_act);
    float* raw_in = reinterpret_cast<float*>(a1 + p.sm_raw_in);
    float* raw_out = reinterpret_cast<float*>(a1 + p.sm_raw_out);
    const float* bias = reinterpret_cast<const float*>(smem + p.bias_off);
    const uint32_t tm = tmem_base + slot * SLOT_COLS + ((uint32_t)((warp & 3) * 32) << 16);
    const int* pi_last = tab + p.tab_pi + p.L * p.K;
    uint32_t it = 0, cnt = 0;
    uint8_t* a1_row = a1 + (t >> 3) * SBO1 + (t & 7) * 16;
    const bool both = SH ? true : (p.nets == 3);
    const int K = KK, tile_elems = TILE_M * KK;
    // truncation shrinks every hidden unit by ~0.72*2^-9 on average; undo it on the fp32 output
    const float comp = (EPI == 1) ? 1.0f + 0.72f / 512.0f : 1.0f;
    const uint32_t one_bits = 0x3f80u;   // bf16 1.0
    // (s, f) walk of e = t + 128*i without integer division
    const int s0 = t / K, f0 = t - s0 * K, ds = TILE_M / K, df = TILE_M - ds * K;
    auto prefetch = [&](int64_t tile) {   // row-major copy of a FULL tile, 16 B per cp.async
      const float* gp = xin + tile * TILE_M * (int64_t)K;
      for (int c = t; c < tile_elems / 4; c += 128) cp_async16(raw_in + 4 * c, gp + 4 * c);
      cp_async_commit();
    };
    // second k-half of the A1 row is constant unless d1 >= 8: zero it once
    *reinterpret_cast<uint4*>(a1_row + LBO1) = make_uint4(0u, 0u, 0u, 0u);
    const int64_t tile0 = blockIdx.x + (int64_t)slot * G;
    const int64_t tstep = (int64_t)TC_SLOTS * G;
    bool prefetched = false;
    if (io16 && tile0 < ntiles && (tile0 + 1) * TILE_M <= N) { prefetch(tile0); prefetched = true; }
    for (int64_t tile = tile0; tile < ntiles; tile += tstep) {
      const int64_t base = tile * TILE_M;
      // ---- tile -> act[slot][sample] (transposing) --------------------------------------------
      if (prefetched) {
        cp_async_wait_all();
        wg_sync(slot);
        int s = s0, f = f0;
        for (int e = t; e < ((CNF_TC_EXP & 4) ? 0 : tile_elems); e += 128) {
          act[(inverse ? pi_last[f] : f) * TILE_M + s] = raw_in[e];
          s += ds; f += df;
          if (f >= K) { f -= K; ++s; }
        }
      } else {
        const float* gp = xin + base * K;
        const int64_t avail = (N - base) * (int64_t)K;
        int s = s0, f = f0;
        for (int e = t; e < tile_elems; e += 128) {
          const float v = (e < avail) ? __ldg(gp + e) : 0.f;
          act[(inverse ? pi_last[f] : f) * TILE_M + s] = v;
          s += ds; f += df;
          if (f >= K) { f -= K; ++s; }
        }
      }
      wg_sync(slot);
      if (TAIL && ta.center) {   // forward only (host-checked): act row f holds logical column f of sample t
        float* col = act + t;
        const float mean = numpy_row_mean([&](int j) -> float { return col[j * TILE_M]; }, K);
        for (int j = 0; j < K; ++j) col[j * TILE_M] = __fsub_rn(col[j * TILE_M], mean);
      }
      {
        const int64_t nxt = tile + tstep;
        prefetched = false;
        if (io16 && nxt < ntiles && (nxt + 1) * TILE_M <= N) { prefetch(nxt); prefetched = true; }
      }
      float ld = 0.f;
      for (int li = 0; li < p.L; ++li, ++it) {
        const int l = inverse ? p.L - 1 - li : li;
        const int* cond = tab + p.tab_cond + l * D1;
        const int* trans = tab + p.tab_trans + l * D0;
        // ---- A1 row: conditioning logits as bf16 + the constant one (independent loads) -------
        {
          float u[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) u[k] = (k < D1) ? act[cond[k] * TILE_M + t] : 0.f;
          uint4 lo;
          lo.x = pack_bf16(u[0], u[1]); lo.y = pack_bf16(u[2], u[3]);
          lo.z = pack_bf16(u[4], u[5]); lo.w = pack_bf16(u[6], u[7]);
          if (D1 < 8) {
            const uint32_t ob = one_bits << ((D1 & 1) * 16);
            const int wi = D1 >> 1;
            lo.x |= (wi == 0) ? ob : 0u; lo.y |= (wi == 1) ? ob : 0u;
            lo.z |= (wi == 2) ? ob : 0u; lo.w |= (wi == 3) ? ob : 0u;
          }
          *reinterpret_cast<uint4*>(a1_row) = lo;
          if (D1 >= 8) {      // K >= 15: second k-half carries u_8.. and the one
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = (k + 8 < D1) ? act[cond[k + 8] * TILE_M + t] : (k + 8 == D1 ? 1.f : 0.f);
            uint4 hi;
            hi.x = pack_bf16(v[0], v[1]); hi.y = pack_bf16(v[2], v[3]);
            hi.z = pack_bf16(v[4], v[5]); hi.w = pack_bf16(v[6], v[7]);
            *reinterpret_cast<uint4*>(a1_row + LBO1) = hi;
          }
        }
        fence_async_smem();
        tc_fence_before();
        mbar_arrive(a1_ready + slot);
        // pre-load what EPI2 needs while the tensor pipe works
        float xv[8];
        int ps[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          ps[q] = (q < D0) ? trans[q] * TILE_M + t : t;
          xv[q] = act[ps[q]];
        }
        const float* bl = bias + l * 16;
        // ---- EPI1 per net phase: relu + bf16 in place, released 64 columns at a time -----------
        for (int ph = 0; ph < n_ph; ++ph, ++cnt) {
          mbar_wait(d1_ready + slot, cnt & 1);
          tc_fence_after();
          if (CNF_TC_EXP & 1) {
            tc_fence_before();
            for (int c = 0; c < Hp; c += 64) mbar_arrive(a2_ready + 2 * slot + (c >> 6));
            continue;
          }
          uint32_t ra[32], rb[32], pk[16];
          tmem_ld32(tm, ra);
          for (int c = 0; c < Hp; c += 64) {
            tmem_wait_ld32(ra);
            if (c + 32 < Hp) tmem_ld32(tm + c + 32, rb);
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = pack_hidden<EPI>(ra[2 * i], ra[2 * i + 1]);
            tmem_st16(tm + c / 2, pk);
            if (c + 32 < Hp) {
              tmem_wait_ld32(rb);
              if (c + 64 < Hp) tmem_ld32(tm + c + 64, ra);
#pragma unroll
              for (int i = 0; i < 16; ++i) pk[i] = pack_hidden<EPI>(rb[2 * i], rb[2 * i + 1]);
              tmem_st16(tm + c / 2 + 16, pk);
            }
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(a2_ready + 2 * slot + (c >> 6));
          }
        }
        // ---- EPI2: coupling update in fp32 --------------------------------------------------
        mbar_wait(d2_ready + slot, it & 1);
        tc_fence_after();
        if (!(CNF_TC_EXP & 2)) {
          uint32_t r[16];
          float svs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          tmem_ld16(tm + d2_col, r);
          tmem_wait_ld16(r);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (q < D0) {
              // columns 0..7 belong to the first present net, 8..15 to the second
              const float first = fmaf(__uint_as_float(r[q]), comp, bl[q]);
              const float second = fmaf(__uint_as_float(r[8 + q]), comp, bl[8 + q]);
              const float sv = (SH || (p.nets & 1)) ? first : 0.f;
              const float tv = both ? second : ((p.nets & 2) ? first : 0.f);
              float yv;
              if (!inverse) { yv = xv[q] * tc_exp(sv) + tv; ld += sv; }
              else          { yv = (xv[q] - tv) * tc_exp(-sv); ld -= sv; }
              act[ps[q]] = yv;
              if (TAPE) svs[q] = sv;
            }
          }
          if (TAPE && base + t < N) {   // training: what autograd would save for this layer
            float4* tp = reinterpret_cast<float4*>(tape + ((size_t)l * N + (base + t)) * 16);
            tp[0] = make_float4(xv[0], xv[1], xv[2], xv[3]);
            tp[1] = make_float4(xv[4], xv[5], xv[6], xv[7]);
            tp[2] = make_float4(svs[0], svs[1], svs[2], svs[3]);
            tp[3] = make_float4(svs[4], svs[5], svs[6], svs[7]);
          }
        }
      }
      if (base + t < N && (!TAIL || logdet != nullptr)) logdet[base + t] = ld;
      if (TAIL && base + t < N) {   // the finished sample in logical class order, still in shared memory
        const float* col = act + t;
        auto get = [&](int jj) -> float { return col[(inverse ? jj : pi_last[jj]) * TILE_M]; };
        tail_row<TAIL == 0 ? CNF_METRICS_LOGITS : TAIL>(get, base + t, K, ta, tsm, cache, a_nll, a_correct, a_n);
      }
      // ---- act -> row-major staging -> coalesced global store ---------------------------------
      if ((!TAIL || zout != nullptr) && !(CNF_TC_EXP & 4)) {
        int s = s0, f = f0;
        const bool full = io16 && (base + TILE_M <= N);
        float* gp = zout + base * K;
        const int64_t avail = (N - base) * (int64_t)K;
        if (full) {
          wg_sync(slot);                       // every row of act is final
          for (int e = t; e < tile_elems; e += 128) {
            raw_out[e] = act[(inverse ? f : pi_last[f]) * TILE_M + s];
            s += ds; f += df;
            if (f >= K) { f -= K; ++s; }
          }
          wg_sync(slot);
          for (int c = t; c < tile_elems / 4; c += 128)
            *reinterpret_cast<float4*>(gp + 4 * c) = *reinterpret_cast<const float4*>(raw_out + 4 * c);
        } else {
          wg_sync(slot);
          for (int e = t; e < tile_elems; e += 128) {
            if (e < avail) gp[e] = act[(inverse ? f : pi_last[f]) * TILE_M + s];
            s += ds; f += df;
            if (f >= K) { f -= K; ++s; }
          }
          wg_sync(slot);
        }
      } else {
        wg_sync(slot);                         // the tail has read act: the next tile may overwrite it
      }
    }
  }
  // ---- teardown -----------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 3) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
  if (TAIL) stats_finish_block(a_nll, a_correct, a_n, cache, tsm.s_cnt, tsm.s_cor, tsm.s_conf, ta.bins, ta.acc, tail_red, tid, TC_THREADS);
}


}  // namespace

long long cnf_tc_blob_bytes(const cnf_flow_desc* desc, const CnfDims& d) {
  TcDims t;
  if (cnf_tc_dims(d, &t)) return (long long)t.blob_bytes;
  return cnf_tcw_blob_bytes(d);
}

// gather_tc: entry i < n_bf16 addresses bf16 element i of the B1/B2 image; the remaining n_f32
// entries address the fp32 last-layer biases.  -1 = zero.  Length = cnf_tc_gather_len().
extern "C" int cnf_tc_gather_len(const cnf_flow_desc* desc, int64_t* n) {
  CnfDims d; TcDims t;
  int rc = cnf_make_dims(desc, &d);
  if (rc) return rc;
  if (!n) { cnf_set_error("null out"); return CNF_E_ARG; }
  *n = cnf_tc_dims(d, &t) ? (int64_t)t.n_bf16 + t.n_f32 : (int64_t)cnf_tcw_gather_len(d);
  return CNF_OK;
}

extern "C" int cnf_plan_build_tc(const cnf_flow_desc* desc, int32_t* g) {
  CnfDims d; TcDims t;
  int rc = cnf_make_dims(desc, &d);
  if (rc) return rc;
  if (!g) { cnf_set_error("null output"); return CNF_E_ARG; }
  if (!cnf_tc_dims(d, &t)) return cnf_tcw_plan_build(d, g);
  const int K = d.K, half = K / 2, H = d.H[0], Hp = d.Hp[0];
  for (int i = 0; i < t.n_bf16 + t.n_f32; ++i) g[i] = -1;
  // canonical flat offsets of one net: W0 [H,K], b0 [H], W1 [K,H], b1 [K]
  const long long net_sz = (long long)H * K + H + (long long)K * H + K;
  for (int l = 0; l < d.L; ++l) {
    int slot = 0;
    for (int net = 0; net < 2; ++net) {
      if (!(d.nets & (1 << net))) continue;
      const long long base = ((long long)l * d.n_nets + slot) * net_sz;
      const long long w0 = base, b0 = base + (long long)H * K, w1 = b0 + H, b1 = w1 + (long long)K * H;
      // B1 image: element (n, k), n = slot*Hp + h
      int32_t* B1 = g + (t.b1_off + l * t.b_layer_bytes) / 2;
      for (int h = 0; h < H; ++h) {
        const int n = slot * Hp + h;
        for (int k = 0; k <= d.d1; ++k) {
          const int byte = (n / 8) * SBO1 + (k / 8) * LBO1 + (n % 8) * 16 + (k % 8) * 2;
          B1[byte / 2] = (int32_t)(k < d.d1 ? w0 + (long long)h * K + half + k : b0 + h);
        }
      }
      // B2 image: element (n2, kk), n2 = slot*8 + q, kk = slot*Hp + h
      int32_t* B2 = g + (t.b2_off + l * t.b_layer_bytes) / 2;
      for (int q = 0; q < d.d0; ++q) {
        const int n2 = slot * 8 + q;
        for (int h = 0; h < H; ++h) {
          const int kk = slot * Hp + h;
          const int byte = (kk / 16) * 512 + ((kk % 16) / 8) * LBO2 + (n2 / 8) * SBO2 + (n2 % 8) * 16 + (kk % 8) * 2;
          B2[byte / 2] = (int32_t)(w1 + (long long)q * H + h);
        }
        g[t.n_bf16 + l * 16 + n2] = (int32_t)(b1 + q);
      }
      ++slot;
    }
  }
  return CNF_OK;
}

namespace {
__global__ void gather_f32_kernel(const float* __restrict__ flat, const int* __restrict__ gather, float* __restrict__ out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const int gi = gather[i]; out[i] = gi >= 0 ? flat[gi] : 0.f; }
}
}  // namespace

extern "C" int cnf_pack_weights_tc(const cnf_flow_desc* desc, const float* flat, const int32_t* gather_tc,
                                   void* packed_tc, void* stream) {
  CnfDims d; TcDims t;
  int rc = cnf_make_dims(desc, &d);
  if (rc) return rc;
  if (!flat || !gather_tc || !packed_tc) { cnf_set_error("cnf_pack_weights_tc: null pointer"); return CNF_E_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  if (!cnf_tc_dims(d, &t)) return cnf_tcw_pack(d, flat, gather_tc, packed_tc, st);
  if ((rc = cnf_pack_bf16(flat, gather_tc, packed_tc, t.n_bf16, st))) return rc;
  gather_f32_kernel<<<(t.n_f32 + 127) / 128, 128, 0, st>>>(flat, gather_tc + t.n_bf16,
                                                           reinterpret_cast<float*>((uint8_t*)packed_tc + t.bias_off), t.n_f32);
  CNF_CHECK_CUDA(cudaGetLastError());
  return CNF_OK;
}

int cnf_tc_apply_ex(const cnf_flow_desc* desc, const void* packed_tc, const int32_t* tables, const float* x, float* z,
                    float* logdet, float* tape, int64_t N, int inverse, const CnfTail* tail, cudaStream_t st);

int cnf_tc_apply(const cnf_flow_desc* desc, const void* packed_tc, const int32_t* tables, const float* x, float* z,
                 float* logdet, int64_t N, int inverse, cudaStream_t st) {
  return cnf_tc_apply_ex(desc, packed_tc, tables, x, z, logdet, nullptr, N, inverse, nullptr, st);
}

// tape (optional, resident-weight kernel only): float32 [L, N, 16]; per layer and sample the pre-layer
// values of the transformed slots (0..7) and the scale-net outputs s (8..15).
int cnf_tc_apply_tape(const cnf_flow_desc* desc, const void* packed_tc, const int32_t* tables, const float* x, float* z,
                      float* logdet, float* tape, int64_t N, int inverse, cudaStream_t st) {
  return cnf_tc_apply_ex(desc, packed_tc, tables, x, z, logdet, tape, N, inverse, nullptr, st);
}

// Fused pass of cnf_flow_predict on the tensor-core path; CNF_E_UNSUPPORTED for shapes only the
// streamed-weight kernel covers (the caller then composes the two-pass form).
int cnf_tc_predict(const cnf_flow_desc* desc, const void* packed_tc, const int32_t* tables, const float* x, float* z,
                   float* logdet, int64_t N, const CnfTail& ta, cudaStream_t st) {
  return cnf_tc_apply_ex(desc, packed_tc, tables, x, z, logdet, nullptr, N, 0, &ta, st);
}

int cnf_tc_apply_ex(const cnf_flow_desc* desc, const void* packed_tc, const int32_t* tables, const float* x, float* z,
                    float* logdet, float* tape, int64_t N, int inverse, const CnfTail* tail, cudaStream_t st) {
  CnfDims d; TcDims t;
  int rc = cnf_make_dims(desc, &d);
  if (rc) return rc;
  if (N == 0) return CNF_OK;
  if (!packed_tc || !tables || !x || N < 0 || (!tail && (!z || !logdet))) { cnf_set_error("null pointer / negative N"); return CNF_E_ARG; }
  if (!cnf_tc_dims(d, &t)) {
    if (tape) { cnf_set_error("the training tape is only produced by the resident-weight tensor-core kernel"); return CNF_E_UNSUPPORTED; }
    if (tail) { cnf_set_error("the fused predict tail is not available on the streamed-weight tensor-core kernel"); return CNF_E_UNSUPPORTED; }
    return cnf_tcw_apply(d, packed_tc, tables, x, z, logdet, N, inverse, st);
  }
  if (tail) {
    t.sm_total += cnf_tail_smem_bytes(tail->bins, d.K);
    if (t.sm_total > 227 * 1024) { cnf_set_error("fused tail: %d bins do not fit shared memory", tail->bins); return CNF_E_SMEM; }
  }
  CnfDevInfo di;
  if ((rc = cnf_dev_info(&di))) return rc;
  const int g_tc_sms = di.sms;
  const int64_t ntiles = (N + TILE_M - 1) / TILE_M;
  const int grid = (int)(ntiles < g_tc_sms ? ntiles : g_tc_sms);
  // 16-byte tile I/O needs 16 B-aligned pointers and whole tiles that are a multiple of 16 B
  const int io16 = (((uintptr_t)x | (uintptr_t)z) % 16 == 0 && (TILE_M * d.K) % 4 == 0) ? 1 : 0;
  int epi = 0;
  if (const char* v = cnf_switch(CNF_SW_TC_EPI)) epi = atoi(v);   // 0: round-to-nearest F2FP, 1: truncate+compensate
  const int sh = (d.K == 10 && t.Hp == 128 && d.nets == 3 && !cnf_switch(CNF_SW_TC_GENERIC)) ? 1 : 0;
  const CnfTail ta = tail ? *tail : CnfTail();
#define LAUNCH_TC_S(E, T, S, M)                                                                                 \
  do {                                                                                                          \
    if ((rc = cnf_kernel_smem(flow_tc_kernel<E, T, S, M>, t.sm_total))) return rc;                                          \
    flow_tc_kernel<E, T, S, M><<<grid, TC_THREADS, t.sm_total, st>>>(t, (const uint8_t*)packed_tc, tables, x, z, logdet, N, \
                                                                     inverse, io16, tape, ta);                  \
  } while (0)
#define LAUNCH_TC(E, T)                                                                                         \
  do {                                                                                                          \
    if (sh) LAUNCH_TC_S(E, T, 1, 0); else LAUNCH_TC_S(E, T, 0, 0);                                              \
  } while (0)
  if (tail) {
    if (tail->mode == CNF_METRICS_LOGITS) { if (sh) LAUNCH_TC_S(0, false, 1, CNF_METRICS_LOGITS); else LAUNCH_TC_S(0, false, 0, CNF_METRICS_LOGITS); }
    else { if (sh) LAUNCH_TC_S(0, false, 1, CNF_METRICS_CALIBRATED); else LAUNCH_TC_S(0, false, 0, CNF_METRICS_CALIBRATED); }
  } else if (tape) { if (epi == 0) LAUNCH_TC(0, true); else LAUNCH_TC(1, true); }
  else      { if (epi == 0) LAUNCH_TC(0, false); else LAUNCH_TC(1, false); }
#undef LAUNCH_TC_S
#undef LAUNCH_TC
  CNF_CHECK_CUDA(cudaGetLastError());
  return CNF_OK;
}
