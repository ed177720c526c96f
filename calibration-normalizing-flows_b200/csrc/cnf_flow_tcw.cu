// "Wide" bf16 tensor-core kernel of the coupling-flow stack for sm_100a (config C4 class:
// K up to 126 classes, hidden widths beyond what shared memory can hold for all layers).
//
// Same math as cnf_flow_tc.cu; differences:
//   * GEMM1 has K1 = pad16(d1+1) <= 64 contraction columns (K1/16 k-steps), GEMM2 writes one
//     accumulator per net, D2_s / D2_t [128 x N2], N2 = pad16(d0) <= 64.
//   * The hidden layer is processed in blocks of 128 units; one "phase" = (net, block):
//       GEMM1 block -> D1[128 x 128] (TMEM) -> EPI1 relu+bf16 in place -> GEMM2 accumulates into D2_net.
//   * Weights do not stay resident: every phase's B1 / B2 block (256*(K1+N2) bytes, contiguous in
//     the packed blob) is streamed into a 2-stage shared-memory ring by one TMA bulk copy
//     (cp.async.bulk + mbarrier complete_tx), issued by a producer warp one phase ahead.
//   * Two tiles per CTA run in lockstep over the phases so that a weight stage is consumed by
//     both before it is recycled; with >= 256 cycles of tensor work per phase and tile the lockstep
//     costs little.
// Reference arithmetic: flows/flows.py:101-126, flows/utils.py:26-31 (see cnf_flow_tc.cu).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdlib>

#include "cnf_common.h"
// e^s as ex2.approx(s * log2 e) in this kernel: 50 outputs per sample and layer make the coupling epilogue its
// largest serial section (+2.6 % at C4); relative error ~1e-6, far inside the bf16 path's stated 1e-2
#ifndef CNF_TC_FASTEXP
#define CNF_TC_FASTEXP 1
#endif
#include "cnf_tc_ptx.cuh"

int cnf_pack_bf16(const float* flat, const int32_t* gather, void* packed, int n, cudaStream_t st);

// Timing-only experiment switches (wrong results; profiles/microbench/build_variant.sh): bit 0 = weights are not
// streamed after the first two phases, bit 1 = EPI1 does no work, bit 2 = no tile load / store,
// bit 3 = no EPI2 math, bit 4 = no A1 build; bit 5 (results stay right) = the scale half of EPI2 stays at the layer boundary.
#ifndef CNF_TCW_STAGGER
#define CNF_TCW_STAGGER 0
#endif
#ifndef CNF_TCW_EXP
#define CNF_TCW_EXP 0
#endif

namespace {

constexpr int W_THREADS = 384;   // warps 0 / 3 MMA issuers (slot 0 / 1), 1 weight producer, 2 TMEM allocator, 4-7 / 8-11 epilogue
constexpr int HB = 128;          // hidden units per phase
constexpr int ACT_LD = TILE_M + 1;   // act[slot][sample] row stride: odd, so the transposing tile load / store is bank-conflict free
constexpr int W_LBO1 = 2048, W_SBO1 = 128;   // A1 / B1 block: [k-block][row-block] 128-byte core matrices
constexpr int W_SBO2 = 128;                  // B2 block per k-step: [k-half][row-block]

struct TcwDims {
  int K, L, d0, d1, H, nets, n_nets, K1, N2, n_blk, n_ph;
  int phase_bytes, b1_bytes;        // one phase of weights: B1 block then B2 block
  int bias_off, n_bf16, n_f32, blob_bytes;
  int tab_pi, tab_cond, tab_trans, n_tables;
  int sm_ring, sm_bias, sm_tab, sm_tp, sm_slot, sm_slot_stride, sm_act, sm_bar, sm_total;
};

bool tcw_dims(const CnfDims& d, TcwDims* t) {
  if (d.m != 1 || d.n_nets < 1) return false;
  if (d.d1 + 1 > 64 || d.d0 > 64) return false;
  if (d.K * ACT_LD > 65535) return false;            // act offsets are kept as uint16
  t->K = d.K; t->L = d.L; t->d0 = d.d0; t->d1 = d.d1; t->H = d.H[0]; t->nets = d.nets; t->n_nets = d.n_nets;
  t->K1 = cnf_round_up(d.d1 + 1, 16);
  t->N2 = cnf_round_up(d.d0, 16);
  t->n_blk = (d.H[0] + HB - 1) / HB;
  t->n_ph = t->n_blk * d.n_nets;
  t->b1_bytes = HB * t->K1 * 2;
  t->phase_bytes = t->b1_bytes + t->N2 * HB * 2;
  const long long wbytes = (long long)d.L * t->n_ph * t->phase_bytes;
  if (wbytes > (1ll << 30)) return false;
  t->bias_off = (int)wbytes;
  t->n_bf16 = (int)(wbytes / 2);
  t->n_f32 = d.L * 2 * t->N2;
  t->blob_bytes = t->bias_off + t->n_f32 * 4;
  t->tab_pi = d.tab_pi; t->tab_cond = d.tab_cond; t->tab_trans = d.tab_trans; t->n_tables = d.n_tables;
  int off = 0;
  t->sm_ring = off; off += 2 * t->phase_bytes;
  t->sm_bias = off; off += (t->n_f32 * 4 + 127) / 128 * 128;
  t->sm_tab = off; off += (d.n_tables * 4 + 127) / 128 * 128;
  t->sm_tp = off; off += (d.L * t->N2 * 2 + 127) / 128 * 128;   // act offsets (uint16) of the transformed slots, padded to N2
  t->sm_slot = off;
  t->sm_act = HB * t->K1 * 2;                       // A1 tile first, then the fp32 tile
  t->sm_slot_stride = t->sm_act + (d.K * ACT_LD * 4 + 127) / 128 * 128;
  off += 2 * t->sm_slot_stride;
  t->sm_bar = off; off += 128;
  t->sm_total = off;
  return t->sm_total <= 227 * 1024;
}

// SH = 1: compile-time shape of BASELINE config C4 (K = 100: d0 = d1 = 50, K1 = N2 = 64; 512 hidden
// units = 4 blocks; both nets = 8 phases); SH = 0: every covered shape, read from TcwDims.
template <int SH>
__global__ void __launch_bounds__(W_THREADS, 1)
flow_tcw_kernel(TcwDims p, const uint8_t* __restrict__ blob, const int* __restrict__ tables,
                const float* __restrict__ xin, float* __restrict__ zout, float* __restrict__ logdet, int64_t N,
                int inverse) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int* tab = reinterpret_cast<int*>(smem + p.sm_tab);
  float* bias = reinterpret_cast<float*>(smem + p.sm_bias);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.sm_bar);
  uint64_t* w_full = bars;          // [2] producer (TMA complete_tx) -> MMA: weight stage landed
  uint64_t* w_empty = bars + 2;     // [2] MMA (one commit per slot) -> producer: stage consumed by both tiles
  uint64_t* a1_ready = bars + 4;    // [2] epilogue -> MMA, once per layer
  uint64_t* d1_ready = bars + 6;    // [2] MMA -> epilogue, once per phase
  uint64_t* d2_ready = bars + 8;    // [2] MMA -> epilogue, once per layer
  uint64_t* a2_ready = bars + 10;   // [2 slots][2 groups] epilogue -> MMA, once per phase
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 14);

  for (int i = tid; i < p.n_tables; i += W_THREADS) tab[i] = tables[i];
  // EPI2 needs no guards: a padded output (q >= d0: zero weights and biases, so s = t = 0 exactly) "updates" the
  // layer's first conditioning slot, which no real output touches, with x * e^0 + 0 = x
  unsigned short* tp = reinterpret_cast<unsigned short*>(smem + p.sm_tp);
  for (int i = tid; i < p.L * p.N2; i += W_THREADS) {
    const int l = i / p.N2, q = i - l * p.N2;
    tp[i] = (unsigned short)((q < p.d0 ? tables[p.tab_trans + l * p.d0 + q] : tables[p.tab_cond + l * p.d1]) * ACT_LD);
  }
  {
    const float* gb = reinterpret_cast<const float*>(blob + p.bias_off);
    for (int i = tid; i < p.n_f32; i += W_THREADS) bias[i] = __ldg(gb + i);
  }
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(w_full + s, 1); mbar_init(w_empty + s, 2);
      mbar_init(a1_ready + s, 128); mbar_init(d1_ready + s, 1); mbar_init(d2_ready + s, 1);
      mbar_init(a2_ready + 2 * s, 128); mbar_init(a2_ready + 2 * s + 1, 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int64_t ntiles = (N + TILE_M - 1) / TILE_M;
  const int G = gridDim.x;
  // tiles of this CTA: cta, cta+G, ...; slot s takes every other one; both slots advance in rounds
  const int64_t mine = blockIdx.x < ntiles ? (ntiles - blockIdx.x + G - 1) / G : 0;
  const int64_t nt[2] = {(mine + 1) / 2, mine / 2};
  const int64_t rounds = nt[0];
  const int n_ph = SH ? 8 : p.n_ph, n_blk = SH ? 4 : p.n_blk;
  const int K1 = SH ? 64 : p.K1, N2 = SH ? 64 : p.N2, D1 = SH ? 50 : p.d1, KK = SH ? 100 : p.K;
  const int NETS = SH ? 3 : p.nets;
  const int k1_steps = K1 / 16;
  const int d2_col0 = 128;

  if (warp == 1) {
    // ================================ weight producer ==========================================
    if (lane == 0) {
      uint32_t g = 0;
      for (int64_t r = 0; r < rounds; ++r)
        for (int li = 0; li < p.L; ++li) {
          const int l = inverse ? p.L - 1 - li : li;
          for (int ph = 0; ph < n_ph; ++ph, ++g) {
            const int st = g & 1;
            if ((CNF_TCW_EXP & 1) && g >= 2) continue;
            if (g >= 2) mbar_wait(w_empty + st, ((g >> 1) - 1) & 1);
            mbar_expect_tx(w_full + st, (uint32_t)p.phase_bytes);
            bulk_copy_g2s(smem + p.sm_ring + st * p.phase_bytes, blob + ((size_t)l * n_ph + ph) * p.phase_bytes,
                          (uint32_t)p.phase_bytes, w_full + st);
          }
        }
    }
    __syncwarp();
  } else if (warp == 0 || warp == 3) {
    // ================================ MMA issuers: warp 0 -> slot 0, warp 3 -> slot 1 ===============
    // One thread per slot: a single thread issuing both slots' 24 MMAs per phase (descriptor arithmetic
    // included) was the kernel's critical path.  Both walk the same sequence of weight stages; a stage is
    // released when both have committed (or, for a slot without a tile in the last round, arrived).
    if (lane == 0) {
      const int s = warp == 0 ? 0 : 1;
      const uint32_t idesc1 = make_idesc(HB), idesc2 = make_idesc(N2);
      const uint32_t smem_base = smem_u32(smem);
      const uint32_t lbo2 = (uint32_t)N2 * 16;
      const uint32_t tm = tmem_base + s * 256;
      const uint64_t a1d = make_desc(smem_base + p.sm_slot + s * p.sm_slot_stride, W_LBO1, W_SBO1);
      const uint64_t kstep1 = (uint64_t)(2 * W_LBO1 / 16);          // descriptor start-address units per GEMM1 k-step
      const uint64_t kstep2 = (uint64_t)(N2 * 32 / 16);
      uint32_t g = 0, lay_cnt = 0, ph_cnt = 0;
      auto gemm1 = [&](uint32_t gg) {
        const uint64_t b1d = make_desc(smem_base + p.sm_ring + (gg & 1) * p.phase_bytes, W_LBO1, W_SBO1);
        for (int j = 0; j < k1_steps; ++j) mma_ss(tm, a1d + j * kstep1, b1d + j * kstep1, idesc1, j > 0 ? 1u : 0u);
        tc_commit(d1_ready + s);
      };
      for (int64_t r = 0; r < rounds; ++r) {
        const bool has_tile = r < nt[s];
        for (int li = 0; li < p.L; ++li) {
          if (!has_tile) {           // keep the ring's arrival count
            for (int ph = 0; ph < n_ph; ++ph, ++g) {
              if ((CNF_TCW_EXP & 1) && g >= 2) continue;
              mbar_wait_backoff(w_full + (g & 1), (g >> 1) & 1);
              mbar_arrive(w_empty + (g & 1));
            }
            continue;
          }
          mbar_wait_backoff(a1_ready + s, lay_cnt & 1);
          ++lay_cnt;
          if (!(CNF_TCW_EXP & 1) || g < 2) mbar_wait_backoff(w_full + (g & 1), (g >> 1) & 1);
          tc_fence_after();
          gemm1(g);
          for (int ph = 0; ph < n_ph; ++ph, ++g) {
            const int st = g & 1;
            const int net = ph / n_blk, blk = ph - net * n_blk;
            const uint64_t b2d = make_desc(smem_base + p.sm_ring + st * p.phase_bytes + p.b1_bytes, lbo2, W_SBO2);
            const uint32_t d2 = tm + d2_col0 + net * N2;
#pragma unroll
            for (int grp = 0; grp < 2; ++grp) {
              mbar_wait_backoff(a2_ready + 2 * s + grp, ph_cnt & 1);
              tc_fence_after();
#pragma unroll
              for (int j = 4 * grp; j < 4 * grp + 4; ++j)
                mma_ts(d2, tm + j * 8, b2d + j * kstep2, idesc2, (blk > 0 || j > 0) ? 1u : 0u);
            }
            ++ph_cnt;
            // this slot's GEMM1 of the next phase goes in right behind its GEMM2 (D1 is free once the GEMM2
            // ahead of it in the pipe has read it): the next EPI1 overlaps the other slot's tensor work
            if (ph + 1 < n_ph) {
              if (!(CNF_TCW_EXP & 1) || g + 1 < 2) mbar_wait_backoff(w_full + ((g + 1) & 1), ((g + 1) >> 1) & 1);
              tc_fence_after();
              gemm1(g + 1);
            }
            tc_commit(w_empty + st);
          }
          tc_commit(d2_ready + s);
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ================================ epilogue warpgroups =====================================
    const int slot = (warp - 4) >> 2;
    const int t = tid - 128 * (1 + slot);
    uint8_t* a1 = smem + p.sm_slot + slot * p.sm_slot_stride;
    float* act = reinterpret_cast<float*>(a1 + p.sm_act);
    const uint32_t tm = tmem_base + slot * 256 + ((uint32_t)((warp & 3) * 32) << 16);
    const int* pi_last = tab + p.tab_pi + p.L * KK;
    uint8_t* a1_row = a1 + (t >> 3) * W_SBO1 + (t & 7) * 16;
    const int K = KK, tile_elems = TILE_M * KK;
    const int s0 = t / K, f0 = t - s0 * K, ds = TILE_M / K, df = TILE_M - ds * K;
    const bool both = (NETS == 3);
    uint32_t lay_cnt = 0, ph_cnt = 0;
    // EPI2 on outputs qc .. qc+15.  mode 0: the whole coupling update; 1: its scale half (x *= e^s, log-det);
    // 2: its shift half (x += t).  Eight outputs at a time, every shared-memory load of the group issued before
    // its first store: bias and act share an element type, so the compiler cannot move a later output's loads
    // above an earlier output's store itself, and one output at a time is a chain of dependent shared-memory
    // round trips (27 % of the kernel when it was written that way).  No guards: see tp.
    auto epi2_chunk = [&](const int qc, const int l, float& ld, const int mode) {
      const float* bl = bias + l * 2 * N2 + qc;
      const unsigned short* tpl = tp + l * N2 + qc;
      uint32_t r1[16], r2[16];
      if (mode != 2) tmem_ld16(tm + d2_col0 + qc, r1);
      if (mode == 2 || (mode == 0 && both)) tmem_ld16(tm + d2_col0 + N2 + qc, r2);
      if (mode != 2) tmem_wait_ld16(r1);
      if (mode == 2 || (mode == 0 && both)) tmem_wait_ld16(r2);
#pragma unroll
      for (int h = 0; h < 16; h += 8) {
        int ps[8];
        float xv[8], bs[8], bt[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) ps[i] = (int)tpl[h + i] + t;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          xv[i] = act[ps[i]];
          bs[i] = (mode != 2) ? bl[h + i] : 0.f;
          bt[i] = (mode == 2 || (mode == 0 && both)) ? bl[N2 + h + i] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (mode == 1) {
            const float sv = __uint_as_float(r1[h + i]) + bs[i];
            xv[i] *= tc_exp(sv); ld += sv;
          } else if (mode == 2) {
            xv[i] += __uint_as_float(r2[h + i]) + bt[i];
          } else {
            const float first = __uint_as_float(r1[h + i]) + bs[i];
            const float second = both ? __uint_as_float(r2[h + i]) + bt[i] : 0.f;
            const float sv = (NETS & 1) ? first : 0.f;
            const float tv = both ? second : ((NETS & 2) ? first : 0.f);
            if (!inverse) { xv[i] = xv[i] * tc_exp(sv) + tv; ld += sv; }
            else          { xv[i] = (xv[i] - tv) * tc_exp(-sv); ld -= sv; }
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) act[ps[i]] = xv[i];
      }
    };
#if CNF_TCW_STAGGER > 0
    if (slot > 0) { const long long c0 = clock64(); while (clock64() - c0 < CNF_TCW_STAGGER) {} }
#endif

    for (int64_t r = 0; r < nt[slot]; ++r) {
      const int64_t tile = blockIdx.x + (2 * r + slot) * (int64_t)G;
      const int64_t base = tile * TILE_M;
      if (!(CNF_TCW_EXP & 4)) {
        // the tile goes straight into the transposed layout with 4-byte cp.async copies (all of a thread's
        // tile_elems / 128 copies in flight at once; rows past N are zero-filled): staging it through registers
        // ten loads at a time cost 14 % of the kernel
        const float* gp = xin + base * K;
        const int64_t avail = (N - base) * (int64_t)K;
        const uint32_t act_s = smem_u32(act);
        int s = s0, f = f0;
        for (int e = t; e < tile_elems; e += 128) {
          const bool ok = e < avail;
          cp_async4_zfill(act_s + (uint32_t)((inverse ? pi_last[f] : f) * ACT_LD + s) * 4u, ok ? gp + e : gp, ok ? 4u : 0u);
          s += ds; f += df;
          while (f >= K) { f -= K; ++s; }
        }
        cp_async_commit();
        cp_async_wait_all();
      }
      wg_sync(slot);
      float ld = 0.f;
      for (int li = 0; li < p.L; ++li, ++lay_cnt) {
        const int l = inverse ? p.L - 1 - li : li;
        const int* cond = tab + p.tab_cond + l * D1;
        // ---- A1 row: K1 bf16 = conditioning logits, the constant one, zero padding ------------
        // (sixteen gathers per pass, all issued before the pass's two stores; for the compile-time shape the
        // k < D1 selects are resolved at compile time)
        auto a1_pass = [&](const int kb) {
          float u[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int k = kb * 8 + i;
            u[i] = (k < D1) ? act[cond[k] * ACT_LD + t] : (k == D1 ? 1.f : 0.f);
          }
          uint4 v, w;
          v.x = pack_bf16(u[0], u[1]); v.y = pack_bf16(u[2], u[3]);
          v.z = pack_bf16(u[4], u[5]); v.w = pack_bf16(u[6], u[7]);
          w.x = pack_bf16(u[8], u[9]); w.y = pack_bf16(u[10], u[11]);
          w.z = pack_bf16(u[12], u[13]); w.w = pack_bf16(u[14], u[15]);
          *reinterpret_cast<uint4*>(a1_row + kb * W_LBO1) = v;
          *reinterpret_cast<uint4*>(a1_row + (kb + 1) * W_LBO1) = w;
        };
        if (!(CNF_TCW_EXP & 16)) {
#pragma unroll
          for (int kb = 0; kb < (SH ? 8 : 2); kb += 2) a1_pass(kb);
          if (!SH) for (int kb = 2; kb < K1 / 8; kb += 2) a1_pass(kb);
        }
        fence_async_smem();
        tc_fence_before();
        mbar_arrive(a1_ready + slot);
        // ---- phases: EPI1 on the 128 hidden units of (net, block) ------------------------------
        // Forward with both nets: the scale half of the coupling update does not wait for the layer's last MMA.
        // D2_s is complete when the first t-net phase's D1 arrives (the commit behind that GEMM1 covers every
        // earlier MMA of the issuer), so it runs in 16-output chunks behind the EPI1s of the t-net phases, where
        // this warpgroup otherwise waits for the tensor pipe; only "+ t" is left for the layer boundary.
        const bool early = !inverse && both && !(CNF_TCW_EXP & 32);
        const int e_chunks = N2 / 16, cpp = (e_chunks + n_blk - 1) / n_blk;
        for (int ph = 0; ph < n_ph; ++ph, ++ph_cnt) {
          mbar_wait(d1_ready + slot, ph_cnt & 1);
          tc_fence_after();
          if (CNF_TCW_EXP & 2) {
            tc_fence_before();
            mbar_arrive(a2_ready + 2 * slot);
            mbar_arrive(a2_ready + 2 * slot + 1);
            continue;
          }
          uint32_t ra[32], rb[32], pk[16];
          tmem_ld32(tm, ra);
#pragma unroll 1
          for (int c = 0; c < HB; c += 64) {
            tmem_wait_ld32(ra);
            tmem_ld32(tm + c + 32, rb);
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = pack_relu_bf16(__uint_as_float(ra[2 * i]), __uint_as_float(ra[2 * i + 1]));
            tmem_st16(tm + c / 2, pk);
            tmem_wait_ld32(rb);
            if (c + 64 < HB) tmem_ld32(tm + c + 64, ra);
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = pack_relu_bf16(__uint_as_float(rb[2 * i]), __uint_as_float(rb[2 * i + 1]));
            tmem_st16(tm + c / 2 + 16, pk);
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(a2_ready + 2 * slot + (c >> 6));
          }
          if (early && ph >= n_blk)
            for (int ch = (ph - n_blk) * cpp; ch < min((ph - n_blk + 1) * cpp, e_chunks); ++ch) epi2_chunk(ch * 16, l, ld, 1);
        }
        // ---- EPI2: coupling update in fp32, 16 outputs at a time (what is left of it) -----------
        mbar_wait(d2_ready + slot, lay_cnt & 1);
        tc_fence_after();
        if (!(CNF_TCW_EXP & 8)) {
          if (early) for (int qc = 0; qc < N2; qc += 16) epi2_chunk(qc, l, ld, 2);
          else       for (int qc = 0; qc < N2; qc += 16) epi2_chunk(qc, l, ld, 0);
        }
      }
      if (base + t < N) logdet[base + t] = ld;
      wg_sync(slot);
      if (!(CNF_TCW_EXP & 4)) {
        float* gp = zout + base * K;
        const int64_t avail = (N - base) * (int64_t)K;
        int s = s0, f = f0;
        for (int e = t; e < tile_elems; e += 128) {
          if (e < avail) gp[e] = act[(inverse ? f : pi_last[f]) * ACT_LD + s];
          s += ds; f += df;
          while (f >= K) { f -= K; ++s; }
        }
      }
      wg_sync(slot);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

__global__ void tcw_gather_f32(const float* __restrict__ flat, const int* __restrict__ gather, float* __restrict__ out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const int gi = gather[i]; out[i] = gi >= 0 ? flat[gi] : 0.f; }
}


}  // namespace

// conditioners with two hidden layers: cnf_flow_tcm.cu (same streamed-weight scheme, one more GEMM per phase)
bool cnf_tcm_supported(const CnfDims& d);
long long cnf_tcm_blob_bytes(const CnfDims& d);
long long cnf_tcm_gather_len(const CnfDims& d);
int cnf_tcm_plan_build(const CnfDims& d, int32_t* g);
int cnf_tcm_pack(const CnfDims& d, const float* flat, const int32_t* gather_tc, void* packed_tc, cudaStream_t st);
int cnf_tcm_apply(const CnfDims& d, const void* packed_tc, const int32_t* tables, const float* x, float* z,
                  float* logdet, int64_t N, int inverse, cudaStream_t st);

bool cnf_tcw_supported(const CnfDims& d) { TcwDims t; return d.m >= 2 ? cnf_tcm_supported(d) : tcw_dims(d, &t); }

long long cnf_tcw_blob_bytes(const CnfDims& d) {
  TcwDims t;
  if (d.m >= 2) return cnf_tcm_blob_bytes(d);
  return tcw_dims(d, &t) ? (long long)t.blob_bytes : 0;
}

long long cnf_tcw_gather_len(const CnfDims& d) {
  TcwDims t;
  if (d.m >= 2) return cnf_tcm_gather_len(d);
  return tcw_dims(d, &t) ? (long long)t.n_bf16 + t.n_f32 : 0;
}

int cnf_tcw_plan_build(const CnfDims& d, int32_t* g) {
  TcwDims t;
  if (d.m >= 2) return cnf_tcm_plan_build(d, g);
  if (!tcw_dims(d, &t)) { cnf_set_error("wide tensor-core path not available for this shape"); return CNF_E_UNSUPPORTED; }
  const int K = d.K, half = K / 2, H = d.H[0];
  const long long total = (long long)t.n_bf16 + t.n_f32;
  for (long long i = 0; i < total; ++i) g[i] = -1;
  const long long net_sz = (long long)H * K + H + (long long)K * H + K;
  const int lbo2 = t.N2 * 16;
  for (int l = 0; l < d.L; ++l) {
    int slot = 0;
    for (int net = 0; net < 2; ++net) {
      if (!(d.nets & (1 << net))) continue;
      const long long base = ((long long)l * d.n_nets + slot) * net_sz;
      const long long w0 = base, b0 = base + (long long)H * K, w1 = b0 + H, b1 = w1 + (long long)K * H;
      for (int blk = 0; blk < t.n_blk; ++blk) {
        const long long phase = ((long long)l * t.n_ph + slot * t.n_blk + blk) * t.phase_bytes;
        int32_t* B1 = g + phase / 2;
        int32_t* B2 = g + (phase + t.b1_bytes) / 2;
        for (int n = 0; n < HB; ++n) {
          const int h = blk * HB + n;
          if (h >= H) break;
          for (int k = 0; k <= d.d1; ++k) {
            const int byte = (k / 8) * W_LBO1 + (n / 8) * W_SBO1 + (n % 8) * 16 + (k % 8) * 2;
            B1[byte / 2] = (int32_t)(k < d.d1 ? w0 + (long long)h * K + half + k : b0 + h);
          }
        }
        for (int q = 0; q < d.d0; ++q)
          for (int kk = 0; kk < HB; ++kk) {
            const int h = blk * HB + kk;
            if (h >= H) break;
            const int byte = (kk / 16) * (t.N2 * 32) + ((kk % 16) / 8) * lbo2 + (q / 8) * W_SBO2 + (q % 8) * 16 + (kk % 8) * 2;
            B2[byte / 2] = (int32_t)(w1 + (long long)q * H + h);
          }
      }
      for (int q = 0; q < d.d0; ++q) g[t.n_bf16 + (l * 2 + slot) * t.N2 + q] = (int32_t)(b1 + q);
      ++slot;
    }
  }
  return CNF_OK;
}

int cnf_tcw_pack(const CnfDims& d, const float* flat, const int32_t* gather_tc, void* packed_tc, cudaStream_t st) {
  TcwDims t;
  if (d.m >= 2) return cnf_tcm_pack(d, flat, gather_tc, packed_tc, st);
  if (!tcw_dims(d, &t)) { cnf_set_error("wide tensor-core path not available for this shape"); return CNF_E_UNSUPPORTED; }
  int rc = cnf_pack_bf16(flat, gather_tc, packed_tc, t.n_bf16, st);
  if (rc) return rc;
  tcw_gather_f32<<<(t.n_f32 + 127) / 128, 128, 0, st>>>(flat, gather_tc + t.n_bf16,
                                                        reinterpret_cast<float*>((uint8_t*)packed_tc + t.bias_off), t.n_f32);
  CNF_CHECK_CUDA(cudaGetLastError());
  return CNF_OK;
}

int cnf_tcw_apply(const CnfDims& d, const void* packed_tc, const int32_t* tables, const float* x, float* z,
                  float* logdet, int64_t N, int inverse, cudaStream_t st) {
  TcwDims t;
  if (d.m >= 2) return cnf_tcm_apply(d, packed_tc, tables, x, z, logdet, N, inverse, st);
  if (!tcw_dims(d, &t)) { cnf_set_error("wide tensor-core path not available for this shape"); return CNF_E_UNSUPPORTED; }
  CnfDevInfo di;
  { const int rc2 = cnf_dev_info(&di); if (rc2) return rc2; }
  const int g_tcw_sms = di.sms;
  if ((uintptr_t)packed_tc % 16 != 0) { cnf_set_error("packed_tc must be 16-byte aligned"); return CNF_E_ARG; }
  const int64_t ntiles = (N + TILE_M - 1) / TILE_M;
  const int grid = (int)(ntiles < g_tcw_sms ? ntiles : g_tcw_sms);
  const bool sh = d.K == 100 && d.H[0] == 512 && d.nets == 3 && !cnf_switch(CNF_SW_TC_GENERIC);
  if (sh) {
    { const int rc2 = cnf_kernel_smem(flow_tcw_kernel<1>, t.sm_total); if (rc2) return rc2; }
    flow_tcw_kernel<1><<<grid, W_THREADS, t.sm_total, st>>>(t, (const uint8_t*)packed_tc, tables, x, z, logdet, N, inverse);
  } else {
    { const int rc2 = cnf_kernel_smem(flow_tcw_kernel<0>, t.sm_total); if (rc2) return rc2; }
    flow_tcw_kernel<0><<<grid, W_THREADS, t.sm_total, st>>>(t, (const uint8_t*)packed_tc, tables, x, z, logdet, N, inverse);
  }
  CNF_CHECK_CUDA(cudaGetLastError());
  return CNF_OK;
}
