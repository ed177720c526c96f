// bf16 tensor-core kernel of the coupling-flow stack for conditioners with TWO TO FOUR hidden layers of 16 .. 128
// units each (flows/utils.py:6-31 allows any depth; flows/flows.py:69 defaults to two hidden layers), K up to 65
// classes, sm_100a.  The H x H middle Linears are the dense contractions of the path: [128 x 128] x [128 x 128] per
// tile and net, eight full-width tcgen05.mma k-steps each.
//
// Per coupling layer and tile of 128 samples (TMEM lane = sample row), one "phase" per conditioner net:
//   GEMM1    D1[128 x H1p] = A1[128 x K1] . B1^T         SS; A1 row = (conditioning logits, 1, 0..): first bias folded
//   EPI1     h1 = relu(D1) -> bf16, in place (A aliases the low half of D1)
//   GEMMmid  Dm[128 x H2p] = h1[128 x H1p] . Bm^T         TS (A from TMEM), H1p/16 k-steps
//   EPImid   h2 = relu(Dm + bm) -> bf16 into the A columns     (GEMMmid / EPImid once per further hidden layer;
//            the last one is released to the issuer in two halves)
//   GEMM2    D2_net[128 x N2] = h_m[128 x Hmp] . B2^T     TS, Hmp/16 k-steps
// then EPI2: s, t = D2 + b2 in fp32; y = x e^s + t, ld += sum s (flows/flows.py:107-109); inverse
// x = (y - t) e^-s, ld -= sum s (:121-125).
// TMEM per tile slot (256 columns): D1 [0,128), A [0,64), Dm [64,192), D2_s / D2_t [192, 192 + 2 N2).  Dm overlaps
// the upper half of D1, so GEMMmid waits for the whole EPI1; GEMM1 of the next net is issued right behind GEMM2 (the
// in-order tensor pipe resolves the WAR on the A columns; Dm has been consumed by then).
// Weights do not stay resident (6 layers of [128,128] nets are 480 KB of bf16): a phase's weights are m - 1 blocks
// -- block b holds the middle image Bm_(b+1), the first one B1 in front of it, the last one B2 behind it --
// contiguous in the packed blob, streamed one block per stage into a 2-stage shared-memory ring by TMA bulk copies
// that a producer warp issues one block ahead; two tiles per CTA walk the blocks in lockstep and share each stage
// (the scheme of cnf_flow_tcw.cu).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdlib>

#include "cnf_common.h"
#ifndef CNF_TC_FASTEXP
#define CNF_TC_FASTEXP 1
#endif
#include "cnf_tc_ptx.cuh"

int cnf_pack_bf16(const float* flat, const int32_t* gather, void* packed, int n, cudaStream_t st);

namespace {

constexpr int M_THREADS = 384;   // warps 0 / 3 MMA issuers (slot 0 / 1), 1 weight producer, 2 TMEM allocator, 4-7 / 8-11 epilogue
constexpr int ACT_LD = TILE_M + 1;           // act[slot][sample] row stride (odd: conflict-free transposing tile I/O)
constexpr int M_LBO1 = 2048, M_SBO = 128;    // A1 / B1: [k-block][row-block] 128-byte core matrices, 128 rows
constexpr int COL_DM = 64, COL_D2 = 192;     // TMEM columns of a slot
constexpr int BM_LD = 128;                   // middle biases per (layer, net) in the fp32 section

constexpr int MAX_MID = CNF_MAX_HIDDEN - 1;   // middle Linears per net

struct TcmDims {
  int K, L, d0, d1, m, Hp[CNF_MAX_HIDDEN], nets, n_nets, K1, N2, n_ph;
  int b1_bytes;                          // B1 image in front of the first block's Bm
  int blk_off[MAX_MID], blk_bytes[MAX_MID], b2_off, stage_bytes, phase_bytes;   // blocks of one phase (bytes inside the phase); B2 inside the last block
  int f32_off, n_bf16, n_f32, bm_floats, blob_bytes;   // fp32 section: [L * n_nets * (m-1)][128] middle biases, [L * 2][N2] last biases
  int tab_pi, tab_cond, tab_trans, n_tables;
  int sm_ring, sm_f32, sm_tab, sm_tp, sm_slot, sm_slot_stride, sm_act, sm_bar, sm_total;
};

bool tcm_dims(const CnfDims& d, TcmDims* t) {
  if (d.m < 2 || d.m > CNF_MAX_HIDDEN || d.n_nets < 1) return false;
  // at least one 16-unit k-step per hidden layer: narrower nets (the reference's default [5, 5]) are no dense
  // contraction and belong on the fp32 register kernels (15 G samples/s there)
  for (int j = 0; j < d.m; ++j) if (d.H[j] < 16 || d.H[j] > 128) return false;
  if (d.d1 + 1 > 64 || d.d0 > 32) return false;
  t->K = d.K; t->L = d.L; t->d0 = d.d0; t->d1 = d.d1; t->m = d.m; t->nets = d.nets; t->n_nets = d.n_nets;
  for (int j = 0; j < CNF_MAX_HIDDEN; ++j) t->Hp[j] = j < d.m ? cnf_round_up(d.H[j], 16) : 0;
  t->K1 = cnf_round_up(d.d1 + 1, 16);
  t->N2 = cnf_round_up(d.d0, 16);
  t->n_ph = d.n_nets;
  t->b1_bytes = TILE_M * t->K1 * 2;
  t->stage_bytes = 0;
  {
    int off = 0;
    for (int b = 0; b < MAX_MID; ++b) { t->blk_off[b] = 0; t->blk_bytes[b] = 0; }
    for (int b = 0; b < d.m - 1; ++b) {
      t->blk_off[b] = off;
      int bytes = t->Hp[b] * t->Hp[b + 1] * 2;
      if (b == 0) bytes += t->b1_bytes;
      if (b == d.m - 2) { t->b2_off = off + bytes; bytes += t->Hp[d.m - 1] * t->N2 * 2; }
      t->blk_bytes[b] = bytes;
      if (bytes > t->stage_bytes) t->stage_bytes = bytes;
      off += bytes;
    }
    t->phase_bytes = off;
  }
  const long long wbytes = (long long)d.L * t->n_ph * t->phase_bytes;
  if (wbytes > (1ll << 30)) return false;
  t->f32_off = (int)wbytes;
  t->n_bf16 = (int)(wbytes / 2);
  t->bm_floats = d.L * d.n_nets * (d.m - 1) * BM_LD;
  t->n_f32 = t->bm_floats + d.L * 2 * t->N2;
  t->blob_bytes = t->f32_off + t->n_f32 * 4;
  t->tab_pi = d.tab_pi; t->tab_cond = d.tab_cond; t->tab_trans = d.tab_trans; t->n_tables = d.n_tables;
  int off = 0;
  t->sm_ring = off; off += 2 * t->stage_bytes;
  t->sm_f32 = off; off += (t->n_f32 * 4 + 127) / 128 * 128;
  t->sm_tab = off; off += (d.n_tables * 4 + 127) / 128 * 128;
  t->sm_tp = off; off += (d.L * t->N2 * 2 + 127) / 128 * 128;   // act offsets (uint16) of the transformed slots, padded to N2
  t->sm_slot = off;
  t->sm_act = TILE_M * t->K1 * 2;                   // A1 tile first, then the fp32 tile
  t->sm_slot_stride = t->sm_act + (d.K * ACT_LD * 4 + 127) / 128 * 128;
  off += 2 * t->sm_slot_stride;
  t->sm_bar = off; off += 256;
  t->sm_total = off;
  return t->sm_total <= 227 * 1024;
}

// fp32 accumulator columns [src, src + Hp) of this thread's TMEM lane -> (+ bias) -> relu -> bf16 pairs at columns
// [0, Hp / 2): the A operand of the next GEMM.  32 columns per step, the next step's load in flight behind the
// current one's conversion; the last step may run into padding columns (never read by an MMA).  bar_half (optional)
// is arrived on once the first ceil(steps / 2) steps are stored, bar_full after the last one.
template <bool BIAS>
__device__ __forceinline__ void hidden_pass(uint32_t tm, int src, int Hp, const float* __restrict__ bz,
                                            uint64_t* bar_half, uint64_t* bar_full) {
  const int nch = (Hp + 31) >> 5, half = (nch + 1) >> 1;
  uint32_t ra[32], rb[32];
  auto emit = [&](const uint32_t (&r)[32], const int ch) {
    uint32_t pk[16];
    if (BIAS) {
      const float4* b4 = reinterpret_cast<const float4*>(bz + ch * 32);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 q = b4[i];
        pk[2 * i] = pack_relu_bf16(__uint_as_float(r[4 * i]) + q.x, __uint_as_float(r[4 * i + 1]) + q.y);
        pk[2 * i + 1] = pack_relu_bf16(__uint_as_float(r[4 * i + 2]) + q.z, __uint_as_float(r[4 * i + 3]) + q.w);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) pk[i] = pack_relu_bf16(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
    }
    tmem_st16(tm + ch * 16, pk);
    const bool at_half = bar_half != nullptr && ch + 1 == half, at_end = ch + 1 == nch;
    if (at_half || at_end) {
      tmem_wait_st();
      tc_fence_before();
      if (at_half) mbar_arrive(bar_half);
      if (at_end) mbar_arrive(bar_full);
    }
  };
  tmem_ld32(tm + src, ra);
#pragma unroll 1
  for (int ch = 0; ch < nch; ch += 2) {
    tmem_wait_ld32(ra);
    if (ch + 1 < nch) tmem_ld32(tm + src + (ch + 1) * 32, rb);
    emit(ra, ch);
    if (ch + 1 < nch) {
      tmem_wait_ld32(rb);
      if (ch + 2 < nch) tmem_ld32(tm + src + (ch + 2) * 32, ra);
      emit(rb, ch + 1);
    }
  }
}

template <int NMID>      // middle Linears per net (hidden layers - 1)
__global__ void __launch_bounds__(M_THREADS, 1)
flow_tcm_kernel(TcmDims p, const uint8_t* __restrict__ blob, const int* __restrict__ tables,
                const float* __restrict__ xin, float* __restrict__ zout, float* __restrict__ logdet, int64_t N,
                int inverse) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int* s_hp = reinterpret_cast<int*>(smem + p.sm_bar + 192);   // padded widths, block offsets / bytes (indexed at run time)
  int* s_blk = s_hp + CNF_MAX_HIDDEN;
  if (tid == 0) {
    s_hp[0] = p.Hp[0]; s_hp[1] = p.Hp[1]; s_hp[2] = p.Hp[2]; s_hp[3] = p.Hp[3];
    s_blk[0] = p.blk_off[0]; s_blk[1] = p.blk_off[1]; s_blk[2] = p.blk_off[2];
    s_blk[3] = p.blk_bytes[0]; s_blk[4] = p.blk_bytes[1]; s_blk[5] = p.blk_bytes[2];
  }
  int* tab = reinterpret_cast<int*>(smem + p.sm_tab);
  float* f32s = reinterpret_cast<float*>(smem + p.sm_f32);
  float* bias = f32s + p.bm_floats;     // last biases [L * 2][N2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.sm_bar);
  uint64_t* w_full = bars;          // [2] producer (TMA complete_tx) -> MMA: weight stage landed
  uint64_t* w_empty = bars + 2;     // [2] MMA (one commit per slot) -> producer: stage consumed by both tiles
  uint64_t* a1_ready = bars + 4;    // [2] epilogue -> MMA, once per layer
  uint64_t* d1_ready = bars + 6;    // [2] MMA -> epilogue, once per phase
  uint64_t* am_ready = bars + 8;    // [2] epilogue -> MMA, m - 1 times per phase: h_1 .. h_(m-1) complete
  uint64_t* dm_ready = bars + 10;   // [2] MMA -> epilogue, m - 1 times per phase
  uint64_t* a2_ready = bars + 12;   // [2 slots][2 groups] epilogue -> MMA, once per phase: halves of h_m
  uint64_t* d2_ready = bars + 16;   // [2] MMA -> epilogue, once per layer
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 18);

  for (int i = tid; i < p.n_tables; i += M_THREADS) tab[i] = tables[i];
  // EPI2 needs no guards: a padded output (q >= d0: zero weights and biases, so s = t = 0 exactly) "updates" the
  // layer's first conditioning slot, which no real output touches, with x * e^0 + 0 = x
  unsigned short* tp = reinterpret_cast<unsigned short*>(smem + p.sm_tp);
  for (int i = tid; i < p.L * p.N2; i += M_THREADS) {
    const int l = i / p.N2, q = i - l * p.N2;
    tp[i] = (unsigned short)((q < p.d0 ? tables[p.tab_trans + l * p.d0 + q] : tables[p.tab_cond + l * p.d1]) * ACT_LD);
  }
  {
    const float* gb = reinterpret_cast<const float*>(blob + p.f32_off);
    for (int i = tid; i < p.n_f32; i += M_THREADS) f32s[i] = __ldg(gb + i);
  }
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(w_full + s, 1); mbar_init(w_empty + s, 2);
      mbar_init(a1_ready + s, 128); mbar_init(d1_ready + s, 1); mbar_init(dm_ready + s, 1); mbar_init(d2_ready + s, 1);
      mbar_init(am_ready + s, 128);
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
  const int64_t nt0 = (mine + 1) / 2, nt1 = mine / 2;
  const int64_t rounds = nt0;
  const int n_ph = p.n_ph, K1 = p.K1, N2 = p.N2, D1 = p.d1, KK = p.K;
  const int NETS = p.nets;
  constexpr int n_mid = NMID;
  const int Hlast = NMID == 1 ? p.Hp[1] : (NMID == 2 ? p.Hp[2] : p.Hp[3]);
  const int nk1 = K1 / 16, nk2 = Hlast / 16;
  const int nk2_g0 = min(2 * ((((Hlast + 31) >> 5) + 1) >> 1), nk2);   // k-steps of GEMM2 covered by the first half of the last EPImid

  if (warp == 1) {
    // ================================ weight producer ==========================================
    if (lane == 0) {
      uint32_t g = 0;
      for (int64_t r = 0; r < rounds; ++r)
        for (int li = 0; li < p.L; ++li) {
          const int l = inverse ? p.L - 1 - li : li;
          for (int ph = 0; ph < n_ph; ++ph)
            for (int b = 0; b < n_mid; ++b, ++g) {
              const int st = g & 1;
              if (g >= 2) mbar_wait(w_empty + st, ((g >> 1) - 1) & 1);
              mbar_expect_tx(w_full + st, (uint32_t)s_blk[MAX_MID + b]);
              bulk_copy_g2s(smem + p.sm_ring + st * p.stage_bytes,
                            blob + ((size_t)l * n_ph + ph) * p.phase_bytes + s_blk[b], (uint32_t)s_blk[MAX_MID + b], w_full + st);
            }
        }
    }
    __syncwarp();
  } else if (warp == 0 || warp == 3) {
    // ================================ MMA issuers: warp 0 -> slot 0, warp 3 -> slot 1 ===============
    if (lane == 0) {
      const int s = warp == 0 ? 0 : 1;
      const uint32_t idesc1 = make_idesc(s_hp[0]), idesc2 = make_idesc(N2);
      const uint32_t smem_base = smem_u32(smem);
      const uint32_t tm = tmem_base + s * 256;
      const uint64_t a1d = make_desc(smem_base + p.sm_slot + s * p.sm_slot_stride, M_LBO1, M_SBO);
      const uint64_t kstep1 = (uint64_t)(2 * M_LBO1 / 16);      // descriptor start-address units per k-step
      const uint64_t kstep2 = (uint64_t)(N2 * 32 / 16);
      uint32_t g = 0, lay_cnt = 0, ph_cnt = 0, mid_cnt = 0;
      auto gemm1 = [&](uint32_t gg) {
        const uint64_t b1d = make_desc(smem_base + p.sm_ring + (gg & 1) * p.stage_bytes, M_LBO1, M_SBO);
        for (int j = 0; j < nk1; ++j) mma_ss(tm, a1d + j * kstep1, b1d + j * kstep1, idesc1, j > 0 ? 1u : 0u);
        tc_commit(d1_ready + s);
      };
      for (int64_t r = 0; r < rounds; ++r) {
        const bool has_tile = r < (s == 0 ? nt0 : nt1);
        for (int li = 0; li < p.L; ++li) {
          if (!has_tile) {           // keep the ring's arrival count
            for (int i = 0; i < n_ph * n_mid; ++i, ++g) {
              mbar_wait_backoff(w_full + (g & 1), (g >> 1) & 1);
              mbar_arrive(w_empty + (g & 1));
            }
            continue;
          }
          mbar_wait_backoff(a1_ready + s, lay_cnt & 1);
          ++lay_cnt;
          mbar_wait_backoff(w_full + (g & 1), (g >> 1) & 1);
          tc_fence_after();
          gemm1(g);
          for (int ph = 0; ph < n_ph; ++ph) {
            const uint32_t d2 = tm + COL_D2 + ph * N2;
            // ---- middle GEMMs: block b holds Bm_(b+1); its stage has landed for b == 0 (waited before GEMM1) ----
            for (int bk = 0; bk < n_mid; ++bk) {
              if (bk > 0) { mbar_wait_backoff(w_full + (g & 1), (g >> 1) & 1); }
              const uint32_t stage = smem_base + p.sm_ring + (g & 1) * p.stage_bytes;
              const int Hin = s_hp[bk], Hout = s_hp[bk + 1];
              const uint64_t bmd = make_desc(stage + (bk == 0 ? p.b1_bytes : 0), (uint32_t)Hout * 16, M_SBO);
              const uint64_t kstepm = (uint64_t)(Hout * 32 / 16);
              const uint32_t idescm = make_idesc(Hout);
              mbar_wait_backoff(am_ready + s, mid_cnt & 1);
              tc_fence_after();
              for (int j = 0; j < Hin / 16; ++j) mma_ts(tm + COL_DM, tm + j * 8, bmd + j * kstepm, idescm, j > 0 ? 1u : 0u);
              tc_commit(dm_ready + s);
              ++mid_cnt;
              if (bk + 1 < n_mid) { tc_commit(w_empty + (g & 1)); ++g; }
            }
            // ---- GEMM2: B2 sits behind the last block's Bm ----
            const uint32_t stage = smem_base + p.sm_ring + (g & 1) * p.stage_bytes;
            const uint64_t b2d = make_desc(stage + (p.b2_off - s_blk[n_mid - 1]), (uint32_t)N2 * 16, M_SBO);
            mbar_wait_backoff(a2_ready + 2 * s, ph_cnt & 1);
            tc_fence_after();
            for (int j = 0; j < nk2_g0; ++j) mma_ts(d2, tm + j * 8, b2d + j * kstep2, idesc2, j > 0 ? 1u : 0u);
            mbar_wait_backoff(a2_ready + 2 * s + 1, ph_cnt & 1);
            tc_fence_after();
            for (int j = nk2_g0; j < nk2; ++j) mma_ts(d2, tm + j * 8, b2d + j * kstep2, idesc2, 1u);
            ++ph_cnt;
            // this slot's GEMM1 of the next net goes in right behind its GEMM2: the next EPI1 overlaps the other
            // slot's tensor work
            if (ph + 1 < n_ph) {
              mbar_wait_backoff(w_full + ((g + 1) & 1), ((g + 1) >> 1) & 1);
              tc_fence_after();
              gemm1(g + 1);
            }
            tc_commit(w_empty + (g & 1));
            ++g;
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
    uint8_t* a1_row = a1 + (t >> 3) * M_SBO + (t & 7) * 16;
    const int K = KK, tile_elems = TILE_M * KK;
    const int s0 = t / K, f0 = t - s0 * K, ds = TILE_M / K, df = TILE_M - ds * K;
    const bool both = (NETS == 3);
    uint32_t lay_cnt = 0, ph_cnt = 0, mid_cnt = 0;
    // EPI2 on outputs qc .. qc+15.  mode 0: the whole coupling update; 1: its scale half (x *= e^s, log-det);
    // 2: its shift half (x += t).  Eight outputs at a time, loads -> math -> stores; no guards (see tp).
    auto epi2_chunk = [&](const int qc, const int l, float& ld, const int mode) {
      const float* bl = bias + l * 2 * N2 + qc;
      const unsigned short* tpl = tp + l * N2 + qc;
      uint32_t r1[16], r2[16];
      if (mode != 2) tmem_ld16(tm + COL_D2 + qc, r1);
      if (mode == 2 || (mode == 0 && both)) tmem_ld16(tm + COL_D2 + N2 + qc, r2);
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

    const int64_t my_tiles = slot == 0 ? nt0 : nt1;
    for (int64_t r = 0; r < my_tiles; ++r) {
      const int64_t tile = blockIdx.x + (2 * r + slot) * (int64_t)G;
      const int64_t base = tile * TILE_M;
      {
        // the tile goes straight into the transposed layout with 4-byte cp.async copies (rows past N zero-filled)
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
        for (int kb = 0; kb < K1 / 8; kb += 2) {
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
          *reinterpret_cast<uint4*>(a1_row + kb * M_LBO1) = v;
          *reinterpret_cast<uint4*>(a1_row + (kb + 1) * M_LBO1) = w;
        }
        fence_async_smem();
        tc_fence_before();
        mbar_arrive(a1_ready + slot);
        // Forward with both nets: the scale half of the coupling update runs while the tensor pipe works on the
        // t-net's middle GEMM (D2_s is complete when the t-net's D1 arrives: the commit behind that GEMM1 covers
        // every earlier MMA of the issuer); only "+ t" is left for the layer boundary.
        const bool early = !inverse && both;
        for (int ph = 0; ph < n_ph; ++ph, ++ph_cnt) {
          mbar_wait(d1_ready + slot, ph_cnt & 1);
          tc_fence_after();
          hidden_pass<false>(tm, 0, s_hp[0], nullptr, nullptr, am_ready + slot);
          if (early && ph == 1)
            for (int qc = 0; qc < N2; qc += 16) epi2_chunk(qc, l, ld, 1);
          for (int bk = 0; bk < n_mid; ++bk, ++mid_cnt) {
            mbar_wait(dm_ready + slot, mid_cnt & 1);
            tc_fence_after();
            const float* bz = f32s + ((l * p.n_nets + ph) * n_mid + bk) * BM_LD;
            const bool last = bk + 1 == n_mid;      // the last hidden layer is released in two halves (GEMM2 starts on the first)
            hidden_pass<true>(tm, COL_DM, s_hp[bk + 1], bz, last ? a2_ready + 2 * slot : nullptr,
                              last ? a2_ready + 2 * slot + 1 : am_ready + slot);
          }
        }
        // ---- EPI2: coupling update in fp32, 16 outputs at a time (what is left of it) -----------
        mbar_wait(d2_ready + slot, lay_cnt & 1);
        tc_fence_after();
        if (early) for (int qc = 0; qc < N2; qc += 16) epi2_chunk(qc, l, ld, 2);
        else       for (int qc = 0; qc < N2; qc += 16) epi2_chunk(qc, l, ld, 0);
      }
      if (base + t < N) logdet[base + t] = ld;
      wg_sync(slot);
      {
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

__global__ void tcm_gather_f32(const float* __restrict__ flat, const int* __restrict__ gather, float* __restrict__ out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const int gi = gather[i]; out[i] = gi >= 0 ? flat[gi] : 0.f; }
}

}  // namespace

bool cnf_tcm_supported(const CnfDims& d) { TcmDims t; return tcm_dims(d, &t); }

long long cnf_tcm_blob_bytes(const CnfDims& d) { TcmDims t; return tcm_dims(d, &t) ? (long long)t.blob_bytes : 0; }

long long cnf_tcm_gather_len(const CnfDims& d) { TcmDims t; return tcm_dims(d, &t) ? (long long)t.n_bf16 + t.n_f32 : 0; }

// gather[i]: index into the canonical flat parameters of bf16 element i of the phase images, then of the fp32
// section's entries; -1 = zero.  One net's canonical block: W0 [H1,K], b0 [H1], W1 [H2,H1], b1 [H2], ...,
// W_m [K,Hm], b_m [K] (flows/utils.py:12-24).
int cnf_tcm_plan_build(const CnfDims& d, int32_t* g) {
  TcmDims t;
  if (!tcm_dims(d, &t)) { cnf_set_error("deep tensor-core path not available for this shape"); return CNF_E_UNSUPPORTED; }
  const int K = d.K, half = K / 2, m = d.m;
  const long long total = (long long)t.n_bf16 + t.n_f32;
  for (long long i = 0; i < total; ++i) g[i] = -1;
  long long w_at[CNF_MAX_HIDDEN + 1], b_at[CNF_MAX_HIDDEN + 1], net_sz = 0;     // offsets of Linear j inside one net
  for (int j = 0; j <= m; ++j) {
    const int n_in = j == 0 ? K : d.H[j - 1], n_out = j == m ? K : d.H[j];
    w_at[j] = net_sz; net_sz += (long long)n_out * n_in;
    b_at[j] = net_sz; net_sz += n_out;
  }
  const int lbo2 = t.N2 * 16, Hl = d.H[m - 1];
  for (int l = 0; l < d.L; ++l) {
    int slot = 0;
    for (int net = 0; net < 2; ++net) {
      if (!(d.nets & (1 << net))) continue;
      const long long base = ((long long)l * d.n_nets + slot) * net_sz;
      const long long phase = ((long long)l * t.n_ph + slot) * t.phase_bytes;
      int32_t* B1 = g + phase / 2;
      for (int n = 0; n < d.H[0]; ++n)
        for (int k = 0; k <= d.d1; ++k) {
          const int byte = (k / 8) * M_LBO1 + (n / 8) * M_SBO + (n % 8) * 16 + (k % 8) * 2;
          B1[byte / 2] = (int32_t)(k < d.d1 ? base + w_at[0] + (long long)n * K + half + k : base + b_at[0] + n);
        }
      for (int bk = 0; bk < m - 1; ++bk) {       // middle Linear bk + 1: [H_(bk+2) x H_(bk+1)]
        int32_t* Bm = g + (phase + t.blk_off[bk] + (bk == 0 ? t.b1_bytes : 0)) / 2;
        const int Hin = d.H[bk], Hout = d.H[bk + 1], Houtp = t.Hp[bk + 1];
        for (int n = 0; n < Hout; ++n)
          for (int kk = 0; kk < Hin; ++kk) {
            const int byte = (kk / 16) * (Houtp * 32) + ((kk % 16) / 8) * (Houtp * 16) + (n / 8) * M_SBO + (n % 8) * 16 + (kk % 8) * 2;
            Bm[byte / 2] = (int32_t)(base + w_at[bk + 1] + (long long)n * Hin + kk);
          }
        for (int n = 0; n < Hout; ++n)
          g[t.n_bf16 + (((long long)l * d.n_nets + slot) * (m - 1) + bk) * BM_LD + n] = (int32_t)(base + b_at[bk + 1] + n);
      }
      int32_t* B2 = g + (phase + t.b2_off) / 2;
      for (int q = 0; q < d.d0; ++q)
        for (int kk = 0; kk < Hl; ++kk) {
          const int byte = (kk / 16) * (t.N2 * 32) + ((kk % 16) / 8) * lbo2 + (q / 8) * M_SBO + (q % 8) * 16 + (kk % 8) * 2;
          B2[byte / 2] = (int32_t)(base + w_at[m] + (long long)q * Hl + kk);
        }
      for (int q = 0; q < d.d0; ++q) g[t.n_bf16 + t.bm_floats + (l * 2 + slot) * t.N2 + q] = (int32_t)(base + b_at[m] + q);
      ++slot;
    }
  }
  return CNF_OK;
}

int cnf_tcm_pack(const CnfDims& d, const float* flat, const int32_t* gather_tc, void* packed_tc, cudaStream_t st) {
  TcmDims t;
  if (!tcm_dims(d, &t)) { cnf_set_error("deep tensor-core path not available for this shape"); return CNF_E_UNSUPPORTED; }
  int rc = cnf_pack_bf16(flat, gather_tc, packed_tc, t.n_bf16, st);
  if (rc) return rc;
  tcm_gather_f32<<<(t.n_f32 + 127) / 128, 128, 0, st>>>(flat, gather_tc + t.n_bf16,
                                                        reinterpret_cast<float*>((uint8_t*)packed_tc + t.f32_off), t.n_f32);
  CNF_CHECK_CUDA(cudaGetLastError());
  return CNF_OK;
}

int cnf_tcm_apply(const CnfDims& d, const void* packed_tc, const int32_t* tables, const float* x, float* z,
                  float* logdet, int64_t N, int inverse, cudaStream_t st) {
  TcmDims t;
  if (!tcm_dims(d, &t)) { cnf_set_error("deep tensor-core path not available for this shape"); return CNF_E_UNSUPPORTED; }
  CnfDevInfo di;
  { const int rc2 = cnf_dev_info(&di); if (rc2) return rc2; }
  if ((uintptr_t)packed_tc % 16 != 0) { cnf_set_error("packed_tc must be 16-byte aligned"); return CNF_E_ARG; }
  const int64_t ntiles = (N + TILE_M - 1) / TILE_M;
  const int grid = (int)(ntiles < di.sms ? ntiles : di.sms);
#define LAUNCH_TCM(NM)                                                                                              \
  do {                                                                                                            \
    const int rc2 = cnf_kernel_smem(flow_tcm_kernel<NM>, t.sm_total);                                             \
    if (rc2) return rc2;                                                                                          \
    flow_tcm_kernel<NM><<<grid, M_THREADS, t.sm_total, st>>>(t, (const uint8_t*)packed_tc, tables, x, z, logdet, N, inverse); \
  } while (0)
  if (t.m == 2) LAUNCH_TCM(1); else if (t.m == 3) LAUNCH_TCM(2); else LAUNCH_TCM(3);
#undef LAUNCH_TCM
  CNF_CHECK_CUDA(cudaGetLastError());
  return CNF_OK;
}
