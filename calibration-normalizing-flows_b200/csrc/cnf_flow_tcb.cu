// bf16 tensor-core TRAINING backward of the coupling-flow stack for sm_100a (tcgen05 + TMEM).
//
// Replaces, for the shapes the resident-weight forward kernel covers (one hidden layer, K <= 14,
// pad16(H) <= 128, L*nets <= 15), the autograd backward that the reference gets from torch
// (loss.backward() at calibrators.py:293-294 / run_experiment3D.py:133-134) for the calibrator's
// loss  -mean(log(softmax(z)[y]+eps) + gamma*log_det)  (calibrators.py:288-291).
//
// Inputs: the forward kernel's outputs for the same samples -- z, log-det and the per-layer tape
// (pre-layer values of the transformed slots and the scale-net outputs s: exactly what autograd would
// have saved) -- plus the labels.  Per tile of 128 samples and per coupling layer l = L-1 .. 0
// (u = conditioning logits, x_t = transformed logits, g_y = upstream gradient on the transformed slots):
//
//   E0    g_s = g_y * x_t * e^s + g_ld ; g_t = g_y ; g_xt = g_y * e^s          (fp32, one thread per sample)
//         shared-memory record per 8 samples: [G2s | G2t | 0 | A1 | 0] (bf16, 8x8 core matrices)
//   per conditioner net p (s, then t), in halves of 64 hidden units:
//   T1    D1 [128 smp x 64] = A1 . B1p^T         (recompute of the hidden pre-activations, bias folded)
//         GH [128 smp x 64] = G2 . W2p           (B operand: the forward's B2 image read MN-major)
//   E1    h = relu(D1), ghm = GH * [D1 > 0]  -> bf16 shared-memory images [128 smp x Hp] whose 8x8 core
//         matrices serve both as a K-major operand (contraction over the hidden units) and as an
//         MN-major operand (contraction over the samples)
//   T2    GU [128 smp x 16] += ghm . W1p         (B: the forward's B1 image read MN-major)
//   then, per net:
//   T4    ACC[l][p] [128 hid x 16] += h^T . [G2p | 0] + ghm^T . [0 | A1]   (contraction over the samples)
//         columns 0..7: d/dW2p (last Linear), 8..15: d/dW1p and, through A1's constant-one column, d/db1p
//   E5    g_u += GU                                                         (fp32)
// The weight-gradient accumulators ACC stay in TMEM for the whole launch (fp32, 16 columns per layer
// and net, zeroed at the start) and are added to the CTA's own row of the partial buffer at the end;
// last-layer bias gradients are warp-reduced in fp32.
//
// One persistent CTA per SM, 352 threads.  TWO tiles are in flight (slots): each has its own epilogue
// warpgroup, MMA-issuer warp, TMEM working set (144 columns) and shared-memory records / images, so one
// tile's epilogue runs under the other tile's MMAs; both accumulate into the same ACC columns.  A
// producer warp streams each layer's B1 / B2 images (the forward kernel's blob) through a 3-stage
// shared-memory ring with TMA bulk copies (full / empty mbarriers, one empty arrival per slot).
//
// Measured on B200 (profiles/microbench/tmem_bw.cu): tcgen05.ld moves 128 B/clk per SM whatever the
// number of warps (stores ride along for free), so the fp32 reads of D1 and GH (2 x 512 clk per net at
// Hp = 128) are a floor of the epilogue; a first version that formed h^T / ghm^T by a second,
// transposed recompute in TMEM doubled those reads and ran at 0.6x the speed of this one.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdlib>

#include "cnf_common.h"
#include "cnf_tc_dims.h"
#include "cnf_tc_ptx.cuh"

// Timing-only experiment switches (wrong results; profiles/microbench/build_variant.sh): bit 0 = E1 does no work,
// bit 1 = no E0 math, bit 2 = no loss head, bit 3 = no E5, bit 4 = no T4 MMAs, bit 5 = no T2 MMAs.
#ifndef CNF_TCB_EXP
#define CNF_TCB_EXP 0
#endif
#ifndef CNF_TCB_STAGGER
#define CNF_TCB_STAGGER 0
#endif
#ifndef CNF_TCB_BATON
#define CNF_TCB_BATON 1
#endif

namespace {

constexpr int TB_SLOTS = 2;                       // tiles in flight per CTA, one epilogue warpgroup each
constexpr int TB_THREADS = 128 * TB_SLOTS + 96;   // + one MMA-issuer warp per slot + the weight-producer / allocator warp
constexpr int SLOT_COLS = 144;                    // TMEM columns per slot: D1 half (64) + GH half (64) + GU (16)
constexpr int COL_D1 = 0, COL_GH = 64, COL_GU = 128, COL_ACC = TB_SLOTS * SLOT_COLS;
constexpr int LBO1 = 128, SBO1 = 256;             // the forward kernel's B1 image (K-major)
constexpr int REC = 640;                          // bytes of one 8-sample record of the A1/G2 image
constexpr int OFF_G2 = 0, OFF_Z = 256, OFF_A1 = 384;
constexpr int MAX_STAGES = 3;

struct TbDims {
  int K, L, d0, d1, Hp, nets, n_nets;
  int b_layer_bytes, b1_off, b2_off;
  int tab_pi, tab_cond, tab_trans, n_tables;
  int n_grad;                                    // floats per partial row
  int n_stages;                                  // weight ring depth (one stage = one layer's B1 and B2 images)
  // shared memory (bytes): ring | tables | per slot {records, h image, ghm image, act, gact} | gb2 | red | barriers
  int sm_tab, sm_slot, sm_slot_stride, sm_h, sm_ghm, sm_act, sm_gact, sm_gb2, sm_red, sm_bar, sm_total;
};

bool tb_dims(const CnfDims& d, TbDims* t) {
  TcDims f;
  if (!cnf_tc_dims(d, &f)) return false;
  if (d.d1 + 1 > 8) return false;                          // A1 must fit one 8-column block (shared accumulator)
  if (COL_ACC + d.L * d.n_nets * 16 > 512) return false;    // weight-gradient accumulators live in TMEM
  t->K = d.K; t->L = d.L; t->d0 = d.d0; t->d1 = d.d1; t->Hp = f.Hp; t->nets = d.nets; t->n_nets = d.n_nets;
  t->b_layer_bytes = f.b_layer_bytes; t->b1_off = f.b1_off; t->b2_off = f.b2_off;
  t->tab_pi = d.tab_pi; t->tab_cond = d.tab_cond; t->tab_trans = d.tab_trans; t->n_tables = d.n_tables;
  t->n_grad = d.L * d.n_nets * 128 * 16 + d.L * 16;
  for (int ns = MAX_STAGES; ns >= 2; --ns) {
    t->n_stages = ns;
    int off = ns * 2 * f.b_layer_bytes;
    off = (off + 127) / 128 * 128;
    t->sm_tab = off; off += (d.n_tables * 4 + 127) / 128 * 128;
    off = (off + 1023) / 1024 * 1024;
    t->sm_slot = off;
    {
      int o = 16 * REC;                         // records first
      o = (o + 1023) / 1024 * 1024;
      t->sm_h = o; o += TILE_M * f.Hp * 2;
      t->sm_ghm = o; o += TILE_M * f.Hp * 2;
      t->sm_act = o; o += d.K * TILE_M * 4;
      t->sm_gact = o; o += d.K * TILE_M * 4;
      t->sm_slot_stride = (o + 1023) / 1024 * 1024;
    }
    off += TB_SLOTS * t->sm_slot_stride;
    t->sm_gb2 = off; off += 4 * TB_SLOTS * d.L * 16 * 4;
    t->sm_red = off; off += 4 * 4 * TB_SLOTS * 8;
    t->sm_bar = off; off += 256;
    t->sm_total = off;
    if (t->sm_total <= 227 * 1024) return true;
  }
  return false;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void epi_sync_all() { asm volatile("bar.sync 3, %0;" ::"n"(128 * TB_SLOTS) : "memory"); }

// SH = 1: compile-time shape of BASELINE configs C2/C3/C5 (K = 10, d0 = d1 = 5, 128 hidden units, both
// nets); SH = 0: every covered shape, read from TbDims.
template <int SH>
__global__ void __launch_bounds__(TB_THREADS, 1)
flow_tcb_kernel(TbDims p, const uint8_t* __restrict__ blob, const int* __restrict__ tables,
                const float* __restrict__ zin, const float* __restrict__ logdet, const float* __restrict__ tape,
                const int64_t* __restrict__ labels, float* __restrict__ partials, double* __restrict__ loss_acc,
                int64_t N, float eps, float gamma, float inv_n, int do_bwd) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int* tab = reinterpret_cast<int*>(smem + p.sm_tab);
  float* gb2 = reinterpret_cast<float*>(smem + p.sm_gb2);     // [epilogue warp][L][16]
  double* red = reinterpret_cast<double*>(smem + p.sm_red);   // [4 sums][epilogue warps]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.sm_bar);
  // per slot s:  ag_ready (128 arrivals, once per layer)   epilogue -> MMA: A1/G2 records written
  //              t1_done  (commit, once per half phase)    MMA -> epilogue: D1 / GH halves complete
  //              hg_ready (128 arrivals, per half phase)   epilogue -> MMA: 64 more columns of the h / ghm images
  //              t4_done  (commit, once per net phase)     MMA -> epilogue: GU update and weight-gradient MMAs done
  // weight ring: full[st] (expect_tx, producer -> MMA), empty[st] (one commit per slot, MMA -> producer)
  uint64_t* ag_ready = bars;
  uint64_t* t1_done = bars + TB_SLOTS;
  uint64_t* hg_ready = bars + 2 * TB_SLOTS;
  uint64_t* t4_done = bars + 3 * TB_SLOTS;
  uint64_t* w_full = bars + 4 * TB_SLOTS;
  uint64_t* w_empty = w_full + MAX_STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(w_empty + MAX_STAGES);
  volatile uint32_t* mid = tmem_ptr + 4;          // [slot] net phases whose first half this slot has passed
  constexpr int MMA_WARP0 = 4 * TB_SLOTS, PROD_WARP = 4 * TB_SLOTS + TB_SLOTS;

  // ---- one-time setup ---------------------------------------------------------------------
  for (int i = tid; i < p.n_tables; i += TB_THREADS) tab[i] = tables[i];
  for (int sl = 0; sl < TB_SLOTS; ++sl) {       // zero blocks of the records stay zero for the whole launch
    uint4* z4 = reinterpret_cast<uint4*>(smem + p.sm_slot + sl * p.sm_slot_stride);
    for (int i = tid; i < 16 * REC / 16; i += TB_THREADS) z4[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  for (int i = tid; i < 4 * TB_SLOTS * p.L * 16; i += TB_THREADS) gb2[i] = 0.f;
  if (tid == 0) {
    mid[0] = 0u; mid[1] = 0u;
    for (int sl = 0; sl < TB_SLOTS; ++sl) {
      mbar_init(ag_ready + sl, 128);
      mbar_init(t1_done + sl, 1);
      mbar_init(hg_ready + sl, 128);
      mbar_init(t4_done + sl, 1);
    }
    for (int st = 0; st < MAX_STAGES; ++st) { mbar_init(w_full + st, 1); mbar_init(w_empty + st, TB_SLOTS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == PROD_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int Hp = SH ? 128 : p.Hp, n_ph = SH ? 2 : p.n_nets;
  const int D0 = SH ? 5 : p.d0, D1 = SH ? 5 : p.d1;
  if (warp < 4 && do_bwd) {                      // the weight-gradient accumulators start at zero
    uint32_t zero[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) zero[i] = 0u;
    for (int a = 0; a < p.L * n_ph; ++a) tmem_st16(tmem_base + ((uint32_t)(warp * 32) << 16) + COL_ACC + a * 16, zero);
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const int64_t ntiles = (N + TILE_M - 1) / TILE_M;
  const int G = gridDim.x;
  const int64_t my_tiles = (ntiles - blockIdx.x + G - 1) / G;          // tiles blockIdx.x, +G, +2G, ...
  const int64_t npairs = (my_tiles + TB_SLOTS - 1) / TB_SLOTS;         // slot s takes tile index TB_SLOTS*i + s
  const int n_half = (Hp + 63) / 64;
  const uint32_t img_sr = (uint32_t)(Hp / 8) * 128;                    // bytes between 8-sample row groups of the images
  const int stage_bytes = 2 * p.b_layer_bytes;

  if (warp == PROD_WARP) {
    // ================================ weight producer ==========================================
    if (lane == 0 && do_bwd) {
      const int64_t n_items = npairs * p.L;
      for (int64_t it = 0; it < n_items; ++it) {
        const int st = (int)(it % p.n_stages);
        const uint32_t round = (uint32_t)(it / p.n_stages);
        if (round > 0) mbar_wait_backoff(w_empty + st, (round - 1) & 1);
        const int l = p.L - 1 - (int)(it % p.L);
        uint8_t* dst = smem + st * stage_bytes;
        mbar_expect_tx(w_full + st, (uint32_t)stage_bytes);
        bulk_copy_g2s(dst, blob + p.b1_off + (size_t)l * p.b_layer_bytes, (uint32_t)p.b_layer_bytes, w_full + st);
        bulk_copy_g2s(dst + p.b_layer_bytes, blob + p.b2_off + (size_t)l * p.b_layer_bytes, (uint32_t)p.b_layer_bytes, w_full + st);
      }
    }
    __syncwarp();
  } else if (warp >= MMA_WARP0) {
    // ================================ MMA issuers: one per slot ================================
    if (lane == 0 && do_bwd) {
      const int sl = warp - MMA_WARP0;
      const uint32_t smem_base = smem_u32(smem);
      const uint32_t slot_base = smem_base + p.sm_slot + sl * p.sm_slot_stride;
      const uint32_t rec = slot_base;
      const uint32_t h_img = slot_base + p.sm_h, ghm_img = slot_base + p.sm_ghm;
      const uint32_t tm = tmem_base + sl * SLOT_COLS;
      // Descriptors are formed once; inside the loops only their start-address field (16-byte units) moves,
      // by plain 64-bit adds: the issuing thread's instruction stream is on the kernel's critical path.
      const uint64_t a1_k = make_desc(rec + OFF_A1, 128, REC);     // [128 smp x 16] K-major (second k-block = zeros)
      const uint64_t g2_k = make_desc(rec + OFF_G2, 128, REC);     // [128 smp x 16] K-major: g_s | g_t
      const uint64_t ghm_k = make_desc(ghm_img, 128, img_sr);      // ghm image K-major: k-step = 2 core matrices = 256 B
      const uint64_t h_mn = make_desc(h_img, img_sr, 128);         // h / ghm images MN-major (M = hidden unit):
      const uint64_t ghm_mn = make_desc(ghm_img, img_sr, 128);     //   S_mn = 128, S_k = img_sr; k-step = 2*img_sr
      const uint64_t g2p_mn[2] = {make_desc(rec + OFF_G2, REC, 256), make_desc(rec + OFF_G2 + 128, REC, 128)};  // [G2p | 0]
      const uint64_t a1_mn = make_desc(rec + OFF_Z, REC, 128);     // [0 | A1]
      const uint64_t img_kstep = (uint64_t)(2 * img_sr / 16), rec_kstep = (uint64_t)(2 * REC / 16);
      const uint32_t id_t2 = make_idesc_ex(16, 0, 1);
      const uint32_t id_t4 = make_idesc_ex(16, 1, 1);
      uint32_t lc = 0, hc = 0;
      // Anti-phase baton (TB_SLOTS == 2).  Started together the two slots stay in phase: both run their E1s at
      // the same time and then queue their T4 blocks behind each other, and the step costs the sum of epilogue
      // and tensor time instead of their maximum.  So slot 1 starts net phase k only once slot 0 is past the first
      // half of its net phase k, and slot 0 starts k once slot 1 is past the first half of k-1 (counters in shared
      // memory, capped by the partner's total so that a slot without tiles never blocks the other).
      const int64_t cnt_other = (my_tiles + sl) / 2;                    // tiles of the other slot
      const uint32_t np_other = (uint32_t)(cnt_other * p.L * n_ph);
      uint32_t kk = 0;
      for (int64_t i = 0; i < npairs; ++i) {
        const bool has_tile = TB_SLOTS * i + sl < my_tiles;
        for (int li = 0; li < p.L; ++li) {
          const int64_t it = i * p.L + li;
          const int st = (int)(it % p.n_stages);
          mbar_wait_backoff(w_full + st, (uint32_t)(it / p.n_stages) & 1);
          if (!has_tile) { mbar_arrive(w_empty + st); continue; }     // keep the ring's arrival count
          const int l = p.L - 1 - li;
          const uint32_t b1 = smem_base + st * stage_bytes;
          const uint32_t b2 = b1 + p.b_layer_bytes;
          mbar_wait_backoff(ag_ready + sl, lc & 1);
          ++lc;
          tc_fence_after();
#pragma unroll
          for (int ph = 0; ph < 2; ++ph) {
            if (ph >= n_ph) break;
            const uint32_t b1p = b1 + ph * (Hp / 8) * SBO1;      // rows ph*Hp.. of the B1 image
            const uint32_t b2p = b2 + ph * (Hp / 16) * 512;      // k-steps of net ph in the B2 image
            const uint64_t b1_k = make_desc(b1p, LBO1, SBO1);    // B1p K-major: 8 hidden rows = SBO1 bytes
            // B2 image, element (n2, hid): (hid/8)*256 + (n2/8)*128 + (n2%8)*16 + (hid%8)*2
            //   -> MN-major with the hidden unit as the MN index: S_mn = 256, S_k = 128; 16 hidden = 512 B
            const uint64_t w2_mn = make_desc(b2p, 128, 256);
            // B1 image (hid, feat) read MN-major over feat: S_mn = 128, S_k = 256; k-step (16 hidden) = 512 B
            const uint64_t w1_mn = make_desc(b1p, 256, 128);
            if (CNF_TCB_BATON) {
              const uint32_t need = min(sl == 0 ? kk : kk + 1, np_other);
              while (mid[sl ^ 1] < need) __nanosleep(20);
            }
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              if (hf >= n_half) break;
              const int h0 = 64 * hf, w = min(64, Hp - h0);
              // ---- T1 (half): D1 = A1 . B1p[h0..]^T ; GH = G2 . W2p[.., h0..]
              mma_ss(tm + COL_D1, a1_k, b1_k + (uint64_t)((h0 / 8) * (SBO1 / 16)), make_idesc_ex(w, 0, 0), 0u);
              mma_ss(tm + COL_GH, g2_k, w2_mn + (uint64_t)((h0 / 16) * 32), make_idesc_ex(w, 0, 1), 0u);
              tc_commit(t1_done + sl);
              ++hc;
              // ---- T2 (half): GU += ghm . W1p
              mbar_wait_backoff(hg_ready + sl, (hc - 1) & 1);
              tc_fence_after();
              if (CNF_TCB_BATON && hf == (CNF_TCB_BATON == 2 ? n_half - 1 : 0)) mid[sl] = ++kk;
#pragma unroll
              for (int jj = 0; jj < ((CNF_TCB_EXP & 32) ? 0 : 4); ++jj) {
                const int j = h0 / 16 + jj;
                if (j < (h0 + w) / 16)
                  mma_ss(tm + COL_GU, ghm_k + (uint64_t)(j * 16), w1_mn + (uint64_t)(j * 32), id_t2, (ph > 0 || j > 0) ? 1u : 0u);
              }
            }
            // ---- T4: weight gradients, contraction over the tile's 128 samples (8 k-steps of 16)
            const uint32_t acc = tmem_base + COL_ACC + (l * n_ph + ph) * 16;
#pragma unroll
            for (int j = 0; j < ((CNF_TCB_EXP & 16) ? 0 : 8); ++j) mma_ss(acc, h_mn + j * img_kstep, g2p_mn[ph] + j * rec_kstep, id_t4, 1u);
#pragma unroll
            for (int j = 0; j < ((CNF_TCB_EXP & 16) ? 0 : 8); ++j) mma_ss(acc, ghm_mn + j * img_kstep, a1_mn + j * rec_kstep, id_t4, 1u);
            tc_commit(t4_done + sl);
          }
          tc_commit(w_empty + st);           // this slot no longer reads the stage once everything above completed
        }
      }
    }
    __syncwarp();
  } else {
    // ================================ epilogue warpgroups =====================================
    const int sl = warp >> 2;
    const int t = tid & 127;                              // sample row of the tile == TMEM lane
    const uint32_t tm = tmem_base + sl * SLOT_COLS + ((uint32_t)((warp & 3) * 32) << 16);
    uint8_t* slot_base = smem + p.sm_slot + sl * p.sm_slot_stride;
    uint8_t* rec_row = slot_base + (t >> 3) * REC + (t & 7) * 16;
    float* act = reinterpret_cast<float*>(slot_base + p.sm_act);
    float* gact = reinterpret_cast<float*>(slot_base + p.sm_gact);
    const uint32_t img_row = (uint32_t)(t >> 3) * img_sr + (t & 7) * 16;   // 16-byte slot in core matrix 0
    const uint32_t h_row = smem_u32(slot_base + p.sm_h) + img_row, ghm_row = smem_u32(slot_base + p.sm_ghm) + img_row;
    const int* pi_last = tab + p.tab_pi + p.L * (SH ? 10 : p.K);
    const int K = SH ? 10 : p.K;
    const int s0 = t / K, f0 = t - s0 * K, ds = TILE_M / K, df = TILE_M - ds * K;
    const bool has_s = SH ? true : (p.nets & 1) != 0, has_t = SH ? true : (p.nets & 2) != 0;
    const uint32_t one_bits = 0x3f80u;
    double a_loss = 0.0, a_ce = 0.0, a_ld = 0.0, a_bad = 0.0;
    uint32_t hc = 0, pc = 0;
    // register prefetch of the next tile's inputs (element e = t + 128*i of the row-major z tile)
    constexpr int KMAX = 14;
    float zr[KMAX], ld_r = 0.f;
    int y_r = 0;
    auto fetch_tile = [&](int64_t tile) {
      const int64_t base = tile * TILE_M;
      const float* gp = zin + base * K;
      const int64_t avail = (N - base) * (int64_t)K;
#pragma unroll
      for (int i = 0; i < KMAX; ++i) {
        const int e = t + 128 * i;
        zr[i] = (i < K && e < avail) ? __ldg(gp + e) : 0.f;
      }
      const bool v = base + t < N;
      ld_r = v ? __ldg(logdet + base + t) : 0.f;
      y_r = v ? (int)labels[base + t] : 0;
    };
    float4 tq[4];                                        // tape record of the next layer to process
    auto fetch_tape = [&](int l, int64_t n) {
      if (n < N) {
        const float4* tp = reinterpret_cast<const float4*>(tape + ((size_t)l * N + n) * 16);
        tq[0] = __ldg(tp); tq[1] = __ldg(tp + 1); tq[2] = __ldg(tp + 2); tq[3] = __ldg(tp + 3);
      } else {
        tq[0] = tq[1] = tq[2] = tq[3] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    const int64_t tile0 = blockIdx.x + (int64_t)sl * G, tstep = (int64_t)TB_SLOTS * G;
    if (tile0 < ntiles) fetch_tile(tile0);
    // Slot 1 starts half a net phase behind slot 0.  Started together, the two slots stay in phase: both run their
    // E1s at the same time (sharing the TMEM read port) and then queue their T4 blocks behind each other, so the
    // step costs the sum of epilogue and tensor time instead of their maximum.  The offset persists (equal
    // periods): +13 % on the backward kernel (profiles/microbench/tcb_speed.py; 1500 clk: nothing, 3000-6500: the same).
    if (sl == 1 && CNF_TCB_STAGGER > 0 && my_tiles >= 2 * TB_SLOTS) {
      const long long c0 = clock64();
      while (clock64() - c0 < CNF_TCB_STAGGER) {}
    }
    for (int64_t tile = tile0; tile < ntiles; tile += tstep) {
      const int64_t base = tile * TILE_M;
      const int64_t n = base + t;
      const bool valid = n < N;
      if (do_bwd) fetch_tape(p.L - 1, n);
      // ---- z tile -> act[physical slot][sample] ------------------------------------------------
      {
        int s = s0, f = f0;
#pragma unroll
        for (int i = 0; i < KMAX; ++i) {
          if (i < K) {
            act[pi_last[f] * TILE_M + s] = zr[i];
            s += ds; f += df;
            if (f >= K) { f -= K; ++s; }
          }
        }
      }
      const float ldv = ld_r;
      const int ylab = y_r;
      wg_sync(sl);
      if (tile + tstep < ntiles) fetch_tile(tile + tstep);
      // ---- loss head (calibrators.py:288-291), one thread per sample ----------------------------
      float gld = 0.f;
      if (!(CNF_TCB_EXP & 4)) {
        float mx = -INFINITY;
        for (int j = 0; j < K; ++j) mx = fmaxf(mx, act[j * TILE_M + t]);
        float se = 0.f;
        for (int j = 0; j < K; ++j) se += expf(act[j * TILE_M + t] - mx);
        const int yy = min(max(ylab, 0), K - 1);
        const int py_slot = pi_last[yy];
        const float zy = act[py_slot * TILE_M + t];
        const float inv_se = 1.f / se;
        const float py = expf(zy - mx) * inv_se;
        float ce, coef;
        if (eps == 0.f) { ce = (zy - mx) - logf(se); coef = 1.f; }
        else            { ce = logf(py + eps); coef = py / (py + eps); }
        if (valid) {
          const float tot = ce + gamma * ldv;
          a_loss += (double)tot; a_ce += (double)ce; a_ld += (double)ldv;
          if (!isfinite(tot)) a_bad += 1.0;
        }
        if (do_bwd) {
          const float sc = valid ? -inv_n * coef : 0.f;
          for (int j = 0; j < K; ++j) {     // j walks physical slots here
            const float pj = expf(act[j * TILE_M + t] - mx) * inv_se;
            gact[j * TILE_M + t] = sc * ((j == py_slot ? 1.f : 0.f) - pj);
          }
          gld = valid ? -gamma * inv_n : 0.f;
        }
      }
      if (do_bwd) {
        for (int li = 0; li < p.L; ++li) {
          const int l = p.L - 1 - li;
          const int* cond = tab + p.tab_cond + l * D1;
          const int* trans = tab + p.tab_trans + l * D0;
          // ---- E0 -------------------------------------------------------------------------------
          {
            float gv[16];                      // 0..7: gradient on the first present net's outputs, 8..15: second
#pragma unroll
            for (int q = 0; q < 16; ++q) gv[q] = 0.f;
            if (!(CNF_TCB_EXP & 2)) {
            const float xt[8] = {tq[0].x, tq[0].y, tq[0].z, tq[0].w, tq[1].x, tq[1].y, tq[1].z, tq[1].w};
            const float sv[8] = {tq[2].x, tq[2].y, tq[2].z, tq[2].w, tq[3].x, tq[3].y, tq[3].z, tq[3].w};
            // the gathers of all transformed slots first, then the math, then the stores: interleaved, every load
            // waited for the store in front of it (act, gact and the loads share an element type)
            int psq[8];
            float gyq[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              psq[q] = (q < D0) ? trans[q] * TILE_M + t : t;
              gyq[q] = (q < D0) ? gact[psq[q]] : 0.f;
            }
            float esq[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              gv[q] = 0.f; gv[8 + q] = 0.f; esq[q] = 1.f;
              if (q < D0) {
                const float gy = gyq[q];
                const float es = has_s ? expf(sv[q]) : 1.f;
                const float gs = gy * xt[q] * es + gld;
                esq[q] = es;
                if (has_s) { gv[q] = gs; gv[8 + q] = has_t ? gy : 0.f; }
                else       { gv[q] = gy; }
              }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (q < D0) {
                act[psq[q]] = xt[q];           // step the tile state back to the input of layer l
                gact[psq[q]] = gyq[q] * esq[q];
              }
            if (l > 0) fetch_tape(l - 1, n);   // in flight while this layer's MMAs and epilogues run
            uint4 v;
            v.x = pack_bf16(gv[0], gv[1]); v.y = pack_bf16(gv[2], gv[3]);
            v.z = pack_bf16(gv[4], gv[5]); v.w = pack_bf16(gv[6], gv[7]);
            *reinterpret_cast<uint4*>(rec_row + OFF_G2) = v;
            v.x = pack_bf16(gv[8], gv[9]); v.y = pack_bf16(gv[10], gv[11]);
            v.z = pack_bf16(gv[12], gv[13]); v.w = pack_bf16(gv[14], gv[15]);
            *reinterpret_cast<uint4*>(rec_row + OFF_G2 + 128) = v;
            float u[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) u[k] = (k < D1) ? act[cond[k] * TILE_M + t] : 0.f;
            v.x = pack_bf16(u[0], u[1]); v.y = pack_bf16(u[2], u[3]);
            v.z = pack_bf16(u[4], u[5]); v.w = pack_bf16(u[6], u[7]);
            {
              const uint32_t ob = one_bits << ((D1 & 1) * 16);
              const int wi = D1 >> 1;
              v.x |= (wi == 0) ? ob : 0u; v.y |= (wi == 1) ? ob : 0u;
              v.z |= (wi == 2) ? ob : 0u; v.w |= (wi == 3) ? ob : 0u;
            }
            *reinterpret_cast<uint4*>(rec_row + OFF_A1) = v;
            } else if (l > 0) fetch_tape(l - 1, n);
            fence_async_smem();
            tc_fence_before();
            mbar_arrive(ag_ready + sl);
            // last-layer bias gradients: the warp's 16 column sums by a halving butterfly (15 shuffles),
            // fp32, no atomics; lane L (even) ends up with the column whose bits are L's bits 4..1
            float w8[8], w4[4], w2[2], w1;
            {
              const bool hi = (lane & 16) != 0;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float send = hi ? gv[i] : gv[8 + i], keep = hi ? gv[8 + i] : gv[i];
                w8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
              }
            }
            {
              const bool hi = (lane & 8) != 0;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float send = hi ? w8[i] : w8[4 + i], keep = hi ? w8[4 + i] : w8[i];
                w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
              }
            }
            {
              const bool hi = (lane & 4) != 0;
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                const float send = hi ? w4[i] : w4[2 + i], keep = hi ? w4[2 + i] : w4[i];
                w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
              }
            }
            {
              const bool hi = (lane & 2) != 0;
              const float send = hi ? w2[0] : w2[1], keep = hi ? w2[1] : w2[0];
              w1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
            }
            w1 += __shfl_xor_sync(0xffffffffu, w1, 1);
            if ((lane & 1) == 0) {
              const int col = ((lane & 16) ? 8 : 0) + ((lane & 8) ? 4 : 0) + ((lane & 4) ? 2 : 0) + ((lane & 2) ? 1 : 0);
              gb2[(warp * p.L + l) * 16 + col] += w1;
            }
          }
          for (int ph = 0; ph < n_ph; ++ph, ++pc) {
            for (int hf = 0; hf < n_half; ++hf, ++hc) {
              // ---- E1 (half): h = relu(D1), ghm = GH * [D1 > 0] -> bf16 shared-memory images -----------
              // The first t1_done of a net phase also says that the previous phase's MMAs no longer read
              // the images.  TMEM loads of the next 16 columns fly while the current 16 are converted.
              const int h0 = 64 * hf, w = min(64, Hp - h0);
              mbar_wait(t1_done + sl, hc & 1);
              tc_fence_after();
              uint32_t rdA[16], rgA[16], rdB[16], rgB[16];
              auto convert = [&](const uint32_t (&rd)[16], const uint32_t (&rg)[16], int c) {
#pragma unroll
                for (int b8 = 0; b8 < 2; ++b8) {
                  uint32_t vh[4], vg[4];
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const float d0v = __uint_as_float(rd[8 * b8 + 2 * i]), d1v = __uint_as_float(rd[8 * b8 + 2 * i + 1]);
                    vh[i] = pack_relu_bf16(d0v, d1v);
                    vg[i] = pack_bf16(d0v > 0.f ? __uint_as_float(rg[8 * b8 + 2 * i]) : 0.f,
                                      d1v > 0.f ? __uint_as_float(rg[8 * b8 + 2 * i + 1]) : 0.f);
                  }
                  const uint32_t off = (uint32_t)((h0 + c) / 8 + b8) * 128;
                  sts128(h_row + off, vh[0], vh[1], vh[2], vh[3]);
                  sts128(ghm_row + off, vg[0], vg[1], vg[2], vg[3]);
                }
              };
              if (!(CNF_TCB_EXP & 1)) {
              tmem_ld16(tm + COL_D1, rdA);
              tmem_ld16(tm + COL_GH, rgA);
              }
              for (int c = 0; c < ((CNF_TCB_EXP & 1) ? 0 : w); c += 32) {
                tmem_wait_ld16(rdA);
                tmem_wait_ld16(rgA);
                if (c + 16 < w) { tmem_ld16(tm + COL_D1 + c + 16, rdB); tmem_ld16(tm + COL_GH + c + 16, rgB); }
                convert(rdA, rgA, c);
                if (c + 16 < w) {
                  tmem_wait_ld16(rdB);
                  tmem_wait_ld16(rgB);
                  if (c + 32 < w) { tmem_ld16(tm + COL_D1 + c + 32, rdA); tmem_ld16(tm + COL_GH + c + 32, rgA); }
                  convert(rdB, rgB, c + 16);
                }
              }
              fence_async_smem();
              tc_fence_before();
              mbar_arrive(hg_ready + sl);
            }
          }
          // ---- E5: gradient on the conditioning logits ---------------------------------------------
          mbar_wait(t4_done + sl, (pc - 1) & 1);
          tc_fence_after();
          if (!(CNF_TCB_EXP & 8)) {
            uint32_t r[16];
            tmem_ld16(tm + COL_GU, r);
            tmem_wait_ld16(r);
            int pk[8];
            float gk[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              pk[k] = (k < D1) ? cond[k] * TILE_M + t : t;
              gk[k] = (k < D1) ? gact[pk[k]] : 0.f;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (k < D1) gact[pk[k]] = gk[k] + __uint_as_float(r[k]);
          }
          tc_fence_before();
        }
      }
      wg_sync(sl);   // every thread is done with act / gact before the next tile overwrites them
    }
    // ---- loss sums -------------------------------------------------------------------------------
    constexpr int EW = 4 * TB_SLOTS;
    if (loss_acc != nullptr) {
      const double v0 = warp_sum_d(a_loss), v1 = warp_sum_d(a_ce), v2 = warp_sum_d(a_ld), v3 = warp_sum_d(a_bad);
      if (lane == 0) { red[0 * EW + warp] = v0; red[1 * EW + warp] = v1; red[2 * EW + warp] = v2; red[3 * EW + warp] = v3; }
    }
    tc_fence_before();
    epi_sync_all();      // both slots are done: every MMA has completed, gb2 and red are final
    tc_fence_after();
    if (loss_acc != nullptr && tid < 4) {
      double a = 0.0;
      for (int wv = 0; wv < EW; ++wv) a += red[tid * EW + wv];
      atomicAdd(loss_acc + tid, a);
    }
    // ---- weight-gradient accumulators -> this CTA's row of the partial buffer ---------------------
    if (do_bwd && sl == 0) {
      float* row = partials + (size_t)blockIdx.x * p.n_grad;
      const uint32_t tacc = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + COL_ACC;
      for (int a = 0; a < p.L * n_ph; ++a) {
        uint32_t r[16];
        tmem_ld16(tacc + a * 16, r);
        tmem_wait_ld16(r);
        float4* dst = reinterpret_cast<float4*>(row + ((size_t)a * 128 + t) * 16);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float4 o = dst[i];
          o.x += __uint_as_float(r[4 * i]); o.y += __uint_as_float(r[4 * i + 1]);
          o.z += __uint_as_float(r[4 * i + 2]); o.w += __uint_as_float(r[4 * i + 3]);
          dst[i] = o;
        }
      }
      float* rb = row + (size_t)p.L * n_ph * 128 * 16;
      for (int i = t; i < p.L * 16; i += 128) {
        float a = 0.f;
        for (int wv = 0; wv < EW; ++wv) a += gb2[(wv * p.L) * 16 + i];
        rb[i] += a;
      }
    }
  }
  // ---- teardown -----------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == PROD_WARP) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

__global__ void tcb_reduce_kernel(const float* __restrict__ partials, const int* __restrict__ gather,
                                  float* __restrict__ flat_grad, int n_grad, int rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_grad) return;
  const int g = gather[i];
  if (g < 0) return;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int r = 0;
  for (; r + 4 <= rows; r += 4) {
    a0 += partials[(size_t)(r + 0) * n_grad + i];
    a1 += partials[(size_t)(r + 1) * n_grad + i];
    a2 += partials[(size_t)(r + 2) * n_grad + i];
    a3 += partials[(size_t)(r + 3) * n_grad + i];
  }
  for (; r < rows; ++r) a0 += partials[(size_t)r * n_grad + i];
  flat_grad[g] = (a0 + a1) + (a2 + a3);
}


// The tail of a single-GPU bf16 training step in one launch: tcb_reduce_kernel's sum (same order), Adam on the flat
// entry the gradient entry maps to (no weight decay: entries without a gradient stay as they are), and the entry's
// place in the tensor-core blob refreshed (bf16 image element or fp32 last-layer bias).
__global__ void tcb_reduce_adam_pack_kernel(const float* __restrict__ partials, const int* __restrict__ gather,
                                            const int* __restrict__ scatter, float* __restrict__ flat,
                                            float* __restrict__ flat_grad, float* __restrict__ m, float* __restrict__ v,
                                            uint8_t* __restrict__ blob, int n_bf16, int bias_off, int n_grad, int rows,
                                            float lr_over_bc1, float inv_sqrt_bc2, float b1, float b2, float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_grad) return;
  const int g = gather[i];
  if (g < 0) return;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int r = 0;
  for (; r + 4 <= rows; r += 4) {
    a0 += partials[(size_t)(r + 0) * n_grad + i];
    a1 += partials[(size_t)(r + 1) * n_grad + i];
    a2 += partials[(size_t)(r + 2) * n_grad + i];
    a3 += partials[(size_t)(r + 3) * n_grad + i];
  }
  for (; r < rows; ++r) a0 += partials[(size_t)r * n_grad + i];
  const float gi = (a0 + a1) + (a2 + a3);
  flat_grad[g] = gi;
  float mi = m[g], vi = v[g];
  const float pn = cnf_adam_entry(flat[g], gi, mi, vi, lr_over_bc1, inv_sqrt_bc2, b1, b2, eps);
  m[g] = mi; v[g] = vi;
  flat[g] = pn;
  const int j = scatter[g];
  if (j >= 0) {
    if (j < n_bf16) reinterpret_cast<__nv_bfloat16*>(blob)[j] = __float2bfloat16_rn(pn);
    else            reinterpret_cast<float*>(blob + bias_off)[j - n_bf16] = pn;
  }
}

int tb_sms() {
  CnfDevInfo di;
  return cnf_dev_info(&di) == CNF_OK ? di.sms : -1;
}

}  // namespace

int cnf_tc_apply_tape(const cnf_flow_desc* desc, const void* packed_tc, const int32_t* tables, const float* x, float* z,
                      float* logdet, float* tape, int64_t N, int inverse, cudaStream_t st);

// Rows of the partial buffer: one per CTA of the persistent grid (fixed upper bound, so that the
// host can size it without a device query).
#define CNF_TCB_ROWS 160

extern "C" int cnf_tc_train_info(const cnf_flow_desc* desc, int64_t* n_grad, int64_t* rows, int64_t* ws_bytes_per_sample) {
  CnfDims d; TbDims t;
  int rc = cnf_make_dims(desc, &d);
  if (rc) return rc;
  if (!n_grad || !rows || !ws_bytes_per_sample) { cnf_set_error("cnf_tc_train_info: null out"); return CNF_E_ARG; }
  if (!tb_dims(d, &t)) { *n_grad = 0; *rows = 0; *ws_bytes_per_sample = 0; return CNF_OK; }
  *n_grad = t.n_grad;
  *rows = CNF_TCB_ROWS;
  *ws_bytes_per_sample = (int64_t)(d.K + 1 + 16 * d.L) * 4;      // z, log-det, tape
  return CNF_OK;
}

// gather[i] = index into flat of partial-row entry i, or -1
extern "C" int cnf_plan_build_tcgrad(const cnf_flow_desc* desc, int32_t* g) {
  CnfDims d; TbDims t;
  int rc = cnf_make_dims(desc, &d);
  if (rc) return rc;
  if (!g) { cnf_set_error("null output"); return CNF_E_ARG; }
  if (!tb_dims(d, &t)) { cnf_set_error("shape not covered by the tensor-core training kernel"); return CNF_E_UNSUPPORTED; }
  const int K = d.K, half = K / 2, H = d.H[0];
  for (int i = 0; i < t.n_grad; ++i) g[i] = -1;
  const long long net_sz = (long long)H * K + H + (long long)K * H + K;
  int32_t* gb = g + (size_t)d.L * d.n_nets * 128 * 16;
  for (int l = 0; l < d.L; ++l) {
    int slot = 0;
    for (int net = 0; net < 2; ++net) {
      if (!(d.nets & (1 << net))) continue;
      const long long base = ((long long)l * d.n_nets + slot) * net_sz;
      const long long w0 = base, b0 = base + (long long)H * K, w1 = b0 + H, b1 = w1 + (long long)K * H;
      int32_t* ga = g + ((size_t)(l * d.n_nets + slot) * 128) * 16;
      for (int h = 0; h < H; ++h) {
        for (int q = 0; q < d.d0; ++q) ga[h * 16 + q] = (int32_t)(w1 + (long long)q * H + h);
        for (int k = 0; k < d.d1; ++k) ga[h * 16 + 8 + k] = (int32_t)(w0 + (long long)h * K + half + k);
        ga[h * 16 + 8 + d.d1] = (int32_t)(b0 + h);
      }
      for (int q = 0; q < d.d0; ++q) gb[l * 16 + slot * 8 + q] = (int32_t)(b1 + q);
      ++slot;
    }
  }
  return CNF_OK;
}

// One training pass on the tensor-core path: per chunk of samples the forward kernel (with tape) and
// the backward kernel above.  workspace: DEVICE, >= chunk * ws_bytes_per_sample; chunk is derived from
// workspace_bytes.  grad_partials_tc: float32 [rows, n_grad], overwritten (NULL = evaluation only).
extern "C" int cnf_nll_train_step_tc(const cnf_flow_desc* desc, const void* packed_tc, const int32_t* tables,
                                     const float* x, const int64_t* y, int64_t N, float eps, float gamma,
                                     float inv_n_total, float* grad_partials_tc, double* loss_acc, void* workspace,
                                     int64_t workspace_bytes, int64_t* rows_used, void* stream) {
  CnfDims d; TbDims t;
  int rc = cnf_make_dims(desc, &d);
  if (rc) return rc;
  if (desc->precision != CNF_PREC_BF16_TC) { cnf_set_error("cnf_nll_train_step_tc needs a CNF_PREC_BF16_TC descriptor"); return CNF_E_ARG; }
  if (!tb_dims(d, &t)) { cnf_set_error("shape not covered by the tensor-core training kernel"); return CNF_E_UNSUPPORTED; }
  if (!packed_tc || !tables || (N > 0 && !y) || !loss_acc || !workspace || N < 0) { cnf_set_error("cnf_nll_train_step_tc: null pointer / negative N"); return CNF_E_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  const int sms = tb_sms();
  if (sms <= 0) { cnf_set_error("no CUDA device"); return CNF_E_CUDA; }
  if (sms > CNF_TCB_ROWS) { cnf_set_error("device has more SMs than partial rows"); return CNF_E_UNSUPPORTED; }
  // rows [0, grid) of the partial buffer are written (one per CTA of the largest launch): only those are
  // cleared here and summed by cnf_grad_reduce_tc -- at calibration-set sizes that is a few dozen rows, not 160
  const int64_t nt_all = (N + TILE_M - 1) / TILE_M;
  const int64_t used = nt_all < sms ? (nt_all > 0 ? nt_all : 1) : sms;
  if (rows_used) *rows_used = used;
  if (grad_partials_tc) CNF_CHECK_CUDA(cudaMemsetAsync(grad_partials_tc, 0, (size_t)used * t.n_grad * sizeof(float), st));
  if (N == 0) return CNF_OK;
  if (!x) { cnf_set_error("null x"); return CNF_E_ARG; }
  const int64_t per_sample = (int64_t)(d.K + 1 + 16 * d.L) * 4;
  int64_t chunk = workspace_bytes / per_sample;
  chunk = chunk / 1024 * 1024;                    // keeps every sub-buffer 16-byte aligned
  if (chunk < 1024) { cnf_set_error("cnf_nll_train_step_tc: workspace smaller than 1024 samples (%lld bytes each)", (long long)per_sample); return CNF_E_ARG; }
  if (chunk > N) chunk = (N + 1023) / 1024 * 1024;
  float* zbuf = reinterpret_cast<float*>(workspace);
  float* ldbuf = zbuf + chunk * d.K;
  float* tapebuf = ldbuf + chunk;
  const bool sh = d.K == 10 && t.Hp == 128 && d.nets == 3 && !cnf_switch(CNF_SW_TC_GENERIC);
  if ((rc = sh ? cnf_kernel_smem(flow_tcb_kernel<1>, t.sm_total) : cnf_kernel_smem(flow_tcb_kernel<0>, t.sm_total))) return rc;
  for (int64_t lo = 0; lo < N; lo += chunk) {
    const int64_t n = (N - lo < chunk) ? N - lo : chunk;
    rc = cnf_tc_apply_tape(desc, packed_tc, tables, x + lo * d.K, zbuf, ldbuf, grad_partials_tc ? tapebuf : nullptr, n, 0, st);
    if (rc) return rc;
    const int64_t ntiles = (n + TILE_M - 1) / TILE_M;
    const int grid = (int)(ntiles < sms ? ntiles : sms);
    if (sh)
      flow_tcb_kernel<1><<<grid, TB_THREADS, t.sm_total, st>>>(t, (const uint8_t*)packed_tc, tables, zbuf, ldbuf, tapebuf,
                                                              y + lo, grad_partials_tc, loss_acc, n, eps, gamma,
                                                              inv_n_total, grad_partials_tc ? 1 : 0);
    else
      flow_tcb_kernel<0><<<grid, TB_THREADS, t.sm_total, st>>>(t, (const uint8_t*)packed_tc, tables, zbuf, ldbuf, tapebuf,
                                                              y + lo, grad_partials_tc, loss_acc, n, eps, gamma,
                                                              inv_n_total, grad_partials_tc ? 1 : 0);
    CNF_CHECK_CUDA(cudaGetLastError());
  }
  return CNF_OK;
}

extern "C" int cnf_grad_reduce_tc(const cnf_flow_desc* desc, const float* grad_partials_tc, int64_t rows_used,
                                  const int32_t* gather_tcgrad, float* flat_grad, void* stream) {
  CnfDims d; TbDims t;
  int rc = cnf_make_dims(desc, &d);
  if (rc) return rc;
  if (!tb_dims(d, &t)) { cnf_set_error("shape not covered by the tensor-core training kernel"); return CNF_E_UNSUPPORTED; }
  if (!grad_partials_tc || !gather_tcgrad || !flat_grad || rows_used < 1 || rows_used > CNF_TCB_ROWS) { cnf_set_error("cnf_grad_reduce_tc: null pointer or bad row count"); return CNF_E_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  CNF_CHECK_CUDA(cudaMemsetAsync(flat_grad, 0, (size_t)d.n_flat * sizeof(float), st));
  tcb_reduce_kernel<<<(t.n_grad + 127) / 128, 128, 0, st>>>(grad_partials_tc, gather_tcgrad, flat_grad, t.n_grad, (int)rows_used);
  CNF_CHECK_CUDA(cudaGetLastError());
  return CNF_OK;
}

extern "C" int cnf_reduce_adam_pack_tc(const cnf_flow_desc* desc, const float* grad_partials_tc, int64_t rows_used,
                                       const int32_t* gather_tcgrad, const int32_t* scatter_tc, float* flat,
                                       float* flat_grad, float* exp_avg, float* exp_avg_sq, void* packed_tc, int64_t step,
                                       float lr, float beta1, float beta2, float eps, void* stream) {
  CnfDims d; TbDims t; TcDims f;
  int rc = cnf_make_dims(desc, &d);
  if (rc) return rc;
  if (!tb_dims(d, &t) || !cnf_tc_dims(d, &f)) { cnf_set_error("shape not covered by the tensor-core training kernel"); return CNF_E_UNSUPPORTED; }
  if (!grad_partials_tc || !gather_tcgrad || !scatter_tc || !flat || !flat_grad || !exp_avg || !exp_avg_sq || !packed_tc ||
      rows_used < 1 || rows_used > CNF_TCB_ROWS || step < 1) { cnf_set_error("cnf_reduce_adam_pack_tc: null pointer, bad row count or step < 1"); return CNF_E_ARG; }
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  tcb_reduce_adam_pack_kernel<<<(t.n_grad + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      grad_partials_tc, gather_tcgrad, scatter_tc, flat, flat_grad, exp_avg, exp_avg_sq, (uint8_t*)packed_tc, f.n_bf16,
      f.bias_off, t.n_grad, (int)rows_used, (float)((double)lr / bc1), (float)(1.0 / sqrt(bc2)), beta1, beta2, eps);
  CNF_CHECK_CUDA(cudaGetLastError());
  return CNF_OK;
}
