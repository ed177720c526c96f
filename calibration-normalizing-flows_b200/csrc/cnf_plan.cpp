// Host-side planner: turns a cnf_flow_desc into the packed-weight gather map and the
// physical-slot index tables.  Pure index logic; runs without a CUDA device.
//
// What it folds (reference file:line):
//   * half-split mask, mask[:, K//2:] = 1                flows/flows.py:81-86
//     -> only the d1 = K - K//2 conditioning columns of each first Linear and the
//        d0 = K//2 transformed rows of each last Linear are kept (the rest never
//        influence the output: inputs are mask*x, outputs are multiplied by 1-mask)
//   * z.flip((1,)) after every layer                      flows/flows.py:112
//   * optional random_flip permutation                    flows/flows.py:92-99,110-111
//     -> data never moves; layer l reads/writes physical slots pi_l(j), with
//        pi_0 = id, pi_{l+1}(j) = pi_l(perm_l[K-1-j])
//   * MLP layout units=[K]+hidden+[K], weight [out,in]    flows/utils.py:13-24
#include <cstdarg>
#include <vector>

#include "cnf_common.h"

static thread_local char g_err[512] = "";

void cnf_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* cnf_last_error(void) { return g_err; }
extern "C" int cnf_version(void) { return CNF_VERSION; }

int cnf_make_dims(const cnf_flow_desc* desc, CnfDims* d) {
  if (!desc || !d) { cnf_set_error("null descriptor"); return CNF_E_ARG; }
  if (desc->K < 2 || desc->K > 4096) { cnf_set_error("K=%d out of range [2,4096]", desc->K); return CNF_E_ARG; }
  if (desc->L < 1 || desc->L > 1024) { cnf_set_error("L=%d out of range [1,1024]", desc->L); return CNF_E_ARG; }
  if (desc->n_hidden < 0 || desc->n_hidden > CNF_MAX_HIDDEN) {
    cnf_set_error("n_hidden=%d out of range [0,%d]", desc->n_hidden, CNF_MAX_HIDDEN);
    return CNF_E_ARG;
  }
  memset(d, 0, sizeof(*d));
  d->K = desc->K; d->L = desc->L; d->m = desc->n_hidden;
  d->d0 = desc->K / 2; d->d1 = desc->K - desc->K / 2;
  d->d0p = cnf_round_up(d->d0, CNF_CH);
  d->nets = (desc->scale ? 1 : 0) | (desc->shift ? 2 : 0);
  d->n_nets = (desc->scale ? 1 : 0) + (desc->shift ? 1 : 0);
  for (int j = 0; j < d->m; ++j) {
    if (desc->hidden[j] < 1 || desc->hidden[j] > 8192) {
      cnf_set_error("hidden[%d]=%d out of range [1,8192]", j, desc->hidden[j]);
      return CNF_E_ARG;
    }
    d->H[j] = desc->hidden[j];
    d->Hp[j] = cnf_round_up(desc->hidden[j], CNF_CH);
    if (d->Hp[j] > d->Hmax) d->Hmax = d->Hp[j];
  }
  // net block layout (floats)
  int off = 0;
  if (d->m == 0) {
    d->w_off[0] = off; off += d->d1 * d->d0p;       // [d1][d0p]   (in, out)
    d->b_off[0] = off; off += d->d0p;
  } else {
    d->w_off[0] = off; off += d->d1 * d->Hp[0];     // [d1][Hp0]   (in, out)
    d->b_off[0] = off; off += d->Hp[0];
    for (int j = 1; j < d->m; ++j) {
      d->w_off[j] = off; off += d->Hp[j - 1] * d->Hp[j];   // [Hp_{j-1}][Hp_j]  (in, out)
      d->b_off[j] = off; off += d->Hp[j];
    }
    d->w_off[d->m] = off; off += d->d0 * d->Hp[d->m - 1];  // [d0][Hp_{m-1}]  (out, in)
    d->b_off[d->m] = off; off += cnf_round_up(d->d0, 4);
  }
  d->net_stride = off;
  d->layer_stride = off * d->n_nets;
  long long np = (long long)d->layer_stride * d->L;
  if (np > (1ll << 30)) { cnf_set_error("model too large"); return CNF_E_ARG; }
  d->n_packed = (int)np;
  d->tab_pi = 0;
  d->tab_cond = (d->L + 1) * d->K;
  d->tab_trans = d->tab_cond + d->L * d->d1;
  d->n_tables = d->tab_trans + d->L * d->d0;
  // canonical flat size
  long long per_net = 0;
  {
    int prev = d->K;
    for (int j = 0; j < d->m; ++j) { per_net += (long long)d->H[j] * prev + d->H[j]; prev = d->H[j]; }
    per_net += (long long)d->K * prev + d->K;
  }
  d->n_flat = (int)(per_net * d->n_nets * d->L);
  long long rows = CNF_GRAD_ROWS_MAX;
  if (d->n_packed > 0 && rows * d->n_packed > CNF_GRAD_BUDGET_FLOATS) {
    rows = CNF_GRAD_BUDGET_FLOATS / d->n_packed;
    if (rows < 1) rows = 1;
  }
  d->grad_rows_max = (int)rows;
  d->grad_rows = (int)(rows < CNF_GRAD_ROWS ? rows : CNF_GRAD_ROWS);
  return CNF_OK;
}

extern "C" int cnf_plan_info_get(const cnf_flow_desc* desc, cnf_plan_info* out) {
  CnfDims d;
  int rc = cnf_make_dims(desc, &d);
  if (rc) return rc;
  if (!out) { cnf_set_error("null out"); return CNF_E_ARG; }
  memset(out, 0, sizeof(*out));
  out->n_flat = d.n_flat;
  out->n_packed = d.n_packed;
  out->n_tables = d.n_tables;
  out->n_grad_rows = d.grad_rows_max;
  out->tc_bytes = 0;
  out->d0 = d.d0; out->d1 = d.d1;
  for (int j = 0; j < d.m; ++j) out->hidden_padded[j] = d.Hp[j];
  out->tc_bytes = cnf_tc_blob_bytes(desc, d);
  return CNF_OK;
}

extern "C" int cnf_plan_build(const cnf_flow_desc* desc, int32_t* gather, int32_t* tables) {
  CnfDims d;
  int rc = cnf_make_dims(desc, &d);
  if (rc) return rc;
  if (!gather || !tables) { cnf_set_error("null output"); return CNF_E_ARG; }
  const int K = d.K, half = K / 2;
  // ---- index tables -------------------------------------------------------
  std::vector<int> pi(K), nxt(K);
  for (int j = 0; j < K; ++j) pi[j] = j;
  for (int l = 0; l <= d.L; ++l) {
    for (int j = 0; j < K; ++j) tables[d.tab_pi + l * K + j] = pi[j];
    if (l == d.L) break;
    for (int c = 0; c < d.d1; ++c) tables[d.tab_cond + l * d.d1 + c] = pi[half + c];
    for (int q = 0; q < d.d0; ++q) tables[d.tab_trans + l * d.d0 + q] = pi[q];
    const int32_t* perm = desc->perm ? desc->perm + (size_t)l * K : nullptr;
    if (perm) {
      std::vector<char> seen(K, 0);
      for (int j = 0; j < K; ++j) {
        if (perm[j] < 0 || perm[j] >= K || seen[perm[j]]) {
          cnf_set_error("perm of layer %d is not a permutation of 0..K-1", l);
          return CNF_E_ARG;
        }
        seen[perm[j]] = 1;
      }
    }
    for (int j = 0; j < K; ++j) {
      int src = K - 1 - j;
      nxt[j] = pi[perm ? perm[src] : src];
    }
    pi.swap(nxt);
  }
  // ---- gather map ---------------------------------------------------------
  for (int i = 0; i < d.n_packed; ++i) gather[i] = -1;
  int units[CNF_MAX_HIDDEN + 2];
  units[0] = K;
  for (int j = 0; j < d.m; ++j) units[j + 1] = d.H[j];
  units[d.m + 1] = K;
  long long flat_off = 0;
  for (int l = 0; l < d.L; ++l) {
    int slot = 0;
    for (int net = 0; net < 2; ++net) {
      if (!(d.nets & (1 << net))) continue;
      int32_t* g = gather + (size_t)l * d.layer_stride + (size_t)slot * d.net_stride;
      ++slot;
      for (int j = 0; j <= d.m; ++j) {
        const int in = units[j], out = units[j + 1];
        const long long wbase = flat_off, bbase = flat_off + (long long)in * out;
        flat_off = bbase + out;
        if (d.m == 0) {
          for (int c = 0; c < d.d1; ++c)
            for (int q = 0; q < d.d0; ++q)
              g[d.w_off[0] + c * d.d0p + q] = (int32_t)(wbase + (long long)q * K + half + c);
          for (int q = 0; q < d.d0; ++q) g[d.b_off[0] + q] = (int32_t)(bbase + q);
        } else if (j == 0) {
          for (int c = 0; c < d.d1; ++c)
            for (int h = 0; h < out; ++h)
              g[d.w_off[0] + c * d.Hp[0] + h] = (int32_t)(wbase + (long long)h * K + half + c);
          for (int h = 0; h < out; ++h) g[d.b_off[0] + h] = (int32_t)(bbase + h);
        } else if (j < d.m) {
          for (int i = 0; i < in; ++i)
            for (int o = 0; o < out; ++o)
              g[d.w_off[j] + i * d.Hp[j] + o] = (int32_t)(wbase + (long long)o * in + i);
          for (int o = 0; o < out; ++o) g[d.b_off[j] + o] = (int32_t)(bbase + o);
        } else {
          for (int q = 0; q < d.d0; ++q)
            for (int r = 0; r < in; ++r)
              g[d.w_off[j] + q * d.Hp[d.m - 1] + r] = (int32_t)(wbase + (long long)q * in + r);
          for (int q = 0; q < d.d0; ++q) g[d.b_off[j] + q] = (int32_t)(bbase + q);
        }
      }
    }
  }
  if (flat_off != d.n_flat) { cnf_set_error("internal: flat size mismatch"); return CNF_E_ARG; }
  return CNF_OK;
}
