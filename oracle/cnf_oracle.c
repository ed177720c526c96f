/* Plain-C restatement of the reference's coupling-flow forward / inverse / log-det and of its
 * ECE / NLL / accuracy statistics.  TEST INFRASTRUCTURE ONLY: nothing in the product package
 * links or loads this file; tests/ use it (through ctypes) as a second, independent checker next
 * to oracle/flow_oracle.py.  Pinned by tests/test_oracle_golden.py against the fixtures that
 * oracle/make_golden.py produced by running the reference itself.
 *
 * Reference lines followed (relative to the reference root):
 *   MLP forward (ReLU between Linears)            flows/utils.py:26-31
 *   coupling forward, mask[:, K//2:] = 1, flip     flows/flows.py:81-86, 101-112
 *   coupling inverse                               flows/flows.py:114-126
 *   stack forward / inverse                        flows/flows.py:17-37
 *   ECE (right-closed bins), NLL, accuracy         utils/metrics.py:35-73, 6-15, 76-80
 *
 * Parameters: `flat` in the framework's canonical order (per layer: s-net then t-net; per
 * Linear: weight [out,in] row-major, then bias) = the reference state_dict tensors end to end.
 * Dense formulation exactly as the reference: full-width mask*x inputs, full-width outputs.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXH 8

typedef struct {
  int K, L, m, H[MAXH], scale, shift;
  const int32_t* perm; /* [L*K] or NULL */
} cfg_t;

static long net_size(const cfg_t* c) {
  long n = 0;
  int prev = c->K;
  for (int j = 0; j < c->m; ++j) { n += (long)c->H[j] * prev + c->H[j]; prev = c->H[j]; }
  return n + (long)c->K * prev + c->K;
}

/* out[K] = MLP(in[K]) ; scratch holds two hidden vectors */
static void mlp(const cfg_t* c, const double* w, const double* in, double* out, double* scratch, int maxh) {
  const double* cur = in;
  int n_in = c->K;
  double* bufs[2] = {scratch, scratch + maxh};
  for (int j = 0; j <= c->m; ++j) {
    const int n_out = (j == c->m) ? c->K : c->H[j];
    double* dst = (j == c->m) ? out : bufs[j & 1];
    const double* W = w;
    const double* b = w + (long)n_out * n_in;
    for (int o = 0; o < n_out; ++o) {
      double acc = b[o];
      for (int i = 0; i < n_in; ++i) acc += W[(long)o * n_in + i] * cur[i];
      dst[o] = (j == c->m) ? acc : (acc > 0 ? acc : 0);
    }
    w = b + n_out;
    cur = dst;
    n_in = n_out;
  }
}

static int run(const cfg_t* c, const double* flat, const double* x, double* z, double* logdet, double* all,
               long N, int inverse) {
  const int K = c->K, half = K / 2;
  int maxh = K;
  for (int j = 0; j < c->m; ++j) if (c->H[j] > maxh) maxh = c->H[j];
  const long nsz = net_size(c);
  const int n_nets = (c->scale ? 1 : 0) + (c->shift ? 1 : 0);
  double* buf = (double*)malloc(sizeof(double) * (5 * K + 2 * maxh));
  if (!buf) return -1;
  double *v = buf, *xb = buf + K, *s = buf + 2 * K, *t = buf + 3 * K, *tmp = buf + 4 * K, *scratch = buf + 5 * K;
  for (long n = 0; n < N; ++n) {
    memcpy(v, x + n * K, sizeof(double) * K);
    double ld = 0;
    for (int li = 0; li < c->L; ++li) {
      const int l = inverse ? c->L - 1 - li : li;
      const double* wl = flat + (long)l * n_nets * nsz;
      const int32_t* perm = c->perm ? c->perm + (long)l * K : NULL;
      if (inverse) { /* flows/flows.py:115-117: flip, then rev_perm */
        for (int j = 0; j < K; ++j) tmp[j] = v[K - 1 - j];
        if (perm) { for (int j = 0; j < K; ++j) v[perm[j]] = tmp[j]; }   /* z[:, rev_perm] */
        else memcpy(v, tmp, sizeof(double) * K);
      }
      for (int j = 0; j < K; ++j) { xb[j] = (j >= half) ? v[j] : 0.0; s[j] = 0; t[j] = 0; }
      const double* w = wl;
      if (c->scale) { mlp(c, w, xb, s, scratch, maxh); w += nsz; }
      if (c->shift) mlp(c, w, xb, t, scratch, maxh);
      for (int j = 0; j < half; ++j) {
        if (!inverse) { v[j] = v[j] * exp(s[j]) + t[j]; ld += s[j]; }
        else          { v[j] = (v[j] - t[j]) * exp(-s[j]); ld -= s[j]; }
      }
      if (!inverse) { /* flows/flows.py:110-112: perm, then flip */
        if (perm) { for (int j = 0; j < K; ++j) tmp[j] = v[perm[j]]; } else memcpy(tmp, v, sizeof(double) * K);
        for (int j = 0; j < K; ++j) v[j] = tmp[K - 1 - j];
      }
      if (all) memcpy(all + ((long)li * N + n) * K, v, sizeof(double) * K);
    }
    memcpy(z + n * K, v, sizeof(double) * K);
    logdet[n] = ld;
  }
  free(buf);
  return 0;
}

int cnf_oracle_forward(int K, int L, int m, const int* H, int scale, int shift, const int32_t* perm,
                       const double* flat, const double* x, double* z, double* logdet, double* zs, long N) {
  cfg_t c; memset(&c, 0, sizeof(c));
  if (m > MAXH) return -2;
  c.K = K; c.L = L; c.m = m; c.scale = scale; c.shift = shift; c.perm = perm;
  for (int j = 0; j < m; ++j) c.H[j] = H[j];
  return run(&c, flat, x, z, logdet, zs, N, 0);
}

int cnf_oracle_inverse(int K, int L, int m, const int* H, int scale, int shift, const int32_t* perm,
                       const double* flat, const double* z, double* x, double* logdet, double* xs, long N) {
  cfg_t c; memset(&c, 0, sizeof(c));
  if (m > MAXH) return -2;
  c.K = K; c.L = L; c.m = m; c.scale = scale; c.shift = shift; c.perm = perm;
  for (int j = 0; j < m; ++j) c.H[j] = H[j];
  return run(&c, flat, z, x, logdet, xs, N, 1);
}

/* stats[3*bins+3]: per bin count, sum conf, sum correct; then sum -log(p_y+1e-7), #correct, N.
 * Bin i holds low < conf <= high with low = i*width, high = (i+1)*width (utils/metrics.py:60-63). */
int cnf_oracle_metrics(const double* probs, const int64_t* y, long N, int K, int bins, double* stats) {
  memset(stats, 0, sizeof(double) * (3 * bins + 3));
  const double width = 1.0 / bins;
  for (long n = 0; n < N; ++n) {
    const double* p = probs + n * K;
    int pred = 0;
    for (int j = 1; j < K; ++j) if (p[j] > p[pred]) pred = j;
    const double conf = p[pred];
    const int ok = (pred == (int)y[n]);
    for (int i = 0; i < bins; ++i) {
      const double low = i * width, high = (i + 1) * width;
      if (low < conf && conf <= high) { stats[i] += 1; stats[bins + i] += conf; stats[2 * bins + i] += ok; }
    }
    stats[3 * bins] -= log(p[y[n]] + 1e-7);
    stats[3 * bins + 1] += ok;
    stats[3 * bins + 2] += 1;
  }
  return 0;
}
