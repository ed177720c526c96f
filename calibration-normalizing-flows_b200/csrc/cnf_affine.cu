// Per-dimension constant affine layer: AffineConstantLayer
// (flows/flows.py:40-65) z = x*exp(s)+t / x = (z-t)*exp(-s), and TempScaler (flows/utils.py:34-48)
// as the special case s_k = -log|T|.  Pure streaming: 8K bytes per sample, HBM-bound.
#include <cuda_runtime.h>

#include "cnf_common.h"

namespace {

// 128-bit loads and stores, grid-stride; the K-periodic parameters sit in shared memory and the column
// index of a thread's first element is advanced by (stride mod K) per iteration, so the loop has no
// integer division (a 64-bit modulo per element held the first version at 50 % of the copy rate).
__global__ void affine_kernel(const float* __restrict__ x, const float* __restrict__ s, const float* __restrict__ t,
                              float* __restrict__ z, int64_t total, int K, int inverse, int vec4) {
  extern __shared__ float sm[];
  float* es = sm;        // exp(+-s)
  float* tt = sm + K;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float sv = s ? s[k] : 0.f;
    es[k] = expf(inverse ? -sv : sv);
    tt[k] = t ? t[k] : 0.f;
  }
  __syncthreads();
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  auto apply = [&](float v, int k) { return inverse ? (v - tt[k]) * es[k] : fmaf(v, es[k], tt[k]); };
  const int64_t nvec = vec4 ? total / 4 : 0;
  if (nvec > 0) {
    int k = (int)((tid * 4) % K);
    const int kstep = (int)((nthreads * 4) % K);
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* z4 = reinterpret_cast<float4*>(z);
    for (int64_t v = tid; v < nvec; v += nthreads) {
      const float4 a = __ldcs(x4 + v);
      int k1 = k + 1; if (k1 >= K) k1 -= K;
      int k2 = k1 + 1; if (k2 >= K) k2 -= K;
      int k3 = k2 + 1; if (k3 >= K) k3 -= K;
      float4 o;
      o.x = apply(a.x, k); o.y = apply(a.y, k1); o.z = apply(a.z, k2); o.w = apply(a.w, k3);
      __stcs(z4 + v, o);
      k += kstep; if (k >= K) k -= K;
    }
  }
  for (int64_t i = nvec * 4 + tid; i < total; i += nthreads) z[i] = apply(x[i], (int)(i % K));
}

// g_x = g_z*exp(s); g_s[k] += sum_n g_z*x*exp(s); g_t[k] += sum_n g_z   (forward direction)
__global__ void affine_backward_kernel(const float* __restrict__ x, const float* __restrict__ gz,
                                       const float* __restrict__ s, float* __restrict__ gx, float* __restrict__ gs,
                                       float* __restrict__ gt, int64_t N, int K) {
  extern __shared__ float sm[];
  float* es = sm;
  float* acc_s = sm + K;
  float* acc_t = sm + 2 * K;
  for (int k = threadIdx.x; k < K; k += blockDim.x) { es[k] = expf(s ? s[k] : 0.f); acc_s[k] = 0.f; acc_t[k] = 0.f; }
  __syncthreads();
  // thread owns column k = tid % K for rows tid / K + j * rows_per_block: coalesced, no per-element modulo
  const int rows_per_block = blockDim.x / K;
  if (rows_per_block > 0 && threadIdx.x < rows_per_block * K) {
    const int k = threadIdx.x % K, r0 = threadIdx.x / K;
    float as = 0.f, at = 0.f;
    for (int64_t n = (int64_t)blockIdx.x * rows_per_block + r0; n < N; n += (int64_t)gridDim.x * rows_per_block) {
      const float g = gz[n * K + k], xv = x[n * K + k];
      if (gx) gx[n * K + k] = g * es[k];
      as = fmaf(g * xv, es[k], as);
      at += g;
    }
    atomicAdd(acc_s + k, as);
    atomicAdd(acc_t + k, at);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    if (gs) atomicAdd(gs + k, acc_s[k]);
    if (gt) atomicAdd(gt + k, acc_t[k]);
  }
}

}  // namespace

extern "C" int cnf_affine_const(const float* x, const float* s, const float* t, float* z, int64_t N, int32_t K,
                                int32_t inverse, void* stream) {
  if (N < 0 || K < 1 || K > 1024) { cnf_set_error("cnf_affine_const: bad argument (1 <= K <= 1024, as the backward pass)"); return CNF_E_ARG; }
  if (N == 0) return CNF_OK;      // an empty batch may come with null data pointers
  if (!x || !z) { cnf_set_error("cnf_affine_const: null pointer"); return CNF_E_ARG; }
  int dev = 0, sms = 0;
  CNF_CHECK_CUDA(cudaGetDevice(&dev));
  CNF_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t total = N * K;
  const int vec4 = (((uintptr_t)x | (uintptr_t)z) % 16 == 0) ? 1 : 0;
  const int64_t want = (total / (vec4 ? 4 : 1) + 255) / 256;
  const int grid = (int)(want < (int64_t)sms * 8 ? (want > 0 ? want : 1) : (int64_t)sms * 8);
  affine_kernel<<<grid, 256, 2 * K * sizeof(float), (cudaStream_t)stream>>>(x, s, t, z, total, K, inverse, vec4);
  CNF_CHECK_CUDA(cudaGetLastError());
  return CNF_OK;
}

extern "C" int cnf_affine_const_backward(const float* x, const float* g_z, const float* s, float* g_x, float* g_s,
                                         float* g_t, int64_t N, int32_t K, void* stream) {
  if ((N > 0 && (!x || !g_z)) || N < 0 || K < 1 || K > 1024) { cnf_set_error("cnf_affine_const_backward: bad argument"); return CNF_E_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  if (g_s) CNF_CHECK_CUDA(cudaMemsetAsync(g_s, 0, K * sizeof(float), st));
  if (g_t) CNF_CHECK_CUDA(cudaMemsetAsync(g_t, 0, K * sizeof(float), st));
  if (N == 0) return CNF_OK;
  int dev = 0, sms = 0;
  CNF_CHECK_CUDA(cudaGetDevice(&dev));
  CNF_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int nt = K <= 256 ? 256 : 1024;
  const int rows = nt / K;
  const int64_t want = (N + rows - 1) / rows;
  const int grid = (int)(want < (int64_t)sms * 8 ? want : (int64_t)sms * 8);
  affine_backward_kernel<<<grid, nt, 3 * K * sizeof(float), st>>>(x, g_z, s, g_x, g_s, g_t, N, K);
  CNF_CHECK_CUDA(cudaGetLastError());
  return CNF_OK;
}
