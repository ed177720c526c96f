// Parameter plumbing around the flow kernels (sm_100a): flat <-> packed gather/scatter,
// gradient-partial reduction, and the optimiser updates.  All O(P) with P <= a few MB;
// one launch each so a whole training step can be captured in a CUDA graph.
//
// Reference arithmetic restated here (third-party torch semantics the reference calls):
//   torch.optim.Adam.step    calibrators.py:259,295; run_experiment3D.py:55-57,135
//   torch.optim.SGD.step     run_experiment3D.py:59-61,135
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "cnf_common.h"

namespace {

__global__ void gather_kernel(const float* __restrict__ flat, const int* __restrict__ gather, float* __restrict__ packed,
                              int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int g = gather[i];
    packed[i] = g >= 0 ? flat[g] : 0.f;
  }
}

__global__ void gather_bf16_kernel(const float* __restrict__ flat, const int* __restrict__ gather,
                                   __nv_bfloat16* __restrict__ packed, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int g = gather[i];
    float v = 0.f;
    if (g >= 0) v = flat[g];
    else if (g == -2) v = 1.f;       // constant-one entries (bias folding)
    packed[i] = __float2bfloat16_rn(v);
  }
}

// flat_grad[gather[i]] = sum_rows partials[row][i]; fixed summation order (deterministic given the
// partials).  flat_grad must be zeroed first (dead entries keep exactly 0).
// Block = 32 columns x 8 row groups: thread (c, rg) sums rows rg, rg+8, ... of column c with 8 loads in
// flight (the row walk is latency-bound), the 8 group sums are folded through shared memory in a fixed order.
constexpr int RED_RG = 8;
__global__ void __launch_bounds__(32 * RED_RG)
grad_reduce_kernel(const float* __restrict__ partials, const int* __restrict__ gather,
                   float* __restrict__ flat_grad, int n_packed, int rows) {
  __shared__ float part[RED_RG][32];
  const int c = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + c;
  float a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = 0.f;
  if (i < n_packed) {
    for (int r = rg; r < rows; r += 8 * RED_RG) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int rr = r + k * RED_RG;
        if (rr < rows) a[k] += partials[(size_t)rr * n_packed + i];
      }
    }
  }
  part[rg][c] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  __syncthreads();
  if (rg == 0 && i < n_packed) {
    const int g = gather[i];
    if (g >= 0) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < RED_RG; ++k) t += part[k][c];
      flat_grad[g] = t;
    }
  }
}

// The whole tail of a single-GPU fp32 training step in one launch: the partial-row reduction above, Adam on the
// entry the column maps to, and the refreshed packed weight (the gather map is one-to-one on live entries, checked
// by the caller).  Same summation order and the same update expression as grad_reduce_kernel + adam_kernel +
// gather_kernel, so the results are bitwise those of the three launches.  Entries of flat the kernels never read
// (masked weight columns) have zero gradient and, without weight decay, Adam leaves them untouched: they are skipped.
__global__ void __launch_bounds__(32 * RED_RG)
reduce_adam_pack_kernel(const float* __restrict__ partials, const int* __restrict__ gather, float* __restrict__ flat,
                        float* __restrict__ flat_grad, float* __restrict__ m, float* __restrict__ v,
                        float* __restrict__ packed, int n_packed, int rows, float lr_over_bc1, float inv_sqrt_bc2,
                        float b1, float b2, float eps) {
  __shared__ float part[RED_RG][32];
  const int c = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + c;
  float a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = 0.f;
  if (i < n_packed) {
    for (int r = rg; r < rows; r += 8 * RED_RG) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int rr = r + k * RED_RG;
        if (rr < rows) a[k] += partials[(size_t)rr * n_packed + i];
      }
    }
  }
  part[rg][c] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  __syncthreads();
  if (rg == 0 && i < n_packed) {
    const int g = gather[i];
    if (g >= 0) {
      float gi = 0.f;
#pragma unroll
      for (int k = 0; k < RED_RG; ++k) gi += part[k][c];
      flat_grad[g] = gi;
      float mi = m[g], vi = v[g];
      const float pn = cnf_adam_entry(flat[g], gi, mi, vi, lr_over_bc1, inv_sqrt_bc2, b1, b2, eps);
      m[g] = mi; v[g] = vi;
      flat[g] = pn;
      packed[i] = pn;
    }
  }
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float lr_over_bc1, float inv_sqrt_bc2, float b1, float b2,
                            float eps, float wd) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float gi = g[i];
  const float pi = p[i];
  if (wd != 0.f) gi = fmaf(wd, pi, gi);
  float mi = m[i], vi = v[i];
  p[i] = cnf_adam_entry(pi, gi, mi, vi, lr_over_bc1, inv_sqrt_bc2, b1, b2, eps);
  m[i] = mi; v[i] = vi;
}

// CUDA-graph friendly Adam: the step count lives in device memory, so a captured training step can be
// replayed (a host-computed bias correction would be baked into the graph at capture time).
__global__ void adam_prep_kernel(long long* __restrict__ step, float* __restrict__ coef, float lr, float b1, float b2) {
  const long long t = *step + 1;
  *step = t;
  const double bc1 = 1.0 - pow((double)b1, (double)t);
  const double bc2 = 1.0 - pow((double)b2, (double)t);
  coef[0] = (float)((double)lr / bc1);
  coef[1] = (float)(1.0 / sqrt(bc2));
}

__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, int64_t n, const float* __restrict__ coef, float b1, float b2,
                                float eps, float wd) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float lr_over_bc1 = coef[0], inv_sqrt_bc2 = coef[1];
  float gi = g[i];
  const float pi = p[i];
  if (wd != 0.f) gi = fmaf(wd, pi, gi);
  float mi = m[i], vi = v[i];
  p[i] = cnf_adam_entry(pi, gi, mi, vi, lr_over_bc1, inv_sqrt_bc2, b1, b2, eps);
  m[i] = mi; v[i] = vi;
}

__global__ void sgd_kernel(float* __restrict__ p, const float* __restrict__ g, int64_t n, float lr, float wd) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float gi = g[i];
  if (wd != 0.f) gi = fmaf(wd, p[i], gi);
  p[i] = p[i] - lr * gi;
}

}  // namespace

extern "C" int cnf_pack_weights(const cnf_flow_desc* desc, const float* flat, const int32_t* gather, float* packed,
                                void* stream) {
  CnfDims d;
  int rc = cnf_make_dims(desc, &d);
  if (rc) return rc;
  if (!flat || !gather || !packed) { cnf_set_error("cnf_pack_weights: null pointer"); return CNF_E_ARG; }
  gather_kernel<<<(d.n_packed + 255) / 256, 256, 0, (cudaStream_t)stream>>>(flat, gather, packed, d.n_packed);
  CNF_CHECK_CUDA(cudaGetLastError());
  return CNF_OK;
}

int cnf_pack_bf16(const float* flat, const int32_t* gather, void* packed, int n, cudaStream_t st) {
  gather_bf16_kernel<<<(n + 255) / 256, 256, 0, st>>>(flat, gather, (__nv_bfloat16*)packed, n);
  CNF_CHECK_CUDA(cudaGetLastError());
  return CNF_OK;
}

extern "C" int cnf_grad_reduce(const cnf_flow_desc* desc, const float* grad_partials, const int32_t* gather,
                               float* flat_grad, void* stream) {
  CnfDims d;
  int rc = cnf_make_dims(desc, &d);
  if (rc) return rc;
  if (!grad_partials || !gather || !flat_grad) { cnf_set_error("cnf_grad_reduce: null pointer"); return CNF_E_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  CNF_CHECK_CUDA(cudaMemsetAsync(flat_grad, 0, (size_t)d.n_flat * sizeof(float), st));
  grad_reduce_kernel<<<(d.n_packed + 31) / 32, 32 * RED_RG, 0, st>>>(grad_partials, gather, flat_grad, d.n_packed, d.grad_rows_max);
  CNF_CHECK_CUDA(cudaGetLastError());
  return CNF_OK;
}

extern "C" int cnf_grad_reduce_rows(const cnf_flow_desc* desc, const float* grad_partials, int64_t rows_used,
                                    const int32_t* gather, float* flat_grad, void* stream) {
  CnfDims d;
  int rc = cnf_make_dims(desc, &d);
  if (rc) return rc;
  if (!grad_partials || !gather || !flat_grad) { cnf_set_error("cnf_grad_reduce_rows: null pointer"); return CNF_E_ARG; }
  if (rows_used < 0 || rows_used > d.grad_rows_max) { cnf_set_error("cnf_grad_reduce_rows: rows_used %lld out of range", (long long)rows_used); return CNF_E_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  CNF_CHECK_CUDA(cudaMemsetAsync(flat_grad, 0, (size_t)d.n_flat * sizeof(float), st));
  if (rows_used == 0) return CNF_OK;
  grad_reduce_kernel<<<(d.n_packed + 31) / 32, 32 * RED_RG, 0, st>>>(grad_partials, gather, flat_grad, d.n_packed, (int)rows_used);
  CNF_CHECK_CUDA(cudaGetLastError());
  return CNF_OK;
}

extern "C" int cnf_reduce_adam_pack_rows(const cnf_flow_desc* desc, const float* grad_partials, int64_t rows_used,
                                        const int32_t* gather, float* flat, float* flat_grad, float* exp_avg,
                                        float* exp_avg_sq, float* packed, int64_t step, float lr, float beta1,
                                        float beta2, float eps, void* stream) {
  CnfDims d;
  int rc = cnf_make_dims(desc, &d);
  if (rc) return rc;
  if (!grad_partials || !gather || !flat || !flat_grad || !exp_avg || !exp_avg_sq || !packed || step < 1) { cnf_set_error("cnf_reduce_adam_pack_rows: null pointer or step < 1"); return CNF_E_ARG; }
  if (rows_used < 1 || rows_used > d.grad_rows_max) { cnf_set_error("cnf_reduce_adam_pack_rows: rows_used %lld out of range", (long long)rows_used); return CNF_E_ARG; }
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  reduce_adam_pack_kernel<<<(d.n_packed + 31) / 32, 32 * RED_RG, 0, (cudaStream_t)stream>>>(
      grad_partials, gather, flat, flat_grad, exp_avg, exp_avg_sq, packed, d.n_packed, (int)rows_used,
      (float)((double)lr / bc1), (float)(1.0 / sqrt(bc2)), beta1, beta2, eps);
  CNF_CHECK_CUDA(cudaGetLastError());
  return CNF_OK;
}

extern "C" int cnf_adam_step(float* params, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                             int64_t step, float lr, float beta1, float beta2, float eps, float weight_decay,
                             void* stream) {
  if (!params || !grad || !exp_avg || !exp_avg_sq || n < 0 || step < 1) { cnf_set_error("cnf_adam_step: bad argument"); return CNF_E_ARG; }
  if (n == 0) return CNF_OK;
  // bias corrections in double on the host, as torch does for python-float steps
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float lr_over_bc1 = (float)((double)lr / bc1);
  const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(params, grad, exp_avg, exp_avg_sq, n,
                                                                             lr_over_bc1, inv_sqrt_bc2, beta1, beta2,
                                                                             eps, weight_decay);
  CNF_CHECK_CUDA(cudaGetLastError());
  return CNF_OK;
}

extern "C" int cnf_adam_step_dev(float* params, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                                 int64_t* step_dev, float* coef_dev, float lr, float beta1, float beta2, float eps,
                                 float weight_decay, void* stream) {
  if (!params || !grad || !exp_avg || !exp_avg_sq || !step_dev || !coef_dev || n < 0) { cnf_set_error("cnf_adam_step_dev: bad argument"); return CNF_E_ARG; }
  if (n == 0) return CNF_OK;
  cudaStream_t st = (cudaStream_t)stream;
  adam_prep_kernel<<<1, 1, 0, st>>>(reinterpret_cast<long long*>(step_dev), coef_dev, lr, beta1, beta2);
  adam_dev_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(params, grad, exp_avg, exp_avg_sq, n, coef_dev, beta1, beta2,
                                                             eps, weight_decay);
  CNF_CHECK_CUDA(cudaGetLastError());
  return CNF_OK;
}

extern "C" int cnf_sgd_step(float* params, const float* grad, int64_t n, float lr, float weight_decay, void* stream) {
  if (!params || !grad || n < 0) { cnf_set_error("cnf_sgd_step: bad argument"); return CNF_E_ARG; }
  if (n == 0) return CNF_OK;
  sgd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(params, grad, n, lr, weight_decay);
  CNF_CHECK_CUDA(cudaGetLastError());
  return CNF_OK;
}


// TorchFlowCalibrator.fit's full-batch epoch loop (calibrators.py:283-317) enqueued from C -- no host-language work
// between the launches: `epochs` times the fused NLL training pass (cnf_nll_train_step_rows) and the one-launch
// optimiser tail (cnf_reduce_adam_pack_rows).  The reference evaluates the whole set after every update (:297-317);
// with the full batch that number IS the loss the next epoch's forward computes on the same weights and samples, so
// epoch e's sums are taken from the training pass of epoch e + 1 and only the last epoch gets its own evaluation pass.
// Same kernels in the same order as the step-by-step calls: the parameters and the optimiser state are bitwise the same
// (the loss sums are float64 atomic adds of per-CTA sums: equal up to their order).
extern "C" int cnf_fit_full_batch(const cnf_flow_desc* desc, void* packed, const int32_t* tables, const float* x,
                                  const int64_t* y, int64_t N, float eps, float gamma, float inv_n_total,
                                  float* grad_partials, const int32_t* gather, float* flat, float* flat_grad,
                                  float* exp_avg, float* exp_avg_sq, int64_t steps_before, float lr, float beta1,
                                  float beta2, float adam_eps, int64_t epochs, double* hist, double* scratch4,
                                  void* stream) {
  if (epochs < 0 || steps_before < 0 || N < 1) { cnf_set_error("cnf_fit_full_batch: epochs, steps_before >= 0 and N >= 1"); return CNF_E_ARG; }
  if (!hist || !scratch4 || !grad_partials) { cnf_set_error("cnf_fit_full_batch: null pointer"); return CNF_E_ARG; }
  if (epochs == 0) return CNF_OK;
  CNF_CHECK_CUDA(cudaMemsetAsync(scratch4, 0, 4 * sizeof(double), (cudaStream_t)stream));
  CNF_CHECK_CUDA(cudaMemsetAsync(hist, 0, (size_t)epochs * 4 * sizeof(double), (cudaStream_t)stream));
  for (int64_t e = 0; e < epochs; ++e) {
    int64_t used = 0;
    int rc = cnf_nll_train_step_rows(desc, packed, tables, x, y, N, eps, gamma, inv_n_total, grad_partials,
                                     e == 0 ? scratch4 : hist + 4 * (e - 1), &used, stream);
    if (rc) return rc;
    rc = cnf_reduce_adam_pack_rows(desc, grad_partials, used, gather, flat, flat_grad, exp_avg, exp_avg_sq,
                                   reinterpret_cast<float*>(packed), steps_before + e + 1, lr, beta1, beta2, adam_eps, stream);
    if (rc) return rc;
  }
  int64_t used = 0;
  return cnf_nll_train_step_rows(desc, packed, tables, x, y, N, eps, gamma, inv_n_total, nullptr, hist + 4 * (epochs - 1),
                                 &used, stream);
}
