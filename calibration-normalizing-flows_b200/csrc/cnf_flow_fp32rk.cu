// Register-resident fp32 NLL training step for K = 2 .. 9 classes (cnf_fp32r_dev.cuh; K = 10 lives in
// cnf_flow_fp32r.cu): the reference's own notebook and script shapes -- K = 3, hidden_size [3, 3] / [5, 5], NICE or
// RealNVP, N = 1,500 full-batch fits (notebooks/simulated-predictions-flows.ipynb:214, 224; run_experiment3D.py:33-38)
// -- ran on the generic one-thread-per-sample tile kernel before (170 us per step at N = 1,500, K = 3, 5 x [3, 3]).
// Same arithmetic as the K = 10 kernel: the K logits sit in the same two register arrays of five, positions K does not
// use carry zeros (RegMap).  Reference arithmetic: calibrators.py:287-293, flows/flows.py:101-112, flows/utils.py:26-31.
#include <cuda_runtime.h>

#include "cnf_common.h"
#include "cnf_fp32r_dev.cuh"

namespace {

template <int KK, int NT, int SPT, int MB, bool M2>
int launch_train_k(const CnfDims& d, const float* packed, const int32_t* tables, const float* x, const int64_t* y,
                   float* partials, double* loss_acc, int64_t N, float eps, float gamma, float inv_n, size_t smem_fwd,
                   int sms, int max_smem, int64_t* rows_out, const float* gz_ext, const float* gld_ext, cudaStream_t st) {
  const size_t smem = smem_fwd + (size_t)3 * SPT * RD * NT * sizeof(float);
  if ((long long)smem > max_smem - 1024) return CNF_E_SMEM;
  int rc = cnf_kernel_smem(train_reg10_kernel<NT, SPT, 2, MB, M2, KK>, smem);
  if (rc) return rc;
  int per_sm = 0;
  CNF_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, train_reg10_kernel<NT, SPT, 2, MB, M2, KK>, NT, smem));
  if (per_sm < 1) per_sm = 1;
  const int64_t ntiles = (N + NT * SPT - 1) / (NT * SPT);
  int64_t cap = (int64_t)sms * per_sm;
  if (partials && cap > d.grad_rows_max / (NT / 32)) cap = d.grad_rows_max / (NT / 32);
  if (cap < 1) { cnf_set_error("partial buffer too small for the register-resident training kernel"); return CNF_E_SMEM; }
  const int grid = (int)(ntiles < cap ? ntiles : cap);
  const int64_t rows = (int64_t)grid * (NT / 32);
  if (rows_out) *rows_out = rows;
  if (partials) CNF_CHECK_CUDA(cudaMemsetAsync(partials, 0, (size_t)(rows_out ? rows : d.grad_rows_max) * d.n_packed * sizeof(float), st));
  train_reg10_kernel<NT, SPT, 2, MB, M2, KK><<<grid, NT, smem, st>>>(d, packed, tables, x, y, partials, loss_acc, N, eps, gamma, inv_n, gz_ext, gld_ext);
  CNF_CHECK_CUDA(cudaGetLastError());
  return CNF_OK;
}

template <int KK>
int train_k(const CnfDims& d, const float* packed, const int32_t* tables, const float* x, const int64_t* y,
            float* partials, double* loss_acc, int64_t N, float eps, float gamma, float inv_n, size_t smem_fwd, int sms,
            int max_smem, int64_t* rows_out, const float* gz_ext, const float* gld_ext, cudaStream_t st) {
  // the largest tile that still gives every SM one: 128 x 1 at calibration-set sizes (as for K = 10, [5, 5])
#define TRY(SPT, MB, M2)                                                                                              \
  do {                                                                                                                \
    const int rc = launch_train_k<KK, 128, SPT, MB, M2>(d, packed, tables, x, y, partials, loss_acc, N, eps, gamma, inv_n, \
                                                        smem_fwd, sms, max_smem, rows_out, gz_ext, gld_ext, st);     \
    if (rc != CNF_E_SMEM) return rc;                                                                                  \
  } while (0)
  if (d.m == 2) {
    if (N < 128LL * 2 * sms) TRY(1, 4, true);
    if (N < 128LL * 4 * sms) TRY(2, 4, true);
    TRY(4, 2, true);
    TRY(1, 4, true);
  } else {
    if (N < 128LL * 2 * sms) TRY(1, 4, false);
    if (N < 128LL * 4 * sms) TRY(2, 4, false);
    TRY(4, 2, false);
    TRY(1, 4, false);
  }
#undef TRY
  cnf_set_error("register-resident training kernel: the weights of %d layers do not fit shared memory", d.L);
  return CNF_E_SMEM;
}

}  // namespace

int cnf_fp32rk_train(const CnfDims& d, const float* packed, const int32_t* tables, const float* x, const int64_t* y,
                     float* partials, double* loss_acc, int64_t N, float eps, float gamma, float inv_n, size_t smem_fwd,
                     int sms, int max_smem, int64_t* rows_out, const float* gz_ext, const float* gld_ext, cudaStream_t st) {
#define K_CASE(KK) case KK: return train_k<KK>(d, packed, tables, x, y, partials, loss_acc, N, eps, gamma, inv_n, smem_fwd, sms, max_smem, rows_out, gz_ext, gld_ext, st)
  switch (d.K) {
    K_CASE(2); K_CASE(3); K_CASE(4); K_CASE(5); K_CASE(6); K_CASE(7); K_CASE(8); K_CASE(9);
    default: break;
  }
#undef K_CASE
  cnf_set_error("register-resident training kernel: K = %d is not instantiated", d.K);
  return CNF_E_UNSUPPORTED;
}
