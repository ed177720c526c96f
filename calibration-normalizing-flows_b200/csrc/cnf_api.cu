// extern "C" entry points of the flow itself: precision dispatch between the fp32 CUDA-core
// kernels (cnf_flow_fp32.cu) and the bf16 tcgen05 kernel (cnf_flow_tc.cu).  See include/cnf.h.
#include <cuda_runtime.h>

#include "cnf_common.h"

int cnf_fp32_apply(const cnf_flow_desc* desc, const float* packed, const int32_t* tables, const float* x, float* z,
                   float* logdet, float* zs, int64_t N, int inverse, cudaStream_t st);
int cnf_fp32_train(const cnf_flow_desc* desc, const float* packed, const int32_t* tables, const float* x,
                   const int64_t* y, const float* gz, const float* gld, float* gx, float* partials, double* loss_acc,
                   int64_t N, float eps, float gamma, float inv_n, int head, cudaStream_t st);
int cnf_tc_apply(const cnf_flow_desc* desc, const void* packed_tc, const int32_t* tables, const float* x, float* z,
                 float* logdet, int64_t N, int inverse, cudaStream_t st);

static int check_prec(const cnf_flow_desc* desc) {
  if (!desc) { cnf_set_error("null descriptor"); return CNF_E_ARG; }
  if (desc->precision != CNF_PREC_FP32 && desc->precision != CNF_PREC_BF16_TC) {
    cnf_set_error("unknown precision %d", desc->precision);
    return CNF_E_ARG;
  }
  return CNF_OK;
}

extern "C" int cnf_flow_forward(const cnf_flow_desc* desc, const void* packed, const int32_t* tables, const float* x,
                                float* z, float* logdet, float* zs, int64_t N, void* stream) {
  int rc = check_prec(desc);
  if (rc) return rc;
  if (desc->precision == CNF_PREC_BF16_TC) {
    if (zs) { cnf_set_error("intermediate outputs (zs) are only produced by the fp32 path"); return CNF_E_UNSUPPORTED; }
    return cnf_tc_apply(desc, packed, tables, x, z, logdet, N, 0, (cudaStream_t)stream);
  }
  return cnf_fp32_apply(desc, (const float*)packed, tables, x, z, logdet, zs, N, 0, (cudaStream_t)stream);
}

extern "C" int cnf_flow_inverse(const cnf_flow_desc* desc, const void* packed, const int32_t* tables, const float* z,
                                float* x, float* logdet, float* xs, int64_t N, void* stream) {
  int rc = check_prec(desc);
  if (rc) return rc;
  if (desc->precision == CNF_PREC_BF16_TC) {
    if (xs) { cnf_set_error("intermediate outputs (xs) are only produced by the fp32 path"); return CNF_E_UNSUPPORTED; }
    return cnf_tc_apply(desc, packed, tables, z, x, logdet, N, 1, (cudaStream_t)stream);
  }
  return cnf_fp32_apply(desc, (const float*)packed, tables, z, x, logdet, xs, N, 1, (cudaStream_t)stream);
}

extern "C" int cnf_nll_train_step(const cnf_flow_desc* desc, const void* packed, const int32_t* tables, const float* x,
                                  const int64_t* y, int64_t N, float eps, float gamma, float inv_n_total,
                                  float* grad_partials, double* loss_acc, void* stream) {
  int rc = check_prec(desc);
  if (rc) return rc;
  if (desc->precision != CNF_PREC_FP32) { cnf_set_error("training runs on the fp32 path"); return CNF_E_UNSUPPORTED; }
  return cnf_fp32_train(desc, (const float*)packed, tables, x, y, nullptr, nullptr, nullptr, grad_partials, loss_acc, N,
                        eps, gamma, inv_n_total, CNF_HEAD_NLL, (cudaStream_t)stream);
}

extern "C" int cnf_flow_backward(const cnf_flow_desc* desc, const void* packed, const int32_t* tables, const float* x,
                                 const float* g_z, const float* g_logdet, float* g_x, float* grad_partials, int64_t N,
                                 void* stream) {
  int rc = check_prec(desc);
  if (rc) return rc;
  if (desc->precision != CNF_PREC_FP32) { cnf_set_error("backward runs on the fp32 path"); return CNF_E_UNSUPPORTED; }
  return cnf_fp32_train(desc, (const float*)packed, tables, x, nullptr, g_z, g_logdet, g_x, grad_partials, nullptr, N,
                        0.f, 0.f, 0.f, CNF_HEAD_EXTERNAL, (cudaStream_t)stream);
}
