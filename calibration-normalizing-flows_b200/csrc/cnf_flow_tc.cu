// bf16 tcgen05 tensor-core kernel of the coupling-flow stack (placeholder until the kernel lands).
#include <cuda_runtime.h>

#include "cnf_common.h"

long long cnf_tc_blob_bytes(const cnf_flow_desc*, const CnfDims&) { return 0; }

extern "C" int cnf_plan_build_tc(const cnf_flow_desc*, int32_t*) {
  cnf_set_error("tensor-core path not available for this shape");
  return CNF_E_UNSUPPORTED;
}
extern "C" int cnf_pack_weights_tc(const cnf_flow_desc*, const float*, const int32_t*, void*, void*) {
  cnf_set_error("tensor-core path not available for this shape");
  return CNF_E_UNSUPPORTED;
}
int cnf_tc_apply(const cnf_flow_desc*, const void*, const int32_t*, const float*, float*, float*, int64_t, int,
                 cudaStream_t) {
  cnf_set_error("tensor-core path not available for this shape");
  return CNF_E_UNSUPPORTED;
}
