// Shape plan of the resident-weight tensor-core kernels (forward/inverse in cnf_flow_tc.cu, training
// backward in cnf_flow_tcb.cu): blob offsets of the bf16 UMMA weight images and the forward kernel's
// shared-memory carve-up.  Not part of the public ABI.
#pragma once
#include "cnf_common.h"

struct TcDims {
  int K, L, d0, d1, Hp, nets, n_nets, N1;
  int b_layer_bytes;                 // bytes of one layer's B1 (== B2): N1*16*2
  int b1_off, b2_off, bias_off;      // byte offsets inside the blob
  int n_bf16, n_f32, blob_bytes;
  int tab_pi, tab_cond, tab_trans, n_tables;
  // shared-memory carve-up (bytes)
  int sm_tab, sm_slot, sm_slot_stride, sm_act, sm_raw_in, sm_raw_out, sm_bar, sm_tail, sm_total;
};

// false when the shape is outside the resident-weight kernel's coverage (see cnf_flow_tc.cu)
bool cnf_tc_dims(const CnfDims& d, TcDims* t);
