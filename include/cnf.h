/*
 * cnf.h -- C ABI of the B200-native coupling-flow calibration hot path.
 *
 * One shared library (libcnf_b200.so, sm_100a) replaces the arithmetic of the
 * reference's PyTorch modules on this path.  Every entry point is extern "C",
 * takes plain pointers and sizes, returns 0 on success or a negative CNF_E_*
 * code (message via cnf_last_error(), thread-local), never throws, never
 * allocates or frees caller-visible memory, and launches asynchronously on the
 * caller's cudaStream_t (passed as void*).  There is no CPU fallback: a call
 * without a usable CUDA device returns CNF_E_CUDA.
 *
 * Reference interfaces replaced (paths relative to the reference root):
 *   cnf_flow_forward     Flow.forward            flows/flows.py:17-25
 *                        NvpCouplingLayer.forward flows/flows.py:101-112
 *                        MLP.forward              flows/utils.py:26-31
 *   cnf_flow_inverse     Flow.backward            flows/flows.py:27-37
 *                        NvpCouplingLayer.backward flows/flows.py:114-126
 *   cnf_nll_train_step   loss + loss.backward()   calibrators.py:287-293,
 *                                                 run_experiment3D.py:102-107,133-134
 *   cnf_flow_backward    autograd of Flow.forward (torch, third party)
 *   cnf_adam_step        torch.optim.Adam.step    calibrators.py:259,295
 *   cnf_sgd_step         torch.optim.SGD.step     run_experiment3D.py:59-61,135
 *   cnf_flow_predict     Calibrator.predict / predict_post + metrics, one pass
 *                                                 calibrators.py:17, 40-44, 350-353
 *   cnf_metrics          expected_calibration_error / neg_log_likelihood / accuracy
 *                                                 utils/metrics.py:35-73, 6-15, 76-80
 *                        (+ Calibrator.predict tail calibrators.py:40-44, 350-353)
 *
 * Data conventions: logits x, z are float32 [N, K] row-major contiguous; labels
 * are int64 [N]; log-det is float32 [N].  "flat" is the trainable parameter
 * vector in canonical order (per coupling layer: s-net then t-net; per Linear:
 * weight [out,in] row-major then bias), i.e. the reference state_dict tensors
 * laid end to end.  "packed" is the kernel-side copy with the half-split mask,
 * flips and permutations folded in (dead rows/columns dropped).
 */
#ifndef CNF_B200_H
#define CNF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CNF_VERSION 100
#define CNF_MAX_HIDDEN 4

enum {
  CNF_OK = 0,
  CNF_E_ARG = -1,      /* bad descriptor / null pointer / unsupported shape */
  CNF_E_CUDA = -2,     /* CUDA runtime error (no device, launch failure, ...) */
  CNF_E_SMEM = -3,     /* model does not fit the kernel's shared-memory plan */
  CNF_E_UNSUPPORTED = -4
};

enum { CNF_PREC_FP32 = 0, CNF_PREC_BF16_TC = 1 };

/* Loss-head selector for cnf_nll_train_step / cnf_flow_backward */
enum { CNF_HEAD_NLL = 0, CNF_HEAD_EXTERNAL = 1 };

/* cnf_metrics input modes */
enum {
  CNF_METRICS_PROBS = 0,       /* input rows are probabilities                       */
  CNF_METRICS_LOGITS = 1,      /* softmax(z) first (predict_post, calibrators.py:352) */
  CNF_METRICS_CALIBRATED = 2   /* softmax(log(softmax(z)+1e-7) - log_priors), :44     */
};

typedef struct cnf_flow_desc {
  int32_t K;                       /* classes = flow dimension                      */
  int32_t L;                       /* coupling layers                               */
  int32_t n_hidden;                /* hidden layers of each conditioner MLP (0..4)  */
  int32_t hidden[CNF_MAX_HIDDEN];  /* their widths                                  */
  int32_t scale;                   /* 1: s-net present (RealNVP), 0: NICE additive  */
  int32_t shift;                   /* 1: t-net present                              */
  int32_t precision;               /* CNF_PREC_*                                    */
  const int32_t* perm;             /* HOST [L*K] random_flip permutations or NULL   */
} cnf_flow_desc;

typedef struct cnf_plan_info {
  int64_t n_flat;          /* floats in the canonical flat parameter vector         */
  int64_t n_packed;        /* floats in the fp32 packed blob                        */
  int64_t n_tables;        /* int32 entries of the index tables                     */
  int64_t n_grad_rows;     /* rows of the gradient partial buffer (one per CTA or   */
                           /* per warp of the training kernels; <= 1184)            */
  int64_t tc_bytes;        /* bytes of the bf16 tensor-core blob, 0 if shape not    */
                           /* supported by the tensor-core path                     */
  int32_t d0, d1;          /* transformed / conditioning dims per layer             */
  int32_t hidden_padded[CNF_MAX_HIDDEN];
} cnf_plan_info;

const char* cnf_last_error(void);
int cnf_version(void);

/* ---- host-side planning (pure index logic, no device needed) ---------------- */
int cnf_plan_info_get(const cnf_flow_desc* desc, cnf_plan_info* out);
/* gather[i] = index into flat for packed float i, or -1 for structural zero.
 * tables = [pi (L+1)*K | cond L*d1 | trans L*d0] physical-slot index maps.     */
int cnf_plan_build(const cnf_flow_desc* desc, int32_t* gather_host, int32_t* tables_host);
/* Gather map of the tensor-core blob (bf16 UMMA images of the Linears, then the fp32
 * biases added in the epilogues); *n = number of int32 entries, 0 if the shape is not
 * covered.  Covered (forward / inverse): one hidden layer with K <= 126 (any width), or two to
 * four hidden layers of 16..128 units each with K <= 65 (flows/utils.py:6-31).          */
int cnf_tc_gather_len(const cnf_flow_desc* desc, int64_t* n);
int cnf_plan_build_tc(const cnf_flow_desc* desc, int32_t* gather_tc_host);

/* ---- device: parameter packing ----------------------------------------------- */
int cnf_pack_weights(const cnf_flow_desc* desc, const float* flat, const int32_t* gather,
                     float* packed, void* stream);
int cnf_pack_weights_tc(const cnf_flow_desc* desc, const float* flat, const int32_t* gather_tc,
                        void* packed_tc, void* stream);

/* ---- device: the flow --------------------------------------------------------- */
/* zs (optional, may be NULL): float32 [L, N, K], every intermediate output.     */
int cnf_flow_forward(const cnf_flow_desc* desc, const void* packed, const int32_t* tables,
                     const float* x, float* z, float* logdet, float* zs, int64_t N, void* stream);
/* xs (optional): float32 [L, N, K]; xs[L-1] == x.                                */
int cnf_flow_inverse(const cnf_flow_desc* desc, const void* packed, const int32_t* tables,
                     const float* z, float* x, float* logdet, float* xs, int64_t N, void* stream);

/* One fused pass replacing Calibrator.predict (calibrators.py:40-44) on top of
 * TorchFlowCalibrator.predict_post (:350-353) and, with labels, the metrics of utils/metrics.py
 * (:35-73 ECE, :6-15 NLL, :76-80 accuracy) without a second pass over z:
 *   center != 0: x <- x - mean(x, axis=1) per row, in numpy's float32 summation order
 *                (calibrators.py:17, 42); K <= 128
 *   flow forward (+ log-det)
 *   mode CNF_METRICS_LOGITS:     p = softmax(z)                       (predict_post)
 *        CNF_METRICS_CALIBRATED: p = softmax(log(softmax(z)+1e-7) - log_priors)   (predict)
 * Outputs, each optional (NULL = not produced, nothing written): z [N,K] and logdet [N] float32;
 * probs_out [N,K] float64 (calibrated mode only); acc [3*bins+3] float64 ACCUMULATED statistics of p
 * against labels y, laid out as in cnf_metrics (needs y, edges [bins+1], 1 <= bins <= 1024).
 * With only acc requested the pass reads 4K+8 bytes per sample and writes nothing.
 * Shapes: the fp32 path covers every shape whose tile plan fits shared memory; the bf16 path the
 * resident-weight tensor-core kernel (CNF_E_UNSUPPORTED otherwise: compose cnf_flow_forward +
 * cnf_metrics / cnf_calibrated_probs).                                                          */
int cnf_flow_predict(const cnf_flow_desc* desc, const void* packed, const int32_t* tables,
                     const float* x, int64_t N, int32_t center, int32_t mode,
                     const double* log_priors, float* z, float* logdet, double* probs_out,
                     const int64_t* y, int32_t bins, const double* edges, double* acc, void* stream);

/* Host-buffer variant of the two calls above: x_host / z_host / logdet_host are HOST pointers
 * (pinned memory gives full PCIe speed).  The samples stream through the GPU in chunks of `chunk`
 * rows on internal streams so H2D, kernel and D2H of neighbouring chunks overlap; `workspace` is a
 * DEVICE buffer of at least chunk*(2K+1)*4 bytes (up to 4 such slots are used).  Asynchronous:
 * completion is ordered on `stream`.  This is what predict_logits (calibrators.py:330-348) does
 * with tensor.to(dev) / .cpu().                                                              */
int cnf_flow_apply_host(const cnf_flow_desc* desc, const void* packed, const int32_t* tables,
                        const float* x_host, float* z_host, float* logdet_host, int64_t N,
                        int32_t inverse, void* workspace, int64_t workspace_bytes, int64_t chunk,
                        void* stream);

/* ---- device: training --------------------------------------------------------- */
/* One fused pass over N local samples: forward, loss head, backward.
 *   loss_i = -( log(softmax(z_i)[y_i] + eps) + gamma * logdet_i );  eps==0 -> log-softmax.
 * grad_partials: float32 [n_grad_rows, n_packed], overwritten; sum of rows is
 *   d(sum_i loss_i * inv_n_total)/d(packed).  NULL = evaluation only.
 * loss_acc: float64 [4] ACCUMULATED (zero it first): sum(ce_i + gamma*ld_i),
 *   sum(ce_i), sum(ld_i), count of non-finite samples; ce_i = log(p_y+eps).      */
int cnf_nll_train_step(const cnf_flow_desc* desc, const void* packed, const int32_t* tables,
                       const float* x, const int64_t* y, int64_t N, float eps, float gamma,
                       float inv_n_total, float* grad_partials, double* loss_acc, void* stream);
/* Backward for arbitrary upstream gradients (autograd of Flow.forward):
 * g_z [N,K], g_logdet [N] -> g_x [N,K] (may be NULL) and grad_partials.          */
int cnf_flow_backward(const cnf_flow_desc* desc, const void* packed, const int32_t* tables,
                      const float* x, const float* g_z, const float* g_logdet, float* g_x,
                      float* grad_partials, int64_t N, void* stream);
/* flat_grad[gather[i]] = sum over rows of grad_partials[:, i]; dead entries = 0. */
int cnf_grad_reduce(const cnf_flow_desc* desc, const float* grad_partials, const int32_t* gather,
                    float* flat_grad, void* stream);
/* Variants for small batches: only the rows of grad_partials the launch wrote, [0, *rows_used), are cleared
 * and later reduced (a batch of a few thousand samples uses a fraction of the n_grad_rows rows; clearing and
 * summing all of them costs more than the training kernel).  rows_used: HOST out, never NULL.
 * Otherwise identical to cnf_nll_train_step / cnf_flow_backward / cnf_grad_reduce.                       */
int cnf_nll_train_step_rows(const cnf_flow_desc* desc, const void* packed, const int32_t* tables,
                            const float* x, const int64_t* y, int64_t N, float eps, float gamma,
                            float inv_n_total, float* grad_partials, double* loss_acc, int64_t* rows_used,
                            void* stream);
int cnf_flow_backward_rows(const cnf_flow_desc* desc, const void* packed, const int32_t* tables,
                           const float* x, const float* g_z, const float* g_logdet, float* g_x,
                           float* grad_partials, int64_t N, int64_t* rows_used, void* stream);
int cnf_grad_reduce_rows(const cnf_flow_desc* desc, const float* grad_partials, int64_t rows_used,
                         const int32_t* gather, float* flat_grad, void* stream);
/* The tail of a single-GPU fp32 Adam step in ONE launch: cnf_grad_reduce_rows, then torch.optim.Adam (no weight decay:
 * calibrators.py:259 defaults) on every entry a partial column maps to, then the refreshed entry of `packed`
 * (= cnf_grad_reduce_rows + cnf_adam_step + cnf_pack_weights, bitwise).  Replaces loss.backward()'s gradient
 * accumulation + optimizer.step() (calibrators.py:293-295).  `gather` must be one-to-one on its entries >= 0 and
 * `packed` must have been fully packed once; flat_grad entries the kernels never touch are left as they are (zero
 * them once); step is 1-based. */
int cnf_reduce_adam_pack_rows(const cnf_flow_desc* desc, const float* grad_partials, int64_t rows_used,
                              const int32_t* gather, float* flat, float* flat_grad, float* exp_avg,
                              float* exp_avg_sq, float* packed, int64_t step, float lr, float beta1, float beta2,
                              float eps, void* stream);
/* The full-batch epoch loop of TorchFlowCalibrator.fit (calibrators.py:283-317, default batch_size = N) enqueued
 * from C: per epoch cnf_nll_train_step_rows + cnf_reduce_adam_pack_rows, nothing else; hist: float64 [epochs, 4]
 * device, row e = (sum(ce + gamma*ld), sum ce, sum ld, #non-finite) of the evaluation after update e (taken from the
 * next epoch's forward; the last row from one evaluation pass at the end).  steps_before: optimiser steps already
 * taken (Adam's bias correction continues from there); scratch4: float64 [4] device.  The same kernels in the same order as
 * the calls issued one by one: parameters and optimiser state come out bitwise equal. */
int cnf_fit_full_batch(const cnf_flow_desc* desc, void* packed, const int32_t* tables, const float* x,
                       const int64_t* y, int64_t N, float eps, float gamma, float inv_n_total,
                       float* grad_partials, const int32_t* gather, float* flat, float* flat_grad, float* exp_avg,
                       float* exp_avg_sq, int64_t steps_before, float lr, float beta1, float beta2, float adam_eps,
                       int64_t epochs, double* hist, double* scratch4, void* stream);
/* The same one-launch tail for the tensor-core training path: cnf_grad_reduce_tc + cnf_adam_step + cnf_pack_weights_tc.
 * scatter_tc: int32 [n_flat], the position of each flat entry in the gather map of cnf_plan_build_tc (-1: not packed);
 * both maps must be one-to-one on their live entries. */
int cnf_reduce_adam_pack_tc(const cnf_flow_desc* desc, const float* grad_partials_tc, int64_t rows_used,
                            const int32_t* gather_tcgrad, const int32_t* scatter_tc, float* flat, float* flat_grad,
                            float* exp_avg, float* exp_avg_sq, void* packed_tc, int64_t step, float lr, float beta1,
                            float beta2, float eps, void* stream);
/* torch.optim.Adam (L2 weight decay, bias correction as torch); step is 1-based. */
int cnf_adam_step(float* params, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                  int64_t step, float lr, float beta1, float beta2, float eps, float weight_decay,
                  void* stream);
/* Same update with the step count kept on the device (step_dev: int64 [1], incremented by the call;
 * coef_dev: float32 [2] scratch), so that a whole training step can be captured in a CUDA graph and
 * replayed: nothing in it depends on host state.                                                    */
int cnf_adam_step_dev(float* params, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                      int64_t* step_dev, float* coef_dev, float lr, float beta1, float beta2, float eps,
                      float weight_decay, void* stream);
int cnf_sgd_step(float* params, const float* grad, int64_t n, float lr, float weight_decay,
                 void* stream);

/* ---- device: training on the bf16 tensor-core path ------------------------------------------------
 * Same loss and gradient as cnf_nll_train_step, computed by the tcgen05 forward kernel (which also
 * writes a per-layer tape) and a tcgen05 backward kernel.  Shapes: one hidden layer, K <= 14,
 * pad16(H) <= 128, L*nets <= 15; cnf_tc_train_info reports n_grad == 0 otherwise (use the fp32 path).
 *   n_grad:  floats per row of the partial buffer; rows: its row count;
 *   ws_bytes_per_sample: bytes of DEVICE workspace per sample of a chunk (z, log-det, tape).      */
int cnf_tc_train_info(const cnf_flow_desc* desc, int64_t* n_grad, int64_t* rows, int64_t* ws_bytes_per_sample);
/* gather_tcgrad[i] = index into flat of partial-row entry i, or -1; length n_grad (host).         */
int cnf_plan_build_tcgrad(const cnf_flow_desc* desc, int32_t* gather_tcgrad_host);
/* packed_tc: blob of cnf_pack_weights_tc.  The samples are processed in chunks of
 * workspace_bytes / ws_bytes_per_sample (rounded down to 1024) samples.  grad_partials_tc:
 * float32 [rows, n_grad], its first *rows_used rows overwritten (rows_used: HOST out, may be NULL);
 * NULL = evaluation only.  loss_acc as cnf_nll_train_step.                                        */
int cnf_nll_train_step_tc(const cnf_flow_desc* desc, const void* packed_tc, const int32_t* tables,
                          const float* x, const int64_t* y, int64_t N, float eps, float gamma,
                          float inv_n_total, float* grad_partials_tc, double* loss_acc,
                          void* workspace, int64_t workspace_bytes, int64_t* rows_used, void* stream);
/* rows_used: what cnf_nll_train_step_tc reported (rows [0, rows_used) of the partial buffer were written). */
int cnf_grad_reduce_tc(const cnf_flow_desc* desc, const float* grad_partials_tc, int64_t rows_used,
                       const int32_t* gather_tcgrad, float* flat_grad, void* stream);

/* ---- device: metrics ----------------------------------------------------------- */
/* in: [N,K] float32 (is_f64 == 0) or float64 (is_f64 == 1); labels int64 [N].
 * edges: DEVICE float64 [bins+1], edge i = i*(1/bins) as the reference forms it.
 * acc: DEVICE float64 [3*bins+3], ACCUMULATED: per bin count, sum conf, sum acc;
 *      then sum -log(p_y+1e-7), number correct, N.                               */
int cnf_metrics(const void* in, int32_t is_f64, const int64_t* y, int64_t N, int32_t K,
                int32_t bins, int32_t mode, const double* log_priors, const double* edges,
                double* acc, void* stream);
/* Full Calibrator.predict tail on device: probs_out float64 [N,K].              */
int cnf_calibrated_probs(const float* z, int64_t N, int32_t K, const double* log_priors,
                         double* probs_out, void* stream);

/* ---- device: per-dimension constant affine layer (next-row 8f/2) ---------------------------------
 * AffineConstantLayer.forward / .backward (flows/flows.py:53-65): z = x*exp(s)+t, or with
 * inverse != 0 x = (z-t)*exp(-s).  s / t are DEVICE float32 [K] or NULL (= zeros).  TempScaler
 * (flows/utils.py:40-48) is s_k = -log|T|, t = NULL.                                              */
int cnf_affine_const(const float* x, const float* s, const float* t, float* z, int64_t N, int32_t K,
                     int32_t inverse, void* stream);
/* Gradients of the forward direction: g_x = g_z*exp(s) (may be NULL), g_s[k] = sum_n g_z*x*exp(s),
 * g_t[k] = sum_n g_z (either may be NULL; both are overwritten).                                    */
int cnf_affine_const_backward(const float* x, const float* g_z, const float* s, float* g_x, float* g_s,
                              float* g_t, int64_t N, int32_t K, void* stream);

/* ---- device: PlanarLayer / RadialLayer (next-row 8f/4 tail) ----------------------------------------
 * PlanarLayer.forward (flows/flows.py:148-164): h = tanh(x.w + b), z = x + h*u_hat,
 * logdet = log|1 + (1-h^2)(w.u_hat)|.  w, u_hat DEVICE float32 [K], b DEVICE float32 [1]; u_hat (the
 * invertibility-constrained u, :150-153) is formed by the caller.  K <= 512 and 32 rows x 4 warps must
 * fit shared memory (CNF_E_SMEM otherwise).                                                          */
int cnf_planar_forward(const float* x, const float* w, const float* u_hat, const float* b, float* z,
                       float* logdet, int64_t N, int32_t K, void* stream);
/* Gradients for upstream g_z [N,K], g_logdet [N] (may be NULL = zeros): g_x [N,K] (may be NULL),
 * g_w [K], g_uhat [K], g_b [1] (overwritten).                                                       */
int cnf_planar_backward(const float* x, const float* g_z, const float* g_logdet, const float* w,
                        const float* u_hat, const float* b, float* g_x, float* g_w, float* g_uhat,
                        float* g_b, int64_t N, int32_t K, void* stream);
/* RadialLayer.forward (flows/flows.py:180-193): z = x + b_hat*(x-z0)/(a + |x-z0|); its log-det is the
 * constant log(1.0) in the reference and is produced by the caller.  z0 [K], a [1], b_hat [1] DEVICE. */
int cnf_radial_forward(const float* x, const float* z0, const float* a, const float* b_hat, float* z,
                       int64_t N, int32_t K, void* stream);
int cnf_radial_backward(const float* x, const float* g_z, const float* z0, const float* a,
                        const float* b_hat, float* g_x, float* g_z0, float* g_a, float* g_bhat,
                        int64_t N, int32_t K, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CNF_B200_H */
