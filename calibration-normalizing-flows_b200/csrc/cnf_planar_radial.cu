// PlanarLayer and RadialLayer (flows/flows.py:129-193) as streaming kernels (SURVEY.md 8f rank 4 tail).
// Both are one dot product / one norm per sample plus an elementwise update: 8K+4 bytes per sample,
// HBM-bound.  A warp owns tiles of 32 rows: the lanes copy the tile's 32*K contiguous floats into the
// warp's shared-memory slice with coalesced loads, lane r then processes row r, and the tile goes back
// out as one coalesced run.  The backward kernels also form the parameter gradients: a second pass over
// the tile with one lane per column accumulates sum_n coef_n * tile[n][k] in registers (K <= 512),
// flushed with one atomicAdd per column and warp at the end.
//
// Reference arithmetic (forward direction only -- neither layer is invertible, flows/flows.py:146,178):
//   planar   h = tanh(x.w + b); z = x + h*u_hat; log_det = log|1 + (1-h^2) * (w.u_hat)|      :148-164
//            (u_hat = u + (m - w.u) w/|w| is formed by the caller with torch ops: K-vector math)
//   radial   h = 1/(a + |x - z0|); z = x + b_hat*h*(x - z0); log_det = log(1.0) (a constant)  :180-193
#include <cuda_runtime.h>

#include "cnf_common.h"
#include "cnf_tc_ptx.cuh"   // mbarrier + cp.async.bulk wrappers

namespace {

constexpr int PR_WARPS = 4;
constexpr int PR_MAXC = 16;      // columns per lane in the parameter-gradient pass: K <= 512

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void tile_load(float* dst, const float* __restrict__ src, int nel, int lane) {
  for (int e = lane; e < nel; e += 32) dst[e] = __ldcs(src + e);
}
__device__ __forceinline__ void tile_store(float* __restrict__ dst, const float* src, int nel, int lane) {
  for (int e = lane; e < nel; e += 32) __stcs(dst + e, src[e]);
}

// Full tiles of the forward kernels move by TMA bulk copies (global -> shared on an mbarrier, shared ->
// global as a bulk group): no per-element copy loop, which was half of the instructions of these kernels.
// bulk != 0 requires 16-byte aligned x / z (a full tile is 128*K bytes).
__device__ __forceinline__ void bulk_store_s2g(void* dst_global, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_global), "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// Common tile pipeline of the two forward kernels: row_fn(row_ptr, n) transforms one row in place.
template <typename RowFn>
__device__ __forceinline__ void forward_tiles(const float* __restrict__ x, float* __restrict__ z, int64_t N, int K,
                                              float* tile, uint64_t* bar, int bulk, RowFn row_fn) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t ntiles = (N + 31) / 32;
  const uint32_t tile_bytes = (uint32_t)(32 * K * sizeof(float));
  uint32_t parity = 0;
  for (int64_t t = (int64_t)blockIdx.x * PR_WARPS + warp; t < ntiles; t += (int64_t)gridDim.x * PR_WARPS) {
    const int64_t n0 = t * 32;
    const int rows = (int)((N - n0) < 32 ? (N - n0) : 32);
    const bool full = bulk && rows == 32;
    if (full) {
      if (lane == 0) {
        bulk_store_wait_read();                       // the previous tile's store has finished reading the buffer
        mbar_expect_tx(bar, tile_bytes);
        bulk_copy_g2s(tile, x + n0 * K, tile_bytes, bar);
      }
      mbar_wait(bar, parity);
      parity ^= 1u;
    } else {
      if (lane == 0) bulk_store_wait_read();
      __syncwarp();
      tile_load(tile, x + n0 * K, rows * K, lane);
      __syncwarp();
    }
    if (lane < rows) row_fn(tile + lane * K, n0 + lane);
    if (full) {
      fence_async_smem();                             // generic-proxy writes -> visible to the bulk store
      __syncwarp();
      if (lane == 0) bulk_store_s2g(z + n0 * K, tile, tile_bytes);
    } else {
      __syncwarp();
      tile_store(z + n0 * K, tile, rows * K, lane);
    }
    __syncwarp();
  }
  if (lane == 0) bulk_store_wait_read();
}

__global__ void __launch_bounds__(PR_WARPS * 32)
planar_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ uhat,
                  const float* __restrict__ b, float* __restrict__ z, float* __restrict__ logdet, int64_t N, int K,
                  int bulk) {
  extern __shared__ __align__(128) float sm[];
  __shared__ uint64_t bars[PR_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* tile = sm + (size_t)warp * 32 * K;                 // tiles first: 128*K bytes each, 16-byte aligned
  float* sw = sm + (size_t)PR_WARPS * 32 * K;
  float* su = sw + K;
  for (int k = threadIdx.x; k < K; k += blockDim.x) { sw[k] = w[k]; su[k] = uhat[k]; }
  if (lane == 0) mbar_init(&bars[warp], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  float wu = 0.f;
  for (int k = lane; k < K; k += 32) wu = fmaf(sw[k], su[k], wu);
  wu = wsum(wu);
  const float bias = b[0];
  forward_tiles(x, z, N, K, tile, &bars[warp], bulk, [&](float* row, int64_t n) {
    float a = bias;
    for (int k = 0; k < K; ++k) a = fmaf(row[k], sw[k], a);
    const float h = tanhf(a);
    for (int k = 0; k < K; ++k) row[k] = fmaf(h, su[k], row[k]);
    logdet[n] = logf(fabsf(1.f + (1.f - h * h) * wu));
  });
}

// g_x = g_z + g_a*w ; g_w = sum_n g_a x_n + (sum_n c_n) u_hat ; g_uhat = sum_n h_n g_z_n + (sum_n c_n) w ;
// g_b = sum_n g_a ; with D = 1 + (1-h^2) wu, g_a = (g_z.u_hat)(1-h^2) - 2 h (1-h^2) g_ld wu / D, c = g_ld (1-h^2)/D
__global__ void __launch_bounds__(PR_WARPS * 32)
planar_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gz, const float* __restrict__ gld,
                  const float* __restrict__ w, const float* __restrict__ uhat, const float* __restrict__ b,
                  float* __restrict__ gx, float* __restrict__ gw, float* __restrict__ gu, float* __restrict__ gb,
                  int64_t N, int K) {
  extern __shared__ float sm[];
  float* sw = sm;
  float* su = sm + K;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* tx = sm + 2 * K + (size_t)warp * (64 * K + 64);
  float* tg = tx + 32 * K;
  float* s_ga = tg + 32 * K;
  float* s_h = s_ga + 32;
  for (int k = threadIdx.x; k < K; k += blockDim.x) { sw[k] = w[k]; su[k] = uhat[k]; }
  __syncthreads();
  float wu = 0.f;
  for (int k = lane; k < K; k += 32) wu = fmaf(sw[k], su[k], wu);
  wu = wsum(wu);
  const float bias = b[0];
  float aw[PR_MAXC], au[PR_MAXC];
#pragma unroll
  for (int i = 0; i < PR_MAXC; ++i) { aw[i] = 0.f; au[i] = 0.f; }
  float acc_gb = 0.f, acc_c = 0.f;
  const int64_t ntiles = (N + 31) / 32;
  for (int64_t t = (int64_t)blockIdx.x * PR_WARPS + warp; t < ntiles; t += (int64_t)gridDim.x * PR_WARPS) {
    const int64_t n0 = t * 32;
    const int rows = (int)((N - n0) < 32 ? (N - n0) : 32);
    tile_load(tx, x + n0 * K, rows * K, lane);
    tile_load(tg, gz + n0 * K, rows * K, lane);
    __syncwarp();
    float ga = 0.f, h = 0.f;
    if (lane < rows) {
      const float* rx = tx + lane * K;
      const float* rg = tg + lane * K;
      float a = bias, gh = 0.f;
      for (int k = 0; k < K; ++k) { a = fmaf(rx[k], sw[k], a); gh = fmaf(rg[k], su[k], gh); }
      h = tanhf(a);
      const float hp = 1.f - h * h;
      const float D = 1.f + hp * wu;
      const float gl = gld ? gld[n0 + lane] : 0.f;
      ga = gh * hp - 2.f * h * hp * (gl * wu / D);
      acc_c += gl * hp / D;
      acc_gb += ga;
    }
    s_ga[lane] = ga;
    s_h[lane] = h;
    __syncwarp();
#pragma unroll
    for (int i = 0; i < PR_MAXC; ++i) {
      const int k = lane + 32 * i;
      if (k < K) {
        float a1 = aw[i], a2 = au[i];
        for (int r = 0; r < rows; ++r) {
          a1 = fmaf(s_ga[r], tx[r * K + k], a1);
          a2 = fmaf(s_h[r], tg[r * K + k], a2);
        }
        aw[i] = a1; au[i] = a2;
        if (gx) {
          const float wk = sw[k];
          for (int r = 0; r < rows; ++r) tg[r * K + k] = fmaf(s_ga[r], wk, tg[r * K + k]);
        }
      }
    }
    __syncwarp();
    if (gx) tile_store(gx + n0 * K, tg, rows * K, lane);
    __syncwarp();
  }
  acc_c = wsum(acc_c);
  acc_gb = wsum(acc_gb);
#pragma unroll
  for (int i = 0; i < PR_MAXC; ++i) {
    const int k = lane + 32 * i;
    if (k < K) {
      atomicAdd(gw + k, aw[i] + acc_c * su[k]);
      atomicAdd(gu + k, au[i] + acc_c * sw[k]);
    }
  }
  if (lane == 0) atomicAdd(gb, acc_gb);
}

__global__ void __launch_bounds__(PR_WARPS * 32)
radial_fwd_kernel(const float* __restrict__ x, const float* __restrict__ z0, const float* __restrict__ a_p,
                  const float* __restrict__ bhat_p, float* __restrict__ z, int64_t N, int K, int bulk) {
  extern __shared__ __align__(128) float sm[];
  __shared__ uint64_t bars[PR_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* tile = sm + (size_t)warp * 32 * K;
  float* s0 = sm + (size_t)PR_WARPS * 32 * K;
  for (int k = threadIdx.x; k < K; k += blockDim.x) s0[k] = z0[k];
  if (lane == 0) mbar_init(&bars[warp], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const float a = a_p[0], bhat = bhat_p[0];
  forward_tiles(x, z, N, K, tile, &bars[warp], bulk, [&](float* row, int64_t) {
    float r2 = 0.f;
    for (int k = 0; k < K; ++k) { const float d = row[k] - s0[k]; r2 = fmaf(d, d, r2); }
    const float c = bhat / (a + sqrtf(r2));
    for (int k = 0; k < K; ++k) row[k] = fmaf(c, row[k] - s0[k], row[k]);
  });
}

// d = x - z0, r = |d|, h = 1/(a+r), s = g_z.d :
//   g_x = g_z (1 + b_hat h) - b_hat h^2 s d / r ;  g_z0 = -(g_x - g_z) ;  g_a = -b_hat sum_n h^2 s ;  g_bhat = sum_n h s
__global__ void __launch_bounds__(PR_WARPS * 32)
radial_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gz, const float* __restrict__ z0,
                  const float* __restrict__ a_p, const float* __restrict__ bhat_p, float* __restrict__ gx,
                  float* __restrict__ gz0, float* __restrict__ g_a, float* __restrict__ g_bhat, int64_t N, int K) {
  extern __shared__ float sm[];
  float* s0 = sm;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* tx = sm + K + (size_t)warp * (64 * K + 64);
  float* tg = tx + 32 * K;
  float* s_c1 = tg + 32 * K;      // b_hat * h per row
  float* s_c2 = s_c1 + 32;        // b_hat * h^2 * s / r per row
  for (int k = threadIdx.x; k < K; k += blockDim.x) s0[k] = z0[k];
  __syncthreads();
  const float a = a_p[0], bhat = bhat_p[0];
  float az[PR_MAXC];
#pragma unroll
  for (int i = 0; i < PR_MAXC; ++i) az[i] = 0.f;
  float acc_a = 0.f, acc_b = 0.f;
  const int64_t ntiles = (N + 31) / 32;
  for (int64_t t = (int64_t)blockIdx.x * PR_WARPS + warp; t < ntiles; t += (int64_t)gridDim.x * PR_WARPS) {
    const int64_t n0 = t * 32;
    const int rows = (int)((N - n0) < 32 ? (N - n0) : 32);
    tile_load(tx, x + n0 * K, rows * K, lane);
    tile_load(tg, gz + n0 * K, rows * K, lane);
    __syncwarp();
    float c1 = 0.f, c2 = 0.f;
    if (lane < rows) {
      const float* rx = tx + lane * K;
      const float* rg = tg + lane * K;
      float r2 = 0.f, s = 0.f;
      for (int k = 0; k < K; ++k) { const float d = rx[k] - s0[k]; r2 = fmaf(d, d, r2); s = fmaf(rg[k], d, s); }
      const float r = sqrtf(r2);
      const float h = 1.f / (a + r);
      c1 = bhat * h;
      c2 = r > 0.f ? bhat * h * h * s / r : 0.f;
      acc_a -= bhat * h * h * s;
      acc_b += h * s;
    }
    s_c1[lane] = c1;
    s_c2[lane] = c2;
    __syncwarp();
#pragma unroll
    for (int i = 0; i < PR_MAXC; ++i) {
      const int k = lane + 32 * i;
      if (k < K) {
        const float z0k = s0[k];
        float acc = az[i];
        for (int r = 0; r < rows; ++r) {
          const float g = tg[r * K + k];
          const float delta = s_c1[r] * g - s_c2[r] * (tx[r * K + k] - z0k);   // g_x - g_z
          acc -= delta;
          tg[r * K + k] = g + delta;
        }
        az[i] = acc;
      }
    }
    __syncwarp();
    if (gx) tile_store(gx + n0 * K, tg, rows * K, lane);
    __syncwarp();
  }
  acc_a = wsum(acc_a);
  acc_b = wsum(acc_b);
#pragma unroll
  for (int i = 0; i < PR_MAXC; ++i) {
    const int k = lane + 32 * i;
    if (k < K) atomicAdd(gz0 + k, az[i]);
  }
  if (lane == 0) { atomicAdd(g_a, acc_a); atomicAdd(g_bhat, acc_b); }
}

int launch_cfg(int64_t N, int K, size_t smem, int* grid) {
  int dev = 0, sms = 0;
  CNF_CHECK_CUDA(cudaGetDevice(&dev));
  CNF_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  (void)K;
  int per_sm = (int)(200 * 1024 / (smem + 1024));
  per_sm = per_sm < 1 ? 1 : (per_sm > 12 ? 12 : per_sm);
  const int64_t want = ((N + 31) / 32 + PR_WARPS - 1) / PR_WARPS;
  const int64_t cap = (int64_t)sms * per_sm;
  *grid = (int)(want < cap ? (want > 0 ? want : 1) : cap);
  return CNF_OK;
}

template <typename Kern>
int set_smem(Kern k, size_t bytes) {
  return cnf_kernel_smem(k, bytes);
}

}  // namespace

extern "C" int cnf_planar_forward(const float* x, const float* w, const float* u_hat, const float* b, float* z,
                                  float* logdet, int64_t N, int32_t K, void* stream) {
  if (N < 0 || K < 1 || K > 32 * PR_MAXC) { cnf_set_error("cnf_planar_forward: bad argument (K <= %d)", 32 * PR_MAXC); return CNF_E_ARG; }
  if (N == 0) return CNF_OK;      // an empty batch may come with null data pointers
  if (!x || !w || !u_hat || !b || !z || !logdet) { cnf_set_error("cnf_planar_forward: null pointer"); return CNF_E_ARG; }
  const size_t smem = (size_t)(2 * K + PR_WARPS * 32 * K) * sizeof(float);
  if (smem > 200 * 1024) { cnf_set_error("cnf_planar_forward: K=%d rows do not fit shared memory", K); return CNF_E_SMEM; }
  int grid = 1, rc;
  if ((rc = launch_cfg(N, K, smem, &grid))) return rc;
  if ((rc = set_smem(planar_fwd_kernel, smem))) return rc;
  planar_fwd_kernel<<<grid, PR_WARPS * 32, smem, (cudaStream_t)stream>>>(x, w, u_hat, b, z, logdet, N, K,
                                                                         (((uintptr_t)x | (uintptr_t)z) % 16 == 0) ? 1 : 0);
  CNF_CHECK_CUDA(cudaGetLastError());
  return CNF_OK;
}

extern "C" int cnf_planar_backward(const float* x, const float* g_z, const float* g_logdet, const float* w,
                                   const float* u_hat, const float* b, float* g_x, float* g_w, float* g_uhat, float* g_b,
                                   int64_t N, int32_t K, void* stream) {
  if (!g_w || !g_uhat || !g_b || N < 0 || K < 1 || K > 32 * PR_MAXC) { cnf_set_error("cnf_planar_backward: bad argument (K <= %d)", 32 * PR_MAXC); return CNF_E_ARG; }
  if (N > 0 && (!x || !g_z || !w || !u_hat || !b)) { cnf_set_error("cnf_planar_backward: null pointer"); return CNF_E_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  CNF_CHECK_CUDA(cudaMemsetAsync(g_w, 0, K * sizeof(float), st));
  CNF_CHECK_CUDA(cudaMemsetAsync(g_uhat, 0, K * sizeof(float), st));
  CNF_CHECK_CUDA(cudaMemsetAsync(g_b, 0, sizeof(float), st));
  if (N == 0) return CNF_OK;
  const size_t smem = (size_t)(2 * K + PR_WARPS * (64 * K + 64)) * sizeof(float);
  if (smem > 200 * 1024) { cnf_set_error("cnf_planar_backward: K=%d rows do not fit shared memory", K); return CNF_E_SMEM; }
  int grid = 1, rc;
  if ((rc = launch_cfg(N, K, smem, &grid))) return rc;
  if ((rc = set_smem(planar_bwd_kernel, smem))) return rc;
  planar_bwd_kernel<<<grid, PR_WARPS * 32, smem, st>>>(x, g_z, g_logdet, w, u_hat, b, g_x, g_w, g_uhat, g_b, N, K);
  CNF_CHECK_CUDA(cudaGetLastError());
  return CNF_OK;
}

extern "C" int cnf_radial_forward(const float* x, const float* z0, const float* a, const float* b_hat, float* z,
                                  int64_t N, int32_t K, void* stream) {
  if (N < 0 || K < 1 || K > 32 * PR_MAXC) { cnf_set_error("cnf_radial_forward: bad argument (K <= %d)", 32 * PR_MAXC); return CNF_E_ARG; }
  if (N == 0) return CNF_OK;
  if (!x || !z0 || !a || !b_hat || !z) { cnf_set_error("cnf_radial_forward: null pointer"); return CNF_E_ARG; }
  const size_t smem = (size_t)(K + PR_WARPS * 32 * K) * sizeof(float);
  if (smem > 200 * 1024) { cnf_set_error("cnf_radial_forward: K=%d rows do not fit shared memory", K); return CNF_E_SMEM; }
  int grid = 1, rc;
  if ((rc = launch_cfg(N, K, smem, &grid))) return rc;
  if ((rc = set_smem(radial_fwd_kernel, smem))) return rc;
  radial_fwd_kernel<<<grid, PR_WARPS * 32, smem, (cudaStream_t)stream>>>(x, z0, a, b_hat, z, N, K,
                                                                         (((uintptr_t)x | (uintptr_t)z) % 16 == 0) ? 1 : 0);
  CNF_CHECK_CUDA(cudaGetLastError());
  return CNF_OK;
}

extern "C" int cnf_radial_backward(const float* x, const float* g_z, const float* z0, const float* a, const float* b_hat,
                                   float* g_x, float* g_z0, float* g_a, float* g_bhat, int64_t N, int32_t K, void* stream) {
  if (!g_z0 || !g_a || !g_bhat || N < 0 || K < 1 || K > 32 * PR_MAXC) { cnf_set_error("cnf_radial_backward: bad argument (K <= %d)", 32 * PR_MAXC); return CNF_E_ARG; }
  if (N > 0 && (!x || !g_z || !z0 || !a || !b_hat)) { cnf_set_error("cnf_radial_backward: null pointer"); return CNF_E_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  CNF_CHECK_CUDA(cudaMemsetAsync(g_z0, 0, K * sizeof(float), st));
  CNF_CHECK_CUDA(cudaMemsetAsync(g_a, 0, sizeof(float), st));
  CNF_CHECK_CUDA(cudaMemsetAsync(g_bhat, 0, sizeof(float), st));
  if (N == 0) return CNF_OK;
  const size_t smem = (size_t)(K + PR_WARPS * (64 * K + 64)) * sizeof(float);
  if (smem > 200 * 1024) { cnf_set_error("cnf_radial_backward: K=%d rows do not fit shared memory", K); return CNF_E_SMEM; }
  int grid = 1, rc;
  if ((rc = launch_cfg(N, K, smem, &grid))) return rc;
  if ((rc = set_smem(radial_bwd_kernel, smem))) return rc;
  radial_bwd_kernel<<<grid, PR_WARPS * 32, smem, st>>>(x, g_z, z0, a, b_hat, g_x, g_z0, g_a, g_bhat, N, K);
  CNF_CHECK_CUDA(cudaGetLastError());
  return CNF_OK;
}
