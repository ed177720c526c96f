// Probe: how does tcgen05.mma kind::f16 lay out an F16 accumulator (idesc c_format = 0) in TMEM?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_f16acc_probe tc_f16acc_probe.cu && ./tc_f16acc_probe
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}

// ab_fmt: 0 = f16 operands, 1 = bf16 operands ; second MMA (optional): D2[128 x 16] = A(TMEM, from D region) . B2
__global__ void __launch_bounds__(128, 1) probe(int N, int ab_fmt, int chain, const uint8_t* a_img, const uint8_t* b_img,
                                               const uint8_t* b2_img, uint32_t* out, float* out2) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* sa = smem; uint8_t* sb = smem + 8192; uint8_t* sb2 = smem + 8192 + 16384;
  for (int i = tid; i < 128 * 32; i += 128) sa[i] = a_img[i];
  for (int i = tid; i < N * 32; i += 128) sb[i] = b_img[i];
  for (int i = tid; i < 16 * N * 2; i += 128) sb2[i] = b2_img[i];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_ptr;
  const uint32_t tl = tm + ((uint32_t)(warp * 32) << 16);
  for (int c = 0; c < 256 + 16; ++c) {
    uint32_t v = 0xdeadbeefu;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tl + c), "r"(v) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) {
    // c_format (bits 4..5) = 0 (F16); a_format (7..9), b_format (10..12): 0 = f16, 1 = bf16
    const uint32_t idesc = (0u << 4) | ((uint32_t)ab_fmt << 7) | ((uint32_t)ab_fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t ad = make_desc(smem_u32(sa), 128, 256), bd = make_desc(smem_u32(sb), 128, 256);
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;\n\t}"
                 ::"r"(tm), "l"(ad), "l"(bd), "r"(idesc), "r"(0u) : "memory");
    if (chain) {
      // D2 (fp32, at column 256) = A(TMEM f16, packed pairs assumed at columns j*8..) . B2 ; B2 image as the flow kernel's
      const uint32_t idesc2 = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      for (int j = 0; j < N / 16; ++j) {
        const uint64_t b2d = make_desc(smem_u32(sb2) + j * 512, 256, 128);
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, q;\n\t}"
                     ::"r"(tm + 256), "r"(tm + j * 8), "l"(b2d), "r"(idesc2), "r"((uint32_t)(j > 0)) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  {
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c = 0; c < N; ++c) {
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(tl + c) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    out[tid * N + c] = v;
  }
  for (int c = 0; c < 16; ++c) {
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(tl + 256 + c) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    out2[tid * 16 + c] = __uint_as_float(v);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}

static uint16_t f2h(float v) { __half h = __float2half_rn(v); uint16_t u; memcpy(&u, &h, 2); return u; }
static float h2f(uint16_t u) { __half h; memcpy(&h, &u, 2); return __half2float(h); }
static uint16_t f2b(float v) { uint32_t u; memcpy(&u, &v, 4); return (uint16_t)((u + 0x7fff + ((u >> 16) & 1)) >> 16); }
static float b2f(uint16_t h) { uint32_t u = (uint32_t)h << 16; float v; memcpy(&v, &u, 4); return v; }

int main() {
  const int N = 64, K = 16;
  uint8_t *da, *db, *db2; uint32_t* dout; float* dout2;
  cudaMalloc(&da, 8192); cudaMalloc(&db, 16384); cudaMalloc(&db2, 16384); cudaMalloc(&dout, 128 * 256 * 4); cudaMalloc(&dout2, 128 * 16 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  auto rnd = [](int i) { return (float)(((i * 2654435761u) >> 20) % 17) / 8.0f - 1.0f; };
  for (int ab = 0; ab < 2; ++ab) {
    std::vector<float> a(128 * K), b(N * K), b2(16 * N);
    for (int i = 0; i < 128 * K; ++i) a[i] = ab ? b2f(f2b(rnd(i + 7))) : h2f(f2h(rnd(i + 7)));
    for (int i = 0; i < N * K; ++i) b[i] = ab ? b2f(f2b(rnd(i * 3 + 1))) : h2f(f2h(rnd(i * 3 + 1)));
    for (int i = 0; i < 16 * N; ++i) b2[i] = h2f(f2h(rnd(i * 7 + 3) * 0.25f));
    std::vector<uint8_t> A(8192, 0), B(16384, 0), B2(16384, 0);
    auto put = [&](std::vector<uint8_t>& img, size_t byte, float v, int fmt) { uint16_t h = fmt ? f2b(v) : f2h(v); memcpy(&img[byte], &h, 2); };
    for (int r = 0; r < 128; ++r) for (int k = 0; k < K; ++k) put(A, (r / 8) * 256 + (k / 8) * 128 + (r % 8) * 16 + (k % 8) * 2, a[r * K + k], ab);
    for (int r = 0; r < N; ++r) for (int k = 0; k < K; ++k) put(B, (r / 8) * 256 + (k / 8) * 128 + (r % 8) * 16 + (k % 8) * 2, b[r * K + k], ab);
    // B2 [16 x N] in the flow kernel's per-k-step layout (always f16 here)
    for (int n2 = 0; n2 < 16; ++n2) for (int kk = 0; kk < N; ++kk)
      put(B2, (kk / 16) * 512 + ((kk % 16) / 8) * 256 + (n2 / 8) * 128 + (n2 % 8) * 16 + (kk % 8) * 2, b2[n2 * N + kk], 0);
    cudaMemcpy(da, A.data(), 8192, cudaMemcpyHostToDevice); cudaMemcpy(db, B.data(), 16384, cudaMemcpyHostToDevice); cudaMemcpy(db2, B2.data(), 16384, cudaMemcpyHostToDevice);
    for (int chain = 0; chain < 2; ++chain) {
      probe<<<1, 128, 65536>>>(N, ab, chain, da, db, db2, dout, dout2);
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<uint32_t> out(128 * N); std::vector<float> out2(128 * 16);
      cudaMemcpy(out.data(), dout, 128 * N * 4, cudaMemcpyDeviceToHost); cudaMemcpy(out2.data(), dout2, 128 * 16 * 4, cudaMemcpyDeviceToHost);
      std::vector<float> D(128 * N);
      for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += a[m * K + k] * b[n * K + k]; D[m * N + n] = s; }
      // hypothesis P: packed pairs, column c = (D[m][2c], D[m][2c+1]) ; hypothesis U: unpacked, low half of column c = D[m][c]
      double errP = 0, errU = 0; int untouchedP = 0;
      for (int m = 0; m < 128; ++m) {
        for (int c = 0; c < N / 2; ++c) {
          const uint32_t v = out[m * N + c];
          errP = fmax(errP, fabs(h2f(v & 0xffff) - D[m * N + 2 * c])); errP = fmax(errP, fabs(h2f(v >> 16) - D[m * N + 2 * c + 1]));
        }
        for (int c = N / 2; c < N; ++c) untouchedP += out[m * N + c] == 0xdeadbeefu;
        for (int c = 0; c < N; ++c) errU = fmax(errU, fabs(h2f(out[m * N + c] & 0xffff) - D[m * N + c]));
      }
      printf("%s operands, F16 accumulator, N=%d: %s | packed-pairs err %.3g (upper-half columns untouched: %d of %d) | unpacked err %.3g\n",
             ab ? "bf16" : "f16", N, cudaGetErrorString(e), errP, untouchedP, 128 * N / 2, errU);
      printf("   lane 0 raw: %08x %08x %08x %08x ... expected D[0][0..3] = %.4f %.4f %.4f %.4f\n", out[0], out[1], out[2], out[3], D[0], D[1], D[2], D[3]);
      if (chain) {
        // D2[m][n2] = sum_kk relu?(no) f16(D[m][kk]) * b2[n2][kk]
        double err2 = 0;
        for (int m = 0; m < 128; ++m) for (int n2 = 0; n2 < 16; ++n2) {
          float s = 0; for (int kk = 0; kk < N; ++kk) s += h2f(f2h(D[m * N + kk])) * b2[n2 * N + kk];
          err2 = fmax(err2, fabs(out2[m * 16 + n2] - s));
        }
        printf("   chained GEMM2 with A = the F16 accumulator read in place from TMEM: err %.3g\n", err2);
      }
    }
  }
  return 0;
}
