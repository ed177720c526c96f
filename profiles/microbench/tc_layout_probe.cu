// Layout probe for tcgen05.mma operand / accumulator layouts that the planned tensor-core training
// kernel needs (B200, sm_100a): M=64 accumulator lane mapping and MN-major shared-memory operands.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_layout_probe tc_layout_probe.cu && ./tc_layout_probe
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
struct Probe {
  int M, N, ksteps;            // MMA shape, number of K=16 steps
  int a_major, b_major;        // 0 = K-major, 1 = MN-major
  uint32_t a_lbo, a_sbo, a_kstride;   // descriptor fields and byte advance per k-step
  uint32_t b_lbo, b_sbo, b_kstride;
  uint32_t a_bytes, b_bytes;
};

__global__ void __launch_bounds__(128, 1) probe(Probe p, const uint8_t* a_img, const uint8_t* b_img, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* sa = smem;
  uint8_t* sb = smem + 65536;
  for (uint32_t i = tid; i < p.a_bytes; i += 128) sa[i] = a_img[i];
  for (uint32_t i = tid; i < p.b_bytes; i += 128) sb[i] = b_img[i];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_ptr;
  // clear the accumulator region first (so untouched lanes read as a sentinel)
  {
    const uint32_t tl = tm + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < p.N; ++c) {
      uint32_t v = 0x7fc00000u;   // NaN sentinel
      asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tl + c), "r"(v) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)p.a_major << 15) | ((uint32_t)p.b_major << 16) |
                           ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(p.M >> 4) << 24);
    for (int j = 0; j < p.ksteps; ++j) {
      const uint64_t ad = make_desc(smem_u32(sa) + j * p.a_kstride, p.a_lbo, p.a_sbo);
      const uint64_t bd = make_desc(smem_u32(sb) + j * p.b_kstride, p.b_lbo, p.b_sbo);
      const uint32_t acc = j > 0;
      asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;\n\t}"
                   ::"r"(tm), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  {
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    const uint32_t tl = tm + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < p.N; ++c) {
      uint32_t v;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(tl + c) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      out[tid * p.N + c] = __uint_as_float(v);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256) : "memory");
}

static uint16_t bf(float v) { uint32_t u; memcpy(&u, &v, 4); return (uint16_t)((u + 0x7fff + ((u >> 16) & 1)) >> 16); }
static float fb(uint16_t h) { uint32_t u = (uint32_t)h << 16; float v; memcpy(&v, &u, 4); return v; }

int main() {
  uint8_t *da, *db; float* dout;
  cudaMalloc(&da, 65536); cudaMalloc(&db, 65536); cudaMalloc(&dout, 128 * 256 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072 + 1024);
  std::vector<float> out(128 * 256);
  auto run = [&](const char* name, Probe p, const std::vector<uint8_t>& A, const std::vector<uint8_t>& B,
                 const std::vector<float>& a, const std::vector<float>& b, int K) {
    p.a_bytes = (uint32_t)A.size(); p.b_bytes = (uint32_t)B.size();
    cudaMemcpy(da, A.data(), A.size(), cudaMemcpyHostToDevice); cudaMemcpy(db, B.data(), B.size(), cudaMemcpyHostToDevice);
    cudaMemset(dout, 0, 128 * 256 * 4);
    probe<<<1, 128, 131072 + 1024>>>(p, da, db, dout);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(out.data(), dout, 128 * p.N * 4, cudaMemcpyDeviceToHost);
    // expected D[m][n] = sum_k a[m][k] b[n][k]
    std::vector<float> D(p.M * p.N);
    for (int m = 0; m < p.M; ++m) for (int n = 0; n < p.N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += a[m * K + k] * b[n * K + k]; D[m * p.N + n] = s; }
    // hypothesis: row m -> lane m
    double err_id = 0; int nan_lanes = 0;
    for (int m = 0; m < p.M; ++m) for (int n = 0; n < p.N; ++n) err_id = fmax(err_id, fabs(out[m * p.N + n] - D[m * p.N + n]));
    // find, for each row m, the lane whose data matches
    printf("%-44s err(row m -> lane m) %.3g (%s)\n", name, err_id, cudaGetErrorString(e));
    if (err_id > 1e-2 || err_id != err_id) {
      printf("   row->lane map:");
      for (int m = 0; m < p.M; m += (p.M >= 64 ? 8 : 1)) {
        int best = -1;
        for (int l = 0; l < 128; ++l) { double er = 0; for (int n = 0; n < p.N; ++n) er = fmax(er, fabs(out[l * p.N + n] - D[m * p.N + n])); if (er < 1e-2) { best = l; break; } }
        printf(" %d->%d", m, best);
      }
      for (int l = 0; l < 128; ++l) if (out[l * p.N] != out[l * p.N]) ++nan_lanes;
      printf("   (untouched lanes: %d)\n", nan_lanes);
    }
  };
  auto rnd = [](int i) { return (float)(((i * 2654435761u) >> 20) % 17) / 8.0f - 1.0f; };
  // ---- case 1: M=128 K-major baseline, N=32, K=32 (2 k-steps) --------------------------------
  for (int M : {128, 64}) {
    const int N = 32, K = 32;
    std::vector<float> a(M * K), b(N * K);
    for (int i = 0; i < M * K; ++i) a[i] = fb(bf(rnd(i + 7)));
    for (int i = 0; i < N * K; ++i) b[i] = fb(bf(rnd(i * 3 + 1)));
    // K-major images: element (r,k): (k/8)*LBO + (r/8)*SBO + (r%8)*16 + (k%8)*2 ; LBO = rows*16, SBO = 128
    std::vector<uint8_t> A(M * K * 2), B(N * K * 2);
    auto put = [](std::vector<uint8_t>& img, size_t byte, float v) { uint16_t h = bf(v); memcpy(&img[byte], &h, 2); };
    const uint32_t a_lbo = M * 16, b_lbo = N * 16;
    for (int r = 0; r < M; ++r) for (int k = 0; k < K; ++k) put(A, (k / 8) * a_lbo + (r / 8) * 128 + (r % 8) * 16 + (k % 8) * 2, a[r * K + k]);
    for (int r = 0; r < N; ++r) for (int k = 0; k < K; ++k) put(B, (k / 8) * b_lbo + (r / 8) * 128 + (r % 8) * 16 + (k % 8) * 2, b[r * K + k]);
    Probe p{M, N, 2, 0, 0, a_lbo, 128, 2 * a_lbo, b_lbo, 128, 2 * b_lbo, 0, 0};
    char nm[64]; snprintf(nm, 64, "K-major A,B  M=%d N=32 K=32", M);
    run(nm, p, A, B, a, b, K);
  }
  // ---- case 2: MN-major operands: image element (r,k): (r/8)*S_mn + (k/8)*S_k + (k%8)*16 + (r%8)*2 -----
  for (int M : {128, 64}) for (int variant = 0; variant < 2; ++variant) for (int which = 0; which < 3; ++which) {
    const int N = 32, K = 32;
    std::vector<float> a(M * K), b(N * K);
    for (int i = 0; i < M * K; ++i) a[i] = fb(bf(rnd(i + 11)));
    for (int i = 0; i < N * K; ++i) b[i] = fb(bf(rnd(i * 5 + 2)));
    std::vector<uint8_t> A(M * K * 2), B(N * K * 2);
    auto put = [](std::vector<uint8_t>& img, size_t byte, float v) { uint16_t h = bf(v); memcpy(&img[byte], &h, 2); };
    const bool a_mn = (which == 0 || which == 2), b_mn = (which == 1 || which == 2);
    // MN-major: k-blocks (8 k-rows = 128 B) adjacent: S_k = 128, S_mn = (K/8)*128
    const uint32_t a_sk = 128, a_smn = (K / 8) * 128, b_sk = 128, b_smn = (K / 8) * 128;
    const uint32_t a_lbo_k = M * 16, b_lbo_k = N * 16;   // K-major fallbacks
    for (int r = 0; r < M; ++r) for (int k = 0; k < K; ++k)
      put(A, a_mn ? (r / 8) * a_smn + (k / 8) * a_sk + (k % 8) * 16 + (r % 8) * 2
                  : (k / 8) * a_lbo_k + (r / 8) * 128 + (r % 8) * 16 + (k % 8) * 2, a[r * K + k]);
    for (int r = 0; r < N; ++r) for (int k = 0; k < K; ++k)
      put(B, b_mn ? (r / 8) * b_smn + (k / 8) * b_sk + (k % 8) * 16 + (r % 8) * 2
                  : (k / 8) * b_lbo_k + (r / 8) * 128 + (r % 8) * 16 + (k % 8) * 2, b[r * K + k]);
    Probe p{};
    p.M = M; p.N = N; p.ksteps = 2; p.a_major = a_mn; p.b_major = b_mn;
    // variant 0: descriptor LBO = K-direction stride, SBO = MN-direction stride; variant 1: swapped
    if (a_mn) { p.a_lbo = variant ? a_smn : a_sk; p.a_sbo = variant ? a_sk : a_smn; p.a_kstride = 2 * a_sk; }
    else      { p.a_lbo = a_lbo_k; p.a_sbo = 128; p.a_kstride = 2 * a_lbo_k; }
    if (b_mn) { p.b_lbo = variant ? b_smn : b_sk; p.b_sbo = variant ? b_sk : b_smn; p.b_kstride = 2 * b_sk; }
    else      { p.b_lbo = b_lbo_k; p.b_sbo = 128; p.b_kstride = 2 * b_lbo_k; }
    char nm[96]; snprintf(nm, 96, "MN-major %s  M=%d  desc variant %d", which == 0 ? "A  " : which == 1 ? "B  " : "A+B", M, variant);
    run(nm, p, A, B, a, b, K);
  }
  return 0;
}
