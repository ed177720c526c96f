// Micro-benchmark of tcgen05 building blocks at the shapes the flow kernel uses (B200, sm_100a).
// One CTA of 128 threads per launch; thread 0 issues MMAs, clock64 brackets issue -> mbarrier wait.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_micro tc_micro.cu && ./tc_micro
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c) : "memory"); }
__device__ __forceinline__ uint32_t mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { while (!mbar_try(bar, parity)) {} }
__device__ __forceinline__ void tc_commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__host__ __device__ inline uint32_t make_idesc(int N, int M = 128) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}

// mode: 0 = 1x SS N=256 ; 1 = nk x TS N=16 one accumulator ; 2 = nk x TS N=16, 4 accumulators ;
//       3 = nk x SS N=16 one accumulator ; 4 = nk x TS N=32 ; 5 = 1x SS N=256 followed by nk x TS N=16 (back to back)
//       6 = nk x SS N=256 K=16 accumulate (plain GEMM pacing)
//       7.. = nk x SS accumulate with the B image rotating over four 8 KB blocks (round 2: is the B port as
//       narrow as the A port?): 7 = M64 N256, 8 = M128 N256, 9 = M64 N128, 10 = M64 N64, 11 = M64 N16
__global__ void __launch_bounds__(128, 1) micro(int mode, int nk, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_ptr;
  const uint32_t sb = smem_u32(smem);
  long long best = 1ll << 60, sum = 0;
  uint32_t parity = 0;
  if (tid == 0) {
    for (int r = 0; r < reps; ++r) {
      const long long t0 = clock64();
      if (mode == 0) {
        mma_ss(tm, make_desc(sb, 128, 256), make_desc(sb + 8192, 128, 256), make_idesc(256), 0);
      } else if (mode == 1) {
        for (int j = 0; j < nk; ++j) mma_ts(tm + 128, tm + j * 8, make_desc(sb + j * 512, 256, 128), make_idesc(16), j > 0);
      } else if (mode == 2) {
        for (int j = 0; j < nk; ++j) mma_ts(tm + 128 + (j & 3) * 16, tm + j * 8, make_desc(sb + j * 512, 256, 128), make_idesc(16), j > 3);
      } else if (mode == 3) {
        for (int j = 0; j < nk; ++j) mma_ss(tm + 128, make_desc(sb + 16384 + j * 4096, 128, 256), make_desc(sb + j * 512, 256, 128), make_idesc(16), j > 0);
      } else if (mode == 4) {
        for (int j = 0; j < nk; ++j) mma_ts(tm + 128, tm + j * 8, make_desc(sb + j * 1024, 512, 128), make_idesc(32), j > 0);
      } else if (mode == 5) {
        mma_ss(tm + 256, make_desc(sb, 128, 256), make_desc(sb + 8192, 128, 256), make_idesc(256), 0);
        for (int j = 0; j < nk; ++j) mma_ts(tm + 128, tm + j * 8, make_desc(sb + j * 512, 256, 128), make_idesc(16), j > 0);
      } else if (mode == 6) {
        for (int j = 0; j < nk; ++j) mma_ss(tm, make_desc(sb, 128, 256), make_desc(sb + 8192, 128, 256), make_idesc(256), j > 0);
      } else if (mode >= 7 && mode <= 11) {
        const int M = mode == 8 ? 128 : 64;
        const int N = mode == 7 || mode == 8 ? 256 : mode == 9 ? 128 : mode == 10 ? 64 : 16;
        const uint32_t id = make_idesc(N, M);
        for (int j = 0; j < nk; ++j) mma_ss(tm, make_desc(sb + (j & 1) * 4096, 128, 256), make_desc(sb + 8192 + (j & 3) * 8192, 128, 256), id, j > 0);
      }
      else if (mode == 12) {          // C4 GEMM1: nk x (4 k-steps SS M128 N128)
        for (int j = 0; j < nk; ++j)
          for (int k = 0; k < 4; ++k) mma_ss(tm, make_desc(sb + k * 4096, 2048, 128), make_desc(sb + 16384 + (j & 1) * 32768 + k * 4096, 2048, 128), make_idesc(128), k > 0);
      } else if (mode == 13) {        // C4 GEMM2: nk x (8 k-steps TS M128 N64)
        for (int j = 0; j < nk; ++j)
          for (int k = 0; k < 8; ++k) mma_ts(tm + 128, tm + k * 8, make_desc(sb + 32768 + (j & 1) * 32768 + k * 2048, 1024, 128), make_idesc(64), k > 0);
      } else if (mode == 14) {        // C4 phase: GEMM1 then GEMM2, nk phases back to back
        for (int j = 0; j < nk; ++j) {
          for (int k = 0; k < 4; ++k) mma_ss(tm + (j & 1) * 256, make_desc(sb + k * 4096, 2048, 128), make_desc(sb + 16384 + (j & 1) * 32768 + k * 4096, 2048, 128), make_idesc(128), k > 0);
          for (int k = 0; k < 8; ++k) mma_ts(tm + 128 + (j & 1) * 256, tm + (j & 1) * 256 + k * 8, make_desc(sb + 32768 + (j & 1) * 32768 + k * 2048, 1024, 128), make_idesc(64), k > 0);
        }
      } else if (mode == 15) {        // N sweep, TS M128 K16: nk MMAs of N = 32 * reps_n (passed through nk high bits)
        const int N = (nk >> 8) * 8, cnt = nk & 255;
        for (int j = 0; j < cnt; ++j) mma_ts(tm + 256, tm + (j & 7) * 8, make_desc(sb + 32768 + (j & 3) * 8192, 128, 256), make_idesc(N), j > 0);
      } else if (mode == 16) {        // N sweep, SS M128 K16
        const int N = (nk >> 8) * 8, cnt = nk & 255;
        for (int j = 0; j < cnt; ++j) mma_ss(tm + 256, make_desc(sb + (j & 3) * 4096, 128, 256), make_desc(sb + 32768 + (j & 3) * 8192, 128, 256), make_idesc(N), j > 0);
      }
      tc_commit(&bar);
      const long long t1 = clock64();
      mbar_wait(&bar, parity);
      parity ^= 1;
      const long long t2 = clock64();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (r > 0) { sum += t2 - t0; if (t2 - t0 < best) best = t2 - t0; }
      if (r == reps - 1) out[2] = t1 - t0;
    }
    out[0] = best; out[1] = sum / (reps - 1);
  }
  __syncthreads();
  // TMEM load/store timing by warp 0..3 (each its own lanes)
  {
    uint32_t ra[32], pk[16];
    const uint32_t tl = tm + ((uint32_t)(warp * 32) << 16);
    long long t0 = clock64();
    for (int r = 0; r < 8; ++r) {
      tmem_ld32(tl + 32 * r, ra);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 16; ++i) pk[i] = ra[2 * i] ^ ra[2 * i + 1];
      tmem_st16(tl + 16 * r, pk);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    long long t1 = clock64();
    if (tid == 0) out[3] = t1 - t0;
    t0 = clock64();
    tmem_ld32(tl, ra);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    t1 = clock64();
    if (tid == 0) out[4] = t1 - t0 + (ra[0] & 1);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 64);
  cudaFuncSetAttribute(micro, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
  struct { int mode, nk; const char* what; } cases[] = {
      {0, 1, "1x SS M128 N256 K16"},          {1, 16, "16x TS N16, one accumulator"}, {1, 2, "2x TS N16, one accumulator"},
      {1, 1, "1x TS N16"},                    {2, 16, "16x TS N16, 4 accumulators"},  {3, 16, "16x SS N16, one accumulator"},
      {4, 8, "8x TS N32"},                    {5, 16, "SS N256 then 16x TS N16"},     {6, 16, "16x SS N256 accumulate"},
      {6, 64, "64x SS N256 accumulate"},      {1, 64, "64x TS N16 one accumulator"},
      {7, 64, "64x SS M64 N256 (rotating B)"}, {8, 64, "64x SS M128 N256 (rotating B)"}, {9, 64, "64x SS M64 N128"},
      {10, 64, "64x SS M64 N64"},             {11, 64, "64x SS M64 N16"},             {7, 16, "16x SS M64 N256"},
      {12, 16, "16x C4 GEMM1 (4x SS N128)"},  {13, 16, "16x C4 GEMM2 (8x TS N64)"},   {14, 16, "16x C4 phase (GEMM1+GEMM2)"},
      {15, (2 << 8) | 64, "64x TS N16"},      {15, (4 << 8) | 64, "64x TS N32"},      {15, (8 << 8) | 64, "64x TS N64"},
      {15, (12 << 8) | 64, "64x TS N96"},     {15, (16 << 8) | 64, "64x TS N128"},    {15, (32 << 8) | 64, "64x TS N256"},
      {16, (2 << 8) | 64, "64x SS N16"},      {16, (8 << 8) | 64, "64x SS N64"},      {16, (16 << 8) | 64, "64x SS N128"},
      {16, (24 << 8) | 64, "64x SS N192"},    {16, (32 << 8) | 64, "64x SS N256"}};
  for (auto& c : cases) {
    for (int i = 0; i < 8; ++i) out[i] = 0;
    micro<<<1, 128, 131072>>>(c.mode, c.nk, 50, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%-34s issue->done min %6lld avg %6lld cyc | issue %5lld | epi1-like 8x(ld32+st16) %5lld | ld32 lat %4lld  (%s)\n", c.what,
           out[0], out[1], out[2], out[3], out[4], cudaGetErrorString(e));
  }
  return 0;
}
