// Do two (or three) dependent MMA chains, issued by different warps, overlap on the tensor pipe?  (B200, sm_100a)
// Each chain: R times { issue a block of MMAs (C4 phase: 8x TS N64 + 4x SS N128) ; commit ; wait for the commit
// [; relay: an "epilogue" warp waits for the commit and arrives on a second mbarrier the issuer waits on] }.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_chain tc_chain.cu && ./tc_chain
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
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
__host__ __device__ inline uint32_t make_idesc(int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24); }

// chains: number of issuer warps (1..3); relay: 0 = issuer waits for its own commit, 1 = via an epilogue warp (32 arrivals),
// 2 = via an epilogue warpgroup, all 128 threads arrive, 3 = via a warpgroup, __syncwarp + one arrival per warp;
// one_thread: all chains issued round-robin by warp 0's thread (waits in order)
__global__ void __launch_bounds__(512, 1) chain(int chains, int relay, int one_thread, int R, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t done[4], go[4];
  __shared__ uint32_t tmem_ptr;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 131072 / 4; i += 512) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) { mbar_init(done + i, 1); mbar_init(go + i, relay == 2 ? 128 : relay == 3 ? 4 : 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 3) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_ptr, sb = smem_u32(smem);
  auto block = [&](int c) {       // one C4 phase of chain c: GEMM2 (8x TS N64) then GEMM1 (4x SS N128)
    const uint32_t t = tm + c * 160;
    for (int k = 0; k < 8; ++k) mma_ts(t + 128, t + k * 8, make_desc(sb + 65536 + k * 2048, 1024, 128), make_idesc(64), k > 0);
    for (int k = 0; k < 4; ++k) mma_ss(t, make_desc(sb + k * 4096, 2048, 128), make_desc(sb + 16384 + k * 4096, 2048, 128), make_idesc(128), k > 0);
  };
  const long long t0 = clock64();
  if (one_thread) {
    if (tid == 0) {
      for (int r = 0; r < R; ++r)
        for (int c = 0; c < chains; ++c) {
          if (r > 0) { mbar_wait(relay ? go + c : done + c, (r - 1) & 1); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
          block(c);
          tc_commit(done + c);
        }
      for (int c = 0; c < chains; ++c) mbar_wait(relay ? go + c : done + c, (R - 1) & 1);
    }
  } else if (warp < chains && lane == 0) {
    const int c = warp;
    for (int r = 0; r < R; ++r) {
      block(c);
      tc_commit(done + c);
      mbar_wait(relay ? go + c : done + c, r & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
  }
  if (relay && warp >= 4 && (warp - 4) / 4 < chains && (relay >= 2 || (warp & 3) == 0)) {      // "epilogue" warps of chain (warp-4)/4
    const int c = (warp - 4) / 4;
    for (int r = 0; r < R; ++r) {
      mbar_wait(done + c, r & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      if (relay == 3) { __syncwarp(); if (lane == 0) mbar_arrive(go + c); }
      else mbar_arrive(go + c);
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (tid == 0) out[0] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 3) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 64);
  cudaFuncSetAttribute(chain, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
  const int R = 200;
  for (int one = 0; one < 1; ++one)
    for (int relay = 0; relay < 4; ++relay)
      for (int chains = 1; chains <= 3; ++chains) {
        out[0] = 0;
        chain<<<1, 512, 131072>>>(chains, relay, one, R, out);
        cudaError_t e = cudaDeviceSynchronize();
        printf("%s issuer(s), relay %d, %d chain(s): %6.0f clk per round of %d block(s)  = %5.0f clk per block (tensor time of a block ~ 720)  (%s)\n",
               one ? "one thread as" : "one warp per ", relay, chains, (double)out[0] / R, chains, (double)out[0] / R / chains, cudaGetErrorString(e));
      }
  return 0;
}
