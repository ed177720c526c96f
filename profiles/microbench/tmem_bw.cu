// TMEM read/write bandwidth probe (B200, sm_100a): W warps of one CTA loop over tcgen05.ld / tcgen05.st
// on their own lane quarter; clock64 around the loop of the slowest warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu && ./tmem_bw
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD32(r, addr) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(addr) : "memory")
#define LD16X256(r, addr) asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(addr) : "memory")
#define ST32(r, addr) asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" \
      ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), \
        "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory")

#define ST16(r, addr) asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" \
      ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory")

// mode 5: 4 x (ld x32 + st x16) then wait both (the EPI1 pattern) ; 6: 4 x st x16 then wait
// mode 0: ld 32x32b.x32, wait after every load ; 1: two loads in flight, then wait ; 2: four loads then wait
// mode 3: st 32x32b.x32 + wait::st each ; 4: ld 16x256b.x8 (32 regs) wait each ; 5: ld.x32 with 8 loads in flight
__global__ void __launch_bounds__(512, 1) bw(int mode, int iters, long long* out, uint32_t* sink) {
  __shared__ uint32_t tmem_ptr;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_ptr + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t a[32], b[32], c[32], d[32];
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) { a[i] = i; b[i] = i; c[i] = i; d[i] = i; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t col = (uint32_t)((it * 128) & 511);
    if (mode == 0) {
      LD32(a, tm + col); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); acc += a[0] + a[31];
      LD32(b, tm + ((col + 32) & 511)); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); acc += b[0] + b[31];
      LD32(c, tm + ((col + 64) & 511)); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); acc += c[0] + c[31];
      LD32(d, tm + ((col + 96) & 511)); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); acc += d[0] + d[31];
    } else if (mode == 1) {
      LD32(a, tm + col); LD32(b, tm + ((col + 32) & 511));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); acc += a[0] + b[31];
      LD32(c, tm + ((col + 64) & 511)); LD32(d, tm + ((col + 96) & 511));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); acc += c[0] + d[31];
    } else if (mode == 2) {
      LD32(a, tm + col); LD32(b, tm + ((col + 32) & 511)); LD32(c, tm + ((col + 64) & 511)); LD32(d, tm + ((col + 96) & 511));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); acc += a[0] + b[31] + c[1] + d[2];
    } else if (mode == 3) {
      ST32(a, tm + col); ST32(b, tm + ((col + 32) & 511)); ST32(c, tm + ((col + 64) & 511)); ST32(d, tm + ((col + 96) & 511));
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    } else if (mode == 5) {
      LD32(a, tm + col); ST16(c, tm + ((col + 256) & 511)); LD32(b, tm + ((col + 32) & 511)); ST16(d, tm + ((col + 272) & 511));
      LD32(a, tm + ((col + 64) & 511)); ST16(c, tm + ((col + 288) & 511)); LD32(b, tm + ((col + 96) & 511)); ST16(d, tm + ((col + 304) & 511));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      acc += a[0] + b[31];
    } else if (mode == 6) {
      ST16(a, tm + col); ST16(b, tm + ((col + 16) & 511)); ST16(c, tm + ((col + 32) & 511)); ST16(d, tm + ((col + 48) & 511));
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    } else if (mode == 4) {
      // 16x256b: one instruction covers 16 lanes x (8 x 8 columns); a warp reads its 32 lanes with two of them
      LD16X256(a, tm + col); LD16X256(b, tm + col + (16u << 16)); LD16X256(c, tm + ((col + 64) & 511)); LD16X256(d, tm + ((col + 64) & 511) + (16u << 16));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); acc += a[0] + b[31] + c[1] + d[2];
    }
  }
  const long long t1 = clock64();
  if ((tid & 31) == 0) out[warp] = t1 - t0;
  sink[tid] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_ptr), "r"(512) : "memory");
}

int main() {
  long long* out; uint32_t* sink;
  cudaMalloc(&out, 16 * 8); cudaMalloc(&sink, 512 * 4);
  const int iters = 2000;
  const char* names[] = {"ld x32, wait each", "ld x32, 2 in flight", "ld x32, 4 in flight", "st x32, 4 then wait", "ld 16x256b.x8, 4 in flight", "4x(ld x32 + st x16)", "st x16 x4 (2 KB ops)"};
  for (int mode = 0; mode < 7; ++mode)
    for (int warps : {1, 4, 8, 16}) {
      bw<<<1, warps * 32, 0>>>(mode, iters, out, sink);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[16];
      cudaMemcpy(h, out, warps * 8, cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
      const double bytes = (double)iters * 4 * 32 * 32 * 4 * warps;    // 4 ops of 4 KB per warp and iteration
      printf("%-28s warps %2d  cycles/op/warp %7.1f   CTA bytes/clk %7.1f   %s\n", names[mode], warps,
             (double)mx / (iters * 4), bytes / mx, cudaGetErrorString(e));
    }
  return 0;
}
