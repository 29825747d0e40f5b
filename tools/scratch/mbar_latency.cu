// Micro-benchmark: latency from mbarrier.arrive (warp 1) to the waiter's first instruction after the wait (warp 0),
// for mbarrier.try_wait (hardware-suspended) and mbarrier.test_wait (spinning), same SM.  Also tcgen05.commit -> wait.
#include <cstdio>
#include "ptx.cuh"
using namespace fpb;
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__global__ void k(long long* out, int mode, int reps, int delay) {
  __shared__ alignas(8) unsigned long long bars[2];
  __shared__ volatile long long t_arrive;
  const uint32_t b0 = ptx::smem_u32(&bars[0]), b1 = ptx::smem_u32(&bars[1]);
  if (threadIdx.x == 0) { ptx::mbar_init(b0, 1); ptx::mbar_init(b1, 1); ptx::fence_barrier_init(); }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long tot = 0, mx = 0;
  for (int r = 0; r < reps; ++r) {
    const uint32_t par = r & 1;
    if (warp == 1) {
      // wait until the waiter is surely parked, then arrive
      long long t0 = clock64();
      while (clock64() - t0 < delay) {}
      if (lane == 0) { t_arrive = clock64(); ptx::mbar_arrive(b0); }
      __syncwarp();
      // wait for the waiter's acknowledgement before the next round
      while (!ptx::mbar_try_wait(b1, par)) {}
    } else if (warp == 0) {
      if (mode == 0) { while (!ptx::mbar_try_wait(b0, par)) {} }
      else { while (!mbar_test_wait(b0, par)) {} }
      const long long t1 = clock64();
      if (lane == 0) { const long long d = t1 - t_arrive; tot += d; if (d > mx) mx = d; ptx::mbar_arrive(b1); }
      __syncwarp();
    }
  }
  if (threadIdx.x == 0) { out[0] = tot; out[1] = mx; }
}
int main() {
  long long* d; cudaMalloc(&d, 16);
  for (int delay : {200, 2000, 20000})
    for (int mode = 0; mode < 2; ++mode) {
      k<<<1, 64>>>(d, mode, 200, delay);
      long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("%s, waiter parked %5d clk: arrive -> resume avg %.0f clk, max %lld  %s\n", mode ? "test_wait spin" : "try_wait      ", delay,
             h[0] / 200.0, h[1], cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
