// Micro-benchmark: latency of tcgen05.ld (32x32b.x32) + wait::ld in one warp while another warp keeps the tensor pipe
// busy with back-to-back tcgen05.mma (M = 128, N = 64 or 256, A in TMEM), vs an idle tensor pipe.
#include <cstdio>
#include "ptx.cuh"
using namespace fpb;
template <int N>
__global__ void k(long long* out, int reps, int busy, int nld_warps, int with_st) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + 65536, slot = base + 65536 + 8;
  __shared__ volatile int stop;
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_barrier_init(); stop = 0; }
  if (threadIdx.x < 32) { ptx::tmem_alloc(slot, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  const uint32_t tm = *reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - ptx::smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    if (busy) {
      constexpr uint32_t idesc = ptx::make_idesc_f16_major(128, N, 0, 0);
      const uint64_t bd = ptx::make_sw128_kmajor_desc(base + 16384);
      if (ptx::elect_one()) {
        out[5] = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::umma_f16_ts(tm + 256, tm + 8 * k, bd + 2 * k, idesc, 1);
        }
        ptx::umma_commit(bar);
      }
      __syncwarp();
      ptx::mbar_wait(bar, 0);
      if (threadIdx.x == 0) out[4] = clock64() - out[5];
    } else {
      long long t0 = clock64();
      while (clock64() - t0 < 200000) {}
    }
    stop = 1;
  } else if (warp >= 4 && warp < 4 + nld_warps) {
    const uint32_t tl = tm + ((uint32_t)((warp & 3) * 32) << 16);
    long long tot = 0, mx = 0; int n = 0; float acc = 0.f;
    while (!stop) {
      uint32_t r[32];
      long long t0 = clock64();
      ptx::tmem_ld_32x32b_x32(tl + 64, r);
      ptx::tmem_ld_wait();
      long long t1 = clock64();
      acc += __uint_as_float(r[0]) + __uint_as_float(r[31]);
      if (with_st) {
        uint32_t w[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) w[e] = r[e];
        ptx::tmem_st_32x32b_x16(tl + 128, w);
        ptx::tmem_st_wait();
      }
      tot += t1 - t0; if (t1 - t0 > mx) mx = t1 - t0; ++n;
    }
    if (threadIdx.x == 128) { out[0] = tot; out[1] = n; out[2] = mx; out[3] = (long long)acc; }
  }
  ptx::tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc(tm, 512); }
}
template <int N>
void run(long long* d, int busy, int nld, int with_st = 0) {
  cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  cudaMemset(d, 0, 64);
  k<N><<<1, 384, 70000>>>(d, 4000 * 64 / N, busy, nld, with_st);
  long long h[6];
  cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
  printf("N=%3d busy=%d ld-warps=%d: tcgen05.ld x32 + wait: avg %.1f clk, max %lld, n=%lld  MMA %.1f clk each (floor %d) st=%d %s\n", N, busy, nld,
         (double)h[0] / (double)(h[1] ? h[1] : 1), h[2], h[1], (double)h[4] / (4.0 * (4000 * 64 / N)), N / 2, with_st, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  long long* d; cudaMalloc(&d, 64);
  run<64>(d, 1, 0); run<64>(d, 1, 1); run<64>(d, 1, 4); run<64>(d, 1, 8); run<64>(d, 1, 8, 1); run<256>(d, 1, 8); run<256>(d, 1, 8, 1); run<16>(d, 1, 8, 1);
  return 0;
}
