// Micro-benchmark: issue rate / duration of tcgen05.mma (kind::f16, M = 128) as a function of N, with A in shared
// memory (SS) or tensor memory (TS).  nvcc -gencode arch=compute_100a,code=sm_100a -I ../../foundationpose_cpp_b200/csrc
#include <cstdio>
#include "ptx.cuh"
using namespace fpb;
template <int N, bool TS>
__global__ void k(long long* out, int reps) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + 65536, slot = base + 65536 + 8;
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_barrier_init(); }
  if (threadIdx.x < 32) { ptx::tmem_alloc(slot, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  const uint32_t tm = *reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - ptx::smem_u32(smem_raw)));
  if (threadIdx.x < 32) {
    constexpr uint32_t idesc = ptx::make_idesc_f16_major(128, N, 0, 0);
    const uint64_t ad = ptx::make_sw128_kmajor_desc(base), bd = ptx::make_sw128_kmajor_desc(base + 16384);
    long long t0 = clock64(), t1 = 0;
    if (ptx::elect_one()) {
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (TS) ptx::umma_f16_ts(tm + 256, tm + 8 * k, bd + 2 * k, idesc, 1);
          else ptx::umma_f16(tm + 256, ad + 2 * k, bd + 2 * k, idesc, 1);
        }
      }
      t1 = clock64();
      ptx::umma_commit(bar);
    }
    __syncwarp();
    ptx::mbar_wait(bar, 0);
    long long t2 = clock64();
    if (t1) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  ptx::tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc(tm, 512); }
}
template <int N, bool TS>
void run(long long* d) {
  const int reps = 250;
  cudaFuncSetAttribute(k<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  k<N, TS><<<1, 64, 70000>>>(d, reps);
  k<N, TS><<<1, 64, 70000>>>(d, reps);
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("N=%3d %s: issue %.1f clk/MMA, complete %.1f clk/MMA (floor %d) %s\n", N, TS ? "TS" : "SS", h[0] / (4.0 * reps),
         h[1] / (4.0 * reps), N / 2, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  long long* d; cudaMalloc(&d, 16);
  run<16, true>(d); run<64, true>(d); run<80, true>(d); run<128, true>(d); run<208, true>(d); run<256, true>(d);
  run<16, false>(d); run<64, false>(d); run<80, false>(d); run<128, false>(d); run<256, false>(d);
  return 0;
}
