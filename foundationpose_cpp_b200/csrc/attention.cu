// Self-attention core of the refine / score transformers: softmax(Q K^T / sqrt(128)) V for
// 400 tokens, 4 heads x 128 dims, per hypothesis (in the reference this runs inside
// the TensorRT engines, easy_deploy_tool/inference_core/trt_core/src/trt_core.cpp:367).
//
// Flash-style: one CTA = 80 query rows of one (hypothesis, head); 5 warps x 16 rows; the 400 keys
// stream through shared memory in 5 double-buffered tiles of 80 (cp.async); scores never leave
// registers.  The 1.4 % of the network's FLOPs that live here run on mma.sync m16n8k16 fragments;
// the 98 % in convolutions / projections run on tcgen05 (conv_gemm.cu).
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "net_ops.h"

namespace fpb {

namespace {

constexpr int kTokens = 400;
constexpr int kHeadDim = 128;
constexpr int kHeads = 4;
constexpr int kBM = 80;  // query rows per CTA
constexpr int kBN = 80;  // keys per tile
constexpr int kWarps = kBM / 16;
constexpr int kThreadsAttn = kWarps * 32;
constexpr int kTileBytes = kBN * kHeadDim * 2;  // 20480
constexpr int kSmemAttn = 5 * kTileBytes;       // Q + 2 x (K, V)

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// byte offset of (row, 16-byte chunk) in a [rows][128] fp16 tile with an XOR swizzle
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {
  return static_cast<uint32_t>(row * 256 + ((chunk ^ (row & 7)) << 4));
}

// Copies an 80 x 128 fp16 tile (row stride ld elements) into swizzled shared memory.
__device__ __forceinline__ void load_tile(uint32_t dst, const __half* src, int ld, int tid) {
#pragma unroll
  for (int i = 0; i < (kBN * 16) / kThreadsAttn; ++i) {
    const int idx = tid + i * kThreadsAttn;
    const int row = idx >> 4, chunk = idx & 15;
    cp_async16(dst + tile_off(row, chunk), src + static_cast<long long>(row) * ld + chunk * 8);
  }
}

__global__ void __launch_bounds__(kThreadsAttn, 2)
attention_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int qkv_ld, int out_ld) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sq = smem_u32(smem);
  const uint32_t sk = sq + kTileBytes;      // 2 stages
  const uint32_t sv = sk + 2 * kTileBytes;  // 2 stages
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qt = blockIdx.x, head = blockIdx.y, hyp = blockIdx.z;
  const long long row0 = static_cast<long long>(hyp) * kTokens;
  const __half* qbase = qkv + (row0 + qt * kBM) * qkv_ld + head * kHeadDim;
  const __half* kbase = qkv + row0 * qkv_ld + kHeads * kHeadDim + head * kHeadDim;
  const __half* vbase = kbase + kHeads * kHeadDim;

  load_tile(sq, qbase, qkv_ld, tid);
  load_tile(sk, kbase, qkv_ld, tid);
  load_tile(sv, vbase, qkv_ld, tid);
  cp_async_commit();

  float o[16][4];
#pragma unroll
  for (int i = 0; i < 16; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  uint32_t qf[8][4];
  const float scale_log2 = 0.08838834764831845f * 1.4426950408889634f;  // 1/sqrt(128) * log2(e)

  constexpr int kTiles = kTokens / kBN;
  for (int j = 0; j < kTiles; ++j) {
    const int st = j & 1;
    if (j + 1 < kTiles) {
      load_tile(sk + (st ^ 1) * kTileBytes, kbase + static_cast<long long>(j + 1) * kBN * qkv_ld, qkv_ld, tid);
      load_tile(sv + (st ^ 1) * kTileBytes, vbase + static_cast<long long>(j + 1) * kBN * qkv_ld, qkv_ld, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (j == 0) {
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const int row = warp * 16 + (lane & 15);
        const int chunk = ks * 2 + (lane >> 4);
        ldsm_x4(sq + tile_off(row, chunk), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
      }
    }
    // S = Q K^T  (16 x 80 per warp)
    float s[10][4];
#pragma unroll
    for (int nb = 0; nb < 10; ++nb) s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f;
    const uint32_t kt = sk + st * kTileBytes;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
      for (int nb = 0; nb < 10; nb += 2) {
        const int m = lane >> 3;
        const int row = (nb + (m >> 1)) * 8 + (lane & 7);
        const int chunk = ks * 2 + (m & 1);
        uint32_t b0, b1, b2, b3;
        ldsm_x4(kt + tile_off(row, chunk), b0, b1, b2, b3);
        mma16816(s[nb], qf[ks], b0, b1);
        mma16816(s[nb + 1], qf[ks], b2, b3);
      }
    }
    // online softmax; this thread holds rows g (regs 0,1) and g+8 (regs 2,3)
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nb = 0; nb < 10; ++nb) {
      mx[0] = fmaxf(mx[0], fmaxf(s[nb][0], s[nb][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nb][2], s[nb][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float corr[2], msc[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float m_new = fmaxf(m_run[r], mx[r]);
      corr[r] = exp2f((m_run[r] - m_new) * scale_log2);
      m_run[r] = m_new;
      msc[r] = m_new * scale_log2;
    }
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int nb = 0; nb < 10; ++nb) {
      s[nb][0] = exp2f(s[nb][0] * scale_log2 - msc[0]);
      s[nb][1] = exp2f(s[nb][1] * scale_log2 - msc[0]);
      s[nb][2] = exp2f(s[nb][2] * scale_log2 - msc[1]);
      s[nb][3] = exp2f(s[nb][3] * scale_log2 - msc[1]);
      rs[0] += s[nb][0] + s[nb][1];
      rs[1] += s[nb][2] + s[nb][3];
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + rs[r];
#pragma unroll
    for (int nb = 0; nb < 16; ++nb) {
      o[nb][0] *= corr[0];
      o[nb][1] *= corr[0];
      o[nb][2] *= corr[1];
      o[nb][3] *= corr[1];
    }
    // O += P V
    const uint32_t vt = sv + st * kTileBytes;
#pragma unroll
    for (int ks = 0; ks < 5; ++ks) {
      uint32_t pa[4];
      pa[0] = pack_half2(s[2 * ks][0], s[2 * ks][1]);
      pa[1] = pack_half2(s[2 * ks][2], s[2 * ks][3]);
      pa[2] = pack_half2(s[2 * ks + 1][0], s[2 * ks + 1][1]);
      pa[3] = pack_half2(s[2 * ks + 1][2], s[2 * ks + 1][3]);
#pragma unroll
      for (int nb = 0; nb < 16; nb += 2) {
        const int m = lane >> 3;
        const int row = ks * 16 + (m & 1) * 8 + (lane & 7);
        const int chunk = nb + (m >> 1);
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(vt + tile_off(row, chunk), b0, b1, b2, b3);
        mma16816(o[nb], pa, b0, b1);
        mma16816(o[nb + 1], pa, b2, b3);
      }
    }
    __syncthreads();  // everyone is done with stage st before it is refilled
  }

  // finalise: quad-reduce the row sums, normalise, store
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
  const int g = lane >> 2, t = lane & 3;
  __half* obase = out + (row0 + qt * kBM + warp * 16) * out_ld + head * kHeadDim;
#pragma unroll
  for (int nb = 0; nb < 16; ++nb) {
    *reinterpret_cast<__half2*>(obase + static_cast<long long>(g) * out_ld + nb * 8 + 2 * t) =
        __floats2half2_rn(o[nb][0] * inv0, o[nb][1] * inv0);
    *reinterpret_cast<__half2*>(obase + static_cast<long long>(g + 8) * out_ld + nb * 8 + 2 * t) =
        __floats2half2_rn(o[nb][2] * inv1, o[nb][3] * inv1);
  }
}

}  // namespace

const char* attention_init() {
  cudaError_t e =
      cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemAttn);
  return e == cudaSuccess ? nullptr : "cudaFuncSetAttribute(attention smem) failed";
}

void launch_attention(const __half* qkv, __half* out, int n_hyp, int qkv_ld, int out_ld, cudaStream_t s) {
  if (n_hyp <= 0) return;
  dim3 grid(kTokens / kBM, kHeads, n_hyp);
  attention_kernel<<<grid, kThreadsAttn, kSmemAttn, s>>>(qkv, out, qkv_ld, out_ld);
}

}  // namespace fpb
