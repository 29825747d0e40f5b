// Self-attention core of the refine / score transformers: softmax(Q K^T / sqrt(128)) V for
// 400 tokens, 4 heads x 128 dims, per hypothesis (in the reference this runs inside
// the TensorRT engines, easy_deploy_tool/inference_core/trt_core/src/trt_core.cpp:367).
//
// tcgen05 version.  One persistent CTA per SM walks (hypothesis, head, 128-query tile) work items:
//   warp 0      TMA producer: the Q tile (double-buffered) and the K / V stream of the head in five
//               80-key chunks each through a shared-memory ring (128B-swizzled boxes of 64 dims).
//   warp 1      MMA issuer.  S = Q K^T: per key chunk 8 x tcgen05.mma 128x80x16 (K-major A and B) into
//               TMEM columns [80c, 80c+80) - the whole 128 x 400 score tile lives in TMEM, so the
//               softmax is exact two-pass, no online rescaling.  O = P V: 25 x tcgen05.mma 128x128x16
//               with A = P read straight from TMEM and B = V as an MN-major shared-memory operand.
//   warps 2..9  softmax + epilogue: a thread owns one query row (its TMEM lane) and half of the key
//               columns; row max and row sum are exchanged through shared memory; P is written back
//               to TMEM as packed fp16 (tcgen05.st) over the score columns it replaces; the output
//               accumulator reuses score columns that are dead by then (400 + 128 > 512 otherwise).
// TMEM map (512 columns): S [0,400) | P keys 0..207 -> [0,104), keys 208..399 -> [400,496) | O [256,384).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <stdexcept>
#include <tuple>

#include "net_ops.h"
#include "ptx.cuh"

namespace fpb {

namespace tc {

constexpr int kTok = 400;
constexpr int kHD = 128;
constexpr int kHeads = 4;
constexpr int kQT = 128;                     // query rows per work item
constexpr int kQTiles = (kTok + kQT - 1) / kQT;
constexpr int kKC = 80;                      // keys per K / V chunk
constexpr int kChunks = kTok / kKC;          // 5
constexpr int kHalfChunk = kKC * 128;        // one 64-dim half of a chunk: 80 rows x 128 B
constexpr int kChunkBytes = 2 * kHalfChunk;  // 20480
constexpr int kQHalf = kQT * 128;            // 16384
constexpr int kQBytes = 2 * kQHalf;          // 32768
constexpr int kRing = 7;
constexpr int kThreads = 320;
constexpr int kSplit = 208;                  // key columns [0,208) belong to softmax half 0
constexpr int kColO = 256;
constexpr int kColP1 = 400;
constexpr int kRedBytes = 2 * 2 * kQT * 4;   // row max / row sum exchange
constexpr size_t kSmem = 1024 + 2 * kQBytes + kRing * kChunkBytes + kRedBytes + 256;

struct Params {
  CUtensorMap tm_q;   // qkv as (1536, 400, n_hyp), box (64, 128, 1)
  CUtensorMap tm_kv;  // same tensor, box (64, 80, 1)
  __half* out;
  int out_ld;
  int n_items;  // n_hyp * heads * query tiles
};

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void named_barrier_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
attention_tc_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_base = smem_base;
  const uint32_t ring_base = q_base + 2 * kQBytes;
  const uint32_t red_base = ring_base + kRing * kChunkBytes;
  const uint32_t bar_base = red_base + kRedBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kRing + s); };
  auto qfull_bar = [&](int s) { return bar_base + 8u * (2 * kRing + s); };
  auto qempty_bar = [&](int s) { return bar_base + 8u * (2 * kRing + 2 + s); };
  const uint32_t sfull_bar = bar_base + 8u * (2 * kRing + 4);
  const uint32_t pfull_bar = bar_base + 8u * (2 * kRing + 5);
  const uint32_t ofull_bar = bar_base + 8u * (2 * kRing + 6);
  const uint32_t oempty_bar = bar_base + 8u * (2 * kRing + 7);
  const uint32_t tmem_slot = bar_base + 8u * (2 * kRing + 8);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));
  float* red = reinterpret_cast<float*>(smem_raw + (red_base - ptx::smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kRing; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(qfull_bar(s), 1);
      ptx::mbar_init(qempty_bar(s), 1);
    }
    ptx::mbar_init(sfull_bar, 1);
    ptx::mbar_init(pfull_bar, 8);
    ptx::mbar_init(ofull_bar, 1);
    ptx::mbar_init(oempty_bar, 8);
    ptx::fence_barrier_init();
    ptx::prefetch_tensormap(&p.tm_q);
    ptx::prefetch_tensormap(&p.tm_kv);
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
      const int qt = item % kQTiles, head = (item / kQTiles) % kHeads, hyp = item / (kQTiles * kHeads);
      const int qb = it & 1;
      ptx::mbar_wait(qempty_bar(qb), ((it >> 1) & 1) ^ 1u);
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(qfull_bar(qb), kQBytes);
        for (int h = 0; h < 2; ++h)
          ptx::tma_load_3d(q_base + qb * kQBytes + h * kQHalf, &p.tm_q, qfull_bar(qb), head * kHD + h * 64,
                           qt * kQT, hyp);
      }
      __syncwarp();
      for (int kv = 0; kv < 2; ++kv) {
        const int col0 = (kv + 1) * kHeads * kHD + head * kHD;
        for (int c = 0; c < kChunks; ++c) {
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          if (ptx::elect_one()) {
            ptx::mbar_expect_tx(full_bar(stage), kChunkBytes);
            for (int h = 0; h < 2; ++h)
              ptx::tma_load_3d(ring_base + stage * kChunkBytes + h * kHalfChunk, &p.tm_kv, full_bar(stage),
                               col0 + h * 64, c * kKC, hyp);
          }
          __syncwarp();
          if (++stage == kRing) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = ptx::make_idesc_f16_major(kQT, kKC, 0, 0);
    constexpr uint32_t idesc_o = ptx::make_idesc_f16_major(kQT, kHD, 0, 1);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
      const int qb = it & 1;
      const uint32_t par = it & 1;
      ptx::mbar_wait(oempty_bar, par ^ 1u);  // the previous item's O (and P) columns are drained
      ptx::mbar_wait(qfull_bar(qb), (it >> 1) & 1);
      ptx::tc_fence_after();
      // ---- S = Q K^T
      for (int c = 0; c < kChunks; ++c) {
        ptx::mbar_wait(full_bar(stage), phase);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint64_t adesc = ptx::make_sw128_kmajor_desc(q_base + qb * kQBytes + h * kQHalf);
            const uint64_t bdesc = ptx::make_sw128_kmajor_desc(ring_base + stage * kChunkBytes + h * kHalfChunk);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_f16(tmem_base + c * kKC, adesc + 2u * k, bdesc + 2u * k, idesc_s, (h | k) != 0);
          }
          ptx::umma_commit(empty_bar(stage));
          if (c == kChunks - 1) {
            ptx::umma_commit(qempty_bar(qb));
            ptx::umma_commit(sfull_bar);
          }
        }
        __syncwarp();
        if (++stage == kRing) {
          stage = 0;
          phase ^= 1u;
        }
      }
      // ---- O = P V
      ptx::mbar_wait(pfull_bar, par);
      ptx::tc_fence_after();
      for (int c = 0; c < kChunks; ++c) {
        ptx::mbar_wait(full_bar(stage), phase);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < kKC / 16; ++k) {
            const int i = c * (kKC / 16) + k;  // 16-key step
            const uint32_t a_tmem = tmem_base + (i < kSplit / 16 ? 8 * i : kColP1 + 8 * (i - kSplit / 16));
            const uint64_t bdesc =
                ptx::make_sw128_mnmajor_desc(ring_base + stage * kChunkBytes + k * 2048, kHalfChunk);
            ptx::umma_f16_ts(tmem_base + kColO, a_tmem, bdesc, idesc_o, i != 0);
          }
          ptx::umma_commit(empty_bar(stage));
          if (c == kChunks - 1) ptx::umma_commit(ofull_bar);
        }
        __syncwarp();
        if (++stage == kRing) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else {
    const int ew = warp - 2;
    const int q = warp & 3;   // TMEM lane quarter this warp may access
    const int half = ew >> 2; // key-column half (softmax) / dim half (output)
    const int row = q * 32 + lane;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t s_col = half ? kSplit : 0;
    const uint32_t p_col = half ? kColP1 : 0;
    float* red_max = red;             // [2][128]
    float* red_sum = red + 2 * kQT;   // [2][128]
    const float c_log2 = 0.08838834764831845f * 1.4426950408889634f;  // 1/sqrt(128) * log2(e)
    int it = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
      const int qt = item % kQTiles, head = (item / kQTiles) % kHeads, hyp = item / (kQTiles * kHeads);
      const uint32_t par = it & 1;
      ptx::mbar_wait(sfull_bar, par);
      ptx::tc_fence_after();
      // ---- pass 1: row maximum over this half's key columns
      float m = -INFINITY;
      {
        uint32_t ra[32], rb[32];
        ptx::tmem_ld_32x32b_x32(tlane + s_col, ra);
#pragma unroll
        for (int j = 0; j < 6; j += 2) {
          ptx::tmem_ld_wait();
          ptx::tmem_ld_32x32b_x32(tlane + s_col + 32 * (j + 1), rb);
#pragma unroll
          for (int e = 0; e < 32; ++e) m = fmaxf(m, __uint_as_float(ra[e]));
          ptx::tmem_ld_wait();
          if (j + 2 < 6) ptx::tmem_ld_32x32b_x32(tlane + s_col + 32 * (j + 2), ra);
#pragma unroll
          for (int e = 0; e < 32; ++e) m = fmaxf(m, __uint_as_float(rb[e]));
        }
        if (half == 0) {
          uint32_t rc[16];
          ptx::tmem_ld_32x32b_x16(tlane + 192, rc);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) m = fmaxf(m, __uint_as_float(rc[e]));
        }
      }
      red_max[half * kQT + row] = m;
      named_barrier_sync(1, 256);
      m = fmaxf(m, red_max[(half ^ 1) * kQT + row]);
      const float msc = m * c_log2;
      // ---- pass 2: P = exp2(S c - m c) as packed fp16 back into TMEM, fp32 row sum
      float sum = 0.f;
      {
        auto emit = [&](const uint32_t (&r)[32], int j) {
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float a = ptx::ex2_approx(fmaf(__uint_as_float(r[2 * e]), c_log2, -msc));
            const float b = ptx::ex2_approx(fmaf(__uint_as_float(r[2 * e + 1]), c_log2, -msc));
            sum += a + b;
            pk[e] = pack_half2(a, b);
          }
          ptx::tmem_st_32x32b_x16(tlane + p_col + 16 * j, pk);
        };
        uint32_t ra[32], rb[32];
        ptx::tmem_ld_32x32b_x32(tlane + s_col, ra);
#pragma unroll
        for (int j = 0; j < 6; j += 2) {
          ptx::tmem_ld_wait();
          ptx::tmem_ld_32x32b_x32(tlane + s_col + 32 * (j + 1), rb);
          emit(ra, j);
          ptx::tmem_ld_wait();
          if (j + 2 < 6) ptx::tmem_ld_32x32b_x32(tlane + s_col + 32 * (j + 2), ra);
          emit(rb, j + 1);
        }
        if (half == 0) {
          uint32_t rc[16];
          ptx::tmem_ld_32x32b_x16(tlane + 192, rc);
          ptx::tmem_ld_wait();
          uint32_t pk[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float a = ptx::ex2_approx(fmaf(__uint_as_float(rc[2 * e]), c_log2, -msc));
            const float b = ptx::ex2_approx(fmaf(__uint_as_float(rc[2 * e + 1]), c_log2, -msc));
            sum += a + b;
            pk[e] = pack_half2(a, b);
          }
          ptx::tmem_st_32x32b_x8(tlane + 96, pk);
        }
      }
      ptx::tmem_st_wait();
      red_sum[half * kQT + row] = sum;
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(pfull_bar);
      named_barrier_sync(1, 256);
      const float inv = 1.f / (sum + red_sum[(half ^ 1) * kQT + row]);
      // ---- epilogue: this thread's row, dims [64 half, 64 half + 64)
      ptx::mbar_wait(ofull_bar, par);
      ptx::tc_fence_after();
      const int tok = qt * kQT + row;
      __half* dst = p.out + (static_cast<long long>(hyp) * kTok + tok) * p.out_ld + head * kHD + half * 64;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(tlane + kColO + half * 64 + cc * 32, r);
        ptx::tmem_ld_wait();
        if (tok < kTok) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            uint4 u;
            u.x = pack_half2(__uint_as_float(r[8 * v + 0]) * inv, __uint_as_float(r[8 * v + 1]) * inv);
            u.y = pack_half2(__uint_as_float(r[8 * v + 2]) * inv, __uint_as_float(r[8 * v + 3]) * inv);
            u.z = pack_half2(__uint_as_float(r[8 * v + 4]) * inv, __uint_as_float(r[8 * v + 5]) * inv);
            u.w = pack_half2(__uint_as_float(r[8 * v + 6]) * inv, __uint_as_float(r[8 * v + 7]) * inv);
            *reinterpret_cast<uint4*>(dst + cc * 32 + v * 8) = u;
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(oempty_bar);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
int g_num_sms = 0;
std::mutex g_mu;
std::map<std::tuple<const void*, void*, int, int, int>, Params> g_params;

const char* init() {
  if (g_encode) return nullptr;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || !fn)
    return "cuTensorMapEncodeTiled not available from the CUDA driver";
  if (cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem) !=
      cudaSuccess)
    return "cudaFuncSetAttribute(attention smem) failed (not an sm_100a device?)";
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return nullptr;
}

const char* make_params(const __half* qkv, __half* out, int n_hyp, int qkv_ld, int out_ld, Params* p) {
  cuuint64_t dims[3] = {(cuuint64_t)qkv_ld, (cuuint64_t)kTok, (cuuint64_t)n_hyp};
  cuuint64_t strides[2] = {(cuuint64_t)qkv_ld * 2, (cuuint64_t)qkv_ld * 2 * kTok};
  cuuint32_t es[3] = {1, 1, 1};
  cuuint32_t box_q[3] = {64, kQT, 1}, box_kv[3] = {64, kKC, 1};
  if (g_encode(&p->tm_q, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)qkv, dims, strides, box_q, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return "attention: cuTensorMapEncodeTiled(Q) failed";
  if (g_encode(&p->tm_kv, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)qkv, dims, strides, box_kv, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return "attention: cuTensorMapEncodeTiled(KV) failed";
  p->out = out;
  p->out_ld = out_ld;
  p->n_items = n_hyp * kHeads * kQTiles;
  return nullptr;
}

}  // namespace tc

const char* attention_init() { return tc::init(); }

void launch_attention(const __half* qkv, __half* out, int n_hyp, int qkv_ld, int out_ld, cudaStream_t s) {
  if (n_hyp <= 0) return;
  if (const char* e = tc::init()) throw std::runtime_error(e);
  tc::Params prm;
  {
    std::lock_guard<std::mutex> lk(tc::g_mu);
    const auto key = std::make_tuple((const void*)qkv, (void*)out, n_hyp, qkv_ld, out_ld);
    auto it = tc::g_params.find(key);
    if (it == tc::g_params.end()) {
      if (const char* e = tc::make_params(qkv, out, n_hyp, qkv_ld, out_ld, &prm)) throw std::runtime_error(e);
      tc::g_params[key] = prm;
    } else {
      prm = it->second;
    }
  }
  const int grid = prm.n_items < tc::g_num_sms ? prm.n_items : tc::g_num_sms;
  tc::attention_tc_kernel<<<grid, tc::kThreads, tc::kSmem, s>>>(prm);
}

}  // namespace fpb
