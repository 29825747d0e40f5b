// Stem convolution (7 x 7, stride 2, 6 -> 64) as a direct tcgen05 convolution.
//
// The general implicit-GEMM kernel (conv_gemm.cu) feeds the stem one 64-wide K slab per filter row: 128 output pixels
// x 8 input pixels x 8 channels, 16 KB of TMA writes per slab for ~600 bytes of distinct input (neighbouring outputs
// share 6 of their 8 pixels).  Seven slabs per 128 x 64 tile made the layer shared-memory-write bound at 37 % of the
// tensor peak.  Here the A operand is not materialised at all:
//   * the packed crop is split by pixel parity by the TMA unit itself (element stride 2 along x), so a tile's input
//     patch lands in shared memory as two dense arrays P[parity][37 rows][11 px][8 ch] (6.5 KB each),
//   * an output pixel ox reads, for filter column kx, stored pixel 2 ox + kx + 1: within one parity that is pixel
//     ox + const - consecutive outputs read consecutive 16-byte pixels.  That is exactly the row pitch of a
//     NON-swizzled K-major UMMA operand (8 rows x 16 bytes per core matrix), so the shared-memory descriptor can point
//     straight into the patch: rows 16 B apart (ox), 8-row groups 2 patch rows apart (oy; stride-byte-offset 352 B),
//     and the second 8-channel K chunk of a K = 16 instruction is the NEXT pixel of the same parity (leading-byte-offset
//     16 B): two filter taps per MMA, operands overlapping in memory.
//   * weights are reordered to that tap order (kx 0, 2, 4, 6 | 1, 3, 5, zero per filter row), stay resident in shared
//     memory (56 KB, 128-byte swizzle) for the whole persistent CTA.
// Per 8 x 16-pixel output tile: 13 KB of TMA traffic (was 114 KB), 28 MMAs of 128 x 64 x 16, fp32 accumulators in
// tensor memory (4 stages), epilogue bias + ReLU + fp16 -> phase-split output rows of 128 bytes.
// Warp roles (192 threads): warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue.  Two CTAs per SM.
#include "stem_conv.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.h"
#include "preprocess.h"
#include "ptx.cuh"

namespace fpb {
namespace {

using namespace ptx;

constexpr int kThreads = 192;
constexpr int kStages = 3;                       // patch ring
constexpr int kAccStages = 4;                    // accumulators of 64 fp32 columns
constexpr int kTileW = 8, kTileH = 16;           // output pixels per tile (M = 128)
constexpr int kPatchRows = 2 * kTileH + 5;       // 37 input rows
constexpr int kPatchPx = 11;                     // pixels of one parity
constexpr int kPatchPitch = kPatchPx * 16;       // 176 B per patch row
constexpr int kPatchBytes = kPatchRows * kPatchPitch;  // 6512
constexpr int kPatchSlot = 6528;                 // 128-byte aligned slot
constexpr int kStageBytes = 2 * kPatchSlot;
constexpr int kWTileBytes = 64 * 128;            // one filter row: 64 outputs x 64 K
constexpr int kWBytes = 7 * kWTileBytes;
constexpr int kStagingOff = kWBytes + kStages * kStageBytes;  // epilogue transpose buffers: 4 warps x 32 rows x 64 B
constexpr int kBiasOff = kStagingOff + 4 * 32 * 64;
constexpr int kBarOff = kBiasOff + 256;
constexpr int kNumBars = 1 + 2 * kStages + 2 * kAccStages;
constexpr int kSmemBytes = kBarOff + kNumBars * 8 + 16 + 1024;  // + tmem slot + alignment slack
constexpr int kTilesPerImage = (80 / kTileW) * (80 / kTileH);   // 50

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// K-major operand without swizzle: core matrices of 8 rows x 16 bytes (rows 16 B apart), K chunks lbo bytes apart,
// 8-row groups sbo bytes apart.
__device__ __forceinline__ uint64_t make_nosw_kmajor_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (sm_100); layout type 0 = no swizzle
  return d;
}

__global__ void __launch_bounds__(kThreads, 2)
stem_conv_kernel(const __grid_constant__ StemConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t s_w = base, s_patch = base + kWBytes;
  const uint32_t bars = base + kBarOff;
  const uint32_t bar_w = bars, bar_full = bars + 8, bar_empty = bar_full + 8 * kStages, bar_accf = bar_empty + 8 * kStages,
                 bar_acce = bar_accf + 8 * kAccStages;
  const uint32_t tmem_slot = bars + 8 * kNumBars;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);
    }
    for (int i = 0; i < kAccStages; ++i) {
      mbar_init(bar_accf + 8 * i, 1);
      mbar_init(bar_acce + 8 * i, 4);  // one arrival per epilogue warp
    }
    fence_barrier_init();
    prefetch_tensormap(&p.tm_in);
    prefetch_tensormap(&p.tm_w);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kAccStages * 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ---------------- TMA producer
    if (elect_one()) {
      mbar_expect_tx(bar_w, kWBytes);
      for (int ky = 0; ky < 7; ++ky) tma_load_2d(s_w + ky * kWTileBytes, &p.tm_w, bar_w, ky * 64, 0);
    }
    int it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int s = it % kStages;
      mbar_wait(bar_empty + 8 * s, ((it / kStages) & 1) ^ 1);
      if (elect_one()) {
        const int img = tile / kTilesPerImage, r = tile - img * kTilesPerImage;
        const int ty = r / (80 / kTileW), tx = r - ty * (80 / kTileW);
        const int px0 = 2 * (tx * kTileW) + 1, row0 = 2 * (ty * kTileH) - 3;
        const uint32_t dst = s_patch + s * kStageBytes;
        if (p.experiment == 3) {
          mbar_arrive(bar_full + 8 * s);
        } else {
          mbar_expect_tx(bar_full + 8 * s, 2 * kPatchBytes);
          tma_load_4d(dst, &p.tm_in, bar_full + 8 * s, 0, px0, row0, img);                   // odd stored pixels
          tma_load_4d(dst + kPatchSlot, &p.tm_in, bar_full + 8 * s, 0, px0 + 1, row0, img);  // even stored pixels
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer
    constexpr uint32_t idesc = make_idesc_f16(128, 64);
    mbar_wait(bar_w, 0);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int s = it % kStages, acc = it % kAccStages;
      mbar_wait(bar_acce + 8 * acc, ((it / kAccStages) & 1) ^ 1);
      mbar_wait(bar_full + 8 * s, (it / kStages) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t patch = s_patch + s * kStageBytes;
        const uint32_t d_tmem = tmem_base + acc * 64;
#pragma unroll
        for (int ky = 0; ky < (p.experiment == 2 ? 1 : 7); ++ky) {
#pragma unroll
          for (int step = 0; step < 4; ++step) {
            // steps 0, 1: odd-parity taps (kx 0, 2) and (kx 4, 6); steps 2, 3: even-parity taps (kx 1, 3) and (kx 5, zero)
            const uint32_t a_addr = patch + (step >> 1) * kPatchSlot + ky * kPatchPitch + (step & 1) * 32;
            const uint64_t a_desc = make_nosw_kmajor_desc(a_addr, 16, 2 * kPatchPitch);
            const uint64_t b_desc = make_sw128_kmajor_desc(s_w + ky * kWTileBytes + step * 32);
            umma_f16(d_tmem, a_desc, b_desc, idesc, (ky | step) != 0);
          }
        }
        umma_commit(bar_empty + 8 * s);
        umma_commit(bar_accf + 8 * acc);
      }
      __syncwarp();
    }
  } else {
    // ---------------- epilogue: warp w reads TMEM lanes 32 (w % 4) .. (lane = GEMM row = oy_local * 8 + ox_local).
    // A lane holds one whole output row (64 channels); storing it directly would make every store instruction touch 32
    // different 128-byte lines with 16 bytes each (measured: 45 % of the kernel).  The 32 x 64-channel block of the warp
    // is therefore transposed through shared memory, half a row at a time, so that four lanes write 64 contiguous bytes
    // (full sectors) of one row per instruction.
    const int q = warp & 3;
    const uint32_t stg = base + kStagingOff + q * (32 * 64);
    const int rrow = lane >> 2, rchunk = lane & 3;  // read-back mapping: rows rrow + 8 k, 16-byte chunk rchunk
    int it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int acc = it % kAccStages;
      mbar_wait(bar_accf + 8 * acc, (it / kAccStages) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 64;
      uint32_t v[2][32];
      tmem_ld_32x32b_x32(taddr, v[0]);
      tmem_ld_32x32b_x32(taddr + 32, v[1]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acce + 8 * acc);
      const int img = tile / kTilesPerImage, r = tile - img * kTilesPerImage;
      const int ty = r / (80 / kTileW), tx = r - ty * (80 / kTileW);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          uint32_t pk[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = h * 32 + c8 * 8 + 2 * j;
            const float a = fmaxf(__uint_as_float(v[h][c8 * 8 + 2 * j]) + p.bias[c], 0.f);
            const float b = fmaxf(__uint_as_float(v[h][c8 * 8 + 2 * j + 1]) + p.bias[c + 1], 0.f);
            const __half2 hv = __floats2half2_rn(a, b);
            pk[j] = *reinterpret_cast<const uint32_t*>(&hv);
          }
          st_shared_v4(stg + lane * 64 + ((c8 ^ ((lane >> 1) & 3)) << 4), pk[0], pk[1], pk[2], pk[3]);
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int row = rrow + 8 * k;  // row of this warp's block
          uint4 val;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                       : "r"(stg + row * 64 + ((rchunk ^ ((row >> 1) & 3)) << 4)));
          const int m = q * 32 + row;
          const int oy = ty * kTileH + (m >> 3), ox = tx * kTileW + (m & 7);
          const int plane = (oy & 1) * 2 + (ox & 1);
          const long long opix = ((static_cast<long long>(img) * 4 + plane) * 40 + (oy >> 1)) * 40 + (ox >> 1);
          if (p.experiment != 1) *reinterpret_cast<uint4*>(p.out + opix * 64 + h * 32 + rchunk * 8) = val;
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kAccStages * 64);
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
int g_num_sms = 0;
std::mutex g_mu;

}  // namespace

const char* stem_conv_init() {
  static unsigned long long inited = 0;
  std::lock_guard<std::mutex> lk(g_mu);
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return "cudaGetDevice failed";
  if (dev < 64 && ((inited >> dev) & 1ull)) return nullptr;
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || !fn)
      return "cuTensorMapEncodeTiled not available from the CUDA driver";
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  if (cudaFuncSetAttribute(stem_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) != cudaSuccess)
    return "cudaFuncSetAttribute(stem_conv smem) failed (not an sm_100a device?)";
  cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  if (dev < 64) inited |= 1ull << dev;
  return nullptr;
}

void reorder_stem_weights(const __half* file_layout, __half* kernel_layout) {
  // file: [co][ky][8 px slots: zero, kx 0..6][8 ch]; kernel: [co][ky][kx 0, 2, 4, 6, 1, 3, 5, zero][8 ch]
  static const int slot_of[8] = {1, 3, 5, 7, 2, 4, 6, 0};  // kernel slot -> file slot (file slot 0 is the zero pixel)
  for (int co = 0; co < 64; ++co)
    for (int ky = 0; ky < 7; ++ky)
      for (int s = 0; s < 8; ++s)
        memcpy(kernel_layout + ((co * 7 + ky) * 8 + s) * 8, file_layout + ((co * 7 + ky) * 8 + slot_of[s]) * 8,
               8 * sizeof(__half));
}

const char* build_stem_conv(const __half* packed, int n_img, const __half* w_kernel_layout, const float* bias_host, __half* out,
                            StemConvPlan* plan) {
  if (const char* e = stem_conv_init()) return e;
  if (!packed || !w_kernel_layout || !bias_host || !out || n_img < 1) return "stem_conv: bad argument";
  StemConvParams& p = plan->p;
  memset(&p, 0, sizeof(p));
  {
    // (channel, stored pixel, row, image); pixels beyond the row / rows beyond the image read as zero (TMA out-of-bounds
    // fill), which is the convolution's zero padding; the 4 pad pixels stored left of every row are zero in memory
    const cuuint64_t pitch = (cuuint64_t)kStemPitch * kPackC * 2;
    cuuint64_t dims[4] = {8, (cuuint64_t)kStemPitch, (cuuint64_t)kCrop, (cuuint64_t)n_img};
    cuuint64_t strides[3] = {16, pitch, pitch * kCrop};
    cuuint32_t box[4] = {8, 2 * kPatchPx, (cuuint32_t)kPatchRows, 1};
    cuuint32_t es[4] = {1, 2, 1, 1};
    const CUresult r = g_encode(&p.tm_in, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(packed), dims, strides, box,
                                es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      static char msg[96];
      snprintf(msg, sizeof(msg), "stem_conv: cuTensorMapEncodeTiled(input) failed with CUresult %d", (int)r);
      return msg;
    }
  }
  {
    cuuint64_t dims[2] = {448, 64};
    cuuint64_t strides[1] = {448 * 2};
    cuuint32_t box[2] = {64, 64};
    cuuint32_t es[2] = {1, 1};
    const CUresult r = g_encode(&p.tm_w, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(w_kernel_layout), dims,
                                strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return "stem_conv: cuTensorMapEncodeTiled(weights) failed";
  }
  memcpy(p.bias, bias_host, sizeof(p.bias));
  p.out = out;
  p.n_tiles = n_img * kTilesPerImage;
  p.experiment = getenv("FPB200_STEM_EXPERIMENT") ? atoi(getenv("FPB200_STEM_EXPERIMENT")) : 0;
  plan->grid = p.n_tiles < 2 * g_num_sms ? p.n_tiles : 2 * g_num_sms;
  plan->flops = 2.0 * n_img * 80.0 * 80.0 * 64.0 * 294.0;
  return nullptr;
}

cudaError_t launch_stem_conv(const StemConvPlan& plan, cudaStream_t stream) {
  stem_conv_kernel<<<plan.grid, kThreads, kSmemBytes, stream>>>(plan.p);
  return cudaGetLastError();
}

}  // namespace fpb
