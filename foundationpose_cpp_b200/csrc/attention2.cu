// Self-attention core, second design (see attention.cu for the first one and for the op itself):
// softmax(Q K^T / sqrt(128)) V for 400 tokens, 4 heads x 128 dims per hypothesis.
//
// What bounded the first kernel: the whole 128 x 400 score tile had to sit in tensor memory between the row-maximum
// pass and the exponential pass, so the tensor pipe (S = Q K^T), the row maxima and the exponentials of one work item
// ran one after the other (10 600 clocks per item for 3 200 clocks of MMA and 3 200 clocks of MUFU), and every item
// streamed its head's K and V (200 KB) from L2 again - 22 B/clock/SM already; twice as fast would exceed what L2
// delivers (~42 B/clock/SM).
//
// This kernel
//   * keeps K and V of one (hypothesis, head) RESIDENT in shared memory (2 x 100 KB, 128B-swizzled TMA boxes) while
//     one CTA walks that head's query tiles: L2 traffic per item drops from 232 KB to 82 KB;
//   * keeps Q in TENSOR MEMORY (the A operand of S = Q K^T is read from there: `tcgen05.mma` with A in TMEM needs no
//     shared-memory bandwidth for it, and shared memory has no room left for Q anyway).  Q goes global -> registers
//     -> `tcgen05.st`: a thread owns one query row, and the packed-fp16 row IS the A-operand layout;
//   * computes S TWICE, in 64-key chunks through double-buffered 64-column windows: pass A (next item) only feeds the
//     row maximum, pass B (current item) feeds exp2 -> P (fp16, written over its own scores) -> O += P V.  No online
//     rescaling, same exact two-pass softmax as before, but nothing waits for a whole score tile any more: pass A of
//     item i+1, pass B of item i, the exponentials and the epilogue of item i-1 all overlap.  Tensor work per item
//     rises from 3 200 to 4 800 clocks - and becomes the bound.
//
// Warps (480 threads, 1 CTA / SM):
//   0      TMA producer: K / V chunks of the next (hypothesis, head) as soon as the last item has released them
//   1      MMA issuer of pass B and P V (one elected thread), owns the tensor-memory allocation
//   2      MMA issuer of pass A
//   3..10  "set B": exponentials of pass B, P back to tensor memory, row sums.  Two warps per scheduler (a thread owns
//          one query row and one 32-key half of every chunk): one warp per scheduler left the exp2 / convert / add
//          chain latency-bound at ~950 clocks per chunk
//   11..14 "set A": Q staging, row maximum of pass A, output epilogue (O / row sum -> fp16 -> global).  The higher
//          warp ids win the issue arbitration: their short, latency-critical work goes first.
// Tensor memory (512 columns): Q[2] 2 x 64 | O 128 | S_A[2] 2 x 64 | S_B[2] 2 x 64 (P in place: keys 0..31 of a chunk
// as packed fp16 in columns 0..15 of its window, keys 32..63 in columns 32..47).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdlib>
#include <mutex>

#include "net_ops.h"
#include "ptx.cuh"

namespace fpb {
namespace tc2 {

constexpr int kTok = 400;
constexpr int kHD = 128;
constexpr int kHeads = 4;
constexpr int kQT = 128;                      // query rows per item
constexpr int kQTiles = 4;                    // 3 full tiles + one with 16 valid rows
constexpr int kC = 64;                        // keys per chunk
constexpr int kNC = 7;                        // 6 x 64 + 16
constexpr int kHalfBytes = kTok * 128;        // one 64-dim half of K (or V): 400 rows x 128 B
constexpr int kKVBytes = 2 * kHalfBytes;      // 102400
constexpr int kChunkStride = kC * 128;        // 8192 B between chunks inside a half
constexpr int kThreads = 480;                  // 15 warps
constexpr uint32_t kColQ = 0, kColO = 128, kColSA = 256, kColSB = 384;
// (no staging block for the output: 208 KB of shared memory leave ~19 KB of L1 for the few local-memory slots)
constexpr size_t kSmem = 1024 + 2 * kKVBytes + 3 * 2 * kQT * 4 + 512;

// barrier indices
constexpr int kBarKFull = 0, kBarVFull = 7, kBarKFree = 14, kBarVFree = 21, kBarQFull = 28, kBarQFree = 30,
              kBarSAFull = 32, kBarSAFree = 34, kBarSBFull = 36, kBarPFull = 38, kBarMax = 40, kBarSum = 42,
              kBarOFull = 44, kBarOFree = 45, kNumBars = 46;

__device__ __forceinline__ int chunk_keys(int c) { return c == kNC - 1 ? kTok - (kNC - 1) * kC : kC; }
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// item `it` of this CTA: unit = it / tpc (a unit = tpc consecutive query tiles of one (hypothesis, head)); tpc = 1, 2, 4
struct Item {
  int hyp, head, qt;
  bool first, last;  // of its unit
};
__device__ __forceinline__ Item decode(int it, const AttentionPlan& p) {
  const int ls = p.tpc_log2;        // log2(tiles per CTA)
  const int uu = it >> ls, t = it & ((1 << ls) - 1);
  const int unit = (int)blockIdx.x + uu * (int)gridDim.x;
  const int group = unit >> (2 - ls);
  Item r;
  r.qt = ((unit & ((4 >> ls) - 1)) << ls) + t;
  r.head = group & (kHeads - 1);
  r.hyp = group >> 2;
  r.first = t == 0;
  r.last = t == (1 << ls) - 1;
  return r;
}

__global__ void __launch_bounds__(kThreads, 1) attention2_kernel(const __grid_constant__ AttentionPlan p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t k_base = smem_base;
  const uint32_t v_base = k_base + kKVBytes;
  const uint32_t red_base = v_base + kKVBytes;
  const uint32_t bar_base = red_base + 3 * 2 * kQT * 4;
  auto bar = [&](int i) { return bar_base + 8u * i; };
  const uint32_t tmem_slot = bar_base + 8u * kNumBars;
  uint8_t* const smem_al = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  float* const row_max = reinterpret_cast<float*>(smem_al + (red_base - smem_base));  // [2][128], already x c_log2
  float* const row_sum = row_max + 2 * kQT;                                           // [2 column halves][2][128]
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_al + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // units of this CTA: blockIdx.x, + gridDim.x, ...
  const int n_units_cta = (p.n_units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int n_it = n_units_cta << p.tpc_log2;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 28; ++i) ptx::mbar_init(bar(i), 1);  // K / V full (TMA) and free (tcgen05.commit)
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(bar(kBarQFull + s), 4);
      ptx::mbar_init(bar(kBarQFree + s), 1);
      ptx::mbar_init(bar(kBarSAFull + s), 1);
      ptx::mbar_init(bar(kBarSAFree + s), 4);
      ptx::mbar_init(bar(kBarSBFull + s), 1);
      ptx::mbar_init(bar(kBarPFull + s), 8);
      ptx::mbar_init(bar(kBarMax + s), 128);  // every thread that wrote a row maximum arrives itself
      ptx::mbar_init(bar(kBarSum + s), 256);  // ... and every thread that wrote a partial row sum
    }
    ptx::mbar_init(bar(kBarOFull), 1);
    ptx::mbar_init(bar(kBarOFree), 4);
    ptx::fence_barrier_init();
    ptx::prefetch_tensormap(&p.tm_kv64);
    ptx::prefetch_tensormap(&p.tm_kv16);
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
    // ------------------------------------------------------------------ TMA producer
    // order: a K chunk is released before the V chunk of the same keys; pass A of the next unit needs K first
    for (int uu = 0; uu < n_units_cta; ++uu) {
      const Item it0 = decode(uu << p.tpc_log2, p);
      const uint32_t par_free = (uint32_t)((uu & 1) ^ 1);  // released by the previous unit (first unit: passes)
      auto load = [&](int kv, int c) {
        ptx::mbar_wait_quiet(bar((kv ? kBarVFree : kBarKFree) + c), par_free);
        if (ptx::elect_one()) {
          const uint32_t full = bar((kv ? kBarVFull : kBarKFull) + c);
          const int rows = chunk_keys(c);
          ptx::mbar_expect_tx(full, 2u * rows * 128u);
          const CUtensorMap* tm = c == kNC - 1 ? &p.tm_kv16 : &p.tm_kv64;
          const int col0 = (kv + 1) * kHeads * kHD + it0.head * kHD;
          for (int h = 0; h < 2; ++h)
            ptx::tma_load_3d((kv ? v_base : k_base) + h * kHalfBytes + c * kChunkStride, tm, full, col0 + h * 64, c * kC,
                             it0.hyp);
        }
        __syncwarp();
      };
      load(0, 0);
      for (int c = 1; c < kNC; ++c) {
        load(0, c);
        load(1, c - 1);
      }
      load(1, kNC - 1);
      // the next unit's K / V into L2 now, a whole unit ahead: its chunk loads (issued one by one as this unit's last
      // item releases them) then cost an L2 hit instead of an HBM round trip each
      if (uu + 1 < n_units_cta && ptx::elect_one()) {
        const Item nx = decode((uu + 1) << p.tpc_log2, p);
        if (nx.hyp != it0.hyp || nx.head != it0.head) {
          for (int kv = 0; kv < 2; ++kv)
            for (int h = 0; h < 2; ++h)
              for (int c = 0; c < kNC; ++c)
                ptx::tma_prefetch_3d(c == kNC - 1 ? &p.tm_kv16 : &p.tm_kv64, (kv + 1) * kHeads * kHD + nx.head * kHD + h * 64,
                                     c * kC, nx.hyp);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1 || warp == 2) {
    // ------------------------------------------------------------------ MMA issuers
    // Two warps feed the tensor pipe: warp 2 issues pass A (S for the row maxima, one item ahead), warp 1 pass B
    // (S again, then O += P V).  The streams share nothing but read-only operands, so neither waits for the other's
    // consumers - one thread issuing both spent 9 500 clocks per item in its own instruction stream and waits.
    constexpr uint32_t idesc_s64 = ptx::make_idesc_f16_major(kQT, 64, 0, 0);
    constexpr uint32_t idesc_s16 = ptx::make_idesc_f16_major(kQT, 16, 0, 0);
    constexpr uint32_t idesc_o = ptx::make_idesc_f16_major(kQT, kHD, 0, 1);
    const uint64_t kdesc0 = ptx::make_sw128_kmajor_desc(k_base);
    const uint64_t vdesc0 = ptx::make_sw128_mnmajor_desc(v_base, kHalfBytes);
    const bool timed = p.phase_clocks != nullptr && blockIdx.x == 0 && lane == 0;
    long long mk0 = 0, mk1 = 0, mk2 = 0, mk3 = 0, mk4 = 0;
    long long m_prev = timed ? clock64() : 0;
#define FPB_MTICK(var)              \
  if (timed) {                      \
    const long long t_ = clock64(); \
    var += t_ - m_prev;             \
    m_prev = t_;                    \
  }
    // S chunk c = Q K_c^T into tensor-memory columns `col`; a0 = the item's Q columns.  (descriptor start addresses
    // are in 16-byte units: + 32 B per 16-dim step, + 8192 B per chunk, + 51200 B for the second 64-dim half)
    auto issue_s = [&](int c, uint32_t a0, uint32_t col) {
      const uint32_t idesc = c == kNC - 1 ? idesc_s16 : idesc_s64;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint64_t bdesc = kdesc0 + (uint64_t)((h * kHalfBytes + c * kChunkStride) >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          ptx::umma_f16_ts(tmem_base + col, a0 + 32u * h + 8u * k, bdesc + 2u * k, idesc, (h | k) != 0);
      }
    };
    if (warp == 2) {
      int ga = 0;  // running chunk counter of pass A (window = ga & 1)
      for (int it = 0; it < n_it; ++it) {
        const uint32_t kpar = (it >> p.tpc_log2) & 1;
        FPB_MTICK(mk2);
        ptx::mbar_wait_quiet(bar(kBarQFull + (it & 1)), (it >> 1) & 1);
        FPB_MTICK(mk0);
        const uint32_t a0 = tmem_base + kColQ + 64u * (it & 1);
#pragma unroll
        for (int c = 0; c < kNC; ++c) {
          ptx::mbar_wait_quiet(bar(kBarKFull + c), kpar);
          FPB_MTICK(mk0);
          ptx::mbar_wait_quiet(bar(kBarSAFree + (ga & 1)), ((ga >> 1) & 1) ^ 1u);
          FPB_MTICK(mk1);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            issue_s(c, a0, kColSA + 64u * (ga & 1));
            ptx::umma_commit(bar(kBarSAFull + (ga & 1)));
          }
          __syncwarp();
          ++ga;
          FPB_MTICK(mk2);
        }
      }
      if (timed) {
        p.phase_clocks[18] = (unsigned long long)mk0;
        p.phase_clocks[19] = (unsigned long long)mk1;
        p.phase_clocks[20] = (unsigned long long)mk2;
      }
    } else {
      int gb = 0;  // running chunk counter of pass B / P V (window = gb & 1; P V of chunk g follows S_B of chunk g + 1)
      // S_B chunk c of item `it` (flags of that item: last of its unit?)
      auto issue_sb = [&](int it, bool last, int c) {
        ptx::mbar_wait_quiet(bar(kBarKFull + c), (it >> p.tpc_log2) & 1);
        // chunk 0: Q staged, and pass A of this item finished (this stream releases the unit's K chunks as it goes:
        // pass A must have read them; the exponentials need the maxima anyway)
        if (c == 0) {
          ptx::mbar_wait_quiet(bar(kBarQFull + (it & 1)), (it >> 1) & 1);
          ptx::mbar_wait_quiet(bar(kBarMax + (it & 1)), (it >> 1) & 1);
        }
        FPB_MTICK(mk0);
        ptx::tc_fence_after();
        // the window (gb & 1) was last read by P V of chunk gb - 2, issued before this: the tensor pipe runs in order
        if (ptx::elect_one()) {
          issue_s(c, tmem_base + kColQ + 64u * (it & 1), kColSB + 64u * (gb & 1));
          ptx::umma_commit(bar(kBarSBFull + (gb & 1)));
          if (c == kNC - 1) ptx::umma_commit(bar(kBarQFree + (it & 1)));
          if (last) ptx::umma_commit(bar(kBarKFree + c));
        }
        __syncwarp();
        ++gb;
        FPB_MTICK(mk4);
      };
      bool pre = false;  // S_B chunk 0 of this item already issued
      int gp = 0;
      for (int it = 0; it < n_it; ++it) {
        const Item im = decode(it, p);
        const uint32_t kpar = (it >> p.tpc_log2) & 1;
        if (!pre) issue_sb(it, im.last, 0);
        pre = false;
#pragma unroll
        for (int c = 0; c < kNC; ++c) {
          if (c + 1 < kNC) {
            issue_sb(it, im.last, c + 1);
          } else if (it + 1 < n_it &&
                     __all_sync(0xffffffffu, ptx::mbar_try_wait(bar(kBarMax + ((it + 1) & 1)), ((it + 1) >> 1) & 1))) {
            // first pass-B chunk of the next item ahead of the last P V of this one (if its pass A is done)
            issue_sb(it + 1, decode(it + 1, p).last, 0);
            pre = true;
          }
          ptx::mbar_wait_quiet(bar(kBarVFull + c), kpar);
          FPB_MTICK(mk1);
          ptx::mbar_wait_quiet(bar(kBarPFull + (gp & 1)), (gp >> 1) & 1);
          FPB_MTICK(mk2);
          if (c == 0) ptx::mbar_wait_quiet(bar(kBarOFree), (it & 1) ^ 1u);
          FPB_MTICK(mk3);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint32_t a0 = tmem_base + kColSB + 64u * (gp & 1);
            const int steps = c == kNC - 1 ? (kTok - (kNC - 1) * kC) / 16 : kC / 16;
#pragma unroll
            for (int k = 0; k < steps; ++k)
              ptx::umma_f16_ts(tmem_base + kColO, a0 + 8u * (k & 1) + 32u * (k >> 1), vdesc0 + (uint64_t)((c * kChunkStride + k * 2048) >> 4), idesc_o,
                               (c | k) != 0);
            if (im.last) ptx::umma_commit(bar(kBarVFree + c));
            if (c == kNC - 1) ptx::umma_commit(bar(kBarOFull));
          }
          __syncwarp();
          ++gp;
          FPB_MTICK(mk4);
        }
      }
      if (timed) {
        p.phase_clocks[0] = (unsigned long long)mk0;  // wait K / Q
        p.phase_clocks[3] = (unsigned long long)mk1;  // wait V
        p.phase_clocks[4] = (unsigned long long)mk2;  // wait P
        p.phase_clocks[5] = (unsigned long long)mk3;  // wait O free
        p.phase_clocks[6] = (unsigned long long)mk4;  // issue / other
        p.phase_clocks[1] = 0;
        p.phase_clocks[2] = 0;
        p.phase_clocks[7] = (unsigned long long)n_it;
      }
    }
#undef FPB_MTICK
  } else if (warp >= 11) {
    // ------------------------------------------------------------------ set A: Q staging, row maximum, epilogue
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float c_log2 = 0.08838834764831845f * 1.4426950408889634f;  // 1/sqrt(128) * log2(e)
    uint4 qr[16];
    const bool timed = p.phase_clocks != nullptr && blockIdx.x == 0 && threadIdx.x == 12 * 32;  // warp 12 = quarter 0 of set A
    long long tk[6] = {0, 0, 0, 0, 0, 0};
    long long t_prev = timed ? clock64() : 0;
    auto tick = [&](int k) {
      if (timed) {
        const long long t = clock64();
        tk[k] += t - t_prev;
        t_prev = t;
      }
    };
    auto q_load = [&](const Item& im) {
      const int tok = im.qt * kQT + row;
      if (tok < kTok) {
        const uint4* src = reinterpret_cast<const uint4*>(p.qkv + ((size_t)im.hyp * kTok + tok) * p.qkv_ld + im.head * kHD);
#pragma unroll
        for (int i = 0; i < 16; ++i) qr[i] = __ldg(src + i);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) qr[i] = make_uint4(0u, 0u, 0u, 0u);
      }
    };
    auto q_store = [&](int it) {  // registers -> tensor memory (A-operand layout = the packed row), then publish
      const uint32_t dst = tlane + kColQ + 64u * (it & 1);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t r[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          r[4 * i + 0] = qr[4 * g + i].x;
          r[4 * i + 1] = qr[4 * g + i].y;
          r[4 * i + 2] = qr[4 * g + i].z;
          r[4 * i + 3] = qr[4 * g + i].w;
        }
        ptx::tmem_st_32x32b_x16(dst + 16u * g, r);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar(kBarQFull + (it & 1)));
    };
    int ga = 0;
    auto pass_a = [&](int it, const Item& im) {
      const bool active = im.qt * kQT + q * 32 < kTok;
      float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
      for (int c = 0; c < kNC; ++c) {
        const int buf = ga & 1;
        tick(1);
        ptx::mbar_wait_quiet(bar(kBarSAFull + buf), (ga >> 1) & 1);
        tick(0);
        ptx::tc_fence_after();
        if (active) {
          const uint32_t src = tlane + kColSA + 64u * buf;
          if (c < kNC - 1) {
            // 16-column loads, two in flight (the Q row after next occupies 64 registers meanwhile), four
            // independent maxima (a single FMNMX chain is 32 dependent instructions per chunk)
            uint32_t ra[16], rb[16];
            ptx::tmem_ld_32x32b_x16(src, ra);
#pragma unroll
            for (int g = 0; g < 4; g += 2) {
              ptx::tmem_ld_32x32b_x16(src + 16u * (g + 1), rb);
              ptx::tmem_ld_wait();  // (waits for both; the second is short)
#pragma unroll
              for (int e = 0; e < 16; e += 4) {
                m0 = fmaxf(m0, __uint_as_float(ra[e]));
                m1 = fmaxf(m1, __uint_as_float(ra[e + 1]));
                m2 = fmaxf(m2, __uint_as_float(ra[e + 2]));
                m3 = fmaxf(m3, __uint_as_float(ra[e + 3]));
              }
              if (g + 2 < 4) ptx::tmem_ld_32x32b_x16(src + 16u * (g + 2), ra);
#pragma unroll
              for (int e = 0; e < 16; e += 4) {
                m0 = fmaxf(m0, __uint_as_float(rb[e]));
                m1 = fmaxf(m1, __uint_as_float(rb[e + 1]));
                m2 = fmaxf(m2, __uint_as_float(rb[e + 2]));
                m3 = fmaxf(m3, __uint_as_float(rb[e + 3]));
              }
            }
          } else {
            uint32_t rc[16];
            ptx::tmem_ld_32x32b_x16(src, rc);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; e += 2) {
              m0 = fmaxf(m0, __uint_as_float(rc[e]));
              m1 = fmaxf(m1, __uint_as_float(rc[e + 1]));
            }
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(bar(kBarSAFree + buf));
        ++ga;
      }
      row_max[(it & 1) * kQT + row] = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * c_log2;
      ptx::mbar_arrive(bar(kBarMax + (it & 1)));  // release by the writer itself
    };
    auto epilogue = [&](int it, const Item& im) {
      const bool active = im.qt * kQT + q * 32 < kTok;
      tick(5);
      ptx::mbar_wait_quiet(bar(kBarOFull), it & 1);
      ptx::mbar_wait_quiet(bar(kBarSum + (it & 1)), (it >> 1) & 1);
      tick(3);
      ptx::tc_fence_after();
      const float inv =
          1.f / ((row_sum[(it & 1) * kQT + row] + row_sum[(2 + (it & 1)) * kQT + row]) * 5.192296858534828e33f);  // x 2^112
      // O leaves tensor memory 32 columns at a time; each group is scaled, packed to fp16 and stored right away (a lane
      // stores its own row, 64 contiguous bytes per group: an instruction touches 32 rows, the halves of every 32-byte
      // sector arrive back to back and merge in L2)
      const bool row_ok = im.qt * kQT + row < kTok;
      uint4* const orow = reinterpret_cast<uint4*>(p.out + ((size_t)im.hyp * kTok + im.qt * kQT + row) * p.out_ld + im.head * kHD);
#pragma unroll
      for (int g = 0; g < 8; ++g) {  // 16 columns at a time (the next Q row is still in 64 registers)
        uint32_t ra[16];
        if (active) {
          ptx::tmem_ld_32x32b_x16(tlane + kColO + 16u * g, ra);
          ptx::tmem_ld_wait();
        }
        if (g == 7) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(bar(kBarOFree));  // O is in registers
        }
        if (active && row_ok) {
#pragma unroll
          for (int v = 0; v < 2; ++v) {
            uint4 u;
            u.x = pack_half2(__uint_as_float(ra[8 * v + 0]) * inv, __uint_as_float(ra[8 * v + 1]) * inv);
            u.y = pack_half2(__uint_as_float(ra[8 * v + 2]) * inv, __uint_as_float(ra[8 * v + 3]) * inv);
            u.z = pack_half2(__uint_as_float(ra[8 * v + 4]) * inv, __uint_as_float(ra[8 * v + 5]) * inv);
            u.w = pack_half2(__uint_as_float(ra[8 * v + 6]) * inv, __uint_as_float(ra[8 * v + 7]) * inv);
            orow[2 * g + v] = u;
          }
        }
      }
    };

    if (n_it > 0) {
      Item i0 = decode(0, p);
      q_load(i0);
      q_store(0);
      Item i1 = i0;
      if (n_it > 1) {
        i1 = decode(1, p);
        q_load(i1);
      }
      pass_a(0, i0);
      if (n_it > 1) q_store(1);
      Item cur = i0, nxt = i1;
      for (int it = 0; it < n_it; ++it) {
        Item nn = nxt;
        if (it + 2 < n_it) {
          nn = decode(it + 2, p);
          q_load(nn);  // in flight during pass A
        }
        if (it + 1 < n_it) pass_a(it + 1, nxt);
        epilogue(it, cur);  // first: the next item's P V waits for O to be drained
        tick(4);
        if (it + 2 < n_it) {
          ptx::mbar_wait_quiet(bar(kBarQFree + (it & 1)), (it >> 1) & 1);  // pass B of item `it` has read Q[it & 1]
          tick(2);
          ptx::tc_fence_after();
          q_store(it + 2);
        }
        cur = nxt;
        nxt = nn;
      }
      if (timed)
        for (int k = 0; k < 6; ++k) p.phase_clocks[12 + k] = (unsigned long long)tk[k];
    }
  } else {
    // ------------------------------------------------------------------ set B: exponentials, P, row sums
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float c_log2 = 0.08838834764831845f * 1.4426950408889634f;
    // P = exp2(s c - m c) is produced as y = P * 2^-112: the fp32 exponent field of y IS the fp16 exponent field of P,
    // so fp16(P) = (bits(y) + 0x1000) >> 13 (round half up; P < 2^-14 flushes to zero) - integer-pipe work instead of
    // F2FP, which shares the quarter-rate unit with MUFU.EX2 (measured: 300 such operations per thread set the pace
    // of the first kernel's softmax).  The row sum is taken over y and rescaled in the epilogue.
    auto pack2 = [](float a, float b) {
      const uint32_t ua = (__float_as_uint(a) + 0x1000u) >> 13, ub = (__float_as_uint(b) + 0x1000u) >> 13;
      return __byte_perm(ua, ub, 0x5410);
    };
    int gb = 0;
    const bool timed = p.phase_clocks != nullptr && blockIdx.x == 0 && threadIdx.x == 4 * 32;  // warp 4 = quarter 0 of set B
    long long tk[4] = {0, 0, 0, 0};
    long long t_prev = timed ? clock64() : 0;
    auto tick = [&](int k) {
      if (timed) {
        const long long t = clock64();
        tk[k] += t - t_prev;
        t_prev = t;
      }
    };
    const int half = (warp - 3) >> 2;  // this thread's 32 keys of every 64-key chunk
    for (int it = 0; it < n_it; ++it) {
      const Item im = decode(it, p);
      const bool active = im.qt * kQT + q * 32 < kTok;
      tick(3);
      ptx::mbar_wait_quiet(bar(kBarMax + (it & 1)), (it >> 1) & 1);
      tick(0);
      const float msc = row_max[(it & 1) * kQT + row] + 112.f;
      float sum = 0.f;
      for (int c = 0; c < kNC; ++c) {
        const int buf = gb & 1;
        tick(3);
        ptx::mbar_wait_quiet(bar(kBarSBFull + buf), (gb >> 1) & 1);
        tick(0);
        ptx::tc_fence_after();
        if (active) {
          const uint32_t win = tlane + kColSB + 64u * buf;
          if (c < kNC - 1) {
            uint32_t ra[32];
            ptx::tmem_ld_32x32b_x32(win + 32u * half, ra);
            ptx::tmem_ld_wait();
            tick(1);
            uint32_t pa[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const float a = ptx::ex2_approx(fmaf(__uint_as_float(ra[2 * e]), c_log2, -msc));
              const float b = ptx::ex2_approx(fmaf(__uint_as_float(ra[2 * e + 1]), c_log2, -msc));
              sum += a + b;
              pa[e] = pack2(a, b);
            }
            tick(2);
            // P over the first 16 columns of this thread's OWN 32 score columns (no other warp reads those)
            ptx::tmem_st_32x32b_x16(win + 32u * half, pa);
          } else if (half == 0) {
            uint32_t rc[16];
            ptx::tmem_ld_32x32b_x16(win, rc);
            ptx::tmem_ld_wait();
            uint32_t pc[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float a = ptx::ex2_approx(fmaf(__uint_as_float(rc[2 * e]), c_log2, -msc));
              const float b = ptx::ex2_approx(fmaf(__uint_as_float(rc[2 * e + 1]), c_log2, -msc));
              sum += a + b;
              pc[e] = pack2(a, b);
            }
            ptx::tmem_st_32x32b_x8(win, pc);
          }
          ptx::tmem_st_wait();
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(bar(kBarPFull + buf));
        ++gb;
      }
      row_sum[(half * 2 + (it & 1)) * kQT + row] = sum;
      ptx::mbar_arrive(bar(kBarSum + (it & 1)));  // release by the writer itself
    }
    tick(3);
    if (timed)
      for (int k = 0; k < 4; ++k) p.phase_clocks[8 + k] = (unsigned long long)tk[k];
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
std::mutex g_mu;
int g_sms[64] = {0};

const char* init() {
  std::lock_guard<std::mutex> lk(g_mu);
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return "cudaGetDevice failed";
  if (dev < 64 && g_sms[dev]) return nullptr;
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || !fn)
      return "cuTensorMapEncodeTiled not available from the CUDA driver";
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  if (cudaFuncSetAttribute(attention2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem) != cudaSuccess)
    return "cudaFuncSetAttribute(attention2 smem) failed (not an sm_100a device?)";
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, attention2_kernel) != cudaSuccess) return "loading attention2_kernel failed";
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (dev < 64) g_sms[dev] = sms > 0 ? sms : 148;
  return nullptr;
}

// SM count of the current device (after init() on it)
int device_sms() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev >= 0 && dev < 64 && g_sms[dev] ? g_sms[dev] : 148;
}

const char* make_params(const __half* qkv, __half* out, int n_hyp, int qkv_ld, int out_ld, AttentionPlan* p) {
  const int sms = device_sms();
  cuuint64_t dims[3] = {(cuuint64_t)qkv_ld, (cuuint64_t)kTok, (cuuint64_t)n_hyp};
  cuuint64_t strides[2] = {(cuuint64_t)qkv_ld * 2, (cuuint64_t)qkv_ld * 2 * kTok};
  cuuint32_t es[3] = {1, 1, 1};
  cuuint32_t box64[3] = {64, 64, 1}, box16[3] = {64, 16, 1};
  if (g_encode(&p->tm_kv64, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)qkv, dims, strides, box64, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return "attention: cuTensorMapEncodeTiled(KV 64) failed";
  if (g_encode(&p->tm_kv16, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)qkv, dims, strides, box16, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return "attention: cuTensorMapEncodeTiled(KV 16) failed";
  p->qkv = qkv;
  p->qkv_ld = qkv_ld;
  p->out = out;
  p->out_ld = out_ld;
  // query tiles per CTA: all four of a (hypothesis, head) share its resident K / V; with few hypotheses the tiles are
  // spread over more CTAs instead.  Cost model (units of 1000 clocks): a unit costs ~4 to fill + ~5 per tile.
  const int groups = n_hyp * kHeads;
  int best = 4;
  double best_cost = 1e30;
  for (int tpc = 4; tpc >= 1; tpc >>= 1) {
    const int units = groups * (kQTiles / tpc);
    const int rounds = (units + sms - 1) / sms;
    const double cost = rounds * (4.0 + 5.0 * tpc);
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = tpc;
    }
  }
  const int force_tpc = getenv("FPB200_ATTN_TPC") ? atoi(getenv("FPB200_ATTN_TPC")) : 0;  // tests; read per plan
  if (force_tpc == 1 || force_tpc == 2 || force_tpc == 4) best = force_tpc;
  p->tiles_per_cta = best;
  p->tpc_log2 = best == 4 ? 2 : best == 2 ? 1 : 0;
  p->n_units = groups * (kQTiles / best);
  p->grid = p->n_units < sms ? p->n_units : sms;
  return nullptr;
}

void launch(const AttentionPlan& plan, cudaStream_t s) {
  attention2_kernel<<<plan.grid, kThreads, kSmem, s>>>(plan);
}

}  // namespace tc2
}  // namespace fpb
