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
// TMEM map (512 columns): S [0,400) | O [384,512) | P written in place over score columns its producer
// has already read: keys 0..207 (half 0, ascending groups of 32) -> [0,104); keys 208..399 (half 1,
// DESCENDING groups of 32) -> the lower 16 columns of each group's own 32, so that columns [384,400) are
// dead after the first step and the P V MMAs can start while the rest of the softmax is still running.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdlib>
#include <mutex>
#include <stdexcept>

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
constexpr int kRing = 6;
constexpr int kThreads = 320;
// V chunks are loaded (and their ring stages released) in the order the P V steps finish with them: the
// two softmax halves walk the keys from both ends towards the middle, so the order is 0, 4, 1, 3, 2.
__device__ __constant__ int kVOrder[5] = {0, 4, 1, 3, 2};
constexpr int kSplit = 208;                  // key columns [0,208) belong to softmax half 0
constexpr int kColO = 384;                   // output accumulator: columns [384,512)
constexpr int kColP1 = 208;                  // P of column half 1, key group g: columns [208 + 32 g, 224 + 32 g)
constexpr int kRedBytes = 2 * 2 * kQT * 4;   // row max / row sum exchange
constexpr int kStageBytes = kQT * kHD * 2;   // output tile staging: 8 warps x (32 rows x 128 B)
constexpr size_t kSmem = 1024 + 2 * kQBytes + kRing * kChunkBytes + kStageBytes + kRedBytes + 256;

using Params = AttentionPlan;

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void named_barrier_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// Work order.  Items are (hyp, head, tile) with the tile index minor, handed out in rounds of gridDim.x
// consecutive items, so the four query tiles of a (hyp, head) run at the same time on neighbouring CTAs and
// share its K / V through L2.  Within round k CTA b takes slot (b + k) mod gridDim.x: a plain b would pin
// a CTA to one tile index (148 % 4 == 0) and leave a quarter of the CTAs with only the cheap 16-row tail
// tiles.  Returns -1 when the CTA is done.
__device__ __forceinline__ int round_item(int k, int n_items) {
  const int base = k * (int)gridDim.x;
  if (base >= n_items) return -1;
  const int item = base + (int)((blockIdx.x + k) % gridDim.x);
  return item < n_items ? item : -1;  // only the last round can be partial
}
__device__ __forceinline__ void decode_item(int item, int& hyp, int& head, int& qt) {
  qt = item % kQTiles;
  head = (item / kQTiles) % kHeads;
  hyp = item / (kQTiles * kHeads);
}

__global__ void __launch_bounds__(kThreads, 1)
attention_tc_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_base = smem_base;
  const uint32_t ring_base = q_base + 2 * kQBytes;
  const uint32_t stage_base = ring_base + kRing * kChunkBytes;
  const uint32_t red_base = stage_base + kStageBytes;
  const uint32_t bar_base = red_base + kRedBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kRing + s); };
  auto qfull_bar = [&](int s) { return bar_base + 8u * (2 * kRing + s); };
  auto qempty_bar = [&](int s) { return bar_base + 8u * (2 * kRing + 2 + s); };
  const uint32_t sfull_bar = bar_base + 8u * (2 * kRing + 4);
  // one barrier per softmax step: the softmax warps may run several steps ahead of the MMA warp, which a
  // single phase bit could not express
  auto pstep_bar = [&](int j) { return bar_base + 8u * (2 * kRing + 9 + j); };
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
    for (int j = 0; j < 6; ++j) ptx::mbar_init(pstep_bar(j), 8);
    ptx::mbar_init(ofull_bar, 1);
    ptx::mbar_init(oempty_bar, 8);
    ptx::fence_barrier_init();
    ptx::prefetch_tensormap(&p.tm_q);
    ptx::prefetch_tensormap(&p.tm_kv);
    ptx::prefetch_tensormap(&p.tm_out);
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // programmatic dependent launch (plan.pdl): the prologue above may overlap the QKV projection's tail; no-ops otherwise
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int item; (item = round_item(it, p.n_items)) >= 0; ++it) {
      int qt, head, hyp;
      decode_item(item, hyp, head, qt);
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
        for (int cc = 0; cc < kChunks; ++cc) {
          const int c = kv ? kVOrder[cc] : cc;
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
    const bool timed = p.phase_clocks != nullptr && blockIdx.x == 0 && lane == 0;
    long long mk[6] = {0, 0, 0, 0, 0, 0};
    long long m_prev = timed ? clock64() : 0;
    auto mtick = [&](int k) {
      if (timed) {
        const long long t = clock64();
        mk[k] += t - m_prev;
        m_prev = t;
      }
    };
    for (int item; (item = round_item(it, p.n_items)) >= 0; ++it) {
      const int qb = it & 1;
      const uint32_t par = it & 1;
      mtick(5);
      ptx::mbar_wait(qfull_bar(qb), (it >> 1) & 1);
      ptx::tc_fence_after();
      mtick(0);
      // ---- S = Q K^T.  Chunks 0..3 land in columns [0,320): the previous item's P there was consumed by
      // its PV MMAs, which this thread issued earlier (the tensor pipe executes them in order), so they
      // overlap the previous epilogue.  Chunk 4 overwrites previous O columns [384,400): wait for the
      // epilogue warps to have drained O.
      for (int c = 0; c < kChunks; ++c) {
        if (c == 4) {
          mtick(5);
          ptx::mbar_wait(oempty_bar, par ^ 1u);
          ptx::tc_fence_after();
          mtick(2);
        }
        mtick(5);
        ptx::mbar_wait(full_bar(stage), phase);
        ptx::tc_fence_after();
        mtick(1);
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
      // ---- O = P V, issued while the softmax warps are still producing P: they publish it in six steps
      // (32 keys of each column half per step, pstep_bar), and the 16-key MMA steps whose P is complete are
      // issued right away (accumulation order is free).  All five V chunks are resident by then (they were
      // loaded behind the K chunks while S / the row maxima were computed); a chunk's ring stage is released
      // at its last use.
      // (everything below is unrolled so that chunk / step indices are compile-time: the issuing thread's own
      // instruction stream is on the critical path of every softmax step)
      int vst[kChunks];
      uint64_t vdesc[kChunks];  // MN-major descriptor of each V chunk's first 16-key step
      {
        int st = stage;
        uint32_t ph = phase;
        mtick(5);
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          constexpr int order[kChunks] = {0, 4, 1, 3, 2};  // = kVOrder
          vst[order[c]] = st;
          vdesc[order[c]] = ptx::make_sw128_mnmajor_desc(ring_base + st * kChunkBytes, kHalfChunk);
          ptx::mbar_wait(full_bar(st), ph);
          mtick(3);
          if (++st == kRing) {
            st = 0;
            ph ^= 1u;
          }
        }
        stage = st;
        phase = ph;
      }
      ptx::tc_fence_after();
      auto issue = [&](int i) {  // 16-key step i: keys [16 i, 16 i + 16)
        const int c = i / (kKC / 16), k = i % (kKC / 16);
        const int i1 = i - kSplit / 16;  // step within half 1: group i1 / 2
        const uint32_t a_tmem = tmem_base + (i1 < 0 ? 8 * i : kColP1 + 32 * (i1 >> 1) + 8 * (i1 & 1));
        // + k * 2048 bytes in the 16-byte units of the start-address field (no carry out of its 14 bits)
        ptx::umma_f16_ts(tmem_base + kColO, a_tmem, vdesc[c] + static_cast<uint64_t>(k * (2048 >> 4)), idesc_o, i != 0);
      };
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        mtick(5);
        ptx::mbar_wait(pstep_bar(j), par);
        ptx::tc_fence_after();
        mtick(4);
        if (ptx::elect_one()) {
          // step j: half 0 published keys [32 j, 32 j + 32) (+ [192,208) at j = 5), half 1 its group 5 - j
          issue(2 * j);
          issue(2 * j + 1);
          if (j == 5) issue(12);
          issue(13 + 2 * (5 - j));
          issue(14 + 2 * (5 - j));
          // last uses: chunks 0 (steps 0-4) and 4 (20-24) at j = 2, chunks 1 (5-9) and 3 (15-19) at j = 4,
          // chunk 2 (10-14) at j = 5
          if (j == 2) {
            ptx::umma_commit(empty_bar(vst[0]));
            ptx::umma_commit(empty_bar(vst[4]));
          }
          if (j == 4) {
            ptx::umma_commit(empty_bar(vst[1]));
            ptx::umma_commit(empty_bar(vst[3]));
          }
          if (j == 5) {
            ptx::umma_commit(empty_bar(vst[2]));
            ptx::umma_commit(ofull_bar);
          }
        }
        __syncwarp();
      }
    }
    if (timed)
      for (int k = 0; k < 6; ++k) p.phase_clocks[16 + k] = (unsigned long long)mk[k];
  } else {
    const int ew = warp - 2;
    const int q = warp & 3;   // TMEM lane quarter this warp may access
    const int half = ew >> 2; // key-column half (softmax) / dim half (output)
    const int row = q * 32 + lane;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t s_col = half ? kSplit : 0;
    float* red_max = red;             // [2][128]
    float* red_sum = red + 2 * kQT;   // [2][128]
    const float c_log2 = 0.08838834764831845f * 1.4426950408889634f;  // 1/sqrt(128) * log2(e)
    int it = 0;
    const bool timed = p.phase_clocks != nullptr && blockIdx.x == 0 && threadIdx.x == 64;
    long long tk[14] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    int tk_off = 0, n_tail_items = 0;
    long long t_prev = timed ? clock64() : 0;
    auto tick = [&](int k) {
      if (timed) {
        const long long t = clock64();
        tk[tk_off + k] += t - t_prev;
        t_prev = t;
      }
    };
    for (int item; (item = round_item(it, p.n_items)) >= 0; ++it) {
      int qt, head, hyp;
      decode_item(item, hyp, head, qt);
      const uint32_t par = it & 1;
      // The last query tile of a head holds 400 - 384 = 16 valid rows: warps whose 32 rows are all padding
      // keep every barrier but skip the TMEM traffic and the exponentials (their P / O rows stay garbage,
      // which is harmless: MMA rows are independent and the rows are never stored).
      const bool active = qt * kQT + q * 32 < kTok;
      tk_off = qt == kQTiles - 1 ? 7 : 0;
      n_tail_items += qt == kQTiles - 1;
      ptx::mbar_wait(sfull_bar, par);
      ptx::tc_fence_after();
      tick(0);  // waiting for S
      // ---- pass 1: row maximum over this half's key columns
      float m = -INFINITY;
      if (active) {
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
      tick(1);  // pass 1 (row max) + exchange
      m = fmaxf(m, red_max[(half ^ 1) * kQT + row]);
      const float msc = m * c_log2;
      // ---- pass 2: P = exp2(S c - m c) as packed fp16 back into TMEM, fp32 row sum
      float sum = 0.f;
      if (!active) {
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0)
          for (int j = 0; j < 6; ++j) ptx::mbar_arrive(pstep_bar(j));
      } else {
        auto publish = [&](int j) {  // this warp's P columns of step j are in TMEM
          ptx::tmem_st_wait();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(pstep_bar(j));
        };
        auto emit = [&](const uint32_t (&r)[32], int j) {
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float a = ptx::ex2_approx(fmaf(__uint_as_float(r[2 * e]), c_log2, -msc));
            const float b = ptx::ex2_approx(fmaf(__uint_as_float(r[2 * e + 1]), c_log2, -msc));
            sum += a + b;
            pk[e] = pack_half2(a, b);
          }
          // half 0: packed to the front of its range; half 1: lower half of the group's own columns
          ptx::tmem_st_32x32b_x16(tlane + (half ? kColP1 + 32 * (5 - j) : 16 * j), pk);
          if (j < 5) publish(j);
        };
        // step j handles key group j of half 0 and key group 5 - j of half 1 (see the TMEM map)
        auto grp_col = [&](int j) { return tlane + s_col + 32 * (half ? 5 - j : j); };
        uint32_t ra[32], rb[32];
        ptx::tmem_ld_32x32b_x32(grp_col(0), ra);
#pragma unroll
        for (int j = 0; j < 6; j += 2) {
          ptx::tmem_ld_wait();
          ptx::tmem_ld_32x32b_x32(grp_col(j + 1), rb);
          emit(ra, j);
          ptx::tmem_ld_wait();
          if (j + 2 < 6) ptx::tmem_ld_32x32b_x32(grp_col(j + 2), ra);
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
        publish(5);  // step 5 (+ keys [192,208) of half 0)
      }
      red_sum[half * kQT + row] = sum;
      named_barrier_sync(1, 256);
      tick(2);  // pass 2 (exp, P store) + exchange
      const float inv = 1.f / (sum + red_sum[(half ^ 1) * kQT + row]);
      // ---- epilogue: this thread's row, dims [64 half, 64 half + 64)
      ptx::mbar_wait(ofull_bar, par);
      ptx::tc_fence_after();
      tick(3);  // waiting for O = P V
      // TMEM -> registers -> this warp's private staging block (32 rows x 128 B, 16-byte chunks XOR-swizzled
      // by row = the 128-byte TMA swizzle) -> global memory with one TMA tensor store per warp (a lane-per-
      // row store would touch 32 rows per instruction; rows past token 399 are clipped by the tensor map).
      const uint32_t stg_addr = stage_base + ew * 4096;
      uint8_t* stg = smem_raw + (stg_addr - ptx::smem_u32(smem_raw));
      // O leaves TMEM first (64 registers per thread) so that the MMA warp can reuse its columns for the next
      // item's scores while this warp is still scaling / staging / storing.
      uint32_t r[2][32];
      if (active) {
        ptx::tmem_ld_32x32b_x32(tlane + kColO + half * 64, r[0]);
        ptx::tmem_ld_32x32b_x32(tlane + kColO + half * 64 + 32, r[1]);
        ptx::tmem_ld_wait();
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(oempty_bar);     // TMEM drained: the next item's S may overwrite O
        ptx::tma_store_wait_read();       // the previous item's store has read the staging block
      }
      __syncwarp();
      tick(4);  // epilogue: TMEM -> registers
      if (active) {
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            uint4 u;
            u.x = pack_half2(__uint_as_float(r[cc][8 * v + 0]) * inv, __uint_as_float(r[cc][8 * v + 1]) * inv);
            u.y = pack_half2(__uint_as_float(r[cc][8 * v + 2]) * inv, __uint_as_float(r[cc][8 * v + 3]) * inv);
            u.z = pack_half2(__uint_as_float(r[cc][8 * v + 4]) * inv, __uint_as_float(r[cc][8 * v + 5]) * inv);
            u.w = pack_half2(__uint_as_float(r[cc][8 * v + 6]) * inv, __uint_as_float(r[cc][8 * v + 7]) * inv);
            *reinterpret_cast<uint4*>(stg + lane * 128 + (((cc * 4 + v) ^ (lane & 7)) << 4)) = u;
          }
        }
        ptx::fence_proxy_async();  // staging writes -> visible to the TMA engine
      }
      __syncwarp();
      if (lane == 0 && active) {
        ptx::tma_store_3d(&p.tm_out, stg_addr, head * kHD + half * 64, qt * kQT + q * 32, hyp);
        ptx::tma_store_commit();
      }
      tick(5);  // epilogue: scale, stage, issue the store
    }
    if (lane == 0) ptx::tma_store_wait_read();
    if (timed) {
      for (int k = 0; k < 14; ++k) p.phase_clocks[k] = (unsigned long long)tk[k];
      p.phase_clocks[6] = (unsigned long long)(it - n_tail_items);
      p.phase_clocks[13] = (unsigned long long)n_tail_items;
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
std::mutex g_mu;

const char* init() {
  // kernel attributes are per device; the driver entry point is process-wide
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
  if (cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem) !=
      cudaSuccess)
    return "cudaFuncSetAttribute(attention smem) failed (not an sm_100a device?)";
  if (dev < 64) inited |= 1ull << dev;
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
  {
    cuuint64_t odims[3] = {(cuuint64_t)out_ld, (cuuint64_t)kTok, (cuuint64_t)n_hyp};
    cuuint64_t ostrides[2] = {(cuuint64_t)out_ld * 2, (cuuint64_t)out_ld * 2 * kTok};
    cuuint32_t box_o[3] = {64, 32, 1};
    if (g_encode(&p->tm_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)out, odims, ostrides, box_o, es,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return "attention: cuTensorMapEncodeTiled(out) failed";
  }
  p->n_items = n_hyp * kHeads * kQTiles;
  p->phase_clocks = nullptr;
  return nullptr;
}

}  // namespace tc

namespace tc2 {
const char* init();
const char* make_params(const __half* qkv, __half* out, int n_hyp, int qkv_ld, int out_ld, AttentionPlan* p);
void launch(const AttentionPlan& plan, cudaStream_t s);
int device_sms();
}  // namespace tc2

const char* attention_init() {
  if (const char* e = tc::init()) return e;
  return tc2::init();
}

namespace {
unsigned long long* g_phase_clocks = nullptr;  // device buffer, only with FPB200_ATTN_PHASES=1
}
// Debug: per-phase clock sums of CTA 0's softmax thread for the last launch, for full 128-row tiles and
// then for the 16-row tail tiles (2 x 7 values: wait S, pass 1, pass 2, wait O, epilogue to staging,
// store issue, items), two unused slots, then 6 sums of the MMA-issuing thread: wait Q, wait K chunks,
// wait O drained, wait V chunks, wait P steps, everything else (issue).  Returns false unless
// FPB200_ATTN_PHASES=1 was set before the launch.
bool attention_phase_clocks(unsigned long long out[22]) {
  if (!g_phase_clocks) return false;
  return cudaMemcpy(out, g_phase_clocks, 22 * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess;
}

const char* build_attention(const __half* qkv, __half* out, int n_hyp, int qkv_ld, int out_ld, AttentionPlan* plan) {
  if (n_hyp <= 0) return "attention: no hypotheses";
  if (const char* e = attention_init()) return e;
  if (const char* e = tc::make_params(qkv, out, n_hyp, qkv_ld, out_ld, plan)) return e;
  if (const char* e = tc2::make_params(qkv, out, n_hyp, qkv_ld, out_ld, plan)) return e;
  plan->grid_v1 = plan->n_items < tc2::device_sms() ? plan->n_items : tc2::device_sms();  // SM count of THIS device
  {
    const char* pdl = getenv("FPB200_PDL");  // see conv_gemm.cu: only launches that leave SMs idle
    plan->pdl = pdl ? pdl[0] == '1' : plan->n_items < tc2::device_sms();
  }
  // FPB200_ATTN_V2=1 selects the second design (attention2.cu: measured slower so far, kept with its tests - see
  // DESIGN.md section 8); read per plan, never per launch
  const char* v2 = getenv("FPB200_ATTN_V2");
  plan->version = v2 && v2[0] == '1' ? 2 : 1;
  return nullptr;
}

void launch_attention(const AttentionPlan& plan, cudaStream_t s) {
  if (plan.n_items <= 0) return;
  static const bool want_phases = getenv("FPB200_ATTN_PHASES") != nullptr;
  if (plan.version == 2 && !want_phases) {
    tc2::launch(plan, s);
    return;
  }
  tc::Params prm = plan;
  if (want_phases) {
    if (!g_phase_clocks) cudaMalloc(&g_phase_clocks, 24 * sizeof(unsigned long long));
    prm.phase_clocks = g_phase_clocks;
  }
  if (plan.version == 2) {
    tc2::launch(prm, s);
    return;
  }
  if (prm.pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(prm.grid_v1);
    cfg.blockDim = dim3(tc::kThreads);
    cfg.dynamicSmemBytes = tc::kSmem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, tc::attention_tc_kernel, prm);
    return;
  }
  tc::attention_tc_kernel<<<prm.grid_v1, tc::kThreads, tc::kSmem, s>>>(prm);
}

}  // namespace fpb
