// Fused tail of a post-LN transformer encoder layer on tcgen05 tensor cores (sm_100a):
//
//   x1   = LayerNorm1(att Wout^T + b_out + x0)          (self-attention out-projection + residual)
//   h    = ReLU(x1 W1^T + b1)                           (FFN, width 512)
//   y    = LayerNorm2(h W2^T + b2 + x1)
//   mean over the 400 tokens of each hypothesis of y    (the pose heads only ever consume the token mean)
//
// for att / x0 : [rows, 512] fp16, rows = hypotheses x 400 tokens.  In the reference these layers run inside the
// refiner's TensorRT engine (easy_deploy_tool/inference_core/trt_core/src/trt_core.cpp:323-371; structure:
// SURVEY.md section 2.3, nn.TransformerEncoderLayer(d=512, heads=4, ffn=512, post-LN, ReLU)).
//
// Unfused this chain was five launches (3 GEMMs + 2 LayerNorm passes) that moved 1.24 GB through HBM per encoder
// layer at N = 252 and ran under the ridge point; here a CTA keeps its 128 token rows ON CHIP through all three
// GEMMs: HBM traffic is the two inputs (206 MB) plus 13 MB of partial sums.
//
// Per CTA (320 threads, persistent over 128-row blocks; PAIR: two CTAs of a cluster issue ONE
// tcgen05.mma.cta_group::2 of M = 256 and each ingests only half of every weight tile):
//   warp 0      TMA producer: the block's attention rows (8 k-block tiles of [128 x 64] into the A buffer), the
//               weight stream (3 x 4 N-tiles x 8 k-blocks of [128 x 64]) through a ring, and per N-tile of GEMM 1
//               the residual slab x0[:, 128 columns].
//   warp 1      MMA issuer.  GEMM 1 and 2 read A from shared memory (SW128 K-major tiles), GEMM 3 reads A = h
//               straight from TENSOR MEMORY (packed fp16), so h never touches shared memory.
//   warps 2..9  epilogue, a thread owns one token row (its TMEM lane); the two warps of a lane quarter split the
//               columns.  Pass A (per 128-column accumulator tile, overlapped with the MMAs of the next tile through
//               two accumulator stages): + bias, + residual, ReLU, round to fp16, row statistics.  Pass B (per GEMM
//               1 / 3): LayerNorm of the whole 512-wide row (row statistics are thread-local: a TMEM lane is a row).
// Tensor memory (512 columns): accumulator stages [0,128) [128,256) | R [256,512) = one [128 x 512] fp16 matrix
// (two values per column) holding, in turn, the pre-LayerNorm sum of GEMM 1 and the FFN activation h.
// Shared memory: A buffer 128 KB (att -> x1 -> pre-LayerNorm sum of GEMM 3, each written in place over its
// predecessor), weight ring, residual slab 32 KB.
// Hand-offs inside a block, all by mbarrier: x1 k-block kb is published by the LayerNorm pass as soon as all rows
// have it (GEMM 2 starts while LayerNorm 1 is still running); h k-block kb likewise from GEMM 2's pass A.
//
// Token mean: rows are summed in groups of 16 consecutive tokens (400 = 25 x 16, so a group never straddles two
// hypotheses and its position inside a 128-row block does not matter) by a fixed recursive-halving shuffle tree;
// encoder_mean_kernel adds the 25 group sums of a hypothesis in index order.  The result is therefore independent
// of how many hypotheses share a launch and of where a hypothesis starts (sharded runs stay bit-identical).
#include "encoder_block.h"

#include "common.h"

#include <cstdio>
#include <cstring>
#include <mutex>
#include <type_traits>

#include "ptx.cuh"

namespace fpb {

namespace {

constexpr int kD = 512;                   // model width = FFN width
constexpr int kBM = 128;                  // token rows per CTA
constexpr int kBN = 128;                  // accumulator tile width
constexpr int kNT = kD / kBN;             // 4 N-tiles per GEMM
constexpr int kKB = kD / 64;              // 8 k-blocks of 64 channels
constexpr int kTileA = kBM * 128;         // one k-block tile of the row block: 128 rows x 128 B
constexpr int kABytes = kKB * kTileA;     // 128 KB
constexpr int kSlabBytes = 2 * kTileA;    // residual slab: 128 rows x 128 columns
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kAccCols = 2 * kBN;         // two accumulator stages
constexpr int kColR = kAccCols;           // R: packed fp16 [128 x 512] in columns [256, 512)
constexpr int kGroup = 16;                // tokens per partial sum
constexpr int kBarBytes = 512 + 4 * kBM * 4;  // mbarriers + TMEM slot, then the row-statistics exchange [4][128] f32
constexpr float kLnEps = 1e-5f;

template <bool PAIR>
struct Cfg {
  static constexpr int kCS = PAIR ? 2 : 1;
  static constexpr int kWTile = kBN * 128 / kCS;   // weight bytes per CTA and (N-tile, k-block)
  // A ring stage holds TWO k-blocks: the single MMA-issuing thread is close to critical (8 tcgen05.mma of 64
  // clocks each per hand-shake instead of 4), and the producer issues half as many waits.
  static constexpr int kWStage = 2 * kWTile;
  static constexpr int kWRing = PAIR ? 4 : 2;      // 64 KB either way
  // no alignment slack: the dynamic shared-memory window of a kernel without static shared memory starts
  // 1024-byte aligned (checked at kernel entry; SWIZZLE_128B tiles need it)
  static constexpr size_t kSmem = kABytes + kSlabBytes + kWRing * kWStage + kBarBytes;
  static_assert(kSmem <= 232448, "shared memory budget");
};

__device__ __forceinline__ void named_barrier_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_half2(uint32_t u) {
  return __half22float2(*reinterpret_cast<const __half2*>(&u));
}
__device__ __forceinline__ void ld_shared_v4u(uint32_t addr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr) : "memory");
}
// D[tmem of both CTAs] (+)= A[tmem, own 128 rows of each CTA] * B[smem halves]; issued by the leader.
__device__ __forceinline__ void umma_f16_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int V>
using ic = std::integral_constant<int, V>;

// TIMED: a second instantiation with phase clocks (tools/bench_encoder.py); the product kernel carries none of it
// (even disabled at run time the extra branches in the MMA-issuing thread cost 30 %).
template <bool PAIR, bool TIMED>
__global__ void __launch_bounds__(kThreads, 1)
encoder_block_kernel(const __grid_constant__ EncoderBlockParams p) {
  using C = Cfg<PAIR>;
  constexpr int CS = C::kCS;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = ptx::smem_u32(smem_raw);
  if (smem_base & 1023u) {
    if (threadIdx.x == 0) printf("fpb200: encoder_block: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  const uint32_t a_base = smem_base;                       // 8 k-block tiles [128 rows x 128 B], SW128
  const uint32_t slab_base = a_base + kABytes;             // 2 tiles
  const uint32_t w_base = slab_base + kSlabBytes;          // weight ring
  const uint32_t bar_base = w_base + C::kWRing * C::kWStage;
  // barriers (8 B each)
  auto afull = [&](int kb) { return bar_base + 8u * kb; };              // [8] attention k-block landed
  auto xready = [&](int kb) { return bar_base + 8u * (8 + kb); };       // [8] x1 k-block written to the A buffer
  auto hready = [&](int kb) { return bar_base + 8u * (16 + kb); };      // [8] h k-block written to tensor memory
  auto wfull = [&](int s) { return bar_base + 8u * (24 + s); };         // [8]
  auto wempty = [&](int s) { return bar_base + 8u * (32 + s); };        // [8]
  auto accfull = [&](int s) { return bar_base + 8u * (40 + s); };       // [2]
  auto accempty = [&](int s) { return bar_base + 8u * (42 + s); };      // [2]
  const uint32_t sfull = bar_base + 8u * 44, sempty = bar_base + 8u * 45, afree = bar_base + 8u * 46;
  const uint32_t tmem_slot = bar_base + 8u * 47;
  const uint32_t red_base = bar_base + 512;  // row-statistics exchange: [2 column halves][128 rows] sum, then sumsq
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int crank = PAIR ? (int)ptx::cluster_ctarank() : 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) {
      ptx::mbar_init(afull(i), CS);          // one expect_tx arrival per CTA's producer (PAIR: on the leader)
      ptx::mbar_init(xready(i), 4 * CS);     // the 4 warps (one per lane quarter) that own the k-block's column half
      ptx::mbar_init(hready(i), 4 * CS);
      ptx::mbar_init(wfull(i), CS);
      ptx::mbar_init(wempty(i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(accfull(i), 1);
      ptx::mbar_init(accempty(i), kEpiWarps * CS);
    }
    ptx::mbar_init(sfull, 1);
    ptx::mbar_init(sempty, kEpiWarps);
    ptx::mbar_init(afree, kEpiWarps);
    ptx::fence_barrier_init();
    ptx::prefetch_tensormap(&p.tm_att);
    ptx::prefetch_tensormap(&p.tm_x0);
    ptx::prefetch_tensormap(&p.tm_w[0]);
    ptx::prefetch_tensormap(&p.tm_w[1]);
    ptx::prefetch_tensormap(&p.tm_w[2]);
  }
  if (warp == 1) {
    if (PAIR) {
      ptx::tmem_alloc_pair(tmem_slot, 512);
      ptx::tmem_relinquish_pair();
    } else {
      ptx::tmem_alloc(tmem_slot, 512);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CS > 1) ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // programmatic dependent launch (plan.pdl): the prologue above may overlap the attention kernel's tail
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int n_blocks = (p.rows + kBM - 1) / kBM;
  const int n_items = (n_blocks + CS - 1) / CS;  // work items of a cluster
  const int cluster_id = blockIdx.x / CS, num_clusters = gridDim.x / CS;
  constexpr uint16_t kAllCtas = (uint16_t)((1u << CS) - 1u);
  // PAIR: barriers that gate the leader's MMA issue live in the leader CTA
  auto on_leader = [&](uint32_t bar) { return PAIR ? ptx::mapa_rank(bar, 0) : bar; };

  if (warp == 0) {
    // ================================ TMA producer ================================
    int ws = 0;
    uint32_t wph = 0;
    int it = 0;
    for (int item = cluster_id; item < n_items; item += num_clusters, ++it) {
      const int row0 = (item * CS + crank) * kBM;  // may lie beyond the tensor: TMA zero-fills, nothing is stored
      ptx::mbar_wait(afree, (it & 1) ^ 1u);
      if (ptx::elect_one()) {
        for (int kb = 0; kb < kKB; ++kb) {
          const uint32_t bar = on_leader(afull(kb));
          if (PAIR) {
            ptx::mbar_expect_tx_cluster(bar, kTileA);
            ptx::tma_load_2d_pair(a_base + kb * kTileA, &p.tm_att, bar, kb * 64, row0);
          } else {
            ptx::mbar_expect_tx(bar, kTileA);
            ptx::tma_load_2d(a_base + kb * kTileA, &p.tm_att, bar, kb * 64, row0);
          }
        }
      }
      __syncwarp();
      for (int g = 0; g < 3; ++g) {
        for (int nt = 0; nt < kNT; ++nt) {
          for (int kb = 0; kb < kKB; kb += 2) {
            ptx::mbar_wait(wempty(ws), wph ^ 1u);
            if (ptx::elect_one()) {
              const uint32_t bar = on_leader(wfull(ws));
              const uint32_t dst = w_base + ws * C::kWStage;
              const int nrow = nt * kBN + crank * (kBN / CS);
              if (PAIR) {
                ptx::mbar_expect_tx_cluster(bar, C::kWStage);
                ptx::tma_load_2d_pair(dst, &p.tm_w[g], bar, kb * 64, nrow);
                ptx::tma_load_2d_pair(dst + C::kWTile, &p.tm_w[g], bar, kb * 64 + 64, nrow);
              } else {
                ptx::mbar_expect_tx(bar, C::kWStage);
                ptx::tma_load_2d(dst, &p.tm_w[g], bar, kb * 64, nrow);
                ptx::tma_load_2d(dst + C::kWTile, &p.tm_w[g], bar, kb * 64 + 64, nrow);
              }
            }
            __syncwarp();
            if (++ws == C::kWRing) {
              ws = 0;
              wph ^= 1u;
            }
          }
          if (g == 0) {
            // residual slab of this N-tile (x0[:, nt*128 .. +128)); issued behind the tile's weights so that
            // waiting for the epilogue of the previous tile never delays the weight stream
            const int use = it * kNT + nt;
            ptx::mbar_wait(sempty, (use & 1) ^ 1u);
            if (ptx::elect_one()) {
              ptx::mbar_expect_tx(sfull, kSlabBytes);
              ptx::tma_load_2d(slab_base, &p.tm_x0, sfull, nt * kBN, row0);
              ptx::tma_load_2d(slab_base + kTileA, &p.tm_x0, sfull, nt * kBN + 64, row0);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (!PAIR || crank == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(PAIR ? 256 : kBM, kBN);
      constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO, version, SWIZZLE_128B
      auto desc = [&](uint32_t lo) { return (static_cast<uint64_t>(kDescHi) << 32) | lo; };
      const uint32_t a_lo0 = ((a_base & 0x3FFFF) >> 4) | (1u << 16);
      const uint32_t w_lo0 = ((w_base & 0x3FFFF) >> 4) | (1u << 16);
      int ws = 0;
      uint32_t wph = 0;
      int it = 0;
      // optional phase clocks (tools/bench_encoder.py): where the issuing thread of CTA 0 waits
      const bool timed = TIMED && p.dbg_clocks != nullptr && blockIdx.x == 0 && lane == 0;
      long long mk[6] = {0, 0, 0, 0, 0, 0};
      long long m_prev = timed ? clock64() : 0;
      auto mtick = [&](int k) {
        if constexpr (!TIMED) return;
        if (timed) {
          const long long now = clock64();
          mk[k] += now - m_prev;
          m_prev = now;
        }
      };
      for (int item = cluster_id; item < n_items; item += num_clusters, ++it) {
        const uint32_t bpar = it & 1;  // parity of the once-per-block barriers
        for (int g = 0; g < 3; ++g) {
          for (int nt = 0; nt < kNT; ++nt) {
            const int t = g * kNT + nt;              // accumulator tile of the block (0..11)
            const int acc = t & 1;
            const int use = it * 6 + (t >> 1);       // how often this stage has been used before
            mtick(5);
            ptx::mbar_wait(accempty(acc), (use & 1) ^ 1u);
            ptx::tc_fence_after();
            mtick(0);
            const uint32_t d_tmem = tmem_base + acc * kBN;
            for (int kb = 0; kb < kKB; kb += 2) {
              if (nt == 0) {  // first use of these two A k-blocks by this GEMM
                mtick(5);
                ptx::mbar_wait(g == 0 ? afull(kb) : (g == 1 ? xready(kb) : hready(kb)), bpar);
                ptx::mbar_wait(g == 0 ? afull(kb + 1) : (g == 1 ? xready(kb + 1) : hready(kb + 1)), bpar);
                mtick(1 + g);
              }
              mtick(5);
              ptx::mbar_wait(wfull(ws), wph);
              ptx::tc_fence_after();
              mtick(4);
              if (ptx::elect_one()) {
                const uint64_t bdesc = desc(w_lo0 + ws * (C::kWStage >> 4));
                if (g < 2) {
                  const uint64_t adesc = desc(a_lo0 + kb * (kTileA >> 4));
#pragma unroll
                  for (int k = 0; k < 8; ++k) {
                    // k-block (k >> 2): next A tile (kTileA bytes on) / next weight tile of the stage
                    const uint64_t ad = adesc + (uint64_t)((k >> 2) * (kTileA >> 4)) + 2u * (k & 3);
                    const uint64_t bd = bdesc + (uint64_t)((k >> 2) * (C::kWTile >> 4)) + 2u * (k & 3);
                    if (PAIR) ptx::umma_f16_pair(d_tmem, ad, bd, idesc, (uint32_t)(kb | k));
                    else ptx::umma_f16(d_tmem, ad, bd, idesc, (uint32_t)(kb | k));
                  }
                } else {
                  // A = h from tensor memory: 16 K values = 8 packed columns per MMA
                  const uint32_t a_tmem = tmem_base + kColR + kb * 32;
#pragma unroll
                  for (int k = 0; k < 8; ++k) {
                    const uint64_t bd = bdesc + (uint64_t)((k >> 2) * (C::kWTile >> 4)) + 2u * (k & 3);
                    if (PAIR) umma_f16_ts_pair(d_tmem, a_tmem + 8u * k, bd, idesc, (uint32_t)(kb | k));
                    else ptx::umma_f16_ts(d_tmem, a_tmem + 8u * k, bd, idesc, (uint32_t)(kb | k));
                  }
                }
                if (PAIR) ptx::umma_commit_pair_mcast(wempty(ws), kAllCtas);
                else ptx::umma_commit(wempty(ws));
                if (kb == kKB - 2) {
                  if (PAIR) ptx::umma_commit_pair_mcast(accfull(acc), kAllCtas);
                  else ptx::umma_commit(accfull(acc));
                }
              }
              __syncwarp();
              if (++ws == C::kWRing) {
                ws = 0;
                wph ^= 1u;
              }
            }
          }
        }
      }
      if (timed) {
        mtick(5);
        for (int k = 0; k < 6; ++k) p.dbg_clocks[k] = (unsigned long long)mk[k];
        p.dbg_clocks[6] = (unsigned long long)it;
      }
    }
  } else {
    // ================================ epilogue ================================
    // Everything below is instantiated twice, once per 64-column half (SIDX) a warp set owns; the tile / k-block
    // loops stay rolled (fully unrolled the kernel is 290 KB of SASS and thrashes the instruction cache: measured
    // 285 us against 208 us) but their counters are warp-uniform, so with SIDX a compile-time constant every column
    // offset is uniform and the bias / LayerNorm vectors in the kernel parameters are fetched with uniform
    // constant loads (LDCU.128: four values per instruction into uniform registers that the FADD / FFMA read
    // directly).  Per-thread indexed LDC loads were 47 % of the warp stalls; global loads before that were worse.
    auto run_epilogue = [&](auto sidx_tag) {
    constexpr int sidx = decltype(sidx_tag)::value;  // which 64-column half of every 128-column tile this warp owns
    const int q = warp & 3;      // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    float* red = reinterpret_cast<float*>(smem_raw + (red_base - ptx::smem_u32(smem_raw)));  // [2][128] sum, [2][128] sumsq
    // byte offset of 16-byte chunk c of this thread's row inside a SW128 tile
    auto chunk_off = [&](int c) { return static_cast<uint32_t>(row * 128 + ((c ^ (row & 7)) << 4)); };
    int it = 0;
    long long row_g = 0;
    float s1 = 0.f, s2 = 0.f;  // row statistics of the pre-LayerNorm sum (GEMM 1 and 3)
    const bool timed = TIMED && p.dbg_clocks != nullptr && blockIdx.x == 0 && threadIdx.x == 64;
    long long tk[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long t_prev = timed ? clock64() : 0;
    auto tick = [&](int k) {
      if constexpr (!TIMED) return;
      if (timed) {
        const long long now = clock64();
        tk[k] += now - t_prev;
        t_prev = now;
      }
    };

    // ---- pass A of one accumulator tile; the GEMM index is a compile-time constant so that the residual / ReLU
    // variants carry no predicated-off instructions.  Bias / LayerNorm vectors come from the constant bank
    // (kernel parameters): a warp-uniform LDC instead of a global load + descriptor set-up per pair.
    auto pass_a = [&](auto g_tag, int nt) {
      constexpr int G = decltype(g_tag)::value;
      const float* bias_g = G == 0 ? p.vec.b_out : (G == 1 ? p.vec.b1 : p.vec.b2);
      const int t = G * kNT + nt;
      const int acc = t & 1;
      const int use = it * 6 + (t >> 1);
      const int kb = 2 * nt + sidx;   // the k-block (64 columns) this warp produces
      const int c0 = kb * 64;         // its first column
      // residual of this thread's row, fetched before the accumulator is waited for
      uint32_t res[32];
      if constexpr (G == 0) {
        ptx::mbar_wait(sfull, (it * kNT + nt) & 1);
        tick(0);  // waiting for the residual slab
#pragma unroll
        for (int c = 0; c < 8; ++c)
          ld_shared_v4u(slab_base + sidx * kTileA + chunk_off(c), res[4 * c], res[4 * c + 1], res[4 * c + 2], res[4 * c + 3]);
      } else if constexpr (G == 2) {
#pragma unroll
        for (int c = 0; c < 8; ++c)
          ld_shared_v4u(a_base + kb * kTileA + chunk_off(c), res[4 * c], res[4 * c + 1], res[4 * c + 2], res[4 * c + 3]);
      }
      ptx::mbar_wait(accfull(acc), use & 1);
      ptx::tc_fence_after();
      tick(1 + G);  // waiting for the accumulator of GEMM G
      uint32_t r0[32], r1[32];
      ptx::tmem_ld_32x32b_x32(tlane + acc * kBN + sidx * 64, r0);
      ptx::tmem_ld_32x32b_x32(tlane + acc * kBN + sidx * 64 + 32, r1);
      ptx::tmem_ld_wait();
      // the accumulator stage is free again as soon as it sits in registers
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) ptx::mbar_arrive_cluster(ptx::mapa_rank(accempty(acc), 0));
        else ptx::mbar_arrive(accempty(acc));
      }
      uint32_t pk[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float v0 = __uint_as_float(j < 16 ? r0[2 * j] : r1[2 * j - 32]) + bias_g[c0 + 2 * j];
        float v1 = __uint_as_float(j < 16 ? r0[2 * j + 1] : r1[2 * j - 31]) + bias_g[c0 + 2 * j + 1];
        if constexpr (G != 1) {
          const float2 rr = unpack_half2(res[j]);
          v0 += rr.x;
          v1 += rr.y;
          s1 += v0 + v1;
          s2 = fmaf(v0, v0, fmaf(v1, v1, s2));
        } else {
          v0 = fmaxf(v0, 0.f);
          v1 = fmaxf(v1, 0.f);
        }
        pk[j] = pack_half2(v0, v1);
      }
      if constexpr (G == 0) {  // the slab has been consumed (its values went into pk[])
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(sempty);
      }
      if constexpr (G < 2) {
        // GEMM 1: pre-LayerNorm sum, GEMM 2: h -> R (packed fp16: K pair per column)
        uint32_t a16[16], b16[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          a16[j] = pk[j];
          b16[j] = pk[16 + j];
        }
        ptx::tmem_st_32x32b_x16(tlane + kColR + kb * 32, a16);
        ptx::tmem_st_32x32b_x16(tlane + kColR + kb * 32 + 16, b16);
        ptx::tmem_st_wait();
        if constexpr (G == 1) {
          if (p.dbg_h && row_g < p.rows) {
            uint4* dst = reinterpret_cast<uint4*>(p.dbg_h + row_g * kD + c0);
#pragma unroll
            for (int c = 0; c < 8; ++c) dst[c] = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
          }
          // publish h k-block kb to the MMA warp (A operand of GEMM 3 in tensor memory)
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR) ptx::mbar_arrive_cluster(ptx::mapa_rank(hready(kb), 0));
            else ptx::mbar_arrive(hready(kb));
          }
        }
      } else {
        // GEMM 3: pre-LayerNorm sum written in place over x1 (this thread's own row of its own k-block)
#pragma unroll
        for (int c = 0; c < 8; ++c)
          ptx::st_shared_v4(a_base + kb * kTileA + chunk_off(c), pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      }
      tick(4 + G);  // pass A of GEMM G
    };
    // ---- row statistics of a 512-wide row: combine the two column halves -> (scale, shift) of the normalisation
    auto row_affine = [&](float& a, float& b) {
      red[sidx * kBM + row] = s1;
      red[(2 + sidx) * kBM + row] = s2;
      named_barrier_sync(1, 32 * kEpiWarps);
      const float t1 = s1 + red[(sidx ^ 1) * kBM + row];
      const float t2 = s2 + red[(2 + (sidx ^ 1)) * kBM + row];
      named_barrier_sync(1, 32 * kEpiWarps);  // red[] may be rewritten
      const float mean = t1 * (1.f / kD);
      a = rsqrtf(fmaxf(t2 * (1.f / kD) - mean * mean, 0.f) + kLnEps);
      b = -mean * a;  // (v - mean) * rstd = v * a + b
      s1 = s2 = 0.f;
      tick(7);  // statistics exchange
    };

    for (int item = cluster_id; item < n_items; item += num_clusters, ++it) {
      row_g = static_cast<long long>(item * CS + crank) * kBM + row;  // this thread's token row
      float na, nb;
      // ================= GEMM 1: out-projection + residual, then LayerNorm 1 -> x1 in the A buffer
#pragma unroll 1
      for (int nt = 0; nt < kNT; ++nt) pass_a(ic<0>{}, nt);
      row_affine(na, nb);
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        const int kb = 2 * j + sidx, c0 = kb * 64;
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(tlane + kColR + kb * 32, r);
        ptx::tmem_ld_wait();
        uint32_t pk[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const float2 v = unpack_half2(r[e]);
          pk[e] = pack_half2(fmaf(fmaf(v.x, na, nb), p.vec.ln1_g[c0 + 2 * e], p.vec.ln1_b[c0 + 2 * e]),
                             fmaf(fmaf(v.y, na, nb), p.vec.ln1_g[c0 + 2 * e + 1], p.vec.ln1_b[c0 + 2 * e + 1]));
        }
#pragma unroll
        for (int c = 0; c < 8; ++c)
          ptx::st_shared_v4(a_base + kb * kTileA + chunk_off(c), pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        if (p.dbg_x1 && row_g < p.rows) {
          uint4* dst = reinterpret_cast<uint4*>(p.dbg_x1 + row_g * kD + c0);
#pragma unroll
          for (int c = 0; c < 8; ++c) dst[c] = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        }
        ptx::fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's operand reads
        ptx::tc_fence_before();    // the R columns just read will be overwritten by GEMM 2's pass A
        __syncwarp();
        if (lane == 0) {
          if (PAIR) ptx::mbar_arrive_cluster(ptx::mapa_rank(xready(kb), 0));
          else ptx::mbar_arrive(xready(kb));
        }
      }
      tick(8);  // LayerNorm 1 pass
      // ================= GEMM 2: FFN first layer + ReLU -> h in tensor memory
#pragma unroll 1
      for (int nt = 0; nt < kNT; ++nt) pass_a(ic<1>{}, nt);
      // ================= GEMM 3: FFN second layer + residual, then LayerNorm 2 + partial token sums
#pragma unroll 1
      for (int nt = 0; nt < kNT; ++nt) pass_a(ic<2>{}, nt);
      row_affine(na, nb);
      {
        // groups of 16 consecutive rows (= half a warp)
        const int hl = lane & 15;
        const long long grp = row_g >> 4;  // global group index of this thread's row
        const bool grp_ok = (row_g - hl + kGroup) <= p.rows;  // rows is a multiple of 16
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
          const int kb = 2 * j + sidx, c0 = kb * 64;
          uint32_t r[32];
#pragma unroll
          for (int c = 0; c < 8; ++c)
            ld_shared_v4u(a_base + kb * kTileA + chunk_off(c), r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]);
          float y[64];
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float2 v = unpack_half2(r[e]);
            y[2 * e] = fmaf(fmaf(v.x, na, nb), p.vec.ln2_g[c0 + 2 * e], p.vec.ln2_b[c0 + 2 * e]);
            y[2 * e + 1] = fmaf(fmaf(v.y, na, nb), p.vec.ln2_g[c0 + 2 * e + 1], p.vec.ln2_b[c0 + 2 * e + 1]);
          }
          if (p.dbg_y && row_g < p.rows) {
            uint4* dst = reinterpret_cast<uint4*>(p.dbg_y + row_g * kD + c0);
#pragma unroll
            for (int c = 0; c < 8; ++c)
              dst[c] = make_uint4(pack_half2(y[8 * c], y[8 * c + 1]), pack_half2(y[8 * c + 2], y[8 * c + 3]),
                                  pack_half2(y[8 * c + 4], y[8 * c + 5]), pack_half2(y[8 * c + 6], y[8 * c + 7]));
          }
          // recursive halving over the 16 lanes of the group: at the level of lane bit b a lane keeps the half of
          // its columns selected by that bit, summed with its partner's copy (fixed tree: deterministic)
#pragma unroll
          for (int lvl = 0; lvl < 4; ++lvl) {
            const int bit = 8 >> lvl;           // lane bit of this level
            const int half = 32 >> lvl;         // columns kept
            const bool up = (lane & bit) != 0;
#pragma unroll
            for (int e = 0; e < half; ++e) {
              const float keep = up ? y[e + half] : y[e];
              const float send = up ? y[e] : y[e + half];
              y[e] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
            }
          }
          if (grp_ok) {
            // kept half at level 0 is [32 b3, +32), at level 1 [16 b2, +16) inside it, ...: first column = 32 b3 + 16 b2 + 8 b1 + 4 b0
            const int col = ((hl & 8) ? 32 : 0) + ((hl & 4) ? 16 : 0) + ((hl & 2) ? 8 : 0) + ((hl & 1) ? 4 : 0);
            *reinterpret_cast<float4*>(p.partial + grp * kD + c0 + col) = make_float4(y[0], y[1], y[2], y[3]);
          }
        }
        // the A buffer may be refilled with the next block's attention rows
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(afree);
        tick(9);  // LayerNorm 2 pass + partial sums
      }
    }
    if (timed) {
      for (int k = 0; k < 12; ++k) p.dbg_clocks[8 + k] = (unsigned long long)tk[k];
      p.dbg_clocks[20] = (unsigned long long)it;
    }
    };  // run_epilogue
    if (((warp - 2) >> 2) == 0) run_epilogue(ic<0>{});
    else run_epilogue(ic<1>{});
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (CS > 1) ptx::cluster_sync();
  if (warp == 1) {
    ptx::tc_fence_after();
    if (PAIR) ptx::tmem_dealloc_pair(tmem_base, 512);
    else ptx::tmem_dealloc(tmem_base, 512);
  }
}

// out[h, c] = (sum of the 25 group sums of hypothesis h) / 400, groups added in index order.
__global__ void __launch_bounds__(256)
encoder_mean_kernel(const float* __restrict__ partial, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * (kD / 4)) return;
  const int h = i / (kD / 4), c4 = i - h * (kD / 4);
  const float4* src = reinterpret_cast<const float4*>(partial) + (long long)h * 25 * (kD / 4) + c4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 5
  for (int g = 0; g < 25; ++g) {
    const float4 v = src[(long long)g * (kD / 4)];
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  reinterpret_cast<float4*>(out)[i] = make_float4(acc.x * (1.f / 400), acc.y * (1.f / 400), acc.z * (1.f / 400), acc.w * (1.f / 400));
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
int g_num_sms = 0;
std::mutex g_mu;

const char* encode_2d(CUtensorMap* tm, const void* base, long long inner, long long outer, int box_inner, int box_outer,
                      CUtensorMapL2promotion promo) {
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)(inner * 2)};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t es[2] = {1, 1};
  const CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? nullptr : "encoder_block: cuTensorMapEncodeTiled failed";
}

}  // namespace

const char* encoder_block_init() {
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
  cudaError_t e = cudaSuccess;
  auto set_attr = [&](auto kernel, size_t smem) {
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  };
  set_attr(encoder_block_kernel<false, false>, Cfg<false>::kSmem);
  set_attr(encoder_block_kernel<true, false>, Cfg<true>::kSmem);
  set_attr(encoder_block_kernel<false, true>, Cfg<false>::kSmem);
  set_attr(encoder_block_kernel<true, true>, Cfg<true>::kSmem);
  if (e == cudaSuccess) e = preload_kernels(encoder_mean_kernel);
  if (e != cudaSuccess)
    return "cudaFuncSetAttribute(encoder_block smem) failed (not an sm_100a device?)";
  cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  if (dev < 64) inited |= 1ull << dev;
  return nullptr;
}

const char* build_encoder_block(const EncoderBlockDesc& d, EncoderBlockPlan* plan) {
  if (const char* e = encoder_block_init()) return e;
  if (d.n_hyp < 1) return "encoder_block: no hypotheses";
  if (!d.att || !d.x0 || !d.w_out || !d.w1 || !d.w2 || !d.partial || !d.mean_out) return "encoder_block: null pointer";
  EncoderBlockParams& p = plan->p;
  memset(&p, 0, sizeof(p));
  const long long rows = (long long)d.n_hyp * 400;
  const int n_blocks = (int)((rows + kBM - 1) / kBM);
  // CTA pairs (half the weight traffic per SM): measured faster at every size (N = 1: 34.9 vs 38.6 us, 32: 39.6 vs
  // 42.7, 63: 67.6 vs 74.5, 252: 197 vs 210), so single CTAs remain only as mode 1 for tests / A-B runs
  const bool pair = d.mode == 0 ? n_blocks >= 2 : d.mode == 2;
  const int cs = pair ? 2 : 1;
  if (const char* e = encode_2d(&p.tm_att, d.att, kD, rows, 64, kBM, CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) return e;
  if (const char* e = encode_2d(&p.tm_x0, d.x0, kD, rows, 64, kBM, CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) return e;
  const __half* w[3] = {d.w_out, d.w1, d.w2};
  for (int g = 0; g < 3; ++g)
    if (const char* e = encode_2d(&p.tm_w[g], w[g], kD, kD, 64, kBN / cs, CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) return e;
  {
    // bias / LayerNorm vectors travel BY VALUE in the kernel parameters (constant bank): 14 KB copied once per plan
    const float* src[7] = {d.b_out, d.b1, d.b2, d.ln1_g, d.ln1_b, d.ln2_g, d.ln2_b};
    float* dst[7] = {p.vec.b_out, p.vec.b1, p.vec.b2, p.vec.ln1_g, p.vec.ln1_b, p.vec.ln2_g, p.vec.ln2_b};
    for (int i = 0; i < 7; ++i) {
      if (!src[i]) return "encoder_block: null bias / LayerNorm vector";
      if (d.vectors_on_host)
        memcpy(dst[i], src[i], sizeof(float) * kD);
      else if (cudaMemcpy(dst[i], src[i], sizeof(float) * kD, cudaMemcpyDeviceToHost) != cudaSuccess)
        return "encoder_block: reading the bias / LayerNorm vectors failed";
    }
  }
  p.partial = d.partial;
  p.rows = (int)rows;
  p.dbg_x1 = d.dbg_x1; p.dbg_h = d.dbg_h; p.dbg_y = d.dbg_y;
  p.dbg_clocks = d.dbg_clocks;
  plan->pair = pair;
  {
    const char* pdl = getenv("FPB200_PDL");  // see conv_gemm.cu: only launches that leave SMs idle
    plan->pdl = pdl ? pdl[0] == '1' : n_blocks < g_num_sms;
  }
  const int items = (n_blocks + cs - 1) / cs;
  const int cap = g_num_sms / cs;
  plan->grid = (items < cap ? items : cap) * cs;
  plan->smem = pair ? Cfg<true>::kSmem : Cfg<false>::kSmem;
  plan->n_hyp = d.n_hyp;
  plan->mean_out = d.mean_out;
  plan->flops = 3.0 * 2.0 * (double)rows * kD * kD;
  return nullptr;
}

cudaError_t launch_encoder_block(const EncoderBlockPlan& plan, cudaStream_t stream) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(plan.grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = plan.smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = plan.pair ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = plan.pdl ? 2 : 1;
  cudaError_t e;
  if (plan.p.dbg_clocks)
    e = plan.pair ? cudaLaunchKernelEx(&cfg, encoder_block_kernel<true, true>, plan.p)
                  : cudaLaunchKernelEx(&cfg, encoder_block_kernel<false, true>, plan.p);
  else
    e = plan.pair ? cudaLaunchKernelEx(&cfg, encoder_block_kernel<true, false>, plan.p)
                  : cudaLaunchKernelEx(&cfg, encoder_block_kernel<false, false>, plan.p);
  if (e != cudaSuccess) return e;
  const int n4 = plan.n_hyp * (kD / 4);
  encoder_mean_kernel<<<(n4 + 255) / 256, 256, 0, stream>>>(plan.p.partial, plan.mean_out, plan.n_hyp);
  return cudaGetLastError();
}

}  // namespace fpb
