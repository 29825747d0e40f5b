// tcgen05 implicit-GEMM convolution kernel (see conv_gemm.h).
//
// Roles inside one persistent CTA (64 + 128 kEpiSets = 320 threads, one CTA per SM):
//   warp 0         : TMA producer.  Two rings: activation BLOCKS (one per tap group x 64 input
//                    channels: bh + span tile rows, loaded once and shared by the group's taps) and
//                    weight stages (one block_n-row tile per tap).
//   warp 1         : MMA issuer.  Four 128 x block_n x 16 tcgen05.mma per tap, the activation operand
//                    descriptor shifted by the tap's row offset inside the block, accumulating in one
//                    of two TMEM accumulator stages.
//   warps 2..9     : epilogue.  tcgen05.ld 32 columns at a time, transpose through shared memory so
//                    that global accesses are row-contiguous, bias / residual / ReLU / pos-table.
// The two TMEM stages let the epilogue of tile i overlap the MMAs of tile i+1.
#include "conv_gemm.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "ptx.cuh"

namespace fpb {

namespace {

// Epilogue warp sets (4 warps each, one per TMEM lane quarter).  Measured with 3 sets (448 threads, 128
// registers): no gain on any layer shape - the short-K linear layers are not paced by the number of
// epilogue instruction streams but by shared-memory bandwidth (per 128 x 256 tile and K = 512: 256 KB of
// TMA writes + 256 KB of MMA operand reads + 256 KB of fp32 staging traffic = 6 k clocks at 128 B/clock
// against 4 k clocks of MMA).
constexpr int kEpiSets = 2;
constexpr int kEpiWarps = 4 * kEpiSets;
constexpr int kThreads = 64 + 32 * kEpiWarps;  // TMA warp + MMA warp + epilogue warps
constexpr int kStageBytesPerWarp = 32 * 32 * 4;         // epilogue staging: 32 rows x 32 fp32 columns
constexpr int kMaxRing = 8;
constexpr int kBarrierBytes = 1024;  // mbarriers, TMEM slot, the issuer's K-loop schedule
constexpr int kSchedMax = 96;        // taps x 64-channel blocks of one tile (per-tap kernels)
constexpr int kSmemBudget = 232448 - 1024 - kEpiWarps * kStageBytesPerWarp - kBarrierBytes;  // for the two rings

// PAIR = false: independent CTAs.  PAIR = true: CTA pairs issuing ONE tcgen05.mma.cta_group::2 of
// M = 256: each SM holds (and ingests) its 128 activation rows and only HALF of the weight tile.
template <int BN, bool PAIR>
struct Cfg {
  static constexpr int kCS = PAIR ? 2 : 1;
  static constexpr int kBTile = (BN * kGemmBlockK * 2) / kCS;  // weight bytes per CTA and tap
  // Accumulator stages in TMEM.  Wide tiles: 2 stages, the 8 epilogue warps split the columns of one
  // tile.  Narrow tiles (short MMA time per tile): 4 stages, the two sets of 4 epilogue warps take
  // alternate tiles, so that the fixed latency of an epilogue pass is overlapped across tiles.
  static constexpr bool kAltTiles = BN <= 128;
  static constexpr int kAccStages = kAltTiles ? 4 : 2;
  static constexpr int kTmemCols = kAccStages * BN;
};

template <int BN, bool PAIR, bool PER_TAP>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
  using C = Cfg<BN, PAIR>;
  constexpr int CS = C::kCS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;                        // [na] activation blocks
  const uint32_t b_base = a_base + p.na * p.a_bytes;        // [nb] weight stages
  const uint32_t stage_base = b_base + p.nb * p.b_stage_bytes;
  const uint32_t bar_base = stage_base + kEpiWarps * kStageBytesPerWarp;
  // barrier layout (8 B each): full[8], empty[8], tfull[4], tempty[4], tmem ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxRing + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kMaxRing + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kMaxRing + 4 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxRing + 8);
  const uint32_t sched_base = bar_base + 256;  // kSchedMax x u32
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
      smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxRing; ++s) {
      // PAIR: the leader's full barriers take one expect_tx arrival from each CTA's producer; the
      // empty barriers one multicast tcgen05.commit by the leader
      ptx::mbar_init(full_bar(s), PAIR ? 2 : 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < C::kAccStages; ++s) {
      ptx::mbar_init(tfull_bar(s), 1);
      // one arrival per epilogue warp working on the tile (PAIR: of both CTAs, on the leader's barrier)
      ptx::mbar_init(tempty_bar(s), (C::kAltTiles ? 4 : kEpiWarps) * CS);
    }
    ptx::fence_barrier_init();
    ptx::prefetch_tensormap(&p.tm_a);
    ptx::prefetch_tensormap(&p.tm_b);
  }
  if (warp == 1) {
    if (PAIR) {
      ptx::tmem_alloc_pair(tmem_slot, C::kTmemCols);
      ptx::tmem_relinquish_pair();
    } else {
      ptx::tmem_alloc(tmem_slot, C::kTmemCols);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CS > 1) ptx::cluster_sync();  // the peer's barriers are initialised before anything lands on them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // Programmatic dependent launch (plan.pdl): everything above (barriers, tensor-memory allocation, descriptor prefetch)
  // touched only this CTA's own state and may run while the producer of this layer's input is still finishing on other
  // SMs; from here on its output is read.  Both instructions are no-ops in an ordinary launch.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int m_tiles = p.tiles_x * p.tiles_y * p.tiles_img;
  const int m_groups = (m_tiles + CS - 1) / CS;
  const int total_tiles = m_groups * p.n_tiles;  // work items of a cluster
  const int crank = CS > 1 ? (int)ptx::cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / CS;
  const int num_clusters = gridDim.x / CS;
  constexpr uint16_t kAllCtas = (uint16_t)((1u << CS) - 1u);

  const int items = p.cblocks * p.num_groups;        // K-loop items of a tile
  if (warp == 0) {
    int as = 0, bs = 0;
    uint32_t bph = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
      const int nt = tile % p.n_tiles;
      const int mt = (tile / p.n_tiles) * CS + crank;  // may be >= m_tiles: TMA zero-fills, epilogue masks
      const int x0 = (mt % p.tiles_x) * p.bw;
      const int y0 = ((mt / p.tiles_x) % p.tiles_y) * p.bh;
      const int n0 = (mt / (p.tiles_x * p.tiles_y)) * p.bn;
      const int nrow = nt * BN + crank * (BN / CS);
      int cb = 0, g = 0;
      for (int j = 0; j < items; ++j) {
        const int ntap = p.g_ntaps[g];
        const uint32_t kidx = p.g_kidx[g];
        const int ax = x0 + p.g_dx[g], ay = y0 + p.g_dy[g], ap = p.g_plane[g];
        const uint32_t a_dst = a_base + as * p.a_bytes;
        as = (as + 1 == p.na) ? 0 : as + 1;
        const int nstage = PER_TAP ? ntap : 1;
        for (int i = 0; i < nstage; ++i) {
          ptx::mbar_wait(empty_bar(bs), bph ^ 1u);
          if (ptx::elect_one()) {
            // PAIR: both CTAs report their bytes to (and land their loads on) the LEADER's barrier
            const uint32_t bar = PAIR ? ptx::mapa_rank(full_bar(bs), 0) : full_bar(bs);
            const uint32_t bytes = (i == 0 ? p.a_bytes : 0) + (PER_TAP ? 1 : ntap) * C::kBTile;
            if (PAIR) ptx::mbar_expect_tx_cluster(bar, bytes);
            else ptx::mbar_expect_tx(bar, bytes);
            if (i == 0) {
              if (PAIR) ptx::tma_load_5d_pair(a_dst, &p.tm_a, bar, cb * kGemmBlockK, ax, n0, ay, ap);
              else ptx::tma_load_5d(a_dst, &p.tm_a, bar, cb * kGemmBlockK, ax, n0, ay, ap);
            }
            const uint32_t b_dst = b_base + bs * p.b_stage_bytes;
            const int t0 = PER_TAP ? i : 0, t1 = PER_TAP ? i + 1 : ntap;
            for (int t = t0; t < t1; ++t) {
              const int kcol = (((kidx >> (8 * t)) & 0xff) * p.cblocks + cb) * kGemmBlockK;
              const uint32_t dst = b_dst + (PER_TAP ? 0 : t) * C::kBTile;
              if (PAIR) ptx::tma_load_2d_pair(dst, &p.tm_b, bar, kcol, nrow);
              else ptx::tma_load_2d(dst, &p.tm_b, bar, kcol, nrow);
            }
          }
          __syncwarp();
          if (++bs == p.nb) {
            bs = 0;
            bph ^= 1u;
          }
        }
        if (++g == p.num_groups) {
          g = 0;
          if (++cb == p.cblocks) cb = 0;
        }
      }
    }
  } else if (warp == 1) {
    if (!PAIR || crank == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(PAIR ? 256 : kGemmBlockM, BN);
      constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO, version, SWIZZLE_128B
      const uint32_t row_units = p.row_bytes >> 4;  // descriptor address units are 16 bytes
      const uint32_t a_units = p.a_bytes >> 4, b_units = p.b_stage_bytes >> 4;
      const uint32_t a_lo0 = ((a_base & 0x3FFFF) >> 4) | (1u << 16);
      const uint32_t b_lo0 = ((b_base & 0x3FFFF) >> 4) | (1u << 16);
      auto desc = [&](uint32_t lo) { return (static_cast<uint64_t>(kDescHi) << 32) | lo; };
      int as = 0, bs = 0;
      uint32_t bph = 0;
      int it = 0;
      if constexpr (PER_TAP) {
        // K-loop schedule of a tile, one entry per tap: descriptor offset of the tap inside its
        // activation block | bit 16: first tap of a group (next activation slot)
        volatile uint32_t* sched = reinterpret_cast<volatile uint32_t*>(
            smem_raw + (sched_base - ptx::smem_u32(smem_raw)));
        const int taps_cb = [&] {
          int n = 0;
          for (int g = 0; g < p.num_groups; ++g) n += p.g_ntaps[g];
          return n;
        }();
        if (lane == 0) {
          int n = 0;
          for (int g = 0; g < p.num_groups; ++g)
            for (int i = 0; i < p.g_ntaps[g]; ++i)
              sched[n++] = (((p.g_aoff[g] >> (8 * i)) & 0xff) * row_units) | (i == 0 ? 1u << 16 : 0u);
        }
        __syncwarp();
        const int n_entries = taps_cb * p.cblocks;
        uint32_t a_cur = a_lo0;
        for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, ++it) {
          const int acc = it % C::kAccStages;
          const uint32_t acc_phase = (it / C::kAccStages) & 1;
          ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * BN;
          uint32_t first = 0;  // 0 until the first MMA of the tile has been issued
          int si = 0;
          uint32_t e = sched[0];
          for (int idx = 0; idx < n_entries; ++idx) {
            if (++si == taps_cb) si = 0;
            const uint32_t e_next = sched[si];
            if (e & (1u << 16)) {
              a_cur = a_lo0 + as * a_units;
              as = (as + 1 == p.na) ? 0 : as + 1;
            }
            const uint64_t adesc = desc(a_cur + (e & 0xffffu));
            const uint64_t bdesc = desc(b_lo0 + bs * b_units);
            ptx::mbar_wait(full_bar(bs), bph);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
#pragma unroll
              for (int k = 0; k < kGemmBlockK / 16; ++k) {
                // advancing 16 fp16 along K inside the 128-byte swizzle row = +32 B = +2 in the
                // descriptor's 16-byte address units
                if (PAIR) ptx::umma_f16_pair(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, first | (uint32_t)k);
                else ptx::umma_f16(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, first | (uint32_t)k);
              }
              if (PAIR) ptx::umma_commit_pair_mcast(empty_bar(bs), kAllCtas);
              else ptx::umma_commit(empty_bar(bs));
              if (idx == n_entries - 1) {
                if (PAIR) ptx::umma_commit_pair_mcast(tfull_bar(acc), kAllCtas);
                else ptx::umma_commit(tfull_bar(acc));
              }
            }
            __syncwarp();
            first = 1;
            e = e_next;
            if (++bs == p.nb) {
              bs = 0;
              bph ^= 1u;
            }
          }
        }
      } else {
        for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, ++it) {
          const int acc = it % C::kAccStages;
          const uint32_t acc_phase = (it / C::kAccStages) & 1;
          ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * BN;
          uint32_t first = 0;
          int g = 0;
          for (int j = 0; j < items; ++j) {
            const int ntap = p.g_ntaps[g];
            const uint32_t aoff = p.g_aoff[g];
            const uint32_t a_cur = a_lo0 + as * a_units;
            as = (as + 1 == p.na) ? 0 : as + 1;
            const uint32_t b_cur = b_lo0 + bs * b_units;
            ptx::mbar_wait(full_bar(bs), bph);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
#pragma unroll
              for (int i = 0; i < kMaxGroupTaps; ++i) {
                if (i < ntap) {
                  const uint64_t adesc = desc(a_cur + ((aoff >> (8 * i)) & 0xff) * row_units);
                  const uint64_t bdesc = desc(b_cur + i * (C::kBTile >> 4));
#pragma unroll
                  for (int k = 0; k < kGemmBlockK / 16; ++k) {
                    const uint32_t accum = first | (uint32_t)(i | k);
                    if (PAIR) ptx::umma_f16_pair(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, accum);
                    else ptx::umma_f16(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, accum);
                  }
                }
              }
              if (PAIR) ptx::umma_commit_pair_mcast(empty_bar(bs), kAllCtas);
              else ptx::umma_commit(empty_bar(bs));
              if (j == items - 1) {
                if (PAIR) ptx::umma_commit_pair_mcast(tfull_bar(acc), kAllCtas);
                else ptx::umma_commit(tfull_bar(acc));
              }
            }
            __syncwarp();
            first = 1;
            if (++bs == p.nb) {
              bs = 0;
              bph ^= 1u;
            }
            if (++g == p.num_groups) g = 0;
          }
        }
      }
    }
  } else {
    // Epilogue, 8 warps.  TMEM hands every thread one accumulator ROW (tcgen05.ld 32x32b; warp w may
    // only read lanes 32*(w%4)..), but global memory wants lanes to cover contiguous bytes of a row.
    // So each warp stages 32 columns of its 32 rows in shared memory (fp32, XOR-swizzled 16-byte
    // chunks: conflict-free both ways) and re-reads them with 4 lanes per row: every global
    // load/store instruction then touches 8 rows x 64 contiguous bytes.  The two warps that share a
    // TMEM lane quarter take alternate 32-column chunks.
    const int ew = warp - 2;       // 0..kEpiWarps-1
    const int q = warp & 3;        // TMEM lane quarter this warp may read
    const int chunk_par = ew >> 2; // warp set: which of the interleaved chunks (wide tiles) / tiles (narrow)
    const uint32_t stg = stage_base + static_cast<uint32_t>(ew) * kStageBytesPerWarp;
    const int sub = lane >> 2;  // row within a group of 8
    const int seg = lane & 3;   // 8-column segment of the 32-column chunk
    // staging addresses (bytes): write = own row, read = rows i*8+sub
    const uint32_t wr_base = stg + lane * 128;
    const uint32_t rd_sw = static_cast<uint32_t>(sub & 7);
    constexpr int kItStep = C::kAltTiles ? kEpiSets : 1;           // narrow tiles: the sets take tiles in turn
    constexpr int kColStep = C::kAltTiles ? 32 : 32 * kEpiSets;    // ... and all of their columns
    for (int it = C::kAltTiles ? chunk_par : 0, tile = cluster_id + it * num_clusters; tile < total_tiles;
         tile += kItStep * num_clusters, it += kItStep) {
      const int nt = tile % p.n_tiles;
      const int mt = (tile / p.n_tiles) * CS + crank;
      const int x0 = (mt % p.tiles_x) << p.bw_log2;
      const int y0 = ((mt / p.tiles_x) % p.tiles_y) * p.bh;
      const int n0 = (mt / (p.tiles_x * p.tiles_y)) << p.bn_log2;
      const int acc = it % C::kAccStages;
      const uint32_t acc_phase = (it / C::kAccStages) & 1;
      // per-lane bookkeeping for the 4 row groups (offsets in units of 8 halves = 16 bytes)
      uint32_t out_off[4], res_off[4], pp_off[4];
      uint32_t vmask = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = q * 32 + i * 8 + sub;  // GEMM row = (y * bn + n) * bw + x
        const int x = x0 + (row & (p.bw - 1));
        const int n = n0 + ((row >> p.bw_log2) & (p.bn - 1));
        const int y = y0 + (row >> (p.bw_log2 + p.bn_log2));
        if ((x < p.W) && (y < p.H) && (n < p.n_img)) vmask |= 1u << i;
        const int nn = n % p.img_mod;
        const int coff = (n / p.img_mod) * p.half_coff + nt * BN;
        long long opix;
        if (p.out_phase_split) {
          const int plane = (y & 1) * 2 + (x & 1);
          opix = ((static_cast<long long>(nn) * 4 + plane) * (p.H >> 1) + (y >> 1)) * (p.W >> 1) + (x >> 1);
        } else {
          opix = (static_cast<long long>(nn) * p.H + y) * p.W + x;
        }
        out_off[i] = static_cast<uint32_t>((opix * p.out_ld + coff) >> 3) + seg;
        res_off[i] = static_cast<uint32_t>((((static_cast<long long>(n) * p.H + y) * p.W + x) * p.res_ld + nt * BN) >> 3) + seg;
        pp_off[i] = static_cast<uint32_t>((y * p.W + x) * p.cout + nt * BN) + seg * 8;
      }
      const float* bptr = p.bias + nt * BN + seg * 8;
      uint4* out16 = reinterpret_cast<uint4*>(p.out);
      const uint4* res16 = reinterpret_cast<const uint4*>(p.residual);
      const int c_first = C::kAltTiles ? 0 : chunk_par * 32;

      // The residual rows of a chunk are fetched one chunk ahead (the first one before the wait on
      // the accumulator), so their DRAM/L2 latency overlaps the TMEM read and the staging.
      uint4 res_cur[4];
      auto fetch_residual = [&](uint4 (&dst)[4], int c0) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          dst[i] = (res16 && ((vmask >> i) & 1u)) ? __ldg(res16 + res_off[i] + (c0 >> 3)) : make_uint4(0, 0, 0, 0);
      };
      fetch_residual(res_cur, c_first);

      ptx::mbar_wait(tfull_bar(acc), acc_phase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
      // The TMEM read of chunk c+1 is issued as soon as chunk c has been staged, so its latency
      // overlaps the arithmetic and stores of chunk c.
      uint32_t r[32];
      ptx::tmem_ld_32x32b_x32(taddr + c_first, r);
      if (!res16 && !p.post_add) {
        // No residual sum and no positional table (the QKV / FFN1 linear layers and half of the convolutions):
        // bias and ReLU are applied in the accumulator's own layout (a thread owns a row, the 32 bias values are
        // broadcast loads) and the chunk is staged as fp16 - half the shared-memory traffic of the fp32
        // transpose below, which is what paces the short-K layers.  Same fp32 operations, same rounding.
        const float* brow = p.bias + nt * BN;
        const uint32_t wr16 = stg + lane * 64;
        const uint32_t wr_sw = static_cast<uint32_t>((lane >> 1) & 3);
#pragma unroll 1
        for (int c0 = c_first; c0 < BN; c0 += kColStep) {
          float4 bb[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) bb[j] = __ldg(reinterpret_cast<const float4*>(brow + c0) + j);
          ptx::tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float v0 = __uint_as_float(r[4 * j + 0]) + bb[j].x, v1 = __uint_as_float(r[4 * j + 1]) + bb[j].y;
            float v2 = __uint_as_float(r[4 * j + 2]) + bb[j].z, v3 = __uint_as_float(r[4 * j + 3]) + bb[j].w;
            if (p.relu) {
              v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f);
            }
            const __half2 h01 = __floats2half2_rn(v0, v1), h23 = __floats2half2_rn(v2, v3);
            pk[2 * j] = *reinterpret_cast<const uint32_t*>(&h01);
            pk[2 * j + 1] = *reinterpret_cast<const uint32_t*>(&h23);
          }
          if (c0 + kColStep < BN) ptx::tmem_ld_32x32b_x32(taddr + c0 + kColStep, r);
          // 64-byte rows, 16-byte units XOR-swizzled by (row / 2) % 4: conflict-free for the row-per-lane
          // writes and for the 4-lanes-per-row reads
#pragma unroll
          for (int c = 0; c < 4; ++c)
            ptx::st_shared_v4(wr16 + ((c ^ wr_sw) << 4), pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int row = i * 8 + sub;
            float f0, f1, f2, f3;
            ptx::ld_shared_v4(stg + row * 64 + ((seg ^ ((row >> 1) & 3)) << 4), f0, f1, f2, f3);
            if ((vmask >> i) & 1u)
              out16[out_off[i] + (c0 >> 3)] = make_uint4(__float_as_uint(f0), __float_as_uint(f1), __float_as_uint(f2),
                                                         __float_as_uint(f3));
          }
          __syncwarp();
        }
      } else
#pragma unroll 1
      for (int c0 = c_first; c0 < BN; c0 += kColStep) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bptr + c0));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(bptr + c0 + 4));
        ptx::tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 8; ++c)
          ptx::st_shared_v4(wr_base + ((c ^ (lane & 7)) << 4), r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]);
        __syncwarp();
        uint4 res_nxt[4];
        if (c0 + kColStep < BN) {
          ptx::tmem_ld_32x32b_x32(taddr + c0 + kColStep, r);
          fetch_residual(res_nxt, c0 + kColStep);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t rd = stg + (i * 8 + sub) * 128;
          float v[8];
          ptx::ld_shared_v4(rd + (((2 * seg) ^ rd_sw) << 4), v[0], v[1], v[2], v[3]);
          ptx::ld_shared_v4(rd + (((2 * seg + 1) ^ rd_sw) << 4), v[4], v[5], v[6], v[7]);
          v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
          v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
          if (res16) {
            const __half2* h2 = reinterpret_cast<const __half2*>(&res_cur[i]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __half22float2(h2[e]);
              v[2 * e] += f.x;
              v[2 * e + 1] += f.y;
            }
          }
          if (p.relu) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
          }
          if (p.post_add && ((vmask >> i) & 1u)) {
            const float4 a0 = __ldg(reinterpret_cast<const float4*>(p.post_add + pp_off[i] + c0));
            const float4 a1 = __ldg(reinterpret_cast<const float4*>(p.post_add + pp_off[i] + c0 + 4));
            v[0] += a0.x; v[1] += a0.y; v[2] += a0.z; v[3] += a0.w;
            v[4] += a1.x; v[5] += a1.y; v[6] += a1.z; v[7] += a1.w;
          }
          uint4 u;
          __half2* h2 = reinterpret_cast<__half2*>(&u);
#pragma unroll
          for (int e = 0; e < 4; ++e) h2[e] = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
          if ((vmask >> i) & 1u) out16[out_off[i] + (c0 >> 3)] = u;
        }
        __syncwarp();
        if (c0 + kColStep < BN) {
#pragma unroll
          for (int i = 0; i < 4; ++i) res_cur[i] = res_nxt[i];
        }
      }
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) ptx::mbar_arrive_cluster(ptx::mapa_rank(tempty_bar(acc), 0));
        else ptx::mbar_arrive(tempty_bar(acc));
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (CS > 1) ptx::cluster_sync();  // no CTA exits while a peer may still signal its barriers
  if (warp == 1) {
    ptx::tc_fence_after();
    if (PAIR) ptx::tmem_dealloc_pair(tmem_base, C::kTmemCols);
    else ptx::tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
int g_num_sms = 0;

}  // namespace

const char* conv_gemm_init() {
  // The driver entry point is process-wide; kernel attributes are per device (a process may hold sessions on
  // several GPUs: BASELINE config 5 in one process), so they are set once for every device that gets here.
  static std::mutex mu;
  static unsigned long long inited = 0;  // bit d: device d done
  std::lock_guard<std::mutex> lk(mu);
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return "cudaGetDevice failed";
  if (dev < 64 && ((inited >> dev) & 1ull)) return nullptr;
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
      return "cuTensorMapEncodeTiled not available from the CUDA driver";
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  cudaError_t e = cudaSuccess;
  auto set_attr = [&](auto kernel) {
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  };
  set_attr(conv_gemm_kernel<64, false, false>);
  set_attr(conv_gemm_kernel<64, false, true>);
  set_attr(conv_gemm_kernel<64, true, false>);
  set_attr(conv_gemm_kernel<64, true, true>);
  set_attr(conv_gemm_kernel<128, false, false>);
  set_attr(conv_gemm_kernel<128, false, true>);
  set_attr(conv_gemm_kernel<128, true, false>);
  set_attr(conv_gemm_kernel<128, true, true>);
  set_attr(conv_gemm_kernel<256, false, true>);
  set_attr(conv_gemm_kernel<256, true, true>);
  if (e != cudaSuccess) return "cudaFuncSetAttribute(max dynamic smem) failed (not an sm_100a device?)";
  // every B200 has the same SM count; the value is only a tiling heuristic (grid size, tile width)
  cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  if (dev < 64) inited |= 1ull << dev;
  return nullptr;
}

void taps_3x3_s1(ConvGemmDesc* d) {
  d->num_taps = 9;
  for (int ky = 0; ky < 3; ++ky)
    for (int kx = 0; kx < 3; ++kx) {
      d->tap_dy[ky * 3 + kx] = (int8_t)(ky - 1);
      d->tap_dx[ky * 3 + kx] = (int8_t)(kx - 1);
      d->tap_plane[ky * 3 + kx] = 0;
    }
}

void taps_3x3_s2_phase(ConvGemmDesc* d) {
  // input pixel (2*oy + ky - 1, 2*ox + kx - 1): parity = (k-1)&1, half-res offset = -1 for k==0
  d->num_taps = 9;
  for (int ky = 0; ky < 3; ++ky)
    for (int kx = 0; kx < 3; ++kx) {
      const int py = (ky == 1) ? 0 : 1, px = (kx == 1) ? 0 : 1;
      d->tap_dy[ky * 3 + kx] = (int8_t)(ky == 0 ? -1 : 0);
      d->tap_dx[ky * 3 + kx] = (int8_t)(kx == 0 ? -1 : 0);
      d->tap_plane[ky * 3 + kx] = (int8_t)(py * 2 + px);
    }
}

void taps_stem7_rows(ConvGemmDesc* d) {
  // input row 2*oy + ky - 3: parity r = (ky-3) mod 2, half-row offset q = floor((ky-3)/2)
  d->num_taps = 7;
  for (int ky = 0; ky < 7; ++ky) {
    const int o = ky - 3;
    const int r = ((o % 2) + 2) % 2;
    const int q = (o - r) / 2;
    d->tap_dy[ky] = (int8_t)q;
    d->tap_dx[ky] = 0;
    d->tap_plane[ky] = (int8_t)r;
  }
}

const char* build_conv_gemm(const ConvGemmDesc& d, ConvGemmPlan* plan) {
  if (const char* e = conv_gemm_init()) return e;
  if (d.bw * d.bh * d.bn != kGemmBlockM) return "conv_gemm: bw*bh*bn must be 128";
  if ((d.bw & (d.bw - 1)) || (d.bh & (d.bh - 1))) return "conv_gemm: bw and bh must be powers of two";
  if (d.cin % kGemmBlockK != 0) return "conv_gemm: cin must be a multiple of 64";
  if (d.num_taps < 1 || d.num_taps > kMaxTaps) return "conv_gemm: bad tap count";
  if (d.cout % 64 != 0) return "conv_gemm: cout must be a multiple of 64";
  // Tile width: the widest of 256 / 128 / 64 that divides cout (fewest activation re-reads), unless the
  // launch then has fewer tiles than SMs (Track: N = 1 hypothesis, 4..100 row tiles per layer): a small
  // launch is bound by the serial K loop of its few CTAs, so narrower tiles (4x more CTAs, 4x shorter
  // MMAs) cut its latency.  Accumulation order per output element is the same for every width.
  int bn_tile = 64;
  {
    const long long m_tiles = (long long)((d.out_w + d.bw - 1) / d.bw) * ((d.out_h + d.bh - 1) / d.bh) *
                              ((d.n_img + d.bn - 1) / d.bn);
    for (int cand : {256, 128})
      if (d.cout % cand == 0 && m_tiles * (d.cout / cand) >= g_num_sms) {
        bn_tile = cand;
        break;
      }
  }
  if (d.block_n) {
    if ((d.block_n != 64 && d.block_n != 128 && d.block_n != 256) || d.cout % d.block_n)
      return "conv_gemm: block_n must be 64, 128 or 256 and divide cout";
    bn_tile = d.block_n;
  } else if (const char* env = getenv("FPB_GEMM_BN")) {
    const int v = atoi(env);
    if ((v == 64 || v == 128 || v == 256) && d.cout % v == 0) bn_tile = v;
  }
  const int out_ld = d.out_ld ? d.out_ld : d.cout;
  if (out_ld % 8 || d.half_coff % 8) return "conv_gemm: output leading dims must be multiples of 8";
  if (d.out_phase_split && ((d.out_h | d.out_w) & 1)) return "conv_gemm: phase-split needs even H, W";

  // pairing mode (see Cfg): 0 = single CTAs, 2 = CTA pairs issuing cta_group::2 MMAs
  const long long m_tiles_all = (long long)((d.out_w + d.bw - 1) / d.bw) * ((d.out_h + d.bh - 1) / d.bh) *
                                ((d.n_img + d.bn - 1) / d.bn);
  int mode = d.cluster > 0 ? (d.cluster == 1 ? 0 : 2) : (m_tiles_all >= 2 * g_num_sms ? 2 : 0);
  // Multi-tap (deep-K) layers between one and a few waves of tiles - a rank's share of a sharded Register, 31..126
  // hypotheses - are dominated by wave quantisation: e.g. the 512-channel layers at 32 hypotheses have 100 row tiles,
  // i.e. 200 tiles of width 256 on 148 SMs = 2 waves for 1.35 waves of work.  Every (width, pairing) candidate is
  // costed as waves x tile width x a measured per-column penalty (B200, tools/scratch sweep of every layer shape at
  // 31 / 32 / 63 / 126 hypotheses: narrow tiles on single CTAs pay 1.5-1.6x per column for the repeated activation
  // traffic, CTA pairs do not) and the rule above is overridden when a candidate is predicted > 5 % faster.
  // Measured: 512 -> 512 at 32 hypotheses 55 -> 43 us (128 wide, paired), 256 -> 512 / s2 at 63: 52 -> 44 us.
  if (d.num_taps > 1 && !d.block_n && d.cluster == 0 && !getenv("FPB_GEMM_BN") && !getenv("FPB_GEMM_MODE") &&
      m_tiles_all * (d.cout / 64) >= 2LL * g_num_sms) {
    auto cost = [&](int bn, int md) -> double {
      const long long items = (md ? (m_tiles_all + 1) / 2 : m_tiles_all) * (d.cout / bn);
      const long long slots = md ? g_num_sms / 2 : g_num_sms;
      const long long waves = (items + slots - 1) / slots;
      double pen = bn == 256 ? (md ? 0.94 : 1.0) : bn == 128 ? (md ? 1.0 : 1.5) : (md ? 1.4 : 1.6);
      if (d.planes > 1 && bn < 256) pen *= 1.3;  // phase-split (stride-2) inputs: more, smaller activation boxes
      return (double)waves * bn * pen;
    };
    double best = cost(bn_tile, mode ? 1 : 0);
    const double keep = best;
    int best_bn = bn_tile, best_mode = mode;
    for (int bn : {256, 128, 64}) {
      if (d.cout % bn) continue;
      for (int md : {1, 0}) {
        if (md && m_tiles_all < 2) continue;
        const double c = cost(bn, md);
        if (c < best) { best = c; best_bn = bn; best_mode = md ? 2 : 0; }
      }
    }
    if (best < 0.95 * keep) { bn_tile = best_bn; mode = best_mode; }
  }
  // Launches with fewer tiles than SMs (Track: 1 - 2 hypotheses) are bound by what ONE SM can ingest during its serial K
  // loop (~1 MB of activation blocks + weights per tile of a 512-channel layer): in a CTA pair each SM ingests only half
  // of every weight tile.  Measured: Track 2 756 -> 2 870 fps.
  if (d.cluster == 0 && !d.block_n && m_tiles_all >= 2 && m_tiles_all * (d.cout / bn_tile) <= g_num_sms) mode = 2;
  if (d.cluster == 0)
    if (const char* env = getenv("FPB_GEMM_MODE")) mode = atoi(env) ? 2 : 0;
  const int cluster = mode == 0 ? 1 : 2;
  if (d.bn & (d.bn - 1)) return "conv_gemm: bn must be a power of two";
  if ((d.bw * d.bn) % 8) return "conv_gemm: a tile row (bw*bn pixels) must be a whole number of swizzle atoms";
  ConvGemmParams& p = plan->p;
  memset(&p, 0, sizeof(p));
  // ---- tap groups: taps with the same (dx, plane) share one activation block
  int span = 0;
  {
    int num_groups = 0;
    int g_dx[kMaxGroups], g_plane[kMaxGroups], g_dy[kMaxGroups], dy_max[kMaxGroups], g_ntaps[kMaxGroups] = {0};
    int g_tap[kMaxGroups][kMaxGroupTaps];
    for (int t = 0; t < d.num_taps; ++t) {
      int g = 0;
      for (; g < num_groups; ++g)
        if (g_dx[g] == d.tap_dx[t] && g_plane[g] == d.tap_plane[t]) break;
      if (g == num_groups) {
        if (g == kMaxGroups) return "conv_gemm: too many tap groups";
        num_groups++;
        g_dx[g] = d.tap_dx[t];
        g_plane[g] = d.tap_plane[t];
        g_dy[g] = dy_max[g] = d.tap_dy[t];
      }
      if (g_ntaps[g] == kMaxGroupTaps) return "conv_gemm: too many taps in a group";
      g_tap[g][g_ntaps[g]++] = t;
      if (d.tap_dy[t] < g_dy[g]) g_dy[g] = d.tap_dy[t];
      if (d.tap_dy[t] > dy_max[g]) dy_max[g] = d.tap_dy[t];
    }
    int max_taps = 1;
    for (int g = 0; g < num_groups; ++g) {
      if (dy_max[g] - g_dy[g] > span) span = dy_max[g] - g_dy[g];
      if (g_ntaps[g] > max_taps) max_taps = g_ntaps[g];
    }
    if (d.bh + span > 256) return "conv_gemm: activation block too tall";
    p.row_bytes = d.bn * d.bw * 128;
    p.a_bytes = (d.bh + span) * p.row_bytes;
    const int b_tile = bn_tile * kGemmBlockK * 2 / cluster;
    // one weight stage per tap when a tap alone covers the producer/issuer hand-shake (>= 16 KB of
    // weights = 4 MMAs of N = 256), else one stage per group
    const bool per_tap = b_tile >= 16384 || max_taps == 1;
    p.b_stage_bytes = (per_tap ? 1 : max_taps) * b_tile;
    int uses[kMaxGroups];
    for (int g = 0; g < num_groups; ++g) uses[g] = per_tap ? g_ntaps[g] : 1;
    // ring depths: most weight stages such that an activation slot is provably free when reused
    p.na = p.nb = 0;
    for (int na = 2; na <= kMaxRing; ++na) {
      const long long left = (long long)kSmemBudget - (long long)na * p.a_bytes;
      if (left < 2LL * p.b_stage_bytes) break;
      int nb = (int)(left / p.b_stage_bytes);
      int cond = 1 << 30;  // weight stages in any na-1 consecutive groups
      for (int g0 = 0; g0 < num_groups; ++g0) {
        int sum = 0;
        for (int j = 0; j < na - 1; ++j) sum += uses[(g0 + j) % num_groups];
        if (sum < cond) cond = sum;
      }
      if (cond < nb) nb = cond;
      if (nb > kMaxRing) nb = kMaxRing;
      if (nb > p.nb) {
        p.nb = nb;
        p.na = na;
      }
    }
    if (p.nb < 2) return "conv_gemm: tile does not fit in shared memory";
    p.num_groups = num_groups;
    p.per_tap = per_tap ? 1 : 0;
    if (per_tap && d.num_taps * (d.cin / kGemmBlockK) > kSchedMax) return "conv_gemm: K loop too long";
    for (int g = 0; g < num_groups; ++g) {
      p.g_dx[g] = (int8_t)g_dx[g];
      p.g_dy[g] = (int8_t)g_dy[g];
      p.g_plane[g] = (int8_t)g_plane[g];
      p.g_ntaps[g] = (int8_t)g_ntaps[g];
      for (int i = 0; i < g_ntaps[g]; ++i) {
        const int t = g_tap[g][i];
        p.g_aoff[g] |= (uint32_t)(d.tap_dy[t] - g_dy[g]) << (8 * i);
        p.g_kidx[g] |= (uint32_t)t << (8 * i);
      }
    }
  }
  {
    const long long sw = d.stride_w ? d.stride_w : 2LL * d.cin;
    const long long sh = d.stride_h ? d.stride_h : sw * d.in_w;
    const long long sp = d.stride_plane ? d.stride_plane : sh * d.in_h;
    const long long si = d.stride_img ? d.stride_img : sp * d.planes;
    // dimension order (C, W, image, H, plane): the box lands in shared memory as [h][image][w] rows
    cuuint64_t dims[5] = {(cuuint64_t)d.cin, (cuuint64_t)d.in_w, (cuuint64_t)d.n_img, (cuuint64_t)d.in_h,
                          (cuuint64_t)d.planes};
    cuuint64_t strides[4] = {(cuuint64_t)sw, (cuuint64_t)si, (cuuint64_t)sh, (cuuint64_t)sp};
    cuuint32_t box[5] = {kGemmBlockK, (cuuint32_t)d.bw, (cuuint32_t)d.bn, (cuuint32_t)(d.bh + span), 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = g_encode(&p.tm_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, (void*)d.in, dims, strides,
                          box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      static char msg[128];
      snprintf(msg, sizeof(msg), "conv_gemm: cuTensorMapEncodeTiled(A) failed with CUresult %d", (int)r);
      return msg;
    }
  }
  {
    const long long ktot = (long long)d.num_taps * d.cin;
    cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)d.cout};
    cuuint64_t strides[1] = {(cuuint64_t)(ktot * 2)};
    cuuint32_t box[2] = {kGemmBlockK, (cuuint32_t)(bn_tile / cluster)};
    cuuint32_t es[2] = {1, 1};
    CUresult r = g_encode(&p.tm_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)d.weight, dims,
                          strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      static char msg[128];
      snprintf(msg, sizeof(msg), "conv_gemm: cuTensorMapEncodeTiled(B) failed with CUresult %d", (int)r);
      return msg;
    }
  }
  p.bw = d.bw;
  p.bh = d.bh;
  p.bn = d.bn;
  for (p.bw_log2 = 0; (1 << p.bw_log2) < d.bw; ++p.bw_log2) {}
  for (p.bn_log2 = 0; (1 << p.bn_log2) < d.bn; ++p.bn_log2) {}
  p.W = d.out_w;
  p.H = d.out_h;
  p.n_img = d.n_img;
  p.tiles_x = (d.out_w + d.bw - 1) / d.bw;
  p.tiles_y = (d.out_h + d.bh - 1) / d.bh;
  p.tiles_img = (d.n_img + d.bn - 1) / d.bn;
  p.n_tiles = d.cout / bn_tile;
  p.cblocks = d.cin / kGemmBlockK;
  p.bias = d.bias;
  p.residual = d.residual;
  p.res_ld = d.res_ld ? d.res_ld : d.cout;
  p.post_add = d.post_add;
  p.out = d.out;
  p.relu = d.relu ? 1 : 0;
  p.cout = d.cout;
  p.out_phase_split = d.out_phase_split ? 1 : 0;
  p.out_ld = out_ld;
  p.img_mod = d.img_mod ? d.img_mod : d.n_img;
  p.half_coff = d.half_coff;

  plan->block_n = bn_tile;
  plan->cluster = cluster;
  plan->mode = mode;
  const long long m_tiles = (long long)p.tiles_x * p.tiles_y * p.tiles_img;
  const long long work = ((m_tiles + cluster - 1) / cluster) * p.n_tiles;  // per cluster
  const int cap = (d.max_ctas > 0 ? d.max_ctas : g_num_sms) / cluster;
  plan->grid = (int)(work < cap ? work : cap) * cluster;
  plan->smem = 1024 + (size_t)p.na * p.a_bytes + (size_t)p.nb * p.b_stage_bytes + kEpiWarps * kStageBytesPerWarp + kBarrierBytes;
  plan->flops = 2.0 * d.n_img * d.out_h * d.out_w * (double)d.cout * d.num_taps * d.cin;
  {
    // Programmatic dependent launch where the launch leaves SMs idle (Track, a few hypotheses): the next kernel's CTAs
    // start there early and hide their launch latency and prologue (Track: +4 %).  Full-size launches gain nothing
    // (measured at N = 252) and stay ordinary.  FPB200_PDL=0 / 1: never / always (measurements).
    const char* env = getenv("FPB200_PDL");
    plan->pdl = env ? env[0] == '1' : plan->grid < g_num_sms;
  }
  return nullptr;
}

namespace {
template <int BN, bool PAIR, bool PER_TAP>
cudaError_t launch_one(const ConvGemmPlan& plan, cudaStream_t stream) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(plan.grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = plan.smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = Cfg<BN, PAIR>::kCS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = plan.pdl ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, conv_gemm_kernel<BN, PAIR, PER_TAP>, plan.p);
}
template <int BN>
cudaError_t launch_bn(const ConvGemmPlan& plan, cudaStream_t stream) {
  const bool pair = plan.mode == 2;
  if constexpr (BN == 256) {
    return pair ? launch_one<BN, true, true>(plan, stream) : launch_one<BN, false, true>(plan, stream);
  } else {
    if (plan.p.per_tap) return pair ? launch_one<BN, true, true>(plan, stream) : launch_one<BN, false, true>(plan, stream);
    return pair ? launch_one<BN, true, false>(plan, stream) : launch_one<BN, false, false>(plan, stream);
  }
}
}  // namespace

cudaError_t launch_conv_gemm(const ConvGemmPlan& plan, cudaStream_t stream) {
  if (plan.grid <= 0) return cudaSuccess;
  switch (plan.block_n) {
    case 256: return launch_bn<256>(plan, stream);
    case 128: return launch_bn<128>(plan, stream);
    default: return launch_bn<64>(plan, stream);
  }
}

}  // namespace fpb
