// tcgen05 implicit-GEMM convolution kernel (see conv_gemm.h).
//
// Roles inside one persistent CTA (320 threads, one CTA per SM):
//   warp 0, lane 0 : TMA producer.  Per k-block (one filter tap x 64 input channels) it loads a
//                    128-row activation box and a block_n-row weight box into a smem ring.
//   warp 1, lane 0 : MMA issuer.  Four 128 x block_n x 16 tcgen05.mma per k-block, accumulating in
//                    one of two TMEM accumulator stages.
//   warps 2..9     : epilogue.  tcgen05.ld 32 columns at a time, transpose through shared memory so
//                    that global accesses are row-contiguous, bias / residual / ReLU / pos-table.
// The two TMEM stages let the epilogue of tile i overlap the MMAs of tile i+1.
#include "conv_gemm.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "ptx.cuh"

namespace fpb {

namespace {

constexpr int kThreads = 320;  // TMA warp + MMA warp + 8 epilogue warps
constexpr int kABytes = kGemmBlockM * kGemmBlockK * 2;  // 16 KB
constexpr int kStageBytesPerWarp = 32 * 32 * 4;         // epilogue staging: 32 rows x 32 fp32 columns

// MODE 0: independent CTAs.  MODE 1: CTA pairs, each CTA loads half of the weight tile and
// multicasts it to both (halves the L2 reads of B; each SM still ingests A + B).  MODE 2: CTA pairs
// issuing ONE tcgen05.mma.cta_group::2 of M = 256: each SM holds (and ingests) its 128 activation
// rows and only HALF of the weight tile, which is what lifts the shared-memory fill-rate bound of the
// narrower tiles.
template <int BN, int MODE>
struct Cfg {
  static constexpr int kCS = MODE == 0 ? 1 : 2;
  static constexpr int kBBytes = (BN * kGemmBlockK * 2) / (MODE == 2 ? 2 : 1);  // per CTA and k-block
  // k-blocks per pipeline stage: the producer / MMA warps synchronise once per stage, and that
  // hand-shake (try_wait + fence + elect + commit, ~500 cycles of warp time) has to be covered by the
  // tensor-core time of the stage: 4 MMAs of N=256 cover it, 4 MMAs of N<=128 do not.
  static constexpr int kKS = BN == 256 ? 1 : 2;
  static constexpr int kStageBytes = kKS * (kABytes + kBBytes);
  static constexpr int kStages = (192 * 1024) / kStageBytes > 10 ? 10 : (192 * 1024) / kStageBytes;
  static constexpr int kTmemCols = 2 * BN;
  static constexpr size_t kSmem = 1024 /*alignment slack*/ + size_t(kStages) * kStageBytes +
                                  8 * kStageBytesPerWarp /*epilogue staging*/ + 256 /*barriers*/;
};

template <int BN, int MODE>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
  using C = Cfg<BN, MODE>;
  constexpr int CS = C::kCS;
  constexpr bool kPairMma = MODE == 2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr int KS = C::kKS;
  const uint32_t a_base = smem_base;                                // [stage][sub] activation tiles
  const uint32_t b_base = smem_base + C::kStages * KS * kABytes;    // [stage][sub] weight tiles
  const uint32_t stage_base = b_base + C::kStages * KS * C::kBBytes;
  const uint32_t bar_base = stage_base + 8 * kStageBytesPerWarp;
  // barrier layout (8 B each): full[kStages], empty[kStages], tfull[2], tempty[2], then tmem ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * C::kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * C::kStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * C::kStages + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
      smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      // MODE 2: the leader's full barrier takes one expect_tx arrival from each CTA's producer
      ptx::mbar_init(full_bar(s), kPairMma ? 2 : 1);
      // MODE 1: one tcgen05.commit arrival from every CTA; MODE 2: one multicast commit by the leader
      ptx::mbar_init(empty_bar(s), MODE == 1 ? 2 : 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(tfull_bar(s), 1);
      // one arrival per epilogue warp (MODE 2: of both CTAs, on the leader's barrier)
      ptx::mbar_init(tempty_bar(s), kPairMma ? 16 : 8);
    }
    ptx::fence_barrier_init();
    ptx::prefetch_tensormap(&p.tm_a);
    ptx::prefetch_tensormap(&p.tm_b);
  }
  if (warp == 1) {
    if (kPairMma) {
      ptx::tmem_alloc_pair(tmem_slot, C::kTmemCols);
      ptx::tmem_relinquish_pair();
    } else {
      ptx::tmem_alloc(tmem_slot, C::kTmemCols);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CS > 1) ptx::cluster_sync();  // peers' barriers are initialised before anything is multicast
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int m_tiles = p.tiles_x * p.tiles_y * p.tiles_img;
  const int m_groups = (m_tiles + CS - 1) / CS;
  const int total_tiles = m_groups * p.n_tiles;  // work items of a cluster
  const int kblocks = p.num_taps * p.cblocks;
  const int crank = CS > 1 ? (int)ptx::cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / CS;
  const int num_clusters = gridDim.x / CS;
  constexpr uint16_t kAllCtas = (uint16_t)((1u << CS) - 1u);
  constexpr int n_stages = C::kStages;

  if (warp == 0) {
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
        const int nt = tile % p.n_tiles;
        const int mt = (tile / p.n_tiles) * CS + crank;  // may be >= m_tiles: TMA zero-fills, epilogue masks
        const int x0 = (mt % p.tiles_x) * p.bw;
        const int y0 = ((mt / p.tiles_x) % p.tiles_y) * p.bh;
        const int n0 = (mt / (p.tiles_x * p.tiles_y)) * p.bn;
        for (int kb0 = 0; kb0 < kblocks; kb0 += KS) {
          const int nsub = min(KS, kblocks - kb0);
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          if (ptx::elect_one()) {
            const uint32_t bytes = nsub * (kABytes + C::kBBytes);
            // MODE 2: both CTAs report their bytes to (and land their loads on) the LEADER's barrier
            const uint32_t bar = kPairMma ? ptx::mapa_rank(full_bar(stage), 0) : full_bar(stage);
            if (kPairMma) ptx::mbar_expect_tx_cluster(bar, bytes);
            else ptx::mbar_expect_tx(bar, bytes);
            for (int sub = 0; sub < nsub; ++sub) {
              const int kb = kb0 + sub;
              const int t = kb / p.cblocks, cb = kb - t * p.cblocks;
              const int cx = x0 + p.tap_dx[t], cy = y0 + p.tap_dy[t], cp = p.tap_plane[t];
              const uint32_t a_dst = a_base + (stage * KS + sub) * kABytes;
              const uint32_t b_dst = b_base + (stage * KS + sub) * C::kBBytes;
              if (kPairMma) {
                ptx::tma_load_5d_pair(a_dst, &p.tm_a, bar, cb * kGemmBlockK, cx, cy, cp, n0);
                ptx::tma_load_2d_pair(b_dst, &p.tm_b, bar, kb * kGemmBlockK, nt * BN + crank * (BN / 2));
              } else {
                ptx::tma_load_5d(a_dst, &p.tm_a, bar, cb * kGemmBlockK, cx, cy, cp, n0);
                if (MODE == 0) ptx::tma_load_2d(b_dst, &p.tm_b, bar, kb * kGemmBlockK, nt * BN);
                else ptx::tma_load_2d_mcast(b_dst + crank * (C::kBBytes / 2), &p.tm_b, bar, kb * kGemmBlockK,
                                            nt * BN + crank * (BN / 2), kAllCtas);
              }
            }
          }
          __syncwarp();
          if (++stage == n_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (!kPairMma || crank == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(kPairMma ? 256 : kGemmBlockM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb0 = 0; kb0 < kblocks; kb0 += KS) {
          const int nsub = min(KS, kblocks - kb0);
          ptx::mbar_wait(full_bar(stage), phase);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            for (int sub = 0; sub < nsub; ++sub) {
              const uint64_t adesc = ptx::make_sw128_kmajor_desc(a_base + (stage * KS + sub) * kABytes);
              const uint64_t bdesc = ptx::make_sw128_kmajor_desc(b_base + (stage * KS + sub) * C::kBBytes);
#pragma unroll
              for (int k = 0; k < kGemmBlockK / 16; ++k) {
                // advancing 16 fp16 along K inside the 128-byte swizzle row = +32 B = +2 in the
                // descriptor's 16-byte address units
                const uint32_t accum = (kb0 | sub | k) != 0;
                if (kPairMma) ptx::umma_f16_pair(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, accum);
                else ptx::umma_f16(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, accum);
              }
            }
            if (MODE == 0) ptx::umma_commit(empty_bar(stage));
            else if (MODE == 1) ptx::umma_commit_mcast(empty_bar(stage), kAllCtas);
            else ptx::umma_commit_pair_mcast(empty_bar(stage), kAllCtas);
            if (kb0 + KS >= kblocks) {
              if (kPairMma) ptx::umma_commit_pair_mcast(tfull_bar(acc), kAllCtas);
              else ptx::umma_commit(tfull_bar(acc));
            }
          }
          __syncwarp();
          if (++stage == n_stages) {
            stage = 0;
            phase ^= 1u;
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
    const int ew = warp - 2;       // 0..7
    const int q = warp & 3;        // TMEM lane quarter this warp may read
    const int chunk_par = ew >> 2; // which of the alternating chunks
    const uint32_t stg = stage_base + static_cast<uint32_t>(ew) * kStageBytesPerWarp;
    const int sub = lane >> 2;  // row within a group of 8
    const int seg = lane & 3;   // 8-column segment of the 32-column chunk
    // staging addresses (bytes): write = own row, read = rows i*8+sub
    const uint32_t wr_base = stg + lane * 128;
    const uint32_t rd_sw = static_cast<uint32_t>(sub & 7);
    int it = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, ++it) {
      const int nt = tile % p.n_tiles;
      const int mt = (tile / p.n_tiles) * CS + crank;
      const int x0 = (mt % p.tiles_x) << p.bw_log2;
      const int y0 = ((mt / p.tiles_x) % p.tiles_y) << p.bh_log2;
      const int n0 = (mt / (p.tiles_x * p.tiles_y)) * p.bn;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      // per-lane bookkeeping for the 4 row groups (offsets in units of 8 halves = 16 bytes)
      uint32_t out_off[4], res_off[4], pp_off[4];
      uint32_t vmask = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = q * 32 + i * 8 + sub;
        const int x = x0 + (row & (p.bw - 1));
        const int y = y0 + ((row >> p.bw_log2) & (p.bh - 1));
        const int n = n0 + (row >> (p.bw_log2 + p.bh_log2));
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
      const int c_first = chunk_par * 32;

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
#pragma unroll 1
      for (int c0 = c_first; c0 < BN; c0 += 64) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bptr + c0));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(bptr + c0 + 4));
        ptx::tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 8; ++c)
          ptx::st_shared_v4(wr_base + ((c ^ (lane & 7)) << 4), r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]);
        __syncwarp();
        uint4 res_nxt[4];
        if (c0 + 64 < BN) {
          ptx::tmem_ld_32x32b_x32(taddr + c0 + 64, r);
          fetch_residual(res_nxt, c0 + 64);
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
        if (c0 + 64 < BN) {
#pragma unroll
          for (int i = 0; i < 4; ++i) res_cur[i] = res_nxt[i];
        }
      }
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kPairMma) ptx::mbar_arrive_cluster(ptx::mapa_rank(tempty_bar(acc), 0));
        else ptx::mbar_arrive(tempty_bar(acc));
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (CS > 1) ptx::cluster_sync();  // no CTA exits while a peer may still signal its barriers
  if (warp == 1) {
    ptx::tc_fence_after();
    if (kPairMma) ptx::tmem_dealloc_pair(tmem_base, C::kTmemCols);
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
  if (g_encode) return nullptr;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
    return "cuTensorMapEncodeTiled not available from the CUDA driver";
  e = cudaSuccess;
  auto set_attr = [&](auto kernel, size_t smem) {
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  };
  set_attr(conv_gemm_kernel<64, 0>, Cfg<64, 0>::kSmem);
  set_attr(conv_gemm_kernel<128, 0>, Cfg<128, 0>::kSmem);
  set_attr(conv_gemm_kernel<256, 0>, Cfg<256, 0>::kSmem);
  set_attr(conv_gemm_kernel<64, 1>, Cfg<64, 1>::kSmem);
  set_attr(conv_gemm_kernel<128, 1>, Cfg<128, 1>::kSmem);
  set_attr(conv_gemm_kernel<256, 1>, Cfg<256, 1>::kSmem);
  set_attr(conv_gemm_kernel<64, 2>, Cfg<64, 2>::kSmem);
  set_attr(conv_gemm_kernel<128, 2>, Cfg<128, 2>::kSmem);
  set_attr(conv_gemm_kernel<256, 2>, Cfg<256, 2>::kSmem);
  if (e != cudaSuccess) return "cudaFuncSetAttribute(max dynamic smem) failed (not an sm_100a device?)";
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
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
  int bn_tile = 0;
  if (d.cout % 256 == 0) bn_tile = 256;
  else if (d.cout % 128 == 0) bn_tile = 128;
  else if (d.cout % 64 == 0) bn_tile = 64;
  else return "conv_gemm: cout must be a multiple of 64";
  const int out_ld = d.out_ld ? d.out_ld : d.cout;
  if (out_ld % 8 || d.half_coff % 8) return "conv_gemm: output leading dims must be multiples of 8";
  if (d.out_phase_split && ((d.out_h | d.out_w) & 1)) return "conv_gemm: phase-split needs even H, W";

  // pairing mode (see Cfg): 0 = single CTAs, 1 = pair + weight multicast, 2 = pair + cta_group::2 MMA
  const long long m_tiles_all = (long long)((d.out_w + d.bw - 1) / d.bw) * ((d.out_h + d.bh - 1) / d.bh) *
                                ((d.n_img + d.bn - 1) / d.bn);
  int mode = d.cluster > 0 ? d.cluster - 1 : (m_tiles_all >= 2 * g_num_sms ? 2 : 0);
  if (const char* env = getenv("FPB_GEMM_MODE")) mode = atoi(env);
  if (mode < 0 || mode > 2) return "conv_gemm: pairing mode must be 0, 1 or 2";
  const int cluster = mode == 0 ? 1 : 2;
  ConvGemmParams& p = plan->p;
  memset(&p, 0, sizeof(p));
  {
    const long long sw = d.stride_w ? d.stride_w : 2LL * d.cin;
    const long long sh = d.stride_h ? d.stride_h : sw * d.in_w;
    const long long sp = d.stride_plane ? d.stride_plane : sh * d.in_h;
    const long long si = d.stride_img ? d.stride_img : sp * d.planes;
    cuuint64_t dims[5] = {(cuuint64_t)d.cin, (cuuint64_t)d.in_w, (cuuint64_t)d.in_h,
                          (cuuint64_t)d.planes, (cuuint64_t)d.n_img};
    cuuint64_t strides[4] = {(cuuint64_t)sw, (cuuint64_t)sh, (cuuint64_t)sp, (cuuint64_t)si};
    cuuint32_t box[5] = {kGemmBlockK, (cuuint32_t)d.bw, (cuuint32_t)d.bh, 1, (cuuint32_t)d.bn};
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
  for (p.bh_log2 = 0; (1 << p.bh_log2) < d.bh; ++p.bh_log2) {}
  p.W = d.out_w;
  p.H = d.out_h;
  p.n_img = d.n_img;
  p.tiles_x = (d.out_w + d.bw - 1) / d.bw;
  p.tiles_y = (d.out_h + d.bh - 1) / d.bh;
  p.tiles_img = (d.n_img + d.bn - 1) / d.bn;
  p.n_tiles = d.cout / bn_tile;
  p.num_taps = d.num_taps;
  p.cblocks = d.cin / kGemmBlockK;
  memcpy(p.tap_dx, d.tap_dx, sizeof(p.tap_dx));
  memcpy(p.tap_dy, d.tap_dy, sizeof(p.tap_dy));
  memcpy(p.tap_plane, d.tap_plane, sizeof(p.tap_plane));
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
  {
    const size_t sm[3][3] = {{Cfg<64, 0>::kSmem, Cfg<64, 1>::kSmem, Cfg<64, 2>::kSmem},
                             {Cfg<128, 0>::kSmem, Cfg<128, 1>::kSmem, Cfg<128, 2>::kSmem},
                             {Cfg<256, 0>::kSmem, Cfg<256, 1>::kSmem, Cfg<256, 2>::kSmem}};
    plan->smem = sm[bn_tile == 256 ? 2 : (bn_tile == 128 ? 1 : 0)][mode];
  }
  plan->flops = 2.0 * d.n_img * d.out_h * d.out_w * (double)d.cout * d.num_taps * d.cin;
  return nullptr;
}

namespace {
template <int BN, int MODE>
cudaError_t launch_one(const ConvGemmPlan& plan, cudaStream_t stream) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(plan.grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = plan.smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = Cfg<BN, MODE>::kCS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, conv_gemm_kernel<BN, MODE>, plan.p);
}
template <int BN>
cudaError_t launch_bn(const ConvGemmPlan& plan, cudaStream_t stream) {
  switch (plan.mode) {
    case 2: return launch_one<BN, 2>(plan, stream);
    case 1: return launch_one<BN, 1>(plan, stream);
    default: return launch_one<BN, 0>(plan, stream);
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
