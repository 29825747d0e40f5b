// Non-GEMM network operators (see net_ops.h).  All HBM-bound or tiny; fp32 math, fp16 storage.
#include "net_ops.h"

#include "common.h"

#include <cfloat>
#include <stdexcept>

namespace fpb {

namespace {

constexpr int kDim = 512;
constexpr int kTok = 400;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Loads this lane's 16 channels of one 512-channel row ([8l, 8l+8) and [256 + 8l, 256 + 8l + 8): two
// coalesced 16-byte loads), normalises the row (LayerNorm, eps 1e-5, fp32 math) in place.
__device__ __forceinline__ void load_row16(const __half* __restrict__ x, long long row, int lane, float (&v)[16]) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const uint4 a = *reinterpret_cast<const uint4*>(x + row * kDim + h * 256 + lane * 8);
    const __half2* ah = reinterpret_cast<const __half2*>(&a);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 fa = __half22float2(ah[e]);
      v[h * 8 + 2 * e] = fa.x;
      v[h * 8 + 2 * e + 1] = fa.y;
    }
  }
}
__device__ __forceinline__ void normalise_row16(float (&v)[16], const float (&g)[16], const float (&b)[16]) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  const float mean = warp_sum(s) * (1.f / kDim);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float d = v[i] - mean;
    q += d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / kDim) + 1e-5f);
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = (v[i] - mean) * rstd * g[i] + b[i];
}
__device__ __forceinline__ void load_affine16(const float* __restrict__ gamma, const float* __restrict__ beta,
                                              int lane, float (&g)[16], float (&b)[16]) {
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float4 gv = __ldg(reinterpret_cast<const float4*>(gamma + h * 256 + lane * 8 + 4 * e));
      const float4 bv = __ldg(reinterpret_cast<const float4*>(beta + h * 256 + lane * 8 + 4 * e));
      g[h * 8 + 4 * e + 0] = gv.x; g[h * 8 + 4 * e + 1] = gv.y; g[h * 8 + 4 * e + 2] = gv.z; g[h * 8 + 4 * e + 3] = gv.w;
      b[h * 8 + 4 * e + 0] = bv.x; b[h * 8 + 4 * e + 1] = bv.y; b[h * 8 + 4 * e + 2] = bv.z; b[h * 8 + 4 * e + 3] = bv.w;
    }
}

// out = LayerNorm(x): one warp per kLnRows consecutive rows (all loads issued before the first
// reduction).  The residual sum x = a + f(a) was formed by the producing GEMM's epilogue.
constexpr int kLnRows = 4;
__global__ void __launch_bounds__(256)
layernorm_kernel(const __half* __restrict__ x, const float* __restrict__ gamma,
                 const float* __restrict__ beta, __half* __restrict__ out, long long rows) {
  const long long row0 = ((static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5) * kLnRows;
  const int lane = threadIdx.x & 31;
  if (row0 >= rows) return;
  float v[kLnRows][16], g[16], b[16];
#pragma unroll
  for (int r = 0; r < kLnRows; ++r)
    if (row0 + r < rows) load_row16(x, row0 + r, lane, v[r]);
  load_affine16(gamma, beta, lane, g, b);
#pragma unroll
  for (int r = 0; r < kLnRows; ++r) {
    if (row0 + r >= rows) break;
    normalise_row16(v[r], g, b);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint4 o;
      __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
      for (int e = 0; e < 4; ++e) oh[e] = __floats2half2_rn(v[r][h * 8 + 2 * e], v[r][h * 8 + 2 * e + 1]);
      *reinterpret_cast<uint4*>(out + (row0 + r) * kDim + h * 256 + lane * 8) = o;
    }
  }
}

// mean over the 400 tokens of LayerNorm(x): the normalised tokens are only ever consumed through
// their mean (pose heads), so they are never written.  One 512-thread CTA per hypothesis: warp w
// normalises tokens w, w+16, ... and keeps a running sum in registers; the 16 partial sums are
// combined through shared memory in a fixed order (deterministic).
__global__ void __launch_bounds__(512)
layernorm_mean_kernel(const __half* __restrict__ x, const float* __restrict__ gamma,
                      const float* __restrict__ beta, float* __restrict__ out, float* __restrict__ partial) {
  // The 400 tokens of a hypothesis are always summed as kLnMeanSlices slices of 50 (warp w of a slice takes
  // tokens w, w + 16, ... of the slice; the 16 warp sums are added in warp order; the slice sums are added in
  // slice order), so the result does not depend on how many hypotheses share the launch.  What varies is
  // how many CTAs share the slices: gridDim.y = 1 (many hypotheses: this CTA does all slices and the final
  // sum) up to kLnMeanSlices (Track: one slice per CTA, layernorm_mean_combine_kernel adds them).
  __shared__ float part[16][kDim];
  __shared__ float slice_sum[kLnMeanSlices][kDim];
  const int n = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kPer = kTok / kLnMeanSlices;
  const int per_cta = kLnMeanSlices / gridDim.y;
  float g[16], b[16];
  load_affine16(gamma, beta, lane, g, b);
  const long long row0 = static_cast<long long>(n) * kTok;
  for (int sl = blockIdx.y * per_cta; sl < (blockIdx.y + 1) * per_cta; ++sl) {
    const int t_begin = sl * kPer, t_end = t_begin + kPer;
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    for (int t = t_begin + warp; t < t_end; t += 32) {
      float v0[16], v1[16];
      const bool two = t + 16 < t_end;
      load_row16(x, row0 + t, lane, v0);
      if (two) load_row16(x, row0 + t + 16, lane, v1);
      normalise_row16(v0, g, b);
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] += v0[i];
      if (two) {
        normalise_row16(v1, g, b);
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] += v1[i];
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int e = 0; e < 8; ++e) part[warp][h * 256 + lane * 8 + e] = acc[h * 8 + e];
    __syncthreads();
    float sum = 0.f;
#pragma unroll 8
    for (int w = 0; w < 16; ++w) sum += part[w][threadIdx.x];
    if (gridDim.y == 1) slice_sum[sl][threadIdx.x] = sum;
    else partial[(static_cast<long long>(n) * kLnMeanSlices + sl) * kDim + threadIdx.x] = sum;
    __syncthreads();  // part[] is rewritten by the next slice
  }
  if (gridDim.y == 1) {
    float sum = 0.f;
    for (int k = 0; k < kLnMeanSlices; ++k) sum += slice_sum[k][threadIdx.x];
    out[static_cast<long long>(n) * kDim + threadIdx.x] = sum * (1.f / kTok);
  }
}
__global__ void __launch_bounds__(512)
layernorm_mean_combine_kernel(const float* __restrict__ partial, float* __restrict__ out) {
  float sum = 0.f;
  for (int k = 0; k < kLnMeanSlices; ++k)
    sum += partial[(static_cast<long long>(blockIdx.x) * kLnMeanSlices + k) * kDim + threadIdx.x];
  out[static_cast<long long>(blockIdx.x) * kDim + threadIdx.x] = sum * (1.f / kTok);
}

// One 1024-thread CTA per hypothesis: four groups of 256 threads each sum 100 tokens (one coalesced
// 1 KB row per step), then the groups are combined in a fixed order (deterministic).
__global__ void __launch_bounds__(1024)
token_mean_kernel(const __half* __restrict__ x, float* __restrict__ out) {
  __shared__ float2 part[4][256];
  const int n = blockIdx.x, grp = threadIdx.x >> 8, t256 = threadIdx.x & 255;
  const __half2* p = reinterpret_cast<const __half2*>(x + (static_cast<long long>(n) * kTok + grp * 100) * kDim) + t256;
  float sx = 0.f, sy = 0.f;
  // 25 independent row loads in flight per thread (the sums stay in token order): with one CTA per hypothesis the
  // kernel is a chain of memory round trips, 4 batches instead of 25
  for (int t0 = 0; t0 < 100; t0 += 25) {
    __half2 v[25];
#pragma unroll
    for (int t = 0; t < 25; ++t) v[t] = p[static_cast<long long>(t0 + t) * (kDim / 2)];
#pragma unroll
    for (int t = 0; t < 25; ++t) {
      const float2 f = __half22float2(v[t]);
      sx += f.x;
      sy += f.y;
    }
  }
  part[grp][t256] = make_float2(sx, sy);
  __syncthreads();
  if (grp == 0) {
    const float ax = ((part[0][t256].x + part[1][t256].x) + part[2][t256].x) + part[3][t256].x;
    const float ay = ((part[0][t256].y + part[1][t256].y) + part[2][t256].y) + part[3][t256].y;
    *reinterpret_cast<float2*>(out + static_cast<long long>(n) * kDim + 2 * t256) =
        make_float2(ax * (1.f / kTok), ay * (1.f / kTok));
  }
}

// out[rows, outs] = in[rows, k] * w[outs, k]^T + b, fp32 (the scorer tail after the token mean: 252 rows).
// 32 x 32 output tile per 256-thread CTA (so a 252 x 512 product spreads over 128 CTAs); thread (ty, tx)
// owns row ty and outputs tx, tx+8, tx+16, tx+24 and walks k four at a time with 16-byte shared loads
// (row pitch 132 floats: the eight lanes of a quarter-warp hit disjoint banks, the row operand is a
// broadcast).  Sums run in ascending k.
// Staging: K chunks of 128 floats through a 4-stage cp.async ring (for K = 512 every load of the CTA is in flight at
// once: the kernel costs one memory round trip plus ~3 us of FMAs instead of 16 dependent chunk round trips).
constexpr int kSGTile = 32, kSGChunk = 128, kSGPitch = kSGChunk + 4, kSGStages = 4;
constexpr size_t kSmallLinearSmem = sizeof(float) * kSGStages * 2 * kSGTile * kSGPitch;
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc, bool valid) {
  const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  const int bytes = valid ? 16 : 0;  // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(bytes) : "memory");
}
__global__ void __launch_bounds__(256)
small_linear_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ b,
                    float* __restrict__ out, int rows, int k, int outs) {
  extern __shared__ __align__(16) float s_lin[];  // [stage][a | w][32 x pitch]
  const int r0 = blockIdx.x * kSGTile, o0 = blockIdx.y * kSGTile;
  const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
  const int n_chunks = (k + kSGChunk - 1) / kSGChunk;
  auto issue = [&](int c) {
    if (c < n_chunks) {
      float* sa = s_lin + (size_t)(c % kSGStages) * 2 * kSGTile * kSGPitch;
      float* sw = sa + kSGTile * kSGPitch;
      const int kc = c * kSGChunk;
#pragma unroll
      for (int i = threadIdx.x; i < kSGTile * (kSGChunk / 4); i += 256) {
        const int lr = i >> 5, lc = (i & 31) * 4;
        const bool in_k = kc + lc < k;  // k is a multiple of 4 (checked by the launcher)
        const bool a_ok = r0 + lr < rows && in_k, w_ok = o0 + lr < outs && in_k;
        cp_async16(sa + lr * kSGPitch + lc, in + static_cast<long long>(a_ok ? r0 + lr : 0) * k + (in_k ? kc + lc : 0), a_ok);
        cp_async16(sw + lr * kSGPitch + lc, w + static_cast<long long>(w_ok ? o0 + lr : 0) * k + (in_k ? kc + lc : 0), w_ok);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");  // (possibly empty: keeps the group count uniform)
  };
#pragma unroll
  for (int c = 0; c < kSGStages; ++c) issue(c);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c = 0; c < n_chunks; ++c) {
    asm volatile("cp.async.wait_group %0;" ::"n"(kSGStages - 1) : "memory");
    __syncthreads();
    const float* pa = s_lin + (size_t)(c % kSGStages) * 2 * kSGTile * kSGPitch + ty * kSGPitch;
    const float* pw = s_lin + (size_t)(c % kSGStages) * 2 * kSGTile * kSGPitch + kSGTile * kSGPitch + tx * kSGPitch;
#pragma unroll 8
    for (int kk = 0; kk < kSGChunk; kk += 4) {
      const float4 a4 = *reinterpret_cast<const float4*>(pa + kk);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 w4 = *reinterpret_cast<const float4*>(pw + j * 8 * kSGPitch + kk);
        acc[j] = fmaf(a4.x, w4.x, acc[j]);
        acc[j] = fmaf(a4.y, w4.y, acc[j]);
        acc[j] = fmaf(a4.z, w4.z, acc[j]);
        acc[j] = fmaf(a4.w, w4.w, acc[j]);
      }
    }
    __syncthreads();  // the slot is refilled by chunk c + stages
    issue(c + kSGStages);
  }
  if (r0 + ty < rows) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int o = o0 + tx + 8 * j;
      if (o < outs) out[static_cast<long long>(r0 + ty) * outs + o] = acc[j] + (b ? b[o] : 0.f);
    }
  }
}

// out[r] = in[r, :] . w + b: one warp per row (the scorer's final Linear(512, 1)).
__global__ void row_dot_kernel(const float* __restrict__ in, const float* __restrict__ w,
                               const float* __restrict__ b, float* __restrict__ out, int rows, int k) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= rows) return;
  float acc = 0.f;
  for (int c = lane; c < k; c += 32) acc = fmaf(in[static_cast<long long>(r) * k + c], w[c], acc);
  acc = warp_sum(acc);
  if (lane == 0) out[r] = acc + (b ? b[0] : 0.f);
}

// Cross-hypothesis attention of the scorer tail: n tokens (one per hypothesis), 4 heads x 128 dims,
// fp32.  One CTA = 8 queries (one per warp) of one head.  Keys, then values, stream through shared
// memory in tiles of 32 rows shared by the 8 queries: for the scores lane j owns key j of the tile
// (128-d dot product from shared memory, no shuffles); for the weighted sum lane l owns dims
// [4l, 4l+4).  Row pitch 132 floats keeps the 16-byte shared loads conflict-free.
constexpr int kCAQ = 8, kCATile = 32, kCAPitch = 132;
constexpr size_t kCrossAttnMaxSmem = 200 * 1024;
__global__ void __launch_bounds__(kCAQ * 32)
cross_attention_kernel(const float* __restrict__ qkv, float* __restrict__ out, int n) {
  extern __shared__ __align__(16) float s_ca[];
  float* s_q = s_ca;                        // [kCAQ][128]
  float* s_t = s_q + kCAQ * 128;            // [kCATile][kCAPitch]
  float* s_p = s_t + kCATile * kCAPitch;    // [kCAQ][n_pad]
  const int n_tiles = (n + kCATile - 1) / kCATile, n_pad = n_tiles * kCATile;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y, qi = blockIdx.x * kCAQ + warp;
  const bool q_ok = qi < n;
  {
    const float4 q = q_ok ? *reinterpret_cast<const float4*>(qkv + static_cast<long long>(qi) * 1536 + head * 128 + lane * 4)
                          : make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(s_q + warp * 128 + lane * 4) = q;
  }
  auto load_tile = [&](int t, int col0) {  // rows [32t, 32t+32) of K (col0 = 512) or V (col0 = 1024)
#pragma unroll
    for (int i = 0; i < (kCATile * 32) / (kCAQ * 32); ++i) {
      const int idx = threadIdx.x + i * kCAQ * 32;
      const int r = idx >> 5, c4 = (idx & 31) * 4;
      const int j = t * kCATile + r;
      const float4 v = j < n ? __ldg(reinterpret_cast<const float4*>(qkv + static_cast<long long>(j) * 1536 + col0 +
                                                                      head * 128 + c4))
                             : make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(s_t + r * kCAPitch + c4) = v;
    }
  };
  const float scale = 0.08838834764831845f;
  float* p = s_p + warp * n_pad;
  float mx = -FLT_MAX;
  for (int t = 0; t < n_tiles; ++t) {
    __syncthreads();  // previous tile fully consumed (and s_q visible on the first trip)
    load_tile(t, 512);
    __syncthreads();
    const float* kr = s_t + lane * kCAPitch;
    const float* qr = s_q + warp * 128;
    float d = 0.f;
#pragma unroll 8
    for (int c = 0; c < 128; c += 4) {
      const float4 k4 = *reinterpret_cast<const float4*>(kr + c);
      const float4 q4 = *reinterpret_cast<const float4*>(qr + c);
      d = fmaf(q4.x, k4.x, d);
      d = fmaf(q4.y, k4.y, d);
      d = fmaf(q4.z, k4.z, d);
      d = fmaf(q4.w, k4.w, d);
    }
    const int j = t * kCATile + lane;
    d = j < n ? d * scale : -FLT_MAX;
    p[j] = d;
    mx = fmaxf(mx, d);
  }
  __syncwarp();  // p[] is read lane-strided below
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int j = lane; j < n_pad; j += 32) {
    const float e = j < n ? __expf(p[j] - mx) : 0.f;
    p[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = 0; t < n_tiles; ++t) {
    __syncthreads();
    load_tile(t, 1024);
    __syncthreads();
    const float* pt = p + t * kCATile;
#pragma unroll 8
    for (int j = 0; j < kCATile; ++j) {
      const float pj = pt[j];
      const float4 v = *reinterpret_cast<const float4*>(s_t + j * kCAPitch + lane * 4);
      acc.x = fmaf(pj, v.x, acc.x);
      acc.y = fmaf(pj, v.y, acc.y);
      acc.z = fmaf(pj, v.z, acc.z);
      acc.w = fmaf(pj, v.w, acc.w);
    }
  }
  if (q_ok) {
    const float inv = 1.f / sum;
    *reinterpret_cast<float4*>(out + static_cast<long long>(qi) * 512 + head * 128 + lane * 4) =
        make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
  }
}

__global__ void argmax_pose_kernel(const float* __restrict__ scores, const float* __restrict__ poses, int n,
                                   int* __restrict__ best_index, float* __restrict__ best_pose) {
  __shared__ float s_v[32];
  __shared__ int s_i[32];
  float bv = -FLT_MAX;
  int bi = INT_MAX;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = scores[i];
    if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = bv; s_i[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int nw = blockDim.x >> 5;
    bv = threadIdx.x < nw ? s_v[threadIdx.x] : -FLT_MAX;
    bi = threadIdx.x < nw ? s_i[threadIdx.x] : INT_MAX;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (bi == INT_MAX) bi = 0;  // all scores NaN: keep index 0 like an empty max_element walk
    if (threadIdx.x == 0) *best_index = bi;
    if (threadIdx.x < 16) best_pose[threadIdx.x] = poses[16 * bi + threadIdx.x];
  }
}

// One warp per hypothesis.
__global__ void refine_update_kernel(const float* __restrict__ mean_t, const float* __restrict__ mean_r,
                                     const float* __restrict__ wt, const float* __restrict__ bt,
                                     const float* __restrict__ wr, const float* __restrict__ br, float half_d,
                                     float* __restrict__ poses, float* __restrict__ trans_out,
                                     float* __restrict__ rot_out, int n) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  float d[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  // (unrolled: all 16 x 8 loads of a lane in flight at once instead of 16 dependent round trips; same summation order)
#pragma unroll
  for (int k = 0; k < kDim / 32; ++k) {
    const int c = lane + 32 * k;
    const float a = mean_t[static_cast<long long>(i) * kDim + c];
    const float b = mean_r[static_cast<long long>(i) * kDim + c];
#pragma unroll
    for (int o = 0; o < 3; ++o) {
      d[o] += a * wt[o * kDim + c];
      d[3 + o] += b * wr[o * kDim + c];
    }
  }
#pragma unroll
  for (int o = 0; o < 6; ++o) d[o] = warp_sum(d[o]);
  if (lane != 0) return;
  float tr[3], ro[3];
#pragma unroll
  for (int o = 0; o < 3; ++o) {
    tr[o] = d[o] + bt[o];
    ro[o] = d[3 + o] + br[o];
    trans_out[3 * i + o] = tr[o];
    rot_out[3 * i + o] = ro[o];
  }
  float* p = poses + 16 * i;  // column-major
  // t += trans * d/2
  p[12] += tr[0] * half_d;
  p[13] += tr[1] * half_d;
  p[14] += tr[2] * half_d;
  // dR = Rodrigues(v)^T, v = tanh(rot) * 20 deg; zero vector -> identity
  const float vx = tanhf(ro[0]) * 0.349065850398865f, vy = tanhf(ro[1]) * 0.349065850398865f,
              vz = tanhf(ro[2]) * 0.349065850398865f;
  const float ang = sqrtf(vx * vx + vy * vy + vz * vz);
  float rm[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  if (ang > 0.f) {
    const float ax = vx / ang, ay = vy / ang, az = vz / ang;
    const float c = cosf(ang), s = sinf(ang), k = 1.f - c;
    rm[0] = c + k * ax * ax;      rm[1] = k * ax * ay - s * az; rm[2] = k * ax * az + s * ay;
    rm[3] = k * ay * ax + s * az; rm[4] = c + k * ay * ay;      rm[5] = k * ay * az - s * ax;
    rm[6] = k * az * ax - s * ay; rm[7] = k * az * ay + s * ax; rm[8] = c + k * az * az;
  }
  float r_old[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) r_old[3 * r + c] = p[4 * c + r];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      // (rm^T * R)(r,c) = sum_k rm[k][r] * R[k][c]
      p[4 * c + r] = rm[0 + r] * r_old[c] + rm[3 + r] * r_old[3 + c] + rm[6 + r] * r_old[6 + c];
    }
}

}  // namespace

const char* net_ops_init() {
  // per device (a process may drive several GPUs), idempotent
  if (cudaFuncSetAttribute(cross_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCrossAttnMaxSmem) !=
      cudaSuccess)
    return "cudaFuncSetAttribute(cross_attention smem) failed (not an sm_100a device?)";
  if (cudaFuncSetAttribute(small_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmallLinearSmem) !=
      cudaSuccess)
    return "cudaFuncSetAttribute(small_linear smem) failed (not an sm_100a device?)";
  if (preload_kernels(layernorm_kernel, layernorm_mean_kernel, layernorm_mean_combine_kernel, token_mean_kernel,
                      row_dot_kernel, argmax_pose_kernel, refine_update_kernel) != cudaSuccess)
    return "loading the network helper kernels failed (not an sm_100a device?)";
  return nullptr;
}

void launch_layernorm(const __half* x, const float* gamma, const float* beta, __half* out, long long rows,
                      cudaStream_t s) {
  if (rows <= 0) return;
  const int threads = 256;
  const long long warps = (rows + kLnRows - 1) / kLnRows;
  const long long blocks = (warps * 32 + threads - 1) / threads;
  layernorm_kernel<<<(unsigned)blocks, threads, 0, s>>>(x, gamma, beta, out, rows);
}

int launch_layernorm_mean(const __half* x, const float* gamma, const float* beta, float* out, int n,
                          float* partial, cudaStream_t s) {
  if (n <= 0) return 0;
  // CTAs per hypothesis: enough to fill the machine, never more than the slices
  int split = 1;
  while (split < kLnMeanSlices && n * split < 128) split *= 2;
  layernorm_mean_kernel<<<dim3(n, split), 512, 0, s>>>(x, gamma, beta, out, partial);
  if (split == 1) return 1;
  layernorm_mean_combine_kernel<<<n, 512, 0, s>>>(partial, out);
  return 2;
}

void launch_token_mean(const __half* x, float* out, int n, cudaStream_t s) {
  if (n <= 0) return;
  token_mean_kernel<<<n, 1024, 0, s>>>(x, out);
}

void launch_small_linear(const float* in, const float* w, const float* b, float* out, int rows, int k, int outs,
                         cudaStream_t s) {
  if (rows <= 0 || outs <= 0) return;
  if (k % 4 != 0) throw std::invalid_argument("small_linear: k must be a multiple of 4");
  if (outs == 1) {
    row_dot_kernel<<<(rows * 32 + 127) / 128, 128, 0, s>>>(in, w, b, out, rows, k);
    return;
  }
  dim3 grid((rows + kSGTile - 1) / kSGTile, (outs + kSGTile - 1) / kSGTile);
  small_linear_kernel<<<grid, 256, kSmallLinearSmem, s>>>(in, w, b, out, rows, k, outs);
}

void launch_cross_attention(const float* qkv, float* out, int n, cudaStream_t s) {
  if (n <= 0) return;
  const int n_pad = (n + kCATile - 1) / kCATile * kCATile;
  const size_t smem = sizeof(float) * (kCAQ * 128 + kCATile * kCAPitch + (size_t)kCAQ * n_pad);
  if (smem > kCrossAttnMaxSmem) throw std::invalid_argument("cross_attention: too many hypotheses for one shared-memory score row");
  dim3 grid((n + kCAQ - 1) / kCAQ, 4);
  cross_attention_kernel<<<grid, kCAQ * 32, smem, s>>>(qkv, out, n);
}

void launch_argmax_pose(const float* scores, const float* poses, int n, int* best_index, float* best_pose,
                        cudaStream_t s) {
  argmax_pose_kernel<<<1, 256, 0, s>>>(scores, poses, n, best_index, best_pose);
}

void launch_refine_update(const float* mean_t, const float* mean_r, const float* wt, const float* bt,
                          const float* wr, const float* br, float diameter, float* poses, float* trans_out,
                          float* rot_out, int n, cudaStream_t s) {
  if (n <= 0) return;
  const int threads = 128;
  refine_update_kernel<<<(n * 32 + threads - 1) / threads, threads, 0, s>>>(mean_t, mean_r, wt, bt, wr, br,
                                                                           diameter / 2.f, poses, trans_out,
                                                                           rot_out, n);
}

}  // namespace fpb
