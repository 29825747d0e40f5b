// Preprocessing kernels of the Register/Track hot path (see preprocess.h).
// Reference behaviour restated per kernel (paths relative to detection_6d_foundationpose/src/).
// This translation unit is compiled with -fmad=false: expressions round exactly as written.
#include "preprocess.h"

#include "common.h"

#include <cfloat>
#include <cstdlib>

namespace fpb {

namespace {

constexpr float kMinDepth = 0.001f;  // foundationpose_render.hpp:26
constexpr float kMaxDepth = 4.0f;    // foundationpose_render.hpp:27 (applied to normalised coords)
constexpr float kZFar = 100.0f;      // foundationpose_sampling.cu.hpp defaults

// ------------------------------------------------------------------ depth -> xyz
// foundationpose_utils.cu:3-32.  Invalid pixels are written as zero (the reference leaves them
// uninitialised; they are zeroed downstream by the z < min_depth test either way).
__global__ void depth_to_xyz_kernel(const float* __restrict__ depth, float* __restrict__ xyz, int h,
                                    int w, float fx, float fy, float cx, float cy) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= h * w) return;
  const int row = idx / w, col = idx - row * w;
  const float d = depth[idx];
  float x = 0.f, y = 0.f, z = 0.f;
  if (!(d < kMinDepth)) {
    x = ((float)col - cx) * d / fx;
    y = ((float)row - cy) * d / fy;
    z = d;
  }
  xyz[3 * idx + 0] = x;
  xyz[3 * idx + 1] = y;
  xyz[3 * idx + 2] = z;
}

// ------------------------------------------------------------------ sampler depth filters
// ErodeDepth (foundationpose_sampling.cu:21-82: radius 2, 1 mm, ratio 0.8, zfar 100) and BilateralFilterDepth
// (:84-164: radius 2, sigmaD 2, sigmaR 1e5) as ONE kernel: a CTA stages its 32 x 8 tile of the depth image with a
// 4-pixel halo in shared memory, erodes the 36 x 12 pixels the bilateral window of the tile can reach, and filters
// from that - the eroded image never touches HBM unless a caller asks for it.  Per pixel the neighbours are visited
// in the reference's order (columns outside, rows inside) and summed in that order, so both results are the same
// floats the two separate kernels of the reference produce.  Block (0, 0) also resets the sampler's scan counters.
constexpr int kFiltTX = 32, kFiltTY = 8;
__device__ __forceinline__ bool depth_valid(float d) { return d >= 0.1f && d < kZFar; }
__global__ void __launch_bounds__(kFiltTX * kFiltTY)
depth_filter_kernel(const float* __restrict__ depth, float* __restrict__ eroded_out, float* __restrict__ filtered, int H,
                    int W, int* __restrict__ meta) {
  constexpr int DW = kFiltTX + 8, DH = kFiltTY + 8, EW = kFiltTX + 4, EH = kFiltTY + 4;
  __shared__ float s_d[DH][DW];
  __shared__ float s_e[EH][EW];
  const int tid = threadIdx.y * kFiltTX + threadIdx.x;
  const int x0 = blockIdx.x * kFiltTX, y0 = blockIdx.y * kFiltTY;
  if (meta && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) {
    meta[0] = 0;
    meta[1] = INT_MAX;
    meta[2] = -1;
    meta[3] = INT_MAX;
    meta[4] = -1;
  }
  for (int i = tid; i < DW * DH; i += kFiltTX * kFiltTY) {
    const int ly = i / DW, lx = i - ly * DW;
    const int gx = x0 - 4 + lx, gy = y0 - 4 + ly;
    s_d[ly][lx] = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? depth[gy * W + gx] : 0.f;
  }
  __syncthreads();
  for (int i = tid; i < EW * EH; i += kFiltTX * kFiltTY) {
    const int ly = i / EW, lx = i - ly * EW;
    const int w = x0 - 2 + lx, h = y0 - 2 + ly;  // image coordinates of this eroded pixel
    float e = 0.f;
    if (w >= 0 && w < W && h >= 0 && h < H) {
      const float d_ori = s_d[ly + 2][lx + 2];
      if (depth_valid(d_ori)) {
        float bad = 0.f, total = 0.f;
        for (int u = w - 2; u <= w + 2; ++u) {
          if (u < 0 || u >= W) continue;
          for (int v = h - 2; v <= h + 2; ++v) {
            if (v < 0 || v >= H) continue;
            const float cur = s_d[v - y0 + 4][u - x0 + 4];
            total += 1.f;
            if (!depth_valid(cur) || fabsf(cur - d_ori) > 0.001f) bad += 1.f;
          }
        }
        e = (bad / total) > 0.8f ? 0.f : d_ori;
      }
      if (eroded_out && lx >= 2 && lx < EW - 2 && ly >= 2 && ly < EH - 2) eroded_out[h * W + w] = e;
    }
    s_e[ly][lx] = e;
  }
  __syncthreads();
  const int w = x0 + threadIdx.x, h = y0 + threadIdx.y;
  if (w >= W || h >= H) return;
  float mean = 0.f;
  int num_valid = 0;
  for (int u = w - 2; u <= w + 2; ++u) {
    if (u < 0 || u >= W) continue;
    for (int v = h - 2; v <= h + 2; ++v) {
      if (v < 0 || v >= H) continue;
      const float cur = s_e[v - y0 + 2][u - x0 + 2];
      if (depth_valid(cur)) {
        ++num_valid;
        mean += cur;
      }
    }
  }
  if (num_valid == 0) {
    filtered[h * W + w] = 0.f;
    return;
  }
  mean /= (float)num_valid;
  const float center = s_e[threadIdx.y + 2][threadIdx.x + 2];
  float sum_w = 0.f, sum = 0.f;
  const float sigma_d = 2.f, sigma_r = 100000.f;
  for (int u = w - 2; u <= w + 2; ++u) {
    if (u < 0 || u >= W) continue;
    for (int v = h - 2; v <= h + 2; ++v) {
      if (v < 0 || v >= H) continue;
      const float cur = s_e[v - y0 + 2][u - x0 + 2];
      if (depth_valid(cur) && fabsf(cur - mean) < 0.01f) {
        const float wgt = expf(-((float)((u - w) * (u - w) + (v - h) * (v - h))) / (2.f * sigma_d * sigma_d) -
                               (center - cur) * (center - cur) / (2.f * sigma_r * sigma_r));
        sum_w += wgt;
        sum += wgt * cur;
      }
    }
  }
  filtered[h * W + w] = (sum_w > 0.f) ? sum / sum_w : 0.f;
}

// ------------------------------------------------------------------ translation guess
// GuessTranslation, foundationpose_sampling.cpp:250-298, in two kernels:
//  (1) mask_scan_kernel (whole grid): mask bounding box (warp-reduced atomics) and compaction of the
//      valid depths (mask > 0 and depth >= min_depth) into a list (warp-aggregated atomicAdd);
//  (2) median_select_kernel (one CTA): exact median of the list by 4-pass radix select on the float
//      bit patterns (depths are positive, so the patterns order like the values; the reference uses
//      std::sort), then K^-1 back-projection.  meta = {count, umin, umax, vmin, vmax}.
__global__ void mask_scan_kernel(const float* __restrict__ depth, const uint8_t* __restrict__ mask, int h, int w,
                                 unsigned* __restrict__ list, int* __restrict__ meta) {
  const int npx = h * w;
  const int lane = threadIdx.x & 31;
  int umin = INT_MAX, umax = -1, vmin = INT_MAX, vmax = -1;
  for (int base = (blockIdx.x * blockDim.x + threadIdx.x) - lane; base < npx; base += gridDim.x * blockDim.x) {
    const int i = base + lane;
    bool valid = false;
    unsigned key = 0;
    if (i < npx && mask[i]) {
      const int v = i / w, u = i - v * w;
      umin = min(umin, u); umax = max(umax, u); vmin = min(vmin, v); vmax = max(vmax, v);
      const float d = depth[i];
      if (d >= kMinDepth) {
        valid = true;
        key = __float_as_uint(d);
      }
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, valid);
    if (ballot) {
      int pos = 0;
      if (lane == 0) pos = atomicAdd(&meta[0], __popc(ballot));
      pos = __shfl_sync(0xffffffffu, pos, 0);
      if (valid) list[pos + __popc(ballot & ((1u << lane) - 1u))] = key;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    umin = min(umin, __shfl_xor_sync(0xffffffffu, umin, o));
    umax = max(umax, __shfl_xor_sync(0xffffffffu, umax, o));
    vmin = min(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
    vmax = max(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
  }
  if (lane == 0 && umax >= 0) {
    atomicMin(&meta[1], umin);
    atomicMax(&meta[2], umax);
    atomicMin(&meta[3], vmin);
    atomicMax(&meta[4], vmax);
  }
}

__global__ void __launch_bounds__(1024, 1)
median_select_kernel(const unsigned* __restrict__ list, const int* __restrict__ meta, Intrinsics K,
                     float* __restrict__ center, int* __restrict__ status, const float* __restrict__ grid,
                     float* __restrict__ poses, int n_poses) {
  __shared__ unsigned hist[256];
  __shared__ unsigned s_prefix, s_rank, s_cnt_less, s_max_less;
  __shared__ float s_center[3];
  const int tid = threadIdx.x;
  if (meta[2] < 0) {
    if (tid == 0) atomicOr(status, kStatusEmptyMask);
    return;
  }
  const unsigned total = (unsigned)meta[0];
  if (total == 0) {
    if (tid == 0) atomicOr(status, kStatusNoDepth);
    return;
  }
  if (tid == 0) {
    s_prefix = 0;
    s_rank = total / 2;  // 0-based rank of the upper median
    s_cnt_less = 0;
    s_max_less = 0;
  }
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = tid; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const unsigned prefix = s_prefix;
    const unsigned himask = shift == 24 ? 0u : (0xFFFFFFFFu << (shift + 8));
    for (unsigned i = tid; i < total; i += blockDim.x) {
      const unsigned key = list[i];
      if ((key & himask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid < 32) {
      // bucket holding rank r: warp scan over 32 groups of 8 bins instead of a serial walk over 256 bins
      unsigned bins[8], group = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        bins[j] = hist[tid * 8 + j];
        group += bins[j];
      }
      unsigned incl = group;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
        if (tid >= o) incl += v;
      }
      const unsigned r = s_rank, before = incl - group;
      __syncwarp();  // every lane has read s_rank before the owning lane overwrites it (racecheck: read / write hazard)
      if (r >= before && r < incl) {  // exactly one lane (total counts of this pass exceed r)
        unsigned rr = r - before;
        int j = 0;
        for (; j < 7; ++j) {
          if (rr < bins[j]) break;
          rr -= bins[j];
        }
        s_rank = rr;
        s_prefix = prefix | ((unsigned)(tid * 8 + j) << shift);
      }
    }
    __syncthreads();
  }
  const unsigned hi_key = s_prefix;
  unsigned cnt_less = 0, max_less = 0;
  for (unsigned i = tid; i < total; i += blockDim.x) {
    const unsigned key = list[i];
    if (key < hi_key) {
      ++cnt_less;
      max_less = max(max_less, key);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt_less += __shfl_xor_sync(0xffffffffu, cnt_less, o);
    max_less = max(max_less, __shfl_xor_sync(0xffffffffu, max_less, o));
  }
  if ((tid & 31) == 0) {
    atomicAdd(&s_cnt_less, cnt_less);
    atomicMax(&s_max_less, max_less);
  }
  __syncthreads();
  if (tid == 0) {
    const float hi = __uint_as_float(hi_key);
    float zc = hi;
    if ((total & 1u) == 0) {
      const float lo = (s_cnt_less == total / 2) ? __uint_as_float(s_max_less) : hi;
      zc = (lo + hi) / 2.f;
    }
    const float uc = (float)(meta[1] + meta[2]) / 2.f;
    const float vc = (float)(meta[3] + meta[4]) / 2.f;
    for (int r = 0; r < 3; ++r) {
      s_center[r] = ((K.kinv[3 * r] * uc + K.kinv[3 * r + 1] * vc) + K.kinv[3 * r + 2]) * zc;
      center[r] = s_center[r];
    }
  }
  __syncthreads();
  // GetHypPoses step 4, foundationpose_sampling.cpp:388-391: every rotation of the grid at the guessed translation
  if (poses)
    for (int i = tid; i < n_poses * 16; i += blockDim.x) {
      const int e = i & 15;
      poses[i] = (e >= 12 && e < 15) ? s_center[e - 12] : grid[i];
    }
}

// ------------------------------------------------------------------ per-pose setup
__device__ __forceinline__ void mat4_mul(const float* a, const float* b, float* out) {
  // row-major, k-sum left to right (matches oracle/fp_oracle.py:_mat4_mul)
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      float s = a[4 * i] * b[j];
      for (int k = 1; k < 4; ++k) s = s + a[4 * i + k] * b[4 * k + j];
      out[4 * i + j] = s;
    }
}

// ComputeCropWindowTF / ComputeTF / ConstructBBox2D / ProjectMatrixFromIntrinsics + the matrix
// product of GeneratePoseClipOnCUDA (foundationpose_render.cpp:25-75,123-186,588-593) and the
// zoom constants of generate_pose_clip_kernel (foundationpose_render.cu:382-386), one thread per pose.
__global__ void pose_setup_kernel(const float* __restrict__ poses, int n, Intrinsics K, float diameter,
                                  float crop_ratio, PoseAux* __restrict__ aux) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = poses + 16 * i;  // column-major
  PoseAux a;
  float pose_rm[16];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) pose_rm[4 * r + c] = p[4 * c + r];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) a.rot[3 * r + c] = pose_rm[4 * r + c];
    a.t[r] = pose_rm[4 * r + 3];
  }
  // crop window
  const float rad3 = diameter * crop_ratio / 2.f;
  const float tx = a.t[0], ty = a.t[1], tz = a.t[2];
  const float u0 = ((K.fx * tx + K.k01 * ty) + K.cx * tz) / tz;
  const float v0 = (K.fy * ty + K.cy * tz) / tz;
  float dv_max = 0.f;  // point 0 contributes v0 - v0 = 0
  {
    const float vy_p = (K.fy * (ty + rad3) + K.cy * tz) / tz;
    const float vy_m = (K.fy * (ty - rad3) + K.cy * tz) / tz;
    dv_max = fmaxf(dv_max, v0 - v0);
    dv_max = fmaxf(dv_max, vy_p - v0);
    dv_max = fmaxf(dv_max, vy_m - v0);
  }
  const float radius = fabsf(dv_max);
  const float left = roundf(u0 - radius), right = roundf(u0 + radius);
  const float top = roundf(v0 - radius), bottom = roundf(v0 + radius);
  const float sx = (float)kCrop / (right - left), sy = (float)kCrop / (bottom - top);
  const float tf02 = sx * (-left), tf12 = sy * (-top);
  a.tf[0] = sx; a.tf[1] = 0.f; a.tf[2] = tf02;
  a.tf[3] = 0.f; a.tf[4] = sy; a.tf[5] = tf12;
  a.tf[6] = 0.f; a.tf[7] = 0.f; a.tf[8] = 1.f;
  a.inv_sx = 1.f / sx;
  a.inv_sy = 1.f / sy;
  a.left = left;
  a.top = top;
  // bbox2d = tf^-1 * {(0,0), (W-1,H-1)}
  const float bl = -tf02 / sx, bt = -tf12 / sy;
  a.bbox[0] = bl;
  a.bbox[1] = bt;
  a.bbox[2] = (float)(kCrop - 1) / sx + bl;
  a.bbox[3] = (float)(kCrop - 1) / sy + bt;
  // clip-space zoom
  const float fh = (float)K.img_h, fw = (float)K.img_w;
  const float l = a.bbox[0], t = fh - a.bbox[1], r = a.bbox[2], b = fh - a.bbox[3];
  a.a00 = fw / (r - l);
  a.a11 = fh / (t - b);
  a.a30 = (fw - r - l) / (r - l);
  a.a31 = (fh - t - b) / (t - b);
  // MVP
  const float gl[16] = {1, 0, 0, 0, 0, -1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 1};
  float tmp[16];
  mat4_mul(gl, pose_rm, tmp);
  mat4_mul(K.proj, tmp, a.mvp);
  aux[i] = a;
}

// ------------------------------------------------------------------ render
constexpr int kRenderThreads = 256;                  // split-path kernels
// band kernel: 2 bands of 80 rows per hypothesis (512 threads, 102 KB z-buffer, 2 CTAs = 32 warps per SM) or one band
// of 160 rows (1024 threads, 205 KB); every band CTA walks all triangles
constexpr unsigned kDepthMin = 2200u << 3;
constexpr unsigned kDepthMax = 0xFFFFFFFFu - (2200u << 3);

// Vertex stage, one thread per (hypothesis, vertex): clip-space position
// (generate_pose_clip_kernel, foundationpose_render.cu:363-398), camera-space position
// (transform_points_kernel, :321-341) and headlight diffuse term (transform_normals_kernel, :418-443).
// The reference keeps these as three [N,V,*] fp32 tensors too (pose_clip, pts_cam, diffuse).
// Output: one 32-byte record per (hypothesis, vertex): {clip position, (camera-space position, diffuse)} - the
// rasteriser reads the first half, a shaded pixel both halves of its three vertices as three 32-byte sectors.
// zfill (may be null): the hypothesis's whole-crop z-buffer of the few-hypotheses path, reset here to "empty" (all ones)
// by the same CTAs instead of by a launch of its own.
__global__ void vertex_stage_kernel(DeviceMesh mesh, const PoseAux* __restrict__ aux_all, float4* __restrict__ vrec,
                                    unsigned long long* __restrict__ zfill) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  const int pose = blockIdx.y;
  if (zfill) {
    unsigned long long* z = zfill + (long long)pose * (kCrop * kCrop);
    for (int i = v; i < kCrop * kCrop; i += gridDim.x * blockDim.x) z[i] = ~0ull;
  }
  if (v >= mesh.num_vertices) return;
  const PoseAux& a = aux_all[pose];
  const float x = mesh.vertices[3 * v], y = mesh.vertices[3 * v + 1], z = mesh.vertices[3 * v + 2];
  const float tx = ((a.mvp[0] * x + a.mvp[1] * y) + a.mvp[2] * z) + a.mvp[3];
  const float ty = ((a.mvp[4] * x + a.mvp[5] * y) + a.mvp[6] * z) + a.mvp[7];
  const float tz = ((a.mvp[8] * x + a.mvp[9] * y) + a.mvp[10] * z) + a.mvp[11];
  const float tw = ((a.mvp[12] * x + a.mvp[13] * y) + a.mvp[14] * z) + a.mvp[15];
  float4 c;
  c.x = tx * a.a00 + tw * a.a30;
  c.y = ty * a.a11 + tw * a.a31;
  c.z = tz;
  c.w = tw;
  float4 q;
  q.x = ((a.rot[0] * x + a.rot[1] * y) + a.rot[2] * z) + a.t[0];
  q.y = ((a.rot[3] * x + a.rot[4] * y) + a.rot[5] * z) + a.t[1];
  q.z = ((a.rot[6] * x + a.rot[7] * y) + a.rot[8] * z) + a.t[2];
  const float nx0 = mesh.normals[3 * v], ny0 = mesh.normals[3 * v + 1], nz0 = mesh.normals[3 * v + 2];
  const float nx = (a.rot[0] * nx0 + a.rot[1] * ny0) + a.rot[2] * nz0;
  const float ny = (a.rot[3] * nx0 + a.rot[4] * ny0) + a.rot[5] * nz0;
  const float nz = (a.rot[6] * nx0 + a.rot[7] * ny0) + a.rot[8] * nz0;
  const float l2 = sqrtf((nx * nx + ny * ny) + nz * nz);
  const float val = l2 == 0.f ? 0.f : -nz / l2;
  q.w = fminf(fmaxf(val, 0.f), 1.f);
  const long long o = 2 * ((long long)pose * mesh.num_vertices + v);
  vrec[o] = c;
  vrec[o + 1] = q;
}

__device__ __forceinline__ int snap(float v) {
  int r;
  asm("cvt.rni.sat.s32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ int min3(int a, int b, int c) { return min(a, min(b, c)); }
__device__ __forceinline__ int max3(int a, int b, int c) { return max(a, max(b, c)); }

// Integer depth plane, cudaraster Util.inl:184-210.
__device__ __forceinline__ void setup_pleq(float z0, float z1, float z2, int v0x, int v0y, int d1x,
                                           int d1y, int d2x, int d2y, float area_rcp, unsigned* pl) {
  const float mx = fmaxf(fmaxf(z0, z1), z2);
  const int sh = min(max((__float_as_int(mx) >> 23) - (127 + 22), 0), 8);
  const int t0 = (int)((unsigned)z0 >> sh);
  const int t1 = (int)((unsigned)z1 >> sh) - t0;
  const int t2 = (int)((unsigned)z2 >> sh) - t0;
  const unsigned rcp_mant = (__float_as_int(area_rcp) & 0x007FFFFF) | 0x00800000;
  const int rcp_shift = (23 + 127) - (__float_as_int(area_rcp) >> 23);
  const long long xc = ((long long)t1 * d2y - (long long)t2 * d1y) * (long long)rcp_mant;
  const long long yc = ((long long)t2 * d1x - (long long)t1 * d2x) * (long long)rcp_mant;
  pl[0] = (unsigned)(xc >> (rcp_shift - (sh + 4)));
  pl[1] = (unsigned)(yc >> (rcp_shift - (sh + 4)));
  const int cx = (v0x * 2 + min3(d1x, d2x, 0) + max3(d1x, d2x, 0)) >> 5;
  const int cy = (v0y * 2 + min3(d1y, d2y, 0) + max3(d1y, d2y, 0)) >> 5;
  const int vcx = v0x - cx * 16;
  const int vcy = v0y - cy * 16;
  unsigned z = (unsigned)t0 << sh;
  z -= (unsigned)(((xc >> 13) * vcx + (yc >> 13) * vcy) >> (rcp_shift - (sh + 13)));
  z -= pl[0] * (unsigned)cx + pl[1] * (unsigned)cy;
  pl[2] = z;
}

// z-buffer update "nearest wins, later triangle wins ties" (FineRaster.inl:152-172,338-351) on the key
// (u32 depth << 32) | ~triangle.  Shared memory has no native 64-bit minimum (atomicMin compiles to a compare-and-swap
// loop that always swaps), so the slot is read first and the swap only attempted while the key still wins: occluded
// samples cost one load.  Global memory: one fire-and-forget REDG.MIN.64.
template <bool kSharedZ>
__device__ __forceinline__ void zbuf_min(unsigned long long* slot, unsigned long long key) {
  if (kSharedZ) {
    unsigned long long old = *reinterpret_cast<volatile unsigned long long*>(slot);
    while (key < old) {
      const unsigned long long assumed = old;
      old = atomicCAS(slot, assumed, key);
      if (old == assumed) break;
    }
  } else {
    atomicMin(slot, key);
  }
}

// One snapped (sub)triangle of triangle t inside GL rows [row0, row0 + nrows): cudaraster's depth plane
// (setupTriangle, TriangleSetup.inl:120-177), exact coverage with its fill rule (Util.inl:304-309) and one z-buffer
// update per covered sample (shared or global memory, row pitch 160), whatever the order in which triangles arrive.
// The samples of the bounding box are walked as ONE flat loop (a warp runs max-over-lanes of the box areas instead
// of the sum of its distinct rows) with the three edge functions and the depth stepped by additions: the same 64-bit
// integers as evaluating them per sample.  `live` = false makes the call a no-op (callers keep their warps converged
// instead of branching around it).
template <bool kSharedZ>
__device__ __forceinline__ void raster_piece(bool live, int p0x, int p0y, int p1x, int p1y, int p2x, int p2y, float z0,
                                             float z1, float z2, float rw0, float rw1, float rw2, int t, int row0,
                                             int nrows, unsigned long long* zbuf) {
  constexpr int W = kCrop, H = kCrop;
  constexpr int bias = (W - 1) << 3;  // sample x of pixel px = px*16 - bias
  const int lox = min3(p0x, p1x, p2x), loy = min3(p0y, p1y, p2y);
  const int hix = max3(p0x, p1x, p2x), hiy = max3(p0y, p1y, p2y);
  const int px0 = max((lox + bias + 15) >> 4, 0), px1 = min((hix + bias) >> 4, W - 1);
  const int py0 = max((loy + bias + 15) >> 4, row0), py1 = min((hiy + bias) >> 4, row0 + nrows - 1);
  if (px0 > px1 || py0 > py1) live = false;  // no sample of this band inside the bounding box
  int d1x = p1x - p0x, d1y = p1y - p0y, d2x = p2x - p0x, d2y = p2y - p0y;
  int area = d1x * d2y - d1y * d2x;
  if (area == 0) live = false;
  if (live) {
    if (area < 0) {
      int ti;
      float tf;
      ti = d1x; d1x = d2x; d2x = ti;
      ti = d1y; d1y = d2y; d2y = ti;
      ti = p1x; p1x = p2x; p2x = ti;
      ti = p1y; p1y = p2y; p2y = ti;
      tf = z1; z1 = z2; z2 = tf;
      tf = rw1; rw1 = rw2; rw2 = tf;
      area = -area;
    }
    const float zcoef = (float)(kDepthMax - kDepthMin) * 0.5f;
    const float zbias = (float)(kDepthMax + kDepthMin) * 0.5f;
    const float zv0 = fmaf(z0 * zcoef, rw0, zbias), zv1 = fmaf(z1 * zcoef, rw1, zbias),
                zv2 = fmaf(z2 * zcoef, rw2, zbias);
    unsigned pl[3];
    setup_pleq(zv0, zv1, zv2, p0x + (W << 3) - 8, p0y + (H << 3) - 8, d1x, d1y, d2x, d2y,
               1.0f / (float)area, pl);
    const int e0x = d1x, e0y = d1y, e1x = d2x - d1x, e1y = d2y - d1y, e2x = -d2x, e2y = -d2y;
    const long long adj0 = (e0y > 0 || (e0y == 0 && e0x <= 0)) ? 1 : 0;
    const long long adj1 = (e1y > 0 || (e1y == 0 && e1x <= 0)) ? 1 : 0;
    const long long adj2 = (e2y > 0 || (e2y == 0 && e2x <= 0)) ? 1 : 0;
    const unsigned long long low = (unsigned long long)(0xFFFFFFFFu - (unsigned)t);
    // edge functions at the first sample of the box; one pixel to the right: sx += 16, one row down: sy += 16
    const int sx0 = px0 * 16 - bias, sy0 = py0 * 16 - bias;
    long long v0 = (long long)(p0x - sx0) * e0y - (long long)(p0y - sy0) * e0x - adj0;
    long long v1 = (long long)(p1x - sx0) * e1y - (long long)(p1y - sy0) * e1x - adj1;
    long long v2 = (long long)(p0x - sx0) * e2y - (long long)(p0y - sy0) * e2x - adj2;
    const int bw = px1 - px0 + 1;
    const long long s0x = -16LL * e0y, s1x = -16LL * e1y, s2x = -16LL * e2y;
    // step at the end of a row: back to the first column and one row down (selected, not branched to: the loop then
    // has one data-dependent branch per sample - the z-buffer update - instead of three)
    const long long s0w = 16LL * e0x - (long long)(bw - 1) * s0x, s1w = 16LL * e1x - (long long)(bw - 1) * s1x,
                    s2w = 16LL * e2x - (long long)(bw - 1) * s2x;
    unsigned depth = pl[0] * (unsigned)px0 + pl[1] * (unsigned)py0 + pl[2];
    const unsigned dstep_w = pl[1] - pl[0] * (unsigned)(bw - 1);
    unsigned long long* slot = zbuf + (py0 - row0) * W + px0;
    const int n_samples = bw * (py1 - py0 + 1);
    int x = 0;
    for (int k = 0; k < n_samples; ++k) {
      // depth > kDepthMax would lie behind the cleared depth (FineRaster.inl:338)
      if ((v0 | v1 | v2) >= 0 && depth <= kDepthMax) zbuf_min<kSharedZ>(slot, ((unsigned long long)depth << 32) | low);
      const bool wrap = ++x == bw;
      v0 += wrap ? s0w : s0x;
      v1 += wrap ? s1w : s1x;
      v2 += wrap ? s2w : s2x;
      depth += wrap ? dstep_w : pl[0];
      slot += wrap ? W - bw + 1 : 1;
      x = wrap ? 0 : x;
    }
  }
}

// cudaraster's polygon clipper (Util.inl:101-160, TriangleSetup.inl:311-390) for the rare triangle that
// crosses the near / far plane or leaves the S16 snap range: the triangle is clipped against the frustum
// planes in barycentric space (Sutherland-Hodgman), the polygon is fan-triangulated, every piece is snapped
// and rasterised with the triangle's id.  Floating-point contraction is written out the way nvcc compiled it
// for the reference (this file is built with -fmad=false): a + b*c + d*e = fma(d, e, fma(b, c, a)),
// x*ac + y*bc = fma(x, ac, y*bc).  Not inlined: its arrays stay out of the fast path's register budget.
__device__ __forceinline__ int clip_poly_plane(float* out, const float* in, int num_in, float v0, float v1, float v2) {
  int num_out = 0;
  if (num_in >= 3) {
    int ai = (num_in - 1) * 2;
    float av = fmaf(v2, in[ai + 1], fmaf(v1, in[ai + 0], v0));
    for (int bi = 0; bi < num_in * 2; bi += 2) {
      const float bv = fmaf(v2, in[bi + 1], fmaf(v1, in[bi + 0], v0));
      if (av * bv < 0.0f) {
        const float bc = av / (av - bv);
        const float ac = 1.0f - bc;
        out[num_out + 0] = fmaf(in[ai + 0], ac, in[bi + 0] * bc);
        out[num_out + 1] = fmaf(in[ai + 1], ac, in[bi + 1] * bc);
        num_out += 2;
      }
      if (bv >= 0.0f) {
        out[num_out + 0] = in[bi + 0];
        out[num_out + 1] = in[bi + 1];
        num_out += 2;
      }
      ai = bi;
      av = bv;
    }
  }
  return num_out >> 1;
}
template <bool kSharedZ>
__device__ __noinline__ void raster_clipped(float4 A, float4 B, float4 C, int t, int row0, int nrows,
                                            unsigned long long* zbuf) {
  const float a[4] = {A.x, A.y, A.z, A.w};
  const float d1[4] = {B.x - A.x, B.y - A.y, B.z - A.z, B.w - A.w};
  const float d2[4] = {C.x - A.x, C.y - A.y, C.z - A.z, C.w - A.w};
  const float bb[4] = {B.x, B.y, B.z, B.w}, cc[4] = {C.x, C.y, C.z, C.w};
  float bary[18], temp[18];
  int num = 3;
  bary[0] = 0.0f, bary[1] = 0.0f;
  bary[2] = 1.0f, bary[3] = 0.0f;
  bary[4] = 0.0f, bary[5] = 1.0f;
  for (int ax = 0; ax < 3; ++ax) {
    if ((a[3] < fabsf(a[ax])) | (bb[3] < fabsf(bb[ax])) | (cc[3] < fabsf(cc[ax]))) {
      num = clip_poly_plane(temp, bary, num, a[3] + a[ax], d1[3] + d1[ax], d2[3] + d2[ax]);
      num = clip_poly_plane(bary, temp, num, a[3] - a[ax], d1[3] - d1[ax], d2[3] - d2[ax]);
    }
  }
  if (num < 3) return;
  const float vs = (float)(kCrop << 3);
  float v0[4], v1[4], v2[4];
  for (int k = 0; k < 4; ++k) {
    v0[k] = fmaf(d2[k], bary[1], fmaf(d1[k], bary[0], a[k]));
    v1[k] = fmaf(d2[k], bary[3], fmaf(d1[k], bary[2], a[k]));
  }
  for (int i = 2; i < num; ++i) {
    for (int k = 0; k < 4; ++k) v2[k] = fmaf(d2[k], bary[i * 2 + 1], fmaf(d1[k], bary[i * 2 + 0], a[k]));
    const float rw0 = 1.0f / v0[3], rw1 = 1.0f / v1[3], rw2 = 1.0f / v2[3];
    raster_piece<kSharedZ>(true, snap(v0[0] * rw0 * vs), snap(v0[1] * rw0 * vs), snap(v1[0] * rw1 * vs),
                           snap(v1[1] * rw1 * vs), snap(v2[0] * rw2 * vs), snap(v2[1] * rw2 * vs), v0[2], v1[2], v2[2], rw0,
                           rw1, rw2, t, row0, nrows, zbuf);
    for (int k = 0; k < 4; ++k) v1[k] = v2[k];
  }
}

// Visibility of the triangles t = t_first + i * t_stride (i = 0, 1, ...) inside GL rows [row0, row0 + nrows)
// of one hypothesis: cudaraster's triangle setup (TriangleSetup.inl:181-391; fast path inline, clip path in
// raster_clipped) and raster_piece per snapped triangle.  Returns whether some triangle took the clip path.
template <bool kSharedZ>
__device__ __forceinline__ bool raster_triangles(const DeviceMesh& mesh, const float4* __restrict__ pc, int t_first,
                                                 int t_stride, int row0, int nrows, unsigned long long* zbuf) {
  constexpr int W = kCrop;
  const float vs = (float)(W << 3);
  bool clip_path = false;
  // Warp-uniform trip count and a single `live` flag instead of `continue`s: the lanes of a warp
  // reconverge at every iteration (with early exits the compiler let them drift apart and the loop
  // ran with ~6 active lanes).
  const int n_iter = (mesh.num_faces + t_stride - 1) / t_stride;
  for (int it = 0; it < n_iter; ++it) {
    const int t = it * t_stride + t_first;
    bool live = t < mesh.num_faces;
    int i0 = 0, i1 = 0, i2 = 0;
    if (live) {
      const int4 f = mesh.faces[t];
      i0 = f.x;
      i1 = f.y;
      i2 = f.z;
      live = (unsigned)i0 < (unsigned)mesh.num_vertices && (unsigned)i1 < (unsigned)mesh.num_vertices &&
             (unsigned)i2 < (unsigned)mesh.num_vertices;
    }
    float4 A = make_float4(0.f, 0.f, 0.f, 1.f), B = A, C = A;
    if (live) {
      A = pc[2 * i0];
      B = pc[2 * i1];
      C = pc[2 * i2];
    }
    if (live && ((A.w < fabsf(A.x)) | (A.w < fabsf(A.y)) | (A.w < fabsf(A.z)))) {
      if (((A.w < +A.x) & (B.w < +B.x) & (C.w < +C.x)) | ((A.w < -A.x) & (B.w < -B.x) & (C.w < -C.x)) |
          ((A.w < +A.y) & (B.w < +B.y) & (C.w < +C.y)) | ((A.w < -A.y) & (B.w < -B.y) & (C.w < -C.y)) |
          ((A.w < +A.z) & (B.w < +B.z) & (C.w < +C.z)) | ((A.w < -A.z) & (B.w < -B.z) & (C.w < -C.z)))
        live = false;
    }
    bool clipped = false;
    if (live && !((A.w >= fabsf(A.z)) & (B.w >= fabsf(B.z)) & (C.w >= fabsf(C.z)))) {
      clipped = true;  // crosses the near / far plane
      live = false;
    }
    const float rw0 = 1.0f / A.w, rw1 = 1.0f / B.w, rw2 = 1.0f / C.w;
    const int p0x = snap(A.x * rw0 * vs), p0y = snap(A.y * rw0 * vs);
    const int p1x = snap(B.x * rw1 * vs), p1y = snap(B.y * rw1 * vs);
    const int p2x = snap(C.x * rw2 * vs), p2y = snap(C.y * rw2 * vs);
    const int lo = min(min3(p0x, p1x, p2x), min3(p0y, p1y, p2y));
    const int hi = max(max3(p0x, p1x, p2x), max3(p0y, p1y, p2y));
    if (live && !(lo >= -32768 && hi <= 32767 && hi - lo <= (1 << 15) - 1)) {
      clipped = true;  // outside the S16 snap range / too large for the guard band
      live = false;
    }
    raster_piece<kSharedZ>(live, p0x, p0y, p1x, p1y, p2x, p2y, A.z, B.z, C.z, rw0, rw1, rw2, t, row0, nrows, zbuf);
    if (clipped) {
      // rare: the vertices are fetched again here (through a different load path, so that the compiler cannot
      // keep the 12 floats above alive across the pixel loop of the fast path)
      clip_path = true;
      const int4 f = __ldcg(mesh.faces + t);
      raster_clipped<kSharedZ>(__ldcg(pc + 2 * f.x), __ldcg(pc + 2 * f.y), __ldcg(pc + 2 * f.z), t, row0, nrows, zbuf);
    }
    __syncwarp();
  }
  return clip_path;
}

// One covered pixel (px, GL row py) of triangle `tri`: nvdiffrast's barycentrics / interpolation / texture
// (rasterize.cu:55-89, interpolate.cu:46-76, texture.cu:19-96), the reference's shading
// (foundationpose_render.cu:459-501), normalisation and packing (foundationpose_render.cu:61-140).  Returns the packed
// 8-channel fp16 pixel.
struct ShadeCtx {
  const float4* rec;  // vertex records of this hypothesis: [v][0] clip position, [v][1] camera-space position + diffuse
  float t_x, t_y, t_z, half_d;
};
__device__ __forceinline__ uint4 pack_pixel(const float* o) {
  uint4 pk;
  __half2* hp = reinterpret_cast<__half2*>(&pk);
  hp[0] = __floats2half2_rn(o[0], o[1]);
  hp[1] = __floats2half2_rn(o[2], o[3]);
  hp[2] = __floats2half2_rn(o[4], o[5]);
  hp[3] = __floats2half2_rn(0.f, 0.f);
  return pk;
}
__device__ __forceinline__ uint4 shade_one(const DeviceMesh& mesh, const ShadeCtx& c, int tri, int px, int py) {
  constexpr int W = kCrop, H = kCrop;
  const float xs = 2.0f / (float)W, xo = 1.0f / (float)W - 1.0f;
  const float ys = 2.0f / (float)H, yo = 1.0f / (float)H - 1.0f;
  const int4 f = mesh.faces[tri];
  const int i0 = f.x, i1 = f.y, i2 = f.z;
  const float4 p0 = c.rec[2 * i0], p1 = c.rec[2 * i1], p2 = c.rec[2 * i2];
  const float4 q0 = c.rec[2 * i0 + 1], q1 = c.rec[2 * i1 + 1], q2 = c.rec[2 * i2 + 1];
  const float2 t0 = *reinterpret_cast<const float2*>(mesh.texcoords + 2 * i0);
  const float2 t1 = *reinterpret_cast<const float2*>(mesh.texcoords + 2 * i1);
  const float2 t2 = *reinterpret_cast<const float2*>(mesh.texcoords + 2 * i2);
  const float fx = xs * (float)px + xo;
  const float fy = ys * (float)py + yo;
  const float p0x = p0.x - fx * p0.w, p0y = p0.y - fy * p0.w;
  const float p1x = p1.x - fx * p1.w, p1y = p1.y - fy * p1.w;
  const float p2x = p2.x - fx * p2.w, p2y = p2.y - fy * p2.w;
  const float a0 = p1x * p2y - p1y * p2x;
  const float a1 = p2x * p0y - p2y * p0x;
  const float a2 = p0x * p1y - p0y * p1x;
  const float iw = 1.f / ((a0 + a1) + a2);
  const float b0 = __saturatef(a0 * iw);
  const float b1 = __saturatef(a1 * iw);
  const float b2 = (1.f - b0) - b1;
  const float xyz0 = (b0 * q0.x + b1 * q1.x) + b2 * q2.x;
  const float xyz1 = (b0 * q0.y + b1 * q1.y) + b2 * q2.y;
  const float xyz2 = (b0 * q0.z + b1 * q1.z) + b2 * q2.z;
  const float dif = (b0 * q0.w + b1 * q1.w) + b2 * q2.w;
  float u = (b0 * t0.x + b1 * t1.x) + b2 * t2.x;
  float v = (b0 * t0.y + b1 * t1.y) + b2 * t2.y;
  // bilinear texture fetch, wrap
  const int tw = mesh.tex_w, th = mesh.tex_h;
  u = u - floorf(u);
  v = v - floorf(v);
  u = u * (float)tw - 0.5f;
  v = v * (float)th - 0.5f;
  int iu0 = (int)floorf(u), iv0 = (int)floorf(v);
  int iu1 = iu0 + 1, iv1 = iv0 + 1;
  const float fu = u - (float)iu0, fv = v - (float)iv0;
  if (iu0 < 0) iu0 += tw;
  if (iv0 < 0) iv0 += th;
  if (iu1 >= tw) iu1 -= tw;
  if (iv1 >= th) iv1 -= th;
  const float4 c00 = mesh.texture[(long long)iv0 * tw + iu0];
  const float4 c10 = mesh.texture[(long long)iv0 * tw + iu1];
  const float4 c01 = mesh.texture[(long long)iv1 * tw + iu0];
  const float4 c11 = mesh.texture[(long long)iv1 * tw + iu1];
  const float light = 0.8f + dif * 0.5f;
  float o[6];
  {
    const float k00[3] = {c00.x, c00.y, c00.z}, k10[3] = {c10.x, c10.y, c10.z};
    const float k01[3] = {c01.x, c01.y, c01.z}, k11[3] = {c11.x, c11.y, c11.z};
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const float top = k00[ch] + fu * (k10[ch] - k00[ch]);
      const float bot = k01[ch] + fu * (k11[ch] - k01[ch]);
      const float col = top + fv * (bot - top);
      o[ch] = fminf(fmaxf(col * light, 0.f), 1.f);
    }
  }
  // normalise (threshold_and_downscale_pointcloud_kernel)
  const bool invalid = xyz2 < kMinDepth;
  const float qx = (xyz0 - c.t_x) / c.half_d, qy = (xyz1 - c.t_y) / c.half_d, qz = (xyz2 - c.t_z) / c.half_d;
  o[3] = (invalid || fabsf(qx) > kMaxDepth) ? 0.f : qx;
  o[4] = (invalid || fabsf(qy) > kMaxDepth) ? 0.f : qy;
  o[5] = (invalid || fabsf(qz) > kMaxDepth) ? 0.f : qz;
  return pack_pixel(o);
}
__device__ __forceinline__ ShadeCtx make_shade_ctx(const DeviceMesh& mesh, const PoseAux* __restrict__ aux_all,
                                                   const float4* pc, int pose) {
  ShadeCtx c;
  c.rec = pc;
  c.t_x = aux_all[pose].t[0];
  c.t_y = aux_all[pose].t[1];
  c.t_z = aux_all[pose].t[2];
  c.half_d = mesh.diameter / 2.f;
  return c;
}

// Pixel i = (GL row row0 + i / 160, column i % 160) of the packed image of `pose`, vertically flipped
// (foundationpose_render.cpp:676-680).
__device__ __forceinline__ uint4* packed_pixel(__half* img, int py, int px) {
  return reinterpret_cast<uint4*>(img + ((long long)(kCrop - 1 - py) * kStemPitch + kStemPad + px) * kPackC);
}

// Split path: every thread shades its pixels of GL rows [row0, row0 + nrows) straight from the global z-buffer.
__device__ __forceinline__ void shade_pixels(const DeviceMesh& mesh, const PoseAux* __restrict__ aux_all,
                                             const float4* __restrict__ pc, __half* __restrict__ packed, int img0,
                                             int32_t* __restrict__ rast_id_debug, int pose, int row0, int nrows,
                                             const unsigned long long* zbuf) {
  constexpr int W = kCrop, H = kCrop;
  const ShadeCtx c = make_shade_ctx(mesh, aux_all, pc, pose);
  __half* img = packed + (long long)(img0 + pose) * kPackImageElems;
  for (int i = threadIdx.x; i < W * nrows; i += blockDim.x) {
    const int ly = i / W, px = i - ly * W;
    const int py = row0 + ly;
    const unsigned long long key = zbuf[i];
    const int tri = key == ~0ull ? -1 : (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFu));
    if (rast_id_debug) rast_id_debug[(long long)pose * W * H + py * W + px] = tri + 1;
    *packed_pixel(img, py, px) = tri >= 0 ? shade_one(mesh, c, tri, px, py) : make_uint4(0u, 0u, 0u, 0u);
  }
}

// Band path (shared-memory z-buffer): covered pixels are COMPACTED first, so that the long dependent-load chain of
// shade_one (z-buffer -> face -> vertices -> texels) is spread evenly over all threads: with a typical 26 % of the crop
// covered, a thread shades ~6 pixels instead of up to 25 while most warps only wrote zeros.  The list of covered pixel
// indices is stored in the UPPER halves of the first `count` z-buffer entries (the depth words, dead after
// rasterisation; the triangle ids in the lower halves stay in place), so it needs no memory of its own.
template <int kThreads, int kRows>
__device__ __forceinline__ void shade_band_compacted(const DeviceMesh& mesh, const PoseAux* __restrict__ aux_all,
                                                     const float4* __restrict__ pc, __half* __restrict__ packed, int img0,
                                                     int32_t* __restrict__ rast_id_debug, int pose, int row0,
                                                     unsigned long long* zbuf, int* count) {
  constexpr int W = kCrop, H = kCrop, kPixels = W * kRows;
  static_assert(kPixels % kThreads == 0 && kPixels / kThreads <= 32, "one coverage bit per owned pixel");
  constexpr int kIter = kPixels / kThreads;
  __half* img = packed + (long long)(img0 + pose) * kPackImageElems;
  // A: coverage bits of the pixels this thread owns; empty pixels are written (zeros) right away
  unsigned mask = 0;
#pragma unroll 5
  for (int it = 0; it < kIter; ++it) {
    const int i = it * kThreads + threadIdx.x;
    const unsigned long long key = zbuf[i];
    const bool covered = key != ~0ull;
    const int ly = i / W, px = i - ly * W;
    if (rast_id_debug)
      rast_id_debug[(long long)pose * W * H + (row0 + ly) * W + px] = covered ? (int)(0xFFFFFFFFu - (unsigned)key) + 1 : 0;
    if (covered) mask |= 1u << it;
    else *packed_pixel(img, row0 + ly, px) = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();  // every key has been read: the depth words may be overwritten
  // B: append the covered pixel indices (warp-aggregated)
  unsigned* words = reinterpret_cast<unsigned*>(zbuf);
  const unsigned lane = threadIdx.x & 31u;
  for (int it = 0; it < kIter; ++it) {
    const bool covered = (mask >> it) & 1u;
    const unsigned vote = __ballot_sync(0xffffffffu, covered);
    if (vote == 0u) continue;
    int base = 0;
    if (lane == 0) base = atomicAdd(count, __popc(vote));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (covered) words[2 * (base + __popc(vote & ((1u << lane) - 1u))) + 1] = (unsigned)(it * kThreads + threadIdx.x);
  }
  __syncthreads();
  // C: shade the list
  const int n_covered = *count;
  const ShadeCtx c = make_shade_ctx(mesh, aux_all, pc, pose);
  for (int j = threadIdx.x; j < n_covered; j += kThreads) {
    const int i = (int)words[2 * j + 1];
    const int tri = (int)(0xFFFFFFFFu - words[2 * i]);
    const int ly = i / W, px = i - ly * W;
    *packed_pixel(img, row0 + ly, px) = shade_one(mesh, c, tri, px, row0 + ly);
  }
}

// One CTA per (band of kRows rows, hypothesis): every thread takes triangles and resolves visibility on a
// shared-memory z-buffer (raster_triangles), then the band is shaded (shade_band_compacted).
template <int kThreads, int kRows>
__global__ void __launch_bounds__(kThreads, 1024 / kThreads)
render_kernel(DeviceMesh mesh, const PoseAux* __restrict__ aux_all, const float4* __restrict__ vrec,
              __half* __restrict__ packed, int img0,
              int32_t* __restrict__ rast_id_debug, int* __restrict__ status, int experiment) {
  extern __shared__ unsigned long long zbuf[];  // [kRows * 160]
  __shared__ int covered_count;
  const int pose = blockIdx.y;
  const int row0 = blockIdx.x * kRows;  // GL rows [row0, row0 + kRows)
  for (int i = threadIdx.x; i < kCrop * kRows; i += kThreads) zbuf[i] = ~0ull;
  if (threadIdx.x == 0) covered_count = 0;
  __syncthreads();
  const float4* pc = vrec + 2 * (long long)pose * mesh.num_vertices;
  if (experiment != 1)
    if (raster_triangles<true>(mesh, pc, threadIdx.x, kThreads, row0, kRows, zbuf)) atomicOr(status, kStatusClipPath);
  __syncthreads();
  if (experiment != 2)
    shade_band_compacted<kThreads, kRows>(mesh, aux_all, pc, packed, img0, rast_id_debug, pose, row0, zbuf,
                                          &covered_count);
}

// Few hypotheses (Track: one): a (band, hypothesis) grid would leave most SMs idle while 4 CTAs walk all
// triangles, so the triangles are spread over blockIdx.x CTAs as well, on a global z-buffer of the whole
// crop (pre-set to ~0), and the crop is shaded by a second kernel.  Same keys, same atomicMin: the
// visibility result is identical to render_kernel's.
__global__ void fill_u64_kernel(unsigned long long* p, long long n, unsigned long long v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void __launch_bounds__(kRenderThreads)
raster_split_kernel(DeviceMesh mesh, const float4* __restrict__ vrec, unsigned long long* __restrict__ zbuf_all,
                    int* __restrict__ status) {
  const int pose = blockIdx.y;
  const float4* pc = vrec + 2 * (long long)pose * mesh.num_vertices;
  if (raster_triangles<false>(mesh, pc, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x, 0, kCrop,
                              zbuf_all + (long long)pose * kCrop * kCrop))
    atomicOr(status, kStatusClipPath);
}
__global__ void __launch_bounds__(kRenderThreads)
shade_split_kernel(DeviceMesh mesh, const PoseAux* __restrict__ aux_all, const float4* __restrict__ vrec,
                   __half* __restrict__ packed, int img0,
                   int32_t* __restrict__ rast_id_debug, const unsigned long long* __restrict__ zbuf_all, int rows) {
  const int pose = blockIdx.y;
  const int row0 = blockIdx.x * rows;
  shade_pixels(mesh, aux_all, vrec + 2 * (long long)pose * mesh.num_vertices, packed, img0, rast_id_debug, pose,
               row0, rows, zbuf_all + (long long)pose * kCrop * kCrop + (long long)row0 * kCrop);
}

// ------------------------------------------------------------------ observed crop
// TransfProcess, foundationpose_render.cpp:731-812: cvcuda::WarpPerspective (rgb bilinear -> u8,
// xyz nearest, constant-0 border) + ConvertTo(1/255) + normalise + pack.
__global__ void warp_kernel(const uint8_t* __restrict__ rgb, const float* __restrict__ xyz, int h, int w,
                            const PoseAux* __restrict__ aux_all, float half_d, __half* __restrict__ packed,
                            int img0) {
  const int pose = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kCrop * kCrop) return;
  const PoseAux& a = aux_all[pose];
  const int y = i / kCrop, x = i - y * kCrop;
  const float sx = a.inv_sx * (float)x + a.left;
  const float sy = a.inv_sy * (float)y + a.top;
  float o[6];
  {
    const float fx0 = floorf(sx), fy0 = floorf(sy);
    const int x0 = (int)fx0, y0 = (int)fy0;
    const float fx = sx - fx0, fy = sy - fy0;
    const float w00 = (1.f - fx) * (1.f - fy), w10 = fx * (1.f - fy), w01 = (1.f - fx) * fy, w11 = fx * fy;
    float acc[3] = {0.f, 0.f, 0.f};
    auto tap = [&](int yy, int xx, float wgt) {
      const bool ok = xx >= 0 && xx < w && yy >= 0 && yy < h;
      const uint8_t* p = rgb + 3 * ((long long)(ok ? yy : 0) * w + (ok ? xx : 0));
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[c] = acc[c] + (ok ? (float)p[c] : 0.f) * wgt;
    };
    tap(y0, x0, w00);
    tap(y0, x0 + 1, w10);
    tap(y0 + 1, x0, w01);
    tap(y0 + 1, x0 + 1, w11);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float q = fminf(fmaxf(rintf(acc[c]), 0.f), 255.f);
      o[c] = q * (1.0f / 255.0f);
    }
  }
  {
    const int xn = (int)floorf(sx + 0.5f), yn = (int)floorf(sy + 0.5f);
    float p[3] = {0.f, 0.f, 0.f};
    if (xn >= 0 && xn < w && yn >= 0 && yn < h) {
      const float* s = xyz + 3 * ((long long)yn * w + xn);
      p[0] = s[0]; p[1] = s[1]; p[2] = s[2];
    }
    const bool invalid = p[2] < kMinDepth;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float q = (p[r] - a.t[r]) / half_d;
      o[3 + r] = (invalid || fabsf(q) > kMaxDepth) ? 0.f : q;
    }
  }
  uint4 pk;
  __half2* hp = reinterpret_cast<__half2*>(&pk);
  hp[0] = __floats2half2_rn(o[0], o[1]);
  hp[1] = __floats2half2_rn(o[2], o[3]);
  hp[2] = __floats2half2_rn(o[4], o[5]);
  hp[3] = __floats2half2_rn(0.f, 0.f);
  __half* img = packed + (long long)(img0 + pose) * kPackImageElems;
  *reinterpret_cast<uint4*>(img + ((long long)y * kStemPitch + kStemPad + x) * kPackC) = pk;
}

__global__ void repack_f32_kernel(const float* __restrict__ blob, long long npix, __half* __restrict__ packed,
                                  int img0) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const long long img = i / (kCrop * kCrop);
  const int r = (int)(i - img * (kCrop * kCrop));
  const int y = r / kCrop, x = r - y * kCrop;
  const float* src = blob + i * 6;
  uint4 pk;
  __half2* hp = reinterpret_cast<__half2*>(&pk);
  hp[0] = __floats2half2_rn(src[0], src[1]);
  hp[1] = __floats2half2_rn(src[2], src[3]);
  hp[2] = __floats2half2_rn(src[4], src[5]);
  hp[3] = __floats2half2_rn(0.f, 0.f);
  *reinterpret_cast<uint4*>(packed + (img0 + img) * kPackImageElems + ((long long)y * kStemPitch + kStemPad + x) * kPackC) = pk;
}

}  // namespace

void launch_repack_f32(const float* blob, int n, __half* packed, int img0, cudaStream_t s) {
  if (n <= 0) return;
  const long long npix = (long long)n * kCrop * kCrop;
  repack_f32_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, s>>>(blob, npix, packed, img0);
}

// ---- mesh diameter --------------------------------------------------------------------------------------
// Tiles of 256 x 256 vertex pairs (upper triangle only): the j tile is staged in shared memory, each thread keeps one
// i vertex in registers.  dx*dx + dy*dy + dz*dz in aiVector3D::Length's order (no contraction in this file).
namespace {
constexpr int kDiamTile = 256;
__global__ void __launch_bounds__(kDiamTile) mesh_diameter_kernel(const float* __restrict__ v, int n, unsigned* out_sq) {
  // linear block index -> (ti, tj) with tj >= ti
  const int nt = (n + kDiamTile - 1) / kDiamTile;
  int ti = 0, rem = blockIdx.x;
  while (rem >= nt - ti) { rem -= nt - ti; ++ti; }
  const int tj = ti + rem;
  __shared__ float sx[kDiamTile], sy[kDiamTile], sz[kDiamTile];
  const int j0 = tj * kDiamTile, jn = min(kDiamTile, n - j0);
  if ((int)threadIdx.x < jn) {
    sx[threadIdx.x] = v[3 * (j0 + threadIdx.x)];
    sy[threadIdx.x] = v[3 * (j0 + threadIdx.x) + 1];
    sz[threadIdx.x] = v[3 * (j0 + threadIdx.x) + 2];
  }
  __syncthreads();
  const int i = ti * kDiamTile + threadIdx.x;
  float best = 0.f;
  if (i < n) {
    const float x = v[3 * i], y = v[3 * i + 1], z = v[3 * i + 2];
    for (int j = 0; j < jn; ++j) {
      const float dx = x - sx[j], dy = y - sy[j], dz = z - sz[j];
      best = fmaxf(best, dx * dx + dy * dy + dz * dz);
    }
  }
  for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
  if ((threadIdx.x & 31) == 0 && best > 0.f) atomicMax(out_sq, __float_as_uint(best));
}
}  // namespace

void launch_mesh_diameter(const float* vertices, int n, unsigned* out_sq, cudaStream_t s) {
  const long long nt = (n + kDiamTile - 1) / kDiamTile;
  const long long blocks = nt * (nt + 1) / 2;
  mesh_diameter_kernel<<<(unsigned)blocks, kDiamTile, 0, s>>>(vertices, n, out_sq);
}

const char* preprocess_init() {
  cudaError_t e = cudaFuncSetAttribute(render_kernel<512, kCrop / 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kCrop * (kCrop / 2) * 8);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(render_kernel<1024, kCrop>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCrop * kCrop * 8);
  if (e != cudaSuccess) return "cudaFuncSetAttribute(render_kernel smem) failed";
  e = preload_kernels(depth_to_xyz_kernel, depth_filter_kernel, mask_scan_kernel,
                      median_select_kernel, pose_setup_kernel, vertex_stage_kernel, fill_u64_kernel,
                      raster_split_kernel, shade_split_kernel, warp_kernel, repack_f32_kernel, mesh_diameter_kernel);
  return e == cudaSuccess ? nullptr : "loading the preprocessing kernels failed (not an sm_100a device?)";
}

void launch_depth_to_xyz(const float* depth, float* xyz, int h, int w, const Intrinsics& K, cudaStream_t s) {
  const int n = h * w;
  depth_to_xyz_kernel<<<(n + 255) / 256, 256, 0, s>>>(depth, xyz, h, w, K.fx, K.fy, K.cx, K.cy);
}

void launch_depth_filters(const float* depth, float* eroded, float* filtered, int h, int w, int* sampler_meta,
                          cudaStream_t s) {
  dim3 block(kFiltTX, kFiltTY), grid((w + kFiltTX - 1) / kFiltTX, (h + kFiltTY - 1) / kFiltTY);
  depth_filter_kernel<<<grid, block, 0, s>>>(depth, eroded, filtered, h, w, sampler_meta);
}

void launch_guess_translation(const float* filtered, const uint8_t* mask, int h, int w, const Intrinsics& K,
                              float* center, int* status, unsigned* scratch, const float* grid, float* poses,
                              int n_poses, cudaStream_t s) {
  // scratch: [8] ints of meta (reset by launch_depth_filters) followed by h*w list entries
  int* meta = reinterpret_cast<int*>(scratch);
  unsigned* list = scratch + 8;
  const int blocks = (h * w + 1023) / 1024;
  mask_scan_kernel<<<blocks, 256, 0, s>>>(filtered, mask, h, w, list, meta);
  median_select_kernel<<<1, 1024, 0, s>>>(list, meta, K, center, status, grid, poses, n_poses);
}

void launch_pose_setup(const float* poses, int n, const Intrinsics& K, float diameter, float crop_ratio,
                       PoseAux* aux, cudaStream_t s) {
  pose_setup_kernel<<<(n + 63) / 64, 64, 0, s>>>(poses, n, K, diameter, crop_ratio, aux);
}

int launch_render(const DeviceMesh& mesh, const PoseAux* aux, int n, float4* vrec, __half* packed, int img0,
                  int32_t* rast_id_debug, int* status, unsigned long long* zbuf_scratch, int zbuf_poses,
                  int num_sms, int split_mode, cudaStream_t s) {
  if (n <= 0) return 0;
  const int sms = num_sms > 0 ? num_sms : 148;
  // band kernel: 2 CTAs per hypothesis and 2 resident CTAs per SM; below ~2 CTAs per SM it is bound by the
  // latency of one CTA walking every triangle (N = 1 and N = 32 take the same ~105 us)
  bool split = n * 4 <= 2 * sms;
  if (split_mode >= 0) split = split_mode != 0;  // tests run both paths on the same input
  const bool use_split = split && zbuf_scratch && n <= zbuf_poses;
  vertex_stage_kernel<<<dim3((mesh.num_vertices + 255) / 256, n), 256, 0, s>>>(mesh, aux, vrec,
                                                                               use_split ? zbuf_scratch : nullptr);
  if (use_split) {
    // few hypotheses: spread the triangles over the machine (see raster_split_kernel); the z-buffers were reset by
    // the vertex stage
    // one or two triangles per thread: the kernel is a chain of dependent gathers, so its duration is the number of
    // triangles a thread walks (FPB200_SPLIT_SLICES: measurements only)
    static const int force_slices = getenv("FPB200_SPLIT_SLICES") ? atoi(getenv("FPB200_SPLIT_SLICES")) : 0;
    const int want = (32 * sms + n - 1) / n, have = (mesh.num_faces + kRenderThreads - 1) / kRenderThreads;
    const int slices = force_slices ? (force_slices < have ? force_slices : have) : (want < have ? want : have);
    raster_split_kernel<<<dim3(slices, n), kRenderThreads, 0, s>>>(mesh, vrec, zbuf_scratch, status);
    static const int shade_rows = getenv("FPB200_SPLIT_SHADE_ROWS") ? atoi(getenv("FPB200_SPLIT_SHADE_ROWS")) : 2;
    shade_split_kernel<<<dim3(kCrop / shade_rows, n), kRenderThreads, 0, s>>>(mesh, aux, vrec, packed, img0,
                                                                             rast_id_debug, zbuf_scratch, shade_rows);
    return 2;
  }
  static const int experiment = getenv("FPB200_RENDER_EXPERIMENT") ? atoi(getenv("FPB200_RENDER_EXPERIMENT")) : 0;
  if (experiment == 3) return 1;
  // more hypotheses than SMs: one 160-row band per CTA (each triangle is set up once; 1.7 waves of 252 CTAs);
  // fewer: two 80-row bands, twice the CTAs
  static const int force_bands = getenv("FPB200_RENDER_BANDS") ? atoi(getenv("FPB200_RENDER_BANDS")) : 0;
  const int bands = force_bands ? force_bands : (n > sms ? 1 : 2);
  if (bands == 1)
    render_kernel<1024, kCrop><<<dim3(1, n), 1024, kCrop * kCrop * 8, s>>>(mesh, aux, vrec, packed, img0,
                                                                           rast_id_debug, status, experiment);
  else
    render_kernel<512, kCrop / 2><<<dim3(2, n), 512, kCrop * (kCrop / 2) * 8, s>>>(mesh, aux, vrec, packed, img0,
                                                                                   rast_id_debug, status, experiment);
  return 2;
}

void launch_warp(const uint8_t* rgb, const float* xyz, int h, int w, const PoseAux* aux, int n, float diameter,
                 __half* packed, int img0, cudaStream_t s) {
  if (n <= 0) return;
  dim3 grid((kCrop * kCrop + 255) / 256, n);
  warp_kernel<<<grid, 256, 0, s>>>(rgb, xyz, h, w, aux, diameter / 2.f, packed, img0);
}

}  // namespace fpb
