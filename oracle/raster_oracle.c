/* ORACLE — test infrastructure only.  Nothing under foundationpose_cpp_b200/ may call this.
 *
 * Sequential CPU restatement of the visibility pass the reference obtains from its vendored
 * nvdiffrast / cudaraster (paths relative to
 * detection_6d_foundationpose/src/nvdiffrast/common/cudaraster/impl/):
 *   frustum cull + fixed-point snapping ... TriangleSetup.inl:11-24, 255-301
 *   degenerate test, vertex swap .......... TriangleSetup.inl:42-54, 120-134
 *   integer depth plane ................... TriangleSetup.inl:136-157, Util.inl:184-210
 *   coverage (edge functions, fill rule) .. Util.inl:304-309 (exact, exclusive rule), FineRaster.inl:75-101
 *   depth test / tie rule ................. FineRaster.inl:152-172, 330-351 (nearest wins; on equal
 *                                           depth the later triangle wins)
 *   polygon clipper ....................... TriangleSetup.inl:303-390, Util.inl:101-160: triangles that
 *                                           cross the near/far plane or leave the S16 snap range are
 *                                           clipped against the frustum planes in barycentric space
 *                                           and fan-triangulated; every piece keeps the triangle's id.
 *                                           Float contraction as nvcc compiled it for the reference
 *                                           (read off the SASS of oracle/_ref): a + b*c + d*e is
 *                                           fma(d, e, fma(b, c, a)); x*ac + y*bc is fma(x, ac, y*bc).
 * *n_clip_path counts the triangles that took the clipper.
 *
 * Parity status: pinned on the GPU box against the reference's own kernels compiled from
 * /root/reference (oracle/_ref, tests/test_ref_raster.py); no golden vectors exist upstream.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SUBPIXEL_LOG2 4
#define MAXVIEWPORT_LOG2 11
#define LERP_ERROR(s) (2200u << (s))
#define DEPTH_MIN LERP_ERROR(3)
#define DEPTH_MAX (0xFFFFFFFFu - LERP_ERROR(3))

static inline int32_t f32_to_s32_sat_rni(float a) {
  if (isnan(a)) return 0;
  if (a >= 2147483648.0f) return INT32_MAX;
  if (a <= -2147483648.0f) return INT32_MIN;
  return (int32_t)lrintf(a); /* default rounding mode: nearest even, like cvt.rni */
}

static inline uint32_t f32_bits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}

static inline uint32_t f32_to_u32_trunc(float a) {
  if (!(a > 0.f)) return 0;
  if (a >= 4294967296.0f) return 0xFFFFFFFFu;
  return (uint32_t)a;
}

static inline int imin3(int a, int b, int c) { return a < b ? (a < c ? a : c) : (b < c ? b : c); }
static inline int imax3(int a, int b, int c) { return a > b ? (a > c ? a : c) : (b > c ? b : c); }

/* Util.inl:184-210 */
static void setup_pleq(const float v[3], int v0x, int v0y, int d1x, int d1y, int d2x, int d2y,
                       float area_rcp, uint32_t pleq[3]) {
  float mx = fmaxf(fmaxf(v[0], v[1]), v[2]);
  int sh = (int)(f32_bits(mx) >> 23) - (127 + 22);
  if (sh < 0) sh = 0;
  if (sh > 8) sh = 8;
  int32_t t0 = (int32_t)(f32_to_u32_trunc(v[0]) >> sh);
  int32_t t1 = (int32_t)(f32_to_u32_trunc(v[1]) >> sh) - t0;
  int32_t t2 = (int32_t)(f32_to_u32_trunc(v[2]) >> sh) - t0;

  uint32_t rcp_mant = (f32_bits(area_rcp) & 0x007FFFFFu) | 0x00800000u;
  int rcp_shift = (23 + 127) - (int)(f32_bits(area_rcp) >> 23);

  /* two's-complement wrap like the GPU's 64-bit integer multiply */
  int64_t xc = (int64_t)((uint64_t)((int64_t)t1 * d2y - (int64_t)t2 * d1y) * (uint64_t)rcp_mant);
  int64_t yc = (int64_t)((uint64_t)((int64_t)t2 * d1x - (int64_t)t1 * d2x) * (uint64_t)rcp_mant);
  pleq[0] = (uint32_t)(xc >> (rcp_shift - (sh + SUBPIXEL_LOG2)));
  pleq[1] = (uint32_t)(yc >> (rcp_shift - (sh + SUBPIXEL_LOG2)));

  int32_t cx = (v0x * 2 + imin3(d1x, d2x, 0) + imax3(d1x, d2x, 0)) >> (SUBPIXEL_LOG2 + 1);
  int32_t cy = (v0y * 2 + imin3(d1y, d2y, 0) + imax3(d1y, d2y, 0)) >> (SUBPIXEL_LOG2 + 1);
  int32_t vcx = v0x - cx * (1 << SUBPIXEL_LOG2);
  int32_t vcy = v0y - cy * (1 << SUBPIXEL_LOG2);

  uint32_t z = (uint32_t)t0 << sh;
  int64_t corr = (int64_t)((uint64_t)(xc >> 13) * (uint64_t)(int64_t)vcx + (uint64_t)(yc >> 13) * (uint64_t)(int64_t)vcy);
  z -= (uint32_t)(corr >> (rcp_shift - (sh + 13)));
  z -= pleq[0] * (uint32_t)cx + pleq[1] * (uint32_t)cy;
  pleq[2] = z;
}

/* One snapped (sub)triangle: setupTriangle's depth plane, exact coverage, depth test (id = t + 1). */
static void raster_piece(int p0x, int p0y, int p1x, int p1y, int p2x, int p2y, float z0, float z1, float z2,
                         float rw0, float rw1, float rw2, int t, int W, int H, int bias_x, int bias_y,
                         int32_t* out_id, uint32_t* zbuf) {
  /* prepareTriangle (only the culls that change the image) */
  int d1x = p1x - p0x, d1y = p1y - p0y, d2x = p2x - p0x, d2y = p2y - p0y;
  int32_t area = d1x * d2y - d1y * d2x;
  if (area == 0) return;
  if (area < 0) { /* setupTriangle: swap vertices 1 and 2 */
    int tx;
    float tf;
    tx = d1x; d1x = d2x; d2x = tx;
    tx = d1y; d1y = d2y; d2y = tx;
    tx = p1x; p1x = p2x; p2x = tx;
    tx = p1y; p1y = p2y; p2y = tx;
    tf = z1; z1 = z2; z2 = tf;
    tf = rw1; rw1 = rw2; rw2 = tf;
    area = -area;
  }
  const float zcoef = (float)(DEPTH_MAX - DEPTH_MIN) * 0.5f;
  const float zbias = (float)(DEPTH_MAX + DEPTH_MIN) * 0.5f;
  /* the reference is compiled with nvcc's default fp contraction: (z*zcoef)*rcpW + zbias is one FMA */
  float zv[3] = {fmaf(z0 * zcoef, rw0, zbias), fmaf(z1 * zcoef, rw1, zbias), fmaf(z2 * zcoef, rw2, zbias)};
  int wv0x = p0x + (W << (SUBPIXEL_LOG2 - 1)), wv0y = p0y + (H << (SUBPIXEL_LOG2 - 1));
  uint32_t pleq[3];
  setup_pleq(zv, wv0x - (1 << (SUBPIXEL_LOG2 - 1)), wv0y - (1 << (SUBPIXEL_LOG2 - 1)), d1x, d1y, d2x, d2y,
             1.0f / (float)area, pleq);

  /* pixel bounding box of the snapped triangle */
  int lo_x = imin3(p0x, p1x, p2x) + bias_x, hi_x = imax3(p0x, p1x, p2x) + bias_x;
  int lo_y = imin3(p0y, p1y, p2y) + bias_y, hi_y = imax3(p0y, p1y, p2y) + bias_y;
  int px0 = (lo_x + 15) >> 4, px1 = hi_x >> 4, py0 = (lo_y + 15) >> 4, py1 = hi_y >> 4;
  if (px0 < 0) px0 = 0;
  if (py0 < 0) py0 = 0;
  if (px1 > W - 1) px1 = W - 1;
  if (py1 > H - 1) py1 = H - 1;
  /* edges: (origin, direction) = (p0, p1-p0), (p1, p2-p1), (p0, p0-p2) */
  const int ex[3] = {d1x, d2x - d1x, -d2x}, ey[3] = {d1y, d2y - d1y, -d2y};
  const int ox[3] = {p0x, p1x, p0x}, oy[3] = {p0y, p1y, p0y};
  for (int py = py0; py <= py1; ++py) {
    int sy = py * 16 - bias_y;
    for (int px = px0; px <= px1; ++px) {
      int sx = px * 16 - bias_x;
      int inside = 1;
      for (int e = 0; e < 3; ++e) {
        int64_t v = (int64_t)(ox[e] - sx) * ey[e] - (int64_t)(oy[e] - sy) * ex[e];
        if (ey[e] > 0 || (ey[e] == 0 && ex[e] <= 0)) v -= 1; /* exclusive edge */
        if (v < 0) { inside = 0; break; }
      }
      if (!inside) continue;
      uint32_t depth = pleq[0] * (uint32_t)px + pleq[1] * (uint32_t)py + pleq[2];
      int idx = py * W + px;
      if (depth <= zbuf[idx]) { /* FineRaster.inl:338-343: killed only when depth > old */
        zbuf[idx] = depth;
        out_id[idx] = t + 1;
      }
    }
  }
}

/* Util.inl:101-130.  Returns the number of output vertices. */
static int clip_poly_plane(float* out, const float* in, int num_in, float v0, float v1, float v2) {
  int num_out = 0;
  if (num_in >= 3) {
    int ai = (num_in - 1) * 2;
    float av = fmaf(v2, in[ai + 1], fmaf(v1, in[ai + 0], v0));
    for (int bi = 0; bi < num_in * 2; bi += 2) {
      float bv = fmaf(v2, in[bi + 1], fmaf(v1, in[bi + 0], v0));
      if (av * bv < 0.0f) {
        float bc = av / (av - bv);
        float ac = 1.0f - bc;
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

/* Util.inl:134-160 */
static int clip_triangle_frustum(float* bary, const float* v0, const float* v1, const float* v2, const float* d1,
                                 const float* d2) {
  int num = 3;
  float temp[18];
  bary[0] = 0.0f, bary[1] = 0.0f;
  bary[2] = 1.0f, bary[3] = 0.0f;
  bary[4] = 0.0f, bary[5] = 1.0f;
  for (int ax = 0; ax < 3; ++ax) {
    if ((v0[3] < fabsf(v0[ax])) | (v1[3] < fabsf(v1[ax])) | (v2[3] < fabsf(v2[ax]))) {
      num = clip_poly_plane(temp, bary, num, v0[3] + v0[ax], d1[3] + d1[ax], d2[3] + d2[ax]);
      num = clip_poly_plane(bary, temp, num, v0[3] - v0[ax], d1[3] - d1[ax], d2[3] - d2[ax]);
    }
  }
  return num;
}

/* TriangleSetup.inl:311-390: clip to the frustum, fan-triangulate, snap and rasterise every piece. */
static void raster_clipped(const float* a, const float* b, const float* c, int t, int W, int H, float vsx, float vsy,
                           int bias_x, int bias_y, int32_t* out_id, uint32_t* zbuf) {
  float bary[18];
  const float d1[4] = {b[0] - a[0], b[1] - a[1], b[2] - a[2], b[3] - a[3]};
  const float d2[4] = {c[0] - a[0], c[1] - a[1], c[2] - a[2], c[3] - a[3]};
  const int nv = clip_triangle_frustum(bary, a, b, c, d1, d2);
  if (nv < 3) return;
  float v0[4], v1[4], v2[4];
  for (int k = 0; k < 4; ++k) {
    v0[k] = fmaf(d2[k], bary[1], fmaf(d1[k], bary[0], a[k]));
    v1[k] = fmaf(d2[k], bary[3], fmaf(d1[k], bary[2], a[k]));
  }
  for (int i = 2; i < nv; ++i) {
    for (int k = 0; k < 4; ++k) v2[k] = fmaf(d2[k], bary[i * 2 + 1], fmaf(d1[k], bary[i * 2 + 0], a[k]));
    const float rw0 = 1.0f / v0[3], rw1 = 1.0f / v1[3], rw2 = 1.0f / v2[3];
    const int p0x = f32_to_s32_sat_rni(v0[0] * rw0 * vsx), p0y = f32_to_s32_sat_rni(v0[1] * rw0 * vsy);
    const int p1x = f32_to_s32_sat_rni(v1[0] * rw1 * vsx), p1y = f32_to_s32_sat_rni(v1[1] * rw1 * vsy);
    const int p2x = f32_to_s32_sat_rni(v2[0] * rw2 * vsx), p2y = f32_to_s32_sat_rni(v2[1] * rw2 * vsy);
    raster_piece(p0x, p0y, p1x, p1y, p2x, p2y, v0[2], v1[2], v2[2], rw0, rw1, rw2, t, W, H, bias_x, bias_y, out_id,
                 zbuf);
    for (int k = 0; k < 4; ++k) v1[k] = v2[k];
  }
}

/* clip: [V][4] clip-space positions; tri: [T][3]; out_id: [H*W] triangle index + 1 (0 = empty),
 * row 0 = bottom row in GL convention (what cudaraster writes); out_depth: [H*W] or NULL. */
int fpo_rasterize(const float* clip, int V, const int32_t* tri, int T, int W, int H,
                  int32_t* out_id, uint32_t* out_depth, int* n_clip_path) {
  uint32_t* zbuf = out_depth ? out_depth : (uint32_t*)malloc((size_t)W * H * sizeof(uint32_t));
  if (!zbuf) return -1;
  for (int i = 0; i < W * H; ++i) {
    out_id[i] = 0;
    zbuf[i] = DEPTH_MAX;
  }
  int clip_path = 0;
  const float vsx = (float)(W << (SUBPIXEL_LOG2 - 1));
  const float vsy = (float)(H << (SUBPIXEL_LOG2 - 1));
  const int bias_x = (W - 1) << (SUBPIXEL_LOG2 - 1); /* sample x of pixel px = px*16 - bias_x */
  const int bias_y = (H - 1) << (SUBPIXEL_LOG2 - 1);

  for (int t = 0; t < T; ++t) {
    int i0 = tri[3 * t], i1 = tri[3 * t + 1], i2 = tri[3 * t + 2];
    if ((unsigned)i0 >= (unsigned)V || (unsigned)i1 >= (unsigned)V || (unsigned)i2 >= (unsigned)V)
      continue;
    const float* a = clip + 4 * i0;
    const float* b = clip + 4 * i1;
    const float* c = clip + 4 * i2;
    /* TriangleSetup.inl:255-268 */
    if ((a[3] < fabsf(a[0])) | (a[3] < fabsf(a[1])) | (a[3] < fabsf(a[2]))) {
      if (((a[3] < +a[0]) & (b[3] < +b[0]) & (c[3] < +c[0])) |
          ((a[3] < -a[0]) & (b[3] < -b[0]) & (c[3] < -c[0])) |
          ((a[3] < +a[1]) & (b[3] < +b[1]) & (c[3] < +c[1])) |
          ((a[3] < -a[1]) & (b[3] < -b[1]) & (c[3] < -c[1])) |
          ((a[3] < +a[2]) & (b[3] < +b[2]) & (c[3] < +c[2])) |
          ((a[3] < -a[2]) & (b[3] < -b[2]) & (c[3] < -c[2])))
        continue;
    }
    if (!((a[3] >= fabsf(a[2])) & (b[3] >= fabsf(b[2])) & (c[3] >= fabsf(c[2])))) {
      ++clip_path;
      raster_clipped(a, b, c, t, W, H, vsx, vsy, bias_x, bias_y, out_id, zbuf);
      continue;
    }
    /* snapTriangle */
    float rw0 = 1.0f / a[3], rw1 = 1.0f / b[3], rw2 = 1.0f / c[3];
    int p0x = f32_to_s32_sat_rni(a[0] * rw0 * vsx), p0y = f32_to_s32_sat_rni(a[1] * rw0 * vsy);
    int p1x = f32_to_s32_sat_rni(b[0] * rw1 * vsx), p1y = f32_to_s32_sat_rni(b[1] * rw1 * vsy);
    int p2x = f32_to_s32_sat_rni(c[0] * rw2 * vsx), p2y = f32_to_s32_sat_rni(c[1] * rw2 * vsy);
    int lox = imin3(p0x, p1x, p2x), loy = imin3(p0y, p1y, p2y);
    int hix = imax3(p0x, p1x, p2x), hiy = imax3(p0y, p1y, p2y);
    int loxy = lox < loy ? lox : loy, hixy = hix > hiy ? hix : hiy;
    int aabb_limit = (1 << (MAXVIEWPORT_LOG2 + SUBPIXEL_LOG2)) - 1;
    if (!(loxy >= -32768 && hixy <= 32767 && hixy - loxy <= aabb_limit)) {
      ++clip_path;
      raster_clipped(a, b, c, t, W, H, vsx, vsy, bias_x, bias_y, out_id, zbuf);
      continue;
    }
    raster_piece(p0x, p0y, p1x, p1y, p2x, p2y, a[2], b[2], c[2], rw0, rw1, rw2, t, W, H, bias_x, bias_y, out_id, zbuf);
  }
  if (n_clip_path) *n_clip_path = clip_path;
  if (!out_depth) free(zbuf);
  return 0;
}
