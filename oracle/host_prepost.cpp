// ORACLE — test infrastructure only (imported by tests/ and by bench.py's cpu_baseline leg; never by the product).
//
// The reference's HOST-side work on the Register path, restated in dependency-free C++ (the reference writes it with
// Eigen, which is not installable here) so that it can be timed on the GPU box's host cores, single-threaded and on
// all cores:
//   GuessTranslation            detection_6d_foundationpose/src/foundationpose_sampling.cpp:250-298
//   ComputeCropWindowTF/ComputeTF   detection_6d_foundationpose/src/foundationpose_render.cpp:25-75
//   ConstructBBox2D             detection_6d_foundationpose/src/foundationpose_render.cpp:123-149
//   ProjectMatrixFromIntrinsics + Proj * GLcam * pose   foundationpose_render.cpp:151-186,588-593
//   RefinePostProcess           detection_6d_foundationpose/src/foundationpose.cpp:360-406
// The product does all of this on the device (csrc/preprocess.cu: guess_translation / pose_setup kernels,
// csrc/net_ops.cu: refine_update_kernel); tests/test_oracle.py checks these functions against oracle/fp_oracle.py.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace {

// ---- GuessTranslation: mask bounding-box centre + median of the valid depths (std::sort), K^-1 back-projection
bool guess_translation(const float* depth, const uint8_t* mask, int h, int w, const float K[9], float min_depth,
                       float center[3]) {
  std::vector<int> vs, us;
  for (int i = 0; i < h; ++i)
    for (int j = 0; j < w; ++j)
      if (mask[(size_t)i * w + j] > 0) {
        vs.push_back(i);
        us.push_back(j);
      }
  if (us.empty()) return false;
  const float uc = (float)((*std::min_element(us.begin(), us.end()) + *std::max_element(us.begin(), us.end())) / 2.0);
  const float vc = (float)((*std::min_element(vs.begin(), vs.end()) + *std::max_element(vs.begin(), vs.end())) / 2.0);
  std::vector<float> valid;
  for (int i = 0; i < h; ++i)
    for (int j = 0; j < w; ++j)
      if (mask[(size_t)i * w + j] > 0 && depth[(size_t)i * w + j] >= min_depth) valid.push_back(depth[(size_t)i * w + j]);
  if (valid.empty()) return false;
  std::sort(valid.begin(), valid.end());
  const int n = (int)valid.size();
  const float zc = (n % 2 == 0) ? (float)((valid[n / 2 - 1] + valid[n / 2]) / 2.0) : valid[n / 2];
  double a[9];
  for (int i = 0; i < 9; ++i) a[i] = K[i];
  const double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
  const double inv[9] = {(a[4] * a[8] - a[5] * a[7]) / det, (a[2] * a[7] - a[1] * a[8]) / det, (a[1] * a[5] - a[2] * a[4]) / det,
                         (a[5] * a[6] - a[3] * a[8]) / det, (a[0] * a[8] - a[2] * a[6]) / det, (a[2] * a[3] - a[0] * a[5]) / det,
                         (a[3] * a[7] - a[4] * a[6]) / det, (a[1] * a[6] - a[0] * a[7]) / det, (a[0] * a[4] - a[1] * a[3]) / det};
  for (int r = 0; r < 3; ++r)
    center[r] = (((float)inv[3 * r] * uc + (float)inv[3 * r + 1] * vc) + (float)inv[3 * r + 2]) * zc;
  return true;
}

void mat4_mul(const float* a, const float* b, float* out) {  // row-major
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      float s = a[4 * i] * b[j];
      for (int k = 1; k < 4; ++k) s = s + a[4 * i + k] * b[4 * k + j];
      out[4 * i + j] = s;
    }
}

void projection_matrix(const float K[9], int h, int w, float out[16]) {
  const float nc = 0.1f, fc = 100.0f, depth = fc - nc;
  const float q = -(fc + nc) / depth, qn = -2.f * (fc * nc) / depth, wf = (float)w, hf = (float)h;
  const float m[16] = {2.f * K[0] / wf, -2.f * K[1] / wf, (-2.f * K[2] + wf) / wf, 0.f, 0.f, 2.f * K[4] / hf,
                       (2.f * K[5] - hf) / hf, 0.f, 0.f, 0.f, q, qn, 0.f, 0.f, -1.f, 0.f};
  memcpy(out, m, sizeof(m));
}

// ---- per pose: crop window transform, bbox2d, clip-space matrix.  pose: row-major 4x4.
void pose_setup(const float* pose, const float K[9], const float* proj, float diameter, float crop_ratio, int crop,
                float tf[9], float bbox[4], float mvp[16]) {
  const float r = diameter * crop_ratio / 2.f;
  const float tx = pose[3], ty = pose[7], tz = pose[11];
  auto proj_uv = [&](float x, float y, float z, float& u, float& v) {
    u = ((K[0] * x + K[1] * y) + K[2] * z) / z;
    v = (K[4] * y + K[5] * z) / z;
  };
  float u0, v0, u, v;
  proj_uv(tx, ty, tz, u0, v0);
  float dv = 0.f;  // the radius uses the vertical extent only (foundationpose_render.cpp:66)
  const float off[4][2] = {{r, 0}, {-r, 0}, {0, r}, {0, -r}};
  for (auto& o : off) {
    proj_uv(tx + o[0], ty + o[1], tz, u, v);
    dv = std::max(dv, v - v0);
  }
  const float radius = std::fabs(dv);
  const float left = std::round(u0 - radius), right = std::round(u0 + radius);
  const float top = std::round(v0 - radius), bottom = std::round(v0 + radius);
  const float sx = (float)crop / (right - left), sy = (float)crop / (bottom - top);
  const float t[9] = {sx, 0, sx * (-left), 0, sy, sy * (-top), 0, 0, 1};
  memcpy(tf, t, sizeof(t));
  const float bl = -t[2] / sx, bt = -t[5] / sy;
  bbox[0] = bl;
  bbox[1] = bt;
  bbox[2] = (float)(crop - 1) / sx + bl;
  bbox[3] = (float)(crop - 1) / sy + bt;
  const float gl[16] = {1, 0, 0, 0, 0, -1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 1};
  float tmp[16];
  mat4_mul(gl, pose, tmp);
  mat4_mul(proj, tmp, mvp);
}

// ---- RefinePostProcess: t += trans * d/2; R <- Rodrigues(tanh(rot) * 0.349066)^T R.  pose row-major, in place.
void refine_post(float* pose, const float trans[3], const float rot[3], float diameter) {
  for (int k = 0; k < 3; ++k) pose[4 * k + 3] += trans[k] * (diameter / 2.f);
  float v[3];
  for (int k = 0; k < 3; ++k) v[k] = std::tanh(rot[k]) * 0.349065850398865f;
  const float ang = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  float rm[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  if (ang > 0.f) {
    const float ax = v[0] / ang, ay = v[1] / ang, az = v[2] / ang, c = std::cos(ang), s = std::sin(ang), k = 1.f - c;
    const float m[9] = {c + k * ax * ax, k * ax * ay - s * az, k * ax * az + s * ay, k * ay * ax + s * az, c + k * ay * ay,
                        k * ay * az - s * ax, k * az * ax - s * ay, k * az * ay + s * ax, c + k * az * az};
    memcpy(rm, m, sizeof(m));
  }
  float old[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) old[3 * r + c] = pose[4 * r + c];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) pose[4 * r + c] = rm[r] * old[c] + rm[3 + r] * old[3 + c] + rm[6 + r] * old[6 + c];
}

}  // namespace

extern "C" {

int host_guess_translation(const float* depth, const uint8_t* mask, int h, int w, const float* K, float* center) {
  return guess_translation(depth, mask, h, w, K, 0.001f, center) ? 0 : -1;
}

// poses: [n][16] row-major.  Outputs tf [n][9], bbox [n][4], mvp [n][16].
void host_pose_setup(const float* poses, int n, const float* K, int img_h, int img_w, float diameter, float crop_ratio,
                     float* tf, float* bbox, float* mvp) {
  float proj[16];
  projection_matrix(K, img_h, img_w, proj);
  for (int i = 0; i < n; ++i) pose_setup(poses + 16 * i, K, proj, diameter, crop_ratio, 160, tf + 9 * i, bbox + 4 * i, mvp + 16 * i);
}

void host_refine_post(float* poses, int n, const float* trans, const float* rot, float diameter) {
  for (int i = 0; i < n; ++i) refine_post(poses + 16 * i, trans + 3 * i, rot + 3 * i, diameter);
}

// One Register's worth of host work as the reference schedules it (refine_itr refine iterations + the score stage):
// GuessTranslation once, then per stage the per-pose window / bbox / clip matrices, and RefinePostProcess per
// iteration; the per-pose loops are split over `threads` host threads.  Returns the average milliseconds over `repeat`.
double host_register_work_ms(const float* depth, const uint8_t* mask, int h, int w, const float* K, const float* grid_poses,
                             int n, float diameter, int refine_itr, int threads, int repeat) {
  std::vector<float> poses((size_t)n * 16), tf((size_t)n * 9), bbox((size_t)n * 4), mvp((size_t)n * 16);
  std::vector<float> trans((size_t)n * 3, 0.01f), rot((size_t)n * 3, 0.02f);
  float proj[16];
  projection_matrix(K, h, w, proj);
  auto parallel = [&](auto&& fn) {
    if (threads <= 1) {
      fn(0, n);
      return;
    }
    std::vector<std::thread> pool;
    const int per = (n + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
      const int lo = t * per, hi = std::min(n, lo + per);
      if (lo < hi) pool.emplace_back([&, lo, hi] { fn(lo, hi); });
    }
    for (auto& th : pool) th.join();
  };
  const auto t0 = std::chrono::steady_clock::now();
  for (int rep = 0; rep < repeat; ++rep) {
    float c[3] = {0, 0, 1};
    guess_translation(depth, mask, h, w, K, 0.001f, c);
    memcpy(poses.data(), grid_poses, sizeof(float) * 16 * n);
    for (int i = 0; i < n; ++i) {
      poses[16 * i + 3] = c[0];
      poses[16 * i + 7] = c[1];
      poses[16 * i + 11] = c[2];
    }
    for (int stage = 0; stage <= refine_itr; ++stage) {
      const float ratio = stage < refine_itr ? 1.2f : 1.1f;
      parallel([&](int lo, int hi) {
        for (int i = lo; i < hi; ++i)
          pose_setup(&poses[16 * i], K, proj, diameter, ratio, 160, &tf[9 * i], &bbox[4 * i], &mvp[16 * i]);
      });
      if (stage < refine_itr)
        parallel([&](int lo, int hi) {
          for (int i = lo; i < hi; ++i) refine_post(&poses[16 * i], &trans[3 * i], &rot[3 * i], diameter);
        });
    }
  }
  const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return ms / repeat;
}

}  // extern "C"
