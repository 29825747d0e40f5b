// C++ host-side test of the drop-in boundary, shaped like the reference's own
// simple_tests/src/test_foundationpose.cpp (Register on the first frame, Track on the following ones,
// then the speed loops of :112-210) but self-contained: a procedural mesh loader and an analytic
// RGB-D frame of a sphere stand where the reference reads mustard0 from disk.
//
//   test_foundationpose_b200 <refiner.fpb2w> <scorer.fpb2w> [track_iterations]
//
// Exit code 0 = all checks passed; 2 = model construction failed (e.g. no GPU: there is no CPU path).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

#include "detection_6d_b200.hpp"

using namespace detection_6d;

namespace {

constexpr int kH = 480, kW = 640;
constexpr float kFx = 600.f, kFy = 600.f, kCx = 320.f, kCy = 240.f;
constexpr float kRadius = 0.06f;
const float kCenter[3] = {0.03f, -0.02f, 0.65f};  // sphere centre in the camera frame

// UV sphere with per-vertex normals / texcoords and a procedural texture: the BaseMeshLoader contract
// (mesh_loader.hpp:15-62) implemented without assimp.
class SphereLoader final : public BaseMeshLoader {
 public:
  SphereLoader(std::string name, int slices, int stacks) : name_(std::move(name)) {
    const float pi = 3.14159265358979f;
    for (int j = 0; j <= stacks; ++j)
      for (int i = 0; i <= slices; ++i) {
        const float th = pi * j / stacks, ph = 2 * pi * i / slices;
        const float n[3] = {std::sin(th) * std::cos(ph), std::sin(th) * std::sin(ph), std::cos(th)};
        verts_.push_back(make3(kRadius * n[0] + 0.2f, kRadius * n[1], kRadius * n[2]));  // off-centre on purpose
        normals_.push_back(make3(n[0], n[1], n[2]));
        uvs_.push_back(make3((float)i / slices, (float)j / stacks, 0.f));
      }
    for (int j = 0; j < stacks; ++j)
      for (int i = 0; i < slices; ++i) {
        const uint32_t a = j * (slices + 1) + i, b = a + 1, c = a + slices + 1, d = c + 1;
        if (j > 0) faces_.push_back(make3u(a, c, b));
        if (j < stacks - 1) faces_.push_back(make3u(b, c, d));
      }
    center_ = make3(0.2f, 0.f, 0.f);
    // a box frame that is NOT the identity (rotation by 90 degrees about z, origin at the centroid), as the reference's
    // PCA box would be: ConvertPoseMesh2BBox must apply it (mesh_loader.hpp:75-81)
    dim_ = make3(2 * kRadius, 2 * kRadius, 2 * kRadius);
    const float ob[16] = {0, 1, 0, 0, -1, 0, 0, 0, 0, 0, 1, 0, 0.2f, 0, 0, 1};  // column-major
    std::memcpy(bounds_.data(), ob, sizeof(ob));
    tex_pixels_.resize(64 * 64 * 3);
    for (int y = 0; y < 64; ++y)
      for (int x = 0; x < 64; ++x) {
        const bool on = ((x / 8) + (y / 8)) & 1;
        tex_pixels_[3 * (y * 64 + x) + 0] = on ? 220 : 40;
        tex_pixels_[3 * (y * 64 + x) + 1] = (uint8_t)(4 * x);
        tex_pixels_[3 * (y * 64 + x) + 2] = (uint8_t)(4 * y);
      }
#ifdef FPB200_HAVE_EIGEN_OPENCV
    tex_ = cv::Mat(64, 64, CV_8UC3, tex_pixels_.data());
#else
    tex_.rows = tex_.cols = 64;
    tex_.channels = 3;
    tex_.data = tex_pixels_.data();
#endif
  }
  std::string GetName() const noexcept override { return name_; }
  float GetMeshDiameter() const noexcept override { return 2 * kRadius; }
  size_t GetMeshNumVertices() const noexcept override { return verts_.size(); }
  size_t GetMeshNumFaces() const noexcept override { return faces_.size(); }
  const std::vector<Vec3f>& GetMeshVertices() const noexcept override { return verts_; }
  const std::vector<Vec3f>& GetMeshVertexNormals() const noexcept override { return normals_; }
  const std::vector<Vec3f>& GetMeshTextureCoords() const noexcept override { return uvs_; }
  const std::vector<Vector3ui>& GetMeshTriangleFaces() const noexcept override { return faces_; }
  const Vec3f& GetMeshModelCenter() const noexcept override { return center_; }
  const Mat4f& GetOrientBounds() const noexcept override { return bounds_; }
  const Vec3f& GetObjectDimension() const noexcept override { return dim_; }
  const Image& GetTextureMap() const noexcept override { return tex_; }

 private:
  static Vec3f make3(float x, float y, float z) {
#ifdef FPB200_HAVE_EIGEN_OPENCV
    return Vec3f(x, y, z);
#else
    return Vec3f{{x, y, z}};
#endif
  }
  static Vector3ui make3u(uint32_t a, uint32_t b, uint32_t c) {
    Vector3ui f;
    f[0] = a; f[1] = b; f[2] = c;
    return f;
  }
  std::string name_;
  std::vector<Vec3f> verts_, normals_, uvs_;
  std::vector<Vector3ui> faces_;
  Vec3f center_, dim_;
  Mat4f bounds_;
  std::vector<uint8_t> tex_pixels_;
  Image tex_;
};

struct Frame {
  std::vector<uint8_t> rgb, mask;
  std::vector<float> depth;
};

// Analytic ray / sphere intersection: depth in metres (0 = background dropout), silhouette mask.
Frame make_frame() {
  Frame f;
  f.rgb.assign(kH * kW * 3, 90);
  f.mask.assign(kH * kW, 0);
  f.depth.assign(kH * kW, 1.2f);
  for (int v = 0; v < kH; ++v)
    for (int u = 0; u < kW; ++u) {
      const float d[3] = {(u - kCx) / kFx, (v - kCy) / kFy, 1.f};
      const float a = d[0] * d[0] + d[1] * d[1] + 1.f;
      const float b = d[0] * kCenter[0] + d[1] * kCenter[1] + kCenter[2];
      const float c = kCenter[0] * kCenter[0] + kCenter[1] * kCenter[1] + kCenter[2] * kCenter[2] - kRadius * kRadius;
      const float disc = b * b - a * c;
      if (disc <= 0) continue;
      const float t = (b - std::sqrt(disc)) / a;
      f.depth[v * kW + u] = t;
      f.mask[v * kW + u] = 255;
      const float nz = (t - kCenter[2]) / kRadius;
      const uint8_t shade = (uint8_t)(60 + 180 * std::fmin(1.f, std::fabs(nz)));
      f.rgb[3 * (v * kW + u) + 0] = shade;
      f.rgb[3 * (v * kW + u) + 1] = (uint8_t)(u & 0xff);
      f.rgb[3 * (v * kW + u) + 2] = (uint8_t)(v & 0xff);
    }
  return f;
}

Image wrap(const void* p, int channels, bool is_float) {
#ifdef FPB200_HAVE_EIGEN_OPENCV
  return cv::Mat(kH, kW, is_float ? CV_32FC1 : (channels == 3 ? CV_8UC3 : CV_8UC1), const_cast<void*>(p));
#else
  Image m;
  m.rows = kH;
  m.cols = kW;
  m.channels = channels;
  m.is_float = is_float;
  m.data = p;
  return m;
#endif
}

#define CHECK(cond)                                                    \
  do {                                                                 \
    if (!(cond)) {                                                     \
      std::fprintf(stderr, "CHECK failed: %s (line %d)\n", #cond, __LINE__); \
      return 1;                                                        \
    }                                                                  \
  } while (0)

}  // namespace

int main(int argc, char** argv) {
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s refiner.fpb2w scorer.fpb2w [track_iterations]\n", argv[0]);
    return 64;
  }
  const int track_iters = argc > 3 ? std::atoi(argv[3]) : 50;
  auto loader = std::make_shared<SphereLoader>("ball", 48, 24);
  Mat3f K;
#ifdef FPB200_HAVE_EIGEN_OPENCV
  K << kFx, 0, kCx, 0, kFy, kCy, 0, 0, 1;
#else
  const float kcm[9] = {kFx, 0, 0, 0, kFy, 0, kCx, kCy, 1};  // column-major
  for (int i = 0; i < 9; ++i) K.v[i] = kcm[i];
#endif
  std::shared_ptr<Base6DofDetectionModel> model;
  try {
    model = CreateFoundationPoseModelB200(argv[1], argv[2], {loader}, K, kH, kW);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "model construction failed: %s\n", e.what());
    return 2;
  }
  const Frame f = make_frame();
  const Image rgb = wrap(f.rgb.data(), 3, false), depth = wrap(f.depth.data(), 1, true),
              mask = wrap(f.mask.data(), 1, false);

  // ---- Register (test_foundationpose.cpp:57-63)
  Mat4f pose;
  CHECK(model->Register(rgb, depth, mask, "ball", pose, 1));
  const float* p = mat_data(pose);  // column-major: translation = elements 12..14
  for (int i = 0; i < 16; ++i) CHECK(std::isfinite(p[i]));
  // pose of the CENTRED mesh in the camera frame: within one refine step of the sphere centre
  CHECK(std::fabs(p[12] - kCenter[0]) < 0.03f && std::fabs(p[13] - kCenter[1]) < 0.03f &&
        std::fabs(p[14] - kCenter[2]) < 0.05f);
  CHECK(p[3] == 0.f && p[7] == 0.f && p[11] == 0.f && p[15] == 1.f);
  // R R^T = I
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      float dot = 0;
      for (int k = 0; k < 3; ++k) dot += p[4 * k + r] * p[4 * k + c];
      CHECK(std::fabs(dot - (r == c ? 1.f : 0.f)) < 1e-3f);
    }
  // ---- ConvertPoseMesh2BBox (mesh_loader.hpp:75-81) = pose * T(-center) * orient_bounds: here the box frame sits at the
  // mesh centre, rotated by 90 degrees about z, so the translation is kept and the box x axis is the pose's y axis
  {
    const Mat4f box = ConvertPoseMesh2BBox(pose, loader);
    const float* b = mat_data(box);
    for (int r = 0; r < 3; ++r) {
      CHECK(std::fabs(b[12 + r] - p[12 + r]) < 1e-5f);
      CHECK(std::fabs(b[r] - p[4 + r]) < 1e-6f && std::fabs(b[4 + r] + p[r]) < 1e-6f && std::fabs(b[8 + r] - p[8 + r]) < 1e-6f);
    }
  }
  // ---- error behaviour: bool false, no throw (foundationpose.cpp:155-179, sampling.cpp:350-353)
  Mat4f dummy;
  CHECK(!model->Register(rgb, depth, mask, "no_such_target", dummy, 1));
  std::vector<uint8_t> empty_mask(kH * kW, 0);
  CHECK(!model->Register(rgb, depth, wrap(empty_mask.data(), 1, false), "ball", dummy, 1));
  CHECK(!model->Register(rgb, rgb, mask, "ball", dummy, 1));  // wrong depth type

  // ---- Track (test_foundationpose.cpp:74-90) + speed loop (:160-210)
  Mat4f cur = pose;
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < track_iters; ++i) {
    Mat4f next;
    CHECK(model->Track(rgb, depth, cur, "ball", next, 1));
    cur = next;
  }
  const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  const float* q = mat_data(cur);
  for (int i = 0; i < 16; ++i) CHECK(std::isfinite(q[i]));
  // (seeded, untrained weights: the tracked pose random-walks, so only its validity is asserted)
  CHECK(q[15] == 1.f && q[3] == 0.f);
  std::printf("Register ok: t = (%.3f, %.3f, %.3f); Track x%d: %.2f ms/call (%.0f fps)\n", p[12], p[13], p[14],
              track_iters, ms / track_iters, 1e3 * track_iters / ms);
  return 0;
}
