// Header-only C++ adaptor: the reference's 6-DoF API on top of the fpb200 C ABI (include/fpb200.h).
//
// With Eigen and OpenCV headers present it exposes exactly the reference's types and signatures
//   detection_6d::Base6DofDetectionModel::{Register,Track}
//     (detection_6d_foundationpose/include/detection_6d_foundationpose/foundationpose.hpp:16-77)
//   detection_6d::BaseMeshLoader (mesh_loader.hpp:15-62: all 12 getters incl. GetOrientBounds / GetObjectDimension,
//     faces as Vector3ui), ConvertPoseMesh2BBox (:75-81: pose * T(-center) * orient_bounds)
//   detection_6d::CreateFoundationPoseModelB200(refiner_weights, scorer_weights, mesh_loaders, K, max_h, max_w)
//     -- the one deliberate difference from CreateFoundationPoseModel (foundationpose.hpp:99-105): two
//     weight-file paths stand where the reference takes two TensorRT BaseInferCore objects.
// Without them (this build image has neither) it falls back to plain-array stand-ins with the same
// member names, so the adaptor can still be compiled and linked here (tests/test_host_logic.py).
// Error behaviour follows the reference: run-time failures return false and log to stderr,
// construction failures throw std::runtime_error.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "fpb200.h"

#if defined(__has_include)
#if __has_include(<Eigen/Dense>) && __has_include(<opencv2/core.hpp>)
#define FPB200_HAVE_EIGEN_OPENCV 1
#endif
#endif

#ifdef FPB200_HAVE_EIGEN_OPENCV
#include <Eigen/Dense>
#include <opencv2/core.hpp>
#endif

namespace detection_6d {

#ifdef FPB200_HAVE_EIGEN_OPENCV
using Mat4f = Eigen::Matrix4f;
using Mat3f = Eigen::Matrix3f;
using Vec3f = Eigen::Vector3f;
using Vec3ui = Eigen::Matrix<uint32_t, 3, 1>;
using Image = cv::Mat;
inline const uint8_t* image_u8(const Image& m) { return m.ptr<uint8_t>(); }
inline const float* image_f32(const Image& m) { return m.ptr<float>(); }
inline int image_rows(const Image& m) { return m.rows; }
inline int image_cols(const Image& m) { return m.cols; }
inline bool image_ok(const Image& m, int channels, bool is_float) {
  return !m.empty() && m.isContinuous() && m.channels() == channels && (m.depth() == (is_float ? CV_32F : CV_8U));
}
inline const float* mat_data(const Mat4f& m) { return m.data(); }  // column-major, like the C ABI
inline float* mat_data(Mat4f& m) { return m.data(); }
#else
// Stand-ins with the storage conventions of the real types.
struct Mat4f { float v[16]; const float* data() const { return v; } float* data() { return v; } };  // column-major
struct Mat3f { float v[9]; float operator()(int r, int c) const { return v[3 * c + r]; } };         // column-major
struct Vec3f { float v[3]; float operator[](int i) const { return v[i]; } float& operator[](int i) { return v[i]; } };
struct Vec3ui { uint32_t v[3]; uint32_t operator[](int i) const { return v[i]; } uint32_t& operator[](int i) { return v[i]; } };
struct Image {  // continuous row-major pixel buffer
  int rows = 0, cols = 0, channels = 1;
  bool is_float = false;
  const void* data = nullptr;
};
inline const uint8_t* image_u8(const Image& m) { return static_cast<const uint8_t*>(m.data); }
inline const float* image_f32(const Image& m) { return static_cast<const float*>(m.data); }
inline int image_rows(const Image& m) { return m.rows; }
inline int image_cols(const Image& m) { return m.cols; }
inline bool image_ok(const Image& m, int channels, bool is_float) {
  return m.data && m.channels == channels && m.is_float == is_float;
}
inline const float* mat_data(const Mat4f& m) { return m.data(); }
inline float* mat_data(Mat4f& m) { return m.data(); }
#endif

// mesh_loader.hpp:15-62
class BaseMeshLoader {
 public:
  virtual ~BaseMeshLoader() = default;
  using Vector3ui = Vec3ui;
  virtual std::string GetName() const noexcept = 0;
  virtual float GetMeshDiameter() const noexcept = 0;
  virtual size_t GetMeshNumVertices() const noexcept = 0;
  virtual size_t GetMeshNumFaces() const noexcept = 0;
  virtual const std::vector<Vec3f>& GetMeshVertices() const noexcept = 0;
  virtual const std::vector<Vec3f>& GetMeshVertexNormals() const noexcept = 0;
  virtual const std::vector<Vec3f>& GetMeshTextureCoords() const noexcept = 0;  // (u, v, unused)
  virtual const std::vector<Vector3ui>& GetMeshTriangleFaces() const noexcept = 0;
  virtual const Vec3f& GetMeshModelCenter() const noexcept = 0;
  virtual const Mat4f& GetOrientBounds() const noexcept = 0;    // PCA box frame, assimp_mesh_loader.cpp:62-113
  virtual const Vec3f& GetObjectDimension() const noexcept = 0;  // extents of the vertices along the box axes
  virtual const Image& GetTextureMap() const noexcept = 0;  // RGB u8
};

// foundationpose.hpp:16-77
class Base6DofDetectionModel {
 public:
  virtual ~Base6DofDetectionModel() = default;
  virtual bool Register(const Image& rgb, const Image& depth, const Image& mask, const std::string& target_name,
                        Mat4f& out_pose_in_mesh, size_t refine_itr = 1) = 0;
  virtual bool Track(const Image& rgb, const Image& depth, const Mat4f& hyp_pose_in_mesh,
                     const std::string& target_name, Mat4f& out_pose_in_mesh, size_t refine_itr = 1) = 0;
};

class FoundationPoseB200 final : public Base6DofDetectionModel {
 public:
  FoundationPoseB200(const std::string& refiner_weights, const std::string& scorer_weights,
                     const std::vector<std::shared_ptr<BaseMeshLoader>>& mesh_loaders, const Mat3f& K, int max_h,
                     int max_w, int device = 0) {
    if (mesh_loaders.empty()) throw std::invalid_argument("[FoundationPoseB200] no mesh loaders");
    std::vector<fpb200_mesh> meshes(mesh_loaders.size());
    std::vector<std::vector<float>> keep_f;
    std::vector<std::vector<int32_t>> keep_i;
    std::vector<std::string> names;
    for (const auto& ld : mesh_loaders) names.push_back(ld->GetName());
    for (size_t i = 0; i < mesh_loaders.size(); ++i) {
      const auto& ld = *mesh_loaders[i];
      const size_t nv = ld.GetMeshNumVertices(), nf = ld.GetMeshNumFaces();
      std::vector<float> v(nv * 3), n(nv * 3), t(nv * 2);
      for (size_t k = 0; k < nv; ++k) {
        for (int c = 0; c < 3; ++c) {
          v[3 * k + c] = ld.GetMeshVertices()[k][c];
          n[3 * k + c] = ld.GetMeshVertexNormals()[k][c];
        }
        t[2 * k] = ld.GetMeshTextureCoords()[k][0];
        t[2 * k + 1] = ld.GetMeshTextureCoords()[k][1];
      }
      std::vector<int32_t> f(nf * 3);
      for (size_t k = 0; k < nf; ++k)
        for (int c = 0; c < 3; ++c) f[3 * k + c] = static_cast<int32_t>(ld.GetMeshTriangleFaces()[k][c]);
      const Image& tex = ld.GetTextureMap();
      if (!image_ok(tex, 3, false)) throw std::runtime_error("[FoundationPoseB200] texture map must be continuous RGB u8");
      keep_f.push_back(std::move(v));
      keep_f.push_back(std::move(n));
      keep_f.push_back(std::move(t));
      keep_i.push_back(std::move(f));
      fpb200_mesh& m = meshes[i];
      m.name = names[i].c_str();
      m.vertices = keep_f[3 * i].data();
      m.normals = keep_f[3 * i + 1].data();
      m.texcoords = keep_f[3 * i + 2].data();
      m.faces = keep_i[i].data();
      m.texture = image_u8(tex);
      m.num_vertices = static_cast<int>(nv);
      m.num_faces = static_cast<int>(nf);
      m.tex_h = image_rows(tex);
      m.tex_w = image_cols(tex);
      m.diameter = ld.GetMeshDiameter();
      for (int c = 0; c < 3; ++c) m.center[c] = ld.GetMeshModelCenter()[c];
    }
    float k9[9];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) k9[3 * r + c] = K(r, c);
    fpb200_options opt;
    fpb200_default_options(&opt);
    opt.device = device;
    if (fpb200_create(meshes.data(), static_cast<int>(meshes.size()), k9, max_h, max_w, refiner_weights.c_str(),
                      scorer_weights.c_str(), &opt, &s_) != FPB200_OK)
      throw std::runtime_error(std::string("[FoundationPoseB200] ") + fpb200_last_error());
  }
  ~FoundationPoseB200() override { fpb200_destroy(s_); }
  FoundationPoseB200(const FoundationPoseB200&) = delete;
  FoundationPoseB200& operator=(const FoundationPoseB200&) = delete;

  bool Register(const Image& rgb, const Image& depth, const Image& mask, const std::string& target_name,
                Mat4f& out_pose_in_mesh, size_t refine_itr = 1) override {
    // CheckInputArguments, foundationpose.cpp:155-179
    if (!image_ok(rgb, 3, false) || !image_ok(depth, 1, true) || !image_ok(mask, 1, false) ||
        image_rows(rgb) != image_rows(depth) || image_cols(rgb) != image_cols(depth) ||
        image_rows(mask) != image_rows(depth) || image_cols(mask) != image_cols(depth))
      return Fail("Register: rgb/depth/mask must be continuous u8x3 / f32 / u8 images of one size");
    const int rc = fpb200_register(s_, image_u8(rgb), image_f32(depth), image_u8(mask), image_rows(depth),
                                   image_cols(depth), target_name.c_str(), static_cast<int>(refine_itr),
                                   mat_data(out_pose_in_mesh));
    return rc == FPB200_OK ? true : Fail(fpb200_last_error());
  }

  bool Track(const Image& rgb, const Image& depth, const Mat4f& hyp_pose_in_mesh, const std::string& target_name,
             Mat4f& out_pose_in_mesh, size_t refine_itr = 1) override {
    if (!image_ok(rgb, 3, false) || !image_ok(depth, 1, true) || image_rows(rgb) != image_rows(depth) ||
        image_cols(rgb) != image_cols(depth))
      return Fail("Track: rgb/depth must be continuous u8x3 / f32 images of one size");
    const int rc = fpb200_track(s_, image_u8(rgb), image_f32(depth), image_rows(depth), image_cols(depth),
                                mat_data(hyp_pose_in_mesh), target_name.c_str(), static_cast<int>(refine_itr),
                                mat_data(out_pose_in_mesh));
    return rc == FPB200_OK ? true : Fail(fpb200_last_error());
  }

  fpb200_session* session() const { return s_; }

 private:
  static bool Fail(const char* msg) {
    std::fprintf(stderr, "[FoundationPoseB200] %s\n", msg);
    return false;
  }
  fpb200_session* s_ = nullptr;
};

inline std::shared_ptr<Base6DofDetectionModel> CreateFoundationPoseModelB200(
    const std::string& refiner_weights, const std::string& scorer_weights,
    const std::vector<std::shared_ptr<BaseMeshLoader>>& mesh_loaders, const Mat3f& K, int max_h = 1080,
    int max_w = 1920) {
  return std::make_shared<FoundationPoseB200>(refiner_weights, scorer_weights, mesh_loaders, K, max_h, max_w);
}

// mesh_loader.hpp:75-81: T_bbox = T_mesh * T(-center) * orient_bounds
#ifdef FPB200_HAVE_EIGEN_OPENCV
inline Mat4f ConvertPoseMesh2BBox(const Mat4f& pose_in_mesh, const std::shared_ptr<BaseMeshLoader>& mesh_loader) {
  Mat4f tf_to_center = Mat4f::Identity();
  tf_to_center.block<3, 1>(0, 3) = -mesh_loader->GetMeshModelCenter();
  return pose_in_mesh * tf_to_center * mesh_loader->GetOrientBounds();
}
#else
inline Mat4f ConvertPoseMesh2BBox(const Mat4f& pose_in_mesh, const std::shared_ptr<BaseMeshLoader>& mesh_loader) {
  auto mul = [](const Mat4f& a, const Mat4f& b) {  // column-major 4x4 product
    Mat4f c;
    for (int col = 0; col < 4; ++col)
      for (int row = 0; row < 4; ++row) {
        float acc = 0.f;
        for (int k = 0; k < 4; ++k) acc += a.v[4 * k + row] * b.v[4 * col + k];
        c.v[4 * col + row] = acc;
      }
    return c;
  };
  Mat4f tf_to_center{};
  for (int i = 0; i < 4; ++i) tf_to_center.v[5 * i] = 1.f;
  for (int c = 0; c < 3; ++c) tf_to_center.v[12 + c] = -mesh_loader->GetMeshModelCenter()[c];
  return mul(mul(pose_in_mesh, tf_to_center), mesh_loader->GetOrientBounds());
}
#endif

}  // namespace detection_6d
