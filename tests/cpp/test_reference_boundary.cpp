// Compile / link check of the drop-in boundary against the REFERENCE's own headers (test_cpp_adaptor.py compiles
// this file with -I<reference>/easy_deploy_tool/deploy_core/include and
// -I<reference>/detection_6d_foundationpose/include, stub glog / OpenCV / Eigen headers from tests/cpp/stubs, the
// reference's deploy_core/src/base_infer_core.cpp compiled in place, and libfpb200.so):
//   * B200InferCore derives from inference_core::BaseInferCore and implements every pure virtual,
//   * CreateB200InferCore has CreateTrtInferCore's parameter list,
//   * detection_6d::CreateFoundationPoseModel(refiner_core, scorer_core, loaders, K, h, w) - the exact signature of
//     foundationpose.hpp:99-105 - is defined by cpp/foundationpose_b200.cpp,
//   * a loader written for the reference's BaseMeshLoader (all 12 getters, Vector3ui faces, GetOrientBounds) works.
// The construction block below is the reference's simple_tests/src/test_foundationpose.cpp:20-46 with the one word
// CreateTrtInferCore -> CreateB200InferCore.  Exit code 0: Register / Track ran; 2: construction failed (no GPU:
// there is no CPU path).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "b200_infer_core.hpp"
#include "detection_6d_foundationpose/foundationpose.hpp"

using namespace detection_6d;
using namespace inference_core;

namespace {
class SphereLoader final : public BaseMeshLoader {
 public:
  SphereLoader() {
    const int slices = 24, stacks = 12;
    const float pi = 3.14159265358979f, radius = 0.06f;
    for (int j = 0; j <= stacks; ++j)
      for (int i = 0; i <= slices; ++i) {
        const float th = pi * j / stacks, ph = 2 * pi * i / slices;
        Eigen::Vector3f n, v, t;
        n[0] = std::sin(th) * std::cos(ph); n[1] = std::sin(th) * std::sin(ph); n[2] = std::cos(th);
        for (int c = 0; c < 3; ++c) v[c] = radius * n[c];
        t[0] = (float)i / slices; t[1] = (float)j / stacks; t[2] = 0.f;
        verts_.push_back(v); normals_.push_back(n); uvs_.push_back(t);
      }
    for (int j = 0; j < stacks; ++j)
      for (int i = 0; i < slices; ++i) {
        const uint32_t a = j * (slices + 1) + i, b = a + 1, c = a + slices + 1, d = c + 1;
        Vector3ui f1, f2;
        f1[0] = a; f1[1] = c; f1[2] = b; f2[0] = b; f2[1] = c; f2[2] = d;
        if (j > 0) faces_.push_back(f1);
        if (j < stacks - 1) faces_.push_back(f2);
      }
    tex_pixels_.assign(16 * 16 * 3, 128);
    tex_ = cv::Mat(16, 16, CV_8UC3, tex_pixels_.data());
    bounds_ = Eigen::Matrix4f::Identity();
    for (int c = 0; c < 3; ++c) dim_[c] = 2 * radius;
  }
  std::string GetName() const noexcept override { return "sphere"; }
  float GetMeshDiameter() const noexcept override { return 0.12f; }
  size_t GetMeshNumVertices() const noexcept override { return verts_.size(); }
  size_t GetMeshNumFaces() const noexcept override { return faces_.size(); }
  const std::vector<Eigen::Vector3f>& GetMeshVertices() const noexcept override { return verts_; }
  const std::vector<Eigen::Vector3f>& GetMeshVertexNormals() const noexcept override { return normals_; }
  const std::vector<Eigen::Vector3f>& GetMeshTextureCoords() const noexcept override { return uvs_; }
  const std::vector<Vector3ui>& GetMeshTriangleFaces() const noexcept override { return faces_; }
  const Eigen::Vector3f& GetMeshModelCenter() const noexcept override { return center_; }
  const Eigen::Matrix4f& GetOrientBounds() const noexcept override { return bounds_; }
  const Eigen::Vector3f& GetObjectDimension() const noexcept override { return dim_; }
  const cv::Mat& GetTextureMap() const noexcept override { return tex_; }

 private:
  std::vector<Eigen::Vector3f> verts_, normals_, uvs_;
  std::vector<Vector3ui> faces_;
  Eigen::Vector3f center_, dim_;
  Eigen::Matrix4f bounds_;
  std::vector<uint8_t> tex_pixels_;
  cv::Mat tex_;
};
}  // namespace

int main(int argc, char** argv) {
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s refiner.fpb2w scorer.fpb2w\n", argv[0]);
    return 1;
  }
  std::shared_ptr<Base6DofDetectionModel> model;
  std::shared_ptr<BaseMeshLoader> mesh_loader = std::make_shared<SphereLoader>();
  bool sync_infer_ok = false;
  try {
    auto refiner_core = CreateB200InferCore(argv[1],
                                            {
                                                {"transf_input", {252, 160, 160, 6}},
                                                {"render_input", {252, 160, 160, 6}},
                                            },
                                            {{"trans", {252, 3}}, {"rot", {252, 3}}}, 1);
    auto scorer_core = CreateB200InferCore(argv[2],
                                           {
                                               {"transf_input", {252, 160, 160, 6}},
                                               {"render_input", {252, 160, 160, 6}},
                                           },
                                           {{"scores", {252, 1}}}, 1);
    // the cores stay usable on their own through the reference's BaseInferCore::SyncInfer (base_infer_core.cpp:33-41),
    // blobs taken from the reference's MemBufferPool, host-located like TrtInferCore's default (dynamic batch of 2)
    {
      auto blobs = refiner_core->GetBuffer(true);
      for (const char* name : {"render_input", "transf_input"}) {
        ITensor* t = blobs->GetTensor(name);
        t->SetShape({2, 160, 160, 6});
        float* p = t->Cast<float>();
        for (size_t i = 0; i < 2 * 160 * 160 * 6; ++i) p[i] = 0.25f * std::sin(0.001f * i + (name[0] == 'r' ? 0.f : 1.f));
      }
      sync_infer_ok = refiner_core->SyncInfer(blobs.get(), 2);
      const float* tr = blobs->GetTensor("trans")->Cast<float>();
      const float* ro = blobs->GetTensor("rot")->Cast<float>();
      for (int i = 0; i < 6; ++i) sync_infer_ok = sync_infer_ok && std::isfinite(tr[i]) && std::isfinite(ro[i]);
      auto sblobs = scorer_core->GetBuffer(true);
      for (const char* name : {"render_input", "transf_input"}) {
        ITensor* t = sblobs->GetTensor(name);
        t->SetShape({2, 160, 160, 6});
        std::memcpy(t->RawPtr(), blobs->GetTensor(name)->RawPtr(), t->GetTensorByteSize());
      }
      sync_infer_ok = sync_infer_ok && scorer_core->SyncInfer(sblobs.get(), 2);
      const float* sc = sblobs->GetTensor("scores")->Cast<float>();
      sync_infer_ok = sync_infer_ok && std::isfinite(sc[0]) && std::isfinite(sc[1]);
      std::printf("SyncInfer %s: trans[0] = %.4f %.4f %.4f, scores = %.4f %.4f\n", sync_infer_ok ? "ok" : "FAILED", tr[0], tr[1],
                  tr[2], sc[0], sc[1]);
    }
    Eigen::Matrix3f K;
    K(0, 0) = 600; K(1, 1) = 600; K(0, 2) = 320; K(1, 2) = 240; K(2, 2) = 1;
    model = CreateFoundationPoseModel(refiner_core, scorer_core, {mesh_loader}, K, 480, 640);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "model construction failed: %s\n", e.what());
    return 2;
  }
  // a frame with the sphere at 0.6 m: analytic depth, flat colour
  const int H = 480, W = 640;
  std::vector<uint8_t> rgb(H * W * 3, 90), mask(H * W, 0);
  std::vector<float> depth(H * W, 1.2f);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const float dx = (x - 320) / 600.f, dy = (y - 240) / 600.f, a = dx * dx + dy * dy + 1, b = -2 * 0.6f,
                  c = 0.36f - 0.0036f, disc = b * b - 4 * a * c;
      if (disc > 0) {
        depth[y * W + x] = (-b - std::sqrt(disc)) / (2 * a);
        mask[y * W + x] = 255;
        rgb[3 * (y * W + x)] = 200;
      }
    }
  cv::Mat rgb_m(H, W, CV_8UC3, rgb.data()), depth_m(H, W, CV_32FC1, depth.data()), mask_m(H, W, CV_8UC1, mask.data());
  Eigen::Matrix4f pose, next;
  if (!model->Register(rgb_m, depth_m, mask_m, "sphere", pose)) return 3;
  if (!model->Track(rgb_m, depth_m, pose, "sphere", next)) return 4;
  const Eigen::Matrix4f box = ConvertPoseMesh2BBox(next, mesh_loader);
  std::printf("Register ok, z = %.3f; Track ok, z = %.3f; bbox pose z = %.3f\n", pose(2, 3), next(2, 3), box(2, 3));
  if (!sync_infer_ok) return 6;
  return std::fabs(next(2, 3) - 0.6f) < 0.1f ? 0 : 5;
}
