// Drop-in definition of the reference's factory, with the reference's EXACT signature:
//
//   std::shared_ptr<Base6DofDetectionModel> detection_6d::CreateFoundationPoseModel(
//       std::shared_ptr<inference_core::BaseInferCore> refiner_core,
//       std::shared_ptr<inference_core::BaseInferCore> scorer_core,
//       const std::vector<std::shared_ptr<BaseMeshLoader>>& mesh_loaders,
//       const Eigen::Matrix3f& intrinsic_in_mat, int max_input_image_height = 1080, int max_input_image_width = 1920)
//   (detection_6d_foundationpose/include/detection_6d_foundationpose/foundationpose.hpp:99-105)
//
// A maintainer compiles THIS file in place of detection_6d_foundationpose/src/foundationpose*.{cpp,cu} (and the
// vendored nvdiffrast), links libfpb200.so, and creates the two cores with CreateB200InferCore instead of
// CreateTrtInferCore; nothing else of simple_tests/src/test_foundationpose.cpp:20-46 changes.  Given two
// B200InferCore objects the factory builds ONE fused session (rasteriser, crop / warp, both networks, pose update,
// ranking as one launch sequence: csrc/session.cu) from their weight files; the cores themselves remain usable through
// SyncInfer for callers that only want the networks.  Any other BaseInferCore is rejected loudly (this library has no
// TensorRT path).  Includes the reference's own headers; needs Eigen and OpenCV like the reference does.
#include <cstdio>
#include <stdexcept>

#include "b200_infer_core.hpp"
#include "detection_6d_foundationpose/foundationpose.hpp"
#include "fpb200.h"

namespace detection_6d {

namespace {

class FoundationPoseB200Model final : public Base6DofDetectionModel {
 public:
  FoundationPoseB200Model(const std::string& refiner_weights, const std::string& scorer_weights,
                          const std::vector<std::shared_ptr<BaseMeshLoader>>& mesh_loaders, const Eigen::Matrix3f& K,
                          int max_h, int max_w) {
    if (mesh_loaders.empty()) throw std::invalid_argument("[FoundationPoseB200] no mesh loaders");
    const size_t n = mesh_loaders.size();
    std::vector<fpb200_mesh> meshes(n);
    std::vector<std::vector<float>> v(n), nr(n), tc(n);
    std::vector<std::vector<int32_t>> f(n);
    std::vector<std::string> names(n);
    for (size_t i = 0; i < n; ++i) {
      const BaseMeshLoader& ld = *mesh_loaders[i];
      names[i] = ld.GetName();
      const size_t nv = ld.GetMeshNumVertices(), nf = ld.GetMeshNumFaces();
      v[i].resize(nv * 3);
      nr[i].resize(nv * 3);
      tc[i].resize(nv * 2);
      f[i].resize(nf * 3);
      for (size_t k = 0; k < nv; ++k) {
        for (int c = 0; c < 3; ++c) {
          v[i][3 * k + c] = ld.GetMeshVertices()[k][c];
          nr[i][3 * k + c] = ld.GetMeshVertexNormals()[k][c];
        }
        tc[i][2 * k] = ld.GetMeshTextureCoords()[k][0];
        tc[i][2 * k + 1] = ld.GetMeshTextureCoords()[k][1];
      }
      for (size_t k = 0; k < nf; ++k)
        for (int c = 0; c < 3; ++c) f[i][3 * k + c] = static_cast<int32_t>(ld.GetMeshTriangleFaces()[k][c]);
      const cv::Mat& tex = ld.GetTextureMap();
      // foundationpose_render.cpp:418-421: the texture must be a continuous 3-channel u8 image
      if (tex.empty() || !tex.isContinuous() || tex.channels() != 3 || tex.depth() != CV_8U)
        throw std::runtime_error("[FoundationPoseB200] texture map must be continuous RGB u8");
      fpb200_mesh& m = meshes[i];
      m.name = names[i].c_str();
      m.vertices = v[i].data();
      m.normals = nr[i].data();
      m.texcoords = tc[i].data();
      m.faces = f[i].data();
      m.texture = tex.ptr<uint8_t>();
      m.num_vertices = static_cast<int>(nv);
      m.num_faces = static_cast<int>(nf);
      m.tex_h = tex.rows;
      m.tex_w = tex.cols;
      m.diameter = ld.GetMeshDiameter();
      for (int c = 0; c < 3; ++c) m.center[c] = ld.GetMeshModelCenter()[c];
    }
    float k9[9];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) k9[3 * r + c] = K(r, c);
    if (fpb200_create(meshes.data(), static_cast<int>(n), k9, max_h, max_w, refiner_weights.c_str(), scorer_weights.c_str(),
                      nullptr, &s_) != FPB200_OK)
      throw std::runtime_error(std::string("[FoundationPoseB200] ") + fpb200_last_error());
  }
  ~FoundationPoseB200Model() override { fpb200_destroy(s_); }

  bool Register(const cv::Mat& rgb, const cv::Mat& depth, const cv::Mat& mask, const std::string& target_name,
                Eigen::Matrix4f& out_pose_in_mesh, size_t refine_itr) override {
    // CheckInputArguments, foundationpose.cpp:155-179
    if (!Ok(rgb, 3, CV_8U) || !Ok(depth, 1, CV_32F) || !Ok(mask, 1, CV_8U) || rgb.rows != depth.rows ||
        rgb.cols != depth.cols || mask.rows != depth.rows || mask.cols != depth.cols)
      return Fail("Register: rgb / depth / mask must be continuous u8x3 / f32 / u8 images of one size");
    const int rc = fpb200_register(s_, rgb.ptr<uint8_t>(), depth.ptr<float>(), mask.ptr<uint8_t>(), depth.rows, depth.cols,
                                   target_name.c_str(), static_cast<int>(refine_itr), out_pose_in_mesh.data());
    return rc == FPB200_OK ? true : Fail(fpb200_last_error());
  }
  bool Track(const cv::Mat& rgb, const cv::Mat& depth, const Eigen::Matrix4f& hyp_pose_in_mesh,
             const std::string& target_name, Eigen::Matrix4f& out_pose_in_mesh, size_t refine_itr) override {
    if (!Ok(rgb, 3, CV_8U) || !Ok(depth, 1, CV_32F) || rgb.rows != depth.rows || rgb.cols != depth.cols)
      return Fail("Track: rgb / depth must be continuous u8x3 / f32 images of one size");
    const int rc = fpb200_track(s_, rgb.ptr<uint8_t>(), depth.ptr<float>(), depth.rows, depth.cols, hyp_pose_in_mesh.data(),
                                target_name.c_str(), static_cast<int>(refine_itr), out_pose_in_mesh.data());
    return rc == FPB200_OK ? true : Fail(fpb200_last_error());
  }

 private:
  static bool Ok(const cv::Mat& m, int channels, int depth) {
    return !m.empty() && m.isContinuous() && m.channels() == channels && m.depth() == depth;
  }
  static bool Fail(const char* msg) {  // the reference logs with glog and returns false
    LOG(ERROR) << "[FoundationPoseB200] " << msg;
    return false;
  }
  fpb200_session* s_ = nullptr;
};

}  // namespace

std::shared_ptr<Base6DofDetectionModel> CreateFoundationPoseModel(
    std::shared_ptr<inference_core::BaseInferCore> refiner_core, std::shared_ptr<inference_core::BaseInferCore> scorer_core,
    const std::vector<std::shared_ptr<BaseMeshLoader>>& mesh_loaders, const Eigen::Matrix3f& intrinsic_in_mat,
    const int max_input_image_height, const int max_input_image_width) {
  auto refiner = std::dynamic_pointer_cast<inference_core::B200InferCore>(refiner_core);
  auto scorer = std::dynamic_pointer_cast<inference_core::B200InferCore>(scorer_core);
  if (!refiner || !scorer)
    throw std::invalid_argument("[FoundationPoseB200] CreateFoundationPoseModel needs two cores made by CreateB200InferCore "
                                "(this library has no TensorRT path)");
  if (refiner->kind() != 0 || scorer->kind() != 1)
    throw std::invalid_argument("[FoundationPoseB200] refiner_core / scorer_core are swapped or hold the wrong network");
  return std::make_shared<FoundationPoseB200Model>(refiner->weight_path(), scorer->weight_path(), mesh_loaders,
                                                   intrinsic_in_mat, max_input_image_height, max_input_image_width);
}

}  // namespace detection_6d
