// B200InferCore: the refine / score networks of this library behind the reference's OWN inference-core plugin
// interface, to be compiled inside the reference tree (it includes the reference's header, not a copy):
//
//   class inference_core::BaseInferCore      easy_deploy_tool/deploy_core/include/deploy_core/base_infer_core.h:179-237
//   class inference_core::ITensor / BlobsTensor   easy_deploy_tool/deploy_core/include/deploy_core/blob_buffer.h:18-94
//   CreateTrtInferCore(path, in_shapes, out_shapes, mem_buf_size)
//                                             easy_deploy_tool/inference_core/trt_core/include/trt_core/trt_core.h:26-30
//
// CreateB200InferCore has CreateTrtInferCore's signature; the flat weight file (foundationpose_cpp_b200/weights.py)
// stands where the serialized TensorRT engine stood.  The stages do what TrtInferCore's do
// (easy_deploy_tool/inference_core/trt_core/src/trt_core.cpp:296-400): PreProcess uploads HOST-located inputs,
// Inference runs the network on the device buffers of the blobs (fpb200_infer_refiner / fpb200_infer_scorer: the
// hand-written sm_100a kernels), PostProcess downloads HOST-located outputs.  Blob names and layouts are the
// reference's (detection_6d_foundationpose/src/foundationpose.cpp:78-83): render_input / transf_input f32 NHWC
// [N,160,160,6]; trans / rot [N,3]; scores [N,1].
//
// Only the C ABI (include/fpb200.h) is used: no CUDA headers are needed to compile this file.
#pragma once
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "deploy_core/base_infer_core.h"
#include "fpb200.h"

namespace inference_core {

class B200Tensor : public ITensor {
 public:
  B200Tensor(std::string name, std::vector<size_t> shape, int device) : name_(std::move(name)), default_shape_(shape), shape_(std::move(shape)) {
    host_.resize(GetBufferMaxByteSize());
    if (fpb200_device_alloc(GetBufferMaxByteSize(), device, &own_dev_) != FPB200_OK)
      throw std::runtime_error(std::string("[B200Tensor] device allocation failed: ") + fpb200_last_error());
    dev_ = own_dev_;
    host_ptr_ = host_.data();
  }
  ~B200Tensor() override { fpb200_device_free(own_dev_); }
  B200Tensor(const B200Tensor&) = delete;
  B200Tensor& operator=(const B200Tensor&) = delete;

  const std::string& GetName() const noexcept override { return name_; }
  void* RawPtr() override { return location_ == DataLocation::HOST ? host_ptr_ : dev_; }
  void SetBufferLocation(DataLocation location) override {
    if (location == DataLocation::UNKOWN) throw std::runtime_error("[B200Tensor] `SetBufferLocation` Got invalid location: UNKOWN !");
    location_ = location;
  }
  void ToLocation(DataLocation location) override {
    if (location == DataLocation::UNKOWN) throw std::runtime_error("[B200Tensor] `ToLocation` Got invalid location: UNKOWN !");
    if (location == location_) return;
    if (location_ == DataLocation::HOST) Copy(dev_, host_ptr_, FPB200_COPY_H2D);
    else Copy(host_ptr_, dev_, FPB200_COPY_D2H);
  }
  DataLocation GetBufferLocation() const noexcept override { return location_; }
  void ZeroCopy(ITensor* tensor) override {
    if (!tensor || !tensor->RawPtr()) throw std::runtime_error("[B200Tensor] `ZeroCopy` Got invalid tensor !");
    const auto loc = tensor->GetBufferLocation();
    if (loc == DataLocation::DEVICE) dev_ = tensor->RawPtr();
    else host_ptr_ = tensor->RawPtr();
    location_ = loc;
  }
  void DeepCopy(ITensor* tensor) override {
    if (!tensor || !tensor->RawPtr()) throw std::runtime_error("[B200Tensor] `DeepCopy` Got invalid tensor !");
    const auto loc = tensor->GetBufferLocation();
    if (loc == DataLocation::DEVICE) {
      dev_ = own_dev_;
      Copy(dev_, tensor->RawPtr(), FPB200_COPY_D2D);
    } else {
      host_ptr_ = host_.data();
      std::memcpy(host_ptr_, tensor->RawPtr(), GetTensorByteSize());
    }
    location_ = loc;
  }
  const std::vector<size_t>& GetDefaultShape() const noexcept override { return default_shape_; }
  const std::vector<size_t>& GetShape() const noexcept override { return shape_; }
  void SetShape(const std::vector<size_t>& shape) override {
    if (sizeof(float) * Count(shape) > GetBufferMaxByteSize())
      throw std::runtime_error("[B200Tensor] `SetShape` Got invalid shape: exceeds max byte size !");
    shape_ = shape;
  }
  size_t GetBufferMaxByteSize() const noexcept override { return sizeof(float) * Count(default_shape_); }
  size_t GetTensorByteSize() const noexcept override { return sizeof(float) * Count(shape_); }
  size_t GetElementByteSize() const noexcept override { return sizeof(float); }

  void* device_ptr() const { return dev_; }
  void* host_ptr() const { return host_ptr_; }
  void Copy(void* dst, const void* src, int kind) {
    if (fpb200_memcpy(dst, src, GetTensorByteSize(), kind) != FPB200_OK)
      throw std::runtime_error(std::string("[B200Tensor] copy failed: ") + fpb200_last_error());
  }

 private:
  static size_t Count(const std::vector<size_t>& s) {
    size_t n = 1;
    for (size_t d : s) n *= d;
    return n;
  }
  std::string name_;
  std::vector<size_t> default_shape_, shape_;
  std::vector<unsigned char> host_;
  void *own_dev_ = nullptr, *dev_ = nullptr, *host_ptr_ = nullptr;
  DataLocation location_ = DataLocation::HOST;
};

class B200InferCore : public BaseInferCore {
 public:
  using ShapeMap = std::unordered_map<std::string, std::vector<uint64_t>>;

  B200InferCore(const std::string& weight_path, const ShapeMap& input_blobs_shape, const ShapeMap& output_blobs_shape,
                int mem_buf_size, int device = 0)
      : weight_path_(weight_path), inputs_(input_blobs_shape), outputs_(output_blobs_shape), device_(device) {
    // the reference validates blob names when the model is built (foundationpose.cpp:126-139); here the kind of the
    // network follows from the output blobs
    for (const char* n : {"render_input", "transf_input"})
      if (!inputs_.count(n)) throw std::invalid_argument(std::string("[B200InferCore] input blob missing: ") + n);
    if (outputs_.count("trans") && outputs_.count("rot")) kind_ = 0;
    else if (outputs_.count("scores")) kind_ = 1;
    else throw std::invalid_argument("[B200InferCore] output blobs must be {trans, rot} (refiner) or {scores} (scorer)");
    const auto& shp = inputs_.at("render_input");
    if (shp.size() != 4 || shp[1] != 160 || shp[2] != 160 || shp[3] != 6 || inputs_.at("transf_input") != shp)
      throw std::invalid_argument("[B200InferCore] input blobs must be [N,160,160,6]");
    max_batch_ = static_cast<int>(shp[0]);
    if (fpb200_infer_create(weight_path.c_str(), kind_, max_batch_, device, &core_) != FPB200_OK)
      throw std::runtime_error(std::string("[B200InferCore] ") + fpb200_last_error());
    Init(mem_buf_size);
  }
  ~B200InferCore() override {
    Release();
    fpb200_infer_destroy(core_);
  }

  std::unique_ptr<BlobsTensor> AllocBlobsBuffer() override {
    std::unordered_map<std::string, std::unique_ptr<ITensor>> map;
    for (const auto* shapes : {&inputs_, &outputs_})
      for (const auto& kv : *shapes)
        map.emplace(kv.first, std::make_unique<B200Tensor>(kv.first, std::vector<size_t>(kv.second.begin(), kv.second.end()), device_));
    return std::make_unique<BlobsTensor>(std::move(map));
  }
  InferCoreType GetType() override { return InferCoreType::NOT_PROVIDED; }
  std::string GetName() override { return "b200_core"; }

  const std::string& weight_path() const { return weight_path_; }
  int kind() const { return kind_; }  // 0 refiner, 1 scorer
  int max_batch() const { return max_batch_; }

 protected:
  bool PreProcess(std::shared_ptr<async_pipeline::IPipelinePackage> unit) override {
    CHECK_STATE(unit != nullptr && unit->GetInferBuffer() != nullptr, "[B200InferCore] PreProcess got invalid buffer!");
    for (const auto& kv : inputs_) {
      auto* t = dynamic_cast<B200Tensor*>(unit->GetInferBuffer()->GetTensor(kv.first));
      CHECK_STATE(t != nullptr, "[B200InferCore] PreProcess got invalid tensor : " + kv.first);
      if (t->GetBufferLocation() == DataLocation::HOST) t->Copy(t->device_ptr(), t->host_ptr(), FPB200_COPY_H2D);
    }
    return true;
  }
  bool Inference(std::shared_ptr<async_pipeline::IPipelinePackage> unit) override {
    CHECK_STATE(unit != nullptr && unit->GetInferBuffer() != nullptr, "[B200InferCore] Inference got invalid buffer!");
    BlobsTensor* blobs = unit->GetInferBuffer();
    auto dev = [&](const char* name) { return static_cast<float*>(dynamic_cast<B200Tensor*>(blobs->GetTensor(name))->device_ptr()); };
    const int n = static_cast<int>(blobs->GetTensor("render_input")->GetShape()[0]);  // dynamic batch, foundationpose.cpp:349-354
    const int rc = kind_ == 0 ? fpb200_infer_refiner(core_, dev("render_input"), dev("transf_input"), n, dev("trans"), dev("rot"))
                              : fpb200_infer_scorer(core_, dev("render_input"), dev("transf_input"), n, dev("scores"));
    CHECK_STATE(rc == FPB200_OK, std::string("[B200InferCore] Inference failed: ") + fpb200_last_error());
    return true;
  }
  bool PostProcess(std::shared_ptr<async_pipeline::IPipelinePackage> unit) override {
    CHECK_STATE(unit != nullptr && unit->GetInferBuffer() != nullptr, "[B200InferCore] PostProcess got invalid buffer!");
    for (const auto& kv : outputs_) {
      auto* t = dynamic_cast<B200Tensor*>(unit->GetInferBuffer()->GetTensor(kv.first));
      CHECK_STATE(t != nullptr, "[B200InferCore] PostProcess got invalid tensor : " + kv.first);
      if (t->GetBufferLocation() == DataLocation::HOST) t->Copy(t->host_ptr(), t->device_ptr(), FPB200_COPY_D2H);
    }
    return true;
  }

 private:
  std::string weight_path_;
  ShapeMap inputs_, outputs_;
  int device_ = 0, kind_ = 0, max_batch_ = 0;
  fpb200_infer_core* core_ = nullptr;
};

// Same signature as CreateTrtInferCore (trt_core.h:26-30).
inline std::shared_ptr<BaseInferCore> CreateB200InferCore(std::string model_path,
                                                          const std::unordered_map<std::string, std::vector<uint64_t>>& input_blobs_shape = {},
                                                          const std::unordered_map<std::string, std::vector<uint64_t>>& output_blobs_shape = {},
                                                          const int mem_buf_size = 5) {
  return std::make_shared<B200InferCore>(model_path, input_blobs_shape, output_blobs_shape, mem_buf_size);
}

}  // namespace inference_core
