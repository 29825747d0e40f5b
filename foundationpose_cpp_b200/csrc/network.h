// Refine / score networks as sequences of tcgen05 implicit-GEMM layers and small fused operators.
// Replaces the two TensorRT engines behind BaseInferCore::SyncInfer in the reference
// (easy_deploy_tool/deploy_core/src/base_infer_core.cpp:33-41, inference_core/trt_core/src/trt_core.cpp:323-371).
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "conv_gemm.h"
#include "preprocess.h"

namespace fpb {

struct WeightTensor {
  void* dev = nullptr;
  int dtype = 0;  // 0 = f16, 1 = f32
  int ndim = 0;
  int dims[4] = {1, 1, 1, 1};
};

// Flat weight file (format: foundationpose_cpp_b200/weights.py) resident in device memory.
class WeightFile {
 public:
  WeightFile(const std::string& path, int expected_kind);
  ~WeightFile();
  const WeightTensor& get(const std::string& name, int dtype, int d0, int d1 = -1) const;
  const __half* f16(const std::string& name, int d0, int d1 = -1) const {
    return static_cast<const __half*>(get(name, 0, d0, d1).dev);
  }
  const float* f32(const std::string& name, int d0, int d1 = -1) const {
    return static_cast<const float*>(get(name, 1, d0, d1).dev);
  }
  // the same tensor in the host copy of the file (plan builders read small vectors without touching the device)
  const __half* host_f16(const std::string& name, int d0, int d1 = -1) const {
    return reinterpret_cast<const __half*>(host_.data() + (static_cast<const uint8_t*>(get(name, 0, d0, d1).dev) -
                                                            static_cast<const uint8_t*>(blob_)));
  }
  // tensor by name without a shape expectation (nullptr when absent); host_ptr: its bytes in the host copy
  const WeightTensor* find(const std::string& name) const {
    auto it = tensors_.find(name);
    return it == tensors_.end() ? nullptr : &it->second;
  }
  const void* host_ptr(const WeightTensor& t) const {
    return host_.data() + (static_cast<const uint8_t*>(t.dev) - static_cast<const uint8_t*>(blob_));
  }
  const float* host_f32(const std::string& name, int d0, int d1 = -1) const {
    return reinterpret_cast<const float*>(host_.data() + (static_cast<const uint8_t*>(get(name, 1, d0, d1).dev) -
                                                           static_cast<const uint8_t*>(blob_)));
  }

 private:
  std::map<std::string, WeightTensor> tensors_;
  void* blob_ = nullptr;
  std::vector<uint8_t> host_;
  std::string path_;
};

// Activation storage shared by both networks (they never run concurrently).
struct Arena {
  int cap = 0;  // hypotheses
  // act[]: four backbone activation buffers handed out by the plan builder (each fits the largest activation of the
  // supported layer lists: the stem output, 2N x 80 x 80 x 64); d0: the tokens [N, 400, 512] the backbone ends in
  __half *packed = nullptr, *act[4] = {nullptr, nullptr, nullptr, nullptr}, *d0 = nullptr, *qkv = nullptr,
         *att = nullptr, *proj = nullptr, *x1 = nullptr, *ff1 = nullptr, *ff2 = nullptr, *x2 = nullptr;
  size_t act_bytes = 0;
  float *mean_t = nullptr, *mean_r = nullptr;
  float* mean_part = nullptr;  // partial sums of the sliced LayerNorm + token mean (few hypotheses)
  void* blob = nullptr;
  size_t bytes = 0;
  explicit Arena(int max_hyps);
  ~Arena();
};

enum class NetKind { kRefiner = 0, kScorer = 1 };

struct Step;  // one launch of the forward sequence

// One backbone layer of the weight file's "arch" table (foundationpose_cpp_b200/weights.py): section 0 = encodeA (on
// the 2N images), 1 = encodeAB (after the channel concat); kind 0 = 7x7 ConvBNReLU on the packed crop, 1 = 3x3
// ConvBNReLU, 2 = ResnetBasicBlock.
struct ArchLayer {
  int sec, kind, cout, stride;
};

class Network {
 public:
  Network(NetKind kind, const std::string& weight_path, Arena* arena);
  ~Network();
  // Runs the network on hypotheses [0,n) of arena->packed (render images 0..n-1, observed n..2n-1).
  // Refiner: leaves token means in arena->mean_t / mean_r.  Scorer: leaves the token mean of the
  // self-attention output (before out_proj) in arena->mean_t.
  void forward(int n, cudaStream_t s);
  double flops_per_hypothesis() const { return flops_per_hyp_; }
  // Profiling of the tcgen05 GEMM launches: when enabled every conv_gemm launch is bracketed by
  // CUDA events on the launching stream; collect_profile() sums their durations and FLOPs since
  // the last collect (it synchronises).
  void set_profile(bool on) { profile_ = on; }
  void collect_profile(double* ms, double* flops, long long* launches);
  int launches_per_forward(int n);
  const WeightFile& weights() const { return *w_; }

 private:
  struct Plan;
  Plan& plan(int n);
  void add_backbone(Plan& p, int n);
  void add_encoder_layer(Plan& p, int n, const std::string& head, float* mean_out);
  NetKind kind_;
  std::vector<ArchLayer> arch_;
  std::unique_ptr<WeightFile> w_;
  Arena* arena_;
  std::map<int, std::unique_ptr<Plan>> plans_;
  double flops_per_hyp_ = 0;
  bool profile_ = false;
  __half* stem_w_ = nullptr;  // stem weights in the direct-convolution kernel's tap order (stem_conv.cu)
  std::vector<cudaEvent_t> ev_pool_;
  size_t ev_used_ = 0;
  std::vector<double> ev_flops_;
};

}  // namespace fpb
