// Stand-alone inference cores: the refine / score networks behind the plugin boundary the reference
// uses for its TensorRT engines — inference_core::BaseInferCore::SyncInfer on device-located blobs
// (easy_deploy_tool/deploy_core/include/deploy_core/base_infer_core.h:179-237; blob names and layout
// detection_6d_foundationpose/src/foundationpose.cpp:78-83,349-354: render_input / transf_input f32
// NHWC [n,160,160,6] -> trans [n,3], rot [n,3] (refiner) or scores [n] (scorer, n <= 252)).
// A `B200InferCore : BaseInferCore` in the reference tree forwards PreProcess/Inference/PostProcess
// to these calls; the session API (session.cu) does not go through them, it keeps everything packed.
#include <memory>
#include <string>

#include "../../include/fpb200.h"
#include "common.h"
#include "net_ops.h"
#include "network.h"
#include "preprocess.h"

using namespace fpb;

struct fpb200_infer_core {
  int kind = 0, max_batch = 0, device = 0;
  cudaStream_t stream = nullptr;
  std::unique_ptr<Arena> arena;
  std::unique_ptr<Network> net;
  float *feats = nullptr, *qkv = nullptr, *att = nullptr, *out = nullptr;
  ~fpb200_infer_core() {
    for (float* p : {feats, qkv, att, out})
      if (p) cudaFree(p);
    if (stream) cudaStreamDestroy(stream);
  }
};

namespace {
template <typename F>
int guarded(F&& f) {
  try {
    return f();
  } catch (const std::invalid_argument& e) {
    return fail(FPB200_ERR_INVALID, "%s", e.what());
  } catch (const CudaError& e) {
    return fail(FPB200_ERR_CUDA, "%s", e.what());
  } catch (const std::exception& e) {
    return fail(FPB200_ERR_IO, "%s", e.what());
  }
}

void forward_inputs(fpb200_infer_core* c, const float* render_input, const float* transf_input, int n) {
  FPB_REQUIRE(c && render_input && transf_input, "null argument");
  FPB_REQUIRE(n >= 1 && n <= c->max_batch, "batch %d outside [1,%d]", n, c->max_batch);
  FPB_CUDA(cudaSetDevice(c->device));
  launch_repack_f32(render_input, n, c->arena->packed, 0, c->stream);
  launch_repack_f32(transf_input, n, c->arena->packed, n, c->stream);
  c->net->forward(n, c->stream);
}
}  // namespace

extern "C" {

int fpb200_infer_create(const char* weights, int kind, int max_batch, int device, fpb200_infer_core** out) {
  return guarded([&]() -> int {
    FPB_REQUIRE(weights && out, "null argument");
    FPB_REQUIRE(kind == 0 || kind == 1, "kind must be 0 (refiner) or 1 (scorer)");
    FPB_REQUIRE(max_batch >= 1, "max_batch must be >= 1");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
      throw CudaError("no CUDA device available: this library has no CPU fallback");
    FPB_CUDA(cudaSetDevice(device));
    std::unique_ptr<fpb200_infer_core> c(new fpb200_infer_core);
    c->kind = kind;
    c->max_batch = max_batch;
    c->device = device;
    FPB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->arena.reset(new Arena(max_batch));
    c->net.reset(new Network(kind == 0 ? NetKind::kRefiner : NetKind::kScorer, weights, c->arena.get()));
    FPB_CUDA(cudaMalloc(&c->feats, sizeof(float) * 512 * max_batch));
    FPB_CUDA(cudaMalloc(&c->qkv, sizeof(float) * 1536 * max_batch));
    FPB_CUDA(cudaMalloc(&c->att, sizeof(float) * 512 * max_batch));
    FPB_CUDA(cudaMalloc(&c->out, sizeof(float) * 512 * max_batch));
    *out = c.release();
    return FPB200_OK;
  });
}

// Device buffers for callers that hold blobs but do not link the CUDA runtime themselves (the B200InferCore plugin
// of cpp/b200_infer_core.hpp is compiled with the host compiler only).
int fpb200_device_alloc(size_t bytes, int device, void** out) {
  return guarded([&]() -> int {
    FPB_REQUIRE(out, "null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
      throw CudaError("no CUDA device available: this library has no CPU fallback");
    FPB_CUDA(cudaSetDevice(device));
    FPB_CUDA(cudaMalloc(out, bytes ? bytes : 1));
    FPB_CUDA(cudaMemset(*out, 0, bytes));
    return FPB200_OK;
  });
}

void fpb200_device_free(void* ptr) {
  if (ptr) cudaFree(ptr);
}

int fpb200_memcpy(void* dst, const void* src, size_t bytes, int kind) {
  return guarded([&]() -> int {
    FPB_REQUIRE(dst && src, "null argument");
    const cudaMemcpyKind k = kind == FPB200_COPY_H2D ? cudaMemcpyHostToDevice
                             : kind == FPB200_COPY_D2H ? cudaMemcpyDeviceToHost
                                                       : cudaMemcpyDeviceToDevice;
    FPB_REQUIRE(kind >= FPB200_COPY_H2D && kind <= FPB200_COPY_D2D, "bad copy kind %d", kind);
    FPB_CUDA(cudaMemcpy(dst, src, bytes, k));
    return FPB200_OK;
  });
}

void fpb200_infer_destroy(fpb200_infer_core* c) {
  if (!c) return;
  cudaStreamSynchronize(c->stream);
  delete c;
}

int fpb200_infer_refiner(fpb200_infer_core* c, const float* render_input, const float* transf_input, int n,
                         float* trans, float* rot) {
  return guarded([&]() -> int {
    FPB_REQUIRE(c && c->kind == 0, "not a refiner core");
    FPB_REQUIRE(trans && rot, "null output");
    forward_inputs(c, render_input, transf_input, n);
    const WeightFile& W = c->net->weights();
    launch_small_linear(c->arena->mean_t, W.f32("trans.head.w", 3, 512), W.f32("trans.head.b", 3), trans, n, 512, 3, c->stream);
    launch_small_linear(c->arena->mean_r, W.f32("rot.head.w", 3, 512), W.f32("rot.head.b", 3), rot, n, 512, 3, c->stream);
    FPB_CUDA(cudaStreamSynchronize(c->stream));  // SyncInfer semantics
    return FPB200_OK;
  });
}

int fpb200_infer_scorer(fpb200_infer_core* c, const float* render_input, const float* transf_input, int n,
                        float* scores) {
  return guarded([&]() -> int {
    FPB_REQUIRE(c && c->kind == 1, "not a scorer core");
    FPB_REQUIRE(scores, "null output");
    forward_inputs(c, render_input, transf_input, n);
    const WeightFile& W = c->net->weights();
    launch_small_linear(c->arena->mean_t, W.f32("att.out.w", 512, 512), W.f32("att.out.b", 512), c->feats, n, 512, 512, c->stream);
    launch_small_linear(c->feats, W.f32("cross.qkv.w", 1536, 512), W.f32("cross.qkv.b", 1536), c->qkv, n, 512, 1536, c->stream);
    launch_cross_attention(c->qkv, c->att, n, c->stream);
    launch_small_linear(c->att, W.f32("cross.out.w", 512, 512), W.f32("cross.out.b", 512), c->out, n, 512, 512, c->stream);
    launch_small_linear(c->out, W.f32("lin.w", 1, 512), W.f32("lin.b", 1), scores, n, 512, 1, c->stream);
    FPB_CUDA(cudaStreamSynchronize(c->stream));
    return FPB200_OK;
  });
}

}  // extern "C"
