// Fused tail of a post-LN transformer encoder layer (out-projection + residual + LayerNorm, FFN, residual +
// LayerNorm, token mean) as ONE tcgen05 kernel that keeps a 128-token row block on chip through its three GEMMs.
// See encoder_block.cu.  In the reference these layers run inside the refiner's TensorRT engine
// (easy_deploy_tool/inference_core/trt_core/src/trt_core.cpp:323-371).
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace fpb {

// Bias / LayerNorm vectors of the layer, passed by value inside the kernel parameters: the epilogue threads all read
// the same element at the same time, which the constant bank serves as one broadcast (LDC) without a global load.
struct EncoderVectors {
  float b_out[512], b1[512], b2[512], ln1_g[512], ln1_b[512], ln2_g[512], ln2_b[512];
};

struct EncoderBlockParams {
  CUtensorMap tm_att;   // attention output [rows, 512] fp16, box (64, 128)
  CUtensorMap tm_x0;    // layer input (residual of the out-projection) [rows, 512] fp16, box (64, 128)
  CUtensorMap tm_w[3];  // Wout, W1, W2: [512 out][512 in] fp16 K-major, box (64, 128 or 64)
  EncoderVectors vec;
  float* partial;       // [rows / 16][512] f32: sums of LayerNorm2 outputs over groups of 16 consecutive tokens
  int rows;             // hypotheses * 400 (a multiple of 16)
  __half *dbg_x1, *dbg_h, *dbg_y;  // tests: LayerNorm1 output, FFN activation, LayerNorm2 output [rows, 512]; or null
  unsigned long long* dbg_clocks;  // optional [24]: phase clock sums of CTA 0 (MMA issuer [0,8), one epilogue thread [8,24))
};

struct EncoderBlockDesc {
  const __half* att = nullptr;
  const __half* x0 = nullptr;
  const __half *w_out = nullptr, *w1 = nullptr, *w2 = nullptr;
  const float *b_out = nullptr, *b1 = nullptr, *b2 = nullptr;
  const float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;
  float* partial = nullptr;   // n_hyp * 25 * 512 floats of scratch
  float* mean_out = nullptr;  // [n_hyp, 512] f32: token mean of the layer output
  int n_hyp = 0;
  int mode = 0;               // 0 automatic, 1 single CTAs, 2 CTA pairs (tcgen05.mma.cta_group::2)
  __half *dbg_x1 = nullptr, *dbg_h = nullptr, *dbg_y = nullptr;
  unsigned long long* dbg_clocks = nullptr;
  bool vectors_on_host = false;  // b_* / ln* point to host memory (plan building then never touches the device)
};

struct EncoderBlockPlan {
  EncoderBlockParams p;
  bool pair = false;
  bool pdl = false;  // programmatic dependent launch (FPB200_PDL=1: measurements)
  int grid = 0;
  size_t smem = 0;
  int n_hyp = 0;
  float* mean_out = nullptr;
  double flops = 0;  // 3 GEMMs over the valid rows
};

const char* encoder_block_init();  // per-device kernel attributes; nullptr on success
const char* build_encoder_block(const EncoderBlockDesc& d, EncoderBlockPlan* plan);
// The fused kernel followed by the 25-group token-mean reduction (2 launches).
cudaError_t launch_encoder_block(const EncoderBlockPlan& plan, cudaStream_t stream);

}  // namespace fpb
