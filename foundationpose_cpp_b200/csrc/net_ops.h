// Non-GEMM network operators: attention core, residual LayerNorm, token means, small fp32 heads,
// pose update and ranking.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace fpb {

// Per-device kernel attributes of this file's kernels and of the attention kernel; nullptr on success.
const char* attention_init();
const char* net_ops_init();
// Self-attention core.  qkv: [n_hyp*400, qkv_ld] fp16 with Q | K | V blocks of 4 heads x 128; out: [n_hyp*400, out_ld].
// The plan (three TMA descriptors) belongs to whoever owns the buffers: a network plan, or a debug call.
struct AttentionPlan {
  CUtensorMap tm_q;    // qkv as (qkv_ld, 400, n_hyp), box (64, 128, 1)
  CUtensorMap tm_kv;   // same tensor, box (64, 80, 1)
  CUtensorMap tm_out;  // out as (out_ld, 400, n_hyp), box (64, 32, 1): one softmax warp's rows x one dim half
  int n_items = 0;     // n_hyp * heads * query tiles
  int grid_v1 = 0;     // CTAs of the first kernel: min(n_items, SMs of the plan's device)
  bool pdl = false;    // programmatic dependent launch (FPB200_PDL=1: measurements)
  unsigned long long* phase_clocks = nullptr;  // optional (FPB200_ATTN_PHASES=1): per-phase clock sums of CTA 0
  // second design (attention2.cu: K / V resident in shared memory, Q in tensor memory, S computed twice)
  CUtensorMap tm_kv64;  // qkv, box (64, 64, 1)
  CUtensorMap tm_kv16;  // qkv, box (64, 16, 1): keys 384..399
  const __half* qkv = nullptr;
  __half* out = nullptr;
  int qkv_ld = 0, out_ld = 0;
  int tiles_per_cta = 4;  // query tiles of one (hypothesis, head) walked by one CTA: 4, 2 or 1
  int tpc_log2 = 2;
  int n_units = 0;        // n_hyp * heads * (4 / tiles_per_cta)
  int grid = 0;
  int version = 2;        // 1: attention.cu (FPB200_ATTN_V1=1), 2: attention2.cu
};
const char* build_attention(const __half* qkv, __half* out, int n_hyp, int qkv_ld, int out_ld, AttentionPlan* plan);
void launch_attention(const AttentionPlan& plan, cudaStream_t s);
bool attention_phase_clocks(unsigned long long out[22]);

// out = LayerNorm(x) * gamma + beta over 512 channels, eps 1e-5 (post-LN of nn.TransformerEncoderLayer;
// the residual sum x = a + sublayer(a) is formed by the producing GEMM's epilogue).  [rows,512] fp16.
void launch_layernorm(const __half* x, const float* gamma, const float* beta, __half* out, long long rows,
                      cudaStream_t s);
// out[n,512] f32 = mean over the 400 tokens of LayerNorm(x[n,400,512]); the normalised tokens are not stored.
// partial: n * kLnMeanSlices * 512 floats of scratch.  The tokens of a hypothesis are always summed as
// kLnMeanSlices slices in a fixed order (the result does not depend on n); with few hypotheses the slices
// are spread over up to kLnMeanSlices CTAs and a second kernel adds them (one hypothesis on one SM took
// 24 us).  Returns the number of kernels launched.
constexpr int kLnMeanSlices = 8;
int launch_layernorm_mean(const __half* x, const float* gamma, const float* beta, float* out, int n,
                          float* partial, cudaStream_t s);
void launch_token_mean(const __half* x, float* out, int n, cudaStream_t s);
// out[r,o] = sum_k in[r,k] * w[o,k] + b[o]   (fp32, small shapes)
void launch_small_linear(const float* in, const float* w, const float* b, float* out, int rows, int k,
                         int outs, cudaStream_t s);
// Multi-head attention over a single sequence of n rows: qkv [n,1536] f32 -> out [n,512] f32.
void launch_cross_attention(const float* qkv, float* out, int n, cudaStream_t s);
// argmax (first maximum, like thrust::max_element in foundationpose_decoder.cu:24-35) and copy of
// the winning pose; best_index/best_pose live in device memory.
void launch_argmax_pose(const float* scores, const float* poses, int n, int* best_index, float* best_pose,
                        cudaStream_t s);
// RefinePostProcess (foundationpose.cpp:360-406) on the device: heads = mean tokens x Linear(512,3),
// then t += trans * d/2, R <- Rodrigues(tanh(rot) * 0.349066)^T R.  mean_t / mean_r: [n,512] f32.
void launch_refine_update(const float* mean_t, const float* mean_r, const float* wt, const float* bt,
                          const float* wr, const float* br, float diameter, float* poses, float* trans_out,
                          float* rot_out, int n, cudaStream_t s);

}  // namespace fpb
