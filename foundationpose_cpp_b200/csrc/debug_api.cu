// Stage-level C-ABI entry points used by the parity tests and micro-benchmarks: they run one
// kernel of the hot path on caller-provided DEVICE pointers.  Declared in include/fpb200.h.
#include <cuda_runtime.h>

#include <stdexcept>
#include <vector>

#include "../../include/fpb200.h"
#include "common.h"
#include "conv_gemm.h"
#include "encoder_block.h"
#include "stem_conv.h"
#include "net_ops.h"

using namespace fpb;

extern "C" int fpb200_debug_conv_gemm(const fpb200_conv_desc* c, int repeat, float* ms_out) {
  try {
    ConvGemmDesc d;
    d.in = static_cast<const __half*>(c->in);
    d.n_img = c->n_img;
    d.planes = c->planes;
    d.in_h = c->in_h;
    d.in_w = c->in_w;
    d.cin = c->cin;
    d.stride_w = c->stride_w;
    d.stride_h = c->stride_h;
    d.stride_plane = c->stride_plane;
    d.stride_img = c->stride_img;
    d.out_h = c->out_h;
    d.out_w = c->out_w;
    d.bw = c->bw;
    d.bh = c->bh;
    d.bn = c->bn;
    switch (c->tap_kind) {
      case FPB200_TAPS_1X1: d.num_taps = 1; break;
      case FPB200_TAPS_3X3_S1: taps_3x3_s1(&d); break;
      case FPB200_TAPS_3X3_S2_PHASE: taps_3x3_s2_phase(&d); break;
      case FPB200_TAPS_STEM7_ROWS: taps_stem7_rows(&d); break;
      default: return fail(FPB200_ERR_INVALID, "unknown tap_kind %d", c->tap_kind);
    }
    d.weight = static_cast<const __half*>(c->weight);
    d.cout = c->cout;
    d.bias = c->bias;
    d.residual = static_cast<const __half*>(c->residual);
    d.res_ld = c->res_ld;
    d.post_add = c->post_add;
    d.out = static_cast<__half*>(c->out);
    d.relu = c->relu != 0;
    d.out_phase_split = c->out_phase_split != 0;
    d.out_ld = c->out_ld;
    d.img_mod = c->img_mod;
    d.half_coff = c->half_coff;
    d.max_ctas = c->max_ctas;
    d.cluster = c->cluster;
    d.block_n = c->block_n;
    ConvGemmPlan plan;
    if (const char* e = build_conv_gemm(d, &plan)) return fail(FPB200_ERR_INVALID, "%s", e);
    cudaEvent_t e0, e1;
    FPB_CUDA(cudaEventCreate(&e0));
    FPB_CUDA(cudaEventCreate(&e1));
    FPB_CUDA(launch_conv_gemm(plan, 0));  // warm-up + the result the caller checks
    FPB_CUDA(cudaEventRecord(e0, 0));
    for (int i = 0; i < repeat; ++i) FPB_CUDA(launch_conv_gemm(plan, 0));
    FPB_CUDA(cudaEventRecord(e1, 0));
    FPB_CUDA(cudaDeviceSynchronize());
    float ms = 0.f;
    FPB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ms_out) *ms_out = repeat > 0 ? ms / repeat : 0.f;
    return FPB200_OK;
  } catch (const std::exception& e) {
    return fail(FPB200_ERR_CUDA, "%s", e.what());
  }
}

extern "C" int fpb200_debug_attention(const void* qkv, void* out, int n_hyp, int repeat, float* ms_out) {
  try {
    if (const char* e = attention_init()) return fail(FPB200_ERR_CUDA, "%s", e);
    const __half* q = static_cast<const __half*>(qkv);
    __half* o = static_cast<__half*>(out);
    cudaEvent_t e0, e1;
    FPB_CUDA(cudaEventCreate(&e0));
    FPB_CUDA(cudaEventCreate(&e1));
    AttentionPlan ap;
    if (const char* e = build_attention(q, o, n_hyp, 1536, 512, &ap)) return fail(FPB200_ERR_INVALID, "%s", e);
    launch_attention(ap, 0);
    FPB_CUDA(cudaGetLastError());
    FPB_CUDA(cudaEventRecord(e0, 0));
    for (int i = 0; i < repeat; ++i) launch_attention(ap, 0);
    FPB_CUDA(cudaEventRecord(e1, 0));
    FPB_CUDA(cudaDeviceSynchronize());
    float ms = 0.f;
    FPB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ms_out) *ms_out = repeat > 0 ? ms / repeat : 0.f;
    return FPB200_OK;
  } catch (const std::exception& e) {
    return fail(FPB200_ERR_CUDA, "%s", e.what());
  }
}

namespace {
template <typename F>
int timed_launches(F&& launch, int repeat, float* ms_out) {
  try {
    cudaEvent_t e0, e1;
    FPB_CUDA(cudaEventCreate(&e0));
    FPB_CUDA(cudaEventCreate(&e1));
    launch();
    FPB_CUDA(cudaGetLastError());
    FPB_CUDA(cudaEventRecord(e0, 0));
    for (int i = 0; i < repeat; ++i) launch();
    FPB_CUDA(cudaEventRecord(e1, 0));
    FPB_CUDA(cudaDeviceSynchronize());
    float ms = 0.f;
    FPB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ms_out) *ms_out = repeat > 0 ? ms / repeat : 0.f;
    return FPB200_OK;
  } catch (const std::invalid_argument& e) {
    return fail(FPB200_ERR_INVALID, "%s", e.what());
  } catch (const std::exception& e) {
    return fail(FPB200_ERR_CUDA, "%s", e.what());
  }
}
}  // namespace

extern "C" int fpb200_debug_small_linear(const float* in, const float* w, const float* b, float* out, int rows, int k,
                                         int outs, int repeat, float* ms_out) {
  return timed_launches([&]() { launch_small_linear(in, w, b, out, rows, k, outs, 0); }, repeat, ms_out);
}

extern "C" int fpb200_debug_cross_attention(const float* qkv, float* out, int n, int repeat, float* ms_out) {
  return timed_launches([&]() { launch_cross_attention(qkv, out, n, 0); }, repeat, ms_out);
}

extern "C" int fpb200_debug_attention_phases(unsigned long long out[22]) {
  return attention_phase_clocks(out) ? FPB200_OK
                                     : fail(FPB200_ERR_STATE, "set FPB200_ATTN_PHASES=1 before the first attention launch");
}

extern "C" const char* fpb200_last_error(void) { return last_error().c_str(); }

extern "C" int fpb200_debug_encoder_block(const fpb200_encoder_desc* c, int repeat, float* ms_out) {
  if (!c) return fail(FPB200_ERR_INVALID, "null descriptor");
  EncoderBlockDesc d;
  d.att = static_cast<const __half*>(c->att);
  d.x0 = static_cast<const __half*>(c->x0);
  d.w_out = static_cast<const __half*>(c->w_out);
  d.w1 = static_cast<const __half*>(c->w1);
  d.w2 = static_cast<const __half*>(c->w2);
  d.b_out = c->b_out; d.b1 = c->b1; d.b2 = c->b2;
  d.ln1_g = c->ln1_g; d.ln1_b = c->ln1_b; d.ln2_g = c->ln2_g; d.ln2_b = c->ln2_b;
  d.partial = c->partial; d.mean_out = c->mean_out; d.n_hyp = c->n_hyp; d.mode = c->mode;
  d.dbg_x1 = static_cast<__half*>(c->dbg_x1);
  d.dbg_h = static_cast<__half*>(c->dbg_h);
  d.dbg_y = static_cast<__half*>(c->dbg_y);
  d.dbg_clocks = static_cast<unsigned long long*>(c->dbg_clocks);
  EncoderBlockPlan plan;
  if (const char* e = build_encoder_block(d, &plan)) return fail(FPB200_ERR_INVALID, "%s", e);
  return timed_launches([&]() { FPB_CUDA(launch_encoder_block(plan, 0)); }, repeat, ms_out);
}

// Stem convolution alone: packed [n_img][160][164][8] fp16 (+ one pad row) -> phase-split [n_img][4][40][40][64] fp16.
// weight: DEVICE pointer, file layout of `enc.stem.w` ([64][448], layouts.pack_stem7); bias: device [64] f32.
extern "C" int fpb200_debug_stem_conv(const void* packed, int n_img, const void* weight, const float* bias, void* out,
                                      int repeat, float* ms_out) {
  try {
    std::vector<__half> w_file(64 * 448), w_kernel(64 * 448);
    FPB_CUDA(cudaMemcpy(w_file.data(), weight, w_file.size() * sizeof(__half), cudaMemcpyDeviceToHost));
    reorder_stem_weights(w_file.data(), w_kernel.data());
    __half* w_dev = nullptr;
    FPB_CUDA(cudaMalloc(&w_dev, w_kernel.size() * sizeof(__half)));
    FPB_CUDA(cudaMemcpy(w_dev, w_kernel.data(), w_kernel.size() * sizeof(__half), cudaMemcpyHostToDevice));
    float bias_host[64];
    FPB_CUDA(cudaMemcpy(bias_host, bias, sizeof(bias_host), cudaMemcpyDeviceToHost));
    StemConvPlan plan;
    const char* e = build_stem_conv(static_cast<const __half*>(packed), n_img, w_dev, bias_host, static_cast<__half*>(out), &plan);
    int rc = FPB200_OK;
    if (e) rc = fail(FPB200_ERR_INVALID, "%s", e);
    else rc = timed_launches([&]() { FPB_CUDA(launch_stem_conv(plan, 0)); }, repeat, ms_out);
    cudaDeviceSynchronize();
    cudaFree(w_dev);
    return rc;
  } catch (const std::exception& ex) {
    return fail(FPB200_ERR_CUDA, "%s", ex.what());
  }
}
