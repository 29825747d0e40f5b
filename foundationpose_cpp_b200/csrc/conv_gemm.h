// Implicit-GEMM convolution / linear layer on tcgen05 tensor cores (sm_100a).
//
// One kernel serves every GEMM-shaped layer of the refine and score networks (SURVEY.md §2.3;
// the reference runs them inside TensorRT: easy_deploy_tool/inference_core/trt_core/src/trt_core.cpp:323-371):
//   * 3x3 / 7x7 convolutions, stride 1 or 2, NHWC fp16 activations, as a sum over filter taps of
//     [128 pixels x 64 channels] x [64 channels x Cout] products.  Taps that differ only in their
//     vertical offset share ONE activation block in shared memory (a TMA box of bh + span tile rows
//     of a 5-D view (C, W, image, H, plane) of the input; zero padding is TMA out-of-bounds fill)
//     and address it through shifted operand descriptors, which cuts the L2 -> SM activation
//     traffic of a 3x3 layer 2.4x.  Stride-2 layers read an input that its producer wrote phase-split
//     (plane = (y&1)*2 + (x&1)), which makes every tap a unit-stride box again.
//   * linear layers (1 tap, W = number of rows).
// Epilogue (fused): + bias, + residual, ReLU, + positional table, fp16 store with optional
// phase-split / channel-concat addressing.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace fpb {

constexpr int kGemmBlockM = 128;
constexpr int kGemmBlockK = 64;
constexpr int kMaxTaps = 16;

constexpr int kMaxGroups = 8;     // tap groups per layer (taps that share dx and plane)
constexpr int kMaxGroupTaps = 4;  // taps per group (differ only in dy)

struct ConvGemmParams {
  CUtensorMap tm_a;  // activations, 5-D (C, W, image, H, plane), box (64, bw, bn, bh + span, 1)
  CUtensorMap tm_b;  // weights, 2-D (K_total, Cout), box (64, block_n or block_n / 2)
  // M tiling: a tile is bw x bh pixels of bn consecutive images (bw*bh*bn == 128); GEMM row
  // r = (y * bn + n) * bw + x, so that one tile row (bn*bw pixels) is a whole number of 1024-byte
  // swizzle atoms and a vertical filter offset is an aligned shift of the operand start address.
  int tiles_x, tiles_y, tiles_img;
  int bw, bh, bn;
  int bw_log2, bn_log2;
  int W, H, n_img;  // output extent (validity of rows)
  int n_tiles;      // Cout / block_n
  int cblocks;      // input channel blocks of 64 per tap
  // K loop = (64-channel block, tap group) items; the taps of a group read the same activation
  // block (bh + span tile rows, loaded ONCE) at different row offsets, each with its own weight tile.
  // Weight stages (one tile, or all tiles of a group) are the only synchronised ring: an activation
  // block rides on the full-barrier of its group's first weight stage, and its slot is known to be
  // free again because the ring depths satisfy (weight stages in any na-1 consecutive groups) >= nb.
  int num_groups;
  int per_tap;      // 1: one weight stage per tap; 0: one per group
  int a_bytes;      // bytes of one activation block
  int row_bytes;    // bytes of one tile row (bn * bw * 128)
  int na, nb;       // ring depths: activation blocks / weight stages
  int b_stage_bytes;
  int8_t g_dx[kMaxGroups], g_dy[kMaxGroups], g_plane[kMaxGroups], g_ntaps[kMaxGroups];
  uint32_t g_aoff[kMaxGroups];  // 4 x 8 bit: tile-row offset of each tap inside the block
  uint32_t g_kidx[kMaxGroups];  // 4 x 8 bit: index of each tap in the weight tensor
  // epilogue
  const float* bias;       // [Cout] fp32
  const __half* residual;  // [n_img, H, W, res_ld] fp16 or nullptr
  const float* post_add;   // [H*W, Cout] fp32, added after ReLU, or nullptr
  __half* out;
  int relu;
  int cout;
  int res_ld;
  int out_phase_split;  // write [img][4 planes][H/2][W/2][out_ld]
  int out_ld;           // channels per pixel of the destination tensor
  int img_mod;          // images per half: image n is written as image n % img_mod ...
  int half_coff;        // ... at channel offset (n / img_mod) * half_coff
};

// Host-side description of a layer; build_conv_gemm() turns it into launchable params.
struct ConvGemmDesc {
  const __half* in = nullptr;  // device
  // Input view [n_img][planes][in_h][in_w][cin].  Byte strides default to dense when 0.
  int n_img = 0, planes = 1, in_h = 0, in_w = 0, cin = 0;
  long long stride_w = 0, stride_h = 0, stride_plane = 0, stride_img = 0;
  // output extent
  int out_h = 0, out_w = 0;
  int bw = 8, bh = 8, bn = 2;
  int num_taps = 1;
  int8_t tap_dx[kMaxTaps] = {0}, tap_dy[kMaxTaps] = {0}, tap_plane[kMaxTaps] = {0};
  const __half* weight = nullptr;  // [cout][num_taps*cin] fp16, K-major
  int cout = 0;
  const float* bias = nullptr;
  const __half* residual = nullptr;
  int res_ld = 0;  // 0 -> cout
  const float* post_add = nullptr;
  __half* out = nullptr;
  bool relu = false;
  bool out_phase_split = false;
  int out_ld = 0;   // 0 -> cout
  int img_mod = 0;  // 0 -> n_img
  int half_coff = 0;
  int max_ctas = 0;  // 0 -> number of SMs
  int cluster = 0;   // 0 -> automatic; 1 = single CTAs; 2, 3 = CTA pairs issuing
                     // tcgen05.mma.cta_group::2 (M = 256)
  int block_n = 0;   // 0 -> automatic; 64 / 128 / 256 forces the tile width (tests run every kernel variant)
};

struct ConvGemmPlan {
  ConvGemmParams p;
  int block_n = 0;
  int cluster = 1;  // CTAs per cluster
  int mode = 0;     // 0 single CTAs, 2 CTA pairs + cta_group::2
  int grid = 0;
  size_t smem = 0;
  bool pdl = false;  // launch with programmatic stream serialisation (the kernel waits itself before reading its input)
  double flops = 0;  // 2*M*N*K over valid rows
};

// Tap tables for the layer kinds of the networks.
void taps_3x3_s1(ConvGemmDesc* d);        // planes == 1
void taps_3x3_s2_phase(ConvGemmDesc* d);  // input phase-split, planes == 4
void taps_stem7_rows(ConvGemmDesc* d);    // 7 row taps over a y-parity-split, x-overlapping view

// Returns nullptr on success, else a static error string.
const char* build_conv_gemm(const ConvGemmDesc& d, ConvGemmPlan* plan);
cudaError_t launch_conv_gemm(const ConvGemmPlan& plan, cudaStream_t stream);
// Sets kernel attributes (max dynamic smem), resolves cuTensorMapEncodeTiled; once per process.
const char* conv_gemm_init();

}  // namespace fpb
