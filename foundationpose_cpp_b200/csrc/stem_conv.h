// Stem of both networks (7 x 7, stride 2, 6 -> 64 channels on the packed 160 x 160 crops) as a dedicated tcgen05
// kernel whose A operand is read IN PLACE from the input patch (stem_conv.cu).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace fpb {

struct StemConvParams {
  CUtensorMap tm_in;  // packed input [img][160 rows][164 px][8 ch] fp16, box (8 ch, 22 px at element stride 2, 37 rows)
  CUtensorMap tm_w;   // reordered weights [64 out][448] fp16 K-major, box (64, 64), 128-byte swizzle
  float bias[64];     // by value: the epilogue reads it through the constant bank (one broadcast operand per add)
  __half* out;        // phase-split [img][(y&1)*2+(x&1)][40][40][64] fp16 (ReLU applied)
  int n_tiles;        // images * 50 (tiles of 8 x 16 output pixels)
  int experiment;     // measurements only: 1 = no output stores, 2 = no MMAs, 3 = no TMA patch loads
};

struct StemConvPlan {
  StemConvParams p;
  int grid = 0;
  double flops = 0;  // algorithmic: 2 * 7*7*6 MACs per output element
};

const char* stem_conv_init();
// file layout of `enc.stem.w` ([64][7 rows][8 px: zero, kx 0..6][8 ch], layouts.pack_stem7) -> the kernel's K order
// ([64][7 rows][kx 0, 2, 4, 6, 1, 3, 5, zero][8 ch]); host buffers of 64 * 448 halves.
void reorder_stem_weights(const __half* file_layout, __half* kernel_layout);
// bias_host: HOST pointer (64 floats); everything else device memory.  Plan building never touches the device.
const char* build_stem_conv(const __half* packed, int n_img, const __half* w_kernel_layout, const float* bias_host, __half* out,
                            StemConvPlan* plan);
cudaError_t launch_stem_conv(const StemConvPlan& plan, cudaStream_t stream);

}  // namespace fpb
