// ORACLE — test infrastructure only.  Thin C entry points over the REFERENCE's own CUDA sources,
// which oracle/Makefile compiles in place from /root/reference (nothing is copied into this
// repository): vendored nvdiffrast (cudaraster + rasterize/interpolate/texture kernels) and the
// reference's depth / sampling / decoder kernels.  Used by tests/test_ref_kernels.py on the GPU box to
// pin the CPU oracle (and the product kernels) against the reference's real arithmetic.
//
// The three launch helpers below set up the same kernel parameters as the reference's wrappers
// foundationpose_render::rasterize / interpolate / texture
// (detection_6d_foundationpose/src/foundationpose_render.cu:184-319); those wrappers themselves
// cannot be compiled here because their header pulls Eigen / assimp / OpenCV.
#include <cstdint>
#include <cuda_runtime.h>

#include "nvdiffrast/common/common.h"
#include "nvdiffrast/common/cudaraster/CudaRaster.hpp"
#include "nvdiffrast/common/interpolate.h"
#include "nvdiffrast/common/rasterize.h"
#include "nvdiffrast/common/texture.h"
#include "foundationpose_decoder.cu.hpp"
#include "foundationpose_sampling.cu.hpp"
#include "foundationpose_utils.cu.hpp"

void RasterizeCudaFwdShaderKernel(const RasterizeCudaFwdShaderParams p);
void InterpolateFwdKernel(const InterpolateKernelParams p);
void TextureFwdKernelLinear1(const TextureKernelParams p);

extern "C" {

// clip [N,V,4] f32, tri [T,3] i32 (device) -> rast [N,H,W,4] f32 (b0, b1, z/w, id+1); GL row order.
int ref_rasterize(float* clip, int32_t* tri, float* rast, int V, int T, int H, int W, int N) {
  CR::CudaRaster cr;
  cr.setViewportSize(W, H, N);
  cr.setVertexBuffer((void*)clip, V);
  cr.setIndexBuffer((void*)tri, T);
  cr.setRenderModeFlags(0);
  cr.deferredClear(0u);
  const bool ok = cr.drawTriangles(nullptr, false, 0);
  RasterizeCudaFwdShaderParams p;
  p.pos = clip;
  p.tri = tri;
  p.in_idx = (const int*)cr.getColorBuffer();
  p.out = rast;
  p.numTriangles = T;
  p.numVertices = V;
  p.width = W;
  p.height = H;
  p.depth = N;
  p.instance_mode = 1;
  p.xs = 2.f / (float)W;
  p.xo = 1.f / (float)W - 1.f;
  p.ys = 2.f / (float)H;
  p.yo = 1.f / (float)H - 1.f;
  dim3 block = getLaunchBlockSize(RAST_CUDA_FWD_SHADER_KERNEL_BLOCK_WIDTH, RAST_CUDA_FWD_SHADER_KERNEL_BLOCK_HEIGHT, W, H);
  dim3 grid = getLaunchGridSize(block, W, H, N);
  void* args[] = {&p};
  cudaLaunchKernel((void*)RasterizeCudaFwdShaderKernel, grid, block, args, 0, 0);
  const cudaError_t e = cudaDeviceSynchronize();
  return (ok && e == cudaSuccess) ? 0 : -1;
}

int ref_interpolate(float* attr, float* rast, int32_t* tri, float* out, int V, int T, int attr_shape_dim,
                    int attr_dim, int H, int W, int N) {
  InterpolateKernelParams p = {};
  p.instance_mode = attr_shape_dim > 2 ? 1 : 0;
  p.numVertices = V;
  p.numAttr = attr_dim;
  p.numTriangles = T;
  p.height = H;
  p.width = W;
  p.depth = N;
  p.attr = attr;
  p.rast = rast;
  p.tri = tri;
  p.attrBC = 0;
  p.out = out;
  dim3 block = getLaunchBlockSize(IP_FWD_MAX_KERNEL_BLOCK_WIDTH, IP_FWD_MAX_KERNEL_BLOCK_HEIGHT, W, H);
  dim3 grid = getLaunchGridSize(block, W, H, N);
  void* args[] = {&p};
  cudaLaunchKernel((void*)InterpolateFwdKernel, grid, block, args, 0, 0);
  return cudaDeviceSynchronize() == cudaSuccess ? 0 : -1;
}

int ref_texture(float* tex, float* uv, float* out, int th, int tw, int H, int W, int N) {
  TextureKernelParams p = {};
  p.enableMip = false;
  p.filterMode = TEX_MODE_LINEAR;
  p.boundaryMode = TEX_BOUNDARY_MODE_WRAP;
  p.texDepth = 1;
  p.texHeight = th;
  p.texWidth = tw;
  p.channels = 3;
  p.n = N;
  p.imgHeight = H;
  p.imgWidth = W;
  p.tex[0] = tex;
  p.uv = uv;
  p.mipLevelBias = NULL;
  p.out = out;
  dim3 block = getLaunchBlockSize(TEX_FWD_MAX_KERNEL_BLOCK_WIDTH, TEX_FWD_MAX_KERNEL_BLOCK_HEIGHT, W, H);
  dim3 grid = getLaunchGridSize(block, W, H, N);
  void* args[] = {&p};
  cudaLaunchKernel((void*)TextureFwdKernelLinear1, grid, block, args, 0, 0);
  return cudaDeviceSynchronize() == cudaSuccess ? 0 : -1;
}

int ref_depth_filters(float* depth, float* eroded, float* filtered, int H, int W) {
  erode_depth(0, depth, eroded, H, W);
  bilateral_filter_depth(0, eroded, filtered, H, W);
  return cudaDeviceSynchronize() == cudaSuccess ? 0 : -1;
}

int ref_depth_to_xyz(const float* depth, float* xyz, int H, int W, float fx, float fy, float cx, float cy,
                     float min_depth) {
  convert_depth_to_xyz_map((cudaStream_t)0, depth, H, W, xyz, fx, fy, cx, cy, min_depth);
  return cudaDeviceSynchronize() == cudaSuccess ? 0 : -1;
}

int ref_argmax(float* scores, int n) { return getMaxScoreIndex(0, scores, n); }

}  // extern "C"
