// ORACLE — test infrastructure only.  Runs the REFERENCE's own render path on caller-provided device buffers:
// the kernels and wrappers of detection_6d_foundationpose/src/foundationpose_render.cu (compiled in place by
// oracle/Makefile with its Eigen/assimp/OpenCV-pulling header skipped through its include guard) in the order of
// FoundationPoseRenderer::NvdiffrastRender + RenderProcess
// (detection_6d_foundationpose/src/foundationpose_render.cpp:611-683,685-729):
//   transform_points (K4) -> generate_pose_clip (K5) -> rasterize (cudaraster K7-K10 + K11) -> interpolate xyz ->
//   interpolate uv -> texture -> transform_normals (K6) -> interpolate diffuse -> refine_color (K14) -> clamp (K15)
//   -> threshold_and_downscale_pointcloud (K16) -> concat (K17).
// The two cvcuda::Flip calls between K15 and K16 are CV-CUDA (absent here): the caller flips rows on the host.
// Used (a) to pin fp_oracle.py and the product's fused render kernel against the reference's arithmetic, and
// (b) by bench.py as the same-box "reference CUDA preprocess" baseline.
#include <cstdint>
#include <cuda_runtime.h>

#include "foundationpose_render.cu.hpp"

extern "C" {

// All pointers are device memory.  pose / mvp: [N,16] column-major (pose = centred mesh -> camera,
// mvp = Proj * GLcam * pose); bbox2d [N,4]; outputs and scratch as named.  With repeat > 0 the whole sequence is
// run `repeat` more times between two CUDA events and the average is returned in *ms_out.
int ref_render_sequence(float* pose, float* mvp, float* bbox2d, float* verts, float* normals, float* texcoords,
                        int32_t* faces, float* texture, float* pts_cam, float* pose_clip, float* diffuse, float* rast,
                        float* xyz_map, float* uv, float* color, float* diffuse_map, float* render_input, int V, int T,
                        int th, int tw, int N, int H, int W, int rgb_H, int rgb_W, float diameter, int repeat,
                        float* ms_out) {
  cudaStream_t s = 0;
  CR::CudaRaster cr;
  auto once = [&]() {
    foundationpose_render::transform_points(s, pose, N, verts, V, pts_cam);
    foundationpose_render::generate_pose_clip(s, mvp, bbox2d, N, verts, V, pose_clip, rgb_H, rgb_W);
    foundationpose_render::rasterize(s, &cr, pose_clip, faces, rast, V, T, H, W, N);
    foundationpose_render::interpolate(s, pts_cam, rast, faces, xyz_map, V, T, 3, 3, H, W, N);
    foundationpose_render::interpolate(s, texcoords, rast, faces, uv, V, T, 2, 2, H, W, N);
    foundationpose_render::texture(s, texture, uv, color, th, tw, 3, 1, H, W, N);
    foundationpose_render::transform_normals(s, pose, N, normals, V, diffuse);
    foundationpose_render::interpolate(s, diffuse, rast, faces, diffuse_map, V, T, 3, 1, H, W, N);
    foundationpose_render::refine_color(s, color, diffuse_map, rast, color, N, 0.8f, 0.5f, H, W);
    foundationpose_render::clamp(s, color, 0.0f, 1.0f, N * H * W * 3);
    // (cvcuda::Flip of color and xyz_map happens here in the reference)
    foundationpose_render::threshold_and_downscale_pointcloud(s, xyz_map, pose, N, H * W, diameter / 2.0f, 0.001f, 4.0f);
    foundationpose_render::concat(s, color, xyz_map, render_input, N, H, W, 3, 3);
  };
  once();
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  if (repeat > 0) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, s);
    for (int i = 0; i < repeat; ++i) once();
    cudaEventRecord(e1, s);
    if (cudaEventSynchronize(e1) != cudaSuccess) return -2;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms_out) *ms_out = ms / repeat;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

}  // extern "C"
