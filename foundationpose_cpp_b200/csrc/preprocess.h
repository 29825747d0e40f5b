// RGB-D preprocessing, hypothesis sampling, fused render and crop/warp kernels (preprocess.cu).
// Everything here is fp32 arithmetic written in the reference's operation order and compiled
// with -fmad=false, so that integer decisions (window rounding, sub-pixel snapping, coverage,
// depth test) are bit-identical to the CPU oracle; the packed result is stored as fp16.
#pragma once
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace fpb {

constexpr int kCrop = 160;
constexpr int kPackC = 8;                 // rgb(3) + xyz(3) + 2 zero channels
constexpr int kStemPad = 4;               // zero pixels stored left of every row
constexpr int kStemPitch = kCrop + kStemPad;  // pixels per stored row
constexpr long long kPackImageElems = (long long)kCrop * kStemPitch * kPackC;

struct Intrinsics {
  float fx, k01, cx, fy, cy;
  float kinv[9];   // row-major K^-1 (computed in double on the host)
  float proj[16];  // row-major GL projection for the current image size
  int img_h, img_w;
};

// Per-hypothesis constants derived from the pose by pose_setup().
struct PoseAux {
  float mvp[16];  // row-major Proj * GLcam * pose
  float a00, a11, a30, a31;
  float inv_sx, inv_sy, left, top;  // inverse crop transform: src = (x*inv_sx + left, y*inv_sy + top)
  float rot[9];                     // row-major R
  float t[3];
  float tf[9];    // row-major crop transform (debug / tests)
  float bbox[4];  // l, t, r, b (debug / tests)
};

struct DeviceMesh {
  const float* vertices;   // [V,3] centred
  const float* normals;    // [V,3]
  const float* texcoords;  // [V,2] already (u, 1-v)
  const int4* faces;       // [T] (i0, i1, i2, unused): one 16-byte load per triangle
  const float4* texture;   // [th,tw] texels (r, g, b, unused), f32 = u8 * (1/255): one 16-byte load per bilinear tap
  int num_vertices, num_faces, tex_h, tex_w;
  float diameter;
};

// status word bits written by the sampler
constexpr int kStatusEmptyMask = 1;
constexpr int kStatusNoDepth = 2;
constexpr int kStatusClipPath = 4;  // informational: some triangle took cudaraster's polygon clipper

void launch_depth_to_xyz(const float* depth, float* xyz, int h, int w, const Intrinsics& K,
                         cudaStream_t s);
// Erode + bilateral filter of the depth image in one kernel.  eroded may be null (the intermediate image is only
// stored on request); sampler_meta (may be null): the 8-int header of the guess-translation scratch, reset here.
void launch_depth_filters(const float* depth, float* eroded, float* filtered, int h, int w, int* sampler_meta,
                          cudaStream_t s);
// Mask bounding box + median depth + K^-1 back-projection; writes center[3] and status, and (poses != null) the
// n_poses hypotheses = rotation grid at that translation (col-major 4x4).  scratch: 8 + h*w 32-bit words whose
// header launch_depth_filters has reset.
void launch_guess_translation(const float* filtered, const uint8_t* mask, int h, int w,
                              const Intrinsics& K, float* center, int* status, unsigned* scratch, const float* grid,
                              float* poses, int n_poses, cudaStream_t s);
void launch_pose_setup(const float* poses, int n, const Intrinsics& K, float diameter,
                       float crop_ratio, PoseAux* aux, cudaStream_t s);
// Vertex stage + rasterise + shade + flip + normalise + pack hypothesis i into packed image
// `img0 + i`.  vrec: scratch of n * num_vertices records of 2 float4.  zbuf_scratch (may be null):
// zbuf_poses * 160 * 160 keys; with n <= zbuf_poses and few enough hypotheses to leave SMs idle the
// triangles are spread over the whole machine on global z-buffers (identical result).
constexpr int kSplitRenderPoses = 80;
// rast_id_debug (may be null: the product path passes null unless a test asked for the ids) receives the
// visibility id map [n,160,160] (triangle + 1, 0 = empty; GL row order).  num_sms: SM count of the session's
// device; split_mode: -1 automatic, 0 band kernel, 1 split kernels.  Returns the number of kernels launched.
int launch_render(const DeviceMesh& mesh, const PoseAux* aux, int n, float4* vrec,
                   __half* packed, int img0, int32_t* rast_id_debug, int* status,
                   unsigned long long* zbuf_scratch, int zbuf_poses, int num_sms, int split_mode, cudaStream_t s);
// Crop/warp the observed rgb + xyz, normalise, pack into packed image `img0 + i`.
void launch_warp(const uint8_t* rgb, const float* xyz, int h, int w, const PoseAux* aux, int n,
                 float diameter, __half* packed, int img0, cudaStream_t s);
// The reference's engine blobs (f32 NHWC [n,160,160,6], foundationpose.cpp:349-354) -> packed fp16
// network input images img0 .. img0+n-1.
void launch_repack_f32(const float* blob, int n, __half* packed, int img0, cudaStream_t s);
// CalcMeshDiameter (mesh_loader/assimp_mesh_loader.cpp:47-60): max pairwise vertex distance, all V^2 / 2 pairs, on the
// device.  out_sq receives the maximum SQUARED distance as float bits (an unsigned max is exact for non-negative floats);
// it must be zeroed by the caller; diameter = sqrtf(out_sq) (sqrt is monotone, so this equals the max of the lengths).
void launch_mesh_diameter(const float* vertices, int n, unsigned* out_sq, cudaStream_t s);
const char* preprocess_init();  // kernel attributes; nullptr on success

}  // namespace fpb
