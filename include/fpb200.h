/* fpb200 — C ABI of the B200-native FoundationPose Register/Track hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  Each entry point
 * names the reference interface it replaces (paths under zz990099/foundationpose_cpp):
 *   fp/  = detection_6d_foundationpose/     edt/ = easy_deploy_tool/
 *
 * Conventions
 *   - 4x4 poses are 16 floats, COLUMN-major (Eigen::Matrix4f storage, as the reference copies them
 *     to the device: fp/src/foundationpose_render.cpp:836); a pose maps the CENTRED mesh into the
 *     OpenCV camera frame (fp/src/foundationpose.cpp:225,262).
 *   - rgb is u8 [h,w,3] RGB, depth is f32 [h,w] metres, mask is u8 [h,w] (>0 = object), all HOST
 *     pointers borrowed for the call (fp/include/detection_6d_foundationpose/foundationpose.hpp:16-77).
 *   - every function returns FPB200_OK (0) or a negative error code; fpb200_last_error() returns
 *     the calling thread's last message.  (The reference returns bool + glog; 0 <-> true.)
 *   - a session is NOT thread-safe per target mesh, like the reference's renderer scratch
 *     (fp/src/foundationpose_render.hpp:129-152).
 */
#ifndef FPB200_H_
#define FPB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  FPB200_OK = 0,
  FPB200_ERR_INVALID = -1, /* bad argument (reference: CHECK_STATE false / std::invalid_argument) */
  FPB200_ERR_CUDA = -2,    /* CUDA failure (reference: CHECK_CUDA -> false) */
  FPB200_ERR_IO = -3,      /* weight file missing/corrupt (reference: engine load failure) */
  FPB200_ERR_STATE = -4,   /* call order violated, e.g. stage call before upload */
  FPB200_ERR_NO_DEPTH = -5 /* empty mask / no valid depth (fp/src/foundationpose_sampling.cpp:250-298) */
};

typedef struct fpb200_session fpb200_session;

/* What a BaseMeshLoader yields (fp/include/detection_6d_foundationpose/mesh_loader.hpp:15-62).
 * vertices are the loader's raw vertices; the library subtracts `center` like
 * fp/src/foundationpose_render.cpp:396-398 and flips texcoords to (u,1-v) like :405-406. */
typedef struct {
  const char* name;
  const float* vertices;  /* [num_vertices,3] */
  const float* normals;   /* [num_vertices,3] */
  const float* texcoords; /* [num_vertices,2] */
  const int32_t* faces;   /* [num_faces,3] */
  const uint8_t* texture; /* [tex_h,tex_w,3] RGB */
  int num_vertices, num_faces, tex_h, tex_w;
  float diameter;
  float center[3];
} fpb200_mesh;

typedef struct {
  int device;        /* CUDA device ordinal */
  int max_hyps;      /* capacity in hypotheses (reference: 252, fp/src/foundationpose.cpp:86) */
  int shard_rank;    /* multi-GPU: this process handles hypotheses of rank shard_rank ... */
  int shard_world;   /* ... out of shard_world equal contiguous shards (1 = everything) */
  int use_cuda_graph; /* capture the per-iteration launch sequence into CUDA graphs */
} fpb200_options;

void fpb200_default_options(fpb200_options* o);

/* CalcMeshDiameter of the reference's loader (fp/src/mesh_loader/assimp_mesh_loader.cpp:47-60: the maximum distance
 * over ALL vertex pairs, O(V^2) on one CPU core there) evaluated on the device.  vertices: HOST [n,3] f32.
 * fpb200_create calls it for every mesh whose `diameter` field is <= 0. */
int fpb200_mesh_diameter(const float* vertices, int num_vertices, int device, float* diameter);

/* Replaces CreateFoundationPoseModel (fp/include/.../foundationpose.hpp:99-105) together with the
 * two CreateTrtInferCore calls that feed it (edt/inference_core/trt_core/include/trt_core/trt_core.h:26-30):
 * the weight files take the place of the serialized TensorRT engines. */
int fpb200_create(const fpb200_mesh* meshes, int num_meshes, const float K[9], int max_h, int max_w,
                  const char* refiner_weights, const char* scorer_weights,
                  const fpb200_options* opts, fpb200_session** out);
void fpb200_destroy(fpb200_session* s);

/* Base6DofDetectionModel::Register (fp/src/foundationpose.cpp:181-228). */
int fpb200_register(fpb200_session* s, const uint8_t* rgb, const float* depth, const uint8_t* mask,
                    int h, int w, const char* target_name, int refine_itr, float out_pose[16]);
/* Register on the frame (and mask) ALREADY resident in device memory (fpb200_upload +
 * fpb200_sample_hypotheses with a mask were called before): enqueues sampler -> refine x itr ->
 * score -> argmax on the session stream (CUDA-graph replay after the second call) and returns without
 * synchronising; fpb200_wait delivers the pose.  Same result as fpb200_register on the same frame. */
int fpb200_register_async(fpb200_session* s, const char* target_name, int refine_itr);

/* Base6DofDetectionModel::Track (fp/src/foundationpose.cpp:230-265). */
int fpb200_track(fpb200_session* s, const uint8_t* rgb, const float* depth, int h, int w,
                 const float hyp_pose[16], const char* target_name, int refine_itr,
                 float out_pose[16]);
/* Track split in two for callers that pipeline several sessions (one per object / camera) from one
 * host thread, the role EasyDeploy's async pipeline plays around the reference
 * (deploy_core/async_pipeline.h:78-242): fpb200_track_async enqueues upload + refine + read-back on the
 * session stream and returns; fpb200_track_wait blocks until the pose is there.  rgb / depth must stay
 * valid until the wait returns (pinned host memory makes the upload truly asynchronous). */
int fpb200_track_async(fpb200_session* s, const uint8_t* rgb, const float* depth, int h, int w,
                       const float hyp_pose[16], const char* target_name, int refine_itr);
int fpb200_track_wait(fpb200_session* s, float out_pose[16]);

/* Tracking as a STREAM of frames (the reference would use EasyDeploy's async pipeline,
 * deploy_core/async_pipeline.h:78-242, around fp/src/foundationpose.cpp:230-265): frame k + 1 is uploaded on a copy
 * stream into a second set of frame buffers while frame k is still being refined, the refined pose stays in device memory
 * as the hypothesis of the next frame (no host round trip), results are delivered in order.  At most two frames are in
 * flight; rgb / depth of a pushed frame must stay valid until its pop returns (pinned memory makes the upload truly
 * asynchronous).  The poses equal those of consecutive fpb200_track calls that feed each result back in. */
int fpb200_track_stream_begin(fpb200_session* s, const float hyp_pose[16], const char* target_name, int refine_itr);
int fpb200_track_stream_push(fpb200_session* s, const uint8_t* rgb, const float* depth, int h, int w);
int fpb200_track_stream_pop(fpb200_session* s, float out_pose[16]);

/* ---- stage-level entry points (the reference's private seams), used by tests / multi-GPU ---- */

/* UploadDataToDevice + depth->xyz (fp/src/foundationpose.cpp:267-315, fp/src/foundationpose_utils.cu:3-75). */
int fpb200_upload(fpb200_session* s, const uint8_t* rgb, const float* depth, int h, int w);
/* FoundationPoseSampler::GetHypPoses (fp/src/foundationpose_sampling.cpp:344-394): fills the
 * session's hypothesis set with the 252-pose grid at the guessed translation.  mask == NULL reuses
 * the mask already resident on the device. */
int fpb200_sample_hypotheses(fpb200_session* s, const uint8_t* mask, const char* target_name);
/* Replace the session's hypothesis set (n col-major 4x4).  Track does this with n = 1. */
int fpb200_set_hypotheses(fpb200_session* s, const float* poses, int n, const char* target_name);
int fpb200_get_hypotheses(fpb200_session* s, float* poses, int capacity, int* n);
/* FoundationPoseRenderer::RenderAndTransform (fp/src/foundationpose_render.hpp:29-37) on the
 * session's current hypotheses: fills the packed network input (render + observed crops). */
int fpb200_render_and_transform(fpb200_session* s, float crop_ratio);
/* One refine iteration = RefinePreProcess + refiner SyncInfer + RefinePostProcess
 * (fp/src/foundationpose.cpp:317-406); hypotheses are updated in place on the device. */
int fpb200_refine_step(fpb200_session* s);
/* ScorePreprocess + scorer SyncInfer up to the per-hypothesis 512-d feature
 * (fp/src/foundationpose.cpp:408-430).  With shard_world > 1 only the local shard is computed. */
int fpb200_score_features(fpb200_session* s);
/* Device pointer + row count of the local shard's features [n_local,512] f32 and of the full
 * buffer [n,512] the tail reads; a multi-GPU host all-gathers into the latter. */
int fpb200_feature_buffers(fpb200_session* s, void** local_dev, int* n_local, int* local_offset,
                           void** all_dev, int* n_all);
/* Cross-hypothesis attention + Linear + argmax (scorer tail) and ScorePostProcess
 * (fp/src/foundationpose.cpp:432-446, fp/src/foundationpose_decoder.cu:24-35). */
int fpb200_score_finish(fpb200_session* s, int* best_index, float out_pose[16]);

/* ---- multi-GPU: the N hypotheses of ONE Register sharded over the GPUs of a node, one process per GPU ----
 * The reference has no multi-GPU path; this is the exchange SURVEY.md section 8(e) derives from the scorer's
 * cross-hypothesis attention: every rank needs all [N,512] features (+ the refined poses, for the winner).
 * It runs over NVLink peer memory inside the library: each session (shard_world > 1) owns an exchange region
 * in its HBM; fpb200_exchange_export yields its CUDA IPC handle, the host program hands every rank the handles
 * of all ranks (any transport: torch.distributed, MPI, a file), fpb200_exchange_connect maps them.  After that
 * fpb200_register / fpb200_register_async work on sharded sessions: each rank renders / refines / scores its
 * shard, stores its rows straight into every peer's region and ranks all N redundantly, so every rank returns
 * the same pose.  All ranks must make the same sequence of Register calls (a missing peer is reported as
 * FPB200_ERR_STATE after ~2 s, never a hang).  fpb200_exchange_connect_local does the same for sessions that
 * live in ONE process (tests; several GPUs driven by one host thread).  Ranks that SHARE a device spin-wait for each
 * other inside kernels, so each of their streams (2 per session) needs a hardware queue of its own: with more than 3
 * such sessions set CUDA_DEVICE_MAX_CONNECTIONS >= 2 * world + 2 before CUDA initialises (default 8), otherwise an
 * aliased queue parks a peer's kernels behind a spinning one and the call reports the exchange time-out. */
#define FPB200_IPC_HANDLE_BYTES 64
int fpb200_exchange_export(fpb200_session* s, unsigned char handle[FPB200_IPC_HANDLE_BYTES]);
int fpb200_exchange_connect(fpb200_session* s, const unsigned char* handles /* [shard_world][64], rank order */);
int fpb200_exchange_connect_local(fpb200_session* const* sessions, int world);

/* Copy a named device buffer of the session to the host (tests): "render_input" / "transf_input"
 * (f16 [n,160,160,8] packed rgb+xyz+2 pad), "xyz_map" (f32 [h,w,3]), "scores" (f32 [n]),
 * "trans" / "rot" (f32 [n,3]), "features" (f32 [n,512]), "rast_id" (i32 [n,160,160]; needs
 * fpb200_debug_options(s, 1, ...)), "crop_tf" (f32 [n,9] row-major) / "bbox2d" (f32 [n,4]) of the last
 * render_and_transform, "depth_filtered" (f32 [h,w]), "tokens" (f16 [n,400,512]), "center" (f32 [3]),
 * "status" (i32), "pose_aux" (raw per-hypothesis constants). */
int fpb200_read_buffer(fpb200_session* s, const char* name, void* host_dst, size_t capacity_bytes,
                       size_t* bytes_written);
/* Test switches.  write_rast_ids != 0: every render also writes the visibility id map read back as "rast_id"
 * (off by default: nothing on the product path reads it).  render_split: -1 automatic, 0 = one CTA per
 * (band, hypothesis) on shared-memory z-buffers, 1 = triangles spread over all SMs on global z-buffers
 * (what few-hypothesis launches use); both give identical ids.  Drops captured CUDA graphs. */
int fpb200_debug_options(fpb200_session* s, int write_rast_ids, int render_split);
/* Average device time (ms, CUDA events on the session stream) of `repeat` runs of one stage on the session's
 * current hypotheses: RENDER = pose set-up + vertex stage + rasterise + shade + pack (refine crop ratio; what
 * the reference does in RenderProcess, fp/src/foundationpose_render.cpp:685-729), WARP = observed crop
 * (TransfProcess, :731-812), EXCHANGE = the peer-memory exchange alone (all ranks must call it together). */
enum { FPB200_STAGE_RENDER = 0, FPB200_STAGE_WARP = 1, FPB200_STAGE_EXCHANGE = 2 };
int fpb200_debug_time_stage(fpb200_session* s, int stage, int repeat, float* ms_out);
/* Device-side duration of the stages of the last Register/Track (CUDA events), in ms:
 * [0] upload+xyz [1] sampler [2] unused [3] refine iterations (render, warp, net, pose update)
 * [4] score preprocess + scorer net [5] scorer tail + argmax [6] unused; returns the number filled. */
int fpb200_last_timings(fpb200_session* s, float* ms, int capacity);
/* Algorithmic FLOPs of one refiner / scorer forward per hypothesis, from the loaded layer list. */
int fpb200_flops_per_hypothesis(fpb200_session* s, double* refiner, double* scorer);
/* Number of kernel launches issued by the last Register/Track call / since session creation. */
int fpb200_last_launch_count(fpb200_session* s);
long long fpb200_launch_count_total(fpb200_session* s);

/* ---- asynchronous use (benchmarks, multi-GPU pipelines) ---- */
/* The CUDA stream (cudaStream_t) every kernel of the session is launched on. */
void* fpb200_stream(fpb200_session* s);
/* on != 0: stage-level calls enqueue their work and return without synchronising; errors that are
 * only known on the device (empty mask, no depth) are then reported by fpb200_wait. */
int fpb200_set_async(fpb200_session* s, int on);
/* fpb200_score_finish without the final synchronisation; pair with fpb200_wait. */
int fpb200_score_finish_async(fpb200_session* s);
int fpb200_wait(fpb200_session* s, int* best_index, float out_pose[16]);
/* Bracket every tcgen05 GEMM launch with CUDA events on the session stream; fpb200_gemm_profile
 * returns (and clears) the summed duration, algorithmic FLOPs and launch count since the last call. */
int fpb200_set_gemm_profile(fpb200_session* s, int on);
int fpb200_gemm_profile(fpb200_session* s, double* ms, double* flops, long long* launches);

/* ---- inference-core boundary: the refine / score networks alone, behind the plugin interface the
 * reference uses for its TensorRT engines.  Replaces CreateTrtInferCore + BaseInferCore::SyncInfer
 * (edt/inference_core/trt_core/include/trt_core/trt_core.h:26-30,
 *  edt/deploy_core/include/deploy_core/base_infer_core.h:179-237) with DEVICE-located blobs in the
 * reference's own layout: render_input / transf_input f32 NHWC [n,160,160,6]
 * (fp/src/foundationpose.cpp:78-83,349-354) -> trans [n,3] + rot [n,3], or scores [n]. ---- */
typedef struct fpb200_infer_core fpb200_infer_core;
int fpb200_infer_create(const char* weights, int kind /*0 refiner, 1 scorer*/, int max_batch, int device,
                        fpb200_infer_core** out);
void fpb200_infer_destroy(fpb200_infer_core* core);
int fpb200_infer_refiner(fpb200_infer_core* core, const float* render_input, const float* transf_input, int n,
                         float* trans, float* rot);
int fpb200_infer_scorer(fpb200_infer_core* core, const float* render_input, const float* transf_input, int n,
                        float* scores);

/* Device buffers and copies for a caller that holds the blobs but is compiled without the CUDA runtime (the
 * `B200InferCore : inference_core::BaseInferCore` plugin in foundationpose_cpp_b200/cpp/b200_infer_core.hpp; the
 * reference's TrtTensor does the same with cudaMalloc / cudaMemcpy, edt/inference_core/trt_core/src/trt_blob_buffer.hpp:50-191). */
enum { FPB200_COPY_H2D = 0, FPB200_COPY_D2H = 1, FPB200_COPY_D2D = 2 };
int fpb200_device_alloc(size_t bytes, int device, void** out);
void fpb200_device_free(void* ptr);
int fpb200_memcpy(void* dst, const void* src, size_t bytes, int kind);

/* ---- kernel-level debug entry points (DEVICE pointers) ---- */
enum {
  FPB200_TAPS_1X1 = 0,
  FPB200_TAPS_3X3_S1 = 1,
  FPB200_TAPS_3X3_S2_PHASE = 2,
  FPB200_TAPS_STEM7_ROWS = 3
};
typedef struct {
  const void* in; /* f16 */
  int n_img, planes, in_h, in_w, cin;
  long long stride_w, stride_h, stride_plane, stride_img; /* bytes; 0 = dense */
  int out_h, out_w, bw, bh, bn, tap_kind;
  const void* weight; /* f16 [cout][taps*cin] */
  int cout;
  const float* bias;
  const void* residual; /* f16 or NULL */
  int res_ld;
  const float* post_add;
  void* out; /* f16 */
  int relu, out_phase_split, out_ld, img_mod, half_coff, max_ctas;
  int cluster; /* 0 = automatic, 1 = single CTAs, 2 or 3 = CTA pairs issuing tcgen05.mma.cta_group::2 (M = 256) */
  int block_n; /* 0 = automatic, 64 / 128 / 256 = forced tile width (parity tests run every kernel variant) */
} fpb200_conv_desc;
/* Runs the tcgen05 implicit-GEMM kernel once (+ `repeat` timed launches; avg ms -> *ms_out). */
int fpb200_debug_conv_gemm(const fpb200_conv_desc* desc, int repeat, float* ms_out);
/* Self-attention core of the refine / score transformers (inside the TensorRT engines in the
 * reference, trt_core.cpp:367): qkv [n_hyp*400, 1536] fp16 (Q | K | V, 4 heads x 128) ->
 * out [n_hyp*400, 512] fp16.  Runs once + `repeat` timed launches on the default stream. */
int fpb200_debug_attention(const void* qkv, void* out, int n_hyp, int repeat, float* ms_out);
/* With FPB200_ATTN_PHASES=1 in the environment: SM-clock sums of the last attention launch spent by one
 * softmax thread of CTA 0 in {wait S, row max, exp + P store, wait O, epilogue TMEM -> staging,
 * store issue}, then its item count; 7 values for the full 128-row tiles, 7 for the 16-row tail tiles,
 * 2 unused, then 6 sums of the MMA-issuing thread {wait Q, wait K, wait O drained, wait V, wait P steps, issue}. */
int fpb200_debug_attention_phases(unsigned long long out[22]);

/* Fused tail of one transformer encoder layer of the refiner (inside the TensorRT engine in the reference):
 * att, x0 [n_hyp*400, 512] fp16 -> mean_out [n_hyp, 512] f32 = token mean of
 * LayerNorm2(x1 + FFN(x1)), x1 = LayerNorm1(x0 + att Wout^T + b_out).  Weights [512][512] fp16 (out x in),
 * biases / LayerNorm parameters f32 [512]; partial: n_hyp*25*512 floats of scratch; dbg_*: optional
 * [n_hyp*400, 512] fp16 dumps of LayerNorm1 output, FFN activation, LayerNorm2 output.
 * mode: 0 automatic, 1 single CTAs, 2 CTA pairs. */
typedef struct {
  const void *att, *x0, *w_out, *w1, *w2;
  const float *b_out, *b1, *b2, *ln1_g, *ln1_b, *ln2_g, *ln2_b;
  float *partial, *mean_out;
  int n_hyp, mode;
  void *dbg_x1, *dbg_h, *dbg_y;
  void* dbg_clocks; /* optional u64[24] (device): SM-clock sums of CTA 0's MMA-issuing thread and one epilogue thread per phase */
} fpb200_encoder_desc;
int fpb200_debug_encoder_block(const fpb200_encoder_desc* desc, int repeat, float* ms_out);

/* Stem convolution alone (7 x 7, stride 2, 6 -> 64; csrc/stem_conv.cu): packed [n_img][160][164][8] fp16 (+ one pad row)
 * -> phase-split [n_img][4][40][40][64] fp16 with bias + ReLU.  weight: file layout of `enc.stem.w` ([64][448] fp16). */
int fpb200_debug_stem_conv(const void* packed, int n_img, const void* weight, const float* bias, void* out, int repeat,
                           float* ms_out);

/* fp32 operators of the scorer tail (after the token mean; inside the scorer engine in the reference):
 * out[rows, outs] = in[rows, k] w[outs, k]^T + b (b may be NULL), and the cross-hypothesis attention
 * qkv [n, 1536] (Q | K | V, 4 heads x 128) -> out [n, 512].  Once + `repeat` timed launches. */
int fpb200_debug_small_linear(const float* in, const float* w, const float* b, float* out, int rows, int k, int outs,
                              int repeat, float* ms_out);
int fpb200_debug_cross_attention(const float* qkv, float* out, int n, int repeat, float* ms_out);

const char* fpb200_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* FPB200_H_ */
