// Session object + C ABI (include/fpb200.h): the orchestration the reference does in
// detection_6d_foundationpose/src/foundationpose.cpp (class FoundationPose), re-designed so that one
// Register is a fixed sequence of launches on one stream with a single device->host read at the end
// (the reference synchronises >= 12 times per call and copies every pose separately).
#include <array>
#include <cmath>
#include <functional>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/fpb200.h"
#include "common.h"
#include "net_ops.h"
#include "network.h"
#include "preprocess.h"

using namespace fpb;

namespace {

constexpr int kGridPoses = 252;  // 42 icosphere views x 6 in-plane rotations
constexpr float kRefineCropRatio = 1.2f, kScoreCropRatio = 1.1f;  // foundationpose.cpp:84-85
constexpr int kExchLd = 512 + 16;  // per-hypothesis exchange row: scorer feature + refined pose
constexpr int kNumStages = 7;

// ---- rotation grid: foundationpose_sampling.cpp:56-121,178-237 (host, once per session)
struct V3 { float x, y, z; };
V3 normalized(V3 v) {
  const float n = std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
  return {v.x / n, v.y / n, v.z / n};
}
V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

std::vector<V3> icosphere(unsigned n_views) {
  std::vector<V3> verts;
  std::vector<std::array<int, 3>> faces;
  const float t = (float)((1.0 + std::sqrt(5.0)) / 2.0);
  const V3 raw[12] = {{-1, t, 0}, {1, t, 0}, {-1, -t, 0}, {1, -t, 0}, {0, -1, t}, {0, 1, t},
                      {0, -1, -t}, {0, 1, -t}, {t, 0, -1}, {t, 0, 1}, {-t, 0, -1}, {-t, 0, 1}};
  for (auto& v : raw) verts.push_back(normalized(v));
  const int f0[20][3] = {{0, 11, 5}, {0, 5, 1}, {0, 1, 7}, {0, 7, 10}, {0, 10, 11}, {1, 5, 9}, {5, 11, 4},
                         {11, 10, 2}, {10, 7, 6}, {7, 1, 8}, {3, 9, 4}, {3, 4, 2}, {3, 2, 6}, {3, 6, 8},
                         {3, 8, 9}, {4, 9, 5}, {2, 4, 11}, {6, 2, 10}, {8, 6, 7}, {9, 8, 1}};
  for (auto& f : f0) faces.push_back({f[0], f[1], f[2]});
  std::map<long long, int> cache;
  auto mid = [&](int i, int j) {
    const long long key = ((long long)std::min(i, j) << 32) + std::max(i, j);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    const V3 a = verts[i], b = verts[j];
    verts.push_back(normalized({(a.x + b.x) / 2.f, (a.y + b.y) / 2.f, (a.z + b.z) / 2.f}));
    return cache[key] = (int)verts.size() - 1;
  };
  while (verts.size() < n_views) {
    std::vector<std::array<int, 3>> nf;
    for (auto& f : faces) {
      const int ab = mid(f[0], f[1]), bc = mid(f[1], f[2]), ca = mid(f[2], f[0]);
      nf.push_back({f[0], ab, ca});
      nf.push_back({f[1], bc, ab});
      nf.push_back({f[2], ca, bc});
      nf.push_back({ab, bc, ca});
    }
    faces.swap(nf);
  }
  return verts;
}

// 252 column-major 4x4 ob_in_cam matrices, icosphere-vertex major, in-plane angle minor.
std::vector<float> make_rotation_grid() {
  std::vector<float> out;
  for (const V3& v : icosphere(40)) {
    const V3 z = normalized({-v.x, -v.y, -v.z});
    V3 x = cross({0, 0, 1}, z);
    if (x.x == 0 && x.y == 0 && x.z == 0) x = {1, 0, 0};
    x = normalized(x);
    const V3 y = normalized(cross(z, x));
    for (int step = 0; step < 6; ++step) {
      const float a = (float)(step * 60.0 * M_PI / 180.0);
      const float c = std::cos(a), s = std::sin(a);
      // cam_in_ob = [x y z | v] * Rz(a)
      const double r[3][3] = {{(double)(x.x * c + y.x * s), (double)(-x.x * s + y.x * c), z.x},
                              {(double)(x.y * c + y.y * s), (double)(-x.y * s + y.y * c), z.y},
                              {(double)(x.z * c + y.z * s), (double)(-x.z * s + y.z * c), z.z}};
      const double tv[3] = {v.x, v.y, v.z};
      // general inverse of the rigid-like matrix in double (the reference calls Matrix4f::inverse())
      const double det = r[0][0] * (r[1][1] * r[2][2] - r[1][2] * r[2][1]) -
                         r[0][1] * (r[1][0] * r[2][2] - r[1][2] * r[2][0]) +
                         r[0][2] * (r[1][0] * r[2][1] - r[1][1] * r[2][0]);
      double inv[3][3];
      inv[0][0] = (r[1][1] * r[2][2] - r[1][2] * r[2][1]) / det;
      inv[0][1] = (r[0][2] * r[2][1] - r[0][1] * r[2][2]) / det;
      inv[0][2] = (r[0][1] * r[1][2] - r[0][2] * r[1][1]) / det;
      inv[1][0] = (r[1][2] * r[2][0] - r[1][0] * r[2][2]) / det;
      inv[1][1] = (r[0][0] * r[2][2] - r[0][2] * r[2][0]) / det;
      inv[1][2] = (r[0][2] * r[1][0] - r[0][0] * r[1][2]) / det;
      inv[2][0] = (r[1][0] * r[2][1] - r[1][1] * r[2][0]) / det;
      inv[2][1] = (r[0][1] * r[2][0] - r[0][0] * r[2][1]) / det;
      inv[2][2] = (r[0][0] * r[1][1] - r[0][1] * r[1][0]) / det;
      float m[16];
      for (int cc = 0; cc < 3; ++cc) {
        for (int rr = 0; rr < 3; ++rr) m[4 * cc + rr] = (float)inv[rr][cc];
        m[4 * cc + 3] = 0.f;
      }
      for (int rr = 0; rr < 3; ++rr)
        m[12 + rr] = (float)(-(inv[rr][0] * tv[0] + inv[rr][1] * tv[1] + inv[rr][2] * tv[2]));
      m[15] = 1.f;
      out.insert(out.end(), m, m + 16);
    }
  }
  return out;
}

// ProjectMatrixFromIntrinsics (y_down), foundationpose_render.cpp:151-186; fp32 like the reference.
void projection_matrix(const float K[9], int h, int w, float out[16]) {
  const float nc = 0.1f, fc = 100.0f;
  const float depth = fc - nc;
  const float q = -(fc + nc) / depth;
  const float qn = -2.f * (fc * nc) / depth;
  const float wf = (float)w, hf = (float)h;
  const float m[16] = {2.f * K[0] / wf, -2.f * K[1] / wf, (-2.f * K[2] + wf) / wf, 0.f,
                       0.f, 2.f * K[4] / hf, (2.f * K[5] - hf) / hf, 0.f,
                       0.f, 0.f, q, qn,
                       0.f, 0.f, -1.f, 0.f};
  memcpy(out, m, sizeof(m));
}

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  void alloc(size_t count) {
    n = count;
    FPB_CUDA(cudaMalloc(&p, sizeof(T) * (count ? count : 1)));
  }
  void upload(const T* src, size_t count) {
    alloc(count);
    FPB_CUDA(cudaMemcpy(p, src, sizeof(T) * count, cudaMemcpyHostToDevice));
  }
  ~DevBuf() { if (p) cudaFree(p); }
};

struct MeshEntry {
  std::string name;
  DevBuf<float> vertices, normals, texcoords;
  DevBuf<float4> texture;  // (r, g, b, 0) texels
  DevBuf<int4> faces;      // (i0, i1, i2, 0)
  DeviceMesh dm{};
};

struct HostResult {
  float pose[16];
  float hyp[16];  // staging of the Track hypothesis
  int best_index;
  int status;
};

// ---- multi-GPU exchange over peer memory -------------------------------------------------------------
// The path has exactly one exchange: the scorer's cross-hypothesis attention needs every
// hypothesis's 512-d feature, and the winner's refined pose must be known everywhere.  One process per GPU;
// each rank owns a REGION in its own HBM that its peers map through CUDA IPC (NVLink / NVSwitch peer
// access).  A Register then contains no library collective at all: the rank's rows are stored straight into
// every peer's region by exchange_push_kernel, followed by a release-store of the call's epoch into the peer's
// arrival flag; exchange_collect_kernel acquires all flags and compacts the rows for the tail.  Everything is
// kernels on the session stream, so the whole sharded Register is one CUDA-graph replay.
// Region layout (floats): rows [2][cap_total][kExchLd] | flags u32[kMaxWorld] | epoch u32.
// Rows are double-buffered by epoch parity: a peer can only be one exchange ahead (it needs this rank's flag of
// epoch e + 1 before it finishes e + 1), so its stores of e + 1 never touch the buffer this rank still reads for e.
constexpr int kMaxWorld = 16;
constexpr int kExchTailWords = kMaxWorld + 16;  // flags + epoch (+ padding)
struct ExchPeers {
  float* rows[kMaxWorld];     // region base of every rank (own region included)
  unsigned* flags[kMaxWorld]; // arrival flags of every rank's region
  unsigned* epoch;            // this rank's exchange counter (device memory: graph replays cannot change arguments)
  int world, rank, cap_total;
};
constexpr int kStatusExchangeTimeout = 8;

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void exchange_begin_kernel(unsigned* epoch) { *epoch += 1; }

// Status reset as a KERNEL, not cudaMemsetAsync: copy-engine operations of all streams share FIFO queues, so a memset
// could wait behind another session's result read-back that is itself waiting for a spinning exchange_wait_kernel
// (sessions of one process sharing a device).  Kernels only wait on their own stream.
__global__ void clear_status_kernel(int* status) { *status = 0; }

// grid = world (one CTA per destination rank): local rows [lo, lo + nl) = 512 feature floats + 16 pose floats.
__global__ void __launch_bounds__(1024)
exchange_push_kernel(ExchPeers P, const float* __restrict__ feats, const float* __restrict__ poses, int lo, int nl) {
  const int peer = blockIdx.x;
  const unsigned e = *P.epoch;
  float4* dst = reinterpret_cast<float4*>(P.rows[peer] + ((size_t)(e & 1u) * P.cap_total + lo) * kExchLd);
  constexpr int kRow4 = kExchLd / 4;  // 132 float4 per row: 128 feature + 4 pose
  const float4* f4 = reinterpret_cast<const float4*>(feats);
  const float4* p4 = reinterpret_cast<const float4*>(poses);
  for (int i = threadIdx.x; i < nl * kRow4; i += blockDim.x) {
    const int r = i / kRow4, c = i - r * kRow4;
    dst[i] = c < 128 ? f4[r * 128 + c] : p4[r * 4 + (c - 128)];
  }
  __threadfence_system();  // every thread's stores are visible system-wide before the flag is
  __syncthreads();
  if (threadIdx.x == 0) st_release_sys(P.flags[peer] + P.rank, e);
}

// exchange_wait_kernel (ONE warp: a spinning CTA per SM would take scheduling slots from co-resident work) waits until
// every rank's rows of this epoch have arrived in the local region and records the verdict; exchange_collect_kernel then
// compacts the rows into the dense [n,512] feature and [n,16] pose arrays the tail reads.  A rank that never arrives
// (a peer that skipped the call, or died) raises kStatusExchangeTimeout after ~2 s instead of hanging the GPU.
__global__ void __launch_bounds__(32) exchange_wait_kernel(ExchPeers P, int* status) {
  const unsigned e = *P.epoch;
  bool late = false;
  if ((int)threadIdx.x < P.world) {
    const unsigned* flag = P.flags[P.rank] + threadIdx.x;
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(flag) - e) < 0) {
      if (clock64() - t0 > 4000000000LL) {
        late = true;
        break;
      }
      __nanosleep(200);
    }
  }
  if (__any_sync(0xffffffffu, late) && threadIdx.x == 0) atomicOr(status, kStatusExchangeTimeout);
}

__global__ void __launch_bounds__(256)
exchange_collect_kernel(ExchPeers P, float* __restrict__ feats, float* __restrict__ poses, int n, const int* status) {
  if (*status & kStatusExchangeTimeout) return;
  const unsigned e = *P.epoch;
  const float4* src = reinterpret_cast<const float4*>(P.rows[P.rank] + (size_t)(e & 1u) * P.cap_total * kExchLd);
  constexpr int kRow4 = kExchLd / 4;
  float4* f4 = reinterpret_cast<float4*>(feats);
  float4* p4 = reinterpret_cast<float4*>(poses);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n * kRow4; i += gridDim.x * blockDim.x) {
    const int r = i / kRow4, c = i - r * kRow4;
    const float4 v = __ldcg(src + i);  // peers wrote these through NVLink: read at L2, never a stale L1 line
    if (c < 128) f4[r * 128 + c] = v; else p4[r * 4 + (c - 128)] = v;
  }
}

}  // namespace

struct fpb200_session {
  fpb200_options opt{};
  float K[9];
  int max_h = 0, max_w = 0;
  Intrinsics intr{};
  cudaStream_t stream = nullptr;
  // side branch: the observed-crop warp does not depend on the rasteriser, so it runs beside it (fork / join with
  // events; inside a capture this becomes a parallel branch of the CUDA graph)
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // streamed tracking (fpb200_track_stream_*): frames alternate between the primary buffers and a second set, uploaded
  // on copy_stream while the previous frame is still being refined; results go through a ring of pinned slots
  static constexpr int kStreamRing = 4;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr}, ev_result[kStreamRing] = {};
  DevBuf<uint8_t> rgb_alt;
  DevBuf<float> depth_alt, xyz_alt;
  uint8_t* f_rgb = nullptr;   // frame buffers the kernels of the current call read (primary or alternate set)
  float *f_depth = nullptr, *f_xyz = nullptr;
  float (*stream_poses)[16] = nullptr;  // pinned [kStreamRing][16]
  long long stream_pushed = 0, stream_popped = 0;
  bool stream_open = false;
  int stream_itr = 1;
  MeshEntry* stream_mesh = nullptr;
  cudaEvent_t ev[kNumStages + 1]{};
  bool ev_valid[kNumStages + 1]{};
  std::vector<std::unique_ptr<MeshEntry>> meshes;
  MeshEntry* cur_mesh = nullptr;
  DevBuf<uint8_t> rgb, mask;
  DevBuf<float> depth, xyz, filtered, grid, poses, center, exch_local, exch_all, feats, tail_qkv,
      tail_att, tail_out, scores, trans, rot, best_pose, all_poses;
  DevBuf<int> status, best_index;
  DevBuf<unsigned> sampler_scratch;
  DevBuf<int32_t> rast_id;
  DevBuf<unsigned long long> zbuf;  // whole-crop z-buffers of the few-hypotheses render path
  DevBuf<PoseAux> aux;
  DevBuf<float4> vrec;  // vertex-stage output [local hypotheses][max vertices][2]
  HostResult* host_result = nullptr;  // pinned
  std::unique_ptr<Arena> arena;
  std::unique_ptr<Network> refiner, scorer;
  int img_h = 0, img_w = 0;
  bool uploaded = false;
  int n_hyp = 0;             // total hypotheses of the current call
  int lo = 0, hi = 0;        // local shard [lo, hi)
  int shard_cap = 0;         // rows per rank in the padded exchange buffer
  bool features_ready = false;
  int launches = 0;        // since the last Register/Track/upload
  long long launches_total = 0;
  float crop_ratio_last = 0.f;
  bool in_graph = false;      // capturing or about to replay: no event marks
  struct GraphEntry {
    bool seen = false;
    cudaGraphExec_t exec = nullptr;
    int launches = 0;
  };
  std::map<std::string, GraphEntry> graphs;  // Register / Track launch sequences, keyed by their static parameters
  bool async_stages = false;  // stage calls return without synchronising the stream
  bool mask_resident = false;
  bool profile_gemm = false;
  int num_sms = 148;
  // test switches (fpb200_debug_options): visibility ids are only written when a test asked for them
  // (25.8 MB per render at N = 252 that nothing on the product path reads)
  bool debug_ids = false;
  int render_split = -1;  // -1 automatic, 0 band kernel, 1 split kernels
  bool pending_async = false;  // an async call's pinned staging / result has not been collected yet
  // Multi-GPU exchange over NVLink peer memory (fpb200_exchange_*): every rank owns one region
  // [2 epochs][max_hyps][528] f32 + arrival flags; ranks store their rows straight into every peer's region.
  DevBuf<float> exch_region;
  ExchPeers peers{};
  std::vector<void*> ipc_opened;
  bool exchange_ready = false;

  void drop_graphs() {
    for (auto& g : graphs) if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
    graphs.clear();
  }

  ~fpb200_session() {
    drop_graphs();
    for (void* p : ipc_opened) cudaIpcCloseMemHandle(p);
    if (host_result) cudaFreeHost(host_result);
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    for (auto e : ev_copied) if (e) cudaEventDestroy(e);
    for (auto e : ev_consumed) if (e) cudaEventDestroy(e);
    for (auto e : ev_result) if (e) cudaEventDestroy(e);
    if (stream_poses) cudaFreeHost(stream_poses);
    if (copy_stream) cudaStreamDestroy(copy_stream);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    if (side_stream) cudaStreamDestroy(side_stream);
    if (stream) cudaStreamDestroy(stream);
  }
  MeshEntry* find_mesh(const char* name) {
    if (!name) return nullptr;
    for (auto& m : meshes) if (m->name == name) return m.get();
    return nullptr;
  }
  void count_launches(int n) {
    launches += n;
    launches_total += n;
  }
  void stage_sync() {
    if (!async_stages) FPB_CUDA(cudaStreamSynchronize(stream));
  }
  void mark(int i) {
    if (in_graph) return;  // stage timings are only taken on the eager path
    FPB_CUDA(cudaEventRecord(ev[i], stream));
    ev_valid[i] = true;
  }
  void set_shard(int n) {
    n_hyp = n;
    const int world = opt.shard_world, rank = opt.shard_rank;
    const int base = n / world, rem = n % world;
    lo = rank * base + std::min(rank, rem);
    hi = lo + base + (rank < rem ? 1 : 0);
    shard_cap = (n + world - 1) / world;
    features_ready = false;
  }
};

namespace {

int guard(const std::function<int()>& f) {
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

void do_upload(fpb200_session* s, const uint8_t* rgb, const float* depth, int h, int w) {
  FPB_NVTX("fpb200/upload");
  FPB_REQUIRE(rgb && depth, "rgb/depth must not be null");
  FPB_REQUIRE(h > 0 && w > 0 && h <= s->max_h && w <= s->max_w, "image %dx%d exceeds the session maximum %dx%d", h,
              w, s->max_h, s->max_w);
  if (h != s->img_h || w != s->img_w) {
    s->img_h = h;
    s->img_w = w;
    s->intr.img_h = h;
    s->intr.img_w = w;
    projection_matrix(s->K, h, w, s->intr.proj);
  }
  s->mark(0);
  s->f_rgb = s->rgb.p; s->f_depth = s->depth.p; s->f_xyz = s->xyz.p;  // (a streamed call may have left the alternates)
  s->stream_open = false;  // an ordinary call ends a frame stream
  FPB_CUDA(cudaMemcpyAsync(s->rgb.p, rgb, (size_t)h * w * 3, cudaMemcpyHostToDevice, s->stream));
  FPB_CUDA(cudaMemcpyAsync(s->depth.p, depth, sizeof(float) * (size_t)h * w, cudaMemcpyHostToDevice, s->stream));
  s->uploaded = true;
}

void do_xyz(fpb200_session* s) {
  FPB_NVTX("fpb200/depth_to_xyz");
  launch_depth_to_xyz(s->f_depth, s->f_xyz, s->img_h, s->img_w, s->intr, s->stream);
  s->count_launches(1);
  s->mark(1);
}

void do_mask_copy(fpb200_session* s, const uint8_t* mask) {
  if (mask) {
    FPB_CUDA(cudaMemcpyAsync(s->mask.p, mask, (size_t)s->img_h * s->img_w, cudaMemcpyHostToDevice, s->stream));
    s->mask_resident = true;
  }
  FPB_REQUIRE(s->mask_resident, "mask must not be null (no mask is resident on the device yet)");
}

void do_sample(fpb200_session* s, MeshEntry* mesh) {
  FPB_NVTX("fpb200/sample_hypotheses");
  const int h = s->img_h, w = s->img_w;
  // one filter kernel (erode + bilateral; the eroded image stays on chip), one scan, one select that also writes
  // the hypotheses
  const int n = std::min(kGridPoses, s->opt.max_hyps);
  launch_depth_filters(s->f_depth, nullptr, s->filtered.p, h, w, reinterpret_cast<int*>(s->sampler_scratch.p), s->stream);
  launch_guess_translation(s->filtered.p, s->mask.p, h, w, s->intr, s->center.p, s->status.p, s->sampler_scratch.p,
                           s->grid.p, s->poses.p, n, s->stream);
  s->count_launches(3);
  s->cur_mesh = mesh;
  s->set_shard(n);
  s->mark(2);
}

void do_render_and_transform(fpb200_session* s, float crop_ratio) {
  FPB_NVTX("fpb200/render_and_transform");
  FPB_REQUIRE(s->uploaded && s->cur_mesh && s->n_hyp > 0, "upload an image and set hypotheses first");
  const int nl = s->hi - s->lo;
  const MeshEntry& m = *s->cur_mesh;
  launch_pose_setup(s->poses.p + 16 * s->lo, nl, s->intr, m.dm.diameter, crop_ratio, s->aux.p, s->stream);
  // fork: warp (HBM-bound, reads the frame) beside the rasteriser (latency-bound, reads the mesh); both only need aux
  FPB_CUDA(cudaEventRecord(s->ev_fork, s->stream));
  FPB_CUDA(cudaStreamWaitEvent(s->side_stream, s->ev_fork, 0));
  launch_warp(s->f_rgb, s->f_xyz, s->img_h, s->img_w, s->aux.p, nl, m.dm.diameter, s->arena->packed, nl, s->side_stream);
  FPB_CUDA(cudaEventRecord(s->ev_join, s->side_stream));
  const int k_render = launch_render(m.dm, s->aux.p, nl, s->vrec.p, s->arena->packed, 0,
                                     s->debug_ids ? s->rast_id.p : nullptr, s->status.p, s->zbuf.p, kSplitRenderPoses,
                                     s->num_sms, s->render_split, s->stream);
  FPB_CUDA(cudaStreamWaitEvent(s->stream, s->ev_join, 0));  // join before the networks read both halves of `packed`
  s->count_launches(2 + k_render);
  s->crop_ratio_last = crop_ratio;
}

void do_refine_step(fpb200_session* s) {
  FPB_NVTX("fpb200/refine_step");
  do_render_and_transform(s, kRefineCropRatio);
  const int nl = s->hi - s->lo;
  s->refiner->forward(nl, s->stream);
  const WeightFile& W = s->refiner->weights();
  launch_refine_update(s->arena->mean_t, s->arena->mean_r, W.f32("trans.head.w", 3, 512), W.f32("trans.head.b", 3),
                       W.f32("rot.head.w", 3, 512), W.f32("rot.head.b", 3), s->cur_mesh->dm.diameter,
                       s->poses.p + 16 * s->lo, s->trans.p, s->rot.p, nl, s->stream);
  s->count_launches(s->refiner->launches_per_forward(nl) + 1);
}

__global__ void pack_exchange_kernel(const float* __restrict__ feats, const float* __restrict__ poses,
                                     float* __restrict__ exch, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * kExchLd) return;
  const int r = i / kExchLd, c = i - r * kExchLd;
  exch[i] = c < 512 ? feats[r * 512 + c] : poses[r * 16 + (c - 512)];
}

// exchange rows of rank r sit at [r * cap, r * cap + count_r); compact to dense [n,512] + [n,16]
__global__ void unpack_exchange_kernel(const float* __restrict__ exch, float* __restrict__ feats,
                                       float* __restrict__ poses, int n, int world, int cap) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * kExchLd) return;
  const int row = i / kExchLd, c = i - row * kExchLd;
  const int base = n / world, rem = n % world;
  // rank owning global row `row`
  int rank, start;
  if (row < rem * (base + 1)) {
    rank = row / (base + 1);
    start = rank * (base + 1);
  } else {
    rank = rem + (row - rem * (base + 1)) / (base > 0 ? base : 1);
    start = rem * (base + 1) + (rank - rem) * base;
  }
  const float v = exch[((long long)rank * cap + (row - start)) * kExchLd + c];
  if (c < 512) feats[row * 512 + c] = v; else poses[row * 16 + (c - 512)] = v;
}

void do_score_features(fpb200_session* s) {
  FPB_NVTX("fpb200/score_features");
  do_render_and_transform(s, kScoreCropRatio);
  const int nl = s->hi - s->lo;
  s->scorer->forward(nl, s->stream);
  const WeightFile& W = s->scorer->weights();
  // out_proj of the self-attention applied to the token mean (linear, so it commutes with the mean)
  launch_small_linear(s->arena->mean_t, W.f32("att.out.w", 512, 512), W.f32("att.out.b", 512), s->feats.p, nl, 512,
                      512, s->stream);
  // The exchange rows only exist to cross GPUs; a single-GPU session ranks straight from feats / poses.
  int extra = 0;
  if (s->exchange_ready) {
    // peer-memory exchange: this rank's rows go straight into every rank's region (NVLink stores)
    exchange_begin_kernel<<<1, 1, 0, s->stream>>>(s->peers.epoch);
    exchange_push_kernel<<<s->peers.world, 1024, 0, s->stream>>>(s->peers, s->feats.p, s->poses.p + 16 * s->lo, s->lo, nl);
    FPB_CUDA(cudaGetLastError());
    extra = 2;
  } else if (s->opt.shard_world > 1) {
    // host-driven exchange: the caller all-gathers the packed rows (fpb200_feature_buffers)
    pack_exchange_kernel<<<(nl * kExchLd + 255) / 256, 256, 0, s->stream>>>(s->feats.p, s->poses.p + 16 * s->lo,
                                                                           s->exch_local.p, nl);
    FPB_CUDA(cudaGetLastError());
    extra = 1;
  }
  s->count_launches(s->scorer->launches_per_forward(nl) + 1 + extra);
  s->features_ready = true;
}

void do_score_finish(fpb200_session* s) {
  FPB_NVTX("fpb200/score_finish");
  FPB_REQUIRE(s->features_ready, "score_features must run before score_finish");
  const int n = s->n_hyp;
  const WeightFile& W = s->scorer->weights();
  const float* all_poses = s->poses.p;
  if (s->exchange_ready) {
    exchange_wait_kernel<<<1, 32, 0, s->stream>>>(s->peers, s->status.p);
    exchange_collect_kernel<<<16, 256, 0, s->stream>>>(s->peers, s->feats.p, s->all_poses.p, n, s->status.p);
    FPB_CUDA(cudaGetLastError());
    all_poses = s->all_poses.p;
  } else if (s->opt.shard_world > 1) {
    unpack_exchange_kernel<<<(n * kExchLd + 255) / 256, 256, 0, s->stream>>>(s->exch_all.p, s->feats.p, s->all_poses.p,
                                                                            n, s->opt.shard_world, s->shard_cap);
    FPB_CUDA(cudaGetLastError());
    all_poses = s->all_poses.p;
  }
  launch_small_linear(s->feats.p, W.f32("cross.qkv.w", 1536, 512), W.f32("cross.qkv.b", 1536), s->tail_qkv.p, n, 512,
                      1536, s->stream);
  launch_cross_attention(s->tail_qkv.p, s->tail_att.p, n, s->stream);
  launch_small_linear(s->tail_att.p, W.f32("cross.out.w", 512, 512), W.f32("cross.out.b", 512), s->tail_out.p, n, 512,
                      512, s->stream);
  launch_small_linear(s->tail_out.p, W.f32("lin.w", 1, 512), W.f32("lin.b", 1), s->scores.p, n, 512, 1, s->stream);
  launch_argmax_pose(s->scores.p, all_poses, n, s->best_index.p, s->best_pose.p, s->stream);
  s->count_launches(s->exchange_ready ? 7 : s->opt.shard_world > 1 ? 6 : 5);
}

// Single device->host read of the call: winning pose, its index and the status word.
void enqueue_result_read(fpb200_session* s) {
  FPB_CUDA(cudaMemcpyAsync(s->host_result->pose, s->best_pose.p, sizeof(float) * 16, cudaMemcpyDeviceToHost, s->stream));
  FPB_CUDA(cudaMemcpyAsync(&s->host_result->best_index, s->best_index.p, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
  FPB_CUDA(cudaMemcpyAsync(&s->host_result->status, s->status.p, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
}

int wait_result(fpb200_session* s, float out_pose[16], int* best_index) {
  FPB_NVTX("fpb200/wait_result");
  FPB_CUDA(cudaStreamSynchronize(s->stream));
  s->pending_async = false;
  const int st = s->host_result->status;
  if (st & kStatusEmptyMask) return fail(FPB200_ERR_NO_DEPTH, "[FoundationposeSampling] Mask is all zero.");
  if (st & kStatusNoDepth) return fail(FPB200_ERR_NO_DEPTH, "[FoundationposeSampling] No valid value in mask.");
  if (st & kStatusExchangeTimeout)
    return fail(FPB200_ERR_STATE, "multi-GPU exchange timed out: a peer rank did not take part in this call");
  if (out_pose) memcpy(out_pose, s->host_result->pose, sizeof(float) * 16);
  if (best_index) *best_index = s->host_result->best_index;
  return FPB200_OK;
}

int finish_and_read(fpb200_session* s, float out_pose[16], int* best_index) {
  enqueue_result_read(s);
  return wait_result(s, out_pose, best_index);
}

// Runs `body` (a fixed launch sequence on s->stream).  The first call with a given key runs it
// eagerly (this also builds every lazily created plan), the second call captures it into a CUDA
// graph, later calls replay the graph: one cudaGraphLaunch instead of ~70 kernel launches.
template <typename F>
void run_maybe_graphed(fpb200_session* s, const std::string& key, F&& body) {
  if (!s->opt.use_cuda_graph || s->profile_gemm) {
    body();
    return;
  }
  fpb200_session::GraphEntry& g = s->graphs[key];
  if (g.exec) {
    FPB_CUDA(cudaGraphLaunch(g.exec, s->stream));
    s->count_launches(g.launches);
    return;
  }
  if (!g.seen) {
    g.seen = true;
    body();
    return;
  }
  const int before = s->launches;
  s->in_graph = true;
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal);
  if (e == cudaSuccess) {
    try {
      body();
    } catch (...) {
      cudaStreamEndCapture(s->stream, &graph);
      if (graph) cudaGraphDestroy(graph);
      s->in_graph = false;
      throw;
    }
    e = cudaStreamEndCapture(s->stream, &graph);
  }
  s->in_graph = false;
  if (e != cudaSuccess || !graph) throw CudaError(std::string("CUDA graph capture failed: ") + cudaGetErrorString(e));
  e = cudaGraphInstantiate(&g.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) {
    g.exec = nullptr;
    throw CudaError(std::string("cudaGraphInstantiate failed: ") + cudaGetErrorString(e));
  }
  g.launches = s->launches - before;
  FPB_CUDA(cudaGraphLaunch(g.exec, s->stream));
}

void reset_call(fpb200_session* s) {
  s->launches = 0;
  for (auto& v : s->ev_valid) v = false;
  clear_status_kernel<<<1, 1, 0, s->stream>>>(s->status.p);
}

}  // namespace

extern "C" {

void fpb200_default_options(fpb200_options* o) {
  if (!o) return;
  o->device = 0;
  o->max_hyps = kGridPoses;
  o->shard_rank = 0;
  o->shard_world = 1;
  o->use_cuda_graph = 1;
}

int fpb200_create(const fpb200_mesh* meshes, int num_meshes, const float K[9], int max_h, int max_w,
                  const char* refiner_weights, const char* scorer_weights, const fpb200_options* opts,
                  fpb200_session** out) {
  return guard([&]() -> int {
    FPB_REQUIRE(out && K && meshes && num_meshes > 0, "fpb200_create: null argument");
    FPB_REQUIRE(refiner_weights && scorer_weights, "fpb200_create: weight paths are required");
    FPB_REQUIRE(max_h > 0 && max_w > 0, "fpb200_create: bad maximum image size");
    std::unique_ptr<fpb200_session> s(new fpb200_session);
    if (opts) s->opt = *opts; else fpb200_default_options(&s->opt);
    FPB_REQUIRE(s->opt.max_hyps >= 1, "max_hyps must be >= 1");
    FPB_REQUIRE(s->opt.shard_world >= 1 && s->opt.shard_rank >= 0 && s->opt.shard_rank < s->opt.shard_world,
                "bad shard rank/world");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
      throw CudaError("no CUDA device available: this library has no CPU fallback");
    FPB_CUDA(cudaSetDevice(s->opt.device));
    cudaDeviceProp prop;
    FPB_CUDA(cudaGetDeviceProperties(&prop, s->opt.device));
    if (prop.major != 10) throw CudaError("fpb200 targets sm_100a (B200); found a different architecture");
    if (const char* e = preprocess_init()) throw CudaError(e);
    // The wait kernel may spin for as long as a peer needs.  An SM whose shared-memory carve-out has to change for the
    // next kernel must drain first, so a spinning CTA configured for the default (small) carve-out would make its SM
    // unusable for the 200+ KB persistent kernels of a co-resident session (seen with 8 sessions on one device: the
    // persistent kernel's last CTA never found an SM and the exchange timed out).  Asking for the maximum carve-out
    // keeps the SM in the configuration those kernels need.
    FPB_CUDA(cudaFuncSetAttribute(exchange_wait_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  (int)cudaSharedmemCarveoutMaxShared));
    FPB_CUDA(preload_kernels(exchange_begin_kernel, clear_status_kernel, exchange_push_kernel, exchange_wait_kernel,
                             exchange_collect_kernel, pack_exchange_kernel, unpack_exchange_kernel));
    memcpy(s->K, K, sizeof(float) * 9);
    s->max_h = max_h;
    s->max_w = max_w;
    s->intr.fx = K[0]; s->intr.k01 = K[1]; s->intr.cx = K[2]; s->intr.fy = K[4]; s->intr.cy = K[5];
    {  // K^-1 in double
      double a[9];
      for (int i = 0; i < 9; ++i) a[i] = K[i];
      const double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) +
                         a[2] * (a[3] * a[7] - a[4] * a[6]);
      FPB_REQUIRE(std::fabs(det) > 1e-12, "camera matrix is singular");
      const double inv[9] = {(a[4] * a[8] - a[5] * a[7]) / det, (a[2] * a[7] - a[1] * a[8]) / det,
                             (a[1] * a[5] - a[2] * a[4]) / det, (a[5] * a[6] - a[3] * a[8]) / det,
                             (a[0] * a[8] - a[2] * a[6]) / det, (a[2] * a[3] - a[0] * a[5]) / det,
                             (a[3] * a[7] - a[4] * a[6]) / det, (a[1] * a[6] - a[0] * a[7]) / det,
                             (a[0] * a[4] - a[1] * a[3]) / det};
      for (int i = 0; i < 9; ++i) s->intr.kinv[i] = (float)inv[i];
    }
    FPB_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    FPB_CUDA(cudaStreamCreateWithFlags(&s->side_stream, cudaStreamNonBlocking));
    FPB_CUDA(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
    FPB_CUDA(cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming));
    for (auto& e : s->ev) FPB_CUDA(cudaEventCreate(&e));
    FPB_CUDA(cudaMallocHost(&s->host_result, sizeof(HostResult)));

    for (int i = 0; i < num_meshes; ++i) {
      const fpb200_mesh& m = meshes[i];
      FPB_REQUIRE(m.name && m.vertices && m.normals && m.texcoords && m.faces && m.texture, "mesh %d: null field", i);
      FPB_REQUIRE(m.num_vertices > 0 && m.num_faces > 0 && m.tex_h > 0 && m.tex_w > 0, "mesh %d: empty geometry / texture", i);
      FPB_REQUIRE(!s->find_mesh(m.name), "duplicate mesh name '%s'", m.name);
      std::unique_ptr<MeshEntry> e(new MeshEntry);
      e->name = m.name;
      std::vector<float> v((size_t)m.num_vertices * 3), tc((size_t)m.num_vertices * 2);
      for (int k = 0; k < m.num_vertices; ++k) {
        for (int c = 0; c < 3; ++c) v[3 * k + c] = m.vertices[3 * k + c] - m.center[c];
        tc[2 * k] = m.texcoords[2 * k];
        tc[2 * k + 1] = 1 - m.texcoords[2 * k + 1];
      }
      std::vector<float4> tex((size_t)m.tex_h * m.tex_w);
      const float sc = 1.0f / 255.0f;  // cvcuda::ConvertTo(scale = 1/255), foundationpose_render.cpp:503-506
      for (size_t k = 0; k < tex.size(); ++k)
        tex[k] = make_float4((float)m.texture[3 * k] * sc, (float)m.texture[3 * k + 1] * sc, (float)m.texture[3 * k + 2] * sc, 0.f);
      std::vector<int4> faces((size_t)m.num_faces);
      for (int k = 0; k < m.num_faces * 3; ++k)
        FPB_REQUIRE(m.faces[k] >= 0 && m.faces[k] < m.num_vertices, "mesh %d: face index out of range", i);
      for (int k = 0; k < m.num_faces; ++k) faces[k] = make_int4(m.faces[3 * k], m.faces[3 * k + 1], m.faces[3 * k + 2], 0);
      e->vertices.upload(v.data(), v.size());
      e->normals.upload(m.normals, (size_t)m.num_vertices * 3);
      e->texcoords.upload(tc.data(), tc.size());
      e->texture.upload(tex.data(), tex.size());
      e->faces.upload(faces.data(), faces.size());
      float diameter = m.diameter;
      if (!(diameter > 0)) {  // not supplied by the loader: all-pairs maximum on the device (translation-invariant)
        DevBuf<unsigned> sq;
        sq.alloc(1);
        FPB_CUDA(cudaMemsetAsync(sq.p, 0, sizeof(unsigned), s->stream));
        launch_mesh_diameter(e->vertices.p, m.num_vertices, sq.p, s->stream);
        unsigned bits = 0;
        FPB_CUDA(cudaMemcpyAsync(&bits, sq.p, sizeof(bits), cudaMemcpyDeviceToHost, s->stream));
        FPB_CUDA(cudaStreamSynchronize(s->stream));
        float d2;
        memcpy(&d2, &bits, sizeof(d2));
        diameter = sqrtf(d2);
        FPB_REQUIRE(diameter > 0, "mesh %d: all vertices coincide", i);
      }
      e->dm = DeviceMesh{e->vertices.p, e->normals.p, e->texcoords.p, e->faces.p, e->texture.p,
                         m.num_vertices, m.num_faces, m.tex_h, m.tex_w, diameter};
      s->meshes.push_back(std::move(e));
    }

    const size_t px = (size_t)max_h * max_w;
    const int cap = s->opt.max_hyps;
    s->rgb.alloc(px * 3); s->mask.alloc(px); s->depth.alloc(px); s->xyz.alloc(px * 3);
    s->f_rgb = s->rgb.p; s->f_depth = s->depth.p; s->f_xyz = s->xyz.p;
    s->filtered.alloc(px); s->sampler_scratch.alloc(px + 8);
    const std::vector<float> grid = make_rotation_grid();
    s->grid.upload(grid.data(), grid.size());
    s->poses.alloc((size_t)cap * 16); s->all_poses.alloc((size_t)cap * 16);
    s->center.alloc(4); s->status.alloc(1); s->best_index.alloc(1); s->best_pose.alloc(16);
    s->aux.alloc(cap);  // rast_id: allocated by fpb200_debug_options when a test asks for the ids
    FPB_CUDA(cudaDeviceGetAttribute(&s->num_sms, cudaDevAttrMultiProcessorCount, s->opt.device));
    s->zbuf.alloc((size_t)std::min(cap, kSplitRenderPoses) * kCrop * kCrop);
    if (s->opt.shard_world > 1) {
      FPB_REQUIRE(s->opt.shard_world <= kMaxWorld, "shard_world > %d is not supported", kMaxWorld);
      s->exch_region.alloc((size_t)2 * cap * kExchLd + kExchTailWords);
      FPB_CUDA(cudaMemset(s->exch_region.p, 0, sizeof(float) * s->exch_region.n));
    }
    s->exch_local.alloc((size_t)cap * kExchLd);
    s->exch_all.alloc(((size_t)cap + s->opt.shard_world) * kExchLd);
    s->feats.alloc((size_t)cap * 512); s->tail_qkv.alloc((size_t)cap * 1536); s->tail_att.alloc((size_t)cap * 512);
    s->tail_out.alloc((size_t)cap * 512); s->scores.alloc(cap); s->trans.alloc((size_t)cap * 3);
    s->rot.alloc((size_t)cap * 3);
    FPB_CUDA(cudaMemset(s->status.p, 0, sizeof(int)));
    const int local_cap = (cap + s->opt.shard_world - 1) / s->opt.shard_world;
    {
      size_t max_v = 0;
      for (auto& m : s->meshes) max_v = std::max(max_v, (size_t)m->dm.num_vertices);
      s->vrec.alloc((size_t)local_cap * max_v * 2);
    }
    s->arena.reset(new Arena(local_cap));
    s->refiner.reset(new Network(NetKind::kRefiner, refiner_weights, s->arena.get()));
    s->scorer.reset(new Network(NetKind::kScorer, scorer_weights, s->arena.get()));
    // the reference validates blob names at construction (foundationpose.cpp:126-139); here: head tensors
    s->refiner->weights().f32("trans.head.w", 3, 512);
    s->refiner->weights().f32("rot.head.w", 3, 512);
    s->scorer->weights().f32("lin.w", 1, 512);
    FPB_CUDA(cudaDeviceSynchronize());
    *out = s.release();
    return FPB200_OK;
  });
}

int fpb200_mesh_diameter(const float* vertices, int num_vertices, int device, float* diameter) {
  return guard([&]() -> int {
    FPB_REQUIRE(vertices && diameter && num_vertices > 0, "null / empty argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
      throw CudaError("no CUDA device available: this library has no CPU fallback");
    FPB_CUDA(cudaSetDevice(device));
    DevBuf<float> v;
    DevBuf<unsigned> sq;
    v.upload(vertices, (size_t)num_vertices * 3);
    sq.alloc(1);
    FPB_CUDA(cudaMemset(sq.p, 0, sizeof(unsigned)));
    launch_mesh_diameter(v.p, num_vertices, sq.p, nullptr);
    FPB_CUDA(cudaGetLastError());
    unsigned bits = 0;
    FPB_CUDA(cudaMemcpy(&bits, sq.p, sizeof(bits), cudaMemcpyDeviceToHost));
    float d2;
    memcpy(&d2, &bits, sizeof(d2));
    *diameter = sqrtf(d2);
    return FPB200_OK;
  });
}

void fpb200_destroy(fpb200_session* s) {
  if (!s) return;
  cudaStreamSynchronize(s->stream);
  delete s;
}

int fpb200_register(fpb200_session* s, const uint8_t* rgb, const float* depth, const uint8_t* mask, int h, int w,
                    const char* target_name, int refine_itr, float out_pose[16]) {
  FPB_NVTX("fpb200_register");
  return guard([&]() -> int {
    FPB_REQUIRE(s && out_pose, "null session / output");
    FPB_REQUIRE(s->opt.shard_world == 1 || s->exchange_ready,
                "a sharded session needs fpb200_exchange_connect before fpb200_register (or the stage API with a "
                "host-driven all-gather)");
    MeshEntry* mesh = s->find_mesh(target_name);
    FPB_REQUIRE(mesh, "unknown target '%s'", target_name ? target_name : "(null)");
    FPB_REQUIRE(refine_itr >= 0, "refine_itr must be >= 0");
    reset_call(s);
    do_upload(s, rgb, depth, h, w);
    do_mask_copy(s, mask);
    // host-side state of the call (not part of the captured launch sequence)
    s->cur_mesh = mesh;
    s->set_shard(std::min(kGridPoses, s->opt.max_hyps));
    char key[160];
    snprintf(key, sizeof(key), "register:%d:%dx%d:%p", refine_itr, h, w, (void*)mesh);
    run_maybe_graphed(s, key, [&]() {
      do_xyz(s);
      do_sample(s, mesh);
      for (int i = 0; i < refine_itr; ++i) do_refine_step(s);
      s->mark(4);
      do_score_features(s);
      s->mark(5);
      do_score_finish(s);
      s->mark(6);
      enqueue_result_read(s);
    });
    s->features_ready = true;
    return wait_result(s, out_pose, nullptr);
  });
}

int fpb200_register_async(fpb200_session* s, const char* target_name, int refine_itr) {
  FPB_NVTX("fpb200_register_async");
  return guard([&]() -> int {
    FPB_REQUIRE(s, "null session");
    FPB_REQUIRE(s->opt.shard_world == 1 || s->exchange_ready,
                "a sharded session needs fpb200_exchange_connect before fpb200_register_async");
    if (!s->uploaded) return fail(FPB200_ERR_STATE, "fpb200_upload must be called first");
    MeshEntry* mesh = s->find_mesh(target_name);
    FPB_REQUIRE(mesh, "unknown target '%s'", target_name ? target_name : "(null)");
    FPB_REQUIRE(refine_itr >= 0, "refine_itr must be >= 0");
    s->launches = 0;
    clear_status_kernel<<<1, 1, 0, s->stream>>>(s->status.p);
    s->cur_mesh = mesh;
    s->set_shard(std::min(kGridPoses, s->opt.max_hyps));
    char key[160];
    snprintf(key, sizeof(key), "register_async:%d:%dx%d:%p", refine_itr, s->img_h, s->img_w, (void*)mesh);
    run_maybe_graphed(s, key, [&]() {
      do_sample(s, mesh);
      for (int i = 0; i < refine_itr; ++i) do_refine_step(s);
      do_score_features(s);
      do_score_finish(s);
      enqueue_result_read(s);
    });
    s->features_ready = true;
    return FPB200_OK;
  });
}

int fpb200_track_async(fpb200_session* s, const uint8_t* rgb, const float* depth, int h, int w,
                       const float hyp_pose[16], const char* target_name, int refine_itr) {
  FPB_NVTX("fpb200_track_async");
  return guard([&]() -> int {
    FPB_REQUIRE(s && hyp_pose, "null session / pose");
    MeshEntry* mesh = s->find_mesh(target_name);
    FPB_REQUIRE(mesh, "unknown target '%s'", target_name ? target_name : "(null)");
    FPB_REQUIRE(refine_itr >= 0, "refine_itr must be >= 0");
    // one call in flight per session: the pinned staging (hypothesis in, pose out) is single-buffered, so a
    // second async call first lets the previous one finish (its result stays readable until then)
    if (s->pending_async) FPB_CUDA(cudaStreamSynchronize(s->stream));
    s->pending_async = true;
    reset_call(s);
    do_upload(s, rgb, depth, h, w);
    memcpy(s->host_result->hyp, hyp_pose, sizeof(float) * 16);  // pinned staging: the caller's array may go away
    FPB_CUDA(cudaMemcpyAsync(s->poses.p, s->host_result->hyp, sizeof(float) * 16, cudaMemcpyHostToDevice, s->stream));
    s->cur_mesh = mesh;
    s->n_hyp = 1; s->lo = 0; s->hi = 1; s->shard_cap = 1; s->features_ready = false;
    char key[160];
    snprintf(key, sizeof(key), "track:%d:%dx%d:%p", refine_itr, h, w, (void*)mesh);
    run_maybe_graphed(s, key, [&]() {
      do_xyz(s);
      s->mark(2);
      for (int i = 0; i < refine_itr; ++i) do_refine_step(s);
      s->mark(4);
      FPB_CUDA(cudaMemcpyAsync(s->host_result->pose, s->poses.p, sizeof(float) * 16, cudaMemcpyDeviceToHost, s->stream));
    });
    return FPB200_OK;
  });
}

int fpb200_track_wait(fpb200_session* s, float out_pose[16]) {
  return guard([&]() -> int {
    FPB_REQUIRE(s && out_pose, "null session / output");
    FPB_CUDA(cudaStreamSynchronize(s->stream));
    s->pending_async = false;
    memcpy(out_pose, s->host_result->pose, sizeof(float) * 16);
    return FPB200_OK;
  });
}

// ---- streamed tracking: SURVEY section 8 f4 (double-buffered upload, frame k + 1 in flight while frame k is refined)
int fpb200_track_stream_begin(fpb200_session* s, const float hyp_pose[16], const char* target_name, int refine_itr) {
  return guard([&]() -> int {
    FPB_REQUIRE(s && hyp_pose, "null session / pose");
    MeshEntry* mesh = s->find_mesh(target_name);
    FPB_REQUIRE(mesh, "unknown target '%s'", target_name ? target_name : "(null)");
    FPB_REQUIRE(refine_itr >= 1, "refine_itr must be >= 1");
    FPB_CUDA(cudaStreamSynchronize(s->stream));
    if (!s->copy_stream) {
      const size_t px = (size_t)s->max_h * s->max_w;
      FPB_CUDA(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
      for (auto& e : s->ev_copied) FPB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      for (auto& e : s->ev_consumed) FPB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      for (auto& e : s->ev_result) FPB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      FPB_CUDA(cudaMallocHost(&s->stream_poses, sizeof(float) * 16 * fpb200_session::kStreamRing));
      s->rgb_alt.alloc(px * 3);
      s->depth_alt.alloc(px);
      s->xyz_alt.alloc(px * 3);
    }
    memcpy(s->host_result->hyp, hyp_pose, sizeof(float) * 16);
    FPB_CUDA(cudaMemcpyAsync(s->poses.p, s->host_result->hyp, sizeof(float) * 16, cudaMemcpyHostToDevice, s->stream));
    // both frame buffer sets are free
    for (auto& e : s->ev_consumed) FPB_CUDA(cudaEventRecord(e, s->stream));
    s->stream_mesh = mesh;
    s->stream_itr = refine_itr;
    s->stream_pushed = s->stream_popped = 0;
    s->stream_open = true;
    s->pending_async = false;
    return FPB200_OK;
  });
}

int fpb200_track_stream_push(fpb200_session* s, const uint8_t* rgb, const float* depth, int h, int w) {
  return guard([&]() -> int {
    FPB_REQUIRE(s && rgb && depth, "null argument");
    if (!s->stream_open) return fail(FPB200_ERR_STATE, "fpb200_track_stream_begin must be called first");
    if (s->stream_pushed - s->stream_popped >= 2)
      return fail(FPB200_ERR_STATE, "two frames are in flight: call fpb200_track_stream_pop first");
    FPB_REQUIRE(h > 0 && w > 0 && h <= s->max_h && w <= s->max_w, "image %dx%d exceeds the session capacity %dx%d", h, w,
                s->max_h, s->max_w);
    if (h != s->img_h || w != s->img_w) {
      FPB_REQUIRE(s->stream_pushed == s->stream_popped, "the image size may only change while no frame is in flight");
      s->img_h = h; s->img_w = w; s->intr.img_h = h; s->intr.img_w = w;
      projection_matrix(s->K, h, w, s->intr.proj);
    }
    const int par = (int)(s->stream_pushed & 1);
    uint8_t* d_rgb = par ? s->rgb_alt.p : s->rgb.p;
    float* d_depth = par ? s->depth_alt.p : s->depth.p;
    // upload on the copy stream as soon as the frame that last used this buffer set has been consumed
    FPB_CUDA(cudaStreamWaitEvent(s->copy_stream, s->ev_consumed[par], 0));
    FPB_CUDA(cudaMemcpyAsync(d_rgb, rgb, (size_t)h * w * 3, cudaMemcpyHostToDevice, s->copy_stream));
    FPB_CUDA(cudaMemcpyAsync(d_depth, depth, sizeof(float) * (size_t)h * w, cudaMemcpyHostToDevice, s->copy_stream));
    FPB_CUDA(cudaEventRecord(s->ev_copied[par], s->copy_stream));
    // refine on the session stream; the hypothesis is the pose the previous frame left in device memory
    FPB_CUDA(cudaStreamWaitEvent(s->stream, s->ev_copied[par], 0));
    s->launches = 0;
    s->f_rgb = d_rgb; s->f_depth = d_depth; s->f_xyz = par ? s->xyz_alt.p : s->xyz.p;
    s->uploaded = true;
    s->cur_mesh = s->stream_mesh;
    s->n_hyp = 1; s->lo = 0; s->hi = 1; s->shard_cap = 1; s->features_ready = false;
    const int slot = (int)(s->stream_pushed % fpb200_session::kStreamRing);
    char key[160];
    snprintf(key, sizeof(key), "track_stream:%d:%dx%d:%p:%d:%d", s->stream_itr, h, w, (void*)s->stream_mesh, par, slot);
    run_maybe_graphed(s, key, [&]() {
      do_xyz(s);
      for (int i = 0; i < s->stream_itr; ++i) do_refine_step(s);
      FPB_CUDA(cudaMemcpyAsync(s->stream_poses[slot], s->poses.p, sizeof(float) * 16, cudaMemcpyDeviceToHost, s->stream));
    });
    FPB_CUDA(cudaEventRecord(s->ev_consumed[par], s->stream));
    FPB_CUDA(cudaEventRecord(s->ev_result[slot], s->stream));
    ++s->stream_pushed;
    return FPB200_OK;
  });
}

int fpb200_track_stream_pop(fpb200_session* s, float out_pose[16]) {
  return guard([&]() -> int {
    FPB_REQUIRE(s && out_pose, "null session / output");
    if (!s->stream_open || s->stream_popped >= s->stream_pushed) return fail(FPB200_ERR_STATE, "no frame is in flight");
    const int slot = (int)(s->stream_popped % fpb200_session::kStreamRing);
    FPB_CUDA(cudaEventSynchronize(s->ev_result[slot]));
    memcpy(out_pose, s->stream_poses[slot], sizeof(float) * 16);
    ++s->stream_popped;
    return FPB200_OK;
  });
}

int fpb200_track(fpb200_session* s, const uint8_t* rgb, const float* depth, int h, int w, const float hyp_pose[16],
                 const char* target_name, int refine_itr, float out_pose[16]) {
  if (!out_pose) return fail(FPB200_ERR_INVALID, "null output");
  const int rc = fpb200_track_async(s, rgb, depth, h, w, hyp_pose, target_name, refine_itr);
  return rc != FPB200_OK ? rc : fpb200_track_wait(s, out_pose);
}

int fpb200_upload(fpb200_session* s, const uint8_t* rgb, const float* depth, int h, int w) {
  return guard([&]() -> int {
    FPB_REQUIRE(s, "null session");
    reset_call(s);
    do_upload(s, rgb, depth, h, w);
    do_xyz(s);
    return FPB200_OK;
  });
}

int fpb200_sample_hypotheses(fpb200_session* s, const uint8_t* mask, const char* target_name) {
  return guard([&]() -> int {
    FPB_REQUIRE(s, "null session");
    if (!s->uploaded) return fail(FPB200_ERR_STATE, "fpb200_upload must be called first");
    MeshEntry* mesh = s->find_mesh(target_name);
    FPB_REQUIRE(mesh, "unknown target '%s'", target_name ? target_name : "(null)");
    do_mask_copy(s, mask);
    do_sample(s, mesh);
    if (s->async_stages) return FPB200_OK;  // status is checked by fpb200_wait
    FPB_CUDA(cudaMemcpyAsync(&s->host_result->status, s->status.p, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    FPB_CUDA(cudaStreamSynchronize(s->stream));
    const int st = s->host_result->status;
    if (st & kStatusEmptyMask) return fail(FPB200_ERR_NO_DEPTH, "[FoundationposeSampling] Mask is all zero.");
    if (st & kStatusNoDepth) return fail(FPB200_ERR_NO_DEPTH, "[FoundationposeSampling] No valid value in mask.");
    return FPB200_OK;
  });
}

int fpb200_set_hypotheses(fpb200_session* s, const float* poses, int n, const char* target_name) {
  return guard([&]() -> int {
    FPB_REQUIRE(s && poses, "null argument");
    FPB_REQUIRE(n >= 1 && n <= s->opt.max_hyps, "hypothesis count %d outside [1,%d]", n, s->opt.max_hyps);
    MeshEntry* mesh = s->find_mesh(target_name);
    FPB_REQUIRE(mesh, "unknown target '%s'", target_name ? target_name : "(null)");
    FPB_CUDA(cudaMemcpyAsync(s->poses.p, poses, sizeof(float) * 16 * n, cudaMemcpyHostToDevice, s->stream));
    FPB_CUDA(cudaStreamSynchronize(s->stream));
    s->cur_mesh = mesh;
    s->set_shard(n);
    return FPB200_OK;
  });
}

int fpb200_get_hypotheses(fpb200_session* s, float* poses, int capacity, int* n) {
  return guard([&]() -> int {
    FPB_REQUIRE(s && poses && n, "null argument");
    FPB_REQUIRE(capacity >= s->n_hyp, "capacity %d < %d hypotheses", capacity, s->n_hyp);
    FPB_CUDA(cudaMemcpyAsync(poses, s->poses.p, sizeof(float) * 16 * s->n_hyp, cudaMemcpyDeviceToHost, s->stream));
    FPB_CUDA(cudaStreamSynchronize(s->stream));
    *n = s->n_hyp;
    return FPB200_OK;
  });
}

int fpb200_render_and_transform(fpb200_session* s, float crop_ratio) {
  return guard([&]() -> int {
    FPB_REQUIRE(s, "null session");
    if (!s->uploaded || !s->cur_mesh) return fail(FPB200_ERR_STATE, "upload an image and set hypotheses first");
    do_render_and_transform(s, crop_ratio);
    s->stage_sync();
    return FPB200_OK;
  });
}

int fpb200_refine_step(fpb200_session* s) {
  return guard([&]() -> int {
    FPB_REQUIRE(s, "null session");
    if (!s->uploaded || !s->cur_mesh) return fail(FPB200_ERR_STATE, "upload an image and set hypotheses first");
    do_refine_step(s);
    s->stage_sync();
    return FPB200_OK;
  });
}

int fpb200_score_features(fpb200_session* s) {
  return guard([&]() -> int {
    FPB_REQUIRE(s, "null session");
    if (!s->uploaded || !s->cur_mesh) return fail(FPB200_ERR_STATE, "upload an image and set hypotheses first");
    do_score_features(s);
    s->stage_sync();
    return FPB200_OK;
  });
}

int fpb200_feature_buffers(fpb200_session* s, void** local_dev, int* n_local, int* local_offset, void** all_dev,
                           int* n_all) {
  return guard([&]() -> int {
    FPB_REQUIRE(s, "null session");
    if (local_dev) *local_dev = s->exch_local.p;
    if (n_local) *n_local = s->hi - s->lo;
    if (local_offset) *local_offset = s->lo;
    if (all_dev) *all_dev = s->exch_all.p;
    if (n_all) *n_all = s->n_hyp;
    return FPB200_OK;
  });
}

int fpb200_score_finish(fpb200_session* s, int* best_index, float out_pose[16]) {
  return guard([&]() -> int {
    FPB_REQUIRE(s, "null session");
    if (!s->features_ready) return fail(FPB200_ERR_STATE, "fpb200_score_features must be called first");
    do_score_finish(s);
    return finish_and_read(s, out_pose, best_index);
  });
}

int fpb200_read_buffer(fpb200_session* s, const char* name, void* host_dst, size_t capacity, size_t* written) {
  return guard([&]() -> int {
    FPB_REQUIRE(s && name && host_dst, "null argument");
    const std::string nm(name);
    const int nl = s->hi - s->lo;
    const void* src = nullptr;
    size_t bytes = 0;
    const size_t img_b = (size_t)kPackImageElems * 2;
    if (nm == "render_input") { src = s->arena->packed; bytes = img_b * nl; }
    else if (nm == "transf_input") { src = s->arena->packed + (size_t)nl * kPackImageElems; bytes = img_b * nl; }
    else if (nm == "xyz_map") { src = s->f_xyz; bytes = sizeof(float) * 3 * s->img_h * s->img_w; }
    else if (nm == "depth_filtered") { src = s->filtered.p; bytes = sizeof(float) * s->img_h * s->img_w; }
    else if (nm == "scores") { src = s->scores.p; bytes = sizeof(float) * s->n_hyp; }
    else if (nm == "trans") { src = s->trans.p; bytes = sizeof(float) * 3 * nl; }
    else if (nm == "rot") { src = s->rot.p; bytes = sizeof(float) * 3 * nl; }
    else if (nm == "features") { src = s->feats.p; bytes = sizeof(float) * 512 * s->n_hyp; }
    else if (nm == "rast_id") {
      FPB_REQUIRE(s->debug_ids && s->rast_id.p, "rast_id is only written after fpb200_debug_options(s, 1, ...)");
      src = s->rast_id.p; bytes = sizeof(int32_t) * kCrop * kCrop * nl;
    }
    else if (nm == "crop_tf" || nm == "bbox2d") {
      // per-hypothesis crop transform (row-major 3x3) / bbox2d (l, t, r, b) of the LAST render_and_transform
      const bool tf = nm == "crop_tf";
      bytes = sizeof(float) * (tf ? 9 : 4) * nl;
      FPB_REQUIRE(capacity >= bytes, "buffer '%s' needs %zu bytes, got %zu", name, bytes, capacity);
      FPB_CUDA(cudaStreamSynchronize(s->stream));
      std::vector<PoseAux> aux(nl);
      FPB_CUDA(cudaMemcpy(aux.data(), s->aux.p, sizeof(PoseAux) * nl, cudaMemcpyDeviceToHost));
      float* dst = static_cast<float*>(host_dst);
      for (int i = 0; i < nl; ++i)
        memcpy(dst + (size_t)i * (tf ? 9 : 4), tf ? aux[i].tf : aux[i].bbox, sizeof(float) * (tf ? 9 : 4));
      if (written) *written = bytes;
      return FPB200_OK;
    }
    else if (nm == "tokens") { src = s->arena->d0; bytes = (size_t)2 * 400 * 512 * nl; }
    else if (nm == "center") { src = s->center.p; bytes = sizeof(float) * 3; }
    else if (nm == "status") { src = s->status.p; bytes = sizeof(int); }
    else if (nm == "pose_aux") { src = s->aux.p; bytes = sizeof(PoseAux) * nl; }
    else return fail(FPB200_ERR_INVALID, "unknown buffer '%s'", name);
    FPB_REQUIRE(capacity >= bytes, "buffer '%s' needs %zu bytes, got %zu", name, bytes, capacity);
    FPB_CUDA(cudaStreamSynchronize(s->stream));
    FPB_CUDA(cudaMemcpy(host_dst, src, bytes, cudaMemcpyDeviceToHost));
    if (written) *written = bytes;
    return FPB200_OK;
  });
}

int fpb200_last_timings(fpb200_session* s, float* ms, int capacity) {
  return guard([&]() -> int {
    FPB_REQUIRE(s && ms, "null argument");
    FPB_CUDA(cudaStreamSynchronize(s->stream));
    int filled = 0;
    int prev = -1;
    for (int i = 0; i < kNumStages && filled < capacity; ++i) ms[filled++] = 0.f;
    for (int i = 0; i <= kNumStages - 1; ++i) {
      if (!s->ev_valid[i]) continue;
      if (prev >= 0 && i - 1 < capacity) {
        float t = 0.f;
        FPB_CUDA(cudaEventElapsedTime(&t, s->ev[prev], s->ev[i]));
        ms[i - 1] = t;
      }
      prev = i;
    }
    return filled;
  });
}

int fpb200_flops_per_hypothesis(fpb200_session* s, double* refiner, double* scorer) {
  return guard([&]() -> int {
    FPB_REQUIRE(s, "null session");
    if (refiner) *refiner = s->refiner->flops_per_hypothesis();
    if (scorer) *scorer = s->scorer->flops_per_hypothesis();
    return FPB200_OK;
  });
}

int fpb200_last_launch_count(fpb200_session* s) { return s ? s->launches : 0; }
long long fpb200_launch_count_total(fpb200_session* s) { return s ? s->launches_total : 0; }
void* fpb200_stream(fpb200_session* s) { return s ? (void*)s->stream : nullptr; }

int fpb200_set_async(fpb200_session* s, int on) {
  return guard([&]() -> int {
    FPB_REQUIRE(s, "null session");
    if (!on) FPB_CUDA(cudaStreamSynchronize(s->stream));
    s->async_stages = on != 0;
    return FPB200_OK;
  });
}

int fpb200_score_finish_async(fpb200_session* s) {
  return guard([&]() -> int {
    FPB_REQUIRE(s, "null session");
    if (!s->features_ready) return fail(FPB200_ERR_STATE, "fpb200_score_features must be called first");
    do_score_finish(s);
    enqueue_result_read(s);
    return FPB200_OK;
  });
}

int fpb200_wait(fpb200_session* s, int* best_index, float out_pose[16]) {
  return guard([&]() -> int {
    FPB_REQUIRE(s, "null session");
    return wait_result(s, out_pose, best_index);
  });
}

namespace {
// Wires the session's view of all regions: base[r] = device pointer to rank r's region (mapped or local).
void wire_exchange(fpb200_session* s, void* const* base) {
  const int world = s->opt.shard_world, cap = s->opt.max_hyps;
  ExchPeers& P = s->peers;
  memset(&P, 0, sizeof(P));
  for (int r = 0; r < world; ++r) {
    P.rows[r] = static_cast<float*>(base[r]);
    P.flags[r] = reinterpret_cast<unsigned*>(P.rows[r] + (size_t)2 * cap * kExchLd);
  }
  P.epoch = P.flags[s->opt.shard_rank] + kMaxWorld;
  P.world = world;
  P.rank = s->opt.shard_rank;
  P.cap_total = cap;
  s->exchange_ready = true;
  s->drop_graphs();
}
}  // namespace

int fpb200_exchange_export(fpb200_session* s, unsigned char handle[FPB200_IPC_HANDLE_BYTES]) {
  return guard([&]() -> int {
    FPB_REQUIRE(s && handle, "null argument");
    FPB_REQUIRE(s->opt.shard_world > 1 && s->exch_region.p, "fpb200_exchange_export needs a session with shard_world > 1");
    static_assert(sizeof(cudaIpcMemHandle_t) == FPB200_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    FPB_CUDA(cudaIpcGetMemHandle(&h, s->exch_region.p));
    memcpy(handle, &h, sizeof(h));
    return FPB200_OK;
  });
}

int fpb200_exchange_connect(fpb200_session* s, const unsigned char* handles) {
  return guard([&]() -> int {
    FPB_REQUIRE(s && handles, "null argument");
    FPB_REQUIRE(s->opt.shard_world > 1 && s->exch_region.p, "fpb200_exchange_connect needs a session with shard_world > 1");
    FPB_REQUIRE(!s->exchange_ready, "the exchange is already connected");
    FPB_CUDA(cudaSetDevice(s->opt.device));
    void* base[kMaxWorld] = {};
    for (int r = 0; r < s->opt.shard_world; ++r) {
      if (r == s->opt.shard_rank) {
        base[r] = s->exch_region.p;
        continue;
      }
      cudaIpcMemHandle_t h;
      memcpy(&h, handles + (size_t)r * FPB200_IPC_HANDLE_BYTES, sizeof(h));
      void* ptr = nullptr;
      FPB_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
      s->ipc_opened.push_back(ptr);
      base[r] = ptr;
    }
    wire_exchange(s, base);
    return FPB200_OK;
  });
}

int fpb200_exchange_connect_local(fpb200_session* const* sessions, int world) {
  return guard([&]() -> int {
    FPB_REQUIRE(sessions && world > 1 && world <= kMaxWorld, "bad session list");
    void* base[kMaxWorld] = {};
    for (int r = 0; r < world; ++r) {
      fpb200_session* s = sessions[r];
      FPB_REQUIRE(s && s->opt.shard_world == world && s->opt.shard_rank == r && s->exch_region.p,
                  "session %d is not rank %d of %d", r, r, world);
      FPB_REQUIRE(s->opt.max_hyps == sessions[0]->opt.max_hyps, "all ranks must have the same capacity");
      base[r] = s->exch_region.p;
    }
    for (int r = 0; r < world; ++r) {
      fpb200_session* s = sessions[r];
      FPB_CUDA(cudaSetDevice(s->opt.device));
      for (int q = 0; q < world; ++q) {
        const int dq = sessions[q]->opt.device;
        if (dq == s->opt.device) continue;
        int can = 0;
        FPB_CUDA(cudaDeviceCanAccessPeer(&can, s->opt.device, dq));
        FPB_REQUIRE(can, "device %d cannot access device %d", s->opt.device, dq);
        const cudaError_t e = cudaDeviceEnablePeerAccess(dq, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) FPB_CUDA(e);
        (void)cudaGetLastError();
      }
      wire_exchange(s, base);
    }
    return FPB200_OK;
  });
}

int fpb200_debug_time_stage(fpb200_session* s, int stage, int repeat, float* ms_out) {
  return guard([&]() -> int {
    FPB_REQUIRE(s && ms_out && repeat >= 1, "bad argument");
    if (!s->uploaded || !s->cur_mesh || s->n_hyp <= 0) return fail(FPB200_ERR_STATE, "upload an image and set hypotheses first");
    const int nl = s->hi - s->lo;
    const MeshEntry& m = *s->cur_mesh;
    auto run = [&]() {
      switch (stage) {
        case FPB200_STAGE_RENDER:
          launch_pose_setup(s->poses.p + 16 * s->lo, nl, s->intr, m.dm.diameter, kRefineCropRatio, s->aux.p, s->stream);
          launch_render(m.dm, s->aux.p, nl, s->vrec.p, s->arena->packed, 0, nullptr, s->status.p, s->zbuf.p,
                        kSplitRenderPoses, s->num_sms, s->render_split, s->stream);
          break;
        case FPB200_STAGE_WARP:
          launch_warp(s->rgb.p, s->xyz.p, s->img_h, s->img_w, s->aux.p, nl, m.dm.diameter, s->arena->packed, nl, s->stream);
          break;
        case FPB200_STAGE_EXCHANGE:
          FPB_REQUIRE(s->exchange_ready, "the exchange is not connected");
          exchange_begin_kernel<<<1, 1, 0, s->stream>>>(s->peers.epoch);
          exchange_push_kernel<<<s->peers.world, 1024, 0, s->stream>>>(s->peers, s->feats.p, s->poses.p + 16 * s->lo, s->lo, nl);
          exchange_wait_kernel<<<1, 32, 0, s->stream>>>(s->peers, s->status.p);
          exchange_collect_kernel<<<16, 256, 0, s->stream>>>(s->peers, s->feats.p, s->all_poses.p, s->n_hyp, s->status.p);
          break;
        default:
          throw std::invalid_argument("unknown stage");
      }
      FPB_CUDA(cudaGetLastError());
    };
    run();  // warm-up
    cudaEvent_t e0, e1;
    FPB_CUDA(cudaEventCreate(&e0));
    FPB_CUDA(cudaEventCreate(&e1));
    FPB_CUDA(cudaEventRecord(e0, s->stream));
    for (int i = 0; i < repeat; ++i) run();
    FPB_CUDA(cudaEventRecord(e1, s->stream));
    FPB_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    FPB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms_out = ms / repeat;
    return FPB200_OK;
  });
}

int fpb200_debug_options(fpb200_session* s, int write_rast_ids, int render_split) {
  return guard([&]() -> int {
    FPB_REQUIRE(s, "null session");
    FPB_REQUIRE(render_split >= -1 && render_split <= 1, "render_split must be -1 (automatic), 0 or 1");
    FPB_CUDA(cudaStreamSynchronize(s->stream));
    if (write_rast_ids && !s->rast_id.p) {
      const int local_cap = (s->opt.max_hyps + s->opt.shard_world - 1) / s->opt.shard_world;
      s->rast_id.alloc((size_t)local_cap * kCrop * kCrop);
    }
    s->debug_ids = write_rast_ids != 0;
    s->render_split = render_split;
    s->drop_graphs();  // captured launch sequences carry the old pointers / kernel choice
    return FPB200_OK;
  });
}

int fpb200_set_gemm_profile(fpb200_session* s, int on) {
  return guard([&]() -> int {
    FPB_REQUIRE(s, "null session");
    s->profile_gemm = on != 0;
    s->refiner->set_profile(on != 0);
    s->scorer->set_profile(on != 0);
    return FPB200_OK;
  });
}

int fpb200_gemm_profile(fpb200_session* s, double* ms, double* flops, long long* launches) {
  return guard([&]() -> int {
    FPB_REQUIRE(s, "null session");
    double m1 = 0, f1 = 0, m2 = 0, f2 = 0;
    long long l1 = 0, l2 = 0;
    s->refiner->collect_profile(&m1, &f1, &l1);
    s->scorer->collect_profile(&m2, &f2, &l2);
    if (ms) *ms = m1 + m2;
    if (flops) *flops = f1 + f2;
    if (launches) *launches = l1 + l2;
    return FPB200_OK;
  });
}

}  // extern "C"
