// Layer graph of the refine / score networks (see network.h, SURVEY.md §2.3 for the structure).
#include "network.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>

#include "common.h"
#include "encoder_block.h"
#include "net_ops.h"
#include "stem_conv.h"

namespace fpb {

// ------------------------------------------------------------------ weight file
namespace {
struct RawEntry {
  char name[48];
  uint32_t dtype, ndim, dims[4];
  uint64_t offset, nbytes;
};
static_assert(sizeof(RawEntry) == 48 + 4 + 4 + 16 + 8 + 8, "packed header entry");
}  // namespace

WeightFile::WeightFile(const std::string& path, int expected_kind) : path_(path) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) throw std::runtime_error("cannot open weight file " + path);
  std::vector<uint8_t> buf;
  fseek(f, 0, SEEK_END);
  const long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  buf.resize(sz > 0 ? sz : 0);
  const size_t got = fread(buf.data(), 1, buf.size(), f);
  fclose(f);
  if (got != buf.size() || buf.size() < 16 || memcmp(buf.data(), "FPB2W001", 8) != 0)
    throw std::runtime_error("not an FPB2W001 weight file: " + path);
  uint32_t kind, count;
  memcpy(&kind, buf.data() + 8, 4);
  memcpy(&count, buf.data() + 12, 4);
  if ((int)kind != expected_kind)
    throw std::runtime_error("weight file " + path + " holds the wrong network kind");
  if (16 + (size_t)count * sizeof(RawEntry) > buf.size()) throw std::runtime_error("truncated weight file " + path);
  FPB_CUDA(cudaMalloc(&blob_, buf.size()));
  FPB_CUDA(cudaMemcpy(blob_, buf.data(), buf.size(), cudaMemcpyHostToDevice));
  for (uint32_t i = 0; i < count; ++i) {
    RawEntry e;
    memcpy(&e, buf.data() + 16 + (size_t)i * sizeof(RawEntry), sizeof(e));
    if (e.offset + e.nbytes > buf.size() || e.ndim > 4) throw std::runtime_error("corrupt weight file " + path);
    WeightTensor t;
    t.dev = static_cast<uint8_t*>(blob_) + e.offset;
    t.dtype = (int)e.dtype;
    t.ndim = (int)e.ndim;
    for (int d = 0; d < 4; ++d) t.dims[d] = (int)e.dims[d];
    tensors_[std::string(e.name, strnlen(e.name, sizeof(e.name)))] = t;
  }
  host_ = std::move(buf);
}

WeightFile::~WeightFile() {
  if (blob_) cudaFree(blob_);
}

const WeightTensor& WeightFile::get(const std::string& name, int dtype, int d0, int d1) const {
  auto it = tensors_.find(name);
  if (it == tensors_.end()) throw std::runtime_error("weight file " + path_ + " has no tensor '" + name + "'");
  const WeightTensor& t = it->second;
  if (t.dtype != dtype || t.dims[0] != d0 || (d1 >= 0 && t.dims[1] != d1)) {
    char msg[256];
    snprintf(msg, sizeof(msg), "tensor '%s' has dtype %d dims [%d,%d], expected dtype %d dims [%d,%d]",
             name.c_str(), t.dtype, t.dims[0], t.dims[1], dtype, d0, d1);
    throw std::runtime_error(msg);
  }
  return t;
}

// ------------------------------------------------------------------ arena
Arena::Arena(int max_hyps) : cap(max_hyps) {
  const size_t n = (size_t)max_hyps;
  struct Item { void** p; size_t bytes; };
  const size_t tok = n * 400 * 512 * 2;
  const size_t stem_out = 2 * n * 80 * 80 * 64 * 2;
  act_bytes = stem_out;
  std::vector<Item> items = {
      {(void**)&packed, (2 * n * (size_t)kPackImageElems + (size_t)kStemPitch * kPackC) * 2},
      {(void**)&act[0], stem_out}, {(void**)&act[1], stem_out}, {(void**)&act[2], stem_out}, {(void**)&act[3], stem_out},
      {(void**)&d0, tok},
      {(void**)&qkv, 3 * tok}, {(void**)&att, tok}, {(void**)&proj, tok}, {(void**)&x1, tok},
      {(void**)&ff1, tok}, {(void**)&ff2, tok}, {(void**)&x2, tok},
      {(void**)&mean_t, n * 512 * 4}, {(void**)&mean_r, n * 512 * 4},
      {(void**)&mean_part, (size_t)n * 25 * 512 * 4},  // 25 group sums of 16 tokens per hypothesis (>= 8 slices)
  };
  size_t total = 0;
  for (auto& it : items) total += (it.bytes + 1023) / 1024 * 1024;
  FPB_CUDA(cudaMalloc(&blob, total));
  bytes = total;
  size_t off = 0;
  for (auto& it : items) {
    *it.p = static_cast<uint8_t*>(blob) + off;
    off += (it.bytes + 1023) / 1024 * 1024;
  }
  // the stem reads zero pad pixels that no kernel ever writes
  FPB_CUDA(cudaMemset(packed, 0, (2 * n * (size_t)kPackImageElems + (size_t)kStemPitch * kPackC) * 2));
}

Arena::~Arena() {
  if (blob) cudaFree(blob);
}

// ------------------------------------------------------------------ network
struct Network::Plan {
  std::vector<std::function<void(cudaStream_t)>> steps;
  std::vector<double> gemm_flops;  // per step; 0 for steps that are not conv_gemm launches
  double flops = 0;
  int launches = 0;                // kernel launches of one forward
};

Network::Network(NetKind kind, const std::string& weight_path, Arena* arena)
    : kind_(kind), w_(new WeightFile(weight_path, (int)kind)), arena_(arena) {
  if (const char* e = conv_gemm_init()) throw std::runtime_error(e);
  if (const char* e = attention_init()) throw std::runtime_error(e);
  if (const char* e = net_ops_init()) throw std::runtime_error(e);
  if (const char* e = encoder_block_init()) throw std::runtime_error(e);
  if (const char* e = stem_conv_init()) throw std::runtime_error(e);
  {
    // the backbone's layer list is data: "arch" f32 [L, 4] rows (section, kind, cout, stride)
    const WeightTensor* a = w_->find("arch");
    if (!a || a->dtype != 1 || a->ndim != 2 || a->dims[1] != 4 || a->dims[0] < 2 || a->dims[0] > 64)
      throw std::runtime_error("weight file " + weight_path + " carries no valid 'arch' table (re-export it with "
                               "foundationpose_cpp_b200/weights.py)");
    const float* v = static_cast<const float*>(w_->host_ptr(*a));
    for (int i = 0; i < a->dims[0]; ++i)
      arch_.push_back(ArchLayer{(int)v[4 * i], (int)v[4 * i + 1], (int)v[4 * i + 2], (int)v[4 * i + 3]});
    const ArchLayer& s0 = arch_[0];
    if (s0.sec != 0 || s0.kind != 0 || s0.cout != 64 || s0.stride != 2)
      throw std::runtime_error("arch: the first layer must be the 7x7 stride-2 stem with 64 outputs");
  }
  {
    std::vector<__half> reordered(64 * 448);
    reorder_stem_weights(w_->host_f16("enc.A.0.w", 64, 448), reordered.data());
    FPB_CUDA(cudaMalloc(&stem_w_, reordered.size() * sizeof(__half)));
    FPB_CUDA(cudaMemcpy(stem_w_, reordered.data(), reordered.size() * sizeof(__half), cudaMemcpyHostToDevice));
  }
  flops_per_hyp_ = plan(1).flops;
}

Network::~Network() {
  if (stem_w_) cudaFree(stem_w_);
  for (auto e : ev_pool_) cudaEventDestroy(e);
}

namespace {
std::vector<double>* g_gemm_flops_sink = nullptr;  // set while a plan is being built
void add_gemm(std::vector<std::function<void(cudaStream_t)>>& steps, double* flops, const ConvGemmDesc& d) {
  ConvGemmPlan pl;
  if (const char* e = build_conv_gemm(d, &pl)) throw std::runtime_error(e);
  *flops += pl.flops;
  steps.push_back([pl](cudaStream_t s) { FPB_CUDA(launch_conv_gemm(pl, s)); });
  if (g_gemm_flops_sink) {
    g_gemm_flops_sink->resize(steps.size(), 0.0);
    (*g_gemm_flops_sink)[steps.size() - 1] = pl.flops;
  }
}
void add_attention(std::vector<std::function<void(cudaStream_t)>>& steps, const __half* qkv, __half* out, int n) {
  AttentionPlan ap;
  if (const char* e = build_attention(qkv, out, n, 1536, 512, &ap)) throw std::runtime_error(e);
  steps.push_back([ap](cudaStream_t s) { launch_attention(ap, s); });
}
}  // namespace

void Network::add_backbone(Plan& p, int n) {
  Arena& A = *arena_;
  const WeightFile& W = *w_;
  const int n2 = 2 * n;
  auto conv = [&](const __half* in, int imgs, int hw, int cin, bool s2, const std::string& name, int cout,
                  __half* out, const __half* res, int res_ld, bool phase_out, int out_ld, int img_mod, int half_coff,
                  const float* post) {
    ConvGemmDesc d;
    d.in = in; d.n_img = imgs; d.cin = cin;
    const int ho = s2 ? hw / 2 : hw;
    d.in_h = ho; d.in_w = ho; d.planes = s2 ? 4 : 1;
    d.out_h = ho; d.out_w = ho;
    // 128-row tiles = bw x bh pixels of bn images.  Maps that are not multiples of 8 pixels (20 x 20) take 4 x 4 pixels
    // of 8 images - unless the launch has fewer images than that (Track, N <= 2 hypotheses): 25 tiles per image
    // GROUP with 1/8 .. 4/8 of their rows in use cost two waves of CTAs on the 512-channel layers, so few images take
    // partial 8 x 8 x 2 (9 tiles per image pair) or 16 x 8 x 1 (6 per image) tiles instead.  The accumulation order of
    // an output element does not depend on the tile shape.
    if (ho % 8 == 0) { d.bw = 8; d.bh = 8; d.bn = 2; }
    else if (imgs == 1) { d.bw = 16; d.bh = 8; d.bn = 1; }
    else if (imgs <= 4) { d.bw = 8; d.bh = 8; d.bn = 2; }
    else { d.bw = 4; d.bh = 4; d.bn = 8; }
    if (s2) taps_3x3_s2_phase(&d); else taps_3x3_s1(&d);
    d.weight = W.f16(name + ".w", cout, 9 * cin); d.cout = cout; d.bias = W.f32(name + ".b", cout);
    d.residual = res; d.res_ld = res_ld; d.post_add = post;
    d.out = out; d.relu = true; d.out_phase_split = phase_out; d.out_ld = out_ld; d.img_mod = img_mod;
    d.half_coff = half_coff;
    add_gemm(p.steps, &p.flops, d);
  };
  // ---- the layer list comes from the weight file (arch_): buffers are handed out from a pool of four, the layout
  // each layer writes follows from its consumer (phase-split in front of a stride-2 convolution, the two branches
  // side by side after the last encodeA layer, tokens + positional table at the very end)
  struct Cur { const __half* buf; int imgs, hw, c; } cur{A.packed, n2, kCrop, 8};
  int pool_next = 0;
  auto take = [&](const __half* live0, const __half* live1) -> __half* {
    for (int k = 0; k < 4; ++k) {
      __half* cand = A.act[(pool_next + k) & 3];
      if (cand != live0 && cand != live1) {
        pool_next = (pool_next + k + 1) & 3;
        return cand;
      }
    }
    throw std::logic_error("backbone: activation pool exhausted");
  };
  auto first_stride_after = [&](size_t li) { return li + 1 < arch_.size() && arch_[li + 1].kind != 2 ? arch_[li + 1].stride : 1; };
  auto fits = [&](int imgs, int hw, int c) {
    if ((size_t)imgs * hw * hw * c * 2 > A.act_bytes) throw std::runtime_error("arch: an activation exceeds the arena buffers");
  };
  static const bool stem_as_gemm = getenv("FPB200_STEM_GEMM") != nullptr;  // A/B measurements only
  size_t n_a = 0;  // layers of encodeA (tensor names count from 0 inside each section, like nn.Sequential)
  while (n_a < arch_.size() && arch_[n_a].sec == 0) ++n_a;
  for (size_t li = 0; li < arch_.size(); ++li) {
    const ArchLayer& L = arch_[li];
    const bool last = li + 1 == arch_.size();
    const bool last_a = L.sec == 0 && (last || arch_[li + 1].sec == 1);
    const bool phase = !last && first_stride_after(li) == 2;
    const std::string name = std::string("enc.") + (L.sec == 0 ? "A." : "AB.") + std::to_string(L.sec == 0 ? li : li - n_a);
    if (L.cout % 64) throw std::runtime_error("arch: channel counts must be multiples of 64");
    if (li > 0 && arch_[li].sec < arch_[li - 1].sec) throw std::runtime_error("arch: encodeA layers must precede encodeAB");
    const int out_ld = last_a ? 2 * L.cout : 0, img_mod = last_a ? n : 0, half_coff = last_a ? L.cout : 0;
    const float* post = last ? W.f32("pos.table", 400, 512) : nullptr;
    if (L.kind == 0) {
      if (li != 0) throw std::runtime_error("arch: the 7x7 stem can only be the first layer");
      if (last_a || last) throw std::runtime_error("arch: the stem cannot be the last layer of a section");
      __half* out = take(nullptr, nullptr);
      fits(n2, 80, 64);
      if (!stem_as_gemm && phase) {
        // 7x7 s2, (6 -> 64), packed input -> phase-split 80x80x64: direct tcgen05 convolution reading its A operand in
        // place from the input patch (stem_conv.cu)
        StemConvPlan sp;
        if (const char* e = build_stem_conv(A.packed, n2, stem_w_, W.host_f32(name + ".b", 64), out, &sp)) throw std::runtime_error(e);
        p.flops += sp.flops;
        p.steps.push_back([sp](cudaStream_t s) { FPB_CUDA(launch_stem_conv(sp, s)); });
        p.gemm_flops.resize(p.steps.size(), 0.0);
        p.gemm_flops.back() = sp.flops;  // timed with the tcgen05 GEMM family (fpb200_set_gemm_profile)
      } else {  // the same layer through the general implicit-GEMM kernel: one 64-wide K slab per filter row
        ConvGemmDesc d;
        d.in = A.packed;
        d.n_img = n2; d.planes = 2; d.in_h = 80; d.in_w = 80; d.cin = 64;
        const long long pitch_b = (long long)kStemPitch * kPackC * 2;
        d.stride_w = 32; d.stride_plane = pitch_b; d.stride_h = 2 * pitch_b; d.stride_img = kCrop * pitch_b;
        d.out_h = 80; d.out_w = 80; d.bw = 16; d.bh = 8; d.bn = 1;
        taps_stem7_rows(&d);
        d.weight = W.f16(name + ".w", 64, 448); d.cout = 64; d.bias = W.f32(name + ".b", 64);
        d.out = out; d.relu = true; d.out_phase_split = phase;
        add_gemm(p.steps, &p.flops, d);
        // count the stem at its algorithmic size (7*7*6 MACs per output), not the padded K of 448
        p.flops -= 2.0 * n2 * 80 * 80 * 64 * (448 - 294);
        p.gemm_flops.back() -= 2.0 * n2 * 80 * 80 * 64 * (448 - 294);
      }
      cur = Cur{out, n2, 80, 64};
    } else if (L.kind == 1) {
      if (L.stride != 1 && L.stride != 2) throw std::runtime_error("arch: stride must be 1 or 2");
      if (cur.buf == A.packed) throw std::runtime_error("arch: the first layer must be the stem");
      const int ho = L.stride == 2 ? cur.hw / 2 : cur.hw;
      if (ho % 4) throw std::runtime_error("arch: feature maps must stay multiples of 4 pixels");
      __half* out = last ? A.d0 : take(cur.buf, nullptr);
      if (!last) fits(cur.imgs, ho, L.cout);
      conv(cur.buf, cur.imgs, cur.hw, cur.c, L.stride == 2, name, L.cout, out, nullptr, 0, phase, out_ld, img_mod, half_coff, post);
      cur = Cur{out, last_a ? n : cur.imgs, ho, last_a ? 2 * L.cout : L.cout};
    } else if (L.kind == 2) {
      if (L.cout != cur.c) throw std::runtime_error("arch: a residual block keeps its channel count");
      __half* t1 = take(cur.buf, nullptr);
      conv(cur.buf, cur.imgs, cur.hw, cur.c, false, name + ".c1", L.cout, t1, nullptr, 0, false, 0, 0, 0, nullptr);
      __half* out = last ? A.d0 : take(cur.buf, t1);
      conv(t1, cur.imgs, cur.hw, cur.c, false, name + ".c2", L.cout, out, cur.buf, cur.c, phase, out_ld, img_mod, half_coff, post);
      cur = Cur{out, last_a ? n : cur.imgs, cur.hw, last_a ? 2 * L.cout : L.cout};
    } else {
      throw std::runtime_error("arch: unknown layer kind");
    }
  }
  if (cur.imgs != n || cur.hw != 20 || cur.c != 512)
    throw std::runtime_error("arch: the backbone must end in 20 x 20 x 512 tokens per hypothesis (the transformer heads are fixed)");
}

namespace {
ConvGemmDesc linear_desc(const __half* in, long long rows, int k, const __half* w, const float* b, int outs,
                         __half* out, bool relu, const __half* residual = nullptr) {
  ConvGemmDesc d;
  d.in = in; d.n_img = 1; d.planes = 1; d.in_h = 1; d.in_w = (int)rows; d.cin = k;
  d.out_h = 1; d.out_w = (int)rows; d.bw = 128; d.bh = 1; d.bn = 1;
  d.num_taps = 1;
  d.weight = w; d.cout = outs; d.bias = b; d.out = out; d.relu = relu;
  d.residual = residual; d.res_ld = outs;
  return d;
}
}  // namespace

void Network::add_encoder_layer(Plan& p, int n, const std::string& h, float* mean_out) {
  Arena& A = *arena_;
  const WeightFile& W = *w_;
  const long long rows = (long long)n * 400;
  add_gemm(p.steps, &p.flops, linear_desc(A.d0, rows, 512, W.f16(h + ".qkv.w", 1536, 512), W.f32(h + ".qkv.b", 1536), 1536, A.qkv, false));
  add_attention(p.steps, A.qkv, A.att, n);
  p.flops += 2.0 * n * 4 * 2 * 400.0 * 400.0 * 128.0;
  const float *g1 = W.f32(h + ".ln1.g", 512), *b1 = W.f32(h + ".ln1.b", 512);
  const float *g2 = W.f32(h + ".ln2.g", 512), *b2 = W.f32(h + ".ln2.b", 512);
  static const bool unfused = getenv("FPB200_UNFUSED_ENCODER") != nullptr;  // A/B measurements only
  if (!unfused) {
    // out-projection + residual + LayerNorm1 + FFN + residual + LayerNorm2 + token mean: one kernel that keeps
    // each 128-token row block on chip through the three GEMMs (encoder_block.cu)
    EncoderBlockDesc d;
    d.att = A.att; d.x0 = A.d0; d.n_hyp = n;
    d.w_out = W.f16(h + ".out.w", 512, 512);
    d.w1 = W.f16(h + ".ff1.w", 512, 512);
    d.w2 = W.f16(h + ".ff2.w", 512, 512);
    // the fused kernel takes the vectors by value: read them from the host copy of the file, so that building a plan
    // never synchronises with the device (a peer session of the same process may be spinning on it)
    d.b_out = W.host_f32(h + ".out.b", 512); d.b1 = W.host_f32(h + ".ff1.b", 512); d.b2 = W.host_f32(h + ".ff2.b", 512);
    d.ln1_g = W.host_f32(h + ".ln1.g", 512); d.ln1_b = W.host_f32(h + ".ln1.b", 512);
    d.ln2_g = W.host_f32(h + ".ln2.g", 512); d.ln2_b = W.host_f32(h + ".ln2.b", 512);
    d.vectors_on_host = true;
    d.partial = A.mean_part; d.mean_out = mean_out;
    EncoderBlockPlan ep;
    if (const char* e = build_encoder_block(d, &ep)) throw std::runtime_error(e);
    p.flops += ep.flops;
    p.steps.push_back([ep](cudaStream_t s) { FPB_CUDA(launch_encoder_block(ep, s)); });
    p.launches += 1;  // the step launches two kernels
  } else {
    // post-LN encoder layer, layer by layer: the two residual sums are formed in the GEMM epilogues
    add_gemm(p.steps, &p.flops, linear_desc(A.att, rows, 512, W.f16(h + ".out.w", 512, 512), W.f32(h + ".out.b", 512), 512, A.proj, false, A.d0));
    p.steps.push_back([&A, rows, g1, b1](cudaStream_t s) { launch_layernorm(A.proj, g1, b1, A.x1, rows, s); });
    add_gemm(p.steps, &p.flops, linear_desc(A.x1, rows, 512, W.f16(h + ".ff1.w", 512, 512), W.f32(h + ".ff1.b", 512), 512, A.ff1, true));
    add_gemm(p.steps, &p.flops, linear_desc(A.ff1, rows, 512, W.f16(h + ".ff2.w", 512, 512), W.f32(h + ".ff2.b", 512), 512, A.ff2, false, A.x1));
    // the second LayerNorm's tokens are consumed only through their mean: fused, never stored
    p.steps.push_back([&A, n, g2, b2, mean_out](cudaStream_t s) { launch_layernorm_mean(A.ff2, g2, b2, mean_out, n, A.mean_part, s); });
    if (n < 128) p.launches += 1;  // sliced LayerNorm + mean: two kernels
  }
  p.flops += 2.0 * n * 512 * 3;  // Linear(512,3) on the token mean
}

Network::Plan& Network::plan(int n) {
  auto it = plans_.find(n);
  if (it != plans_.end()) return *it->second;
  if (n < 1 || n > arena_->cap) throw std::invalid_argument("network: hypothesis count exceeds the session capacity");
  std::unique_ptr<Plan> p(new Plan);
  g_gemm_flops_sink = &p->gemm_flops;
  add_backbone(*p, n);
  Arena& A = *arena_;
  if (kind_ == NetKind::kRefiner) {
    add_encoder_layer(*p, n, "trans", A.mean_t);
    add_encoder_layer(*p, n, "rot", A.mean_r);
  } else {
    const WeightFile& W = *w_;
    const long long rows = (long long)n * 400;
    add_gemm(p->steps, &p->flops, linear_desc(A.d0, rows, 512, W.f16("att.qkv.w", 1536, 512), W.f32("att.qkv.b", 1536), 1536, A.qkv, false));
    add_attention(p->steps, A.qkv, A.att, n);
    p->flops += 2.0 * n * 4 * 2 * 400.0 * 400.0 * 128.0;
    p->steps.push_back([&A, n](cudaStream_t s) { launch_token_mean(A.att, A.mean_t, n, s); });
    // out_proj is applied to the token mean by the session (it commutes with the mean)
    p->flops += 2.0 * n * 512 * 512;
  }
  g_gemm_flops_sink = nullptr;
  p->gemm_flops.resize(p->steps.size(), 0.0);
  p->launches += (int)p->steps.size();
  p->flops /= n;
  Plan& ref = *p;
  plans_[n] = std::move(p);
  return ref;
}

void Network::forward(int n, cudaStream_t s) {
  Plan& p = plan(n);
  if (!profile_) {
    for (auto& st : p.steps) st(s);
    return;
  }
  p.gemm_flops.resize(p.steps.size(), 0.0);
  for (size_t i = 0; i < p.steps.size(); ++i) {
    if (p.gemm_flops[i] <= 0) {
      p.steps[i](s);
      continue;
    }
    while (ev_pool_.size() < ev_used_ + 2) {
      cudaEvent_t e;
      FPB_CUDA(cudaEventCreate(&e));
      ev_pool_.push_back(e);
    }
    FPB_CUDA(cudaEventRecord(ev_pool_[ev_used_], s));
    p.steps[i](s);
    FPB_CUDA(cudaEventRecord(ev_pool_[ev_used_ + 1], s));
    ev_used_ += 2;
    ev_flops_.push_back(p.gemm_flops[i]);
  }
}

void Network::collect_profile(double* ms, double* flops, long long* launches) {
  double t = 0, f = 0;
  for (size_t i = 0; i + 1 < ev_used_; i += 2) {
    FPB_CUDA(cudaEventSynchronize(ev_pool_[i + 1]));
    float e = 0.f;
    FPB_CUDA(cudaEventElapsedTime(&e, ev_pool_[i], ev_pool_[i + 1]));
    t += e;
    f += ev_flops_[i / 2];
  }
  if (ms) *ms = t;
  if (flops) *flops = f;
  if (launches) *launches = (long long)(ev_used_ / 2);
  ev_used_ = 0;
  ev_flops_.clear();
}

int Network::launches_per_forward(int n) { return plan(n).launches; }

}  // namespace fpb
