// ViT frame-encoder plan: packed weights + workspace + the forward schedule (C ABI of seam A).
//
// Forward (reference: CLIPModel.forward, D/train/train_vid_score/video/clip.py:141-163, and the timm
// flavour configured at D/train/train_v68/.../sscd.py:78):
//   im2row -> patch GEMM (+bias +pos, rows remapped around the class slot) -> class rows -> [ln_pre]
//   L x { LN1 -> QKV GEMM -> fused MHSA -> proj GEMM (+residual) -> LN2 -> fc1 GEMM (+act) -> fc2 GEMM (+residual) }
//   -> tail: ln_post tokens | ln_post+GeM+Linear (fused) | ln_post -> conv1x1 GEMM -> GeM+Linear.
// Residual stream, LayerNorm statistics, softmax, GeM and all accumulators are fp32; GEMM operands bf16.
// spec.precision == VSCB200_PRECISION_FP32: every bf16 operand buffer (weights and activations) holds two planes, hi at
// the base and lo = bf16(x - hi) `lo_off` elements behind it; the projections run the split-bf16 GEMM and attention runs
// in fp32 (attention_fp32.cu).
// No allocation after create(): the workspace and the host-API staging are sized for max_frames.
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "host_util.h"
#include "kernels.h"

using namespace vscb200;

struct LayerW {
  float *ln1_w = nullptr, *ln1_b = nullptr, *qkv_b = nullptr, *proj_b = nullptr, *ln2_w = nullptr, *ln2_b = nullptr,
        *fc1_b = nullptr, *fc2_b = nullptr;
  void *qkv_w = nullptr, *proj_w = nullptr, *fc1_w = nullptr, *fc2_w = nullptr;   // bf16
};

struct vscb200_vit {
  vscb200_vit_spec spec;
  int max_frames = 0, T = 0, P = 0, W = 0, Kp = 0, Kraw = 0;
  void* patch_w = nullptr;   // bf16 [W, Kp]
  float *patch_b = nullptr, *cls = nullptr, *pos = nullptr, *ln_pre_w = nullptr, *ln_pre_b = nullptr;
  std::vector<LayerW> layers;
  float *ln_post_w = nullptr, *ln_post_b = nullptr;
  void* gem_conv_w = nullptr;   // bf16 [gem_hidden, W]
  float *gem_conv_b = nullptr, *head_w = nullptr, *head_b = nullptr;
  // workspace
  float* x = nullptr;        // [max_frames*T, W] fp32 residual stream
  void* h = nullptr;         // bf16 [M, W]
  void* qkv = nullptr;       // bf16 [M, 3W]
  void* ao = nullptr;        // bf16 [M, W]
  void* u = nullptr;         // bf16 [M, 4W]
  void* patches = nullptr;   // bf16 [max_frames*P, Kp]
  float* y = nullptr;        // fp32 [M, gem_hidden] (gem_conv tail only)
  // host-API staging
  float* frames_stage2[2] = {nullptr, nullptr};
  float* out_stage2[2] = {nullptr, nullptr};
  float* out_pinned2[2] = {nullptr, nullptr};   // page-locked landing buffers: a D2H into pageable memory would block the host
  cudaStream_t own_stream = nullptr, in_stream = nullptr, out_stream = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
  std::vector<void*> allocs;
  std::map<std::string, bool> loaded;
  bool exact = false;        // fp32-equivalent mode: bf16 buffers are (hi, lo) plane pairs
  int planes = 1;
};

namespace {

template <typename Tp>
int dev_alloc(vscb200_vit* m, Tp** p, size_t bytes) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, bytes ? bytes : 16);
  if (e != cudaSuccess) {
    set_last_error(std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
    return VSCB200_ERR_NOMEM;
  }
  m->allocs.push_back(q);
  *p = reinterpret_cast<Tp*>(q);
  return VSCB200_OK;
}

int64_t out_elems(const vscb200_vit* m) {
  return m->spec.tail == VSCB200_TAIL_TOKENS ? static_cast<int64_t>(m->T) * m->W : m->spec.out_dim;
}

// lo plane of a bf16 buffer of `elems` elements per plane (nullptr in the bf16 mode)
inline void* lo_plane(const vscb200_vit* m, void* hi, int64_t elems) {
  return m->exact ? static_cast<void*>(static_cast<uint16_t*>(hi) + elems) : nullptr;
}

int host_staging(vscb200_vit* m) {
  const int64_t in_per = 3LL * m->spec.img * m->spec.img, out_per = out_elems(m);
  VSCB_CUDA_OK(cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking));
  VSCB_CUDA_OK(cudaStreamCreateWithFlags(&m->in_stream, cudaStreamNonBlocking));
  VSCB_CUDA_OK(cudaStreamCreateWithFlags(&m->out_stream, cudaStreamNonBlocking));
  for (int b = 0; b < 2; ++b) {
    VSCB_CUDA_OK(cudaMalloc(&m->frames_stage2[b], static_cast<size_t>(m->max_frames) * in_per * 4));
    VSCB_CUDA_OK(cudaMalloc(&m->out_stage2[b], static_cast<size_t>(m->max_frames) * out_per * 4));
    VSCB_CUDA_OK(cudaMallocHost(&m->out_pinned2[b], static_cast<size_t>(m->max_frames) * out_per * 4));
    VSCB_CUDA_OK(cudaEventCreateWithFlags(&m->ev_in[b], cudaEventDisableTiming));
    VSCB_CUDA_OK(cudaEventCreateWithFlags(&m->ev_comp[b], cudaEventDisableTiming));
    VSCB_CUDA_OK(cudaEventCreateWithFlags(&m->ev_out[b], cudaEventDisableTiming));
  }
  return VSCB200_OK;
}

}  // namespace

extern "C" {

int vscb200_vit_create(const vscb200_vit_spec* spec, int max_frames, vscb200_vit** out) {
  VSCB_REQUIRE(spec && out, "vit_create: null argument");
  VSCB_REQUIRE(spec->img > 0 && spec->patch > 0 && spec->img % spec->patch == 0 && spec->patch % 2 == 0,
               "vit_create: img must be a multiple of an even patch size");
  VSCB_REQUIRE(spec->heads > 0 && spec->width == spec->heads * 64, "vit_create: width must be heads*64");
  VSCB_REQUIRE(spec->width % 8 == 0 && spec->width <= 1024, "vit_create: width must be <= 1024");
  VSCB_REQUIRE(spec->layers > 0 && max_frames > 0, "vit_create: layers/max_frames must be positive");
  VSCB_REQUIRE(spec->tail >= 0 && spec->tail <= 2, "vit_create: unknown tail");
  VSCB_REQUIRE(spec->tail == VSCB200_TAIL_TOKENS || spec->out_dim > 0, "vit_create: out_dim must be positive");
  VSCB_REQUIRE(spec->precision == VSCB200_PRECISION_BF16 || spec->precision == VSCB200_PRECISION_FP32,
               "vit_create: precision must be VSCB200_PRECISION_BF16 or VSCB200_PRECISION_FP32");
  vscb200_vit* m = new vscb200_vit();
  m->spec = *spec;
  m->max_frames = max_frames;
  m->exact = spec->precision == VSCB200_PRECISION_FP32;
  m->planes = m->exact ? 2 : 1;
  const int g = spec->img / spec->patch;
  m->P = g * g;
  m->T = m->P + 1;
  m->W = spec->width;
  m->Kraw = 3 * spec->patch * spec->patch;
  m->Kp = (m->Kraw + 63) / 64 * 64;
  m->layers.resize(spec->layers);
  const int W = m->W;
  const size_t M = static_cast<size_t>(max_frames) * m->T;
  int rc = 0;
#define A(ptr, bytes) if ((rc = dev_alloc(m, &(ptr), (bytes)))) { vscb200_vit_destroy(m); return rc; }
  const size_t P2 = 2 * static_cast<size_t>(m->planes);      // bytes per bf16 operand element (both planes)
  A(m->patch_w, static_cast<size_t>(W) * m->Kp * P2);
  A(m->patch_b, W * 4);
  A(m->cls, W * 4);
  A(m->pos, static_cast<size_t>(m->T) * W * 4);
  A(m->ln_pre_w, W * 4); A(m->ln_pre_b, W * 4);
  A(m->ln_post_w, W * 4); A(m->ln_post_b, W * 4);
  for (auto& l : m->layers) {
    A(l.ln1_w, W * 4); A(l.ln1_b, W * 4); A(l.ln2_w, W * 4); A(l.ln2_b, W * 4);
    A(l.qkv_b, 3 * W * 4); A(l.proj_b, W * 4); A(l.fc1_b, 4 * W * 4); A(l.fc2_b, W * 4);
    A(l.qkv_w, static_cast<size_t>(3) * W * W * P2); A(l.proj_w, static_cast<size_t>(W) * W * P2);
    A(l.fc1_w, static_cast<size_t>(4) * W * W * P2); A(l.fc2_w, static_cast<size_t>(4) * W * W * P2);
  }
  if (spec->tail == VSCB200_TAIL_GEM_LINEAR) {
    A(m->head_w, static_cast<size_t>(spec->out_dim) * W * 4); A(m->head_b, spec->out_dim * 4);
  } else if (spec->tail == VSCB200_TAIL_GEM_CONV_LINEAR) {
    VSCB_REQUIRE(spec->gem_hidden > 0 && spec->gem_hidden % 8 == 0, "vit_create: gem_hidden must be a multiple of 8");
    A(m->gem_conv_w, static_cast<size_t>(spec->gem_hidden) * W * P2); A(m->gem_conv_b, spec->gem_hidden * 4);
    A(m->head_w, static_cast<size_t>(spec->out_dim) * spec->gem_hidden * 4); A(m->head_b, spec->out_dim * 4);
    A(m->y, M * spec->gem_hidden * 4);
  }
  A(m->x, M * W * 4);
  A(m->h, M * W * P2);
  A(m->qkv, M * 3 * W * P2);
  A(m->ao, M * W * P2);
  A(m->u, M * 4 * W * P2);
  A(m->patches, static_cast<size_t>(max_frames) * m->P * m->Kp * P2);
#undef A
  // zero-initialised optional parameters (patch bias absent in the CLIP flavour)
  cudaMemset(m->patch_b, 0, W * 4);
  if ((rc = host_staging(m))) { vscb200_vit_destroy(m); return rc; }
  *out = m;
  return VSCB200_OK;
}

void vscb200_vit_destroy(vscb200_vit* m) {
  if (!m) return;
  for (void* p : m->allocs) cudaFree(p);
  for (int b = 0; b < 2; ++b) {
    if (m->frames_stage2[b]) cudaFree(m->frames_stage2[b]);
    if (m->out_stage2[b]) cudaFree(m->out_stage2[b]);
    if (m->out_pinned2[b]) cudaFreeHost(m->out_pinned2[b]);
    if (m->ev_in[b]) cudaEventDestroy(m->ev_in[b]);
    if (m->ev_comp[b]) cudaEventDestroy(m->ev_comp[b]);
    if (m->ev_out[b]) cudaEventDestroy(m->ev_out[b]);
  }
  if (m->own_stream) cudaStreamDestroy(m->own_stream);
  if (m->in_stream) cudaStreamDestroy(m->in_stream);
  if (m->out_stream) cudaStreamDestroy(m->out_stream);
  delete m;
}

int64_t vscb200_vit_out_elems_per_frame(const vscb200_vit* m) { return m ? out_elems(m) : 0; }

int vscb200_vit_set_param(vscb200_vit* m, const char* name_c, const float* w, int64_t count, void* stream_v) {
  VSCB_REQUIRE(m && name_c && w, "vit_set_param: null argument");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const std::string name(name_c);
  const int W = m->W;
  struct Target { void* dst; int64_t rows; int cols; int ld; bool bf16; };
  Target t{nullptr, 0, 0, 0, false};
  auto f32 = [&](float* dst, int64_t n) { t = Target{dst, 1, static_cast<int>(n), static_cast<int>(n), false}; };
  auto b16 = [&](void* dst, int64_t rows, int cols, int ld) { t = Target{dst, rows, cols, ld, true}; };
  if (name == "patch_w") b16(m->patch_w, W, m->Kraw, m->Kp);
  else if (name == "patch_b") f32(m->patch_b, W);
  else if (name == "cls") f32(m->cls, W);
  else if (name == "pos") f32(m->pos, static_cast<int64_t>(m->T) * W);
  else if (name == "ln_pre_w") f32(m->ln_pre_w, W);
  else if (name == "ln_pre_b") f32(m->ln_pre_b, W);
  else if (name == "ln_post_w") f32(m->ln_post_w, W);
  else if (name == "ln_post_b") f32(m->ln_post_b, W);
  else if (name == "head_w" && m->head_w)
    f32(m->head_w, static_cast<int64_t>(m->spec.out_dim) *
                       (m->spec.tail == VSCB200_TAIL_GEM_CONV_LINEAR ? m->spec.gem_hidden : W));
  else if (name == "head_b" && m->head_b) f32(m->head_b, m->spec.out_dim);
  else if (name == "gem_conv_w" && m->gem_conv_w) b16(m->gem_conv_w, m->spec.gem_hidden, W, W);
  else if (name == "gem_conv_b" && m->gem_conv_b) f32(m->gem_conv_b, m->spec.gem_hidden);
  else if (name.size() > 2 && name[0] == 'l' && name.find('.') != std::string::npos) {
    const size_t dot = name.find('.');
    const int li = atoi(name.substr(1, dot - 1).c_str());
    VSCB_REQUIRE(li >= 0 && li < static_cast<int>(m->layers.size()), "vit_set_param: layer index out of range");
    LayerW& l = m->layers[li];
    const std::string f = name.substr(dot + 1);
    if (f == "ln1_w") f32(l.ln1_w, W); else if (f == "ln1_b") f32(l.ln1_b, W);
    else if (f == "ln2_w") f32(l.ln2_w, W); else if (f == "ln2_b") f32(l.ln2_b, W);
    else if (f == "qkv_b") f32(l.qkv_b, 3 * W); else if (f == "proj_b") f32(l.proj_b, W);
    else if (f == "fc1_b") f32(l.fc1_b, 4 * W); else if (f == "fc2_b") f32(l.fc2_b, W);
    else if (f == "qkv_w") b16(l.qkv_w, 3 * W, W, W); else if (f == "proj_w") b16(l.proj_w, W, W, W);
    else if (f == "fc1_w") b16(l.fc1_w, 4 * W, W, W); else if (f == "fc2_w") b16(l.fc2_w, W, 4 * W, 4 * W);
  }
  if (!t.dst) {
    set_last_error("vit_set_param: unknown or inapplicable parameter '" + name + "'");
    return VSCB200_ERR_INVALID;
  }
  const int64_t expect = t.rows * t.cols;
  if (count != expect) {
    set_last_error("vit_set_param: '" + name + "' expects " + std::to_string(expect) + " elements, got " +
                   std::to_string(count));
    return VSCB200_ERR_INVALID;
  }
  if (t.bf16) {
    int rc = cast_f32_bf16_padded(w, t.dst, t.rows, t.cols, t.ld, stream, lo_plane(m, t.dst, t.rows * t.ld));
    if (rc) return rc;
  } else {
    VSCB_CUDA_OK(cudaMemcpyAsync(t.dst, w, count * 4, cudaMemcpyDeviceToDevice, stream));
  }
  m->loaded[name] = true;
  return VSCB200_OK;
}

static int forward_chunk(vscb200_vit* m, const float* frames, int n, float* out, cudaStream_t s) {
  const vscb200_vit_spec& sp = m->spec;
  const int W = m->W, T = m->T, P = m->P;
  const int64_t M = static_cast<int64_t>(n) * T;
  // plane strides (elements) of the operand buffers, fixed by the plan's capacity; 0 in the bf16 mode
  const int64_t Mcap = static_cast<int64_t>(m->max_frames) * T;
  const int64_t h_lo = m->exact ? Mcap * W : 0, qkv_lo = m->exact ? Mcap * 3 * W : 0, u_lo = m->exact ? Mcap * 4 * W : 0;
  const int64_t pat_lo = m->exact ? static_cast<int64_t>(m->max_frames) * P * m->Kp : 0;
  auto lo = [&](void* hi, int64_t off) -> void* { return off ? static_cast<void*>(static_cast<uint16_t*>(hi) + off) : nullptr; };
  auto wlo = [&](void* w, int64_t rows, int64_t ld) -> const void* { return lo_plane(m, w, rows * ld); };
  int rc;
#define R(call) if ((rc = (call))) return rc
  R(im2row(frames, m->patches, n, sp.img, sp.patch, m->Kp, s, pat_lo));
  R(gemm_bf16(m->patches, m->patch_w, sp.patch_bias ? m->patch_b : nullptr, m->x, static_cast<int64_t>(n) * P, W, m->Kp,
              m->Kp, m->Kp, W, VSCB_EPI_PATCH_F32_ID, -1, s, m->pos, P, false, 0, nullptr, lo(m->patches, pat_lo),
              wlo(m->patch_w, W, m->Kp), nullptr));
  R(cls_rows(m->cls, m->pos, m->x, n, T, W, s));
  if (sp.pre_norm) R(layernorm(m->x, m->ln_pre_w, m->ln_pre_b, m->x, M, W, sp.ln_eps, 0, s));
  // zig-zag walk (kernels.h): every kernel consumes its input starting from the rows its producer wrote last
  static const bool zigzag = [] { const char* e = getenv("VSCB200_ZIGZAG"); return e ? atoi(e) != 0 : true; }();
  bool rev = false;
  auto dir = [&]() { const bool r = rev; if (zigzag) rev = !rev; return r; };
  for (const LayerW& l : m->layers) {
    R(layernorm(m->x, l.ln1_w, l.ln1_b, m->h, M, W, sp.ln_eps, 1, s, dir(), h_lo));
    R(gemm_bf16(m->h, l.qkv_w, l.qkv_b, m->qkv, M, 3 * W, W, W, W, 3 * W, VSCB200_EPI_BF16, -1, s, nullptr, 0, dir(), 0, nullptr,
                lo(m->h, h_lo), wlo(l.qkv_w, 3 * W, W), lo(m->qkv, qkv_lo)));
    if (m->exact) {
      dir();
      R(attention_fp32(m->qkv, qkv_lo, m->ao, h_lo, n, T, sp.heads, 64, 0.125f, nullptr, 0, 0, 0, 0, 0, s));
    } else {
      R(attention(m->qkv, m->ao, n, T, sp.heads, 64, s, dir()));
    }
    R(gemm_bf16(m->ao, l.proj_w, l.proj_b, m->x, M, W, W, W, W, W, VSCB200_EPI_RESIDUAL_F32, -1, s, nullptr, 0, dir(), 0, nullptr,
                lo(m->ao, h_lo), wlo(l.proj_w, W, W), nullptr));
    R(layernorm(m->x, l.ln2_w, l.ln2_b, m->h, M, W, sp.ln_eps, 1, s, dir(), h_lo));
    R(gemm_bf16(m->h, l.fc1_w, l.fc1_b, m->u, M, 4 * W, W, W, W, 4 * W, VSCB200_EPI_BF16, sp.act, s, nullptr, 0, dir(), 0, nullptr,
                lo(m->h, h_lo), wlo(l.fc1_w, 4 * W, W), lo(m->u, u_lo)));
    R(gemm_bf16(m->u, l.fc2_w, l.fc2_b, m->x, M, W, 4 * W, 4 * W, 4 * W, W, VSCB200_EPI_RESIDUAL_F32, -1, s, nullptr, 0, dir(), 0,
                nullptr, lo(m->u, u_lo), wlo(l.fc2_w, W, 4 * W), nullptr));
  }
  if (sp.tail == VSCB200_TAIL_TOKENS) {
    R(layernorm(m->x, m->ln_post_w, m->ln_post_b, out, M, W, sp.ln_eps, 0, s));
  } else if (sp.tail == VSCB200_TAIL_GEM_LINEAR) {
    R(gem_head(m->x, m->ln_post_w, m->ln_post_b, m->head_w, m->head_b, out, n, T, W, sp.out_dim, sp.ln_eps, sp.gem_p,
               true, s));
  } else {
    R(layernorm(m->x, m->ln_post_w, m->ln_post_b, m->h, M, W, sp.ln_eps, 1, s, false, h_lo));
    R(gemm_bf16(m->h, m->gem_conv_w, m->gem_conv_b, m->y, M, sp.gem_hidden, W, W, W, sp.gem_hidden, VSCB200_EPI_F32, -1,
                s, nullptr, 0, false, 0, nullptr, lo(m->h, h_lo), wlo(m->gem_conv_w, sp.gem_hidden, W), nullptr));
    R(gem_head(m->y, nullptr, nullptr, m->head_w, m->head_b, out, n, T, sp.gem_hidden, sp.out_dim, sp.ln_eps,
               sp.gem_p, false, s));
  }
#undef R
  return VSCB200_OK;
}

int vscb200_vit_forward(vscb200_vit* m, const float* frames, int64_t n, float* out, void* stream_v) {
  VSCB_REQUIRE(m && (n == 0 || (frames && out)), "vit_forward: null argument");
  VSCB_REQUIRE(n >= 0, "vit_forward: negative frame count");
  cudaStream_t s = static_cast<cudaStream_t>(stream_v);
  const int64_t in_per = 3LL * m->spec.img * m->spec.img, out_per = out_elems(m);
  for (int64_t f0 = 0; f0 < n; f0 += m->max_frames) {
    const int nc = static_cast<int>(n - f0 < m->max_frames ? n - f0 : m->max_frames);
    int rc = forward_chunk(m, frames + f0 * in_per, nc, out + f0 * out_per, s);
    if (rc) return rc;
  }
  return VSCB200_OK;
}

int vscb200_vit_forward_host(vscb200_vit* m, const float* frames_host, int64_t n, float* out_host) {
  // Double-buffered pipeline: H2D of chunk i+1 (copy-in stream) and D2H of chunk i-1 (copy-out
  // stream) overlap the encoder kernels of chunk i (compute stream).  Overlap needs pinned host
  // buffers; with pageable memory the copies serialise but the result is the same.
  VSCB_REQUIRE(m && (n == 0 || (frames_host && out_host)), "vit_forward_host: null argument");
  const int64_t in_per = 3LL * m->spec.img * m->spec.img, out_per = out_elems(m);
  int64_t chunk = 0, last_f0 = 0;
  int last_nc = 0, last_b = 0;
  for (int64_t f0 = 0; f0 < n; f0 += m->max_frames, ++chunk) {
    const int b = static_cast<int>(chunk & 1);
    const int nc = static_cast<int>(n - f0 < m->max_frames ? n - f0 : m->max_frames);
    if (chunk >= 2) VSCB_CUDA_OK(cudaStreamWaitEvent(m->in_stream, m->ev_comp[b], 0));    // input buffer b is free
    VSCB_CUDA_OK(cudaMemcpyAsync(m->frames_stage2[b], frames_host + f0 * in_per, static_cast<size_t>(nc) * in_per * 4,
                                 cudaMemcpyHostToDevice, m->in_stream));
    VSCB_CUDA_OK(cudaEventRecord(m->ev_in[b], m->in_stream));
    VSCB_CUDA_OK(cudaStreamWaitEvent(m->own_stream, m->ev_in[b], 0));
    if (chunk >= 2) VSCB_CUDA_OK(cudaStreamWaitEvent(m->own_stream, m->ev_out[b], 0));    // output buffer b is drained
    int rc = forward_chunk(m, m->frames_stage2[b], nc, m->out_stage2[b], m->own_stream);
    if (rc) return rc;
    VSCB_CUDA_OK(cudaEventRecord(m->ev_comp[b], m->own_stream));
    VSCB_CUDA_OK(cudaStreamWaitEvent(m->out_stream, m->ev_comp[b], 0));
    // D2H lands in a page-locked buffer (a copy into the caller's pageable array would block this thread until the
    // chunk's kernels finish, so the next chunk's H2D would no longer overlap them); the previous chunk's descriptors
    // are handed to the caller while this chunk computes.
    VSCB_CUDA_OK(cudaMemcpyAsync(m->out_pinned2[b], m->out_stage2[b], static_cast<size_t>(nc) * out_per * 4,
                                 cudaMemcpyDeviceToHost, m->out_stream));
    VSCB_CUDA_OK(cudaEventRecord(m->ev_out[b], m->out_stream));
    if (chunk >= 1) {
      VSCB_CUDA_OK(cudaEventSynchronize(m->ev_out[b ^ 1]));
      memcpy(out_host + (f0 - m->max_frames) * out_per, m->out_pinned2[b ^ 1], static_cast<size_t>(m->max_frames) * out_per * 4);
    }
    last_f0 = f0; last_nc = nc; last_b = b;
  }
  if (last_nc > 0) {
    VSCB_CUDA_OK(cudaEventSynchronize(m->ev_out[last_b]));
    memcpy(out_host + last_f0 * out_per, m->out_pinned2[last_b], static_cast<size_t>(last_nc) * out_per * 4);
  }
  VSCB_CUDA_OK(cudaStreamSynchronize(m->own_stream));
  return VSCB200_OK;
}

}  // extern "C"
