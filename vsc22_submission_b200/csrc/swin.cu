// Swin-V2 frame-encoder plan: packed weights + workspace + forward schedule (C ABI of seam A, second family).
//
// Reference: SwinTransformerV2.forward_features, swinv2.py:619-633, as configured by config_v106.py:8-24
// (SwinV2-B, 256x256, window 16, depths 2/2/18/2, heads 4/8/16/32 -> head_dim 32 in every stage, 512-D out):
//   im2row -> patch GEMM (+bias) -> LN                                               (PatchEmbed :460-498)
//   per stage, per block (shift 0 / ws/2 alternating):
//     gather rows into (shifted) window order as bf16 -> QKV GEMM (+(q_bias,0,v_bias), q/k L2-normalised per head,
//     q * exp(min(logit_scale, log 100)) in the epilogue) -> tcgen05 window attention (+CPB bias table, +region mask)
//     -> proj GEMM -> x += LN(.) scattered back through the inverse window map        (res-post-norm, :273-303)
//     cast -> fc1 GEMM (+erf-GELU) -> fc2 GEMM -> x += LN(.)                           (:306)
//   per stage end: 2x2 gather -> reduction GEMM (no bias) -> LN                       (PatchMerging :353-368)
//   tail: LN -> GeM(p) over the tokens -> Linear                                      (:630-632, :664-665)
// Residual stream, LayerNorm, softmax and accumulators fp32; GEMM / attention operands bf16.
// spec.precision == VSCB200_PRECISION_FP32: bf16 operand buffers hold (hi, lo) plane pairs, projections run the
// split-bf16 GEMM, window attention runs in fp32 (attention_fp32.cu) -- as in vit.cu.  Window sides other than 4 / 8 / 16
// (24: SwinV2-L@384, BASELINE configs[3]) use the fp32 attention kernel in both modes.
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "host_util.h"
#include "kernels.h"

using namespace vscb200;

namespace vscb200 {
int window_gather_bf16(const float* x, void* h, int64_t n, int res, int ws, int shift, int C, cudaStream_t stream,
                       int64_t lo_off = 0);
// h_next != nullptr: the updated rows are also written as bf16 in the row order of the next consumer (ws_next x ws_next
// windows shifted by shift_next; ws_next == res, shift 0: token order)
int ln_residual_scatter(const float* y, const float* gamma, const float* beta, float* x, int64_t n, int res, int ws, int shift,
                        int C, float eps, cudaStream_t stream, void* h_next, int ws_next, int shift_next, int64_t lo_off = 0);
int patch_merge_gather(const float* x, void* out, int64_t n, int res, int C, cudaStream_t stream, int64_t lo_off = 0);
int cpb_table(const float* w0, const float* b0, const float* w2, float* table, int ws, int pretrained_ws, int heads,
              cudaStream_t stream);
int swin_prep(const float* logit_scale, const float* q_bias, const float* v_bias, float* qscale, float* qkv_bias, int heads,
              int C, cudaStream_t stream);
int swin_attention(const void* qkv, void* out, const float* tables, int64_t n_windows_total, int nW_per_frame, int nWx, int ws,
                   int shift, int heads, cudaStream_t stream);
}  // namespace vscb200

struct SwinBlockW {
  float *norm1_w = nullptr, *norm1_b = nullptr, *norm2_w = nullptr, *norm2_b = nullptr;
  float *logit_scale = nullptr, *cpb0_w = nullptr, *cpb0_b = nullptr, *cpb2_w = nullptr, *q_bias = nullptr, *v_bias = nullptr;
  float *proj_b = nullptr, *fc1_b = nullptr, *fc2_b = nullptr;
  void *qkv_w = nullptr, *proj_w = nullptr, *fc1_w = nullptr, *fc2_w = nullptr;   // bf16
  // derived at finalize()
  float *table = nullptr, *qscale = nullptr, *qkv_bias = nullptr;
};

struct SwinStageW {
  int C = 0, res = 0, ws = 0, heads = 0, pretrained_ws = 0;
  std::vector<SwinBlockW> blocks;
  void* red_w = nullptr;                      // bf16 [2C, 4C]
  float *red_norm_w = nullptr, *red_norm_b = nullptr;
};

struct vscb200_swin {
  vscb200_swin_spec spec;
  int max_frames = 0, Kp = 0, Kraw = 0, Cf = 0;
  void* patch_w = nullptr;                     // bf16 [embed, Kp]
  float *patch_b = nullptr, *patch_norm_w = nullptr, *patch_norm_b = nullptr;
  std::vector<SwinStageW> stages;
  float *norm_w = nullptr, *norm_b = nullptr, *head_w = nullptr, *head_b = nullptr;
  // workspace (sized by stage 0, where rows * C is largest)
  float* x = nullptr;       // [M0, C0] fp32 residual stream
  float* y = nullptr;       // [M0, C0] fp32 branch output before its post-norm
  void* h = nullptr;        // bf16 [M0, C0]   (also the 2x2-merged rows [M0/4, 4 C0])
  void* qkv = nullptr;      // bf16 [M0, 3 C0]
  void* ao = nullptr;       // bf16 [M0, C0]
  void* u = nullptr;        // bf16 [M0, 4 C0]
  void* patches = nullptr;  // bf16 [M0, Kp]
  float* frames_stage2[2] = {nullptr, nullptr};
  float* out_stage2[2] = {nullptr, nullptr};
  float* out_pinned2[2] = {nullptr, nullptr};
  cudaStream_t own_stream = nullptr, in_stream = nullptr, out_stream = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
  bool finalized = false;
  bool exact = false;        // fp32-equivalent mode: bf16 buffers are (hi, lo) plane pairs
  int planes = 1;
  std::vector<void*> allocs;
  std::map<std::string, bool> loaded;
};

namespace {

template <typename Tp>
int sw_alloc(vscb200_swin* m, Tp** p, size_t bytes) {
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

inline void* lo_plane(const vscb200_swin* m, void* hi, int64_t elems) {
  return m->exact ? static_cast<void*>(static_cast<uint16_t*>(hi) + elems) : nullptr;
}

int host_staging(vscb200_swin* m) {
  const int64_t in_per = 3LL * m->spec.img * m->spec.img, out_per = m->spec.out_dim;
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

int finalize(vscb200_swin* m, cudaStream_t s) {
  if (m->finalized) return VSCB200_OK;
  for (auto& st : m->stages) {
    for (auto& b : st.blocks) {
      int rc = cpb_table(b.cpb0_w, b.cpb0_b, b.cpb2_w, b.table, st.ws, st.pretrained_ws, st.heads, s);
      if (rc) return rc;
      if ((rc = swin_prep(b.logit_scale, b.q_bias, b.v_bias, b.qscale, b.qkv_bias, st.heads, st.C, s))) return rc;
    }
  }
  m->finalized = true;
  return VSCB200_OK;
}

}  // namespace

extern "C" {

int vscb200_swin_create(const vscb200_swin_spec* spec, int max_frames, vscb200_swin** out) {
  VSCB_REQUIRE(spec && out, "swin_create: null argument");
  VSCB_REQUIRE(spec->n_stages >= 1 && spec->n_stages <= 4, "swin_create: 1..4 stages");
  VSCB_REQUIRE(spec->img > 0 && spec->patch > 0 && spec->patch % 2 == 0 && spec->img % spec->patch == 0,
               "swin_create: img must be a multiple of an even patch size");
  VSCB_REQUIRE(spec->embed % 64 == 0, "swin_create: embed dim must be a multiple of 64 (head pairs of 2 x 32)");
  VSCB_REQUIRE(max_frames > 0 && spec->out_dim > 0, "swin_create: max_frames / out_dim must be positive");
  VSCB_REQUIRE(spec->precision == VSCB200_PRECISION_BF16 || spec->precision == VSCB200_PRECISION_FP32,
               "swin_create: precision must be VSCB200_PRECISION_BF16 or VSCB200_PRECISION_FP32");
  vscb200_swin* m = new vscb200_swin();
  m->spec = *spec;
  m->max_frames = max_frames;
  m->exact = spec->precision == VSCB200_PRECISION_FP32;
  m->planes = m->exact ? 2 : 1;
  const size_t P2 = 2 * static_cast<size_t>(m->planes);      // bytes per bf16 operand element (both planes)
  m->Kraw = 3 * spec->patch * spec->patch;
  m->Kp = (m->Kraw + 63) / 64 * 64;
  const int res0 = spec->img / spec->patch;
  int rc = 0;
#define A(ptr, bytes) if ((rc = sw_alloc(m, &(ptr), (bytes)))) { vscb200_swin_destroy(m); return rc; }
  m->stages.resize(spec->n_stages);
  for (int i = 0; i < spec->n_stages; ++i) {
    SwinStageW& st = m->stages[i];
    st.C = spec->embed << i;
    st.res = res0 >> i;
    st.ws = spec->window < st.res ? spec->window : st.res;
    st.heads = spec->heads[i];
    st.pretrained_ws = spec->pretrained_windows[i];
    if (!(st.heads * 32 == st.C && st.ws >= 2 && st.ws <= 32 && st.res % st.ws == 0 && st.C <= 2048 && (st.C <= 1024 || st.C % 256 == 0) &&
          (res0 % (1 << i)) == 0 && spec->depths[i] > 0 && (st.res <= spec->window || spec->window % 2 == 0))) {
      set_last_error("swin_create: every stage needs head_dim 32, a window side in [2, 32] dividing the token map (even when shifted), width <= 2048");
      vscb200_swin_destroy(m);
      return VSCB200_ERR_INVALID;
    }
    const size_t C = st.C;
    const int ts = 2 * st.ws - 1;
    st.blocks.resize(spec->depths[i]);
    for (auto& b : st.blocks) {
      A(b.norm1_w, C * 4); A(b.norm1_b, C * 4); A(b.norm2_w, C * 4); A(b.norm2_b, C * 4);
      A(b.logit_scale, st.heads * 4); A(b.cpb0_w, 512 * 2 * 4); A(b.cpb0_b, 512 * 4); A(b.cpb2_w, st.heads * 512 * 4);
      A(b.q_bias, C * 4); A(b.v_bias, C * 4); A(b.proj_b, C * 4); A(b.fc1_b, 4 * C * 4); A(b.fc2_b, C * 4);
      A(b.qkv_w, 3 * C * C * P2); A(b.proj_w, C * C * P2); A(b.fc1_w, 4 * C * C * P2); A(b.fc2_w, 4 * C * C * P2);
      A(b.table, static_cast<size_t>(st.heads) * ts * ts * 4); A(b.qscale, st.heads * 4); A(b.qkv_bias, 3 * C * 4);
    }
    if (i + 1 < spec->n_stages) {
      A(st.red_w, 2 * C * 4 * C * P2); A(st.red_norm_w, 2 * C * 4); A(st.red_norm_b, 2 * C * 4);
    }
  }
  m->Cf = spec->embed << (spec->n_stages - 1);
  A(m->patch_w, static_cast<size_t>(spec->embed) * m->Kp * P2);
  A(m->patch_b, spec->embed * 4); A(m->patch_norm_w, spec->embed * 4); A(m->patch_norm_b, spec->embed * 4);
  A(m->norm_w, m->Cf * 4); A(m->norm_b, m->Cf * 4);
  A(m->head_w, static_cast<size_t>(spec->out_dim) * m->Cf * 4); A(m->head_b, spec->out_dim * 4);
  const size_t M0 = static_cast<size_t>(max_frames) * res0 * res0, C0 = spec->embed;
  A(m->x, M0 * C0 * 4); A(m->y, M0 * C0 * 4); A(m->h, M0 * C0 * P2); A(m->qkv, M0 * 3 * C0 * P2); A(m->ao, M0 * C0 * P2);
  A(m->u, M0 * 4 * C0 * P2); A(m->patches, M0 * m->Kp * P2);
#undef A
  if ((rc = host_staging(m))) { vscb200_swin_destroy(m); return rc; }
  *out = m;
  return VSCB200_OK;
}

void vscb200_swin_destroy(vscb200_swin* m) {
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

int vscb200_swin_out_dim(const vscb200_swin* m) { return m ? m->spec.out_dim : 0; }

/* names = the reference's state-dict parameter names (swinv2.py) */
int vscb200_swin_set_param(vscb200_swin* m, const char* name_c, const float* w, int64_t count, void* stream_v) {
  VSCB_REQUIRE(m && name_c && w, "swin_set_param: null argument");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const std::string name(name_c);
  struct Target { void* dst; int64_t rows; int cols; int ld; bool bf16; };
  Target t{nullptr, 0, 0, 0, false};
  auto f32 = [&](float* dst, int64_t n) { t = Target{dst, 1, static_cast<int>(n), static_cast<int>(n), false}; };
  auto b16 = [&](void* dst, int64_t rows, int cols, int ld) { t = Target{dst, rows, cols, ld, true}; };
  const int E = m->spec.embed;
  if (name == "patch_embed.proj.weight") b16(m->patch_w, E, m->Kraw, m->Kp);
  else if (name == "patch_embed.proj.bias") f32(m->patch_b, E);
  else if (name == "patch_embed.norm.weight") f32(m->patch_norm_w, E);
  else if (name == "patch_embed.norm.bias") f32(m->patch_norm_b, E);
  else if (name == "norm.weight") f32(m->norm_w, m->Cf);
  else if (name == "norm.bias") f32(m->norm_b, m->Cf);
  else if (name == "output_proj.weight") f32(m->head_w, static_cast<int64_t>(m->spec.out_dim) * m->Cf);
  else if (name == "output_proj.bias") f32(m->head_b, m->spec.out_dim);
  else if (name.rfind("layers.", 0) == 0) {
    int si = -1, bi = -1, pos = 0;
    if (sscanf(name_c, "layers.%d.blocks.%d.%n", &si, &bi, &pos) == 2 && pos > 0 && si >= 0 &&
        si < static_cast<int>(m->stages.size()) && bi >= 0 && bi < static_cast<int>(m->stages[si].blocks.size())) {
      SwinStageW& st = m->stages[si];
      SwinBlockW& b = st.blocks[bi];
      const int C = st.C;
      const std::string f = name.substr(pos);
      if (f == "norm1.weight") f32(b.norm1_w, C); else if (f == "norm1.bias") f32(b.norm1_b, C);
      else if (f == "norm2.weight") f32(b.norm2_w, C); else if (f == "norm2.bias") f32(b.norm2_b, C);
      else if (f == "attn.logit_scale") f32(b.logit_scale, st.heads);
      else if (f == "attn.cpb_mlp.0.weight") f32(b.cpb0_w, 1024); else if (f == "attn.cpb_mlp.0.bias") f32(b.cpb0_b, 512);
      else if (f == "attn.cpb_mlp.2.weight") f32(b.cpb2_w, st.heads * 512);
      else if (f == "attn.q_bias") f32(b.q_bias, C); else if (f == "attn.v_bias") f32(b.v_bias, C);
      else if (f == "attn.qkv.weight") b16(b.qkv_w, 3 * C, C, C);
      else if (f == "attn.proj.weight") b16(b.proj_w, C, C, C); else if (f == "attn.proj.bias") f32(b.proj_b, C);
      else if (f == "mlp.fc1.weight") b16(b.fc1_w, 4 * C, C, C); else if (f == "mlp.fc1.bias") f32(b.fc1_b, 4 * C);
      else if (f == "mlp.fc2.weight") b16(b.fc2_w, C, 4 * C, 4 * C); else if (f == "mlp.fc2.bias") f32(b.fc2_b, C);
    } else if (sscanf(name_c, "layers.%d.downsample.%n", &si, &pos) == 1 && pos > 0 && si >= 0 &&
               si + 1 < static_cast<int>(m->stages.size())) {
      SwinStageW& st = m->stages[si];
      const std::string f = name.substr(pos);
      if (f == "reduction.weight") b16(st.red_w, 2 * st.C, 4 * st.C, 4 * st.C);
      else if (f == "norm.weight") f32(st.red_norm_w, 2 * st.C); else if (f == "norm.bias") f32(st.red_norm_b, 2 * st.C);
    }
  }
  if (!t.dst) {
    set_last_error("swin_set_param: unknown or inapplicable parameter '" + name + "'");
    return VSCB200_ERR_INVALID;
  }
  const int64_t expect = t.rows * t.cols;
  if (count != expect) {
    set_last_error("swin_set_param: '" + name + "' expects " + std::to_string(expect) + " elements, got " + std::to_string(count));
    return VSCB200_ERR_INVALID;
  }
  if (t.bf16) {
    int rc = cast_f32_bf16_padded(w, t.dst, t.rows, t.cols, t.ld, stream, lo_plane(m, t.dst, t.rows * t.ld));
    if (rc) return rc;
  } else {
    VSCB_CUDA_OK(cudaMemcpyAsync(t.dst, w, count * 4, cudaMemcpyDeviceToDevice, stream));
  }
  m->loaded[name] = true;
  m->finalized = false;
  return VSCB200_OK;
}

static int swin_forward_chunk(vscb200_swin* m, const float* frames, int n, float* out, cudaStream_t s) {
  const vscb200_swin_spec& sp = m->spec;
  int rc;
#define R(call) if ((rc = (call))) return rc
  R(finalize(m, s));
  const int res0 = sp.img / sp.patch;
  const float eps = sp.ln_eps;
  // plane strides (elements) of the operand buffers: fixed by the plan's capacity (stage 0 sizes); 0 in the bf16 mode
  const int64_t M0cap = static_cast<int64_t>(m->max_frames) * res0 * res0;
  const int64_t h_lo = m->exact ? M0cap * sp.embed : 0, qkv_lo = 3 * h_lo, u_lo = 4 * h_lo, pat_lo = m->exact ? M0cap * m->Kp : 0;
  auto lo = [&](void* hi, int64_t off) -> void* { return off ? static_cast<void*>(static_cast<uint16_t*>(hi) + off) : nullptr; };
  auto wlo = [&](void* w, int64_t rows, int64_t ld) -> const void* { return lo_plane(m, w, rows * ld); };
  {
    const int64_t M0 = static_cast<int64_t>(n) * res0 * res0;
    R(im2row(frames, m->patches, n, sp.img, sp.patch, m->Kp, s, pat_lo));
    R(gemm_bf16(m->patches, m->patch_w, m->patch_b, m->x, M0, sp.embed, m->Kp, m->Kp, m->Kp, sp.embed, VSCB200_EPI_F32, -1, s,
                nullptr, 0, false, 0, nullptr, lo(m->patches, pat_lo), wlo(m->patch_w, sp.embed, m->Kp), nullptr));
    R(layernorm(m->x, m->patch_norm_w, m->patch_norm_b, m->x, M0, sp.embed, eps, 0, s));
  }
  for (size_t i = 0; i < m->stages.size(); ++i) {
    const SwinStageW& st = m->stages[i];
    const int C = st.C, res = st.res, ws = st.ws;
    const int64_t M = static_cast<int64_t>(n) * res * res;
    const int nWx = res / ws, nW = nWx * nWx;
    // window attention kernels of the bf16 mode: S-resident kernel (swin_attention.cu) for 4 x 4 / 8 x 8 / 16 x 16 windows
    // (measured at 16 x 16: 25 ms vs 38 ms per 1024 SwinV2-B frames for the streaming kernel), streaming kernel
    // (attention_ws.cu) for larger windows or odd sides (12 x 12, 24 x 24), K-blocked kernel as its fallback
    const bool ws_attn = !m->exact && st.heads % 2 == 0 && ws != 4 && ws != 8 && ws != 16 && ws * ws >= 64 &&
                         attention_ws_supported(ws * ws, 32);
    const bool tc_attn = !m->exact && !ws_attn && (ws == 4 || ws == 8 || ws == 16);
    const bool kb_attn = !m->exact && !tc_attn && !ws_attn && st.heads % 2 == 0 && attention_kb_supported(ws * ws, 32);   // 24 x 24, 12 x 12
    for (size_t j = 0; j < st.blocks.size(); ++j) {
      const SwinBlockW& b = st.blocks[j];
      const int shift = (res > sp.window && (j & 1)) ? sp.window / 2 : 0;     // swinv2.py:223-226, 411
      // ---- (shifted) window attention branch.  Only the first block of a stage gathers its input rows itself: every
      //      later block finds them written, in its own window order, by the previous block's res-post-norm kernel.
      if (j == 0) R(window_gather_bf16(m->x, m->h, n, res, ws, shift, C, s, h_lo));
      R(gemm_bf16(m->h, b.qkv_w, b.qkv_bias, m->qkv, M, 3 * C, C, C, C, 3 * C, VSCB200_EPI_BF16, -1, s, nullptr, 0, false,
                  2 * C, b.qscale, lo(m->h, h_lo), wlo(b.qkv_w, 3 * C, C), lo(m->qkv, qkv_lo)));
      if (ws_attn) {
        R(attention_ws_swin(m->qkv, m->ao, static_cast<int64_t>(n) * nW, ws, st.heads, b.table, shift, nWx, nW, s));
      } else if (tc_attn) {
        R(swin_attention(m->qkv, m->ao, b.table, static_cast<int64_t>(n) * nW, nW, nWx, ws, shift, st.heads, s));
      } else if (kb_attn) {
        R(attention_kb(m->qkv, m->ao, static_cast<int64_t>(n) * nW, ws * ws, st.heads, 32, 1.0f, b.table, ws, shift, nWx, nW, s));
      } else {
        R(attention_fp32(m->qkv, qkv_lo, m->ao, h_lo, static_cast<int64_t>(n) * nW, ws * ws, st.heads, 32, 1.0f, b.table, ws, res,
                         shift, nWx, nW, s));
      }
      R(gemm_bf16(m->ao, b.proj_w, b.proj_b, m->y, M, C, C, C, C, C, VSCB200_EPI_F32, -1, s, nullptr, 0, false, 0, nullptr,
                  lo(m->ao, h_lo), wlo(b.proj_w, C, C), nullptr));
      R(ln_residual_scatter(m->y, b.norm1_w, b.norm1_b, m->x, n, res, ws, shift, C, eps, s, m->h, res, 0, h_lo));   // h: MLP input
      // ---- MLP branch (token order: identity map)
      R(gemm_bf16(m->h, b.fc1_w, b.fc1_b, m->u, M, 4 * C, C, C, C, 4 * C, VSCB200_EPI_BF16, VSCB200_ACT_GELU, s, nullptr, 0, false,
                  0, nullptr, lo(m->h, h_lo), wlo(b.fc1_w, 4 * C, C), lo(m->u, u_lo)));
      R(gemm_bf16(m->u, b.fc2_w, b.fc2_b, m->y, M, C, 4 * C, 4 * C, 4 * C, C, VSCB200_EPI_F32, -1, s, nullptr, 0, false, 0, nullptr,
                  lo(m->u, u_lo), wlo(b.fc2_w, C, 4 * C), nullptr));
      const bool more = j + 1 < st.blocks.size();
      const int next_shift = (res > sp.window && ((j + 1) & 1)) ? sp.window / 2 : 0;
      R(ln_residual_scatter(m->y, b.norm2_w, b.norm2_b, m->x, n, res, res, 0, C, eps, s, more ? m->h : nullptr, ws, next_shift,
                            h_lo));
    }
    if (i + 1 < m->stages.size()) {
      R(patch_merge_gather(m->x, m->h, n, res, C, s, h_lo));
      R(gemm_bf16(m->h, st.red_w, nullptr, m->y, M / 4, 2 * C, 4 * C, 4 * C, 4 * C, 2 * C, VSCB200_EPI_F32, -1, s, nullptr, 0, false,
                  0, nullptr, lo(m->h, h_lo), wlo(st.red_w, 2 * C, 4 * C), nullptr));
      R(layernorm(m->y, st.red_norm_w, st.red_norm_b, m->x, M / 4, 2 * C, eps, 0, s));
    }
  }
  const SwinStageW& last = m->stages.back();
  R(gem_head(m->x, m->norm_w, m->norm_b, m->head_w, m->head_b, out, n, last.res * last.res, last.C, sp.out_dim, eps, sp.gem_p,
             true, s));
#undef R
  return VSCB200_OK;
}

int vscb200_swin_forward(vscb200_swin* m, const float* frames, int64_t n, float* out, void* stream_v) {
  VSCB_REQUIRE(m && (n == 0 || (frames && out)), "swin_forward: null argument");
  VSCB_REQUIRE(n >= 0, "swin_forward: negative frame count");
  cudaStream_t s = static_cast<cudaStream_t>(stream_v);
  const int64_t in_per = 3LL * m->spec.img * m->spec.img;
  for (int64_t f0 = 0; f0 < n; f0 += m->max_frames) {
    const int nc = static_cast<int>(n - f0 < m->max_frames ? n - f0 : m->max_frames);
    int rc = swin_forward_chunk(m, frames + f0 * in_per, nc, out + f0 * m->spec.out_dim, s);
    if (rc) return rc;
  }
  return VSCB200_OK;
}

int vscb200_swin_forward_host(vscb200_swin* m, const float* frames_host, int64_t n, float* out_host) {
  // Same double-buffered pipeline as vscb200_vit_forward_host: H2D of chunk i+1 and D2H of chunk i-1 overlap the kernels
  // of chunk i; descriptors land in page-locked buffers and are handed to the caller one chunk later.
  VSCB_REQUIRE(m && (n == 0 || (frames_host && out_host)), "swin_forward_host: null argument");
  const int64_t in_per = 3LL * m->spec.img * m->spec.img, out_per = m->spec.out_dim;
  int64_t chunk = 0, last_f0 = 0;
  int last_nc = 0, last_b = 0;
  for (int64_t f0 = 0; f0 < n; f0 += m->max_frames, ++chunk) {
    const int b = static_cast<int>(chunk & 1);
    const int nc = static_cast<int>(n - f0 < m->max_frames ? n - f0 : m->max_frames);
    if (chunk >= 2) VSCB_CUDA_OK(cudaStreamWaitEvent(m->in_stream, m->ev_comp[b], 0));
    VSCB_CUDA_OK(cudaMemcpyAsync(m->frames_stage2[b], frames_host + f0 * in_per, static_cast<size_t>(nc) * in_per * 4,
                                 cudaMemcpyHostToDevice, m->in_stream));
    VSCB_CUDA_OK(cudaEventRecord(m->ev_in[b], m->in_stream));
    VSCB_CUDA_OK(cudaStreamWaitEvent(m->own_stream, m->ev_in[b], 0));
    if (chunk >= 2) VSCB_CUDA_OK(cudaStreamWaitEvent(m->own_stream, m->ev_out[b], 0));
    int rc = swin_forward_chunk(m, m->frames_stage2[b], nc, m->out_stage2[b], m->own_stream);
    if (rc) return rc;
    VSCB_CUDA_OK(cudaEventRecord(m->ev_comp[b], m->own_stream));
    VSCB_CUDA_OK(cudaStreamWaitEvent(m->out_stream, m->ev_comp[b], 0));
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
