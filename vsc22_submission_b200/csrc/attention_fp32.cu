// fp32 attention over segments of consecutive rows (a ViT frame of T tokens, or a Swin-V2 window of ws*ws tokens), any
// segment length, head_dim 32 or 64.  Two uses:
//   * the encoders' fp32-equivalent mode: q / k / v arrive as hi + lo bf16 planes (16 mantissa bits), scores, softmax
//     and P.V are computed in fp32 on the FMA pipe, the output leaves as hi + lo planes again;
//   * segment lengths the tcgen05 kernels do not cover in the bf16 mode (Swin windows that are not 4 / 8 / 16 wide:
//     the 24 x 24 windows of SwinV2-L@384, BASELINE configs[3]) -- single-plane input and output.
//
// One CTA = 128 query rows of one (segment, head); thread = query row: q and the output accumulator live in
// registers, keys / values stream through shared memory in blocks of 32 rows (converted to fp32 once), every thread of
// a warp reads the same K / V element (a broadcast LDS.128 per four FMAs), online softmax over 8-key chunks.
//
// Reference: nn.MultiheadAttention at D/train/train_vid_score/video/clip.py:45 (q scaled by d^-0.5) and
// WindowAttention.forward, swinv2.py:147-185 (cosine attention: q, k arrive normalised, q carries the logit scale; the
// relative-position bias is read from the per-head (2ws-1)^2 table, the shifted-window mask -- swinv2.py:232-255 --
// is recomputed from the token coordinates).
#include "host_util.h"
#include "kernels.h"
#include "ptx.cuh"

namespace vscb200 {

constexpr int kAfThreads = 128;
constexpr int kAfKeys = 32;

struct AttnFp32Params {
  const __nv_bfloat16* qkv;     // [M, 3C] rows = segments back to back; columns q | k | v, heads HD-wide contiguous
  int64_t qkv_lo_off;           // element offset of the lo plane (0: single plane)
  __nv_bfloat16* out;           // [M, C]
  int64_t out_lo_off;
  int C, heads, N;              // N: rows per segment
  int qtiles;                   // ceil(N / 128)
  float scale;                  // multiplies q.k
  const float* tables;          // Swin: [heads][(2ws-1)^2] relative-position bias, else nullptr
  int ws, res, shift, nWx, nW_per_frame;
};

__device__ __forceinline__ float bf16_bits_to_f32(uint32_t h) { return __uint_as_float(h << 16); }

// 8 consecutive elements (16 bytes per plane) of a hi (+ lo) bf16 row -> fp32
__device__ __forceinline__ void load8(const __nv_bfloat16* p, int64_t lo_off, float (&f)[8]) {
  const uint4 h = *reinterpret_cast<const uint4*>(p);
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(hw[i] << 16);
    f[2 * i + 1] = __uint_as_float(hw[i] & 0xFFFF0000u);
  }
  if (lo_off) {
    const uint4 l = *reinterpret_cast<const uint4*>(p + lo_off);
    const uint32_t lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] += __uint_as_float(lw[i] << 16);
      f[2 * i + 1] += __uint_as_float(lw[i] & 0xFFFF0000u);
    }
  }
}

// region index of a coordinate of the SHIFTED token map (swinv2.py:238-243: slices [0,-ws), [-ws,-shift), [-shift, end))
__device__ __forceinline__ int shift_region(int c, int res, int ws, int shift) {
  return c < res - ws ? 0 : (c < res - shift ? 1 : 2);
}

template <int HD>
__global__ void __launch_bounds__(kAfThreads)
attention_fp32_kernel(AttnFp32Params p) {
  extern __shared__ __align__(16) float af_smem[];
  float* sK = af_smem;                          // [kAfKeys][HD]
  float* sV = af_smem + kAfKeys * HD;           // [kAfKeys][HD]
  float* sTab = sV + kAfKeys * HD;              // Swin: (2ws-1)^2 bias entries of this head
  const int tid = threadIdx.x;
  int64_t b = blockIdx.x;
  const int qt = static_cast<int>(b % p.qtiles); b /= p.qtiles;
  const int head = static_cast<int>(b % p.heads);
  const int64_t seg = b / p.heads;
  const int64_t row0 = seg * p.N;
  const int ld = 3 * p.C;
  const int i_row = qt * kAfThreads + tid;                    // this thread's query row inside the segment
  const bool row_ok = i_row < p.N;

  // ---- Swin geometry of this thread's query token and of the window
  const bool swin = p.tables != nullptr;
  const int TS = 2 * p.ws - 1;
  int base_i = 0, reg_i = 0, wy0 = 0, wx0 = 0;
  if (swin) {
    for (int e = tid; e < TS * TS; e += kAfThreads) sTab[e] = p.tables[static_cast<int64_t>(head) * TS * TS + e];
    const int ii = row_ok ? i_row : 0;
    const int yi = ii / p.ws, xi = ii % p.ws;
    base_i = (yi + p.ws - 1) * TS + xi + p.ws - 1;
    const int wf = static_cast<int>(seg % p.nW_per_frame);
    wy0 = (wf / p.nWx) * p.ws; wx0 = (wf % p.nWx) * p.ws;
    if (p.shift > 0) reg_i = shift_region(wy0 + yi, p.res, p.ws, p.shift) * 3 + shift_region(wx0 + xi, p.res, p.ws, p.shift);
  }

  // ---- q row -> registers (fp32, pre-multiplied by the score scale)
  float q[HD], o[HD];
#pragma unroll
  for (int c = 0; c < HD; ++c) { q[c] = 0.f; o[c] = 0.f; }
  if (row_ok) {
    const __nv_bfloat16* qp = p.qkv + (row0 + i_row) * ld + head * HD;
#pragma unroll
    for (int c8 = 0; c8 < HD / 8; ++c8) {
      float f[8];
      load8(qp + c8 * 8, p.qkv_lo_off, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) q[c8 * 8 + i] = f[i] * p.scale;
    }
  }
  float m_run = -INFINITY, l_run = 0.f;

  for (int k0 = 0; k0 < p.N; k0 += kAfKeys) {
    __syncthreads();                                          // the previous block has been consumed (and sTab is written)
    // ---- stage 32 keys and values as fp32: 32 rows x HD/8 chunks x {K, V}
    for (int e = tid; e < kAfKeys * (HD / 8) * 2; e += kAfThreads) {
      const int which = e / (kAfKeys * (HD / 8));
      const int r = (e / (HD / 8)) % kAfKeys, c8 = e % (HD / 8);
      float f[8];
      if (k0 + r < p.N) {
        load8(p.qkv + (row0 + k0 + r) * ld + (which + 1) * p.C + head * HD + c8 * 8, p.qkv_lo_off, f);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = 0.f;
      }
      float* dst = (which ? sV : sK) + r * HD + c8 * 8;
      *reinterpret_cast<float4*>(dst) = make_float4(f[0], f[1], f[2], f[3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(f[4], f[5], f[6], f[7]);
    }
    __syncthreads();
#pragma unroll 1
    for (int c0 = 0; c0 < kAfKeys; c0 += 8) {
      if (k0 + c0 >= p.N) break;                              // uniform
      float s[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4* kr = reinterpret_cast<const float4*>(sK + (c0 + j) * HD);
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int c4 = 0; c4 < HD / 4; c4 += 2) {
          const float4 ka = kr[c4], kb = kr[c4 + 1];
          a0 = fmaf(q[4 * c4], ka.x, a0); a0 = fmaf(q[4 * c4 + 1], ka.y, a0);
          a0 = fmaf(q[4 * c4 + 2], ka.z, a0); a0 = fmaf(q[4 * c4 + 3], ka.w, a0);
          a1 = fmaf(q[4 * c4 + 4], kb.x, a1); a1 = fmaf(q[4 * c4 + 5], kb.y, a1);
          a1 = fmaf(q[4 * c4 + 6], kb.z, a1); a1 = fmaf(q[4 * c4 + 7], kb.w, a1);
        }
        float sj = a0 + a1;
        const int kj = k0 + c0 + j;                           // uniform
        if (swin) {
          const int yj = kj / p.ws, xj = kj % p.ws;
          if (kj < p.N) sj += sTab[base_i - (yj * TS + xj)];
          if (p.shift > 0) {
            const int reg_j = shift_region(wy0 + yj, p.res, p.ws, p.shift) * 3 + shift_region(wx0 + xj, p.res, p.ws, p.shift);
            if (reg_j != reg_i) sj += -100.0f;
          }
        }
        s[j] = kj < p.N ? sj : -INFINITY;
      }
      float mx = s[0];
#pragma unroll
      for (int j = 1; j < 8; ++j) mx = fmaxf(mx, s[j]);
      const float m_new = fmaxf(m_run, mx);                   // finite: the first key of a chunk is always valid
      const float alpha = __expf(m_run - m_new);
      float psum = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j] = expf(s[j] - m_new); psum += s[j]; }
      l_run = fmaf(l_run, alpha, psum);
      m_run = m_new;
#pragma unroll
      for (int c = 0; c < HD; ++c) o[c] *= alpha;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4* vr = reinterpret_cast<const float4*>(sV + (c0 + j) * HD);
#pragma unroll
        for (int c4 = 0; c4 < HD / 4; ++c4) {
          const float4 vv = vr[c4];
          o[4 * c4] = fmaf(s[j], vv.x, o[4 * c4]); o[4 * c4 + 1] = fmaf(s[j], vv.y, o[4 * c4 + 1]);
          o[4 * c4 + 2] = fmaf(s[j], vv.z, o[4 * c4 + 2]); o[4 * c4 + 3] = fmaf(s[j], vv.w, o[4 * c4 + 3]);
        }
      }
    }
  }
  if (!row_ok) return;
  const float inv_l = 1.0f / l_run;
  __nv_bfloat16* op = p.out + (row0 + i_row) * p.C + head * HD;
#pragma unroll
  for (int c8 = 0; c8 < HD / 8; ++c8) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float a = o[c8 * 8 + 2 * i] * inv_l, c = o[c8 * 8 + 2 * i + 1] * inv_l;
      hi[i] = pack_bf16x2(a, c);
      lo[i] = pack_bf16x2(a - __uint_as_float(hi[i] << 16), c - __uint_as_float(hi[i] & 0xFFFF0000u));
    }
    *reinterpret_cast<uint4*>(op + c8 * 8) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (p.out_lo_off) *reinterpret_cast<uint4*>(op + p.out_lo_off + c8 * 8) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// qkv / out: see AttnFp32Params.  tables == nullptr: plain scaled-dot-product attention (ViT).
int attention_fp32(const void* qkv, int64_t qkv_lo_off, void* out, int64_t out_lo_off, int64_t n_segs, int N, int heads,
                   int head_dim, float scale, const float* tables, int ws, int res, int shift, int nWx, int nW_per_frame,
                   cudaStream_t stream) {
  VSCB_REQUIRE(head_dim == 32 || head_dim == 64, "attention_fp32: head_dim must be 32 or 64");
  VSCB_REQUIRE(N > 0 && heads > 0, "attention_fp32: empty segment");
  if (n_segs == 0) return VSCB200_OK;
  AttnFp32Params p;
  p.qkv = reinterpret_cast<const __nv_bfloat16*>(qkv); p.qkv_lo_off = qkv_lo_off;
  p.out = reinterpret_cast<__nv_bfloat16*>(out); p.out_lo_off = out_lo_off;
  p.C = heads * head_dim; p.heads = heads; p.N = N; p.qtiles = (N + kAfThreads - 1) / kAfThreads; p.scale = scale;
  p.tables = tables; p.ws = ws > 0 ? ws : 1; p.res = res; p.shift = shift; p.nWx = nWx > 0 ? nWx : 1;
  p.nW_per_frame = nW_per_frame > 0 ? nW_per_frame : 1;
  const int64_t blocks = n_segs * heads * p.qtiles;
  VSCB_REQUIRE(blocks < (1ll << 31), "attention_fp32: problem too large");
  const int TS = 2 * p.ws - 1;
  const size_t smem = static_cast<size_t>(2 * kAfKeys * head_dim + (tables ? TS * TS : 0)) * sizeof(float);
  VSCB_REQUIRE(smem <= 200 * 1024, "attention_fp32: window too large");
  ProfScope prof(kProfAttention, stream, 4.0 * static_cast<double>(n_segs) * heads * N * N * head_dim);
  if (head_dim == 64) {
    VSCB_CUDA_OK(cudaFuncSetAttribute(attention_fp32_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attention_fp32_kernel<64><<<static_cast<unsigned>(blocks), kAfThreads, smem, stream>>>(p);
  } else {
    VSCB_CUDA_OK(cudaFuncSetAttribute(attention_fp32_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attention_fp32_kernel<32><<<static_cast<unsigned>(blocks), kAfThreads, smem, stream>>>(p);
  }
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

}  // namespace vscb200

extern "C" int vscb200_attention_fp32(const void* qkv_hi, const void* qkv_lo, void* out_hi, void* out_lo, int64_t n_segs, int N,
                                      int heads, int head_dim, void* stream) {
  if (!qkv_hi || !out_hi || (qkv_lo != nullptr) != (out_lo != nullptr)) {
    vscb200::set_last_error("vscb200_attention_fp32: null argument (lo planes come in pairs)");
    return VSCB200_ERR_INVALID;
  }
  const int64_t qoff = qkv_lo ? static_cast<const __nv_bfloat16*>(qkv_lo) - static_cast<const __nv_bfloat16*>(qkv_hi) : 0;
  const int64_t ooff = out_lo ? static_cast<__nv_bfloat16*>(out_lo) - static_cast<__nv_bfloat16*>(out_hi) : 0;
  return vscb200::attention_fp32(qkv_hi, qoff, out_hi, ooff, n_segs, N, heads, head_dim, 1.0f / sqrtf(static_cast<float>(head_dim)),
                                 nullptr, 0, 0, 0, 0, 0, static_cast<cudaStream_t>(stream));
}
