// Swin-V2 glue kernels (memory-bound; one warp per token row, 128-bit accesses):
//   window_gather_bf16   x fp32 [n, res*res, C] -> bf16 rows in (shifted) window order  (torch.roll + window_partition,
//                        swinv2.py:280-288; the reference materialises both as copies)
//   ln_residual_scatter  x[token] += LayerNorm(y[row])  with the inverse mapping        (window_reverse + roll back +
//                        `shortcut + norm1(x)` / `x + norm2(mlp(x))`, swinv2.py:293-306: res-post-norm)
//   patch_merge_gather   2x2 neighbourhood concat -> bf16 [n*(res/2)^2, 4C]              (PatchMerging.forward :353-365)
//   cpb_table            16*sigmoid(cpb_mlp(log-spaced relative coordinates)) per head   (swinv2.py:99-113,165-170; the
//                        reference re-runs this MLP every forward, here once per weight load)
#include "host_util.h"
#include "kernels.h"
#include "ptx.cuh"

namespace vscb200 {

struct WinMap {
  int res, ws, shift, nWx;      // token map side, window side, cyclic shift, windows per row
};

// window-order row -> token index inside the frame
__device__ __forceinline__ int win_row_to_token(const WinMap& m, int r_in_frame) {
  const int N = m.ws * m.ws;
  const int win = r_in_frame / N, pos = r_in_frame % N;
  const int wy = win / m.nWx, wx = win % m.nWx;
  int y = wy * m.ws + pos / m.ws + m.shift, x = wx * m.ws + pos % m.ws + m.shift;   // roll by -shift: shifted[y] = x[y + shift]
  if (y >= m.res) y -= m.res;
  if (x >= m.res) x -= m.res;
  return y * m.res + x;
}

// token index inside the frame -> its row in (shifted) window order (inverse of win_row_to_token)
__device__ __forceinline__ int token_to_win_row(const WinMap& m, int tok) {
  int y = tok / m.res - m.shift, x = tok % m.res - m.shift;
  if (y < 0) y += m.res;
  if (x < 0) x += m.res;
  const int N = m.ws * m.ws;
  return ((y / m.ws) * m.nWx + x / m.ws) * N + (y % m.ws) * m.ws + x % m.ws;
}

__global__ void __launch_bounds__(256)
window_gather_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ h, int64_t rows, int C, WinMap m,
                          int64_t lo_off) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int L = m.res * m.res;
  const int64_t frame = row / L;
  const int tok = win_row_to_token(m, static_cast<int>(row % L));
  const float4* src = reinterpret_cast<const float4*>(x + (frame * L + tok) * C);
  uint2* dst = reinterpret_cast<uint2*>(h + row * C);
  uint2* dst_lo = reinterpret_cast<uint2*>(h + lo_off + row * C);
  for (int c = lane; c < (C >> 2); c += 32) {
    const float4 v = src[c];
    const uint2 hi = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    dst[c] = hi;
    if (lo_off) dst_lo[c] = make_uint2(pack_bf16x2_lo(v.x, v.y, hi.x), pack_bf16x2_lo(v.z, v.w, hi.y));
  }
}

int window_gather_bf16(const float* x, void* h, int64_t n, int res, int ws, int shift, int C, cudaStream_t stream,
                       int64_t lo_off) {
  VSCB_REQUIRE(C % 4 == 0 && res % ws == 0, "window_gather: C % 4 and res % ws must be 0");
  const int64_t rows = n * res * res;
  if (rows == 0) return VSCB200_OK;
  WinMap m{res, ws, shift, res / ws};
  ProfScope prof(kProfVitOther, stream, static_cast<double>(rows) * C * 6);
  window_gather_bf16_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(x, reinterpret_cast<__nv_bfloat16*>(h),
                                                                                       rows, C, m, lo_off);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

constexpr int kSwLnMaxVec = 8;    // float4 per lane -> C <= 1024 (kSwLnWideVec: <= 2048, SwinV2-L's 1536-wide last stage)
constexpr int kSwLnWideVec = 16;

template <int kVec>
__global__ void __launch_bounds__(256)
ln_residual_scatter_kernel(const float* __restrict__ y, const float* __restrict__ gamma, const float* __restrict__ beta,
                           float* __restrict__ x, int64_t rows, int C, float eps, WinMap m, __nv_bfloat16* __restrict__ h_next,
                           WinMap m_next, int64_t lo_off) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = C >> 2;
  const float4* yr = reinterpret_cast<const float4*>(y + row * C);
  float4 v[kVec];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
      v[i] = yr[c];
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / C;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
      const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
      sq += (a * a + b * b) + (cc * cc + d * d);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / C + eps);
  const int L = m.res * m.res;
  const int64_t frame = row / L;
  const int tok = win_row_to_token(m, static_cast<int>(row % L));
  float4* xr = reinterpret_cast<float4*>(x + (frame * L + tok) * C);
  // the updated row is also the next GEMM's A operand: written as bf16 straight into that consumer's row order
  // (token order for the MLP, the next block's shifted-window order for its QKV projection) -- no gather/cast pass
  uint2* hr = h_next ? reinterpret_cast<uint2*>(h_next + (frame * L + token_to_win_row(m_next, tok)) * C) : nullptr;
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
      const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + c);
      const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + c);
      float4 r = xr[c];
      r.x += (v[i].x - mean) * rstd * gm.x + bt.x;
      r.y += (v[i].y - mean) * rstd * gm.y + bt.y;
      r.z += (v[i].z - mean) * rstd * gm.z + bt.z;
      r.w += (v[i].w - mean) * rstd * gm.w + bt.w;
      xr[c] = r;
      if (hr) {
        const uint2 hi = make_uint2(pack_bf16x2(r.x, r.y), pack_bf16x2(r.z, r.w));
        hr[c] = hi;
        if (lo_off) (hr + (lo_off >> 2))[c] = make_uint2(pack_bf16x2_lo(r.x, r.y, hi.x), pack_bf16x2_lo(r.z, r.w, hi.y));
      }
    }
  }
}

int ln_residual_scatter(const float* y, const float* gamma, const float* beta, float* x, int64_t n, int res, int ws, int shift,
                        int C, float eps, cudaStream_t stream, void* h_next, int ws_next, int shift_next, int64_t lo_off) {
  VSCB_REQUIRE(C % 4 == 0 && C <= 128 * kSwLnWideVec && res % ws == 0, "ln_residual_scatter: C must be a multiple of 4, <= 2048");
  VSCB_REQUIRE(lo_off % 4 == 0, "ln_residual_scatter: lo plane offset must be a multiple of 4 elements");
  const int64_t rows = n * res * res;
  if (rows == 0) return VSCB200_OK;
  WinMap m{res, ws, shift, res / ws};
  ProfScope prof(kProfLayerNorm, stream, static_cast<double>(rows) * C * 12);
  WinMap mn{res, ws_next > 0 ? ws_next : res, shift_next, ws_next > 0 ? res / ws_next : 1};
  if (C <= 128 * kSwLnMaxVec)
    ln_residual_scatter_kernel<kSwLnMaxVec><<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(
        y, gamma, beta, x, rows, C, eps, m, reinterpret_cast<__nv_bfloat16*>(h_next), mn, lo_off);
  else
    ln_residual_scatter_kernel<kSwLnWideVec><<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(
        y, gamma, beta, x, rows, C, eps, m, reinterpret_cast<__nv_bfloat16*>(h_next), mn, lo_off);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

// out row (frame, y2, x2) = [x(2y2, 2x2) | x(2y2+1, 2x2) | x(2y2, 2x2+1) | x(2y2+1, 2x2+1)]
__global__ void __launch_bounds__(256)
patch_merge_gather_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int64_t rows, int res, int C,
                          int64_t lo_off) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int r2 = res >> 1;
  const int64_t frame = row / (r2 * r2);
  const int rem = static_cast<int>(row % (r2 * r2));
  const int y2 = rem / r2, x2 = rem % r2;
  uint2* dst = reinterpret_cast<uint2*>(out + row * 4 * C);
#pragma unroll
  for (int part = 0; part < 4; ++part) {
    const int yy = 2 * y2 + (part & 1), xx = 2 * x2 + (part >> 1);
    const float4* src = reinterpret_cast<const float4*>(x + ((frame * res + yy) * res + xx) * C);
    for (int c = lane; c < (C >> 2); c += 32) {
      const float4 v = src[c];
      const uint2 hi = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
      dst[part * (C >> 2) + c] = hi;
      if (lo_off) (dst + (lo_off >> 2))[part * (C >> 2) + c] = make_uint2(pack_bf16x2_lo(v.x, v.y, hi.x), pack_bf16x2_lo(v.z, v.w, hi.y));
    }
  }
}

int patch_merge_gather(const float* x, void* out, int64_t n, int res, int C, cudaStream_t stream, int64_t lo_off) {
  VSCB_REQUIRE(C % 4 == 0 && res % 2 == 0, "patch_merge: C % 4 and even resolution required");
  const int64_t rows = n * (res / 2) * (res / 2);
  if (rows == 0) return VSCB200_OK;
  ProfScope prof(kProfVitOther, stream, static_cast<double>(rows) * 4 * C * 6);
  patch_merge_gather_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(x, reinterpret_cast<__nv_bfloat16*>(out),
                                                                                      rows, res, C, lo_off);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

// table[h][e] for e = dy*(2ws-1) + dx over relative offsets (dy, dx) in [-(ws-1), ws-1]^2
__global__ void cpb_table_kernel(const float* __restrict__ w0, const float* __restrict__ b0, const float* __restrict__ w2,
                                 float* __restrict__ table, int ws, int pretrained_ws, int heads) {
  const int ts = 2 * ws - 1;
  const int e = blockIdx.x;                    // one block per table entry
  const float denom = static_cast<float>(pretrained_ws > 0 ? pretrained_ws - 1 : ws - 1);
  auto coord = [&](int d) {
    const float c = static_cast<float>(d - (ws - 1)) / denom * 8.0f;
    const float s = c > 0.f ? 1.f : (c < 0.f ? -1.f : 0.f);
    return s * log2f(fabsf(c) + 1.0f) / 3.0f;    // / log2(8)
  };
  const float cy = coord(e / ts), cx = coord(e % ts);
  extern __shared__ float hid[];               // [512]
  for (int j = threadIdx.x; j < 512; j += blockDim.x) hid[j] = fmaxf(fmaf(w0[2 * j], cy, fmaf(w0[2 * j + 1], cx, b0[j])), 0.f);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int h = warp; h < heads; h += blockDim.x >> 5) {
    float acc = 0.f;
    for (int j = lane; j < 512; j += 32) acc = fmaf(hid[j], w2[h * 512 + j], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) table[h * ts * ts + e] = 16.0f / (1.0f + expf(-acc));
  }
}

int cpb_table(const float* w0, const float* b0, const float* w2, float* table, int ws, int pretrained_ws, int heads,
              cudaStream_t stream) {
  const int ts = 2 * ws - 1;
  cpb_table_kernel<<<ts * ts, 128, 512 * sizeof(float), stream>>>(w0, b0, w2, table, ws, pretrained_ws, heads);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

// qscale[h] = exp(min(logit_scale[h], log(100)))  (swinv2.py:162);  qkv_bias = (q_bias, 0, v_bias)  (:155-156)
__global__ void swin_prep_kernel(const float* __restrict__ logit_scale, const float* __restrict__ q_bias,
                                 const float* __restrict__ v_bias, float* __restrict__ qscale, float* __restrict__ qkv_bias,
                                 int heads, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < heads) qscale[i] = expf(fminf(logit_scale[i], 4.605170185988092f));
  if (i < C) {
    qkv_bias[i] = q_bias[i];
    qkv_bias[C + i] = 0.f;
    qkv_bias[2 * C + i] = v_bias[i];
  }
}

int swin_prep(const float* logit_scale, const float* q_bias, const float* v_bias, float* qscale, float* qkv_bias, int heads,
              int C, cudaStream_t stream) {
  swin_prep_kernel<<<(C + 255) / 256, 256, 0, stream>>>(logit_scale, q_bias, v_bias, qscale, qkv_bias, heads, C);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

}  // namespace vscb200
