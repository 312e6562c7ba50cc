// Memory-bound kernels of the ViT encoder: LayerNorm (fp32 residual stream -> bf16 GEMM operand),
// im2row patch gather, class-token rows, GeM pooling + projection tail, fp32->bf16 weight packing.
// Reference ops: clip.py:11-16,143-152,158 (LayerNorm / patch conv / cls+pos), backbones/vit.py:56-58
// (gem), sscd.py:30-40,86 (GlobalGeMPool2d + Linear).
#include "host_util.h"
#include "ptx.cuh"

namespace vscb200 {

// ------------------------------------------------------------------ LayerNorm: one warp per row
constexpr int kLnMaxVec = 8;   // float4 per lane -> width <= 1024

template <bool kOutBf16>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 void* __restrict__ y, int64_t rows, int width, float eps, int reverse) {
  const int lane = threadIdx.x & 31;
  const unsigned bid = reverse ? gridDim.x - 1 - blockIdx.x : blockIdx.x;
  const int64_t row = static_cast<int64_t>(bid) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = width >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + row * width);
  float4 v[kLnMaxVec];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
      v[i] = xr[c];
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / width;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
      const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
      sq += (a * a + b * b) + (cc * cc + d * d);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / width + eps);
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
      const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + c);
      const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + c);
      float4 r;
      r.x = (v[i].x - mean) * rstd * gm.x + bt.x;
      r.y = (v[i].y - mean) * rstd * gm.y + bt.y;
      r.z = (v[i].z - mean) * rstd * gm.z + bt.z;
      r.w = (v[i].w - mean) * rstd * gm.w + bt.w;
      if (kOutBf16) {
        uint2 o2 = make_uint2(pack_bf16x2(r.x, r.y), pack_bf16x2(r.z, r.w));
        reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(y) + row * width)[c] = o2;
      } else {
        reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + row * width)[c] = r;
      }
    }
  }
}

int layernorm(const float* x, const float* gamma, const float* beta, void* y, int64_t rows, int width, float eps,
              int out_bf16, cudaStream_t stream, bool reverse) {
  VSCB_REQUIRE(width % 4 == 0 && width <= 128 * kLnMaxVec, "layernorm: width must be a multiple of 4 and <= 1024");
  if (rows == 0) return VSCB200_OK;
  const int rows_per_block = 8;
  const unsigned grid = static_cast<unsigned>((rows + rows_per_block - 1) / rows_per_block);
  ProfScope prof(kProfLayerNorm, stream, static_cast<double>(rows) * width * (out_bf16 ? 6 : 8));
  if (out_bf16)
    layernorm_kernel<true><<<grid, 256, 0, stream>>>(x, gamma, beta, y, rows, width, eps, reverse ? 1 : 0);
  else
    layernorm_kernel<false><<<grid, 256, 0, stream>>>(x, gamma, beta, y, rows, width, eps, reverse ? 1 : 0);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

// ------------------------------------------------------------------ fp32 -> bf16 (weights, with row padding)
__global__ void cast_pad_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int64_t rows, int cols,
                                int ld_out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * ld_out) return;
  const int64_t r = i / ld_out;
  const int c = static_cast<int>(i % ld_out);
  y[i] = __float2bfloat16_rn(c < cols ? x[r * cols + c] : 0.f);
}

int cast_f32_bf16_padded(const float* x, void* y, int64_t rows, int cols, int ld_out, cudaStream_t stream) {
  const int64_t n = rows * ld_out;
  if (n == 0) return VSCB200_OK;
  cast_pad_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
      x, reinterpret_cast<__nv_bfloat16*>(y), rows, cols, ld_out);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

// ------------------------------------------------------------------ im2row for the stride==kernel patch conv
// frames [n,3,H,H] fp32 NCHW -> patches [n*P, Kp] bf16, column k = c*p*p + i*p + j (the flattening of
// conv1.weight [W,3,p,p], clip.py:105), zero padded to Kp.
__global__ void im2row_kernel(const float* __restrict__ frames, __nv_bfloat16* __restrict__ patches, int64_t n,
                              int img, int patch, int Kp) {
  const int grid = img / patch;
  const int P = grid * grid;
  const int kpairs = Kp >> 1;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n * P * kpairs) return;
  const int kp = static_cast<int>(idx % kpairs);
  const int64_t m = idx / kpairs;
  const int pidx = static_cast<int>(m % P);
  const int64_t f = m / P;
  const int k = kp * 2;
  float2 v = make_float2(0.f, 0.f);
  if (k < 3 * patch * patch) {
    const int c = k / (patch * patch), rem = k % (patch * patch);
    const int i = rem / patch, j = rem % patch;
    const int py = pidx / grid, px = pidx % grid;
    const float* src = frames + ((f * 3 + c) * img + (py * patch + i)) * static_cast<int64_t>(img) + px * patch + j;
    v = *reinterpret_cast<const float2*>(src);
  }
  reinterpret_cast<uint32_t*>(patches)[idx] = pack_bf16x2(v.x, v.y);
}

int im2row(const float* frames, void* patches, int64_t n, int img, int patch, int Kp, cudaStream_t stream) {
  VSCB_REQUIRE(patch % 2 == 0 && img % patch == 0 && Kp % 2 == 0, "im2row: patch must be even and divide img");
  const int P = (img / patch) * (img / patch);
  const int64_t total = n * P * (Kp / 2);
  if (total == 0) return VSCB200_OK;
  ProfScope prof(kProfVitOther, stream, static_cast<double>(total) * 2 * 6);
  im2row_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      frames, reinterpret_cast<__nv_bfloat16*>(patches), n, img, patch, Kp);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

// ------------------------------------------------------------------ class-token rows: x[f*T + 0] = cls + pos[0]
__global__ void cls_rows_kernel(const float* __restrict__ cls, const float* __restrict__ pos, float* __restrict__ x,
                                int64_t n, int T, int W) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n * W) return;
  const int c = static_cast<int>(i % W);
  const int64_t f = i / W;
  x[f * T * W + c] = cls[c] + pos[c];
}

int cls_rows(const float* cls, const float* pos, float* x, int64_t n, int T, int W, cudaStream_t stream) {
  cls_rows_kernel<<<static_cast<unsigned>((n * W + 255) / 256), 256, 0, stream>>>(cls, pos, x, n, T, W);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

// ------------------------------------------------------------------ GeM pooling + Linear tail, one CTA per frame
// y: [n, T, C] fp32.  kLN: apply LayerNorm(gamma, beta, eps) to each token row first (ln_post fused).
// g[c] = (mean_t clamp(y[t,c], 1e-6)^p)^(1/p) ; out[o] = head_b[o] + sum_c head_w[o,c] * g[c].
constexpr int kTailThreads = 256;

template <bool kLN>
__global__ void __launch_bounds__(kTailThreads)
gem_head_kernel(const float* __restrict__ y, const float* __restrict__ gamma, const float* __restrict__ beta,
                const float* __restrict__ head_w, const float* __restrict__ head_b, float* __restrict__ out, int T,
                int C, int out_dim, float eps, float p) {
  extern __shared__ float tail_smem[];   // [C] pooled sums / g
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarp = kTailThreads >> 5;
  const int64_t f = blockIdx.x;
  const float* yf = y + f * T * C;
  for (int c = tid; c < C; c += kTailThreads) tail_smem[c] = 0.f;
  __syncthreads();
  const bool cube = (p == 3.0f);
  if (kLN) {
    // one warp per token row; lane owns columns lane*4 + 128*i
    const int nvec = C >> 2;
    float4 acc[kLnMaxVec];
#pragma unroll
    for (int i = 0; i < kLnMaxVec; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = warp; t < T; t += nwarp) {
      const float4* xr = reinterpret_cast<const float4*>(yf + static_cast<int64_t>(t) * C);
      float4 v[kLnMaxVec];
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < kLnMaxVec; ++i) {
        const int c = lane + 32 * i;
        if (c < nvec) { v[i] = xr[c]; sum += (v[i].x + v[i].y) + (v[i].z + v[i].w); }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float mean = sum / C;
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < kLnMaxVec; ++i) {
        const int c = lane + 32 * i;
        if (c < nvec) {
          const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
          sq += (a * a + b * b) + (cc * cc + d * d);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      const float rstd = rsqrtf(sq / C + eps);
#pragma unroll
      for (int i = 0; i < kLnMaxVec; ++i) {
        const int c = lane + 32 * i;
        if (c < nvec) {
          const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + c);
          const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + c);
          float e[4] = {(v[i].x - mean) * rstd * gm.x + bt.x, (v[i].y - mean) * rstd * gm.y + bt.y,
                        (v[i].z - mean) * rstd * gm.z + bt.z, (v[i].w - mean) * rstd * gm.w + bt.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float z = fmaxf(e[q], 1e-6f);
            e[q] = cube ? z * z * z : powf(z, p);
          }
          acc[i].x += e[0]; acc[i].y += e[1]; acc[i].z += e[2]; acc[i].w += e[3];
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kLnMaxVec; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        atomicAdd(&tail_smem[4 * c + 0], acc[i].x);
        atomicAdd(&tail_smem[4 * c + 1], acc[i].y);
        atomicAdd(&tail_smem[4 * c + 2], acc[i].z);
        atomicAdd(&tail_smem[4 * c + 3], acc[i].w);
      }
    }
  } else {
    // thread per column, coalesced across the row
    for (int c = tid; c < C; c += kTailThreads) {
      float a = 0.f;
      for (int t = 0; t < T; ++t) {
        const float z = fmaxf(yf[static_cast<int64_t>(t) * C + c], 1e-6f);
        a += cube ? z * z * z : powf(z, p);
      }
      tail_smem[c] = a;
    }
  }
  __syncthreads();
  for (int c = tid; c < C; c += kTailThreads) {
    const float m = tail_smem[c] / T;
    tail_smem[c] = cube ? cbrtf(m) : powf(m, 1.0f / p);
  }
  __syncthreads();
  for (int o = warp; o < out_dim; o += nwarp) {
    const float* wr = head_w + static_cast<int64_t>(o) * C;
    float a = 0.f;
    for (int c = lane; c < C; c += 32) a += wr[c] * tail_smem[c];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) a += __shfl_xor_sync(0xffffffffu, a, s);
    if (lane == 0) out[f * out_dim + o] = a + head_b[o];
  }
}

int gem_head(const float* y, const float* gamma, const float* beta, const float* head_w, const float* head_b,
             float* out, int64_t n, int T, int C, int out_dim, float eps, float p, bool fuse_ln,
             cudaStream_t stream) {
  if (n == 0) return VSCB200_OK;
  VSCB_REQUIRE(!fuse_ln || (C % 4 == 0 && C <= 128 * kLnMaxVec), "gem_head: fused LN needs width % 4 == 0 and <= 1024");
  const size_t smem = static_cast<size_t>(C) * sizeof(float);
  ProfScope prof(kProfVitOther, stream, static_cast<double>(n) * T * C * 4);
  if (fuse_ln)
    gem_head_kernel<true><<<static_cast<unsigned>(n), kTailThreads, smem, stream>>>(y, gamma, beta, head_w, head_b,
                                                                                   out, T, C, out_dim, eps, p);
  else
    gem_head_kernel<false><<<static_cast<unsigned>(n), kTailThreads, smem, stream>>>(y, gamma, beta, head_w, head_b,
                                                                                    out, T, C, out_dim, eps, p);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

}  // namespace vscb200

extern "C" {
int vscb200_layernorm(const float* x, const float* gamma, const float* beta, void* y, int64_t rows, int width,
                      float eps, int out_bf16, void* stream) {
  return vscb200::layernorm(x, gamma, beta, y, rows, width, eps, out_bf16, static_cast<cudaStream_t>(stream), false);
}
int vscb200_cast_f32_bf16(const float* x, void* y_bf16, int64_t count, void* stream) {
  return vscb200::cast_f32_bf16_padded(x, y_bf16, count, 1, 1, static_cast<cudaStream_t>(stream));
}
}
