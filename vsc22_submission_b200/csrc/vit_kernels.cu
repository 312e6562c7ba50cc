// Memory-bound kernels of the ViT encoder: LayerNorm (fp32 residual stream -> bf16 GEMM operand),
// im2row patch gather, class-token rows, GeM pooling + projection tail, fp32->bf16 weight packing.
// Reference ops: clip.py:11-16,143-152,158 (LayerNorm / patch conv / cls+pos), backbones/vit.py:56-58
// (gem), sscd.py:30-40,86 (GlobalGeMPool2d + Linear).
#include "host_util.h"
#include "ptx.cuh"

namespace vscb200 {

// ------------------------------------------------------------------ LayerNorm: one warp per row
constexpr int kLnMaxVec = 8;   // float4 per lane -> width <= 1024 (kLnWideVec: <= 2048, SwinV2-L's 1536-wide last stage)
constexpr int kLnWideVec = 16;

template <bool kOutBf16, int kVec>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 void* __restrict__ y, int64_t rows, int width, float eps, int reverse, int64_t lo_off) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const unsigned bid = reverse ? gridDim.x - 1 - blockIdx.x : blockIdx.x;
  const int64_t row = static_cast<int64_t>(bid) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = width >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + row * width);
  float4 v[kVec];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
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
  for (int i = 0; i < kVec; ++i) {
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
  for (int i = 0; i < kVec; ++i) {
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
        if (lo_off)       // fp32-equivalent mode: second bf16 plane with the rounding residual
          reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(y) + lo_off + row * width)[c] =
              make_uint2(pack_bf16x2_lo(r.x, r.y, o2.x), pack_bf16x2_lo(r.z, r.w, o2.y));
      } else {
        reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + row * width)[c] = r;
      }
    }
  }
}

int layernorm(const float* x, const float* gamma, const float* beta, void* y, int64_t rows, int width, float eps,
              int out_bf16, cudaStream_t stream, bool reverse, int64_t lo_off) {
  VSCB_REQUIRE(width % 4 == 0 && width <= 128 * kLnWideVec, "layernorm: width must be a multiple of 4 and <= 2048");
  const bool wide = width > 128 * kLnMaxVec;
  if (rows == 0) return VSCB200_OK;
  const int rows_per_block = 8;
  const unsigned grid = static_cast<unsigned>((rows + rows_per_block - 1) / rows_per_block);
  ProfScope prof(kProfLayerNorm, stream, static_cast<double>(rows) * width * (out_bf16 ? 6 : 8));
  const int rv = reverse ? 1 : 0;
  const int64_t no_lo = 0;
  if (out_bf16 && !wide)
    VSCB_CUDA_OK(launch_pdl(layernorm_kernel<true, kLnMaxVec>, dim3(grid), dim3(256), 0, stream, x, gamma, beta, y, rows, width, eps, rv, lo_off));
  else if (out_bf16)
    VSCB_CUDA_OK(launch_pdl(layernorm_kernel<true, kLnWideVec>, dim3(grid), dim3(256), 0, stream, x, gamma, beta, y, rows, width, eps, rv, lo_off));
  else if (!wide)
    VSCB_CUDA_OK(launch_pdl(layernorm_kernel<false, kLnMaxVec>, dim3(grid), dim3(256), 0, stream, x, gamma, beta, y, rows, width, eps, rv, no_lo));
  else
    VSCB_CUDA_OK(launch_pdl(layernorm_kernel<false, kLnWideVec>, dim3(grid), dim3(256), 0, stream, x, gamma, beta, y, rows, width, eps, rv, no_lo));
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

// ------------------------------------------------------------------ fp32 -> bf16 (weights, with row padding)
__global__ void cast_pad_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int64_t rows, int cols,
                                int ld_out, __nv_bfloat16* __restrict__ y_lo) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * ld_out) return;
  const int64_t r = i / ld_out;
  const int c = static_cast<int>(i % ld_out);
  const float v = c < cols ? x[r * cols + c] : 0.f;
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  y[i] = h;
  if (y_lo) y_lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}

int cast_f32_bf16_padded(const float* x, void* y, int64_t rows, int cols, int ld_out, cudaStream_t stream, void* y_lo) {
  const int64_t n = rows * ld_out;
  if (n == 0) return VSCB200_OK;
  cast_pad_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
      x, reinterpret_cast<__nv_bfloat16*>(y), rows, cols, ld_out, reinterpret_cast<__nv_bfloat16*>(y_lo));
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

// ------------------------------------------------------------------ im2row for the stride==kernel patch conv
// frames [n,3,H,H] fp32 NCHW -> patches [n*P, Kp] bf16, column k = c*p*p + i*p + j (the flattening of
// conv1.weight [W,3,p,p], clip.py:105), zero padded to Kp.
__global__ void im2row_kernel(const float* __restrict__ frames, __nv_bfloat16* __restrict__ patches, int64_t n,
                              int img, int patch, int Kp, int64_t lo_off) {
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
  const uint32_t hi = pack_bf16x2(v.x, v.y);
  reinterpret_cast<uint32_t*>(patches)[idx] = hi;
  if (lo_off) reinterpret_cast<uint32_t*>(patches + lo_off)[idx] = pack_bf16x2_lo(v.x, v.y, hi);
}

// Fast path (patch % 4 == 0, no K padding): one thread per 4 consecutive pixels of an image row -- a 128-bit
// coalesced read of the frame, one 8-byte store into the patch row (4 threads fill a 32-byte sector).
__global__ void __launch_bounds__(256)
im2row_vec4_kernel(const float4* __restrict__ frames, uint2* __restrict__ patches, int64_t total, int img, int patch,
                   int64_t lo_off4) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int q = img >> 2;                                   // float4 per image row
  const int x4 = static_cast<int>(idx % q);
  const int64_t t = idx / q;
  const int y = static_cast<int>(t % img);
  const int64_t fc = t / img;
  const int c = static_cast<int>(fc % 3);
  const int64_t f = fc / 3;
  const int grid = img / patch, x = x4 * 4;
  const int py = y / patch, i = y % patch, px = x / patch, j = x % patch;
  const int K = 3 * patch * patch;
  const int64_t dst = ((f * grid + py) * grid + px) * K + (c * patch + i) * patch + j;     // bf16 elements
  const float4 v = ldg_nc_f4(frames + idx);
  const uint2 hi = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  patches[dst >> 2] = hi;
  if (lo_off4) patches[lo_off4 + (dst >> 2)] = make_uint2(pack_bf16x2_lo(v.x, v.y, hi.x), pack_bf16x2_lo(v.z, v.w, hi.y));
}

int im2row(const float* frames, void* patches, int64_t n, int img, int patch, int Kp, cudaStream_t stream, int64_t lo_off) {
  VSCB_REQUIRE(patch % 2 == 0 && img % patch == 0 && Kp % 2 == 0, "im2row: patch must be even and divide img");
  VSCB_REQUIRE(lo_off % 4 == 0, "im2row: lo plane offset must be a multiple of 4 elements");
  const int P = (img / patch) * (img / patch);
  if (n == 0) return VSCB200_OK;
  if (patch % 4 == 0 && Kp == 3 * patch * patch && (reinterpret_cast<uintptr_t>(frames) & 15) == 0) {
    const int64_t total4 = n * 3 * img * (img / 4);
    ProfScope prof(kProfVitOther, stream, static_cast<double>(total4) * 24);
    im2row_vec4_kernel<<<static_cast<unsigned>((total4 + 255) / 256), 256, 0, stream>>>(
        reinterpret_cast<const float4*>(frames), reinterpret_cast<uint2*>(patches), total4, img, patch, lo_off / 4);
    count_launch();
    VSCB_CUDA_OK(cudaGetLastError());
    return VSCB200_OK;
  }
  const int64_t total = n * P * (Kp / 2);
  ProfScope prof(kProfVitOther, stream, static_cast<double>(total) * 2 * 6);
  im2row_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      frames, reinterpret_cast<__nv_bfloat16*>(patches), n, img, patch, Kp, lo_off);
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

template <bool kLN, int kVec>
__global__ void __launch_bounds__(kTailThreads)
gem_head_kernel(const float* __restrict__ y, const float* __restrict__ gamma, const float* __restrict__ beta,
                float* __restrict__ pooled, int T, int C, float eps, float p) {
  extern __shared__ float tail_smem[];   // [C] pooled sums / g, then (LN form) [nwarp][C] per-warp partial sums
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarp = kTailThreads >> 5;
  const int64_t f = blockIdx.x;
  const float* yf = y + f * T * C;
  for (int c = tid; c < C; c += kTailThreads) tail_smem[c] = 0.f;
  __syncthreads();
  const bool cube = (p == 3.0f);
  if (kLN) {
    // one warp per token row; lane owns columns lane*4 + 128*i
    const int nvec = C >> 2;
    float4 acc[kVec];
#pragma unroll
    for (int i = 0; i < kVec; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = warp; t < T; t += nwarp) {
      const float4* xr = reinterpret_cast<const float4*>(yf + static_cast<int64_t>(t) * C);
      float4 v[kVec];
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < kVec; ++i) {
        const int c = lane + 32 * i;
        if (c < nvec) { v[i] = xr[c]; sum += (v[i].x + v[i].y) + (v[i].z + v[i].w); }
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
#pragma unroll
      for (int i = 0; i < kVec; ++i) {
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
    // per-warp partial sums, then a FIXED-order sum over the warps: the pooled descriptor is bit-reproducible
    float* part = tail_smem + C + warp * C;
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) *reinterpret_cast<float4*>(part + 4 * c) = acc[i];
    }
    __syncthreads();
    for (int c = tid; c < C; c += kTailThreads) {
      float a = 0.f;
      for (int w2 = 0; w2 < nwarp; ++w2) a += tail_smem[C + w2 * C + c];
      tail_smem[c] = a;
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
  // pooled descriptor g[f, :] (fp32); the Linear runs as its own kernel that reuses every weight row over 8 frames
  for (int c = tid; c < C; c += kTailThreads) {
    const float m = tail_smem[c] / T;
    pooled[f * C + c] = cube ? cbrtf(m) : powf(m, 1.0f / p);
  }
}

// out[f, o] = head_b[o] + sum_c head_w[o, c] * g[f, c]; CTA = 8 frames x 64 outputs, g of the 8 frames in smem,
// one warp per output row at a time (coalesced weight read, 8 accumulators)
constexpr int kHlFrames = 8, kHlOuts = 64;
__global__ void __launch_bounds__(256)
head_linear_kernel(const float* __restrict__ g, const float* __restrict__ head_w, const float* __restrict__ head_b,
                   float* __restrict__ out, int64_t n, int C, int out_dim) {
  extern __shared__ float hl_smem[];                       // [kHlFrames][C]
  const int64_t f0 = static_cast<int64_t>(blockIdx.x) * kHlFrames;
  const int o0 = blockIdx.y * kHlOuts;
  const int nf = static_cast<int>(n - f0 < kHlFrames ? n - f0 : kHlFrames);
  for (int i = threadIdx.x; i < kHlFrames * C; i += 256) hl_smem[i] = (i / C < nf) ? g[f0 * C + i] : 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int o = o0 + warp; o < o0 + kHlOuts && o < out_dim; o += 8) {
    const float* wr = head_w + static_cast<int64_t>(o) * C;
    float a[kHlFrames];
#pragma unroll
    for (int u = 0; u < kHlFrames; ++u) a[u] = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float wv = __ldg(wr + c);
#pragma unroll
      for (int u = 0; u < kHlFrames; ++u) a[u] += wv * hl_smem[u * C + c];
    }
#pragma unroll
    for (int u = 0; u < kHlFrames; ++u) {
#pragma unroll
      for (int s2 = 16; s2 > 0; s2 >>= 1) a[u] += __shfl_xor_sync(0xffffffffu, a[u], s2);
    }
    if (lane == 0) {
      const float b = head_b[o];
      for (int u = 0; u < nf; ++u) out[(f0 + u) * out_dim + o] = a[u] + b;
    }
  }
}

int gem_head(const float* y, const float* gamma, const float* beta, const float* head_w, const float* head_b,
             float* out, int64_t n, int T, int C, int out_dim, float eps, float p, bool fuse_ln,
             cudaStream_t stream) {
  if (n == 0) return VSCB200_OK;
  VSCB_REQUIRE(!fuse_ln || (C % 4 == 0 && C <= 128 * kLnWideVec), "gem_head: fused LN needs width % 4 == 0 and <= 2048");
  // pooled sums [C] (+ the per-warp partial sums of the fused-LN form: deterministic pooling, no atomics)
  const size_t smem = static_cast<size_t>(C) * sizeof(float) * (fuse_ln ? 1 + kTailThreads / 32 : 1);
  if (fuse_ln) {
    VSCB_CUDA_OK(cudaFuncSetAttribute(gem_head_kernel<true, kLnMaxVec>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    VSCB_CUDA_OK(cudaFuncSetAttribute(gem_head_kernel<true, kLnWideVec>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  }
  const size_t hl_smem_bytes = static_cast<size_t>(kHlFrames) * C * sizeof(float);
  VSCB_REQUIRE(hl_smem_bytes <= 200 * 1024, "gem_head: width too large for the head kernel");
  VSCB_CUDA_OK(cudaFuncSetAttribute(head_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(hl_smem_bytes)));
  float* pooled = nullptr;
  int rc = pool_alloc(reinterpret_cast<void**>(&pooled), static_cast<size_t>(n) * C * sizeof(float), stream);
  if (rc) return rc;
  ProfScope prof(kProfVitOther, stream, static_cast<double>(n) * T * C * 4);
  if (fuse_ln && C <= 128 * kLnMaxVec)
    gem_head_kernel<true, kLnMaxVec><<<static_cast<unsigned>(n), kTailThreads, smem, stream>>>(y, gamma, beta, pooled, T, C, eps, p);
  else if (fuse_ln)
    gem_head_kernel<true, kLnWideVec><<<static_cast<unsigned>(n), kTailThreads, smem, stream>>>(y, gamma, beta, pooled, T, C, eps, p);
  else
    gem_head_kernel<false, kLnMaxVec><<<static_cast<unsigned>(n), kTailThreads, smem, stream>>>(y, gamma, beta, pooled, T, C, eps, p);
  count_launch();
  dim3 grid(static_cast<unsigned>((n + kHlFrames - 1) / kHlFrames), static_cast<unsigned>((out_dim + kHlOuts - 1) / kHlOuts));
  head_linear_kernel<<<grid, 256, hl_smem_bytes, stream>>>(pooled, head_w, head_b, out, n, C,
                                                                                              out_dim);
  count_launch();
  pool_free(pooled, stream);
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

}  // namespace vscb200

extern "C" {
int vscb200_layernorm(const float* x, const float* gamma, const float* beta, void* y, int64_t rows, int width,
                      float eps, int out_bf16, void* stream) {
  return vscb200::layernorm(x, gamma, beta, y, rows, width, eps, out_bf16, static_cast<cudaStream_t>(stream), false, 0);
}
int vscb200_split_f32_bf16(const float* x, void* hi_bf16, void* lo_bf16, int64_t count, void* stream) {
  return vscb200::cast_f32_bf16_padded(x, hi_bf16, count, 1, 1, static_cast<cudaStream_t>(stream), lo_bf16);
}
int vscb200_cast_f32_bf16(const float* x, void* y_bf16, int64_t count, void* stream) {
  return vscb200::cast_f32_bf16_padded(x, y_bf16, count, 1, 1, static_cast<cudaStream_t>(stream), nullptr);
}
}
