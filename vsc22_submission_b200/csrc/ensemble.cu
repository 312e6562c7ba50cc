// Ensemble tail between the encoders and the index:
//   per-model sklearn `normalize` (row L2) -> concatenate -> PCA.transform = (X - mean_) @ components_.T
// Reference: D/infer/concat_pca_sn.py:56-64 (reference bank), D/infer/extract_query_feats.py:169-204 and
// M/infer/infer_matching.py:140-145 (queries) -- numpy/sklearn on the host, one call per video.  Here: one
// gather kernel (normalise + concatenate + centre, warp per frame) writing split-bf16 operand planes and the fp32-equivalent
// tcgen05 GEMM of gemm.cu for the [n, D] x [out, D]^T projection (batches under 256 frames: the exact-fp32 FFMA tile kernel
// of sim.cu); every descriptor of a batch in one call, nothing leaves the device.
#include <stdlib.h>

#include "host_util.h"
#include "kernels.h"
#include "ptx.cuh"

namespace vscb200 {

constexpr int kEnsMaxParts = 8;
struct EnsParts {
  const float* ptr[kEnsMaxParts];
  int dim[kEnsMaxParts];
  int off[kEnsMaxParts];
  int n_parts;
};

// xc[row, off_p + c] = parts[p][row, c] / max(||parts[p][row]||, tiny) - mean[off_p + c]
// (sklearn.preprocessing.normalize leaves all-zero rows untouched: norm 0 -> divide by 1)
__global__ void __launch_bounds__(256)
ensemble_center_kernel(EnsParts parts, const float* __restrict__ mean, float* __restrict__ xc, int64_t n, int D) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= n) return;
  for (int p = 0; p < parts.n_parts; ++p) {
    const float* src = parts.ptr[p] + row * parts.dim[p];
    float ss = 0.f;
    for (int c = lane; c < parts.dim[p]; c += 32) { const float v = src[c]; ss = fmaf(v, v, ss); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float nrm = sqrtf(ss);
    const float inv = nrm > 0.f ? 1.0f / nrm : 1.0f;
    float* dst = xc + row * D + parts.off[p];
    const float* mu = mean + parts.off[p];
    for (int c = lane; c < parts.dim[p]; c += 32) dst[c] = src[c] * inv - mu[c];
  }
}

// the same, written as split-bf16 operand planes (hi = bf16(v), lo = bf16(v - hi)) for the tensor-core projection
__global__ void __launch_bounds__(256)
ensemble_center_planes_kernel(EnsParts parts, const float* __restrict__ mean, __nv_bfloat16* __restrict__ hi,
                              __nv_bfloat16* __restrict__ lo, int64_t n, int D) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= n) return;
  for (int p = 0; p < parts.n_parts; ++p) {
    const float* src = parts.ptr[p] + row * parts.dim[p];
    float ss = 0.f;
    for (int c = lane; c < parts.dim[p]; c += 32) { const float v = src[c]; ss = fmaf(v, v, ss); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float nrm = sqrtf(ss);
    const float inv = nrm > 0.f ? 1.0f / nrm : 1.0f;
    const int64_t base = row * D + parts.off[p];
    const float* mu = mean + parts.off[p];
    for (int c = lane; c < parts.dim[p]; c += 32) {
      const float v = src[c] * inv - mu[c];
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      hi[base + c] = h;
      lo[base + c] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
}

// ---- near-duplicate frame filter of one query video (extract_query_feats.py:190-200, infer_matching.py query side) -------
// feat = rows / ||row||; sim = feat feat^T - I (float64 from float32 products, as numpy promotes it); frames are visited
// by descending column mean of sim; a visited frame that is still alive removes every frame whose similarity to it
// exceeds the threshold.  One CTA: means -> ranks -> the greedy sweep (sequential in the visiting order by definition).
__global__ void __launch_bounds__(256)
row_normalize_kernel(const float* __restrict__ x, int64_t n, int d, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= n) return;
  const float* src = x + row * d;
  float ss = 0.f;
  for (int c = lane; c < d; c += 32) { const float v = src[c]; ss = fmaf(v, v, ss); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float nrm = sqrtf(ss);
  for (int c = lane; c < d; c += 32) out[row * d + c] = src[c] / nrm;
}

constexpr int kDupMaxFrames = 4096;

__global__ void __launch_bounds__(1024)
near_dup_kernel(const float* __restrict__ sim, int n, double thr, double* __restrict__ mean, int* __restrict__ order,
                uint8_t* __restrict__ keep) {
  __shared__ uint8_t removed[kDupMaxFrames];
  __shared__ int alive;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    double sum = 0.0;
    for (int i = 0; i < n; ++i) sum += static_cast<double>(sim[static_cast<int64_t>(i) * n + j]) - (i == j ? 1.0 : 0.0);
    mean[j] = sum / n;
    removed[j] = 0;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    // argsort()[::-1]: descending, equal values by descending index.  A NaN mean (an all-zero descriptor row) ranks
    // as -inf so that the ranks stay a permutation.
    const double mi = mean[i] == mean[i] ? mean[i] : -INFINITY;
    int pos = 0;
    for (int j = 0; j < n; ++j) {
      const double mj = mean[j] == mean[j] ? mean[j] : -INFINITY;
      pos += (mj > mi || (mj == mi && j > i)) ? 1 : 0;
    }
    order[pos] = i;
  }
  __syncthreads();
  for (int p = 0; p < n; ++p) {
    const int i = order[p];
    if (threadIdx.x == 0) alive = removed[i] ? 0 : 1;
    __syncthreads();
    if (alive) {
      for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const double v = static_cast<double>(sim[static_cast<int64_t>(i) * n + j]) - (i == j ? 1.0 : 0.0);
        if (v > thr) removed[j] = 1;
      }
    }
    __syncthreads();
  }
  for (int j = threadIdx.x; j < n; j += blockDim.x) keep[j] = removed[j] ? 0 : 1;
}

}  // namespace vscb200

using namespace vscb200;

extern "C" int vscb200_near_dup_keep(const float* feat_dev, int64_t n, int d, double threshold, uint8_t* keep_dev,
                                     void* stream_v) {
  VSCB_REQUIRE(n >= 0 && d > 0, "near_dup_keep: bad shape");
  if (n == 0) return VSCB200_OK;
  VSCB_REQUIRE(feat_dev && keep_dev, "near_dup_keep: null argument");
  VSCB_REQUIRE(n <= kDupMaxFrames, "near_dup_keep: more than 4096 frames in one video");
  cudaStream_t s = static_cast<cudaStream_t>(stream_v);
  float* fn = nullptr;
  const size_t bytes = static_cast<size_t>(n) * d * sizeof(float) + static_cast<size_t>(n) * n * sizeof(float) +
                       static_cast<size_t>(n) * (sizeof(double) + sizeof(int)) + 64;
  int rc = pool_alloc(reinterpret_cast<void**>(&fn), bytes, s);
  if (rc) return rc;
  float* sim = fn + n * d;
  double* mean = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(sim + n * n) + 15) & ~uintptr_t(15));
  int* order = reinterpret_cast<int*>(mean + n);
  row_normalize_kernel<<<static_cast<unsigned>((n + 7) / 8), 256, 0, s>>>(feat_dev, n, d, fn);
  count_launch();
  rc = scores_simt(fn, fn, sim, n, n, d, n, false, nullptr, nullptr, s);
  if (rc == VSCB200_OK) {
    near_dup_kernel<<<1, 1024, 0, s>>>(sim, static_cast<int>(n), threshold, mean, order, keep_dev);
    count_launch();
  }
  pool_free(fn, s);
  if (rc) return rc;
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

extern "C" int vscb200_ensemble_pca(const float* const* parts_dev, const int* dims, int n_parts, int64_t n,
                                    const float* mean_dev, const float* components_dev, int out_dim, float* out_dev,
                                    void* stream_v) {
  VSCB_REQUIRE(parts_dev && dims && n_parts >= 1 && n_parts <= kEnsMaxParts, "ensemble_pca: 1..8 descriptor sets");
  VSCB_REQUIRE(n >= 0 && out_dim > 0 && mean_dev && components_dev && (n == 0 || out_dev), "ensemble_pca: null argument");
  if (n == 0) return VSCB200_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream_v);
  EnsParts parts = {};
  parts.n_parts = n_parts;
  int D = 0;
  for (int p = 0; p < n_parts; ++p) {
    VSCB_REQUIRE(parts_dev[p] != nullptr && dims[p] > 0, "ensemble_pca: bad descriptor set");
    parts.ptr[p] = parts_dev[p];
    parts.dim[p] = dims[p];
    parts.off[p] = D;
    D += dims[p];
  }
  // Batches of frames: the [n, D] x [out, D]^T projection on the tensor cores in the split-bf16 form (hi.hi + lo.hi + hi.lo,
  // fp32 accumulation: fp32-equivalent to ~1e-6 relative, gemm.cu) -- 2 n D out FLOP leave the FFMA pipe.
  static const int pca_simt = [] { const char* e = getenv("VSCB200_PCA_SIMT"); return e ? atoi(e) : 0; }();
  if (!pca_simt && n >= 256 && D % 8 == 0 && out_dim % 8 == 0 && (reinterpret_cast<uintptr_t>(out_dev) & 15) == 0) {
    const size_t xe = static_cast<size_t>(n) * D, we = static_cast<size_t>(out_dim) * D;
    uint16_t* planes = nullptr;
    int rc = pool_alloc(reinterpret_cast<void**>(&planes), (2 * xe + 2 * we) * sizeof(uint16_t), s);
    if (rc) return rc;
    uint16_t *xh = planes, *xl = planes + xe, *wh = planes + 2 * xe, *wl = wh + we;
    {
      ProfScope prof(kProfVitOther, s, static_cast<double>(n) * D * 8);
      ensemble_center_planes_kernel<<<static_cast<unsigned>((n + 7) / 8), 256, 0, s>>>(
          parts, mean_dev, reinterpret_cast<__nv_bfloat16*>(xh), reinterpret_cast<__nv_bfloat16*>(xl), n, D);
      count_launch();
    }
    rc = split_planes(components_dev, wh, wl, out_dim, D, D, s);
    if (rc == VSCB200_OK)
      rc = gemm_bf16(xh, wh, nullptr, out_dev, n, out_dim, D, D, D, out_dim, VSCB200_EPI_F32, -1, s, nullptr, 0, false, 0, nullptr,
                     xl, wl, nullptr);
    pool_free(planes, s);
    if (rc) return rc;
    VSCB_CUDA_OK(cudaGetLastError());
    return VSCB200_OK;
  }
  float* xc = nullptr;
  int rc = pool_alloc(reinterpret_cast<void**>(&xc), static_cast<size_t>(n) * D * sizeof(float), s);
  if (rc) return rc;
  {
    ProfScope prof(kProfVitOther, s, static_cast<double>(n) * D * 8);
    ensemble_center_kernel<<<static_cast<unsigned>((n + 7) / 8), 256, 0, s>>>(parts, mean_dev, xc, n, D);
    count_launch();
  }
  for (int64_t r0 = 0; r0 < n && rc == VSCB200_OK; r0 += 65535ll * 64) {
    const int64_t nb = n - r0 < 65535ll * 64 ? n - r0 : 65535ll * 64;
    rc = scores_simt(xc + r0 * D, components_dev, out_dev + r0 * out_dim, nb, out_dim, D, out_dim, false, nullptr, nullptr, s);
  }
  pool_free(xc, s);
  if (rc) return rc;
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}
