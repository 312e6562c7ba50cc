// Frame-similarity matrices of ALL candidate video pairs in one launch, plus their per-row top-k.
//
// Reference: LocalizationWithMetadata.similarity, vsc/baseline/localization.py:32-35 (`np.matmul(a, b.T)` per candidate
// pair, + similarity_bias :52-57, looped over ~10^4..10^5 pairs and fed to a 16-process pool, sscd_baseline.py:107-152)
// and the first step of the temporal-network alignment, vcsl/vta.py:262-265 (`np.argsort(-sims)[:, :tn_top_k]`).
// Exact fp32 FFMA arithmetic (the reference's is numpy sgemm): a pair's matrix is a few thousand products, far below
// a tensor-core tile, so the grouped form is a gather-bound batch of small GEMMs: one CTA per pair walks its 64x64
// tiles; descriptors of every video are read from the shared feature arrays (no per-pair copies).
#include <float.h>

#include "host_util.h"
#include "kernels.h"
#include "ptx.cuh"

namespace vscb200 {

struct PairDesc {
  const int64_t* q_off; const int32_t* q_len; const int64_t* r_off; const int32_t* r_len; const int64_t* s_off;
};

constexpr int kPT = 64, kPK = 16;

__global__ void __launch_bounds__(256)
pair_sims_kernel(const float* __restrict__ Q, const float* __restrict__ R, int d, PairDesc pd, float bias,
                 float* __restrict__ S) {
  __shared__ float sQ[kPK][kPT + 4];
  __shared__ float sR[kPK][kPT + 4];
  const int64_t pair = blockIdx.x;
  const int nq = pd.q_len[pair], nr = pd.r_len[pair];
  const float* q = Q + pd.q_off[pair] * d;
  const float* r = R + pd.r_off[pair] * d;
  float* s = S + pd.s_off[pair];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int tiles_r = (nr + kPT - 1) / kPT, tiles = ((nq + kPT - 1) / kPT) * tiles_r;
  for (int t = blockIdx.y; t < tiles; t += gridDim.y) {
    const int q0 = (t / tiles_r) * kPT, r0 = (t % tiles_r) * kPT;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < d; k0 += kPK) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int idx = tid + e * 256;
        const int rr = idx >> 4, kk = idx & 15;
        const bool kin = k0 + kk < d;
        sQ[kk][rr] = (kin && q0 + rr < nq) ? q[static_cast<int64_t>(q0 + rr) * d + k0 + kk] : 0.f;
        sR[kk][rr] = (kin && r0 + rr < nr) ? r[static_cast<int64_t>(r0 + rr) * d + k0 + kk] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < kPK; ++kk) {
        const float4 a = *reinterpret_cast<const float4*>(&sQ[kk][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&sR[kk][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int qi = q0 + ty * 4 + i;
      if (qi >= nq) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int rj = r0 + tx * 4 + j;
        if (rj < nr) s[static_cast<int64_t>(qi) * nr + rj] = acc[i][j] + bias;
      }
    }
  }
}

// per row of every pair's matrix: the k largest entries, best first, ties to the lower column; one warp per row
__global__ void __launch_bounds__(256)
pair_topk_kernel(const float* __restrict__ S, PairDesc pd, const int64_t* __restrict__ row_off, int k,
                 float* __restrict__ topv, int32_t* __restrict__ topi) {
  const int64_t pair = blockIdx.x;
  const int nq = pd.q_len[pair], nr = pd.r_len[pair];
  const float* s = S + pd.s_off[pair];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int row = blockIdx.y * 8 + warp; row < nq; row += gridDim.y * 8) {
    const float* sr = s + static_cast<int64_t>(row) * nr;
    float* ov = topv + (row_off[pair] + row) * k;
    int32_t* oi = topi + (row_off[pair] + row) * k;
    float prev_v = INFINITY;
    int prev_i = -1;
    for (int rnk = 0; rnk < k; ++rnk) {
      // best entry strictly after (prev_v, prev_i) in (value desc, index asc) order
      float bv = -INFINITY;
      int bi = 0x7FFFFFFF;
      for (int c = lane; c < nr; c += 32) {
        const float v = sr[c];
        const bool after = v < prev_v || (v == prev_v && c > prev_i);
        if (after && (v > bv || (v == bv && c < bi))) { bv = v; bi = c; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov2 = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi2 = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov2 > bv || (ov2 == bv && oi2 < bi)) { bv = ov2; bi = oi2; }
      }
      if (lane == 0) {
        ov[rnk] = bi == 0x7FFFFFFF ? -FLT_MAX : bv;
        oi[rnk] = bi == 0x7FFFFFFF ? -1 : bi;
      }
      prev_v = bv;
      prev_i = bi;
      if (bi == 0x7FFFFFFF) prev_v = -INFINITY;
    }
  }
}

// ---- matching-track candidate features ----------------------------------------------------
// M/infer/src/utils.py:18-47 / :50-73: a query video may hold several `seg_len`-frame copies of itself (one per detected
// sub-image); the copy whose rows match the reference best -- mean of the 10 largest row maxima, utils.py:33-41 -- is kept,
// and its similarity block (and the transposed block, utils.py:44-46) is zero-padded / cropped to an H x W image
// (M/infer/src/dataset.py:103-144).  One CTA per pair: row maxima -> per-segment score -> argmax -> images.
constexpr int kSegTop = 10;

// numpy's float32 pairwise add.reduce over a[0..m) for m <= 10 (8 partial sums, combined as a tree, then the tail)
__device__ __forceinline__ float numpy_sum_f32(const float* a, int m) {
  if (m < 8) {
    float r = 0.f;          // -0.0 start in numpy; identical for the sums that occur here
    for (int i = 0; i < m; ++i) r = __fadd_rn(r, a[i]);
    return r;
  }
  float res = __fadd_rn(__fadd_rn(__fadd_rn(a[0], a[1]), __fadd_rn(a[2], a[3])),
                        __fadd_rn(__fadd_rn(a[4], a[5]), __fadd_rn(a[6], a[7])));
  for (int i = 8; i < m; ++i) res = __fadd_rn(res, a[i]);
  return res;
}

__global__ void __launch_bounds__(256)
pair_segment_images_kernel(const float* __restrict__ S, PairDesc pd, const int64_t* __restrict__ row_off,
                           const int32_t* __restrict__ seg_len, float* __restrict__ rowmax, int H, int W, int with_transpose,
                           float* __restrict__ images, int32_t* __restrict__ info) {
  __shared__ float seg_score[256];
  __shared__ int best_seg;
  const int64_t pair = blockIdx.x;
  const int nq = pd.q_len[pair], nr = pd.r_len[pair];
  const float* s = S + pd.s_off[pair];
  float* rm = rowmax + row_off[pair];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sl = seg_len[pair] > 0 ? seg_len[pair] : nq;
  const int nseg = min((nq + sl - 1) / sl, 256);          // seg_score[] capacity; real inputs hold a handful of copies
  if (nseg > 1) {
    for (int row = warp; row < nq; row += 8) {
      float m = -INFINITY;
      for (int c = lane; c < nr; c += 32) m = fmaxf(m, s[static_cast<int64_t>(row) * nr + c]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      if (lane == 0) rm[row] = m;
    }
    __syncthreads();
    for (int sg = warp; sg < nseg; sg += 8) {
      const int a = sg * sl, b = min(nq, a + sl);
      // the (up to) 10 largest row maxima of the segment, found largest first
      float topv[kSegTop];
      float prev_v = INFINITY;
      int prev_i = -1, m = 0;
      for (int rnk = 0; rnk < kSegTop && rnk < b - a; ++rnk) {
        float bv = -INFINITY;
        int bi = 0x7FFFFFFF;
        for (int c = a + lane; c < b; c += 32) {
          const float v = rm[c];
          const bool after = v < prev_v || (v == prev_v && c > prev_i);
          if (after && (v > bv || (v == bv && c < bi))) { bv = v; bi = c; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        topv[rnk] = bv;
        prev_v = bv;
        prev_i = bi;
        ++m;
      }
      if (lane == 0) {
        float asc[kSegTop];
        for (int i = 0; i < m; ++i) asc[i] = topv[m - 1 - i];          // maxs.sort(); maxs[-10:]
        seg_score[sg] = __fdiv_rn(numpy_sum_f32(asc, m), static_cast<float>(m));
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int bs = 0;
      for (int sg = 1; sg < nseg && sg < 256; ++sg)
        if (seg_score[sg] > seg_score[bs]) bs = sg;                     // np.argmax: first maximum
      best_seg = bs;
    }
  } else if (threadIdx.x == 0) {
    best_seg = 0;
  }
  __syncthreads();
  const int start = best_seg * sl, len = min(nq, start + sl) - start;
  const int n_img = with_transpose ? 2 : 1;
  float* img = images + pair * n_img * static_cast<int64_t>(H) * W;
  const int h = min(len, H), w = min(nr, W);
  for (int e = threadIdx.x; e < H * W; e += 256) {
    const int i = e / W, j = e % W;
    img[e] = (i < h && j < w) ? s[static_cast<int64_t>(start + i) * nr + j] : 0.f;
  }
  if (with_transpose) {
    float* imt = img + static_cast<int64_t>(H) * W;
    const int ht = min(nr, H), wt = min(len, W);
    for (int e = threadIdx.x; e < H * W; e += 256) {
      const int i = e / W, j = e % W;                                   // (ref frame, query frame)
      imt[e] = (i < ht && j < wt) ? s[static_cast<int64_t>(start + j) * nr + i] : 0.f;
    }
  }
  if (threadIdx.x == 0) {
    info[pair * 4 + 0] = best_seg; info[pair * 4 + 1] = len; info[pair * 4 + 2] = h; info[pair * 4 + 3] = w;
  }
}

}  // namespace vscb200

using namespace vscb200;

extern "C" {

int vscb200_pair_sims(const float* q_dev, const float* r_dev, int d, int64_t n_pairs, const int64_t* q_off_dev,
                      const int32_t* q_len_dev, const int64_t* r_off_dev, const int32_t* r_len_dev, const int64_t* s_off_dev,
                      float bias, float* sims_dev, void* stream_v) {
  VSCB_REQUIRE(n_pairs >= 0 && d > 0, "pair_sims: bad shape");
  if (n_pairs == 0) return VSCB200_OK;
  VSCB_REQUIRE(q_dev && r_dev && q_off_dev && q_len_dev && r_off_dev && r_len_dev && s_off_dev && sims_dev, "pair_sims: null argument");
  VSCB_REQUIRE(n_pairs < (1ll << 31), "pair_sims: too many pairs in one call");
  cudaStream_t s = static_cast<cudaStream_t>(stream_v);
  PairDesc pd{q_off_dev, q_len_dev, r_off_dev, r_len_dev, s_off_dev};
  ProfScope prof(kProfScores, s, 0.0);
  pair_sims_kernel<<<dim3(static_cast<unsigned>(n_pairs), 4), 256, 0, s>>>(q_dev, r_dev, d, pd, bias, sims_dev);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

int vscb200_pair_topk(const float* sims_dev, int64_t n_pairs, const int32_t* q_len_dev, const int32_t* r_len_dev,
                      const int64_t* s_off_dev, const int64_t* row_off_dev, int k, float* topv_dev, int32_t* topi_dev,
                      void* stream_v) {
  VSCB_REQUIRE(n_pairs >= 0 && k >= 1 && k <= 64, "pair_topk: k must be in [1, 64]");
  if (n_pairs == 0) return VSCB200_OK;
  VSCB_REQUIRE(sims_dev && q_len_dev && r_len_dev && s_off_dev && row_off_dev && topv_dev && topi_dev, "pair_topk: null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream_v);
  PairDesc pd{nullptr, q_len_dev, nullptr, r_len_dev, s_off_dev};
  ProfScope prof(kProfSelect, s, 0.0);
  pair_topk_kernel<<<dim3(static_cast<unsigned>(n_pairs), 4), 256, 0, s>>>(sims_dev, pd, row_off_dev, k, topv_dev, topi_dev);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

int vscb200_pair_segment_images(const float* sims_dev, int64_t n_pairs, const int32_t* q_len_dev, const int32_t* r_len_dev,
                                const int64_t* s_off_dev, const int64_t* row_off_dev, const int32_t* seg_len_dev,
                                float* rowmax_scratch_dev, int H, int W, int with_transpose, float* images_dev,
                                int32_t* info_dev, void* stream_v) {
  VSCB_REQUIRE(n_pairs >= 0 && H >= 1 && W >= 1, "pair_segment_images: bad shape");
  if (n_pairs == 0) return VSCB200_OK;
  VSCB_REQUIRE(sims_dev && q_len_dev && r_len_dev && s_off_dev && row_off_dev && seg_len_dev && rowmax_scratch_dev &&
               images_dev && info_dev, "pair_segment_images: null argument");
  VSCB_REQUIRE(n_pairs < (1ll << 31), "pair_segment_images: too many pairs in one call");
  cudaStream_t s = static_cast<cudaStream_t>(stream_v);
  PairDesc pd{nullptr, q_len_dev, nullptr, r_len_dev, s_off_dev};
  pair_segment_images_kernel<<<static_cast<unsigned>(n_pairs), 256, 0, s>>>(sims_dev, pd, row_off_dev, seg_len_dev,
                                                                          rowmax_scratch_dev, H, W, with_transpose ? 1 : 0,
                                                                          images_dev, info_dev);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

}  // extern "C"
