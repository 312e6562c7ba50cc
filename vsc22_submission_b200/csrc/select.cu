// Row-wise selection over a dense score block S[nq, n] (float32, device):
//   * exact top-k per row (radix select of the k-th key + ordered tie handling + bitonic sort)
//   * range (threshold) search per row: count -> host/device scan -> ordered fill (CSR)
// faiss semantics restated in oracle/faiss_np.py: best-first, ties to the lower id, strict
// threshold, ascending ids inside a range row.  Reference call sites: vsc/index.py:174,
// vsc/exhaustive_search.py:62,74; score_normalization.py:95; M/infer/infer_matching.py:232-235.
#include <float.h>

#include "host_util.h"
#include "kernels.h"
#include "ptx.cuh"

namespace vscb200 {

constexpr int kSelThreads = 512;

// order-preserving float -> uint32 (ascending).  keep_max: larger score is better; else smaller.
__device__ __forceinline__ uint32_t okey(float f, bool keep_max) {
  const uint32_t u = __float_as_uint(f);
  const uint32_t k = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return keep_max ? k : ~k;
}

__global__ void __launch_bounds__(kSelThreads)
topk_select_kernel(const float* __restrict__ S, int64_t ldS, int64_t n, int k, int kpad, int keep_max_i,
                   float* __restrict__ D, int64_t* __restrict__ I, int64_t id_offset) {
  extern __shared__ unsigned long long cand[];     // [kpad] (key << 32) | ~idx
  __shared__ unsigned int hist[256];
  __shared__ unsigned int s_prefix, s_remaining, s_gt_slots, s_warp_tot[kSelThreads / 32];
  const bool keep_max = keep_max_i != 0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* row = S + static_cast<int64_t>(blockIdx.x) * ldS;
  float* Dr = D + static_cast<int64_t>(blockIdx.x) * k;
  int64_t* Ir = I + static_cast<int64_t>(blockIdx.x) * k;
  const int kk = static_cast<int>(n < k ? n : k);    // entries that exist

  for (int i = tid; i < kpad; i += kSelThreads) cand[i] = 0ull;
  if (tid == 0) { s_prefix = 0; s_remaining = kk; s_gt_slots = 0; }
  __syncthreads();

  if (kk > 0) {
    // ---- radix select: key of the kk-th best element
    uint32_t mask = 0;
    for (int shift = 24; shift >= 0; shift -= 8) {
      for (int b = tid; b < 256; b += kSelThreads) hist[b] = 0;
      __syncthreads();
      const uint32_t prefix = s_prefix;
      for (int64_t i = tid; i < n; i += kSelThreads) {
        const uint32_t key = okey(row[i], keep_max);
        if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xFFu], 1u);
      }
      __syncthreads();
      if (tid == 0) {
        unsigned int remaining = s_remaining, cum = 0;
        int b = 255;
        for (; b > 0; --b) {
          if (cum + hist[b] >= remaining) break;
          cum += hist[b];
        }
        s_remaining = remaining - cum;             // how many to take inside bin b
        s_prefix = prefix | (static_cast<uint32_t>(b) << shift);
      }
      mask |= 0xFFu << shift;
      __syncthreads();
    }
    const uint32_t kth = s_prefix;
    const unsigned int need_eq = s_remaining;        // >= 1 elements equal to kth, lowest ids first
    const unsigned int n_gt = kk - need_eq;
    // ---- collect: strictly better (any order) + ties in id order
    unsigned int eq_seen = 0;
    for (int64_t base = 0; base < n; base += kSelThreads) {
      const int64_t i = base + tid;
      uint32_t key = 0;
      bool gt = false, eq = false;
      if (i < n) {
        key = okey(row[i], keep_max);
        gt = key > kth;
        eq = (key == kth) && (eq_seen < need_eq);
      }
      if (gt) {
        const unsigned int slot = atomicAdd(&s_gt_slots, 1u);
        cand[slot] = (static_cast<unsigned long long>(key) << 32) | static_cast<uint32_t>(~static_cast<uint32_t>(i));
      }
      const unsigned int bal = __ballot_sync(0xffffffffu, eq);
      if (lane == 0) s_warp_tot[warp] = __popc(bal);
      __syncthreads();
      unsigned int before = 0, total = 0;
#pragma unroll
      for (int w = 0; w < kSelThreads / 32; ++w) {
        const unsigned int c = s_warp_tot[w];
        if (w < warp) before += c;
        total += c;
      }
      if (eq) {
        const unsigned int rank = eq_seen + before + __popc(bal & ((1u << lane) - 1u));
        if (rank < need_eq)
          cand[n_gt + rank] = (static_cast<unsigned long long>(key) << 32) | static_cast<uint32_t>(~static_cast<uint32_t>(i));
      }
      eq_seen += total;
      __syncthreads();
    }
  }
  __syncthreads();
  // ---- bitonic sort, descending on (key, ~idx)  => best first, ties to the lower id
  for (int size = 2; size <= kpad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < (kpad >> 1); i += kSelThreads) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const unsigned long long a = cand[lo], b = cand[hi];
        if ((a < b) == desc) { cand[lo] = b; cand[hi] = a; }
      }
      __syncthreads();
    }
  }
  for (int j = tid; j < k; j += kSelThreads) {
    if (j < kk) {
      const uint32_t idx = ~static_cast<uint32_t>(cand[j] & 0xFFFFFFFFull);
      Dr[j] = row[idx];
      Ir[j] = id_offset + idx;
    } else {
      Dr[j] = keep_max ? -FLT_MAX : FLT_MAX;
      Ir[j] = -1;
    }
  }
}

int topk_rows(const float* S, int64_t ldS, int64_t nq, int64_t n, int k, bool keep_max, float* D, int64_t* I,
              int64_t id_offset, cudaStream_t stream) {
  VSCB_REQUIRE(k >= 1 && k <= 2048, "search: k must be in [1, 2048]");
  VSCB_REQUIRE(n < (1ll << 32), "search: at most 2^32-1 bank rows per device");
  if (nq == 0) return VSCB200_OK;
  int kpad = 2;
  while (kpad < k) kpad <<= 1;
  ProfScope prof(kProfSelect, stream, static_cast<double>(nq) * n * 4);
  topk_select_kernel<<<static_cast<unsigned>(nq), kSelThreads, kpad * sizeof(unsigned long long), stream>>>(
      S, ldS, n, k, kpad, keep_max ? 1 : 0, D, I, id_offset);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

// ------------------------------------------------------------------ range search
constexpr int kRangeThreads = 256;

__global__ void __launch_bounds__(kRangeThreads)
range_count_kernel(const float* __restrict__ S, int64_t ldS, int64_t n, float thr, int keep_max,
                   unsigned long long* __restrict__ counts) {
  const float* row = S + static_cast<int64_t>(blockIdx.x) * ldS;
  unsigned int c = 0;
  for (int64_t i = threadIdx.x; i < n; i += kRangeThreads) {
    const float v = row[i];
    c += keep_max ? (v > thr) : (v < thr);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  __shared__ unsigned int wsum[kRangeThreads / 32];
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < kRangeThreads / 32; ++w) t += wsum[w];
    counts[blockIdx.x] = t;
  }
}

// offsets[row] = exclusive prefix of counts (element offsets inside this block's output)
__global__ void __launch_bounds__(kRangeThreads)
range_fill_kernel(const float* __restrict__ S, int64_t ldS, int64_t n, float thr, int keep_max,
                  const unsigned long long* __restrict__ offsets, float* __restrict__ D, int64_t* __restrict__ I,
                  int64_t id_offset) {
  __shared__ unsigned int wtot[kRangeThreads / 32];
  const float* row = S + static_cast<int64_t>(blockIdx.x) * ldS;
  unsigned long long pos = offsets[blockIdx.x];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t base = 0; base < n; base += kRangeThreads) {
    const int64_t i = base + threadIdx.x;
    float v = 0.f;
    bool hit = false;
    if (i < n) {
      v = row[i];
      hit = keep_max ? (v > thr) : (v < thr);
    }
    const unsigned int bal = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) wtot[warp] = __popc(bal);
    __syncthreads();
    unsigned int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kRangeThreads / 32; ++w) {
      const unsigned int c = wtot[w];
      if (w < warp) before += c;
      total += c;
    }
    if (hit) {
      const unsigned long long o = pos + before + __popc(bal & ((1u << lane) - 1u));
      D[o] = v;
      I[o] = id_offset + i;
    }
    pos += total;
    __syncthreads();
  }
}

int range_count(const float* S, int64_t ldS, int64_t nq, int64_t n, float thr, bool keep_max,
                unsigned long long* counts, cudaStream_t stream) {
  if (nq == 0) return VSCB200_OK;
  range_count_kernel<<<static_cast<unsigned>(nq), kRangeThreads, 0, stream>>>(S, ldS, n, thr, keep_max ? 1 : 0, counts);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

int range_fill(const float* S, int64_t ldS, int64_t nq, int64_t n, float thr, bool keep_max,
               const unsigned long long* offsets, float* D, int64_t* I, int64_t id_offset, cudaStream_t stream) {
  if (nq == 0) return VSCB200_OK;
  range_fill_kernel<<<static_cast<unsigned>(nq), kRangeThreads, 0, stream>>>(S, ldS, n, thr, keep_max ? 1 : 0, offsets,
                                                                             D, I, id_offset);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

}  // namespace vscb200
