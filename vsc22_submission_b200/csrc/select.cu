// Row-wise selection over a dense score block S[nq, n] (float32, device):
//   * exact top-k per row (radix select of the k-th key + ordered tie handling + bitonic sort)
//   * range (threshold) search per row: count -> host/device scan -> ordered fill (CSR)
// faiss semantics restated in oracle/faiss_np.py: best-first, ties to the lower id, strict
// threshold, ascending ids inside a range row.  Reference call sites: vsc/index.py:174,
// vsc/exhaustive_search.py:62,74; score_normalization.py:95; M/infer/infer_matching.py:232-235.
#include <float.h>

#include "exact.cuh"
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
topk_select_kernel(const float* __restrict__ S, int64_t ldS, int64_t es, int64_t n_all, int64_t seg_len, int k, int kpad,
                   int keep_max_i, float* __restrict__ D, int64_t* __restrict__ I, int64_t id_offset,
                   const int64_t* __restrict__ ids_in) {
  extern __shared__ unsigned long long cand[];     // [kpad] (key << 32) | ~idx
  __shared__ unsigned int hist[256];
  __shared__ unsigned int s_prefix, s_remaining, s_gt_slots, s_warp_tot[kSelThreads / 32];
  const bool keep_max = keep_max_i != 0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // blockIdx.y = segment of the row (two-level selection for few rows x many columns): this CTA selects among
  // columns [seg0, seg0 + n) and writes its k results at slot (row * gridDim.y + segment)
  const int64_t seg0 = static_cast<int64_t>(blockIdx.y) * seg_len;
  const int64_t n = n_all - seg0 < seg_len ? n_all - seg0 : seg_len;
  const float* row = S + static_cast<int64_t>(blockIdx.x) * ldS + seg0 * es;
  const int64_t* ids_row = ids_in ? ids_in + static_cast<int64_t>(blockIdx.x) * n_all + seg0 : nullptr;
  float* Dr = D + (static_cast<int64_t>(blockIdx.x) * gridDim.y + blockIdx.y) * k;
  int64_t* Ir = I + (static_cast<int64_t>(blockIdx.x) * gridDim.y + blockIdx.y) * k;
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
        const uint32_t key = okey(row[i * es], keep_max);
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
        key = okey(row[i * es], keep_max);
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
      Dr[j] = row[static_cast<int64_t>(idx) * es];
      Ir[j] = ids_row ? ids_row[idx] : id_offset + seg0 + idx;
    } else {
      Dr[j] = keep_max ? -FLT_MAX : FLT_MAX;
      Ir[j] = -1;
    }
  }
}

int topk_rows(const float* S, int64_t ldS, int64_t nq, int64_t n, int k, bool keep_max, float* D, int64_t* I,
              int64_t id_offset, cudaStream_t stream, int64_t es) {
  VSCB_REQUIRE(k >= 1 && k <= 2048, "search: k must be in [1, 2048]");
  VSCB_REQUIRE(n < (1ll << 32), "search: at most 2^32-1 bank rows per device");
  if (nq == 0) return VSCB200_OK;
  int kpad = 2;
  while (kpad < k) kpad <<= 1;
  ProfScope prof(kProfSelect, stream, static_cast<double>(nq) * n * 4);
  // Few rows x many columns (the per-video k = 1024 searches of M/infer/infer_matching.py:232): one CTA per row
  // leaves the GPU idle.  Two levels: every row is cut into segments selected by their own CTAs, then the
  // nsplit * k survivors per row are selected again.  Ties keep resolving to the lower id: a lower segment holds
  // lower ids and each segment's output is already ordered.
  int64_t nsplit = 1;
  if (nq < 2 * device_sm_count() && n >= 65536 && n >= 16ll * k) {
    nsplit = (4ll * device_sm_count() + nq - 1) / nq;
    if (nsplit > n / (8ll * k)) nsplit = n / (8ll * k);
    if (nsplit > 64) nsplit = 64;
  }
  if (nsplit <= 1) {
    topk_select_kernel<<<static_cast<unsigned>(nq), kSelThreads, kpad * sizeof(unsigned long long), stream>>>(
        S, ldS, es, n, n, k, kpad, keep_max ? 1 : 0, D, I, id_offset, nullptr);
    count_launch();
    VSCB_CUDA_OK(cudaGetLastError());
    return VSCB200_OK;
  }
  const int64_t seg_len = (n + nsplit - 1) / nsplit;
  nsplit = (n + seg_len - 1) / seg_len;
  float* Dp = nullptr;
  int64_t* Ip = nullptr;
  int rc = pool_alloc(reinterpret_cast<void**>(&Dp), static_cast<size_t>(nq) * nsplit * k * sizeof(float), stream);
  if (rc) return rc;
  if ((rc = pool_alloc(reinterpret_cast<void**>(&Ip), static_cast<size_t>(nq) * nsplit * k * sizeof(int64_t), stream))) {
    pool_free(Dp, stream);
    return rc;
  }
  topk_select_kernel<<<dim3(static_cast<unsigned>(nq), static_cast<unsigned>(nsplit)), kSelThreads,
                       kpad * sizeof(unsigned long long), stream>>>(S, ldS, es, n, seg_len, k, kpad, keep_max ? 1 : 0, Dp, Ip,
                                                                    id_offset, nullptr);
  count_launch();
  const int64_t n2 = nsplit * k;
  topk_select_kernel<<<static_cast<unsigned>(nq), kSelThreads, kpad * sizeof(unsigned long long), stream>>>(
      Dp, n2, 1, n2, n2, k, kpad, keep_max ? 1 : 0, D, I, 0, Ip);
  count_launch();
  pool_free(Dp, stream);
  pool_free(Ip, stream);
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

// ------------------------------------------------------------------ exact rescoring of the survivors
// The tensor-core score block is fp32-equivalent only to ~1e-6 (2-way bf16 split).  search() therefore
// selects k' = k + kRescoreSlack survivors per row on those scores, recomputes their scores here in plain
// fp32 from the original fp32 descriptors (the arithmetic faiss IndexFlat performs), re-sorts and keeps k:
// reported scores are fp32 dot products and the order can only differ from an fp32 brute force where two
// scores agree to fp32 rounding.
constexpr int kRescoreThreads = 256;

__global__ void __launch_bounds__(kRescoreThreads)
rescore_sort_kernel(const float* __restrict__ Q, const float* __restrict__ bank, int d, int l2,
                    const int64_t* __restrict__ Iin, int kin, int kpad, int k, float* __restrict__ D,
                    int64_t* __restrict__ I, int64_t id_offset) {
  extern __shared__ unsigned long long rs_smem[];        // [kpad] composite keys, then [d] floats
  unsigned long long* cand = rs_smem;
  float* sq = reinterpret_cast<float*>(rs_smem + kpad);
  float* sval = sq + d;                                   // [kin] exact scores
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t row = blockIdx.x;
  const bool keep_max = !l2;
  for (int c = tid; c < d; c += kRescoreThreads) sq[c] = Q[row * d + c];
  for (int i = tid; i < kpad; i += kRescoreThreads) cand[i] = 0ull;
  __syncthreads();
  // four survivors per warp per round: independent row reads, so their HBM latencies overlap
  for (int c0 = warp * 4; c0 < kin; c0 += (kRescoreThreads / 32) * 4) {
    int64_t ids[4];
    const float* rp[4];
    float acc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      ids[u] = (c0 + u < kin) ? Iin[row * kin + c0 + u] : -1;     // -1: padding (k > ntotal)
      rp[u] = bank + (ids[u] >= 0 ? ids[u] : 0) * d;
    }
    exact_rows_warp<4>(sq, rp, d, lane, l2 != 0, acc);            // the shared summation order (exact.cuh)
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (lane == 0 && ids[u] >= 0) {
        sval[c0 + u] = acc[u];
        // key 0 is reserved for empty slots: okey() of a finite float is never 0
        cand[c0 + u] = (static_cast<unsigned long long>(okey(acc[u], keep_max)) << 32) |
                       static_cast<uint32_t>(~static_cast<uint32_t>(ids[u]));
      }
    }
  }
  __syncthreads();
  for (int size = 2; size <= kpad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < (kpad >> 1); i += kRescoreThreads) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const unsigned long long a = cand[lo], b = cand[hi];
        if ((a < b) == desc) { cand[lo] = b; cand[hi] = a; }
      }
      __syncthreads();
    }
  }
  for (int j = tid; j < k; j += kRescoreThreads) {
    const unsigned long long c = cand[j];
    if (c != 0ull) {
      const uint32_t key = static_cast<uint32_t>(c >> 32);
      const uint32_t ok = keep_max ? key : ~key;
      const uint32_t u = (ok & 0x80000000u) ? (ok ^ 0x80000000u) : ~ok;     // inverse of okey()
      D[row * k + j] = __uint_as_float(u);
      I[row * k + j] = id_offset + static_cast<int64_t>(~static_cast<uint32_t>(c & 0xFFFFFFFFull));
    } else {
      D[row * k + j] = keep_max ? -FLT_MAX : FLT_MAX;
      I[row * k + j] = -1;
    }
  }
}

int rescore_sort(const float* Q, const float* bank, int d, bool l2, const int64_t* Iin, int kin, int64_t nq, int k,
                 float* D, int64_t* I, int64_t id_offset, cudaStream_t stream) {
  if (nq == 0) return VSCB200_OK;
  int kpad = 2;
  while (kpad < kin) kpad <<= 1;
  const size_t smem = kpad * sizeof(unsigned long long) + (static_cast<size_t>(d) + kin) * sizeof(float);
  VSCB_REQUIRE(smem <= 200 * 1024, "rescore: descriptor dimension too large");
  VSCB_CUDA_OK(cudaFuncSetAttribute(rescore_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  ProfScope prof(kProfSelect, stream, static_cast<double>(nq) * kin * d * 4);
  rescore_sort_kernel<<<static_cast<unsigned>(nq), kRescoreThreads, smem, stream>>>(Q, bank, d, l2 ? 1 : 0, Iin, kin, kpad, k,
                                                                                    D, I, id_offset);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

// ------------------------------------------------------------------ range search
// S holds tensor-core scores (fp32-equivalent to ~1e-6 relative to |q||r|).  Pairs whose score lies
// within that error of the threshold are decided on an exact fp32 recomputation, and every reported
// distance is the exact fp32 value, so the CSR result equals an fp32 brute force.  `exact == 0`
// (S already exact, SIMT path) skips both.
constexpr int kRangeThreads = 256;

struct RangeArgs {
  const float* S; int64_t ldS; int64_t n; float thr; int keep_max;
  int exact; const float* Q; const float* bank; int d; const float* qn; const float* rn;
};

__device__ __forceinline__ float exact_pair(const float* __restrict__ q, const float* __restrict__ r, int d, bool l2) {
  float acc = 0.f;
  for (int j = 0; j < d; ++j) {
    if (l2) { const float df = q[j] - r[j]; acc = fmaf(df, df, acc); }
    else acc = fmaf(q[j], r[j], acc);
  }
  return acc;
}

__device__ __forceinline__ bool range_hit(const RangeArgs& a, const float* sq, float qn, int64_t i, float v) {
  if (a.exact) {
    const float margin = 1.5e-5f * sqrtf(qn * a.rn[i]) * (a.keep_max ? 1.f : 2.f);
    if (fabsf(v - a.thr) <= margin) v = exact_pair(sq, a.bank + i * a.d, a.d, !a.keep_max);
  }
  return a.keep_max ? (v > a.thr) : (v < a.thr);
}

__global__ void __launch_bounds__(kRangeThreads)
range_count_kernel(RangeArgs a, unsigned long long* __restrict__ counts) {
  extern __shared__ float rg_q[];
  const int64_t row_id = blockIdx.x;
  const float* row = a.S + row_id * a.ldS;
  float qn = 0.f;
  if (a.exact) {
    for (int c = threadIdx.x; c < a.d; c += kRangeThreads) rg_q[c] = a.Q[row_id * a.d + c];
    qn = a.qn[row_id];
    __syncthreads();
  }
  unsigned int c = 0;
  for (int64_t i = threadIdx.x; i < a.n; i += kRangeThreads) c += range_hit(a, rg_q, qn, i, row[i]) ? 1u : 0u;
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
range_fill_kernel(RangeArgs a, const unsigned long long* __restrict__ offsets, float* __restrict__ D,
                  int64_t* __restrict__ I, int64_t id_offset) {
  extern __shared__ float rg_q[];
  __shared__ unsigned int wtot[kRangeThreads / 32];
  const int64_t row_id = blockIdx.x;
  const float* row = a.S + row_id * a.ldS;
  const unsigned long long pos0 = offsets[blockIdx.x];
  unsigned long long pos = pos0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float qn = 0.f;
  if (a.exact) {
    for (int c = threadIdx.x; c < a.d; c += kRangeThreads) rg_q[c] = a.Q[row_id * a.d + c];
    qn = a.qn[row_id];
    __syncthreads();
  }
  for (int64_t base = 0; base < a.n; base += kRangeThreads) {
    const int64_t i = base + threadIdx.x;
    float v = 0.f;
    bool hit = false;
    if (i < a.n) {
      v = row[i];
      hit = range_hit(a, rg_q, qn, i, v);
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
      I[o] = a.exact ? i : id_offset + i;
    }
    pos += total;
    __syncthreads();
  }
  if (a.exact) {
    // second phase: exact fp32 value of every hit of this row, one warp per hit (coalesced row reads)
    __threadfence_block();
    __syncthreads();
    const unsigned long long cnt = pos - pos0;
    for (unsigned long long j = warp; j < cnt; j += kRangeThreads / 32) {
      const int64_t id = I[pos0 + j];
      const float* const rp1[1] = {a.bank + id * a.d};
      float acc1[1];
      exact_rows_warp<1>(rg_q, rp1, a.d, lane, !a.keep_max, acc1);
      const float acc = acc1[0];
      if (lane == 0) {
        D[pos0 + j] = acc;
        I[pos0 + j] = id_offset + id;
      }
    }
  }
}

int range_count(const float* S, int64_t ldS, int64_t nq, int64_t n, float thr, bool keep_max,
                unsigned long long* counts, cudaStream_t stream, const float* Q, const float* bank, int d,
                const float* qn, const float* rn) {
  if (nq == 0) return VSCB200_OK;
  RangeArgs a{S, ldS, n, thr, keep_max ? 1 : 0, Q != nullptr ? 1 : 0, Q, bank, d, qn, rn};
  const size_t smem = a.exact ? static_cast<size_t>(d) * sizeof(float) : 0;
  VSCB_REQUIRE(smem <= 40 * 1024, "range_search: descriptor dimension too large");
  range_count_kernel<<<static_cast<unsigned>(nq), kRangeThreads, smem, stream>>>(a, counts);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

int range_fill(const float* S, int64_t ldS, int64_t nq, int64_t n, float thr, bool keep_max,
               const unsigned long long* offsets, float* D, int64_t* I, int64_t id_offset, cudaStream_t stream,
               const float* Q, const float* bank, int d, const float* qn, const float* rn) {
  if (nq == 0) return VSCB200_OK;
  RangeArgs a{S, ldS, n, thr, keep_max ? 1 : 0, Q != nullptr ? 1 : 0, Q, bank, d, qn, rn};
  const size_t smem = a.exact ? static_cast<size_t>(d) * sizeof(float) : 0;
  range_fill_kernel<<<static_cast<unsigned>(nq), kRangeThreads, smem, stream>>>(a, offsets, D, I, id_offset);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

}  // namespace vscb200
