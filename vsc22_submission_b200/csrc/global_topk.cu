// Global (cross-query) candidate search on the device: the `global_k` best (query row, bank row) pairs over ALL
// queries, and their reduction to (query video, reference video) candidates.
//
// Replaces, behind one C-ABI call each:
//   * vsc/index.py:142-165 `_global_threshold_knn_search` + vsc/exhaustive_search.py:206-292
//     `range_search_max_results` / :178-203 `apply_maxres` / :149-161 `threshold_radius_nres` -- the adaptive-radius
//     range search whose result handling is Python loops and numpy partitions on the host;
//   * vsc/index.py:119-140 (regrouping hits by video pair) + vsc/candidates.py:24-40 (MaxScoreAggregation, sort);
//   * M/infer/infer_matching.py:229-256 (per-video search + CPU range_search(thr) + dict max-reduce + sort).
//
// Algorithm (same idea as the reference's: keep <= 2K survivors, raise the radius when the table overflows):
//   per query block: dense tensor-core score block S (score_block) -> `gt_emit_kernel` appends every pair better
//   than the current radius (minus the tensor-core error margin) to a survivor buffer.  When the buffer would
//   overflow, an exact 3-level radix select (11+11+10 bits of the order-preserving key, shared-memory histograms)
//   over {buffer, block} finds the K-th best value, which becomes the new radius; the buffer is compacted and the
//   block emitted again.  At the end every survivor is rescored in exact fp32 (exact.cuh: the one summation order of
//   all search paths), sorted by (score, query row, bank row) with a stable LSD radix sort (sort.cu) and cut to K.
//   The result is therefore the exact fp32 global top-K, not an approximation of it.
//   Single-pass form (default when the bank has operand planes): no dense block at all.  The radius of a block comes
//   from a 1/64 COLUMN sample scored by a small one-pass GEMM; the block is then scored by ONE bf16 MMA per product
//   with the emission in the GEMM epilogue (sim_tc.cu emit mode), under the proven one-pass margin of margin.cuh; the
//   K-th best survivor verifies the estimate, and only when it fails (or no sample is possible) the dense block of
//   the first form is built.  Margins only decide what is rescored; the scores and the order stay exact.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "exact.cuh"
#include "host_util.h"
#include "index_internal.h"
#include "kernels.h"
#include "margin.cuh"

namespace vscb200 {
size_t radix_sort_scratch_bytes(int64_t n);
int radix_sort_pairs(uint64_t* k0, uint64_t* v0, uint64_t* k1, uint64_t* v1, int64_t n, int bit_lo, int bit_hi,
                     void* scratch, int* result_in_alt, cudaStream_t stream);
}  // namespace vscb200

using namespace vscb200;

namespace {

constexpr int kGtThreads = 256;
constexpr int kGtChunk = 4096;            // columns of one row handled per work item
constexpr float kTcMargin = 1.5e-5f;      // tensor-core score error bound relative to |q||r| (select.cu range_hit)

// order-preserving key, larger = better (IP: larger score; L2: smaller distance)
__host__ __device__ __forceinline__ uint32_t better_key(float v, bool keep_max) {
#ifdef __CUDA_ARCH__
  uint32_t b = __float_as_uint(v);
#else
  uint32_t b;
  memcpy(&b, &v, 4);
#endif
  b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  return keep_max ? b : ~b;
}
__host__ __device__ __forceinline__ float key_value(uint32_t key, bool keep_max) {
  uint32_t b = keep_max ? key : ~key;
  b = (b & 0x80000000u) ? (b & 0x7fffffffu) : ~b;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  float f;
  memcpy(&f, &b, 4);
  return f;
#endif
}

struct GtBlock {
  const float* S; int64_t ldS; int64_t nb; int64_t n;     // score block [nb, n]
  int64_t q0;                                             // global query row of block row 0
  const float* marg;                                      // [nq_total] two-sided scoring-error margin of a query row's scores
  float radius; int has_radius; int keep_max;
};

// a pair survives when it is better than the radius by more than the scoring error could hide
__device__ __forceinline__ float row_threshold(const GtBlock& a, int64_t row) {
  if (!a.has_radius) return a.keep_max ? -INFINITY : INFINITY;
  const float m = a.marg[a.q0 + row];
  return a.keep_max ? a.radius - m : a.radius + m;
}
__device__ __forceinline__ bool better(float v, float thr, int keep_max) { return keep_max ? v > thr : v < thr; }

// ---- radix-select histograms over {block entries better than the radius} U {buffer} --------------------------
// level 0: digit = key >> 21; level 1: (key >> 21) == prefix, digit = (key >> 10) & 2047;
// level 2: (key >> 10) == prefix, digit = key & 1023
__device__ __forceinline__ void hist_add(uint32_t* h, uint32_t key, int level, uint32_t prefix, bool pass) {
  uint32_t digit;
  if (level == 0) digit = key >> 21;
  else if (level == 1) { pass = pass && (key >> 21) == prefix; digit = (key >> 10) & 2047u; }
  else { pass = pass && (key >> 10) == prefix; digit = key & 1023u; }
  const uint32_t mask = __ballot_sync(0xffffffffu, pass);
  if (pass) {
    const uint32_t peers = __match_any_sync(mask, digit);
    if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&h[digit], __popc(peers));
  }
}

__global__ void __launch_bounds__(kGtThreads)
gt_hist_kernel(GtBlock a, const float* __restrict__ bufv, int64_t nbuf, int level, const unsigned long long* __restrict__ sel,
               unsigned long long* __restrict__ hist) {
  const uint32_t prefix = level ? static_cast<uint32_t>(sel[0]) : 0u;      // key bits fixed by the levels before (gt_pick_kernel)
  __shared__ uint32_t h[2048];
  for (int i = threadIdx.x; i < 2048; i += kGtThreads) h[i] = 0;
  __syncthreads();
  const int64_t chunks = (a.n + kGtChunk - 1) / kGtChunk;
  const int64_t work = a.nb * chunks;
  for (int64_t w = blockIdx.x; w < work; w += gridDim.x) {
    const int64_t row = w / chunks, c0 = (w % chunks) * kGtChunk;
    const float thr = row_threshold(a, row);
    const float* p = a.S + row * a.ldS + c0;
    const int cnt = a.n - c0 < kGtChunk ? static_cast<int>(a.n - c0) : kGtChunk;
    for (int j0 = 0; j0 < cnt; j0 += kGtThreads * 4) {
      const int j = j0 + threadIdx.x * 4;
      float v[4];
      bool ok[4];
      if (j + 3 < cnt) {
        const float4 t = *reinterpret_cast<const float4*>(p + j);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        ok[0] = ok[1] = ok[2] = ok[3] = true;
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) { ok[u] = j + u < cnt; v[u] = ok[u] ? p[j + u] : 0.f; }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        hist_add(h, better_key(v[u], a.keep_max), level, prefix, ok[u] && better(v[u], thr, a.keep_max));
    }
  }
  for (int64_t i0 = static_cast<int64_t>(blockIdx.x) * kGtThreads; i0 < nbuf; i0 += static_cast<int64_t>(gridDim.x) * kGtThreads) {
    const int64_t i = i0 + threadIdx.x;
    const bool ok = i < nbuf;
    hist_add(h, better_key(ok ? bufv[i] : 0.f, a.keep_max), level, prefix, ok);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2048; i += kGtThreads)
    if (h[i]) atomicAdd(&hist[i], static_cast<unsigned long long>(h[i]));
}

// One level of the radix select on the device: the digit in which the `want`-th best key falls (scanning the histogram
// from the best digit down), appended to the prefix; `want` becomes the rank inside that digit.  sel = {prefix, want}.
// One warp: lane L owns nd / 32 consecutive digits, best first.  (The host used to fetch every histogram: three round
// trips per select.)
__global__ void gt_pick_kernel(const unsigned long long* __restrict__ hist, int level, unsigned long long* __restrict__ sel) {
  const int lane = threadIdx.x;
  const int nd = level == 2 ? 1024 : 2048, per = nd / 32;
  const unsigned long long want = sel[1];
  const int top = nd - 1 - lane * per;                       // this lane's best digit
  unsigned long long mine = 0ull;
  for (int i = 0; i < per; ++i) mine += hist[top - i];
  unsigned long long incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  const uint32_t cross = __ballot_sync(0xffffffffu, incl >= want);
  const int owner = cross ? __ffs(cross) - 1 : 31;           // no digit reaches the rank: the scan ends at digit 0
  if (lane == owner) {
    unsigned long long cum = incl - mine;
    int dgt = top;
    for (; dgt > 0 && dgt > top - per; --dgt) {
      if (cum + hist[dgt] >= want) break;
      cum += hist[dgt];
    }
    if (dgt == top - per) dgt = top - per + 1;               // (only lane 31 without a crossing: stop at its last digit, 0)
    const unsigned long long prefix = sel[0];
    sel[0] = level == 2 ? ((prefix << 10) | static_cast<unsigned long long>(dgt)) : ((prefix << 11) | static_cast<unsigned long long>(dgt));
    sel[1] = want - cum;
  }
}

// ---- emit: append the survivors of one score block -----------------------------------------------------------
// One atomicAdd on the global counter per CTA and work item (4096 scores): under a loose bootstrap radius a block emits
// ~10^6 survivors, and per-warp atomics on the one counter address (~1 per ns) were the whole cost of the pass.
// The counter keeps counting past `cap`, so the host sees by how much a block overflowed.
__global__ void __launch_bounds__(kGtThreads)
gt_emit_kernel(GtBlock a, int64_t ntotal, float* __restrict__ bufv, uint64_t* __restrict__ bufp,
               unsigned long long* __restrict__ counter, unsigned long long cap) {
  __shared__ int warp_cnt[kGtThreads / 32];
  __shared__ unsigned long long cta_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t chunks = (a.n + kGtChunk - 1) / kGtChunk;
  const int64_t work = a.nb * chunks;
  for (int64_t w = blockIdx.x; w < work; w += gridDim.x) {
    const int64_t row = w / chunks, c0 = (w % chunks) * kGtChunk;
    const float thr = row_threshold(a, row);
    const float* p = a.S + row * a.ldS + c0;
    const uint64_t pbase = static_cast<uint64_t>(a.q0 + row) * static_cast<uint64_t>(ntotal) + static_cast<uint64_t>(c0);
    const int cnt = a.n - c0 < kGtChunk ? static_cast<int>(a.n - c0) : kGtChunk;
    // one work item = 4096 columns = 4 float4 per thread: all loads are issued before any is consumed
    float v[16];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = (u * kGtThreads + threadIdx.x) * 4;
      if (j + 3 < cnt) {
        const float4 t = *reinterpret_cast<const float4*>(p + j);
        v[4 * u] = t.x; v[4 * u + 1] = t.y; v[4 * u + 2] = t.z; v[4 * u + 3] = t.w;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[4 * u + e] = j + e < cnt ? p[j + e] : (a.keep_max ? -INFINITY : INFINITY);
      }
    }
    uint32_t hits = 0;                                   // bit 4u+e: element e of load u survives
#pragma unroll
    for (int e = 0; e < 16; ++e) hits |= better(v[e], thr, a.keep_max) ? (1u << e) : 0u;
    const int mine = __popc(hits);
    int incl = mine;                                     // warp inclusive scan of the per-thread survivor counts
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) warp_cnt[warp] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
      int total = 0;
#pragma unroll
      for (int i = 0; i < kGtThreads / 32; ++i) { const int c = warp_cnt[i]; warp_cnt[i] = total; total += c; }
      cta_base = total ? atomicAdd(counter, static_cast<unsigned long long>(total)) : 0ull;
    }
    __syncthreads();
    if (mine) {
      unsigned long long pos = cta_base + static_cast<unsigned long long>(warp_cnt[warp] + incl - mine);
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        if (hits & (1u << e)) {
          const int j = ((e >> 2) * kGtThreads + threadIdx.x) * 4 + (e & 3);
          if (pos < cap) { bufv[pos] = v[e]; bufp[pos] = pbase + static_cast<uint64_t>(j); }
          ++pos;
        }
      }
    }
    __syncthreads();                                     // warp_cnt / cta_base are reused by the next work item
  }
}

// ---- bootstrap sample: 1/64 of the block's scores for a first radius estimate -- runs of 64 consecutive scores (256 B:
// whole sectors) out of every 4096, the run's position inside its 4096 rotated from row to row
constexpr int kGtSampleStride = 64;
constexpr int kGtSampleRun = 64;
__global__ void __launch_bounds__(kGtThreads)
gt_sample_kernel(const float* __restrict__ S, int64_t ldS, int64_t nb, int64_t n, int64_t per_row, int keep_max,
                 float* __restrict__ out) {
  const int64_t total = nb * per_row;
  constexpr int64_t span = static_cast<int64_t>(kGtSampleStride) * kGtSampleRun;      // 4096
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kGtThreads + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * kGtThreads) {
    const int64_t row = i / per_row, t = i % per_row;
    const int64_t col = (t / kGtSampleRun) * span + (row * kGtSampleRun) % span + t % kGtSampleRun;
    out[i] = col < n ? S[row * ldS + col] : (keep_max ? -INFINITY : INFINITY);
  }
}

// ---- compaction of the survivor buffer against a raised radius -----------------------------------------------
__global__ void __launch_bounds__(kGtThreads)
gt_compact_kernel(const float* __restrict__ inv, const uint64_t* __restrict__ inp, int64_t n, int64_t ntotal,
                  const float* __restrict__ marg, float radius, int keep_max,
                  float* __restrict__ outv, uint64_t* __restrict__ outp, unsigned long long* __restrict__ counter) {
  const int lane = threadIdx.x & 31;
  for (int64_t i0 = static_cast<int64_t>(blockIdx.x) * kGtThreads; i0 < n; i0 += static_cast<int64_t>(gridDim.x) * kGtThreads) {
    const int64_t i = i0 + threadIdx.x;
    bool keep = false;
    float v = 0.f;
    uint64_t p = 0;
    if (i < n) {
      v = inv[i];
      p = inp[i];
      const float m = marg[p / static_cast<uint64_t>(ntotal)];
      keep = keep_max ? v > radius - m : v < radius + m;
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, keep);
    if (!bal) continue;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(counter, static_cast<unsigned long long>(__popc(bal)));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (keep) {
      const unsigned long long pos = base + __popc(bal & ((1u << lane) - 1u));
      outv[pos] = v;
      outp[pos] = p;
    }
  }
}

__global__ void __launch_bounds__(1024)
max_reduce_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
  __shared__ float ws[32];
  float m = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += 1024) m = fmaxf(m, x[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = ws[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) *out = m;
  }
}

// scoring-error bound per query row.  one_pass: margin.cuh; else the split-bf16 bound relative to |q||r|
__global__ void __launch_bounds__(kGtThreads)
gt_eps_kernel(const float* __restrict__ qn, const float* __restrict__ qn_lo, int64_t nq, const float* __restrict__ rn_max,
              const unsigned int* __restrict__ bank_max_bits, int d, int keep_max, int one_pass, float* __restrict__ eps) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kGtThreads + threadIdx.x;
  if (i >= nq) return;
  eps[i] = one_pass ? sim1_eps(qn[i], qn_lo[i], bank_max_bits, d, keep_max ? 0 : 1)
                    : kTcMargin * sqrtf(qn[i] * *rn_max) * (keep_max ? 1.f : 2.f);
}
__global__ void __launch_bounds__(kGtThreads)
gt_margin_kernel(float* __restrict__ eps_to_marg, int64_t nq, const float* __restrict__ eps_max) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kGtThreads + threadIdx.x;
  if (i < nq) eps_to_marg[i] = (eps_to_marg[i] + *eps_max) * 1.0001f;
}

// ---- exact rescoring + sort keys ---------------------------------------------------------------------------
// one warp per survivor; key_p = pair id (tie order), key_s = order key of the exact score, ascending = best first.
// fixed_thr: pairs that fail the caller's threshold on the exact score get the worst key and are counted.
__global__ void __launch_bounds__(kGtThreads)
gt_rescore_kernel(const uint64_t* __restrict__ bufp, int64_t n, const float* __restrict__ Q, const float* __restrict__ bank,
                  int d, int64_t ntotal, int keep_max, int use_thresh, float thresh, uint64_t* __restrict__ key_p,
                  uint64_t* __restrict__ key_s, unsigned long long* __restrict__ n_fail) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * kGtThreads + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * kGtThreads) >> 5;
  for (int64_t i = warp0; i < n; i += nwarps) {
    const uint64_t p = bufp[i];
    const uint64_t qi = p / static_cast<uint64_t>(ntotal), ri = p % static_cast<uint64_t>(ntotal);
    const float* const rp[1] = {bank + ri * d};
    float acc[1];
    exact_rows_warp<1>(Q + qi * d, rp, d, lane, !keep_max, acc);
    if (lane == 0) {
      uint32_t k = ~better_key(acc[0], keep_max);       // ascending = best first
      if (use_thresh && !(keep_max ? acc[0] > thresh : acc[0] < thresh)) {
        k = 0xffffffffu;
        atomicAdd(n_fail, 1ull);
      }
      key_p[i] = p;
      key_s[i] = k;
    }
  }
}

__global__ void __launch_bounds__(kGtThreads)
gt_output_kernel(const uint64_t* __restrict__ key_s, const uint64_t* __restrict__ key_p, int64_t n, int64_t ntotal,
                 int64_t id_offset, int keep_max, float* __restrict__ score, int64_t* __restrict__ qrow,
                 int64_t* __restrict__ brow) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kGtThreads + threadIdx.x;
  if (i >= n) return;
  const uint64_t p = key_p[i];
  score[i] = key_value(~static_cast<uint32_t>(key_s[i]), keep_max);
  qrow[i] = static_cast<int64_t>(p / static_cast<uint64_t>(ntotal));
  brow[i] = static_cast<int64_t>(p % static_cast<uint64_t>(ntotal)) + id_offset;
}

// ---- video-pair reduction -----------------------------------------------------------------------------------
__device__ __forceinline__ int64_t segment_of(const int64_t* __restrict__ off, int64_t nseg, int64_t row) {
  int64_t lo = 0, hi = nseg;             // largest s with off[s] <= row
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (off[mid] <= row) lo = mid; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}

// the input is sorted best first, so the smallest index of a (query video, ref video) key is its best pair
__global__ void __launch_bounds__(kGtThreads)
vp_insert_kernel(const int64_t* __restrict__ qrow, const int64_t* __restrict__ brow, int64_t n, int64_t id_offset,
                 const int64_t* __restrict__ q_off, int64_t nqv, const int64_t* __restrict__ r_off, int64_t nrv,
                 unsigned long long* __restrict__ tkey, uint32_t* __restrict__ tmin, uint64_t mask) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kGtThreads + threadIdx.x;
  if (i >= n) return;
  const int64_t qv = segment_of(q_off, nqv, qrow[i]);
  const int64_t rv = segment_of(r_off, nrv, brow[i] - id_offset);
  const unsigned long long key = static_cast<unsigned long long>(qv) * static_cast<unsigned long long>(nrv) +
                                 static_cast<unsigned long long>(rv);
  uint64_t slot = mix64(key) & mask;
  while (true) {
    const unsigned long long prev = atomicCAS(&tkey[slot], ~0ull, key);
    if (prev == ~0ull || prev == key) {
      atomicMin(&tmin[slot], static_cast<uint32_t>(i));
      return;
    }
    slot = (slot + 1) & mask;
  }
}

__global__ void __launch_bounds__(kGtThreads)
vp_extract_kernel(const unsigned long long* __restrict__ tkey, const uint32_t* __restrict__ tmin, uint64_t slots,
                  uint64_t* __restrict__ first_idx, uint64_t* __restrict__ keys, unsigned long long* __restrict__ counter) {
  const uint64_t s = static_cast<uint64_t>(blockIdx.x) * kGtThreads + threadIdx.x;
  if (s >= slots || tkey[s] == ~0ull) return;
  const unsigned long long pos = atomicAdd(counter, 1ull);
  first_idx[pos] = tmin[s];
  keys[pos] = tkey[s];
}

__global__ void __launch_bounds__(kGtThreads)
vp_output_kernel(const uint64_t* __restrict__ first_idx, const uint64_t* __restrict__ keys, int64_t m, int64_t nrv,
                 const float* __restrict__ g_score, float* __restrict__ score, int64_t* __restrict__ qv,
                 int64_t* __restrict__ rv) {
  const int64_t j = static_cast<int64_t>(blockIdx.x) * kGtThreads + threadIdx.x;
  if (j >= m) return;
  score[j] = g_score[first_idx[j]];
  qv[j] = static_cast<int64_t>(keys[j] / static_cast<uint64_t>(nrv));
  rv[j] = static_cast<int64_t>(keys[j] % static_cast<uint64_t>(nrv));
}

int bits_for(uint64_t n) {   // bits needed to represent values < n
  int b = 0;
  while (b < 64 && (n > (1ull << b))) ++b;
  return std::max(b, 1);
}

struct Scratch {               // pool blocks released on every exit path
  std::vector<void*> blocks;
  cudaStream_t s;
  explicit Scratch(cudaStream_t st) : s(st) {}
  ~Scratch() { for (void* b : blocks) pool_free(b, s); }
  template <typename Tp>
  int get(Tp** p, size_t bytes) {
    void* v = nullptr;
    int rc = pool_alloc(&v, std::max<size_t>(bytes, 256), s);
    if (rc) return rc;
    blocks.push_back(v);
    *p = static_cast<Tp*>(v);
    return VSCB200_OK;
  }
  void drop(void* p) {
    if (!p) return;
    for (auto& b : blocks) if (b == p) { pool_free(b, s); b = nullptr; }
  }
};

int grid_for(int64_t work) {
  const int64_t g = std::min<int64_t>(std::max<int64_t>(work, 1), static_cast<int64_t>(device_sm_count()) * 8);
  return static_cast<int>(g);
}

}  // namespace

extern "C" {

int vscb200_index_global_search(vscb200_index* ix, const float* q, int64_t nq, int64_t global_k, int use_thresh,
                                float thresh, int64_t* n_found, void* stream) {
  VSCB_REQUIRE(ix && n_found && (nq == 0 || q), "index_global_search: null argument");
  VSCB_REQUIRE(nq >= 0, "index_global_search: negative query count");
  VSCB_REQUIRE(global_k > 0 || use_thresh, "index_global_search: need global_k > 0 or a threshold");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc = flush_pending(ix, s);
  if (rc) return rc;
  *n_found = 0;
  ix->g_n = 0;
  ix->vp_n = 0;
  const int64_t n = ix->ntotal;
  if (nq == 0 || n == 0) return VSCB200_OK;
  VSCB_REQUIRE(static_cast<double>(nq) * static_cast<double>(n) < 9.0e18, "index_global_search: nq * ntotal overflows");
  const bool keep_max = ix->metric == VSCB200_METRIC_INNER_PRODUCT;
  const bool limited = global_k > 0;
  const uint64_t total_pairs = static_cast<uint64_t>(nq) * static_cast<uint64_t>(n);
  const uint64_t K = limited ? std::min<uint64_t>(static_cast<uint64_t>(global_k), total_pairs) : total_pairs;

  Scratch sc(s);
  // One-pass form: the bank has bf16 operand planes and their norm maxima (index.cu add()).  VSCB200_GT_DENSE=1 forces the
  // dense-block form.
  static const int force_dense = [] { const char* e = getenv("VSCB200_GT_DENSE"); return e ? atoi(e) : 0; }();
  const bool one_pass = !force_dense && !ix->force_simt && ix->bank_hi != nullptr && ix->rmax2_bits != nullptr;
  static const int trace = [] { const char* e = getenv("VSCB200_GT_TRACE"); return e ? atoi(e) : 0; }();
#define GT_TRACE(...) do { if (trace) fprintf(stderr, __VA_ARGS__); } while (0)
  GT_TRACE("[gt] nq=%lld n=%lld K=%llu one_pass=%d\n", (long long)nq, (long long)n, (unsigned long long)K, one_pass ? 1 : 0);
  // squared norms of ALL query rows, the largest bank norm, and the scoring-error margin of every query row.  A true
  // top-K pair of row i scores >= (approximate K-th best) - eps_max - eps_i: the K-th best may sit in the row with the
  // largest error.
  float *qn_all = nullptr, *qlo_all = nullptr, *rn_max = nullptr, *marg = nullptr, *eps_max = nullptr;
  uint16_t* qh_all = nullptr;          // one-pass form: bf16 plane of all query rows
  if ((rc = sc.get(&qn_all, static_cast<size_t>(nq) * sizeof(float)))) return rc;
  if ((rc = sc.get(&marg, static_cast<size_t>(nq) * sizeof(float)))) return rc;
  if ((rc = sc.get(&rn_max, sizeof(float))) || (rc = sc.get(&eps_max, sizeof(float)))) return rc;
  if (one_pass) {
    if ((rc = sc.get(&qlo_all, static_cast<size_t>(nq) * sizeof(float)))) return rc;
    if ((rc = sc.get(&qh_all, static_cast<size_t>(nq) * ix->dp * sizeof(uint16_t)))) return rc;
    if ((rc = q_hi_norm(q, qh_all, qn_all, qlo_all, nq, ix->d, ix->dp, s))) return rc;
  } else {
    if ((rc = row_sqnorm(q, nq, ix->d, qn_all, s))) return rc;
  }
  max_reduce_kernel<<<1, 1024, 0, s>>>(ix->rnorm, n, rn_max);
  gt_eps_kernel<<<static_cast<unsigned>((nq + kGtThreads - 1) / kGtThreads), kGtThreads, 0, s>>>(
      qn_all, qlo_all, nq, rn_max, ix->rmax2_bits, ix->d, keep_max ? 1 : 0, one_pass ? 1 : 0, marg);
  max_reduce_kernel<<<1, 1024, 0, s>>>(marg, nq, eps_max);
  gt_margin_kernel<<<static_cast<unsigned>((nq + kGtThreads - 1) / kGtThreads), kGtThreads, 0, s>>>(marg, nq, eps_max);
  count_launch(4);

  // survivor buffers (ping-pong for compaction).  Logical bound 2K like the reference's max_results: above it the
  // radius is raised to the K-th best survivor; the physical capacity is larger so that a block emitted under a loose
  // (bootstrap) radius still fits.
  const uint64_t logical_cap = limited ? (K > (1ull << 60) ? total_pairs : std::min<uint64_t>(2 * K + 4096, total_pairs)) : ~0ull;
  uint64_t cap = limited ? std::min<uint64_t>(total_pairs, std::max<uint64_t>(logical_cap, (K > (1ull << 32) ? (1ull << 27) : std::min<uint64_t>(16 * K + (1u << 20), 1u << 27))))
                         : std::min<uint64_t>(total_pairs, 1u << 20);
  float* bufv[2] = {nullptr, nullptr};
  uint64_t* bufp[2] = {nullptr, nullptr};
  auto alloc_bufs = [&](uint64_t c, float** v, uint64_t** p) {
    int r = sc.get(v, c * sizeof(float));
    if (r) return r;
    return sc.get(p, c * sizeof(uint64_t));
  };
  if ((rc = alloc_bufs(cap, &bufv[0], &bufp[0]))) return rc;
  unsigned long long *counter = nullptr, *hist = nullptr, *sel = nullptr;
  unsigned long long sel_host[2] = {0ull, 0ull};        // outlives every copy: each select ends with a synchronise
  if ((rc = sc.get(&sel, 2 * sizeof(unsigned long long)))) return rc;
  if ((rc = sc.get(&counter, sizeof(unsigned long long)))) return rc;
  if ((rc = sc.get(&hist, 2048 * sizeof(unsigned long long)))) return rc;
  VSCB_CUDA_OK(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), s));

  const int64_t blk = block_rows(ix, nq);
  const int64_t ldS = (n + 3) & ~3ll;
  const int64_t per_row = ((n + kGtSampleStride * kGtSampleRun - 1) / (kGtSampleStride * kGtSampleRun)) * kGtSampleRun;
  float* sample = nullptr;

  unsigned long long count = 0;     // host mirror of the buffer fill
  float radius = use_thresh ? thresh : 0.f;
  bool has_radius = use_thresh != 0;

  auto read_counter = [&](unsigned long long* out) {
    VSCB_CUDA_OK(cudaMemcpyAsync(out, counter, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    VSCB_CUDA_OK(cudaStreamSynchronize(s));
    return static_cast<int>(VSCB200_OK);
  };
  auto set_counter = [&](unsigned long long v) {
    if (v == 0) {                                  // the common case (compaction, rescoring): no host value, no round trip
      VSCB_CUDA_OK(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), s));
      return static_cast<int>(VSCB200_OK);
    }
    VSCB_CUDA_OK(cudaMemcpyAsync(counter, &v, sizeof(unsigned long long), cudaMemcpyHostToDevice, s));
    VSCB_CUDA_OK(cudaStreamSynchronize(s));      // v lives on this frame
    return static_cast<int>(VSCB200_OK);
  };
  // exact `want`-th best value over {entries of block `a` better than its radius} U {vals[0:nvals)}: 3-level radix select
  auto select_kth = [&](const GtBlock& a, const float* vals, int64_t nvals, uint64_t want, float* kth) {
    const int64_t work = a.nb * ((a.n + kGtChunk - 1) / kGtChunk);
    const int grid = grid_for(std::max<int64_t>(work, (nvals + kGtThreads - 1) / kGtThreads));
    // {prefix, want} live on the device; every level's digit is picked there (gt_pick_kernel): one host round trip per
    // select instead of three histogram downloads
    sel_host[0] = 0ull;
    sel_host[1] = want;
    VSCB_CUDA_OK(cudaMemcpyAsync(sel, sel_host, 2 * sizeof(unsigned long long), cudaMemcpyHostToDevice, s));
    for (int level = 0; level < 3; ++level) {
      VSCB_CUDA_OK(cudaMemsetAsync(hist, 0, 2048 * sizeof(unsigned long long), s));
      gt_hist_kernel<<<grid, kGtThreads, 0, s>>>(a, vals, nvals, level, sel, hist);
      gt_pick_kernel<<<1, 32, 0, s>>>(hist, level, sel);
      count_launch(2);
    }
    VSCB_CUDA_OK(cudaMemcpyAsync(sel_host, sel, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    VSCB_CUDA_OK(cudaStreamSynchronize(s));
    *kth = key_value(static_cast<uint32_t>(sel_host[0]), keep_max);
    return static_cast<int>(VSCB200_OK);
  };
  // keep the buffer entries that can still be among the K best under `radius`; updates count
  auto compact = [&]() {
    if (!bufv[1] && alloc_bufs(cap, &bufv[1], &bufp[1])) return static_cast<int>(VSCB200_ERR_NOMEM);
    int r = set_counter(0);
    if (r) return r;
    if (count) {
      gt_compact_kernel<<<grid_for((static_cast<int64_t>(count) + kGtThreads - 1) / kGtThreads), kGtThreads, 0, s>>>(
          bufv[0], bufp[0], static_cast<int64_t>(count), n, marg, radius, keep_max ? 1 : 0, bufv[1], bufp[1], counter);
      count_launch();
      std::swap(bufv[0], bufv[1]);
      std::swap(bufp[0], bufp[1]);
    }
    return read_counter(&count);
  };
  auto raise_radius = [&](float kth) {
    if (!has_radius || (keep_max ? kth > radius : kth < radius)) radius = kth;
    has_radius = true;
  };
  const GtBlock no_block{nullptr, 0, 0, n, 0, marg, 0.f, 0, keep_max ? 1 : 0};

  for (int64_t q0 = 0; q0 < nq; q0 += blk) {
    const int64_t nb = std::min(blk, nq - q0);
    GtBlock a{nullptr, ldS, nb, n, q0, marg, radius, has_radius ? 1 : 0, keep_max ? 1 : 0};
    // the dense score block (split-bf16, fp32-equivalent): always in the dense form, on demand in the one-pass form
    auto ensure_dense = [&]() -> int {
      if (a.S) return VSCB200_OK;
      GT_TRACE("[gt]   block q0=%lld: dense score block\n", (long long)q0);
      int r = grow(&ix->ws, &ix->ws_bytes, static_cast<size_t>(blk) * std::max<int64_t>(ldS, 4) * sizeof(float), s);
      if (r) return r;
      if ((r = score_block(ix, q + q0 * ix->d, nb, ix->ws, ldS, s))) return r;
      a.S = ix->ws;
      return VSCB200_OK;
    };
    if (!one_pass && (rc = ensure_dense())) return rc;
    const int64_t work = nb * ((n + kGtChunk - 1) / kGtChunk);
    const int grid = grid_for(work);
    // append the block's survivors under a.radius: from the dense block when there is one, else in the epilogue of a
    // one-pass GEMM
    auto emit = [&]() -> int {
      if (a.S) {
        gt_emit_kernel<<<grid, kGtThreads, 0, s>>>(a, n, bufv[0], bufp[0], counter, cap);
        count_launch();
        return VSCB200_OK;
      }
      return scores_tc_emit(qh_all + static_cast<size_t>(q0) * ix->dp, ix->bank_hi, nb, n, ix->dp, !keep_max, qn_all + q0, ix->rnorm,
                            marg + q0, a.radius, a.has_radius != 0, q0, bufv[0], bufp[0], counter, cap, s);
    };
    const uint64_t blk_pairs = static_cast<uint64_t>(nb) * static_cast<uint64_t>(n);
    unsigned long long after = 0;
    bool emitted = false, tentative = false;
    float r_hat = 0.f;
    if (!has_radius && count + blk_pairs > cap) {
      // No radius yet and the block does not fit: estimate one from a 1/64 sample of the block -- the value that about
      // 4K entries of the block should exceed -- emit under it, then VERIFY below (the K-th best survivor must not be
      // worse than the estimate, otherwise pairs between the two were never emitted and the exact select runs instead).
      const uint64_t want_s = std::max<uint64_t>((4 * K) / kGtSampleStride, 16);     // >= ~1000 pairs even for a tiny K
      const int64_t ncs = (n / kGtSampleStride) & ~3ll;                                // one-pass form: sampled bank rows
      if (one_pass && keep_max && ncs >= 4 && static_cast<uint64_t>(nb * ncs) > 4 * want_s) {
        // every 64th bank row against the block's queries: a [nb, ncs] one-pass GEMM (the strided rows come from the
        // tensor map, nothing is gathered)
        const int64_t ns = nb * ncs;
        if (!sample && (rc = sc.get(&sample, static_cast<size_t>(blk) * std::max(per_row, ncs) * sizeof(float)))) return rc;
        if ((rc = scores_tc_planes(qh_all + static_cast<size_t>(q0) * ix->dp, nullptr, ix->bank_hi, nullptr, sample, nb, ncs, ix->dp, ncs,
                                   false, nullptr, nullptr, s, 1, kGtSampleStride))) return rc;
        if ((rc = select_kth(no_block, sample, ns, want_s, &r_hat))) return rc;
        a.radius = r_hat;
        a.has_radius = 1;
        tentative = true;
      } else {
        if ((rc = ensure_dense())) return rc;
        const int64_t ns = nb * per_row;
        if (static_cast<uint64_t>(ns) > 4 * want_s) {
          if (!sample && (rc = sc.get(&sample, static_cast<size_t>(blk) * std::max(per_row, ncs) * sizeof(float)))) return rc;
          gt_sample_kernel<<<grid_for((ns + kGtThreads - 1) / kGtThreads), kGtThreads, 0, s>>>(ix->ws, ldS, nb, n, per_row,
                                                                                            keep_max ? 1 : 0, sample);
          count_launch();
          if ((rc = select_kth(no_block, sample, ns, want_s, &r_hat))) return rc;
          a.radius = r_hat;
          a.has_radius = 1;
          tentative = true;
        }
      }
    }
    if (a.has_radius || count + blk_pairs <= cap) {
      if (!a.has_radius && (rc = ensure_dense())) return rc;      // everything is emitted: no point in a threshold epilogue
      if ((rc = emit())) return rc;
      if ((rc = read_counter(&after))) return rc;
      emitted = after <= cap;
      if (emitted && tentative) {
        float kth = 0.f;
        const bool enough = after >= K;
        if (enough && (rc = select_kth(no_block, bufv[0], static_cast<int64_t>(after), K, &kth))) return rc;
        GT_TRACE("[gt]   block q0=%lld: bootstrap radius %.6f emitted %llu, K-th %.6f -> %s\n", (long long)q0, r_hat, after, kth,
                 (enough && (keep_max ? kth >= r_hat : kth <= r_hat)) ? "verified" : "rejected");
        if (enough && (keep_max ? kth >= r_hat : kth <= r_hat)) {
          count = after;                     // every pair at least as good as kth was emitted: kth is exact
          raise_radius(kth);
          if ((rc = compact())) return rc;
          after = count;
        } else {
          emitted = false;                   // estimate too tight: forget this emission, take the exact path
          if ((rc = set_counter(count))) return rc;
          a.radius = radius;
          a.has_radius = has_radius ? 1 : 0;
          after = count + blk_pairs;
        }
      }
    } else {
      after = count + blk_pairs;             // no radius at all: everything would be emitted
    }
    if (!emitted && limited) {
      // raise the radius: exact K-th best value over {buffer, block entries better than the radius}
      float kth = 0.f;
      if ((rc = ensure_dense())) return rc;  // the select below walks the block's scores
      if ((rc = select_kth(a, bufv[0], static_cast<int64_t>(count), K, &kth))) return rc;
      raise_radius(kth);
      a.radius = radius;
      a.has_radius = 1;
      if ((rc = compact())) return rc;       // what the buffer held before this block
      if ((rc = emit())) return rc;
      if ((rc = read_counter(&after))) return rc;
      emitted = after <= cap;
    }
    while (!emitted) {
      // unlimited (threshold) mode, or more ties at the radius than the buffer holds: grow and emit again
      const uint64_t ncap = std::min<uint64_t>(total_pairs, std::max<uint64_t>(after + after / 4, cap * 2));
      float* nv = nullptr;
      uint64_t* np = nullptr;
      if ((rc = alloc_bufs(ncap, &nv, &np))) return rc;
      if (count) {
        VSCB_CUDA_OK(cudaMemcpyAsync(nv, bufv[0], count * sizeof(float), cudaMemcpyDeviceToDevice, s));
        VSCB_CUDA_OK(cudaMemcpyAsync(np, bufp[0], count * sizeof(uint64_t), cudaMemcpyDeviceToDevice, s));
      }
      sc.drop(bufv[0]); sc.drop(bufp[0]); sc.drop(bufv[1]); sc.drop(bufp[1]);
      bufv[0] = nv; bufp[0] = np; bufv[1] = nullptr; bufp[1] = nullptr;
      cap = ncap;
      if ((rc = set_counter(count))) return rc;
      if ((rc = emit())) return rc;
      if ((rc = read_counter(&after))) return rc;
      emitted = after <= cap;
    }
    count = after;
    if (limited && count > logical_cap) {
      // more than 2K survivors: the K-th best of them is the new radius (apply_maxres, exhaustive_search.py:178-203)
      float kth = 0.f;
      if ((rc = select_kth(no_block, bufv[0], static_cast<int64_t>(count), K, &kth))) return rc;
      raise_radius(kth);
      if ((rc = compact())) return rc;
    }
  }

  if (count == 0) return VSCB200_OK;
  // exact scores + sort keys; the alternate survivor buffer is no longer needed
  sc.drop(bufv[1]); sc.drop(bufp[1]);
  const int64_t m = static_cast<int64_t>(count);
  uint64_t *kp = nullptr, *ks = nullptr, *kp2 = nullptr, *ks2 = nullptr;
  void* sort_ws = nullptr;
  if ((rc = sc.get(&kp, m * sizeof(uint64_t))) || (rc = sc.get(&ks, m * sizeof(uint64_t))) ||
      (rc = sc.get(&kp2, m * sizeof(uint64_t))) || (rc = sc.get(&ks2, m * sizeof(uint64_t))) ||
      (rc = sc.get(&sort_ws, radix_sort_scratch_bytes(m))))
    return rc;
  if ((rc = set_counter(0))) return rc;
  gt_rescore_kernel<<<grid_for((m + 7) / 8), kGtThreads, 0, s>>>(bufp[0], m, q, ix->bank, ix->d, n, keep_max ? 1 : 0,
                                                                 use_thresh, thresh, kp, ks, counter);
  count_launch();
  // (1) by pair id, (2) stably by score key: best first, ties in (query row, bank row) order
  int alt = 0;
  if ((rc = radix_sort_pairs(kp, ks, kp2, ks2, m, 0, bits_for(total_pairs), sort_ws, &alt, s))) return rc;
  uint64_t *p_sorted = alt ? kp2 : kp, *s_by_p = alt ? ks2 : ks;
  uint64_t *p_other = alt ? kp : kp2, *s_other = alt ? ks : ks2;
  if ((rc = radix_sort_pairs(s_by_p, p_sorted, s_other, p_other, m, 0, 32, sort_ws, &alt, s))) return rc;
  const uint64_t* fs = alt ? s_other : s_by_p;
  const uint64_t* fp = alt ? p_other : p_sorted;
  unsigned long long n_fail = 0;
  if ((rc = read_counter(&n_fail))) return rc;
  int64_t keep = m - static_cast<int64_t>(n_fail);
  if (limited) keep = std::min<int64_t>(keep, static_cast<int64_t>(K));
  if (keep > 0) {
    if ((rc = grow(&ix->g_score, &ix->g_score_bytes, keep * sizeof(float), s))) return rc;
    if ((rc = grow(&ix->g_q, &ix->g_q_bytes, keep * sizeof(int64_t), s))) return rc;
    if ((rc = grow(&ix->g_r, &ix->g_r_bytes, keep * sizeof(int64_t), s))) return rc;
    gt_output_kernel<<<static_cast<unsigned>((keep + kGtThreads - 1) / kGtThreads), kGtThreads, 0, s>>>(
        fs, fp, keep, n, ix->id_offset, keep_max ? 1 : 0, ix->g_score, ix->g_q, ix->g_r);
    count_launch();
  }
  VSCB_CUDA_OK(cudaGetLastError());
  VSCB_CUDA_OK(cudaStreamSynchronize(s));
  ix->g_n = keep;
  *n_found = keep;
  return VSCB200_OK;
}

int vscb200_index_global_results(vscb200_index* ix, float* scores, int64_t* qrows, int64_t* brows, void* stream) {
  VSCB_REQUIRE(ix, "index_global_results: null index");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (ix->g_n == 0) return VSCB200_OK;
  VSCB_REQUIRE(scores && qrows && brows, "index_global_results: null output");
  VSCB_CUDA_OK(cudaMemcpyAsync(scores, ix->g_score, ix->g_n * sizeof(float), cudaMemcpyDefault, s));
  VSCB_CUDA_OK(cudaMemcpyAsync(qrows, ix->g_q, ix->g_n * sizeof(int64_t), cudaMemcpyDefault, s));
  VSCB_CUDA_OK(cudaMemcpyAsync(brows, ix->g_r, ix->g_n * sizeof(int64_t), cudaMemcpyDefault, s));
  VSCB_CUDA_OK(cudaStreamSynchronize(s));
  return VSCB200_OK;
}

int vscb200_index_global_video_pairs(vscb200_index* ix, const int64_t* q_offsets, int64_t nqv, const int64_t* r_offsets,
                                     int64_t nrv, int64_t* n_pairs, void* stream) {
  VSCB_REQUIRE(ix && n_pairs && q_offsets && r_offsets && nqv > 0 && nrv > 0, "index_global_video_pairs: bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  *n_pairs = 0;
  ix->vp_n = 0;
  const int64_t n = ix->g_n;
  if (n == 0) return VSCB200_OK;
  VSCB_REQUIRE(n < (1ll << 31), "index_global_video_pairs: too many frame pairs");
  Scratch sc(s);
  uint64_t slots = 1024;
  while (slots < static_cast<uint64_t>(n) * 2) slots <<= 1;
  unsigned long long *tkey = nullptr, *counter = nullptr;
  uint32_t* tmin = nullptr;
  uint64_t *fi = nullptr, *kk = nullptr, *fi2 = nullptr, *kk2 = nullptr;
  void* sort_ws = nullptr;
  int rc;
  if ((rc = sc.get(&tkey, slots * sizeof(unsigned long long))) || (rc = sc.get(&tmin, slots * sizeof(uint32_t))) ||
      (rc = sc.get(&counter, sizeof(unsigned long long))) || (rc = sc.get(&fi, n * sizeof(uint64_t))) ||
      (rc = sc.get(&kk, n * sizeof(uint64_t))) || (rc = sc.get(&fi2, n * sizeof(uint64_t))) ||
      (rc = sc.get(&kk2, n * sizeof(uint64_t))) || (rc = sc.get(&sort_ws, radix_sort_scratch_bytes(n))))
    return rc;
  VSCB_CUDA_OK(cudaMemsetAsync(tkey, 0xff, slots * sizeof(unsigned long long), s));
  VSCB_CUDA_OK(cudaMemsetAsync(tmin, 0xff, slots * sizeof(uint32_t), s));
  VSCB_CUDA_OK(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), s));
  const unsigned g = static_cast<unsigned>((n + kGtThreads - 1) / kGtThreads);
  vp_insert_kernel<<<g, kGtThreads, 0, s>>>(ix->g_q, ix->g_r, n, ix->id_offset, q_offsets, nqv, r_offsets, nrv, tkey, tmin,
                                            slots - 1);
  vp_extract_kernel<<<static_cast<unsigned>((slots + kGtThreads - 1) / kGtThreads), kGtThreads, 0, s>>>(tkey, tmin, slots, fi,
                                                                                                        kk, counter);
  count_launch(2);
  unsigned long long m = 0;
  VSCB_CUDA_OK(cudaMemcpyAsync(&m, counter, sizeof(m), cudaMemcpyDeviceToHost, s));
  VSCB_CUDA_OK(cudaStreamSynchronize(s));
  int alt = 0;
  if ((rc = radix_sort_pairs(fi, kk, fi2, kk2, static_cast<int64_t>(m), 0, bits_for(static_cast<uint64_t>(n)), sort_ws, &alt, s)))
    return rc;
  if ((rc = grow(&ix->vp_score, &ix->vp_score_bytes, m * sizeof(float), s))) return rc;
  if ((rc = grow(&ix->vp_q, &ix->vp_q_bytes, m * sizeof(int64_t), s))) return rc;
  if ((rc = grow(&ix->vp_r, &ix->vp_r_bytes, m * sizeof(int64_t), s))) return rc;
  vp_output_kernel<<<static_cast<unsigned>((m + kGtThreads - 1) / kGtThreads), kGtThreads, 0, s>>>(
      alt ? fi2 : fi, alt ? kk2 : kk, static_cast<int64_t>(m), nrv, ix->g_score, ix->vp_score, ix->vp_q, ix->vp_r);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  VSCB_CUDA_OK(cudaStreamSynchronize(s));
  ix->vp_n = static_cast<int64_t>(m);
  *n_pairs = ix->vp_n;
  return VSCB200_OK;
}

int vscb200_index_video_pair_results(vscb200_index* ix, float* scores, int64_t* qv, int64_t* rv, void* stream) {
  VSCB_REQUIRE(ix, "index_video_pair_results: null index");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (ix->vp_n == 0) return VSCB200_OK;
  VSCB_REQUIRE(scores && qv && rv, "index_video_pair_results: null output");
  VSCB_CUDA_OK(cudaMemcpyAsync(scores, ix->vp_score, ix->vp_n * sizeof(float), cudaMemcpyDefault, s));
  VSCB_CUDA_OK(cudaMemcpyAsync(qv, ix->vp_q, ix->vp_n * sizeof(int64_t), cudaMemcpyDefault, s));
  VSCB_CUDA_OK(cudaMemcpyAsync(rv, ix->vp_r, ix->vp_n * sizeof(int64_t), cudaMemcpyDefault, s));
  VSCB_CUDA_OK(cudaStreamSynchronize(s));
  return VSCB200_OK;
}

}  // extern "C"
