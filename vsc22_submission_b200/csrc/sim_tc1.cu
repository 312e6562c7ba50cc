// Batch top-k similarity search at ONE bf16 MMA per product (k <= 10): selection on approximate scores with a proven
// error margin, exact fp32 scores for the survivors.
//
//   sim1_topk_kernel     a ~ q.r from the hi planes only (bf16(q).bf16(r), fp32 accumulation) on tcgen05 with CTA pairs
//                        (cta_group::2, 256 x 256 tiles, the skeleton of gemm.cu); the epilogue keeps, per thread (= one
//                        query row x one 128-column half of the tiles of a bank slab), the 16 best approximate scores
//                        with their bank rows -- no score ever reaches HBM.
//   row_rescore_kernel   one warp per query: A_k = k-th best approximate score over the thread lists; every bank row with
//                        a >= A_k - 2 eps survives (<= 64), is rescored in exact fp32 (exact.cuh, the one summation order
//                        every search path reports) and the k best by (score, lower id) are emitted.
//   exact_row_topk_kernel  brute-force fp32 search of the flagged rows (survivor overflow), normally none.
//
// Why this is exact.  Let |a(r) - s(r)| <= eps for every bank row r of a query (s = fp32 score).  bf16 rounding moves
// every element by a relative 2^-9 at most, so by Cauchy-Schwarz |bf16(q).bf16(r) - q.r| <= (2^-8 + 2^-17) |q| |r|; the
// fp32 accumulation of the exact bf16 products adds at most d 2^-24 |q| |r|.  eps = 0.00415 |q| max_r |r| covers both for
// d <= 4096.  The k rows with the best a have s >= A_k - eps, so the k-th best exact score s_k >= A_k - eps, and a row
// of the exact top-k has a >= s_k - eps >= A_k - 2 eps: it is a survivor.  A thread list that is full with its 16th
// entry inside the margin may have dropped a survivor: such queries (and queries with more than 64 survivors) are
// flagged and searched exhaustively in fp32.  Reference semantics: faiss IndexFlat.search behind vsc/index.py:174 and
// vsc/baseline/score_normalization.py:93-98 (exact scores, best first, ties to the lower id).
#include <float.h>

#include "exact.cuh"
#include "host_util.h"
#include "kernels.h"
#include "ptx.cuh"

namespace vscb200 {

constexpr int kS1BM = 128, kS1BN = 256, kS1BK = 64;     // per-CTA rows, pair-tile columns, K block
constexpr int kS1Stages = 6;
constexpr int kS1EpiWarps = 8;
constexpr int kS1Threads = 128 + 32 * kS1EpiWarps;
constexpr int kS1ATile = kS1BM * kS1BK * 2;              // 16 KB
constexpr int kS1BTile = (kS1BN / 2) * kS1BK * 2;        // 16 KB: each CTA stages half of the bank tile
constexpr int kS1Smem = kS1Stages * (kS1ATile + kS1BTile) + 256 + 1024;
constexpr int kS1List = 16;                              // per-thread candidate list
constexpr int kS1MaxSurv = 64;                           // survivors rescored per query
constexpr float kS1EpsCoef = 0.00415f;                   // (2^-8 + 2^-12): bf16 rounding of both operands + accumulation

struct Sim1Params {
  int64_t nq, nr;
  int K;                 // padded feature dim
  int l2;
  const float* qn;       // squared query norms (L2 keys)
  const float* rn;       // squared bank norms
  int tiles_m2, tiles_n, slabs;     // pair tiles along M, 256-wide tiles along N, bank slabs
  float* cand_v;         // [nq, slabs * 2 * kS1List] selection keys (larger = better; -distance for L2)
  int32_t* cand_i;       // bank rows (-1 = empty)
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kS1Threads, 1)
sim1_topk_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmR, Sim1Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + kS1Stages * kS1ATile;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + kS1Stages * kS1BTile);
  uint64_t* empty_bar = full_bar + kS1Stages;
  uint64_t* tfull_bar = empty_bar + kS1Stages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) { prefetch_tmap(&tmQ); prefetch_tmap(&tmR); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kS1Stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], kS1EpiWarps * 2); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_cg2<2 * kS1BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t crank = cluster_ctarank();
  const int num_items = p.tiles_m2 * p.slabs;
  const int first = static_cast<int>(blockIdx.x >> 1), step = static_cast<int>(gridDim.x >> 1);
  const int kblocks = (p.K + kS1BK - 1) / kS1BK;
  auto slab_n0 = [&](int sl) { return static_cast<int>(static_cast<int64_t>(sl) * p.tiles_n / p.slabs); };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = first; item < num_items; item += step) {
        const int pm = item / p.slabs, sl = item % p.slabs;
        const int m_blk = pm * 2 + static_cast<int>(crank);
        for (int n_blk = slab_n0(sl), n_end = slab_n0(sl + 1); n_blk < n_end; ++n_blk) {
          for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            const uint32_t lead_full = mapa_u32(smem_u32(&full_bar[stage]), 0u);
            if (crank == 0) mbar_expect_tx(&full_bar[stage], 2 * (kS1ATile + kS1BTile));
            tma_load_2d_cg2(sA + stage * kS1ATile, &tmQ, lead_full, kb * kS1BK, m_blk * kS1BM, kEvictLast);
            tma_load_2d_cg2(sB + stage * kS1BTile, &tmR, lead_full, kb * kS1BK, n_blk * kS1BN + static_cast<int>(crank) * (kS1BN / 2),
                            kEvictNormal);
            if (++stage == kS1Stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && crank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(2 * kS1BM, kS1BN);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int item = first; item < num_items; item += step) {
        const int sl = item % p.slabs;
        for (int n_blk = slab_n0(sl), n_end = slab_n0(sl + 1); n_blk < n_end; ++n_blk) {
          mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * kS1BN;
          for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint64_t a_desc = make_desc_k_sw128(smem_u32(sA + stage * kS1ATile));
            const uint64_t b_desc = make_desc_k_sw128(smem_u32(sB + stage * kS1BTile));
#pragma unroll
            for (int k = 0; k < kS1BK / 16; ++k)
              umma_bf16_ss_cg2(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit_cg2_mcast(&empty_bar[stage], static_cast<uint16_t>(0x3));
            if (++stage == kS1Stages) { stage = 0; phase ^= 1; }
          }
          umma_commit_cg2_mcast(&tfull_bar[acc], static_cast<uint16_t>(0x3));
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    const int ew = warp - 4, quad = warp & 3, half = ew >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = first; item < num_items; item += step) {
      const int pm = item / p.slabs, sl = item % p.slabs;
      const int64_t grow = static_cast<int64_t>(pm * 2 + static_cast<int>(crank)) * kS1BM + quad * 32 + lane;
      // this thread's running top-kS1List of query row `grow` over its columns of the slab, best first; strict
      // comparisons keep the earlier (lower) bank row among equal keys
      float ls[kS1List];
      int32_t li[kS1List];
#pragma unroll
      for (int j = 0; j < kS1List; ++j) { ls[j] = -INFINITY; li[j] = -1; }
      const float qn_row = (p.l2 && grow < p.nq) ? p.qn[grow] : 0.f;
      for (int n_blk = slab_n0(sl), n_end = slab_n0(sl + 1); n_blk < n_end; ++n_blk) {
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < kS1BN / 2 / 32; ++c) {
          const int col0 = half * (kS1BN / 2) + c * 32;
          const int64_t gcol = static_cast<int64_t>(n_blk) * kS1BN + col0;
          uint32_t v[32];
          __syncwarp();
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kS1BN + col0, v);
          tmem_ld_wait();
          if (gcol >= p.nr) continue;                                   // warp-uniform
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (p.l2) {                                                   // key = -(|q|^2 + |r|^2 - 2 q.r)
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = 2.0f * f[j] - __ldg(p.rn + min(gcol + j, p.nr - 1)) - qn_row;
          }
          if (gcol + 32 > p.nr) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (gcol + j >= p.nr) f[j] = -INFINITY;
          }
          float m0 = fmaxf(f[0], f[1]), m1 = fmaxf(f[2], f[3]), m2 = fmaxf(f[4], f[5]), m3 = fmaxf(f[6], f[7]);
#pragma unroll
          for (int j = 8; j < 32; j += 4) {
            m0 = fmaxf(m0, f[j]); m1 = fmaxf(m1, f[j + 1]); m2 = fmaxf(m2, f[j + 2]); m3 = fmaxf(m3, f[j + 3]);
          }
          if (fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) > ls[kS1List - 1]) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (f[j] > ls[kS1List - 1]) {
                ls[kS1List - 1] = f[j];
                li[kS1List - 1] = static_cast<int32_t>(gcol + j);
#pragma unroll
                for (int t = kS1List - 1; t > 0; --t) {
                  if (ls[t] > ls[t - 1]) {
                    const float ts = ls[t]; ls[t] = ls[t - 1]; ls[t - 1] = ts;
                    const int32_t ti = li[t]; li[t] = li[t - 1]; li[t - 1] = ti;
                  }
                }
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(mapa_u32(smem_u32(&tempty_bar[acc]), 0u));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (grow < p.nq) {
        const int64_t base = (grow * p.slabs + sl) * (2 * kS1List) + half * kS1List;
#pragma unroll
        for (int j = 0; j < kS1List; ++j) {
          p.cand_v[base + j] = ls[j];
          p.cand_i[base + j] = li[j];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_cg2<2 * kS1BN>(tmem_base);
}

// ------------------------------------------------------------------ query prologue: hi plane + squared norm in one pass
__global__ void __launch_bounds__(256)
q_hi_norm_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi, float* __restrict__ sq, int64_t n, int d, int dp) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  float s = 0.f;
  for (int c = lane; c < dp; c += 32) {
    const float v = c < d ? x[row * d + c] : 0.f;
    s = fmaf(v, v, s);
    hi[row * dp + c] = __float2bfloat16_rn(v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) sq[row] = s;
}

int q_hi_norm(const float* x, void* hi, float* sq, int64_t n, int d, int dp, cudaStream_t stream) {
  if (n == 0) return VSCB200_OK;
  q_hi_norm_kernel<<<static_cast<unsigned>((n * 32 + 255) / 256), 256, 0, stream>>>(x, reinterpret_cast<__nv_bfloat16*>(hi), sq, n, d, dp);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

// ------------------------------------------------------------------ survivors -> exact scores -> top-k
__device__ __forceinline__ uint32_t okey32(float f) {          // order-preserving float -> uint (larger float = larger key)
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float okey32_inv(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}
// composite key: (score key, ~id): larger = better, equal scores -> the lower id wins; 0 = empty
__device__ __forceinline__ unsigned long long ckey(float key, uint32_t id) {
  return (static_cast<unsigned long long>(okey32(key)) << 32) | static_cast<uint32_t>(~id);
}
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
  const uint32_t hi = static_cast<uint32_t>(v >> 32);
  const uint32_t mh = __reduce_max_sync(0xffffffffu, hi);
  const uint32_t lo = hi == mh ? static_cast<uint32_t>(v) : 0u;
  const uint32_t ml = __reduce_max_sync(0xffffffffu, lo);
  return (static_cast<unsigned long long>(mh) << 32) | ml;
}

constexpr int kRrWarps = 8;

// cand_v / cand_i: [nq, nlists * kS1List] (each list best first, -1 = empty).  rmax2: max squared bank norm (device).
__global__ void __launch_bounds__(kRrWarps * 32)
row_rescore_kernel(const float* __restrict__ Q, const float* __restrict__ bank, int d, int l2, int64_t nq,
                   const float* __restrict__ cand_v, const int32_t* __restrict__ cand_i, int nlists, int k,
                   const float* __restrict__ qn, const unsigned int* __restrict__ rmax2_bits, float* __restrict__ D,
                   int64_t* __restrict__ I, int64_t id_offset, int* __restrict__ flags, int* __restrict__ n_flagged) {
  extern __shared__ __align__(16) uint8_t rr_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sq = reinterpret_cast<float*>(rr_smem) + warp * d;                                   // this warp's query row
  unsigned long long* surv = reinterpret_cast<unsigned long long*>(rr_smem + static_cast<size_t>(kRrWarps) * d * 4) + warp * kS1MaxSurv;
  const int64_t qrow = static_cast<int64_t>(blockIdx.x) * kRrWarps + warp;
  if (qrow >= nq) return;
  const bool keep_max = !l2;
  const int ncand = nlists * kS1List;
  const float* cv = cand_v + qrow * ncand;
  const int32_t* ci = cand_i + qrow * ncand;
  for (int c = lane; c < d; c += 32) sq[c] = Q[qrow * d + c];
  // ---- A_k: k-th best approximate key over all lists (k rounds of "best key below the previous one")
  unsigned long long prev = ~0ull;
  int found = 0;
  for (int r = 0; r < k; ++r) {
    unsigned long long best = 0ull;
    for (int c = lane; c < ncand; c += 32) {
      const int32_t id = ci[c];
      if (id < 0) continue;
      const unsigned long long key = ckey(cv[c], static_cast<uint32_t>(id));
      if (key < prev && key > best) best = key;
    }
    best = warp_max_u64(best);
    if (best == 0ull) break;
    prev = best;
    ++found;
  }
  // ---- margin
  const float qnorm = sqrtf(qn[qrow]), rmax = sqrtf(__uint_as_float(*rmax2_bits));
  float eps = kS1EpsCoef * qnorm * rmax * 1.001f;
  if (l2) eps = 2.0f * eps + 1e-6f * (qn[qrow] + __uint_as_float(*rmax2_bits));
  const float thr = found == k ? okey32_inv(static_cast<uint32_t>(prev >> 32)) - 2.0f * eps : -INFINITY;
  // ---- survivors (+ overflow detection: a full list whose last entry is inside the margin may have dropped one)
  bool overflow = false;
  int ns = 0;
  for (int c0 = 0; c0 < ncand; c0 += 32) {
    const int c = c0 + lane;
    const int32_t id = c < ncand ? ci[c] : -1;
    const float v = c < ncand ? cv[c] : 0.f;
    const bool take = id >= 0 && v >= thr;
    if (take && (c % kS1List) == kS1List - 1) overflow = true;
    const uint32_t m = __ballot_sync(0xffffffffu, take);
    const int pos = ns + __popc(m & ((1u << lane) - 1u));
    if (take && pos < kS1MaxSurv) surv[pos] = static_cast<uint32_t>(id);
    ns += __popc(m);
  }
  overflow = __any_sync(0xffffffffu, overflow) || ns > kS1MaxSurv;
  if (ns > kS1MaxSurv) ns = kS1MaxSurv;
  __syncwarp();
  // ---- exact fp32 scores of the survivors, four rows per round
  for (int c0 = 0; c0 < ns; c0 += 4) {
    const float* rp[4];
    uint32_t ids[4];
    float acc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      ids[u] = static_cast<uint32_t>(surv[min(c0 + u, ns - 1)]);
      rp[u] = bank + static_cast<int64_t>(ids[u]) * d;
    }
    exact_rows_warp<4>(sq, rp, d, lane, l2 != 0, acc);
    __syncwarp();
    if (lane < 4 && c0 + lane < ns) {
      const float a = lane == 0 ? acc[0] : (lane == 1 ? acc[1] : (lane == 2 ? acc[2] : acc[3]));
      const uint32_t id = lane == 0 ? ids[0] : (lane == 1 ? ids[1] : (lane == 2 ? ids[2] : ids[3]));
      surv[c0 + lane] = ckey(keep_max ? a : -a, id);
    }
  }
  __syncwarp();
  // ---- the k best by (exact score, lower id)
  prev = ~0ull;
  for (int r = 0; r < k; ++r) {
    unsigned long long best = 0ull;
    for (int c = lane; c < ns; c += 32) {
      const unsigned long long key = surv[c];
      if (key < prev && key > best) best = key;
    }
    best = warp_max_u64(best);
    if (lane == 0) {
      if (best != 0ull) {
        const float sc = okey32_inv(static_cast<uint32_t>(best >> 32));
        D[qrow * k + r] = keep_max ? sc : -sc;
        I[qrow * k + r] = id_offset + static_cast<int64_t>(~static_cast<uint32_t>(best & 0xFFFFFFFFull));
      } else {
        D[qrow * k + r] = keep_max ? -FLT_MAX : FLT_MAX;
        I[qrow * k + r] = -1;
      }
    }
    if (best != 0ull) prev = best;
    else prev = 0ull;
  }
  if (lane == 0) {
    flags[qrow] = overflow ? 1 : 0;
    if (overflow) atomicAdd(n_flagged, 1);
  }
}

// ------------------------------------------------------------------ exhaustive fp32 search of the flagged queries
// One CTA per flagged query at a time; every warp keeps its k best rows as a sorted list spread over its lanes
// (lane j = rank j), the warp lists are merged through shared memory.  k <= 32.
constexpr int kExWarps = 8;

__global__ void __launch_bounds__(kExWarps * 32)
exact_row_topk_kernel(const float* __restrict__ Q, const float* __restrict__ bank, int d, int l2, int64_t nq, int64_t nr, int k,
                      float* __restrict__ D, int64_t* __restrict__ I, int64_t id_offset, const int* __restrict__ flags,
                      const int* __restrict__ n_flagged) {
  extern __shared__ __align__(16) uint8_t ex_smem[];
  if (*n_flagged == 0) return;
  float* sq = reinterpret_cast<float*>(ex_smem);                                              // [d]
  unsigned long long* merged = reinterpret_cast<unsigned long long*>(ex_smem + static_cast<size_t>(d) * 4);   // [kExWarps * 32]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool keep_max = !l2;
  for (int64_t qrow = blockIdx.x; qrow < nq; qrow += gridDim.x) {
    if (!flags[qrow]) continue;                                       // CTA-uniform
    __syncthreads();
    for (int c = threadIdx.x; c < d; c += kExWarps * 32) sq[c] = Q[qrow * d + c];
    __syncthreads();
    unsigned long long mine = 0ull;                                   // lane j: the warp's j-th best composite key
    for (int64_t r0 = static_cast<int64_t>(warp) * 4; r0 < nr; r0 += kExWarps * 4) {
      const float* rp[4];
      float acc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) rp[u] = bank + (r0 + u < nr ? r0 + u : nr - 1) * d;
      exact_rows_warp<4>(sq, rp, d, lane, l2 != 0, acc);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (r0 + u >= nr) continue;
        const unsigned long long key = ckey(keep_max ? acc[u] : -acc[u], static_cast<uint32_t>(r0 + u));
        const unsigned long long last = __shfl_sync(0xffffffffu, mine, k - 1);
        if (key <= last) continue;                                    // warp-uniform
        const int pos = __popc(__ballot_sync(0xffffffffu, mine > key));
        const unsigned long long up = __shfl_up_sync(0xffffffffu, mine, 1);
        if (lane == pos) mine = key;
        else if (lane > pos) mine = up;
      }
    }
    merged[warp * 32 + lane] = lane < k ? mine : 0ull;
    __syncthreads();
    if (warp == 0) {
      unsigned long long prev = ~0ull;
      for (int r = 0; r < k; ++r) {
        unsigned long long best = 0ull;
        for (int c = lane; c < kExWarps * 32; c += 32) {
          const unsigned long long key = merged[c];
          if (key < prev && key > best) best = key;
        }
        best = warp_max_u64(best);
        if (lane == 0) {
          if (best != 0ull) {
            const float sc = okey32_inv(static_cast<uint32_t>(best >> 32));
            D[qrow * k + r] = keep_max ? sc : -sc;
            I[qrow * k + r] = id_offset + static_cast<int64_t>(~static_cast<uint32_t>(best & 0xFFFFFFFFull));
          } else {
            D[qrow * k + r] = keep_max ? -FLT_MAX : FLT_MAX;
            I[qrow * k + r] = -1;
          }
        }
        prev = best;
      }
    }
  }
}

// ------------------------------------------------------------------ host side
// Bank slabs per pair tile row: enough (pair tile, slab) items to fill the 74 CTA pairs about three times over, chosen
// among [s_min, 3 s_min] for the most even split of the items over the pairs.
int sim1_slabs(int64_t nq, int64_t nr) {
  const int pairs = device_sm_count() / 2;
  const int64_t tm2 = (nq + 2 * kS1BM - 1) / (2 * kS1BM), tn = (nr + kS1BN - 1) / kS1BN;
  int64_t s_min = (3ll * pairs + tm2 - 1) / tm2;
  if (s_min < 1) s_min = 1;
  int64_t s_max = 3 * s_min;
  if (s_max > 32) s_max = 32;
  if (s_max > tn) s_max = tn;
  if (s_min > s_max) s_min = s_max;
  int64_t best = s_min;
  double best_eff = 0.0;
  for (int64_t s = s_min; s <= s_max; ++s) {
    const int64_t items = tm2 * s;
    const double eff = static_cast<double>(items) / static_cast<double>((items + pairs - 1) / pairs * pairs);
    if (eff > best_eff + 0.02) { best_eff = eff; best = s; }
  }
  return static_cast<int>(best);
}
int sim1_list_len() { return kS1List; }

// Qh [nq, dp] / Rh [nr, dp]: bf16 hi planes.  cand_v / cand_i: [nq, slabs * 2 * kS1List].
int sim1_topk(const void* Qh, const void* Rh, int64_t nq, int64_t nr, int dp, bool l2, const float* qn, const float* rn, int slabs,
              float* cand_v, int32_t* cand_i, cudaStream_t stream) {
  if (nq == 0 || nr == 0) return VSCB200_OK;
  VSCB_REQUIRE(dp % 8 == 0 && nr < (1ll << 31), "sim1_topk: dp must be a multiple of 8 and nr < 2^31");
  CUtensorMap tQ, tR;
  int rc;
  if ((rc = make_tmap_2d(&tQ, Qh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, nq, dp, dp, kS1BM, kS1BK, true))) return rc;
  if ((rc = make_tmap_2d(&tR, Rh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, nr, dp, dp, kS1BN / 2, kS1BK, true))) return rc;
  Sim1Params p = {};
  p.nq = nq; p.nr = nr; p.K = dp; p.l2 = l2 ? 1 : 0; p.qn = qn; p.rn = rn;
  p.tiles_m2 = static_cast<int>((nq + 2 * kS1BM - 1) / (2 * kS1BM));
  const int64_t tn = (nr + kS1BN - 1) / kS1BN;
  VSCB_REQUIRE(static_cast<int64_t>(p.tiles_m2) * slabs < (1ll << 31) && tn < (1ll << 31), "sim1_topk: too many tiles");
  p.tiles_n = static_cast<int>(tn);
  p.slabs = slabs; p.cand_v = cand_v; p.cand_i = cand_i;
  const int64_t items = static_cast<int64_t>(p.tiles_m2) * slabs;
  const int pairs = device_sm_count() / 2;
  const int grid = 2 * static_cast<int>(items < pairs ? items : pairs);
  VSCB_CUDA_OK(cudaFuncSetAttribute(sim1_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kS1Smem));
  ProfScope prof(kProfScores, stream, 2.0 * static_cast<double>(nq) * nr * dp);
  sim1_topk_kernel<<<grid, kS1Threads, kS1Smem, stream>>>(tQ, tR, p);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

// flags [nq] / n_flagged [1]: device scratch (n_flagged must be zero on entry; the fallback kernel reads it)
int sim1_rescore(const float* Q, const float* bank, int d, bool l2, int64_t nq, int64_t nr, const float* cand_v,
                 const int32_t* cand_i, int slabs, int k, const float* qn, const unsigned int* rmax2_bits, float* D, int64_t* I,
                 int64_t id_offset, int* flags, int* n_flagged, cudaStream_t stream) {
  if (nq == 0) return VSCB200_OK;
  VSCB_REQUIRE(k >= 1 && k <= kS1List, "sim1_rescore: k must be <= the candidate list length");
  const size_t smem = static_cast<size_t>(kRrWarps) * d * 4 + static_cast<size_t>(kRrWarps) * kS1MaxSurv * 8;
  VSCB_REQUIRE(smem <= 200 * 1024 && d % 4 == 0, "sim1_rescore: dimension must be a multiple of 4 and <= 6144");
  VSCB_CUDA_OK(cudaFuncSetAttribute(row_rescore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  {
    ProfScope prof(kProfSelect, stream, static_cast<double>(nq) * slabs * 2 * kS1List * 8);
    row_rescore_kernel<<<static_cast<unsigned>((nq + kRrWarps - 1) / kRrWarps), kRrWarps * 32, smem, stream>>>(
        Q, bank, d, l2 ? 1 : 0, nq, cand_v, cand_i, slabs * 2, k, qn, rmax2_bits, D, I, id_offset, flags, n_flagged);
    count_launch();
  }
  const size_t smem_ex = static_cast<size_t>(d) * 4 + static_cast<size_t>(kExWarps) * 32 * 8;
  VSCB_CUDA_OK(cudaFuncSetAttribute(exact_row_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_ex)));
  const int64_t grid_ex = nq < 8ll * device_sm_count() ? nq : 8ll * device_sm_count();
  {
    ProfScope prof(kProfSelect, stream, 0.0);
    exact_row_topk_kernel<<<static_cast<unsigned>(grid_ex), kExWarps * 32, smem_ex, stream>>>(Q, bank, d, l2 ? 1 : 0, nq, nr, k, D, I,
                                                                                             id_offset, flags, n_flagged);
    count_launch();
  }
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

}  // namespace vscb200
