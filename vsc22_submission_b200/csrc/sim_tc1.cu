// Batch top-k similarity search at ONE bf16 MMA per product (k <= 10): selection on approximate scores with a proven
// error margin, exact fp32 scores for the survivors.
//
//   sim1_topk_kernel     a ~ q.r from the hi planes only (bf16(q).bf16(r), fp32 accumulation) on tcgen05 with CTA pairs
//                        (cta_group::2, 256 x 256 tiles, the skeleton of gemm.cu).  Every CTA pair walks ONE contiguous
//                        range of the (query pair-tile, bank tile) sequence: perfectly balanced, and a query row meets as
//                        few pairs as possible (<= 3 at config 3).  Epilogue: each warp transposes its 32 rows x 32
//                        columns of scores through a 4 KB staging tile and then works ROW BY ROW -- lane = column for
//                        the comparison against the row's threshold, lane = rank for the row's running top-32 list,
//                        which lives in one register pair per row across the warp (insert = ballot + shuffle-up).  A
//                        thread-per-row list would make the warp execute the union of 32 rows' insertions; here a chunk
//                        without hits costs 5 instructions per row.  No score reaches HBM.
//   row_rescore_kernel   one warp per query: A_k = k-th best approximate score over the row's lists; every bank row with
//                        a >= A_k - 2 eps is rescored in exact fp32 (exact.cuh, the one summation order every search
//                        path reports) and streamed into a warp-wide top-k by (score, lower id).
//   exact_row_topk_kernel  brute-force fp32 search of the flagged rows (a full list inside the margin), normally none.
//
// Why this is exact.  Write q = qh + ql, r = rh + rl (h = bf16 rounding, l = its residual).  The kernel computes
// a = fl(qh.rh); s = q.r = qh.rh + ql.r + qh.rl, so by Cauchy-Schwarz |a - s| <= |ql| |r| + |qh| |rl| + d 2^-23 |q| |r|
// (last term: fp32 accumulation of the exact bf16 products).  |ql| and |qh| are measured per query, max |r| and
// max |rl| per bank at add(): eps(q) = |ql| Rmax + |qh| RLmax + d 2^-23 |q| Rmax -- about 0.0026 |q| Rmax in practice, below
// the worst case 2^-8.  The k rows with the best a have s >= A_k - eps, so the k-th best exact score s_k >= A_k - eps, and
// a row of the exact top-k has a >= s_k - eps >= A_k - 2 eps: it is a survivor.  A list that is full with its last entry
// inside the margin may have dropped a survivor: such queries are flagged and searched exhaustively in fp32.  Reference
// semantics: faiss IndexFlat.search behind vsc/index.py:174 and vsc/baseline/score_normalization.py:93-98 (exact scores,
// best first, ties to the lower id).
#include <float.h>

#include "exact.cuh"
#include "host_util.h"
#include "kernels.h"
#include "ptx.cuh"

namespace vscb200 {

constexpr int kS1BM = 128, kS1BN = 256, kS1BK = 64;     // per-CTA rows, pair-tile columns, K block
constexpr int kS1Stages = 5;
constexpr int kS1EpiWarps = 8;
constexpr int kS1Threads = 128 + 32 * kS1EpiWarps;
constexpr int kS1ATile = kS1BM * kS1BK * 2;              // 16 KB
constexpr int kS1BTile = (kS1BN / 2) * kS1BK * 2;        // 16 KB: each CTA stages half of the bank tile
constexpr int kS1Smem = kS1Stages * (kS1ATile + kS1BTile) + kS1EpiWarps * 4096 + 256 + 1024;
constexpr int kS1List = 32;                              // candidate list per (query row, column half, pair): one entry per lane

struct Sim1Params {
  int64_t nq, nr;
  int K;                 // padded feature dim
  int l2;
  const float* qn;       // squared query norms (L2 keys)
  const float* rn;       // squared bank norms
  int tiles_m2, tiles_n;            // pair tiles along M, 256-wide tiles along N
  int pairs, slots;                 // CTA pairs launched; candidate-list slots per query row (pairs that may meet a row)
  int depth;                        // entries kept per list (<= kS1List): 8 / 16 / 32 by k
  float* cand_v;         // [nq, slots * 2 * kS1List] selection keys (larger = better; -distance for L2)
  int32_t* cand_i;       // bank rows (-1 = empty; the buffer is pre-filled with -1)
};

// Work split: pair p owns tiles [tile_start(p), tile_start(p + 1)) of the sequence t = pm * tiles_n + n_blk.
__host__ __device__ inline int64_t s1_tile_start(int64_t p, int64_t T, int64_t P) { return p * T / P; }
// the pair whose range holds tile t
__host__ __device__ inline int64_t s1_pair_of(int64_t t, int64_t T, int64_t P) { return ((t + 1) * P - 1) / T; }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kS1Threads, 1)
sim1_topk_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmR, Sim1Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + kS1Stages * kS1ATile;
  uint8_t* stage_all = sB + kS1Stages * kS1BTile;          // per-warp 4 KB staging rows of the epilogue
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stage_all + kS1EpiWarps * 4096);
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
  const int64_t T = static_cast<int64_t>(p.tiles_m2) * p.tiles_n;
  const int64_t pair = blockIdx.x >> 1;
  const int64_t t_begin = s1_tile_start(pair, T, p.pairs), t_end = s1_tile_start(pair + 1, T, p.pairs);
  const int kblocks = (p.K + kS1BK - 1) / kS1BK;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t t = t_begin; t < t_end; ++t) {
        const int m_blk = static_cast<int>(t / p.tiles_n) * 2 + static_cast<int>(crank), n_blk = static_cast<int>(t % p.tiles_n);
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
  } else if (warp == 1) {
    if (lane == 0 && crank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(2 * kS1BM, kS1BN);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int64_t t = t_begin; t < t_end; ++t) {
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
  } else if (warp >= 4) {
    const int ew = warp - 4, quad = warp & 3, half = ew >> 2;
    uint8_t* tile = stage_all + ew * 4096;             // 32 rows x 128 B, 16-byte chunk q of row r at q ^ (r & 7)
    int acc = 0;
    uint32_t acc_phase = 0;
    // ls[r] / li[r]: the running top-32 (key, bank row) of the warp's r-th query row over this warp's columns of the
    // pair's range -- lane = rank, best first; strict comparisons keep the earlier (lower) bank row among equal keys
    float ls[32];
    int32_t li[32];
    int cur_pm = -1;
    int64_t row0 = 0;
    float qn_lane = 0.f;
    const int tl = p.depth - 1;                        // lane holding a list's threshold (its last kept entry)
    auto flush = [&]() {
      if (cur_pm < 0) return;
      const int slot = static_cast<int>(pair - s1_pair_of(static_cast<int64_t>(cur_pm) * p.tiles_n, T, p.pairs));
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        if (row0 + r < p.nq) {
          const int64_t base = (((row0 + r) * p.slots + slot) * 2 + half) * kS1List;
          p.cand_v[base + lane] = ls[r];
          p.cand_i[base + lane] = lane <= tl ? li[r] : -1;
        }
      }
    };
    for (int64_t t = t_begin; t < t_end; ++t) {
      const int pm = static_cast<int>(t / p.tiles_n), n_blk = static_cast<int>(t % p.tiles_n);
      if (pm != cur_pm) {
        flush();
        cur_pm = pm;
        row0 = static_cast<int64_t>(pm * 2 + static_cast<int>(crank)) * kS1BM + quad * 32;
        qn_lane = (p.l2 && row0 + lane < p.nq) ? p.qn[row0 + lane] : 0.f;
#pragma unroll
        for (int r = 0; r < 32; ++r) { ls[r] = -INFINITY; li[r] = -1; }
      }
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
        if (p.l2) {                                                   // key = -(|q|^2 + |r|^2 - 2 q.r); lane = query row here
#pragma unroll
          for (int j = 0; j < 32; ++j)
            v[j] = __float_as_uint(2.0f * __uint_as_float(v[j]) - __ldg(p.rn + min(gcol + j, p.nr - 1)) - qn_lane);
        }
        if (gcol + 32 > p.nr) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (gcol + j >= p.nr) v[j] = 0xFF800000u;                 // -inf
        }
        // transpose through the staging tile: written lane = row, read lane = column
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<uint4*>(tile + lane * 128 + ((q ^ (lane & 7)) << 4)) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        __syncwarp();
        // Pass 1, branch-free and fully pipelined: which of the warp's 32 rows have a score above their list's threshold?
        float val[32];
#pragma unroll
        for (int r = 0; r < 32; ++r)
          val[r] = *reinterpret_cast<const float*>(tile + r * 128 + ((((lane >> 2) ^ (r & 7)) << 4) | ((lane & 3) << 2)));
        uint32_t lanehit = 0;
#pragma unroll
        for (int r = 0; r < 32; ++r) lanehit |= val[r] > __shfl_sync(0xffffffffu, ls[r], tl) ? (1u << r) : 0u;
        const uint32_t rowmask = __reduce_or_sync(0xffffffffu, lanehit);
        if (rowmask == 0u) continue;                                  // the common case once the lists have warmed up
        // Pass 2: insert the hits, four rows at a time in straight-line predicated code -- an insertion is a chain of
        // dependent warp shuffles / votes (~70 cycles); four independent chains issued back to back hide it
#pragma unroll
        for (int g4 = 0; g4 < 8; ++g4) {
          if (((rowmask >> (4 * g4)) & 0xFu) == 0u) continue;         // warp-uniform
          uint32_t m[4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            m[i] = __ballot_sync(0xffffffffu, val[4 * g4 + i] > __shfl_sync(0xffffffffu, ls[4 * g4 + i], tl));
          while ((m[0] | m[1] | m[2] | m[3]) != 0u) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = 4 * g4 + i;
              const bool act = m[i] != 0u;
              const int b = act ? __ffs(m[i]) - 1 : 0;
              m[i] &= m[i] - 1u;
              const float vb = __shfl_sync(0xffffffffu, val[r], b);
              const bool ins = act && vb > __shfl_sync(0xffffffffu, ls[r], tl);
              const int pos = __popc(__ballot_sync(0xffffffffu, ls[r] >= vb));
              const float us = __shfl_up_sync(0xffffffffu, ls[r], 1);
              const int32_t ui = __shfl_up_sync(0xffffffffu, li[r], 1);
              const bool shift = ins && lane > pos, put = ins && lane == pos;
              ls[r] = shift ? us : (put ? vb : ls[r]);
              li[r] = shift ? ui : (put ? static_cast<int32_t>(gcol + b) : li[r]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(mapa_u32(smem_u32(&tempty_bar[acc]), 0u));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    flush();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_cg2<2 * kS1BN>(tmem_base);
}

// ------------------------------------------------------------------ query prologue: hi plane + norms in one pass
// sq[row] = |q|^2 ; sq_lo[row] = |q - bf16(q)|^2
__global__ void __launch_bounds__(256)
q_hi_norm_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi, float* __restrict__ sq, float* __restrict__ sq_lo,
                 int64_t n, int d, int dp) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  float s = 0.f, sl = 0.f;
  for (int c = lane; c < dp; c += 32) {
    const float v = c < d ? x[row * d + c] : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const float e = v - __bfloat162float(h);
    s = fmaf(v, v, s);
    sl = fmaf(e, e, sl);
    hi[row * dp + c] = h;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    sl += __shfl_xor_sync(0xffffffffu, sl, o);
  }
  if (lane == 0) { sq[row] = s; sq_lo[row] = sl; }
}

int q_hi_norm(const float* x, void* hi, float* sq, float* sq_lo, int64_t n, int d, int dp, cudaStream_t stream) {
  if (n == 0) return VSCB200_OK;
  q_hi_norm_kernel<<<static_cast<unsigned>((n * 32 + 255) / 256), 256, 0, stream>>>(x, reinterpret_cast<__nv_bfloat16*>(hi), sq, sq_lo,
                                                                                   n, d, dp);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

// bank side, at add(): running maxima (float bits; the values are >= 0) of |r|^2 and |r - bf16(r)|^2
__global__ void __launch_bounds__(256)
bank_norm_max_kernel(const float* __restrict__ x, int64_t n, int d, unsigned int* __restrict__ max_bits) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  float s = 0.f, sl = 0.f;
  for (int c = lane; c < d; c += 32) {
    const float v = x[row * d + c];
    const float e = v - __bfloat162float(__float2bfloat16_rn(v));
    s = fmaf(v, v, s);
    sl = fmaf(e, e, sl);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    sl += __shfl_xor_sync(0xffffffffu, sl, o);
  }
  if (lane == 0) {      // read first: after the first rows almost no row raises a maximum, and contended atomics serialise
    if (__float_as_uint(s) > *reinterpret_cast<volatile unsigned int*>(max_bits)) atomicMax(max_bits, __float_as_uint(s));
    if (__float_as_uint(sl) > *reinterpret_cast<volatile unsigned int*>(max_bits + 1)) atomicMax(max_bits + 1, __float_as_uint(sl));
  }
}

int bank_norm_max(const float* x, int64_t n, int d, unsigned int* max_bits, cudaStream_t stream) {
  if (n == 0) return VSCB200_OK;
  bank_norm_max_kernel<<<static_cast<unsigned>((n * 32 + 255) / 256), 256, 0, stream>>>(x, n, d, max_bits);
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

// insert `key` into the warp's sorted list (lane = rank, best first, k entries); all lanes hold the same key
__device__ __forceinline__ void warp_list_insert(unsigned long long& mine, unsigned long long key, int k, int lane) {
  const unsigned long long last = __shfl_sync(0xffffffffu, mine, k - 1);
  if (key <= last) return;                                            // warp-uniform
  const int pos = __popc(__ballot_sync(0xffffffffu, mine > key));
  const unsigned long long up = __shfl_up_sync(0xffffffffu, mine, 1);
  if (lane == pos) mine = key;
  else if (lane > pos) mine = up;
}

// cand_v / cand_i: [nq, nlists * kS1List] (each list best first, -1 = empty).  qn / qn_lo: |q|^2, |q - bf16(q)|^2.
// bank_max_bits[0..1]: max |r|^2, max |r - bf16(r)|^2 over the bank (float bits).  k <= 32.
__global__ void __launch_bounds__(kRrWarps * 32)
row_rescore_kernel(const float* __restrict__ Q, const float* __restrict__ bank, int d, int l2, int64_t nq,
                   const float* __restrict__ cand_v, const int32_t* __restrict__ cand_i, int nlists, int depth, int k,
                   const float* __restrict__ qn, const float* __restrict__ qn_lo, const unsigned int* __restrict__ bank_max_bits,
                   float* __restrict__ D, int64_t* __restrict__ I, int64_t id_offset, int* __restrict__ flags,
                   int* __restrict__ n_flagged) {
  extern __shared__ __align__(16) uint8_t rr_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sq = reinterpret_cast<float*>(rr_smem) + warp * d;                                   // this warp's query row
  uint32_t* ids = reinterpret_cast<uint32_t*>(rr_smem + static_cast<size_t>(kRrWarps) * d * 4) + warp * 32;
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
  // ---- margin (see the header): eps = |ql| Rmax + |qh| RLmax + d 2^-23 |q| Rmax, |qh| <= |q| + |ql|
  const float qnorm = sqrtf(qn[qrow]), qlo = sqrtf(qn_lo[qrow]);
  const float rmax = sqrtf(__uint_as_float(bank_max_bits[0])), rlmax = sqrtf(__uint_as_float(bank_max_bits[1]));
  float eps = (qlo * rmax + (qnorm + qlo) * rlmax + static_cast<float>(d) * 1.1920929e-7f * qnorm * rmax) * 1.001f;
  if (l2) eps = 2.0f * eps + 1e-6f * (qn[qrow] + __uint_as_float(bank_max_bits[0]));
  const float thr = found == k ? okey32_inv(static_cast<uint32_t>(prev >> 32)) - 2.0f * eps : -INFINITY;
  // ---- survivors, 32 candidates at a time: exact fp32 score, streamed into the warp's top-k (lane = rank).
  //      A full list whose last entry is inside the margin may have dropped a survivor: flag the query.
  bool overflow = false;
  unsigned long long mine = 0ull;
  __syncwarp();
  for (int c0 = 0; c0 < ncand; c0 += 32) {
    const int c = c0 + lane;
    const int32_t id = c < ncand ? ci[c] : -1;
    const bool take = id >= 0 && cv[c < ncand ? c : 0] >= thr;
    if (take && (c % kS1List) == depth - 1) overflow = true;
    const uint32_t m = __ballot_sync(0xffffffffu, take);
    if (m == 0u) continue;
    const int ns = __popc(m);
    if (take) ids[__popc(m & ((1u << lane) - 1u))] = static_cast<uint32_t>(id);
    __syncwarp();
    for (int s0 = 0; s0 < ns; s0 += 4) {
      const float* rp[4];
      uint32_t rid[4];
      float acc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        rid[u] = ids[min(s0 + u, ns - 1)];
        rp[u] = bank + static_cast<int64_t>(rid[u]) * d;
      }
      exact_rows_warp<4>(sq, rp, d, lane, l2 != 0, acc);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (s0 + u < ns) warp_list_insert(mine, ckey(keep_max ? acc[u] : -acc[u], rid[u]), k, lane);
    }
    __syncwarp();
  }
  overflow = __any_sync(0xffffffffu, overflow);
  if (lane < k) {
    if (mine != 0ull) {
      const float sc = okey32_inv(static_cast<uint32_t>(mine >> 32));
      D[qrow * k + lane] = keep_max ? sc : -sc;
      I[qrow * k + lane] = id_offset + static_cast<int64_t>(~static_cast<uint32_t>(mine & 0xFFFFFFFFull));
    } else {
      D[qrow * k + lane] = keep_max ? -FLT_MAX : FLT_MAX;
      I[qrow * k + lane] = -1;
    }
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
        warp_list_insert(mine, ckey(keep_max ? acc[u] : -acc[u], static_cast<uint32_t>(r0 + u)), k, lane);
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
// Work split of one search (see sim1_topk_kernel): CTA pairs launched and candidate-list slots per query row = the
// largest number of pairs whose tile range touches one pair-tile row.
void sim1_plan(int64_t nq, int64_t nr, int* pairs_out, int* slots_out) {
  const int64_t tm2 = (nq + 2 * kS1BM - 1) / (2 * kS1BM), tn = (nr + kS1BN - 1) / kS1BN;
  const int64_t T = tm2 * tn;
  int64_t P = device_sm_count() / 2;
  if (P > T) P = T;
  if (P < 1) P = 1;
  int64_t slots = 1;
  for (int64_t pm = 0; pm < tm2; ++pm) {
    const int64_t n = s1_pair_of(pm * tn + tn - 1, T, P) - s1_pair_of(pm * tn, T, P) + 1;
    if (n > slots) slots = n;
  }
  *pairs_out = static_cast<int>(P);
  *slots_out = static_cast<int>(slots);
}
int sim1_list_len() { return kS1List; }

// Qh [nq, dp] / Rh [nr, dp]: bf16 hi planes.  cand_v / cand_i: [nq, slots * 2 * kS1List] (pairs / slots from sim1_plan).
// Entries kept per candidate list: the cost of a running top-L is ~L (1 + ln(n / L)) insertions per list, and a list has to
// hold its share of a query's survivors (~2 k on random data)
int sim1_depth(int k) { return k <= 1 ? 8 : (k <= 3 ? 16 : kS1List); }

int sim1_topk(const void* Qh, const void* Rh, int64_t nq, int64_t nr, int dp, bool l2, const float* qn, const float* rn, int pairs,
              int slots, int depth, float* cand_v, int32_t* cand_i, cudaStream_t stream) {
  if (nq == 0 || nr == 0) return VSCB200_OK;
  VSCB_REQUIRE(dp % 8 == 0 && nr < (1ll << 31), "sim1_topk: dp must be a multiple of 8 and nr < 2^31");
  CUtensorMap tQ, tR;
  int rc;
  if ((rc = make_tmap_2d(&tQ, Qh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, nq, dp, dp, kS1BM, kS1BK, true))) return rc;
  if ((rc = make_tmap_2d(&tR, Rh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, nr, dp, dp, kS1BN / 2, kS1BK, true))) return rc;
  Sim1Params p = {};
  p.nq = nq; p.nr = nr; p.K = dp; p.l2 = l2 ? 1 : 0; p.qn = qn; p.rn = rn;
  const int64_t tm2 = (nq + 2 * kS1BM - 1) / (2 * kS1BM), tn = (nr + kS1BN - 1) / kS1BN;
  VSCB_REQUIRE(tm2 < (1ll << 30) && tn < (1ll << 30), "sim1_topk: too many tiles");
  p.tiles_m2 = static_cast<int>(tm2);
  p.tiles_n = static_cast<int>(tn);
  VSCB_REQUIRE(depth >= 1 && depth <= kS1List, "sim1_topk: bad list depth");
  p.pairs = pairs; p.slots = slots; p.depth = depth; p.cand_v = cand_v; p.cand_i = cand_i;
  // slots a pair never writes must read as empty
  VSCB_CUDA_OK(cudaMemsetAsync(cand_i, 0xFF, static_cast<size_t>(nq) * slots * 2 * kS1List * sizeof(int32_t), stream));
  VSCB_CUDA_OK(cudaFuncSetAttribute(sim1_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kS1Smem));
  ProfScope prof(kProfScores, stream, 2.0 * static_cast<double>(nq) * nr * dp);
  sim1_topk_kernel<<<2 * pairs, kS1Threads, kS1Smem, stream>>>(tQ, tR, p);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

// flags [nq] / n_flagged [1]: device scratch (n_flagged must be zero on entry; the fallback kernel reads it)
int sim1_rescore(const float* Q, const float* bank, int d, bool l2, int64_t nq, int64_t nr, const float* cand_v,
                 const int32_t* cand_i, int slots, int depth, int k, const float* qn, const float* qn_lo,
                 const unsigned int* bank_max_bits, float* D, int64_t* I, int64_t id_offset, int* flags, int* n_flagged,
                 cudaStream_t stream) {
  if (nq == 0) return VSCB200_OK;
  VSCB_REQUIRE(k >= 1 && k <= depth && depth <= kS1List, "sim1_rescore: k must be <= the candidate list depth");
  const size_t smem = static_cast<size_t>(kRrWarps) * d * 4 + static_cast<size_t>(kRrWarps) * 32 * 4;
  VSCB_REQUIRE(smem <= 200 * 1024 && d % 4 == 0, "sim1_rescore: dimension must be a multiple of 4 and <= 6144");
  VSCB_CUDA_OK(cudaFuncSetAttribute(row_rescore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  {
    ProfScope prof(kProfSelect, stream, static_cast<double>(nq) * slots * 2 * kS1List * 8);
    row_rescore_kernel<<<static_cast<unsigned>((nq + kRrWarps - 1) / kRrWarps), kRrWarps * 32, smem, stream>>>(
        Q, bank, d, l2 ? 1 : 0, nq, cand_v, cand_i, slots * 2, depth, k, qn, qn_lo, bank_max_bits, D, I, id_offset, flags, n_flagged);
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
