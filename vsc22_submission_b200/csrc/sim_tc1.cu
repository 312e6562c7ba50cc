// Batch top-k similarity search at ONE bf16 MMA per product (k <= 12): selection on approximate scores with a proven
// error margin, exact fp32 scores for the survivors.
//
//   sim1_topk_kernel     a ~ q.r from the hi planes only (bf16(q).bf16(r), fp32 accumulation) on tcgen05 with CTA pairs
//                        (cta_group::2, 256 x 256 tiles, the skeleton of gemm.cu).  Every CTA pair walks ONE contiguous
//                        range of the (query pair-tile, bank tile) sequence: perfectly balanced, and a query row meets as
//                        few pairs as possible (<= 3 at config 3).  Epilogue, THREAD = QUERY ROW (the layout tcgen05.ld
//                        delivers, no transpose), 16 columns per step with the next step's TMEM load in flight: four
//                        4-column maxima against the row's threshold tau, one warp-wide OR, one uniform branch; where
//                        some lane hit, the owners APPEND (key, column) to their row's 48-entry buffer in shared memory
//                        with predicated stores (count in a register, no atomics; a buffer below 40 entries cannot
//                        overflow within 8 columns, so there is no capacity test per append).  A buffer that reaches 40
//                        entries is compacted by the whole warp: a lower bound of its k-th best key by 7 bisection votes
//                        over the lanes' maxima (k = 1, 2: REDUX rounds; k = 1 also tightens tau on every append),
//                        tau = that - 2 eps, entries below tau dropped; a buffer still holding > 32 entries (a crowded
//                        margin) is moved to the query's list in global memory.  tau is published per query in global
//                        memory (atomicMax) and re-read once per tile, so the lists of one query held by different
//                        warps / CTA pairs tighten each other.  The accumulator is handed back to the MMA as soon as its
//                        last 16 columns are in registers.  No score reaches HBM.
//   row_rescore_kernel   one warp per query: A_k = k-th best approximate score over the row's lists; every bank row with
//                        a >= A_k - 2 eps is rescored in exact fp32 (exact.cuh, the one summation order every search
//                        path reports) and streamed into a warp-wide top-k by (score, lower id).
//   exact_row_topk_kernel  brute-force fp32 search of the flagged rows (a buffer saturated by entries inside the margin),
//                        normally none.
//
// Why this is exact.  Write q = qh + ql, r = rh + rl (h = bf16 rounding, l = its residual).  The kernel computes
// a = fl(qh.rh); s = q.r = qh.rh + ql.r + qh.rl, so by Cauchy-Schwarz |a - s| <= |ql| |r| + |qh| |rl| + d 2^-23 |q| |r|
// (last term: fp32 accumulation of the exact bf16 products).  |ql| and |qh| are measured per query, max |r| and
// max |rl| per bank at add(): eps(q) = |ql| Rmax + |qh| RLmax + d 2^-23 |q| Rmax -- about 0.0026 |q| Rmax in practice, below
// the worst case 2^-8.  The k rows with the best a have s >= A_k - eps, so the k-th best exact score s_k >= A_k - eps, and
// a row of the exact top-k has a >= s_k - eps >= A_k - 2 eps: it is a survivor.  A list only drops entries below
// (k-th best key of a SUBSET of the row's columns) - 2 eps <= A_k - 2 eps, so every survivor reaches the rescoring pass
// unless the query's global list overflows (several hundred entries inside the margin): such queries are flagged and
// searched exhaustively in fp32.  Reference semantics: faiss IndexFlat.search behind vsc/index.py:174 and
// vsc/baseline/score_normalization.py:93-98 (exact scores, best first, ties to the lower id).
#include <float.h>

#include "exact.cuh"
#include "host_util.h"
#include "kernels.h"
#include "ptx.cuh"
#include "margin.cuh"

namespace vscb200 {

constexpr int kS1BM = 128, kS1BN = 256, kS1BK = 64;     // per-CTA rows, pair-tile columns, K block
#ifndef S1_STAGES
#define S1_STAGES 4
#endif
#ifndef S1_CAP
#define S1_CAP 48
#endif
constexpr int kS1Stages = S1_STAGES;
constexpr int kS1EpiWarps = 8;
constexpr int kS1Threads = 128 + 32 * kS1EpiWarps;
constexpr int kS1ATile = kS1BM * kS1BK * 2;              // 16 KB
constexpr int kS1BTile = (kS1BN / 2) * kS1BK * 2;        // 16 KB: each CTA stages half of the bank tile
constexpr int kS1Cap = S1_CAP;                               // entries per candidate list = per (query row, column half, pair)
constexpr int kS1Keep = kS1Cap - 16;                     // a compaction that leaves more than this saturates the list
constexpr int kS1Trig = kS1Cap - 8;                      // a half-granule (8 columns) never overflows a buffer below this fill
constexpr int kS1BufBytes = kS1Cap * 32 * 8;             // one epilogue warp's 32 row buffers
constexpr int kS1Smem = kS1Stages * (kS1ATile + kS1BTile) + kS1EpiWarps * (kS1BufBytes + 128) + 256 + 1024;
constexpr int kS1MaxK = 12;
constexpr int kS1ListCap = 384;                          // entries of a query's global candidate list
static_assert(kS1Smem <= 232448, "sim1: shared memory");

struct Sim1Params {
  int64_t nq, nr;
  int K;                 // padded feature dim
  int d;                 // feature dim (error bound)
  int l2;
  int k;                 // neighbours wanted
  const float* qn;       // squared query norms
  const float* qn_lo;    // squared norms of q - bf16(q)
  const float* rn;       // squared bank norms (L2 keys)
  const unsigned int* bank_max_bits;    // max |r|^2, max |r - bf16(r)|^2 (float bits)
  int tiles_m2, tiles_n;            // pair tiles along M, 256-wide tiles along N
  int pairs;                        // CTA pairs launched
  uint2* cand;           // [nq, cand_cap] (selection key bits: larger = better, -distance for L2; bank row)
  int* cand_n;           // [nq] entries appended per query (zero on entry; may exceed cand_cap: the excess was dropped)
  int cand_cap;
  int* flags;            // [nq] set to 1 when a query's list overflowed cand_cap (zero on entry)
  unsigned int* tau_g;   // [nq] shared threshold per query as an ordered key (zero on entry = none)
};

// Work split: pair p owns tiles [tile_start(p), tile_start(p + 1)) of the sequence t = pm * tiles_n + n_blk.
__host__ __device__ inline int64_t s1_tile_start(int64_t p, int64_t T, int64_t P) { return p * T / P; }

__device__ __forceinline__ uint32_t okey32(float f) {          // order-preserving float -> uint (larger float = larger key)
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float okey32_inv(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}
__device__ __forceinline__ uint32_t ld_volatile_u32(const unsigned int* p) {      // asynchronous: the consumer waits, not the issue
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kS1Threads, 1)
sim1_topk_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmR, Sim1Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + kS1Stages * kS1ATile;
  uint8_t* buf_all = sB + kS1Stages * kS1BTile;            // per-warp candidate buffers of the epilogue
  float* lmax_all = reinterpret_cast<float*>(buf_all + kS1EpiWarps * kS1BufBytes);     // per-warp 32 floats (compaction)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(buf_all + kS1EpiWarps * (kS1BufBytes + 128));
  uint64_t* empty_bar = full_bar + kS1Stages;
  uint64_t* tfull_bar = empty_bar + kS1Stages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // uniform for the compiler
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
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

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
    if (__shfl_sync(0xffffffffu, crank, 0) == 0) {         // whole warp, warp-uniform operands, one elected lane issues
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
            umma_bf16_ss_cg2_warp(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_cg2_mcast_warp(&empty_bar[stage], static_cast<uint16_t>(0x3));
          if (++stage == kS1Stages) { stage = 0; phase ^= 1; }
        }
        umma_commit_cg2_mcast_warp(&tfull_bar[acc], static_cast<uint16_t>(0x3));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int ew = warp - 4, quad = warp & 3, half = ew >> 2;
    // buf[e][row]: entry e of the thread's row -- the owner's append address advances by one line (256 B) per entry; the
    // warp's cooperative reads of one row (compaction, lanes = entries) take the bank conflicts instead
    uint2* buf = reinterpret_cast<uint2*>(buf_all + ew * kS1BufBytes);
    float* lmax = lmax_all + ew * 32;
    auto slot_of = [&](int e, int row) { return e * 32 + row; };
    int acc = 0;
    uint32_t acc_phase = 0;
    float tau = INFINITY, eps2 = 0.f, qn_lane = 0.f;   // tau: keys <= tau cannot be survivors of this row
    int cnt = 0;                                        // entries in the buffer; compaction at kS1Trig
    int cur_pm = -1;
    int64_t row0 = 0;
    uint32_t g_next = 0u;
    float tau_pub = INFINITY;                          // the threshold this thread last published
    const uint32_t lt_mask = (1u << lane) - 1u;

    // The warp compacts row r's buffer: the k-th best key of the buffer minus the margin becomes the row's threshold.
    auto compact = [&](int r) {
      __syncwarp();
      const int n = __shfl_sync(0xffffffffu, cnt, r);
      const float e2 = __shfl_sync(0xffffffffu, eps2, r);
      const uint2 e0 = lane < n ? buf[slot_of(lane, r)] : make_uint2(0u, 0u);
      const uint2 e1 = lane + 32 < n ? buf[slot_of(lane + 32, r)] : make_uint2(0u, 0u);
      uint32_t a = lane < n ? okey32(__uint_as_float(e0.x)) : 0u;
      uint32_t b = lane + 32 < n ? okey32(__uint_as_float(e1.x)) : 0u;
      uint32_t kth = 0u;
      if (p.k <= 2) {                                   // k rounds of "largest remaining key"
        for (int i = 0; i < p.k; ++i) {
          kth = __reduce_max_sync(0xffffffffu, max(a, b));
          if (kth == 0u) break;
          const uint32_t ba = __ballot_sync(0xffffffffu, a == kth);
          if (ba != 0u) {
            if (lane == __ffs(ba) - 1) a = 0u;
          } else {
            const uint32_t bb = __ballot_sync(0xffffffffu, b == kth);
            if (lane == __ffs(bb) - 1) b = 0u;
          }
        }
      } else {
        // A sound lower bound of the buffer's k-th best key from 32 values: the lanes' maxima M = max(entry l, entry
        // l + 32) -- k distinct entries are at least as large as the k-th best M.  Bisection between min M and max M with
        // one vote per step keeps "at least k lanes above lo" invariant; 7 steps resolve 1/128 of the spread, far below
        // the margin.  (n >= kS1Trig >= 32: every lane holds an entry.)
        const float fm = fmaxf(__uint_as_float(e0.x), lane + 32 < n ? __uint_as_float(e1.x) : -INFINITY);
        const uint32_t km = okey32(fm);
        float lo = okey32_inv(__reduce_min_sync(0xffffffffu, km)), hi = okey32_inv(__reduce_max_sync(0xffffffffu, km));
        bool any_ok = false;                                          // lo itself has 31 lanes above it unless ties: verify
#pragma unroll
        for (int it = 0; it < 7; ++it) {
          const float mid = 0.5f * (lo + hi);
          const bool ok = __popc(__ballot_sync(0xffffffffu, fm > mid)) >= p.k;
          lo = ok ? mid : lo;
          hi = ok ? hi : mid;
          any_ok |= ok;
        }
        // no step succeeded (many equal maxima): fall back to "everything at or above the minimum"
        kth = any_ok ? okey32(lo) : (okey32(lo) - 1u);
      }
      // threshold of the row: this buffer's k-th best minus the margin, or what the query's other lists have reached
      const float tnew = fmaxf(kth != 0u ? okey32_inv(kth) - e2 : -INFINITY, __shfl_sync(0xffffffffu, tau, r));
      const bool keep0 = lane < n && __uint_as_float(e0.x) > tnew, keep1 = lane + 32 < n && __uint_as_float(e1.x) > tnew;
      const uint32_t m0 = __ballot_sync(0xffffffffu, keep0), m1 = __ballot_sync(0xffffffffu, keep1);
      const int n0 = __popc(m0), nn = n0 + __popc(m1);
      const bool spill = nn > kS1Keep;     // saturated by entries inside the margin: move them to the query's global list
      int64_t dst = 0;
      if (spill) {
        if (lane == 0) dst = atomicAdd(p.cand_n + row0 + r, nn);
        dst = __shfl_sync(0xffffffffu, dst, 0);
        uint2* out = p.cand + (row0 + r) * p.cand_cap;
        const int64_t i0 = dst + __popc(m0 & lt_mask), i1 = dst + n0 + __popc(m1 & lt_mask);
        if (keep0 && i0 < p.cand_cap) out[i0] = e0;
        if (keep1 && i1 < p.cand_cap) out[i1] = e1;
      } else {
        __syncwarp();
        if (keep0) buf[slot_of(__popc(m0 & lt_mask), r)] = e0;
        if (keep1) buf[slot_of(n0 + __popc(m1 & lt_mask), r)] = e1;
        __syncwarp();
      }
      if (lane == r) {
        const int64_t row = row0 + r;
        if (spill && dst + nn > p.cand_cap) p.flags[row] = 1;
        if (tnew > -INFINITY) atomicMax(p.tau_g + row, okey32(tnew));     // fire and forget; read back at the next tile
        tau = tnew;
        tau_pub = tnew;
        cnt = spill ? 0 : nn;
      }
    };
    // every row's buffer is appended to its query's global list (end of the pair's range over these rows)
    auto flush = [&]() {
      if (cur_pm < 0) return;
      __syncwarp();
      const int nmine = min(cnt, kS1Cap);
      int64_t dst_mine = 0;
      if (row0 + lane < p.nq && nmine > 0) {
        dst_mine = atomicAdd(p.cand_n + row0 + lane, nmine);
        if (dst_mine + nmine > p.cand_cap) p.flags[row0 + lane] = 1;
      }
#pragma unroll 1
      for (int r = 0; r < 32; ++r) {
        const int64_t row = row0 + r;
        if (row >= p.nq) break;
        const int n = __shfl_sync(0xffffffffu, nmine, r);
        const int64_t dst = __shfl_sync(0xffffffffu, dst_mine, r);
        uint2* out = p.cand + row * p.cand_cap;
        if (lane < n && dst + lane < p.cand_cap) out[dst + lane] = buf[slot_of(lane, r)];
        if (lane + 32 < n && dst + lane + 32 < p.cand_cap) out[dst + lane + 32] = buf[slot_of(lane + 32, r)];
      }
      __syncwarp();
    };
    // One granule = this thread's row x 16 columns.  Main path: four 4-column maxima against the row's threshold, one
    // warp-wide OR, one uniform branch.  Hit path: per 8 columns a uniform test per group some lane hit, the owners
    // append with predicated stores (a buffer below kS1Trig cannot overflow within 8 columns), then one vote.
    const bool k1 = p.k == 1;
    auto granule = [&](uint32_t (&v)[16], int64_t gcol, bool plain) {
      if (!plain) {                                                   // warp-uniform: L2 keys and / or the bank's last tile
        if (gcol >= p.nr) return;
        if (p.l2) {                                                   // key = -(|q|^2 + |r|^2 - 2 q.r)
#pragma unroll
          for (int j = 0; j < 16; ++j)
            v[j] = __float_as_uint(2.0f * __uint_as_float(v[j]) - __ldg(p.rn + min(gcol + j, p.nr - 1)) - qn_lane);
        }
        if (gcol + 16 > p.nr) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (gcol + j >= p.nr) v[j] = 0xFF800000u;                 // -inf
        }
      }
      uint32_t gm = 0u;                                               // groups of 4 columns with a key above the threshold
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float m4 = fmaxf(fmaxf(fmaxf(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1])), __uint_as_float(v[4 * g + 2])),
                               __uint_as_float(v[4 * g + 3]));
        gm |= m4 > tau ? (1u << g) : 0u;
      }
      const uint32_t um = __reduce_or_sync(0xffffffffu, gm);
      if (um == 0u) return;                                           // warp-uniform
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (((um >> (2 * h)) & 3u) == 0u) continue;                   // warp-uniform
#pragma unroll
        for (int g = 2 * h; g < 2 * h + 2; ++g) {
          if ((um >> g) & 1u) {                                       // warp-uniform
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int j = 4 * g + i;
              const float val = __uint_as_float(v[j]);
              if (val > tau) {
                buf[slot_of(cnt, lane)] = make_uint2(v[j], static_cast<uint32_t>(gcol + j));
                ++cnt;
                if (k1) tau = fmaxf(tau, val - eps2);                 // k = 1: the running threshold is exact
              }
            }
          }
        }
        uint32_t need = __ballot_sync(0xffffffffu, cnt >= kS1Trig);
        while (need != 0u) {
          const int r = __ffs(need) - 1;
          need &= need - 1u;
          compact(r);
        }
      }
    };
    for (int64_t t = t_begin; t < t_end; ++t) {
      const int pm = static_cast<int>(t / p.tiles_n), n_blk = static_cast<int>(t % p.tiles_n);
      if (pm != cur_pm) {
        flush();
        cur_pm = pm;
        row0 = static_cast<int64_t>(pm * 2 + static_cast<int>(crank)) * kS1BM + quad * 32;
        const bool live = row0 + lane < p.nq;
        qn_lane = live ? p.qn[row0 + lane] : 0.f;
        eps2 = live ? 2.0002f * sim1_eps(qn_lane, p.qn_lo[row0 + lane], p.bank_max_bits, p.d, p.l2) : 0.f;
        // the query's shared threshold: what sim1_boot_tau derived from a column sample (or other pairs have reached)
        const uint32_t g0 = live ? ld_volatile_u32(p.tau_g + row0 + lane) : 0u;
        tau = live ? (g0 != 0u ? okey32_inv(g0) : -INFINITY) : INFINITY;
        cnt = 0;
        g_next = 0u;
        tau_pub = tau;
      }
      // thresholds other warps / pairs have reached: the load was issued a tile ago (g_next), so its L2 latency is hidden
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      if (g_next != 0u && tau < INFINITY) tau = fmaxf(tau, okey32_inv(g_next));
      g_next = 0u;
      if (tau < INFINITY) g_next = ld_volatile_u32(p.tau_g + row0 + lane);
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kS1BN + half * (kS1BN / 2);
      const int64_t gcol_t = static_cast<int64_t>(n_blk) * kS1BN + half * (kS1BN / 2);
      const bool plain = !p.l2 && gcol_t + kS1BN / 2 <= p.nr;
      // the TMEM load of the next granule is in flight while this one is scanned
      uint32_t v[16], w[16];
      tmem_ld_32x16(t_addr, v);
      tmem_ld_wait();
#pragma unroll 1
      for (int gi = 0; gi < kS1BN / 2 / 16; ++gi) {
        if (gi + 1 < kS1BN / 2 / 16) {
          tmem_ld_32x16(t_addr + (gi + 1) * 16, w);
        } else {                                       // the whole tile is in registers: hand the accumulator back now
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(mapa_u32(smem_u32(&tempty_bar[acc]), 0u));
        }
        granule(v, gcol_t + gi * 16, plain);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = w[j];
      }
      if (tau > tau_pub) {                                            // k = 1 tightens without compacting: publish per tile
        atomicMax(p.tau_g + row0 + lane, okey32(tau));
        tau_pub = tau;
      }
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
// sq[row] = |q|^2 ; sq_lo[row] = |q - bf16(q)|^2 (optional) ; lo plane = bf16(q - hi) (optional: the split-bf16 kernels)
__global__ void __launch_bounds__(256)
q_hi_norm_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, float* __restrict__ sq,
                 float* __restrict__ sq_lo, int64_t n, int d, int dp, unsigned int* __restrict__ max_bits) {
  pdl_launch_dependents();                                // the streaming search launches its kernel behind this one early
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
    if (lo) lo[row * dp + c] = __float2bfloat16_rn(e);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    sl += __shfl_xor_sync(0xffffffffu, sl, o);
  }
  if (lane == 0) {
    sq[row] = s;
    if (sq_lo) sq_lo[row] = sl;
    if (max_bits) {     // bank rows: running maxima of |r|^2 and |r - bf16(r)|^2 (float bits; the values are >= 0).  Read
                        // first: after the first rows almost no row raises a maximum, and contended atomics serialise
      if (__float_as_uint(s) > *reinterpret_cast<volatile unsigned int*>(max_bits)) atomicMax(max_bits, __float_as_uint(s));
      if (__float_as_uint(sl) > *reinterpret_cast<volatile unsigned int*>(max_bits + 1)) atomicMax(max_bits + 1, __float_as_uint(sl));
    }
  }
}

// one pass over query (or bank) rows: hi plane (+ lo plane), squared norms (+ squared norms of the bf16 residual; + their
// running maxima for bank rows: everything add() needs in ONE read of the rows -- three kernels before: 126 us per
// 40k x 512 rows -> one)
int q_hi_norm(const float* x, void* hi, float* sq, float* sq_lo, int64_t n, int d, int dp, cudaStream_t stream, void* lo,
              unsigned int* max_bits) {
  if (n == 0) return VSCB200_OK;
  q_hi_norm_kernel<<<static_cast<unsigned>((n * 32 + 255) / 256), 256, 0, stream>>>(
      x, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), sq, sq_lo, n, d, dp, max_bits);
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

// cand [nq, cand_cap] / cand_n [nq]: the queries' candidate lists (unordered).  qn / qn_lo: |q|^2, |q - bf16(q)|^2.
// bank_max_bits[0..1]: max |r|^2, max |r - bf16(r)|^2 over the bank (float bits).  flags[q] != 0: the query's list
// overflowed in sim1_topk_kernel (the exhaustive kernel redoes the query).  k <= 32.
__global__ void __launch_bounds__(kRrWarps * 32)
row_rescore_kernel(const float* __restrict__ Q, const float* __restrict__ bank, int d, int l2, int64_t nq,
                   const uint2* __restrict__ cand, const int* __restrict__ cand_n, int cand_cap, int k,
                   const float* __restrict__ qn, const float* __restrict__ qn_lo, const unsigned int* __restrict__ bank_max_bits,
                   float* __restrict__ D, int64_t* __restrict__ I, int64_t id_offset, const int* __restrict__ flags,
                   int* __restrict__ n_flagged, int* __restrict__ flagged) {
  extern __shared__ __align__(16) uint8_t rr_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sq = reinterpret_cast<float*>(rr_smem) + warp * d;                                   // this warp's query row
  uint32_t* ids = reinterpret_cast<uint32_t*>(rr_smem + static_cast<size_t>(kRrWarps) * d * 4) + warp * 32;
  const int64_t qrow = static_cast<int64_t>(blockIdx.x) * kRrWarps + warp;
  if (qrow >= nq) return;
  const bool keep_max = !l2;
  const int ncand = min(cand_n[qrow], cand_cap);
  const uint2* ce = cand + qrow * cand_cap;
  for (int c = lane; c < d; c += 32) sq[c] = Q[qrow * d + c];
  // ---- A_k: k-th best approximate key over the list (k rounds of "best key below the previous one")
  unsigned long long prev = ~0ull;
  int found = 0;
  for (int r = 0; r < k; ++r) {
    unsigned long long best = 0ull;
    for (int c = lane; c < ncand; c += 32) {
      const uint2 e = ce[c];
      const unsigned long long key = ckey(__uint_as_float(e.x), e.y);
      if (key < prev && key > best) best = key;
    }
    best = warp_max_u64(best);
    if (best == 0ull) break;
    prev = best;
    ++found;
  }
  // ---- margin (see the header)
  const float eps = sim1_eps(qn[qrow], qn_lo[qrow], bank_max_bits, d, l2);
  const float thr = found == k ? okey32_inv(static_cast<uint32_t>(prev >> 32)) - 2.0f * eps : -INFINITY;
  // ---- survivors, 32 candidates at a time: exact fp32 score, streamed into the warp's top-k (lane = rank).
  const bool overflow = flags[qrow] != 0;                             // warp-uniform
  unsigned long long mine = 0ull;
  __syncwarp();
  for (int c0 = 0; c0 < ncand; c0 += 32) {
    const int c = c0 + lane;
    const uint2 e = c < ncand ? ce[c] : make_uint2(0u, 0u);
    const bool take = c < ncand && __uint_as_float(e.x) >= thr;
    const uint32_t m = __ballot_sync(0xffffffffu, take);
    if (m == 0u) continue;
    const int ns = __popc(m);
    if (take) ids[__popc(m & ((1u << lane) - 1u))] = e.y;
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
  if (lane == 0 && overflow) flagged[atomicAdd(n_flagged, 1)] = static_cast<int>(qrow);
}

// ------------------------------------------------------------------ exhaustive fp32 search of the flagged queries
// Work unit = (flagged query, one of kExSlices slices of the bank): every warp keeps its k best rows of the slice as a
// sorted list spread over its lanes (lane j = rank j), the warp lists are merged through shared memory into the unit's
// k best keys; the CTA that completes a query's last slice merges the kExSlices partial lists.  Flagged queries beyond
// the scratch capacity (f_max) are searched whole by one CTA each.  k <= 32.
constexpr int kExWarps = 8;
constexpr int kExSlices = 64;
constexpr int kExMaxFlagged = 1024;

// k best composite keys of bank rows [r_begin, r_end) for the query row in sq -> best[0..k) (shared memory), best first
__device__ __forceinline__ void exact_slice_topk(const float* sq, const float* __restrict__ bank, int d, bool l2, int64_t r_begin,
                                                 int64_t r_end, int k, unsigned long long* merged, unsigned long long* best) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool keep_max = !l2;
  unsigned long long mine = 0ull;                                     // lane j: the warp's j-th best composite key
  for (int64_t r0 = r_begin + static_cast<int64_t>(warp) * 4; r0 < r_end; r0 += kExWarps * 4) {
    const float* rp[4];
    float acc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) rp[u] = bank + (r0 + u < r_end ? r0 + u : r_end - 1) * d;
    exact_rows_warp<4>(sq, rp, d, lane, l2, acc);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (r0 + u >= r_end) continue;
      warp_list_insert(mine, ckey(keep_max ? acc[u] : -acc[u], static_cast<uint32_t>(r0 + u)), k, lane);
    }
  }
  merged[warp * 32 + lane] = lane < k ? mine : 0ull;
  __syncthreads();
  if (warp == 0) {
    unsigned long long prev = ~0ull;
    for (int r = 0; r < k; ++r) {
      unsigned long long b = 0ull;
      for (int c = lane; c < kExWarps * 32; c += 32) {
        const unsigned long long key = merged[c];
        if (key < prev && key > b) b = key;
      }
      b = warp_max_u64(b);
      if (lane == 0) best[r] = b;
      prev = b;
      if (b == 0ull) prev = 0ull;                                     // nothing left: the remaining entries stay empty
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void exact_write_result(unsigned long long key, bool keep_max, float* D, int64_t* I, int64_t at,
                                                   int64_t id_offset) {
  if (key != 0ull) {
    const float sc = okey32_inv(static_cast<uint32_t>(key >> 32));
    D[at] = keep_max ? sc : -sc;
    I[at] = id_offset + static_cast<int64_t>(~static_cast<uint32_t>(key & 0xFFFFFFFFull));
  } else {
    D[at] = keep_max ? -FLT_MAX : FLT_MAX;
    I[at] = -1;
  }
}

// flagged [n_flagged]: query rows to redo.  partial [f_max, kExSlices, k] keys, done [f_max] (zero on entry).
__global__ void __launch_bounds__(kExWarps * 32)
exact_flagged_kernel(const float* __restrict__ Q, const float* __restrict__ bank, int d, int l2, int64_t nr, int k,
                     float* __restrict__ D, int64_t* __restrict__ I, int64_t id_offset, const int* __restrict__ flagged,
                     const int* __restrict__ n_flagged, unsigned long long* __restrict__ partial, int* __restrict__ done, int f_max) {
  extern __shared__ __align__(16) uint8_t ex_smem[];
  const int nf_all = *n_flagged;
  if (nf_all == 0) return;
  float* sq = reinterpret_cast<float*>(ex_smem);                                              // [d]
  unsigned long long* merged = reinterpret_cast<unsigned long long*>(ex_smem + static_cast<size_t>(d) * 4);   // [kExWarps * 32]
  unsigned long long* best = merged + kExWarps * 32;                                          // [32]
  __shared__ int s_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool keep_max = !l2;
  const int nf = min(nf_all, f_max);
  for (int64_t u = blockIdx.x; u < static_cast<int64_t>(nf) * kExSlices; u += gridDim.x) {
    const int f = static_cast<int>(u / kExSlices), sl = static_cast<int>(u % kExSlices);
    const int64_t qrow = flagged[f];
    __syncthreads();
    for (int c = threadIdx.x; c < d; c += kExWarps * 32) sq[c] = Q[qrow * d + c];
    __syncthreads();
    exact_slice_topk(sq, bank, d, l2 != 0, nr * sl / kExSlices, nr * (sl + 1) / kExSlices, k, merged, best);
    unsigned long long* mine = partial + (static_cast<int64_t>(f) * kExSlices + sl) * k;
    if (threadIdx.x < k) mine[threadIdx.x] = best[threadIdx.x];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(done + f, 1) == kExSlices - 1 ? 1 : 0;
    __syncthreads();
    if (s_last && warp == 0) {                                        // this CTA completed the query: merge its slices
      __threadfence();
      const unsigned long long* all = partial + static_cast<int64_t>(f) * kExSlices * k;
      unsigned long long prev = ~0ull;
      for (int r = 0; r < k; ++r) {
        unsigned long long b = 0ull;
        for (int c = lane; c < kExSlices * k; c += 32) {
          const unsigned long long key = *reinterpret_cast<const volatile unsigned long long*>(all + c);
          if (key < prev && key > b) b = key;
        }
        b = warp_max_u64(b);
        if (lane == 0) exact_write_result(b, keep_max, D, I, qrow * k + r, id_offset);
        prev = b;
      }
    }
  }
  for (int f = f_max + blockIdx.x; f < nf_all; f += gridDim.x) {      // beyond the scratch capacity: one CTA per query
    const int64_t qrow = flagged[f];
    __syncthreads();
    for (int c = threadIdx.x; c < d; c += kExWarps * 32) sq[c] = Q[qrow * d + c];
    __syncthreads();
    exact_slice_topk(sq, bank, d, l2 != 0, 0, nr, k, merged, best);
    if (threadIdx.x < k) exact_write_result(best[threadIdx.x], keep_max, D, I, qrow * k + threadIdx.x, id_offset);
  }
}

// ------------------------------------------------------------------ host side
// Work split of one search (see sim1_topk_kernel): CTA pairs launched.
int sim1_pairs(int64_t nq, int64_t nr) {
  const int64_t tm2 = (nq + 2 * kS1BM - 1) / (2 * kS1BM), tn = (nr + kS1BN - 1) / kS1BN;
  const int64_t T = tm2 * tn;
  int64_t P = device_sm_count() / 2;
  if (P > T) P = T;
  if (P < 1) P = 1;
  return static_cast<int>(P);
}
int sim1_list_cap() { return kS1ListCap; }
int sim1_max_k() { return kS1MaxK; }

// ------------------------------------------------------------------ threshold bootstrap from a column sample
// S [nq, ncs]: one-pass scores of every query against a sample of ncs <= 1024 bank rows (sim_tc.cu, strided tensor map).
// The k-th best of a SUBSET of the bank cannot exceed the k-th best of the bank, so tau0 = (k-th best of the sample) - 2 eps
// is a valid initial threshold for the query: sim1_topk_kernel starts every list of the row from it instead of from
// -infinity (without it ~20 appends per 1024 scores while the thresholds warm up, per (row, pair) list).
// One warp per query row: the row in registers (32 per lane), k rounds of warp arg-max.
__global__ void __launch_bounds__(256)
sim1_boot_tau_kernel(const float* __restrict__ S, int64_t nq, int ncs, int k, const float* __restrict__ qn,
                     const float* __restrict__ qn_lo, const unsigned int* __restrict__ bank_max_bits, int d,
                     unsigned int* __restrict__ tau_g) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= nq) return;
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = i * 32 + lane < ncs ? S[row * ncs + i * 32 + lane] : -INFINITY;
  float kth = -INFINITY;
  for (int r = 0; r < k; ++r) {
    float lm = v[0];
#pragma unroll
    for (int i = 1; i < 32; ++i) lm = fmaxf(lm, v[i]);
    const uint32_t best = __reduce_max_sync(0xffffffffu, okey32(lm));
    const uint32_t owners = __ballot_sync(0xffffffffu, okey32(lm) == best);
    if (lane == __ffs(owners) - 1) {                      // one instance leaves the row
      bool done = false;
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (!done && okey32(v[i]) == best) { v[i] = -INFINITY; done = true; }
    }
    kth = okey32_inv(best);
  }
  if (lane == 0 && kth > -INFINITY) {
    const float tau0 = kth - 2.0002f * sim1_eps(qn[row], qn_lo[row], bank_max_bits, d, 0);
    tau_g[row] = okey32(tau0);
  }
}

// scratch: the zeroed scratch of sim1_topk (the thresholds live at scratch + nq + 1).  Inner-product metric only.
int sim1_boot_tau(const float* S, int64_t nq, int ncs, int k, const float* qn, const float* qn_lo,
                  const unsigned int* bank_max_bits, int d, int* scratch, cudaStream_t stream) {
  if (nq == 0) return VSCB200_OK;
  VSCB_REQUIRE(ncs >= k && ncs <= 1024 && k >= 1, "sim1_boot_tau: need k <= ncs <= 1024");
  sim1_boot_tau_kernel<<<static_cast<unsigned>((nq * 32 + 255) / 256), 256, 0, stream>>>(
      S, nq, ncs, k, qn, qn_lo, bank_max_bits, d, reinterpret_cast<unsigned int*>(scratch + nq + 1));
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

// Qh [nq, dp] / Rh [nr, dp]: bf16 hi planes.  cand: [nq, sim1_list_cap()] entries of 8 bytes.
// qn / qn_lo [nq]: |q|^2, |q - bf16(q)|^2.  scratch: see sim1_rescore (zero on entry).
int sim1_topk(const void* Qh, const void* Rh, int64_t nq, int64_t nr, int d, int dp, bool l2, int k, const float* qn,
              const float* qn_lo, const float* rn, const unsigned int* bank_max_bits, void* cand, int* scratch,
              cudaStream_t stream) {
  if (nq == 0 || nr == 0) return VSCB200_OK;
  VSCB_REQUIRE(dp % 8 == 0 && nr < (1ll << 31), "sim1_topk: dp must be a multiple of 8 and nr < 2^31");
  CUtensorMap tQ, tR;
  int rc;
  if ((rc = make_tmap_2d(&tQ, Qh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, nq, dp, dp, kS1BM, kS1BK, true))) return rc;
  if ((rc = make_tmap_2d(&tR, Rh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, nr, dp, dp, kS1BN / 2, kS1BK, true))) return rc;
  Sim1Params p = {};
  VSCB_REQUIRE(k >= 1 && k <= kS1MaxK, "sim1_topk: k out of range");
  p.nq = nq; p.nr = nr; p.K = dp; p.d = d; p.l2 = l2 ? 1 : 0; p.k = k; p.qn = qn; p.qn_lo = qn_lo; p.rn = rn;
  p.bank_max_bits = bank_max_bits; p.flags = scratch; p.tau_g = reinterpret_cast<unsigned int*>(scratch + nq + 1);
  const int64_t tm2 = (nq + 2 * kS1BM - 1) / (2 * kS1BM), tn = (nr + kS1BN - 1) / kS1BN;
  VSCB_REQUIRE(tm2 < (1ll << 30) && tn < (1ll << 30), "sim1_topk: too many tiles");
  p.tiles_m2 = static_cast<int>(tm2);
  p.tiles_n = static_cast<int>(tn);
  p.pairs = sim1_pairs(nq, nr);
  p.cand = reinterpret_cast<uint2*>(cand); p.cand_n = scratch + 2 * nq + 1; p.cand_cap = kS1ListCap;
  VSCB_CUDA_OK(cudaFuncSetAttribute(sim1_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kS1Smem));
  ProfScope prof(kProfScores, stream, 2.0 * static_cast<double>(nq) * nr * dp);
  sim1_topk_kernel<<<2 * p.pairs, kS1Threads, kS1Smem, stream>>>(tQ, tR, p);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

int sim1_scratch_ints(int64_t nq) { return static_cast<int>(4 * nq + 1 + kExMaxFlagged); }
size_t sim1_partial_bytes() { return static_cast<size_t>(kExMaxFlagged) * kExSlices * kS1MaxK * sizeof(unsigned long long); }

// scratch [sim1_scratch_ints(nq)] (zero on entry, shared with sim1_topk): flags [nq] | flagged count [1] | shared
// thresholds [nq] | list lengths [nq] | flagged query rows [nq] | slice counters [kExMaxFlagged].
// partial: sim1_partial_bytes() of device scratch for the exhaustive pass.
int sim1_rescore(const float* Q, const float* bank, int d, bool l2, int64_t nq, int64_t nr, const void* cand, int k,
                 const float* qn, const float* qn_lo, const unsigned int* bank_max_bits, float* D, int64_t* I, int64_t id_offset,
                 int* scratch, void* partial, cudaStream_t stream) {
  if (nq == 0) return VSCB200_OK;
  VSCB_REQUIRE(k >= 1 && k <= kS1MaxK, "sim1_rescore: k out of range");
  VSCB_REQUIRE(nq < (1ll << 29), "sim1_rescore: too many query rows in one block");
  const int* flags = scratch;
  int* n_flagged = scratch + nq;
  const int* cand_n = scratch + 2 * nq + 1;
  int* flagged = scratch + 3 * nq + 1;
  int* done = scratch + 4 * nq + 1;
  const size_t smem = static_cast<size_t>(kRrWarps) * d * 4 + static_cast<size_t>(kRrWarps) * 32 * 4;
  VSCB_REQUIRE(smem <= 200 * 1024 && d % 4 == 0, "sim1_rescore: dimension must be a multiple of 4 and <= 6144");
  VSCB_CUDA_OK(cudaFuncSetAttribute(row_rescore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  {
    ProfScope prof(kProfSelect, stream, static_cast<double>(nq) * 128 * 8);
    row_rescore_kernel<<<static_cast<unsigned>((nq + kRrWarps - 1) / kRrWarps), kRrWarps * 32, smem, stream>>>(
        Q, bank, d, l2 ? 1 : 0, nq, reinterpret_cast<const uint2*>(cand), cand_n, kS1ListCap, k, qn, qn_lo, bank_max_bits, D, I,
        id_offset, flags, n_flagged, flagged);
    count_launch();
  }
  const size_t smem_ex = static_cast<size_t>(d) * 4 + static_cast<size_t>(kExWarps) * 32 * 8 + 32 * 8;
  VSCB_CUDA_OK(cudaFuncSetAttribute(exact_flagged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_ex)));
  {
    ProfScope prof(kProfSelect, stream, 0.0);
    exact_flagged_kernel<<<4 * device_sm_count(), kExWarps * 32, smem_ex, stream>>>(
        Q, bank, d, l2 ? 1 : 0, nr, k, D, I, id_offset, flagged, n_flagged, reinterpret_cast<unsigned long long*>(partial), done,
        kExMaxFlagged);
    count_launch();
  }
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

}  // namespace vscb200
