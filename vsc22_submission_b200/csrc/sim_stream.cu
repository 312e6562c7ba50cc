// Streaming similarity search: a handful of query rows (nq <= 128) against a large resident bank -- the
// reference's real call pattern (one index.search per query VIDEO: score_normalization.py:93-98, ~40 rows
// against 1.26 M bank rows, 8 295 times; M/infer/infer_matching.py:232).  The bank is read from HBM exactly
// once per call and nothing of size nq x nr is ever written: this form is HBM-bound (20 FLOP/B), its
// roofline is nr * d * 4 bytes / HBM bandwidth.
//
// Roles are swapped with respect to sim_tc.cu: the BANK tile (128 rows) is the M operand and the query
// block (Npad = nq rounded up to 16) the N operand, so the tensor work per bank byte scales with nq instead
// of being padded to 128 query rows.  Same fp32-equivalent 3-MMA bf16 split (hi.hi + lo.hi + hi.lo) and
// two K segments as sim_tc.cu.  The accumulator tile has one bank row per TMEM lane / epilogue thread and
// one query per column; the epilogue reduces every column over the 32 rows of a warp with one REDUX
// (max of order-preserving integer keys) and stores one float per (32-row group, query):
//     gmax[group][q] = max over the group's rows of the selection key (score, or -squared distance)
// 1/32 of a float per pair.  The top-k rows of a query lie in the <= k groups with the largest maxima
// (each such group holds at least one of the k best rows), so search() selects k + slack groups per
// query and rescores their 32 rows each exactly in fp32 -- bit-exact ids, fp32 scores, no dense score block.
// The selection and rescoring run as one CTA per 256-group chunk (group_topk_kernel) and one CTA per (query, group)
// (group_rescore_few_kernel, last arriver of a query sorts); group_rescore_kernel (one CTA per query) serves the
// many-query callers.  The launches of a call are chained by programmatic dependent launch (griddepcontrol).
#include <float.h>

#include "exact.cuh"
#include "host_util.h"
#include "kernels.h"
#include "ptx.cuh"

namespace vscb200 {

constexpr int kSsThreads = 256;       // warp 0 TMA, warp 1 MMA, warp 2 TMEM alloc, warps 4-7 epilogue
constexpr int kSsMaxStages = 6;
constexpr int kSsTile = 128 * 64 * 2;  // one 128x64 bf16 bank tile: 16 KB

struct StreamParams {
  float* gmax;          // [groups][Npad]
  int64_t nr;
  int nq, Npad, K, l2, stages, tiles;
  const float* qn;
  const float* rn;
};

// float bits <-> int whose signed order equals the float order (an involution on the bit pattern)
__device__ __forceinline__ int ordered_int(int i) { return i ^ ((i >> 31) & 0x7FFFFFFF); }

__global__ void __launch_bounds__(kSsThreads, 1)
sim_stream_kernel(const __grid_constant__ CUtensorMap tmRh, const __grid_constant__ CUtensorMap tmRl,
                  const __grid_constant__ CUtensorMap tmQh, const __grid_constant__ CUtensorMap tmQl, StreamParams p) {
  extern __shared__ uint8_t smem_raw[];
  pdl_launch_dependents();                               // the selection kernel may be scheduled early; it waits for this grid
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const int q_tile = p.Npad * 128;                       // Npad rows x 64 bf16
  const int stage_bytes = 2 * kSsTile + 2 * q_tile;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + kSsMaxStages;
  uint64_t* tfull_bar = empty_bar + kSsMaxStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmRh); prefetch_tmap(&tmRl); prefetch_tmap(&tmQh); prefetch_tmap(&tmQl);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 4); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                            // the query planes and norms come from the kernel before this one

  const int kblocks = (p.K + 63) / 64;
  const int nseg = kblocks < 2 ? 1 : 2;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], stage_bytes);
          uint8_t* st = smem + stage * stage_bytes;
          tma_load_2d(st, &tmRh, &full_bar[stage], kb * 64, tile * 128, kEvictFirst);             // streamed once
          tma_load_2d(st + kSsTile, &tmRl, &full_bar[stage], kb * 64, tile * 128, kEvictFirst);
          tma_load_2d(st + 2 * kSsTile, &tmQh, &full_bar[stage], kb * 64, 0, kEvictLast);         // L2 resident
          tma_load_2d(st + 2 * kSsTile + q_tile, &tmQl, &full_bar[stage], kb * 64, 0, kEvictLast);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16_f32(128, p.Npad);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < kblocks; ++kb) {
          const int seg = (nseg == 2 && kb >= kblocks / 2) ? 1 : 0;
          const bool fresh = kb == 0 || (nseg == 2 && kb == kblocks / 2);
          const uint32_t d_tmem = tmem_base + (acc * 2 + seg) * 128;
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t st = smem_u32(smem + stage * stage_bytes);
          const uint64_t rh = make_desc_k_sw128(st), rl = make_desc_k_sw128(st + kSsTile);
          const uint64_t qh = make_desc_k_sw128(st + 2 * kSsTile), ql = make_desc_k_sw128(st + 2 * kSsTile + q_tile);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_bf16_ss(d_tmem, rh + 2 * k, qh + 2 * k, idesc, (fresh && k == 0) ? 0u : 1u);
            umma_bf16_ss(d_tmem, rl + 2 * k, qh + 2 * k, idesc, 1u);
            umma_bf16_ss(d_tmem, rh + 2 * k, ql + 2 * k, idesc, 1u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int quad = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
      const int64_t row = static_cast<int64_t>(tile) * 128 + quad * 32 + lane;     // this thread's bank row
      const bool row_ok = row < p.nr;
      const float rn = (p.l2 && row_ok) ? p.rn[row] : 0.f;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * 256;
      float mine[4] = {-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};      // column 32*i + lane of this warp's group
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c0 = 32 * i + 16 * h;
          if (c0 < p.Npad) {                              // warp-uniform
            uint32_t v[16], w[16];
            tmem_ld_32x16(taddr + c0, v);
            if (nseg == 2) tmem_ld_32x16(taddr + 128 + c0, w);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float s = __uint_as_float(v[j]);
              if (nseg == 2) s += __uint_as_float(w[j]);
              if (p.l2) s = -fmaxf(p.qn[min(c0 + j, p.nq - 1)] + rn - 2.0f * s, 0.f);
              const int key = row_ok ? ordered_int(__float_as_int(s)) : static_cast<int>(0x80000000u);
              const int best = __reduce_max_sync(0xffffffffu, key);
              if (16 * h + j == lane) mine[i] = __int_as_float(ordered_int(best));
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      const int64_t group = static_cast<int64_t>(tile) * 4 + quad;
      if (group * 32 < p.nr) {
        float* dst = p.gmax + group * p.Npad;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (32 * i + lane < p.Npad) dst[32 * i + lane] = mine[i];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

// gmax: [groups][Npad]; planes as in sim_tc.cu.  nq <= 128.
int sim_stream_groupmax(const void* Qh, const void* Ql, const void* Rh, const void* Rl, int64_t nq, int64_t nr, int dp,
                        bool l2, const float* qn, const float* rn, float* gmax, int Npad, cudaStream_t stream) {
  VSCB_REQUIRE(nq >= 1 && nq <= 128 && Npad % 16 == 0 && Npad >= nq && Npad <= 128, "sim_stream: nq must be in [1, 128]");
  VSCB_REQUIRE(dp % 8 == 0 && nr > 0 && nr < (1ll << 31) - 256, "sim_stream: bad bank shape");
  CUtensorMap tRh, tRl, tQh, tQl;
  int rc;
  if ((rc = make_tmap_2d(&tRh, Rh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, nr, dp, dp, 128, 64, true))) return rc;
  if ((rc = make_tmap_2d(&tRl, Rl, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, nr, dp, dp, 128, 64, true))) return rc;
  if ((rc = make_tmap_2d(&tQh, Qh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, nq, dp, dp, Npad, 64, true))) return rc;
  if ((rc = make_tmap_2d(&tQl, Ql, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, nq, dp, dp, Npad, 64, true))) return rc;
  StreamParams p = {};
  p.gmax = gmax; p.nr = nr; p.nq = static_cast<int>(nq); p.Npad = Npad; p.K = dp; p.l2 = l2 ? 1 : 0;
  p.qn = qn; p.rn = rn;
  p.tiles = static_cast<int>((nr + 127) / 128);
  const int stage_bytes = 2 * kSsTile + 2 * Npad * 128;
  int stages = (220 * 1024) / stage_bytes;
  if (stages > kSsMaxStages) stages = kSsMaxStages;
  p.stages = stages;
  const int smem = stages * stage_bytes + 512 + 1024;
  VSCB_CUDA_OK(cudaFuncSetAttribute(sim_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int grid = p.tiles < device_sm_count() ? p.tiles : device_sm_count();
  ProfScope prof(kProfScores, stream, 2.0 * static_cast<double>(nq) * nr * dp);
  VSCB_CUDA_OK(launch_pdl_chain(sim_stream_kernel, dim3(grid), dim3(kSsThreads), static_cast<size_t>(smem), stream, tRh, tRl, tQh, tQl, p));
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

// ------------------------------------------------------------------ per-chunk top groups
// gmax [G][Npad] -> for every (query, chunk of 256 groups) the kg best (key, group id) pairs, best first, ties to
// the lower group id.  The chunk is staged through shared memory with coalesced loads; one warp per query holds
// the query's 256 keys in registers (8 per lane) and extracts kg maxima with two REDUX per round.
constexpr int kGtChunk = 256;
constexpr int kGtThreads = 1024;    // 32 warps: a query per warp, at most two rounds of queries at nq <= 64

__global__ void __launch_bounds__(kGtThreads)
group_topk_kernel(const float* __restrict__ gmax, int64_t G, int Npad, int nq, int kg, int chunks,
                  float* __restrict__ cand_v, int32_t* __restrict__ cand_g) {
  extern __shared__ float gt_tile[];                      // [kGtChunk][Npad + 1]
  pdl_launch_dependents();
  pdl_wait();                                             // gmax comes from the kernel before this one in the stream
  const int chunk = blockIdx.x;
  const int64_t g0 = static_cast<int64_t>(chunk) * kGtChunk;
  const int ng = static_cast<int>(G - g0 < kGtChunk ? G - g0 : kGtChunk);
  const int ld = Npad + 1;
  const float* src = gmax + g0 * Npad;
  for (int idx = threadIdx.x; idx < ng * Npad; idx += kGtThreads) {
    const int r = idx / Npad, c = idx - r * Npad;
    gt_tile[r * ld + c] = src[idx];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kNone = 0x7FFFFFFF;
  for (int q = warp; q < nq; q += kGtThreads / 32) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (i * 32 + lane < ng) ? gt_tile[(i * 32 + lane) * ld + q] : -INFINITY;
    float* ov = cand_v + (static_cast<int64_t>(q) * chunks + chunk) * kg;
    int32_t* og = cand_g + (static_cast<int64_t>(q) * chunks + chunk) * kg;
    for (int r = 0; r < kg; ++r) {
      float lm = v[0];
#pragma unroll
      for (int i = 1; i < 8; ++i) lm = fmaxf(lm, v[i]);
      const int best = __reduce_max_sync(0xffffffffu, ordered_int(__float_as_int(lm)));
      int mine = kNone;
#pragma unroll
      for (int i = 7; i >= 0; --i)
        if (ordered_int(__float_as_int(v[i])) == best) mine = i * 32 + lane;
      const int win = __reduce_min_sync(0xffffffffu, mine);
      const float bv = __int_as_float(ordered_int(best));
      const bool none = bv == -INFINITY;
      if (lane == 0) {
        ov[r] = bv;
        og[r] = none ? -1 : static_cast<int32_t>(g0 + win);
      }
      if (!none && mine == win) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (i == (win >> 5)) v[i] = -INFINITY;
      }
    }
  }
}

int group_topk(const float* gmax, int64_t G, int Npad, int nq, int kg, int chunks, float* cand_v, int32_t* cand_g,
               cudaStream_t stream) {
  const size_t smem = static_cast<size_t>(kGtChunk) * (Npad + 1) * sizeof(float);
  VSCB_CUDA_OK(cudaFuncSetAttribute(group_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  ProfScope prof(kProfSelect, stream, static_cast<double>(G) * Npad * 4);
  VSCB_CUDA_OK(launch_pdl_chain(group_topk_kernel, dim3(chunks), dim3(kGtThreads), smem, stream, gmax, G, Npad, nq, kg, chunks, cand_v, cand_g));
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}
int group_topk_chunks(int64_t G) { return static_cast<int>((G + kGtChunk - 1) / kGtChunk); }

// ------------------------------------------------------------------ exact rescoring of the selected groups
// cand_v / cand_g: [nq, ncg] group keys and ids (-1 = none) from group_topk_kernel (ncg = chunks * kg), or
// cand_v == nullptr and gsel64 [nq, kg] already selected.  Per query: keep the kg best groups (key, lower id
// first), exact fp32 scores of their 32 rows each (one warp per row, coalesced 128-bit reads), sort (score,
// lower id first), emit k.
constexpr int kGrThreads = 512;

__device__ __forceinline__ uint32_t okey_u(float f, bool keep_max) {
  const uint32_t u = __float_as_uint(f);
  const uint32_t k = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return keep_max ? k : ~k;
}

__global__ void __launch_bounds__(kGrThreads)
group_rescore_kernel(const float* __restrict__ Q, const float* __restrict__ bank, int d, int l2, int64_t nr,
                     const int64_t* __restrict__ gsel64, const float* __restrict__ cand_v,
                     const int32_t* __restrict__ cand_g, int ncg, int gpad, int kg, int gs, int cpad, int k,
                     float* __restrict__ D, int64_t* __restrict__ I, int64_t id_offset) {
  extern __shared__ unsigned long long gr_smem[];        // [max(cpad, gpad)] composite keys, [d] floats, [kg] group ids
  unsigned long long* cand = gr_smem;
  const int csz = cpad > gpad ? cpad : gpad;
  float* sq = reinterpret_cast<float*>(gr_smem + csz);
  int32_t* gsel = reinterpret_cast<int32_t*>(sq + d);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t qrow = blockIdx.x;
  const bool keep_max = !l2;
  for (int c = tid; c < d; c += kGrThreads) sq[c] = Q[qrow * d + c];
  if (cand_v != nullptr) {
    // ---- the kg best groups of this query over all chunks
    for (int i = tid; i < gpad; i += kGrThreads) {
      unsigned long long c = 0ull;
      if (i < ncg) {
        const int32_t g = cand_g[qrow * ncg + i];
        if (g >= 0) c = (static_cast<unsigned long long>(okey_u(cand_v[qrow * ncg + i], true)) << 32) |
                        static_cast<uint32_t>(~static_cast<uint32_t>(g));
      }
      cand[i] = c;
    }
    __syncthreads();
    for (int size = 2; size <= gpad; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int i = tid; i < (gpad >> 1); i += kGrThreads) {
          const int lo = 2 * i - (i & (stride - 1));
          const int hi = lo + stride;
          const bool desc = (lo & size) == 0;
          const unsigned long long a = cand[lo], b = cand[hi];
          if ((a < b) == desc) { cand[lo] = b; cand[hi] = a; }
        }
        __syncthreads();
      }
    }
    if (tid < kg) gsel[tid] = cand[tid] != 0ull ? static_cast<int32_t>(~static_cast<uint32_t>(cand[tid] & 0xFFFFFFFFull)) : -1;
  } else {
    if (tid < kg) gsel[tid] = static_cast<int32_t>(gsel64[qrow * kg + tid]);
  }
  __syncthreads();
  for (int i = tid; i < csz; i += kGrThreads) cand[i] = 0ull;
  __syncthreads();
  const int ncand = kg * gs;                    // gs rows per group
  // four rows per warp per round: their loads are independent, so four HBM latencies overlap instead of queueing
  for (int c0 = warp * 4; c0 < ncand; c0 += (kGrThreads / 32) * 4) {
    const float* rp[4];
    int64_t ids[4];
    bool ok[4];
    float acc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = c0 + u;
      const int64_t g = c < ncand ? gsel[c / gs] : -1;
      ids[u] = g * gs + (c % gs);
      ok[u] = g >= 0 && ids[u] < nr;
      rp[u] = bank + (ok[u] ? ids[u] : 0) * d;
    }
    exact_rows_warp<4>(sq, rp, d, lane, l2 != 0, acc);       // the shared summation order (exact.cuh)
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (lane == 0 && ok[u])
        cand[c0 + u] = (static_cast<unsigned long long>(okey_u(acc[u], keep_max)) << 32) |
                       static_cast<uint32_t>(~static_cast<uint32_t>(ids[u]));
    }
  }
  __syncthreads();
  for (int size = 2; size <= cpad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < (cpad >> 1); i += kGrThreads) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const unsigned long long a = cand[lo], b = cand[hi];
        if ((a < b) == desc) { cand[lo] = b; cand[hi] = a; }
      }
      __syncthreads();
    }
  }
  for (int j = tid; j < k; j += kGrThreads) {
    const unsigned long long c = cand[j];
    if (c != 0ull) {
      const uint32_t key = static_cast<uint32_t>(c >> 32);
      const uint32_t ok = keep_max ? key : ~key;
      const uint32_t u = (ok & 0x80000000u) ? (ok ^ 0x80000000u) : ~ok;
      D[qrow * k + j] = __uint_as_float(u);
      I[qrow * k + j] = id_offset + static_cast<int64_t>(~static_cast<uint32_t>(c & 0xFFFFFFFFull));
    } else {
      D[qrow * k + j] = keep_max ? -FLT_MAX : FLT_MAX;
      I[qrow * k + j] = -1;
    }
  }
}

int group_rescore(const float* Q, const float* bank, int d, bool l2, int64_t nr, const int64_t* gsel, const float* cand_v,
                  const int32_t* cand_g, int ncg, int kg, int64_t nq, int k, float* D, int64_t* I, int64_t id_offset,
                  cudaStream_t stream, int gs) {
  if (nq == 0) return VSCB200_OK;
  int cpad = 2, gpad = 2;
  while (cpad < kg * gs) cpad <<= 1;
  while (cand_v != nullptr && gpad < ncg) gpad <<= 1;
  const size_t smem = static_cast<size_t>(cpad > gpad ? cpad : gpad) * sizeof(unsigned long long) +
                      static_cast<size_t>(d) * sizeof(float) + static_cast<size_t>(kg) * sizeof(int32_t);
  VSCB_REQUIRE(smem <= 200 * 1024, "group_rescore: too many groups / dimension too large");
  VSCB_CUDA_OK(cudaFuncSetAttribute(group_rescore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  ProfScope prof(kProfSelect, stream, static_cast<double>(nq) * kg * gs * d * 4);
  group_rescore_kernel<<<static_cast<unsigned>(nq), kGrThreads, smem, stream>>>(Q, bank, d, l2 ? 1 : 0, nr, gsel, cand_v, cand_g,
                                                                                ncg, gpad, kg, gs, cpad, k, D, I, id_offset);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

// ------------------------------------------------------------------ the same for a few query rows: one CTA per (query, group)
// With nq <= 128 the one-CTA-per-query kernel above leaves most SMs idle and serialises 16 groups' worth of HBM
// latency.  Here every CTA (j, q) repeats the cheap selection of the query's kg best groups (two levels of warp
// arg-max rounds over the composite keys in shared memory), rescores group j's rows, parks their keys in global
// scratch, and the last CTA of the query to arrive sorts the kg * gs keys and writes the k results.  Same keys,
// same tie rule (score, then lower id), same exact_rows_warp summation order: results are identical to the kernel above.
constexpr int kGfThreads = 256;

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
  const uint32_t hi = static_cast<uint32_t>(v >> 32);
  const uint32_t mh = __reduce_max_sync(0xffffffffu, hi);
  const uint32_t ml = __reduce_max_sync(0xffffffffu, hi == mh ? static_cast<uint32_t>(v) : 0u);
  return (static_cast<unsigned long long>(mh) << 32) | ml;
}

// the `rounds` largest keys of a[0:n) (destroyed: winners are zeroed) -> out[0:rounds), best first; warp-collective
__device__ __forceinline__ void warp_select_u64(unsigned long long* a, int n, int rounds, unsigned long long* out, int lane) {
  for (int r = 0; r < rounds; ++r) {
    unsigned long long lm = 0ull;
    int li = -1;
    for (int i = lane; i < n; i += 32) {
      const unsigned long long c = a[i];
      if (c > lm) { lm = c; li = i; }
    }
    const unsigned long long best = warp_max_u64(lm);
    if (best != 0ull && lm == best) a[li] = 0ull;          // keys are unique (the id is part of them)
    if (lane == 0) out[r] = best;
    __syncwarp();
  }
}

__global__ void __launch_bounds__(kGfThreads)
group_rescore_few_kernel(const float* __restrict__ Q, const float* __restrict__ bank, int d, int l2, int64_t nr,
                         const float* __restrict__ cand_v, const int32_t* __restrict__ cand_g, int ncg, int kg, int gs,
                         int cpad, int k, unsigned long long* __restrict__ keys, int* __restrict__ counters,
                         float* __restrict__ D, int64_t* __restrict__ I, int64_t id_offset) {
  extern __shared__ unsigned long long gf_smem[];        // [max(ncg, cpad)] composite keys | [8 * kg] warp winners | [d] floats
  pdl_launch_dependents();
  unsigned long long* cand = gf_smem;
  const int csz = ncg > cpad ? ncg : cpad;
  unsigned long long* wins = cand + csz;
  float* sq = reinterpret_cast<float*>(wins + (kGfThreads / 32) * kg);
  __shared__ int32_t my_group;
  __shared__ int is_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j = blockIdx.x;
  const int64_t qrow = blockIdx.y;
  const bool keep_max = !l2;
  for (int c = tid; c < d; c += kGfThreads) sq[c] = Q[qrow * d + c];     // an input of the call: safe before the wait
  pdl_wait();
  // ---- the kg best groups of this query over all chunks; this CTA needs the j-th
  for (int i = tid; i < ncg; i += kGfThreads) {
    const int32_t g = cand_g[qrow * ncg + i];
    cand[i] = g >= 0 ? (static_cast<unsigned long long>(okey_u(cand_v[qrow * ncg + i], true)) << 32) |
                           static_cast<uint32_t>(~static_cast<uint32_t>(g))
                     : 0ull;
  }
  __syncthreads();
  {
    const int per = (ncg + kGfThreads / 32 - 1) / (kGfThreads / 32);
    const int a0 = warp * per;
    const int n = a0 < ncg ? (ncg - a0 < per ? ncg - a0 : per) : 0;
    warp_select_u64(cand + a0, n, j + 1, wins + warp * kg, lane);       // only the first j + 1 of each slice can matter
  }
  __syncthreads();
  if (warp == 0) {
    // compact the warps' first j + 1 winners, then the j-th best of them
    const int m = j + 1;
    for (int i = lane; i < (kGfThreads / 32) * m; i += 32) cand[i] = wins[(i / m) * kg + (i % m)];
    __syncwarp();
    unsigned long long best = 0ull;
    for (int r = 0; r <= j; ++r) {
      unsigned long long lm = 0ull;
      int li = -1;
      for (int i = lane; i < (kGfThreads / 32) * m; i += 32) {
        const unsigned long long c = cand[i];
        if (c > lm) { lm = c; li = i; }
      }
      best = warp_max_u64(lm);
      if (best != 0ull && lm == best) cand[li] = 0ull;
      __syncwarp();
    }
    if (lane == 0) my_group = best != 0ull ? static_cast<int32_t>(~static_cast<uint32_t>(best & 0xFFFFFFFFull)) : -1;
  }
  __syncthreads();
  // ---- exact scores of the group's rows: four rows per warp, all loads of the CTA in flight at once
  const int64_t g = my_group;
  unsigned long long* out = keys + (qrow * kg + j) * gs;
  for (int r0 = warp * 4; r0 < gs; r0 += (kGfThreads / 32) * 4) {
    const float* rp[4];
    int64_t ids[4];
    bool ok[4];
    float acc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      ids[u] = g * gs + r0 + u;
      ok[u] = g >= 0 && r0 + u < gs && ids[u] < nr;
      rp[u] = bank + (ok[u] ? ids[u] : 0) * d;
    }
    exact_rows_warp<4>(sq, rp, d, lane, l2 != 0, acc);       // the shared summation order (exact.cuh)
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (lane == 0 && r0 + u < gs)
        out[r0 + u] = ok[u] ? (static_cast<unsigned long long>(okey_u(acc[u], keep_max)) << 32) |
                                  static_cast<uint32_t>(~static_cast<uint32_t>(ids[u]))
                            : 0ull;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const int prev = atomicAdd(&counters[qrow], 1);
    is_last = prev == kg - 1;
    if (is_last) counters[qrow] = 0;                          // every CTA of the query has arrived: ready for the next call
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // ---- last CTA of the query: sort the kg * gs keys, emit k
  const int ncand = kg * gs;
  for (int i = tid; i < cpad; i += kGfThreads) cand[i] = i < ncand ? __ldcg(keys + qrow * ncand + i) : 0ull;
  __syncthreads();
  for (int size = 2; size <= cpad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < (cpad >> 1); i += kGfThreads) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const unsigned long long a = cand[lo], b = cand[hi];
        if ((a < b) == desc) { cand[lo] = b; cand[hi] = a; }
      }
      __syncthreads();
    }
  }
  for (int jj = tid; jj < k; jj += kGfThreads) {
    const unsigned long long c = cand[jj];
    if (c != 0ull) {
      const uint32_t key = static_cast<uint32_t>(c >> 32);
      const uint32_t ok = keep_max ? key : ~key;
      const uint32_t u = (ok & 0x80000000u) ? (ok ^ 0x80000000u) : ~ok;
      D[qrow * k + jj] = __uint_as_float(u);
      I[qrow * k + jj] = id_offset + static_cast<int64_t>(~static_cast<uint32_t>(c & 0xFFFFFFFFull));
    } else {
      D[qrow * k + jj] = keep_max ? -FLT_MAX : FLT_MAX;
      I[qrow * k + jj] = -1;
    }
  }
}

size_t group_rescore_few_key_bytes(int64_t nq, int kg, int gs) { return static_cast<size_t>(nq) * kg * gs * sizeof(unsigned long long); }

// keys: group_rescore_few_key_bytes of scratch; counters: [nq] ints, zero before the first call (the kernel leaves them zero)
int group_rescore_few(const float* Q, const float* bank, int d, bool l2, int64_t nr, const float* cand_v, const int32_t* cand_g,
                      int ncg, int kg, int64_t nq, int k, float* D, int64_t* I, int64_t id_offset, void* keys, int* counters,
                      cudaStream_t stream, int gs) {
  if (nq == 0) return VSCB200_OK;
  VSCB_REQUIRE(cand_v && cand_g && keys && counters && nq <= 65535 && kg >= 1, "group_rescore_few: bad arguments");
  int cpad = 2;
  while (cpad < kg * gs) cpad <<= 1;
  const size_t smem = static_cast<size_t>(ncg > cpad ? ncg : cpad) * 8 + static_cast<size_t>(kGfThreads / 32) * kg * 8 +
                      static_cast<size_t>(d) * sizeof(float);
  VSCB_REQUIRE(smem <= 200 * 1024, "group_rescore_few: too many groups / dimension too large");
  VSCB_CUDA_OK(cudaFuncSetAttribute(group_rescore_few_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  ProfScope prof(kProfSelect, stream, static_cast<double>(nq) * kg * gs * d * 4);
  VSCB_CUDA_OK(launch_pdl_chain(group_rescore_few_kernel, dim3(kg, static_cast<unsigned>(nq)), dim3(kGfThreads), smem, stream, Q, bank, d,
                          l2 ? 1 : 0, nr, cand_v, cand_g, ncg, kg, gs, cpad, k, reinterpret_cast<unsigned long long*>(keys),
                          counters, D, I, id_offset));
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

}  // namespace vscb200
