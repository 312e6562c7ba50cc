// Tensor-core similarity scoring, fp32-equivalent: S[nq, nr] = Q . R^T on tcgen05 with a 2-way bf16
// split of both operands (x = hi + lo, hi = bf16(x), lo = bf16(x - hi)) and three MMAs per K step:
//     S = Qh.Rh + Ql.Rh + Qh.Rl        (the dropped lo.lo term is <= 2^-18 |q.r| per product)
// accumulated in fp32 in TMEM.  The K loop is cut into kSeg independent accumulator segments that are
// summed in the epilogue, which shortens every tensor-core accumulation chain (fewer roundings of the
// running sum).  Operand planes are produced once per bank row at add() time and once per query block;
// hi+lo is 4 bytes per element, i.e. the HBM traffic of the fp32 bank.
//
// Same pipeline skeleton as gemm.cu: TMA producer warp / single-thread MMA issuer / TMEM allocator /
// 8 epilogue warps, 3-stage smem ring of {Qh, Ql, Rh, Rl} 128x64 tiles (64 KB per stage), persistent
// over 128x128 output tiles.  Epilogue: segment sum -> optional squared-L2 transform -> coalesced fp32
// stores of the dense score block (consumed by select.cu: top-k, range search, dense sim matrices).
//
// Round 2: the same skeleton also runs ONE pass (Qh.Rh only; `passes = 1`, six-stage ring of {Qh, Rh} tiles, one accumulator
// segment) where an approximate score with a proven margin is enough (margin.cuh): the column-sample GEMM of the candidate
// search / the top-k threshold bootstrap, and mode 2, whose epilogue emits the pairs better than a per-row threshold
// instead of storing anything dense (global_topk.cu).  The MMA-issuing warp runs on warp-uniform values.
//
// Reference: faiss IndexFlat scoring (cuBLAS SGEMM on GPU / sgemm on CPU) behind vsc/index.py:174,
// vsc/exhaustive_search.py:62,74, vsc/baseline/score_normalization.py:95.
#include "host_util.h"
#include "kernels.h"
#include "ptx.cuh"

namespace vscb200 {

constexpr int kTM = 128, kTN = 128, kTK = 64;
constexpr int kTStages = 3;
constexpr int kTSeg = 2;
constexpr int kTEpiWarps = 8;
constexpr int kTThreads = 128 + 32 * kTEpiWarps;
constexpr int kTTile = kTM * kTK * 2;                 // one 128x64 bf16 tile: 16 KB
constexpr int kTStageBytes = 4 * kTTile;              // Qh, Ql, Rh, Rl
constexpr int kTTmemCols = 2 * kTSeg * kTN;           // 2 buffers x kSeg segments x 128 columns = 512
constexpr int kTSmem = kTStages * kTStageBytes + 256 + kTEpiWarps * 4096 + 1024;

struct SimParams {
  float* S;
  int64_t ldS;
  int64_t nq, nr;
  int K;            // padded feature dim (multiple of 8)
  int l2;
  int vec4;         // S rows are 16-byte aligned (ldS % 4 == 0): 128-bit stores
  const float* qn;
  const float* rn;
  int tiles_m, tiles_n;
  // fused top-k mode (kFused): work items = (query block, bank slab); per-thread running top-kFK lists are
  // flushed per item to cand_d / cand_i [nq, slabs*2*kFK]
  int slabs;
  float* cand_d;
  int32_t* cand_i;
  // passes = 3: the split above; passes = 1: Qh.Rh only (bf16-rounded operands, error bounded by margin.cuh) -- the lo
  // planes are neither loaded nor multiplied
  int passes;
  // emit mode (kMode == 2, global_topk.cu): nothing dense is written; every score better than its row's threshold
  // radius -/+ marg[row] is appended as (value, pair id = (q0 + row) * ntotal + column) to bufv / bufp through a
  // per-CTA shared-memory buffer flushed once per tile (one atomicAdd on `counter` per tile instead of one per hit).
  // `counter` keeps counting past `cap` so the host sees by how much a pass overflowed.
  const float* marg;
  float radius;
  int has_radius;
  int64_t q0, ntotal;
  float* bufv;
  uint64_t* bufp;
  unsigned long long* counter;
  unsigned long long cap;
};

constexpr int kEmitCap = 2048;    // entries of the per-CTA emit buffer (8 KB values + 16 KB pair ids of the 32 KB epilogue staging)

constexpr int kFK = 16;     // per-thread list length of the fused top-k epilogue (k + rescoring slack <= kFK)
constexpr int kFG = 8;      // rows per candidate group of the fused epilogue

// Work decomposition shared by the three warp roles.  Dense mode: item = one 128x128 tile, n fastest.
// Fused mode: item = (m_blk, slab) and the tiles of the slab are walked consecutively, so a thread's
// running top-k list lives in registers for the whole item.
struct ItemIter {
  int item, step, num_items, slabs, tiles_n;
  __device__ ItemIter(const SimParams& p, bool fused)
      : item(blockIdx.x), step(gridDim.x), num_items(fused ? p.tiles_m * p.slabs : p.tiles_m * p.tiles_n),
        slabs(fused ? p.slabs : p.tiles_n), tiles_n(p.tiles_n) {}
  __device__ bool valid() const { return item < num_items; }
  __device__ void advance() { item += step; }
  __device__ int m_blk() const { return item / slabs; }
  __device__ int slab() const { return item % slabs; }
  // tile range [n0, n1) of this item
  __device__ int n0(bool fused) const { return fused ? static_cast<int>(static_cast<int64_t>(slab()) * tiles_n / slabs) : slab(); }
  __device__ int n1(bool fused) const { return fused ? static_cast<int>(static_cast<int64_t>(slab() + 1) * tiles_n / slabs) : slab() + 1; }
};

template <int kMode>      // 0: dense score block, 1: fused per-row top-k lists, 2: threshold emission
__global__ void __launch_bounds__(kTThreads, 1)
sim3_kernel(const __grid_constant__ CUtensorMap tmQh, const __grid_constant__ CUtensorMap tmQl,
            const __grid_constant__ CUtensorMap tmRh, const __grid_constant__ CUtensorMap tmRl, SimParams p) {
  constexpr bool kFused = kMode == 1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  // the ring holds 3 stages of {Qh, Ql, Rh, Rl} or, with one pass, 6 stages of {Qh, Rh}: the mainloop is bound by TMA
  // round trips per k-block, not by bytes, so the one-pass form needs the deeper ring to go faster at all
  const int nstages = p.passes == 3 ? kTStages : 2 * kTStages;
  const int stage_bytes = p.passes == 3 ? kTStageBytes : kTStageBytes / 2;
  const int r_off = p.passes == 3 ? 2 * kTTile : kTTile;      // Rh inside a stage
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kTStages * kTStageBytes);
  uint64_t* empty_bar = full_bar + 2 * kTStages;
  uint64_t* tfull_bar = empty_bar + 2 * kTStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float4* stage_all = reinterpret_cast<float4*>(smem + kTStages * kTStageBytes + 256);
  // emit mode reuses the staging area: [kEmitCap] values | [kEmitCap] pair ids | fill count | reserved base
  float* emit_val = reinterpret_cast<float*>(stage_all);
  uint64_t* emit_pid = reinterpret_cast<uint64_t*>(emit_val + kEmitCap);
  unsigned long long* emit_base = reinterpret_cast<unsigned long long*>(emit_pid + kEmitCap);
  uint32_t* emit_cnt = reinterpret_cast<uint32_t*>(emit_base + 1);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // uniform for the compiler
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQh); prefetch_tmap(&tmQl); prefetch_tmap(&tmRh); prefetch_tmap(&tmRl);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < nstages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], kTEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<kTTmemCols>(tmem_slot);
  if (kMode == 2 && threadIdx.x == 0) *emit_cnt = 0u;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  const int kblocks = (p.K + kTK - 1) / kTK;
  // accumulator segments shorten the rounding chains of the fp32-equivalent form; the one-pass form (an approximate score
  // with a proven margin) accumulates in one
  const int nseg = p.passes != 3 ? 1 : (kblocks < kTSeg ? kblocks : kTSeg);

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (ItemIter it(p, kFused); it.valid(); it.advance())
      for (int n_blk = it.n0(kFused), n_end = it.n1(kFused), m_blk = it.m_blk(); n_blk < n_end; ++n_blk) {
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], stage_bytes);
          uint8_t* st = smem + stage * stage_bytes;
          tma_load_2d(st, &tmQh, &full_bar[stage], kb * kTK, m_blk * kTM, kEvictLast);
          tma_load_2d(st + r_off, &tmRh, &full_bar[stage], kb * kTK, n_blk * kTN, kEvictNormal);
          if (p.passes == 3) {
            tma_load_2d(st + kTTile, &tmQl, &full_bar[stage], kb * kTK, m_blk * kTM, kEvictLast);
            tma_load_2d(st + 3 * kTTile, &tmRl, &full_bar[stage], kb * kTK, n_blk * kTN, kEvictNormal);
          }
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // MMA issuer: the WHOLE warp runs the loop on warp-uniform values (uniform registers), one elected lane issues inside
    // the umma_*_warp wrappers -- no per-MMA election loop, no register -> uniform-register moves
    {
      constexpr uint32_t idesc = make_idesc_bf16_f32(kTM, kTN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (ItemIter it(p, kFused); it.valid(); it.advance())
      for (int n_blk = it.n0(kFused), n_end = it.n1(kFused); n_blk < n_end; ++n_blk) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        int seg = 0, seg_end = kblocks / nseg;        // segment s covers [s*kblocks/nseg, (s+1)*kblocks/nseg)
        bool fresh = true;
        for (int kb = 0; kb < kblocks; ++kb) {
          if (kb == seg_end) {
            ++seg;
            seg_end = (seg + 1) * kblocks / nseg;
            fresh = true;
          }
          const uint32_t d_tmem = tmem_base + (acc * kTSeg + seg) * kTN;
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t st = smem_u32(smem + stage * stage_bytes);
          const uint64_t qh = make_desc_k_sw128(st), rh = make_desc_k_sw128(st + r_off);
          if (p.passes == 3) {
            const uint64_t ql = make_desc_k_sw128(st + kTTile), rl = make_desc_k_sw128(st + 3 * kTTile);
#pragma unroll
            for (int k = 0; k < kTK / 16; ++k) {
              umma_bf16_ss_warp(d_tmem, qh + 2 * k, rh + 2 * k, idesc, (fresh && k == 0) ? 0u : 1u);
              umma_bf16_ss_warp(d_tmem, ql + 2 * k, rh + 2 * k, idesc, 1u);
              umma_bf16_ss_warp(d_tmem, qh + 2 * k, rl + 2 * k, idesc, 1u);
            }
          } else {
#pragma unroll
            for (int k = 0; k < kTK / 16; ++k)
              umma_bf16_ss_warp(d_tmem, qh + 2 * k, rh + 2 * k, idesc, (fresh && k == 0) ? 0u : 1u);
          }
          fresh = false;
          umma_commit_warp(&empty_bar[stage]);
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
        umma_commit_warp(&tfull_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int ew = warp - 4, quad = warp & 3, half = ew >> 2;
    float4* stage = stage_all + ew * 256;
    // emit mode: all epilogue threads have passed a barrier since the last append and agree that a flush is due
    auto emit_flush = [&]() {
      const int et = threadIdx.x - 128;
      const uint32_t filled = *emit_cnt;
      const uint32_t nflush = filled < static_cast<uint32_t>(kEmitCap) ? filled : static_cast<uint32_t>(kEmitCap);
      if (et == 0) *emit_base = nflush ? atomicAdd(p.counter, static_cast<unsigned long long>(nflush)) : 0ull;
      named_bar_sync(2, 32 * kTEpiWarps);
      const unsigned long long base = *emit_base;
      for (uint32_t i = et; i < nflush; i += 32 * kTEpiWarps)
        if (base + i < p.cap) { p.bufv[base + i] = emit_val[i]; p.bufp[base + i] = emit_pid[i]; }
      if (et == 0) *emit_cnt = 0u;
      named_bar_sync(2, 32 * kTEpiWarps);
    };
    int acc = 0;
    uint32_t acc_phase = 0;
    for (ItemIter it(p, kFused); it.valid(); it.advance()) {
      const int m_blk = it.m_blk();
      const int64_t row_base = static_cast<int64_t>(m_blk) * kTM + quad * 32;
      // fused mode: this thread's running top-kFK of (query row row_base+lane) over its columns of the slab,
      // sorted best-first; strict comparisons keep the earlier (lower) id among equal scores
      float ls[kFK];
      int32_t li[kFK];
      float qn_row = 0.f;
      float thr_row = 0.f;
      if (kMode == 2) {
        const int64_t grow = row_base + lane;
        const bool live = grow < p.nq;
        if (p.l2 && live) qn_row = p.qn[grow];
        const float m = live ? p.marg[grow] : 0.f;
        thr_row = !p.has_radius ? (p.l2 ? INFINITY : -INFINITY) : (p.l2 ? p.radius + m : p.radius - m);
        if (!live) thr_row = p.l2 ? -INFINITY : INFINITY;      // rows past the block never emit
      }
      if (kFused) {
#pragma unroll
        for (int j = 0; j < kFK; ++j) { ls[j] = -INFINITY; li[j] = -1; }
        if (p.l2 && row_base + lane < p.nq) qn_row = p.qn[row_base + lane];
      }
      for (int n_blk = it.n0(kFused), n_end = it.n1(kFused); n_blk < n_end; ++n_blk) {
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < kTN / 2 / 32; ++c) {
        const int col0 = half * (kTN / 2) + c * 32;
        const int64_t gcol = static_cast<int64_t>(n_blk) * kTN + col0;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kTSeg * kTN + col0;
        uint32_t v[32];
        float f[32];
        __syncwarp();
        tmem_ld_32x32(taddr, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (nseg > 1) {
          tmem_ld_32x32(taddr + kTN, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] += __uint_as_float(v[j]);
        }
        if (gcol >= p.nr) continue;   // warp-uniform
        if (kMode == 2) {
          uint32_t hits = 0;
          if (!p.l2 && gcol + 32 <= p.nr) {                // (warp-uniform) interior chunk, inner product: compare only
#pragma unroll
            for (int j = 0; j < 32; ++j) hits |= (f[j] > thr_row ? 1u : 0u) << j;
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float v = f[j];
              const bool in = gcol + j < p.nr;
              if (p.l2) v = fmaxf(qn_row + (in ? p.rn[gcol + j] : 0.f) - 2.0f * v, 0.f);
              f[j] = v;
              if (in && (p.l2 ? v < thr_row : v > thr_row)) hits |= 1u << j;
            }
          }
          // survivors are rare (a fraction of a percent): ONE copy of the append code, run per set bit, with the value
          // picked out of the register array by a 31-select tree (an unrolled copy per column thrashed the
          // instruction cache: 1.66 ms -> see profiles/README.md)
          const uint64_t pid0 = static_cast<uint64_t>(p.q0 + row_base + lane) * static_cast<uint64_t>(p.ntotal) + static_cast<uint64_t>(gcol);
#pragma unroll 1
          while (hits) {
            const int j = __ffs(hits) - 1;
            hits &= hits - 1u;
            float t16[16], t8[8], t4[4];
#pragma unroll
            for (int i = 0; i < 16; ++i) t16[i] = (j & 1) ? f[2 * i + 1] : f[2 * i];
#pragma unroll
            for (int i = 0; i < 8; ++i) t8[i] = (j & 2) ? t16[2 * i + 1] : t16[2 * i];
#pragma unroll
            for (int i = 0; i < 4; ++i) t4[i] = (j & 4) ? t8[2 * i + 1] : t8[2 * i];
            const float u0 = (j & 8) ? t4[1] : t4[0], u1 = (j & 8) ? t4[3] : t4[2];
            const float v = (j & 16) ? u1 : u0;
            const uint32_t slot = atomicAdd(emit_cnt, 1u);
            if (slot < static_cast<uint32_t>(kEmitCap)) {
              emit_val[slot] = v;
              emit_pid[slot] = pid0 + j;
            } else {                                     // buffer full before the tile ended: straight to the global list
              const unsigned long long pos = atomicAdd(p.counter, 1ull);
              if (pos < p.cap) { p.bufv[pos] = v; p.bufp[pos] = pid0 + j; }
            }
          }
          continue;
        }
        if (kFused) {
          // Candidates are GROUPS of kFG = 8 consecutive bank rows, keyed by the group's best selection key
          // (larger is better; negated squared distance for L2): 8x fewer list updates than per-pair
          // candidates, and the k best rows are always inside the k groups with the best maxima, which
          // group_rescore_kernel rescans exactly.
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            float gm = -INFINITY;
#pragma unroll
            for (int j = 8 * g8; j < 8 * g8 + 8; ++j) {
              float key = f[j];
              if (p.l2) key = 2.0f * key - p.rn[min(gcol + j, p.nr - 1)];
              if (gcol + j < p.nr) gm = fmaxf(gm, key);
            }
            if (p.l2) gm = -fmaxf(qn_row - gm, 0.f);
            if (gm > ls[kFK - 1]) {
              ls[kFK - 1] = gm;
              li[kFK - 1] = static_cast<int32_t>((gcol >> 3) + g8);
#pragma unroll
              for (int t = kFK - 1; t > 0; --t) {
                if (ls[t] > ls[t - 1]) {
                  const float ts = ls[t]; ls[t] = ls[t - 1]; ls[t - 1] = ts;
                  const int32_t ti = li[t]; li[t] = li[t - 1]; li[t - 1] = ti;
                }
              }
            }
          }
          continue;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
          stage[lane * 8 + (q ^ (lane & 7))] = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
        __syncwarp();
        const int q = lane & 7;
        const int64_t gc = gcol + q * 4;
        float4 rn4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.l2) {
          rn4.x = gc < p.nr ? p.rn[gc] : 0.f;
          rn4.y = gc + 1 < p.nr ? p.rn[gc + 1] : 0.f;
          rn4.z = gc + 2 < p.nr ? p.rn[gc + 2] : 0.f;
          rn4.w = gc + 3 < p.nr ? p.rn[gc + 3] : 0.f;
        }
#pragma unroll
        for (int it8 = 0; it8 < 8; ++it8) {
          const int r = it8 * 4 + (lane >> 3);
          const int64_t grow = row_base + r;
          float4 o = stage[r * 8 + (q ^ (r & 7))];
          if (grow < p.nq && gc < p.nr) {
            if (p.l2) {
              const float qn = p.qn[grow];
              o.x = fmaxf(qn + rn4.x - 2.0f * o.x, 0.f);
              o.y = fmaxf(qn + rn4.y - 2.0f * o.y, 0.f);
              o.z = fmaxf(qn + rn4.z - 2.0f * o.z, 0.f);
              o.w = fmaxf(qn + rn4.w - 2.0f * o.w, 0.f);
            }
            float* dst = p.S + grow * p.ldS + gc;
            if (p.vec4 && gc + 3 < p.nr) {
              *reinterpret_cast<float4*>(dst) = o;          // ldS % 4 == 0 and gc % 4 == 0: 16-byte aligned
            } else {
              dst[0] = o.x;
              if (gc + 1 < p.nr) dst[1] = o.y;
              if (gc + 2 < p.nr) dst[2] = o.z;
              if (gc + 3 < p.nr) dst[3] = o.w;
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      if (kMode == 2) {
        // flush the survivors once the buffer is half full (a tile adds a few dozen): one barrier per tile, whose OR
        // reduction makes the decision the same for every thread (a fast warp may already be appending the next
        // tile's survivors when a slow one looks); three more and one global reservation per flush
        if (named_bar_or(2, 32 * kTEpiWarps, *emit_cnt > static_cast<uint32_t>(kEmitCap / 2))) emit_flush();
      }
      }   // tiles of the item
      if (kFused) {
        const int64_t grow = row_base + lane;
        if (grow < p.nq) {
          const int64_t base = (grow * p.slabs + it.slab()) * (2 * kFK) + half * kFK;
#pragma unroll
          for (int j = 0; j < kFK; ++j) {
            p.cand_d[base + j] = ls[j];
            p.cand_i[base + j] = li[j];
          }
        }
      }
    }
    if (kMode == 2) {                                      // what the last tiles left in the buffer
      named_bar_sync(2, 32 * kTEpiWarps);
      emit_flush();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<kTTmemCols>(tmem_base);
}

// ------------------------------------------------------------------ operand planes
// x [n, d] fp32 -> hi, lo [n, dp] bf16 (dp = d rounded up to 8, zero padded)
__global__ void split_planes_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                                    __nv_bfloat16* __restrict__ lo, int64_t n, int d, int dp) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n * dp) return;
  const int64_t r = i / dp;
  const int c = static_cast<int>(i % dp);
  const float v = c < d ? x[r * d + c] : 0.f;
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi[i] = h;
  lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}

int split_planes(const float* x, void* hi, void* lo, int64_t n, int d, int dp, cudaStream_t stream) {
  const int64_t total = n * dp;
  if (total == 0) return VSCB200_OK;
  split_planes_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      x, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), n, d, dp);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

static int launch_sim3(const void* Qh, const void* Ql, const void* Rh, const void* Rl, SimParams p, int dp, int mode,
                       cudaStream_t stream, int64_t r_stride = 1) {
  CUtensorMap tQh, tQl, tRh, tRl;
  int rc;
  if (p.passes != 3) { p.passes = 1; Ql = Qh; Rl = Rh; }      // the lo maps are never dereferenced
  // r_stride > 1: bank row j of this launch is row j * r_stride of the planes (a column sample of the score block)
  if ((rc = make_tmap_2d(&tQh, Qh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p.nq, dp, dp, kTM, kTK, true))) return rc;
  if ((rc = make_tmap_2d(&tQl, Ql, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p.nq, dp, dp, kTM, kTK, true))) return rc;
  if ((rc = make_tmap_2d(&tRh, Rh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p.nr, dp, dp * r_stride, kTN, kTK, true))) return rc;
  if ((rc = make_tmap_2d(&tRl, Rl, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p.nr, dp, dp * r_stride, kTN, kTK, true))) return rc;
  const bool fused = mode == 1;
  p.tiles_m = static_cast<int>((p.nq + kTM - 1) / kTM);
  const int64_t tn = (p.nr + kTN - 1) / kTN;
  VSCB_REQUIRE(static_cast<int64_t>(p.tiles_m) * tn < (1ll << 31), "scores_tc: too many tiles");
  p.tiles_n = static_cast<int>(tn);
  const int64_t items = fused ? static_cast<int64_t>(p.tiles_m) * p.slabs : static_cast<int64_t>(p.tiles_m) * p.tiles_n;
  const int grid = static_cast<int>(items < device_sm_count() ? items : device_sm_count());
  ProfScope prof(kProfScores, stream, 2.0 * static_cast<double>(p.nq) * p.nr * dp);
  if (mode == 1) {
    VSCB_CUDA_OK(cudaFuncSetAttribute(sim3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTSmem));
    sim3_kernel<1><<<grid, kTThreads, kTSmem, stream>>>(tQh, tQl, tRh, tRl, p);
  } else if (mode == 2) {
    VSCB_CUDA_OK(cudaFuncSetAttribute(sim3_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTSmem));
    sim3_kernel<2><<<grid, kTThreads, kTSmem, stream>>>(tQh, tQl, tRh, tRl, p);
  } else {
    VSCB_CUDA_OK(cudaFuncSetAttribute(sim3_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTSmem));
    sim3_kernel<0><<<grid, kTThreads, kTSmem, stream>>>(tQh, tQl, tRh, tRl, p);
  }
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

int scores_tc_planes(const void* Qh, const void* Ql, const void* Rh, const void* Rl, float* S, int64_t nq, int64_t nr,
                     int dp, int64_t ldS, bool l2, const float* qn, const float* rn, cudaStream_t stream, int passes,
                     int64_t r_stride) {
  if (nq == 0 || nr == 0) return VSCB200_OK;
  VSCB_REQUIRE(dp % 8 == 0, "scores_tc: dp must be a multiple of 8");
  SimParams p = {};
  p.S = S; p.ldS = ldS; p.nq = nq; p.nr = nr; p.K = dp; p.l2 = l2 ? 1 : 0;
  p.vec4 = (ldS % 4 == 0 && (reinterpret_cast<uintptr_t>(S) & 15) == 0) ? 1 : 0;
  p.qn = qn; p.rn = rn;
  p.passes = passes;
  VSCB_REQUIRE(r_stride == 1 || !l2, "scores_tc: a strided bank sample is inner-product only (rn is not strided)");
  return launch_sim3(Qh, Ql, Rh, Rl, p, dp, 0, stream, r_stride);
}

// Threshold emission for the global candidate search (global_topk.cu): one bf16 pass over the [nq, nr] block, survivors
// (value better than radius -/+ marg[row]; everything when !has_radius) appended to bufv / bufp, `counter` advanced.
// qn: squared norms of the block's rows (L2), marg: their margins, q0: global row of block row 0.
int scores_tc_emit(const void* Qh, const void* Rh, int64_t nq, int64_t nr, int dp, bool l2, const float* qn, const float* rn,
                   const float* marg, float radius, bool has_radius, int64_t q0, float* bufv, uint64_t* bufp,
                   unsigned long long* counter, unsigned long long cap, cudaStream_t stream) {
  if (nq == 0 || nr == 0) return VSCB200_OK;
  VSCB_REQUIRE(dp % 8 == 0, "scores_tc_emit: dp must be a multiple of 8");
  SimParams p = {};
  p.nq = nq; p.nr = nr; p.K = dp; p.l2 = l2 ? 1 : 0; p.qn = qn; p.rn = rn;
  p.passes = 1;
  p.marg = marg; p.radius = radius; p.has_radius = has_radius ? 1 : 0; p.q0 = q0; p.ntotal = nr;
  p.bufv = bufv; p.bufp = bufp; p.counter = counter; p.cap = cap;
  return launch_sim3(Qh, nullptr, Rh, nullptr, p, dp, 2, stream);
}

// Number of bank slabs per query block for the fused top-k mode: enough work items to fill the GPU
// (>= 4 per SM when the bank allows), at most 148 (candidate list <= 148*2*kFK per query).
int fused_topk_slabs(int64_t nq, int64_t nr) {
  const int64_t tiles_m = (nq + kTM - 1) / kTM, tiles_n = (nr + kTN - 1) / kTN;
  int64_t s = (4ll * device_sm_count() + tiles_m - 1) / tiles_m;
  if (s > tiles_n) s = tiles_n;
  if (s > 148) s = 148;
  if (s < 1) s = 1;
  return static_cast<int>(s);
}
int fused_topk_list_len() { return kFK; }
int fused_topk_group_rows() { return kFG; }

// cand_d / cand_i: [nq, slabs * 2 * kFK] group keys (larger = better; -distance for L2) and group ids = row / kFG
// (-1 = empty)
int topk_tc_fused(const void* Qh, const void* Ql, const void* Rh, const void* Rl, int64_t nq, int64_t nr, int dp, bool l2,
                  const float* qn, const float* rn, int slabs, float* cand_d, int32_t* cand_i, cudaStream_t stream) {
  if (nq == 0 || nr == 0) return VSCB200_OK;
  VSCB_REQUIRE(dp % 8 == 0 && nr < (1ll << 31), "topk_tc_fused: dp must be a multiple of 8 and nr < 2^31");
  SimParams p = {};
  p.nq = nq; p.nr = nr; p.K = dp; p.l2 = l2 ? 1 : 0; p.qn = qn; p.rn = rn;
  p.slabs = slabs; p.cand_d = cand_d; p.cand_i = cand_i;
  p.passes = 3;
  return launch_sim3(Qh, Ql, Rh, Rl, p, dp, 1, stream);
}

}  // namespace vscb200
