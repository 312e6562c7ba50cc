// Persistent warp-specialised bf16 GEMM for sm_100a:  C[M,N] = A[M,K] * W[N,K]^T (+bias, act, residual)
//
//   warp 0      TMA producer   (cp.async.bulk.tensor, 128B-swizzled 128x64 / BNx64 bf16 tiles, 4-stage ring)
//   warp 1      MMA issuer     (tcgen05.mma kind::f16, M=128, N=BN, K=16; fp32 accumulators in TMEM)
//   warp 2      TMEM allocator (2 accumulator stages x BN columns)
//   warps 4-11  epilogue       (tcgen05.ld -> bias / QuickGELU / GELU in registers -> 128B-swizzled 32-row box in
//                               smem -> TMA store (bf16 / fp32) or TMA fp32 reduce-add into the residual stream)
//
// Three mbarrier pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue), and a
// static persistent tile schedule (tile = blockIdx.x + i*gridDim.x, N-blocks fastest so that
// concurrently resident CTAs share A tiles through L2 while W stays L2 resident).
//
// kSplit (the encoders' fp32-equivalent mode): both operands arrive as two bf16 planes, x = hi + lo with hi = bf16(x),
// lo = bf16(x - hi), and every K step issues three MMAs  hi.hi + lo.hi + hi.lo  into the same fp32 accumulator (the
// dropped lo.lo term is <= 2^-18 |a.w| per product) -- 16 mantissa bits per operand instead of 8.  bf16 outputs leave
// as two planes as well (a second tensor map).  128 x 128 tiles, one CTA per tile, 3-stage ring of {Ah, Al, Wh, Wl}.
//
// This is the dense-projection engine of the ViT encoder (reference: every nn.Linear /
// nn.MultiheadAttention projection of D/train/train_vid_score/video/clip.py:33-39,45-49 and the
// conv patch-embed :105,143 as an im2row GEMM), replacing the cuBLAS calls of torch 1.11.
#include <stdlib.h>

#include "host_util.h"
#include "ptx.cuh"

namespace vscb200 {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kEpiWarps = 8;
constexpr int kGemmThreads = 128 + 32 * kEpiWarps;

struct GemmParams {
  const float* bias;
  void* C;
  int64_t M;
  int N, K;
  int64_t ldc;
  int epilogue;
  int act;
  int tiles_m, tiles_n;
  // patch-embed epilogue (VSCB_EPI_PATCH_F32): GEMM row m = frame*P + patch  ->  token row m + m/P + 1
  // (slot 0 of every frame is the class token), plus the positional embedding pos[(m % P) + 1, :].
  const float* pos;
  int patch_P;
  int reverse;      // walk the tile list from the last M block to the first (kernels.h)
  // Swin-V2 cosine attention (swinv2.py:160-163), fused into the QKV projection: every 32-column chunk with
  // column < qk_norm_cols is one head's q or k vector of this row -> L2-normalised (F.normalize, eps 1e-12);
  // q chunks (column < qk_norm_cols / 2) are also multiplied by their head's clamped exp(logit_scale).
  int qk_norm_cols;
  const float* qscale;
};

constexpr int VSCB_EPI_PATCH_F32 = 3;

template <int BN, int kCluster, bool kSplit>
struct GemmCfg {
  static constexpr int kStages = kSplit ? 3 : ((kCluster == 2) ? 5 : ((BN == 256) ? 3 : 5));
  static constexpr int kATile = kBM * kBK * 2;
  static constexpr int kBTile = (BN / kCluster) * kBK * 2;        // pair mode: each CTA stages half of the W tile
  static constexpr int kABytes = kATile * (kSplit ? 2 : 1);       // split mode: hi plane tile, then lo plane tile
  static constexpr int kBBytes = kBTile * (kSplit ? 2 : 1);
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kEpiBoxes = kSplit ? 1 : 2;                // 4 KB epilogue boxes per warp (double-buffered unless split)
  static constexpr int kEpiWarpBytes = 4096 * kEpiBoxes;
  static constexpr int kStageBytes = kEpiWarps * kEpiWarpBytes;
  static constexpr int kSmemBytes =
      kStages * (kABytes + kBBytes) + kStageBytes + 256 /*barriers*/ + 1024 /*align slack*/;
};

__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <bool kPrecise>
__device__ __forceinline__ float apply_act(float x, int act) {
  // QuickGELU x*sigmoid(1.702x) = 0.5x*tanh(0.851x) + 0.5x : ONE MUFU op per element (ex2 + rcp would be
  // two, and the epilogue of the K=768 fc1 GEMM is MUFU-paced).  tanh.approx error 2^-11 << bf16 output ulp.
  // kPrecise (fp32-equivalent mode): the reference's formula with expf.
  if (act == VSCB200_ACT_QUICK_GELU) {
    if (kPrecise) return x / (1.0f + expf(-1.702f * x));
    const float h = 0.5f * x;
    return fmaf(h, tanh_approx(0.851f * x), h);
  }
  if (act == VSCB200_ACT_GELU) return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
  return x;
}

// kCluster == 2: a CTA pair (two SMs of one TPC) computes one 256 x BN tile with tcgen05.mma.cta_group::2
// (M = 256): each CTA stages its own 128 rows of A and HALF of the W tile (BN/2 rows) -- 32 KB instead of
// 48 KB of shared-memory fill and operand reads per 64-wide K block, which is what bounds the single-CTA
// kernel (A 16 KB + W 32 KB read per 512 tensor cycles on a 128 B/clk port that also takes the TMA
// writes and the epilogue staging).  Rank 0 issues every MMA; both producers' TMA loads complete on
// rank 0's full barrier; stages and accumulators are released to both CTAs by multicast tcgen05.commit;
// each CTA drains its own 128 accumulator lanes and the peer's epilogue warps arrive remotely on rank
// 0's TMEM-empty barrier.
template <int BN, int kCluster, bool kSplit>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmA2,
                 const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmC2, GemmParams p) {
  static_assert(!kSplit || kCluster == 1, "the split-bf16 mode runs one CTA per tile");
  using Cfg = GemmCfg<BN, kCluster, kSplit>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStages * Cfg::kABytes;
  uint8_t* stage_all = sB + kStages * Cfg::kBBytes;        // 1024-byte aligned (every operand stage is a multiple of 8 KB)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stage_all + Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);     // uniform for the compiler
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmC);
    if (kSplit) { prefetch_tmap(&tmA2); prefetch_tmap(&tmB2); prefetch_tmap(&tmC2); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], kEpiWarps * kCluster);   // rank 0's barrier collects both CTAs' epilogue warps
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (kCluster > 1) tmem_alloc_cg2<Cfg::kTmemCols>(tmem_slot);
    else tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();    // peer barriers are initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_wait();        // everything above overlapped the previous kernel's tail; its results are visible from here

  // Tile schedule.  kCluster == 1: tile = blockIdx.x + i*gridDim.x.  kCluster == 2: the pair walks
  // "pair tiles" (two consecutive M blocks x one N block); rank r takes M block 2*pm + r.
  const uint32_t crank = kCluster > 1 ? cluster_ctarank() : 0u;
  const int sched_m = kCluster > 1 ? (p.tiles_m + 1) / 2 : p.tiles_m;
  const int num_tiles = sched_m * p.tiles_n;
  const int sched_first = kCluster > 1 ? static_cast<int>(blockIdx.x / kCluster) : static_cast<int>(blockIdx.x);
  const int sched_step = kCluster > 1 ? static_cast<int>(gridDim.x / kCluster) : static_cast<int>(gridDim.x);
  const int kblocks = (p.K + kBK - 1) / kBK;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = sched_first; tile < num_tiles; tile += sched_step) {
        const int vt = p.reverse ? num_tiles - 1 - tile : tile;
        const int m_blk = (vt / p.tiles_n) * kCluster + static_cast<int>(crank), n_blk = vt % p.tiles_n;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (kCluster > 1) {
            // both CTAs' boxes complete on rank 0's barrier, which expects the bytes of the whole pair
            const uint32_t lead_full = mapa_u32(smem_u32(&full_bar[stage]), 0u);
            if (crank == 0) mbar_expect_tx(&full_bar[stage], 2 * (Cfg::kABytes + Cfg::kBBytes));
            tma_load_2d_cg2(sA + stage * Cfg::kABytes, &tmA, lead_full, kb * kBK, m_blk * kBM, kEvictNormal);
            tma_load_2d_cg2(sB + stage * Cfg::kBBytes, &tmB, lead_full, kb * kBK,
                            n_blk * BN + static_cast<int>(crank) * (BN / 2), kEvictLast);
          } else {
            mbar_expect_tx(&full_bar[stage], Cfg::kABytes + Cfg::kBBytes);
            tma_load_2d(sA + stage * Cfg::kABytes, &tmA, &full_bar[stage], kb * kBK, m_blk * kBM, kEvictNormal);
            tma_load_2d(sB + stage * Cfg::kBBytes, &tmB, &full_bar[stage], kb * kBK, n_blk * BN, kEvictLast);
            if (kSplit) {
              tma_load_2d(sA + stage * Cfg::kABytes + Cfg::kATile, &tmA2, &full_bar[stage], kb * kBK, m_blk * kBM, kEvictNormal);
              tma_load_2d(sB + stage * Cfg::kBBytes + Cfg::kBTile, &tmB2, &full_bar[stage], kb * kBK, n_blk * BN, kEvictLast);
            }
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // the WHOLE warp runs the issue loop on warp-uniform values (uniform registers); one elected lane issues inside the
    // umma_*_warp wrappers -- no per-MMA R2UR broadcasts (isolated GEMMs +1..5 %)
    if (__shfl_sync(0xffffffffu, crank, 0) == 0) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(kBM * kCluster, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = sched_first; tile < num_tiles; tile += sched_step) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t a_desc = make_desc_k_sw128(smem_u32(sA + stage * Cfg::kABytes));
          const uint64_t b_desc = make_desc_k_sw128(smem_u32(sB + stage * Cfg::kBBytes));
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            // +32 bytes per K=16 step inside the 128B swizzle row: +2 in the (addr>>4) field
            if (kCluster > 1) umma_bf16_ss_cg2_warp(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            else umma_bf16_ss_warp(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            if (kSplit) {       // + lo.hi + hi.lo (the lo tiles follow their hi tiles inside the stage)
              const uint64_t a_lo = a_desc + (Cfg::kATile >> 4), b_lo = b_desc + (Cfg::kBTile >> 4);
              umma_bf16_ss_warp(d_tmem, a_lo + 2 * k, b_desc + 2 * k, idesc, 1u);
              umma_bf16_ss_warp(d_tmem, a_desc + 2 * k, b_lo + 2 * k, idesc, 1u);
            }
          }
          if (kCluster > 1) umma_commit_cg2_mcast_warp(&empty_bar[stage], static_cast<uint16_t>(0x3));
          else umma_commit_warp(&empty_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (kCluster > 1) umma_commit_cg2_mcast_warp(&tfull_bar[acc], static_cast<uint16_t>(0x3));
        else umma_commit_warp(&tfull_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int ew = warp - 4;
    const int quad = warp & 3;          // TMEM lane quadrant this warp may access
    const int half = ew >> 2;           // column half of the tile
    uint8_t* wstage = stage_all + ew * Cfg::kEpiWarpBytes;   // 4 KB boxes (1024-byte aligned), or one 32x32 fp32 tile
    constexpr int kChunks = BN / 2 / 32;
    const bool tma_epi = p.epilogue != VSCB_EPI_PATCH_F32;
    const bool out_bf16 = p.epilogue == VSCB200_EPI_BF16;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t box_no = 0;
    for (int tile = sched_first; tile < num_tiles; tile += sched_step) {
      const int vt = p.reverse ? num_tiles - 1 - tile : tile;
        const int m_blk = (vt / p.tiles_n) * kCluster + static_cast<int>(crank), n_blk = vt % p.tiles_n;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int64_t row_base = static_cast<int64_t>(m_blk) * kBM + quad * 32;
      if (tma_epi) {
        // Tensor-map epilogue: this thread holds 32 consecutive columns of its accumulator row; bias and
        // activation are applied in registers, the warp's 32-row box (32 fp32 or 64 bf16 columns = 128 B
        // rows, 128B-swizzled) is written to shared memory once and leaves through the TMA engine -- a
        // plain store, or an fp32 reduce-add into the residual stream (x += tile is done by the L2, the
        // SM never reads x).  Ragged M / N edges are clipped by the tensor map.  Boxes are double-buffered.
        const int step = out_bf16 ? 2 : 1;
#pragma unroll 1
        for (int c = 0; c < kChunks; c += step) {
          const int col0 = half * (BN / 2) + c * 32;
          const int gcol = n_blk * BN + col0;
          if (gcol >= p.N || row_base >= p.M) continue;   // warp-uniform
          uint8_t* box = wstage + (Cfg::kEpiBoxes > 1 ? (box_no & 1u) * 4096u : 0u);
          ++box_no;
          if (lane == 0) tma_store_wait_read<Cfg::kEpiBoxes - 1>();        // the store that last read this buffer has drained
          __syncwarp();
          uint32_t lo_pk[kSplit ? 32 : 1];                // split mode, bf16 output: the lo plane of this thread's 64 columns
          for (int sub = 0; sub < step; ++sub) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN + col0 + sub * 32, v);
            tmem_ld_wait();
            const int gc0 = gcol + sub * 32;
            float rnorm = 1.0f;
            const bool qk_norm = gc0 < p.qk_norm_cols;        // warp-uniform
            if (qk_norm) {
              float ss = 0.f;
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.bias != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + gc0 + 4 * q));
                const float a = __uint_as_float(v[4 * q]) + b4.x, b = __uint_as_float(v[4 * q + 1]) + b4.y;
                const float c2 = __uint_as_float(v[4 * q + 2]) + b4.z, d2 = __uint_as_float(v[4 * q + 3]) + b4.w;
                ss = fmaf(a, a, ss); ss = fmaf(b, b, ss); ss = fmaf(c2, c2, ss); ss = fmaf(d2, d2, ss);
              }
              rnorm = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
              if (gc0 < (p.qk_norm_cols >> 1)) rnorm *= __ldg(p.qscale + (gc0 >> 5));
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float4 o = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                     __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
              if (p.bias != nullptr && gc0 + 4 * q < p.N) {      // N % 8 == 0: a float4 of columns is all-in or all-out
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + gc0 + 4 * q));
                o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
              }
              if (qk_norm) { o.x *= rnorm; o.y *= rnorm; o.z *= rnorm; o.w *= rnorm; }
              if (p.act >= 0) {
                o.x = apply_act<kSplit>(o.x, p.act); o.y = apply_act<kSplit>(o.y, p.act);
                o.z = apply_act<kSplit>(o.z, p.act); o.w = apply_act<kSplit>(o.w, p.act);
              }
              if (out_bf16) {
                v[2 * q] = pack_bf16x2(o.x, o.y);
                v[2 * q + 1] = pack_bf16x2(o.z, o.w);
                if (kSplit) {
                  lo_pk[sub * 16 + 2 * q] = pack_bf16x2(o.x - __uint_as_float(v[2 * q] << 16),
                                                        o.y - __uint_as_float(v[2 * q] & 0xFFFF0000u));
                  lo_pk[sub * 16 + 2 * q + 1] = pack_bf16x2(o.z - __uint_as_float(v[2 * q + 1] << 16),
                                                            o.w - __uint_as_float(v[2 * q + 1] & 0xFFFF0000u));
                }
              } else {
                *reinterpret_cast<float4*>(box + lane * 128 + ((q ^ (lane & 7)) << 4)) = o;
              }
            }
            if (out_bf16) {
#pragma unroll
              for (int q = 0; q < 4; ++q)
                *reinterpret_cast<uint4*>(box + lane * 128 + (((sub * 4 + q) ^ (lane & 7)) << 4)) =
                    make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (p.epilogue == VSCB200_EPI_RESIDUAL_F32) tma_reduce_add_2d(&tmC, box, gcol, static_cast<int>(row_base));
            else tma_store_2d(&tmC, box, gcol, static_cast<int>(row_base));
            tma_store_commit();
          }
          if (kSplit && out_bf16) {                       // the lo plane through the same box, once the hi store has read it
            if (lane == 0) tma_store_wait_read<0>();
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<uint4*>(box + lane * 128 + ((q ^ (lane & 7)) << 4)) =
                  make_uint4(lo_pk[4 * q], lo_pk[4 * q + 1], lo_pk[4 * q + 2], lo_pk[4 * q + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmC2, box, gcol, static_cast<int>(row_base));
              tma_store_commit();
            }
          }
        }
      } else {
      float4* stage = reinterpret_cast<float4*>(wstage);   // 32 rows x 8 float4
#pragma unroll 1
      for (int c = 0; c < kChunks; ++c) {
        const int col0 = half * (BN / 2) + c * 32;
        const int gcol = n_blk * BN + col0;
        uint32_t v[32];
        __syncwarp();                 // tcgen05.ld is .sync.aligned: the warp must be converged
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN + col0, v);
        tmem_ld_wait();
        if (gcol >= p.N) continue;   // warp-uniform
        // Patch-embed epilogue (rows remapped around the class slot, + positional embedding): transpose
        // through the warp's staging tile so that every global access covers whole 128-byte row segments.
#pragma unroll
        for (int q = 0; q < 8; ++q)
          stage[lane * 8 + (q ^ (lane & 7))] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                                           __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
        __syncwarp();
        const int q = lane & 7;
        const int gc = gcol + q * 4;
        const bool col_ok = gc < p.N;
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias != nullptr && col_ok) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + gc));
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = it * 4 + (lane >> 3);
          const int64_t grow = row_base + r;
          float4 o = stage[r * 8 + (q ^ (r & 7))];
          o.x += bias4.x; o.y += bias4.y; o.z += bias4.z; o.w += bias4.w;
          if (grow < p.M && col_ok) {
            const int64_t orow = grow + grow / p.patch_P + 1;
            const float4 pe = __ldg(reinterpret_cast<const float4*>(p.pos + (grow % p.patch_P + 1) * p.N + gc));
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.C) + orow * p.ldc + gc) =
                make_float4(o.x + pe.x, o.y + pe.y, o.z + pe.z, o.w + pe.w);
          }
        }
        __syncwarp();                 // the staging tile is rewritten by the next chunk
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kCluster > 1) mbar_arrive_remote(mapa_u32(smem_u32(&tempty_bar[acc]), 0u));
        else mbar_arrive(&tempty_bar[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) tma_store_wait_all<0>();   // shared memory stays valid until the TMA engine has read every box
  }

  tc_fence_before();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();    // no CTA exits while its peer may still signal into it
  if (warp == 2) {
    if (kCluster > 1) tmem_dealloc_cg2<Cfg::kTmemCols>(tmem_base);
    else tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

struct GemmMaps {
  CUtensorMap A, B, C, A2, B2, C2;     // *2: the lo planes (split mode); copies of the hi maps otherwise
};

template <int BN, int kCluster, bool kSplit>
static int launch_gemm(const GemmMaps& tm, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, kCluster, kSplit>;
  static_assert(Cfg::kSmemBytes <= 227 * 1024, "gemm: shared memory budget");
  VSCB_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_kernel<BN, kCluster, kSplit>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Cfg::kSmemBytes));
  const int sched_tiles = ((p.tiles_m + kCluster - 1) / kCluster) * p.tiles_n;     // (pair) tiles
  const int max_groups = device_sm_count() / kCluster;
  const int groups = sched_tiles < max_groups ? sched_tiles : max_groups;
  ProfScope prof(kProfGemm, stream, 2.0 * static_cast<double>(p.M) * p.N * p.K);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(groups * kCluster);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  VSCB_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_bf16_kernel<BN, kCluster, kSplit>, tm.A, tm.B, tm.C, tm.A2, tm.B2, tm.C2, p));
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

int gemm_bf16(const void* A, const void* W, const float* bias, void* C, int64_t M, int N, int K, int64_t lda,
              int64_t ldw, int64_t ldc, int epilogue, int act, cudaStream_t stream, const float* pos, int patch_P,
              bool reverse, int qk_norm_cols, const float* qscale, const void* A_lo, const void* W_lo, void* C_lo) {
  VSCB_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem");
  VSCB_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "gemm: K/lda/ldw must be multiples of 8 (16-byte TMA strides)");
  VSCB_REQUIRE(N % 8 == 0 && ldc % 8 == 0, "gemm: N/ldc must be multiples of 8");
  VSCB_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(C) & 15) == 0,
               "gemm: operands must be 16-byte aligned");
  VSCB_REQUIRE(epilogue >= 0 && epilogue <= 3, "gemm: bad epilogue");
  VSCB_REQUIRE(epilogue != VSCB_EPI_PATCH_F32 || (pos != nullptr && patch_P > 0), "gemm: patch epilogue needs pos/P");
  const bool split = A_lo != nullptr || W_lo != nullptr;
  VSCB_REQUIRE(!split || (A_lo != nullptr && W_lo != nullptr && (epilogue != VSCB200_EPI_BF16 || C_lo != nullptr)),
               "gemm: the split-bf16 mode needs the lo planes of A and W (and of a bf16 output)");
  VSCB_REQUIRE(!split || ((reinterpret_cast<uintptr_t>(A_lo) & 15) == 0 && (reinterpret_cast<uintptr_t>(W_lo) & 15) == 0 &&
                          (reinterpret_cast<uintptr_t>(C_lo) & 15) == 0), "gemm: lo planes must be 16-byte aligned");
  const int BN = split ? 128 : ((N >= 256 || N > 128) ? 256 : 128);
  const int tiles_m_all = static_cast<int>((M + kBM - 1) / kBM);
  static const int force_cluster = [] { const char* e = getenv("VSCB200_GEMM_CLUSTER"); return e ? atoi(e) : 0; }();
  const int tiles_n_all = (N + BN - 1) / BN;
  const bool pair = split ? false
                          : force_cluster ? (force_cluster == 2 && tiles_m_all >= 2)
                                          : (BN == 256 && tiles_m_all >= 2 &&
                                             static_cast<int64_t>(tiles_m_all) * tiles_n_all >= 2 * device_sm_count());
  GemmMaps tm;
  int rc = make_tmap_2d(&tm.A, A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, K, lda, kBM, kBK, true);
  if (rc) return rc;
  // in pair mode each CTA's TMA box is its half of the W tile (BN/2 rows)
  rc = make_tmap_2d(&tm.B, W, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, N, K, ldw, pair ? BN / 2 : BN, kBK, true);
  if (rc) return rc;
  tm.A2 = tm.A; tm.B2 = tm.B;
  if (split) {
    if ((rc = make_tmap_2d(&tm.A2, A_lo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, K, lda, kBM, kBK, true))) return rc;
    if ((rc = make_tmap_2d(&tm.B2, W_lo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, N, K, ldw, BN, kBK, true))) return rc;
  }
  GemmParams p;
  p.bias = bias; p.C = C; p.M = M; p.N = N; p.K = K; p.ldc = ldc; p.epilogue = epilogue; p.act = act;
  p.pos = pos; p.patch_P = patch_P; p.reverse = reverse ? 1 : 0;
  VSCB_REQUIRE(qk_norm_cols == 0 || (epilogue == VSCB200_EPI_BF16 && qk_norm_cols % 64 == 0 && qk_norm_cols <= N && qscale != nullptr),
               "gemm: the q/k normalisation epilogue needs bf16 output and whole 32-column heads");
  p.qk_norm_cols = qk_norm_cols; p.qscale = qscale;
  p.tiles_m = static_cast<int>((M + kBM - 1) / kBM);
  p.tiles_n = (N + BN - 1) / BN;
  // output tensor map: 32-row boxes of 128 B (64 bf16 / 32 fp32 columns); unused by the patch-embed epilogue
  tm.C = tm.A;
  if (epilogue == VSCB200_EPI_BF16) {
    rc = make_tmap_2d(&tm.C, C, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, N, ldc, 32, 64, true);
  } else if (epilogue != VSCB_EPI_PATCH_F32) {
    rc = make_tmap_2d(&tm.C, C, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, M, N, ldc, 32, 32, true);
  }
  if (rc) return rc;
  tm.C2 = tm.C;
  if (split && epilogue == VSCB200_EPI_BF16 &&
      (rc = make_tmap_2d(&tm.C2, C_lo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, N, ldc, 32, 64, true))) return rc;
  if (split) return launch_gemm<128, 1, true>(tm, p, stream);
  if (pair) return BN == 256 ? launch_gemm<256, 2, false>(tm, p, stream) : launch_gemm<128, 2, false>(tm, p, stream);
  return BN == 256 ? launch_gemm<256, 1, false>(tm, p, stream) : launch_gemm<128, 1, false>(tm, p, stream);
}

}  // namespace vscb200

extern "C" int vscb200_gemm_bf16(const void* A, const void* W, const float* bias, void* C, int64_t M, int N, int K,
                                 int64_t lda, int64_t ldw, int64_t ldc, int epilogue, int act, void* stream) {
  if (epilogue < 0 || epilogue > 2) {
    vscb200::set_last_error("vscb200_gemm_bf16: epilogue must be VSCB200_EPI_{BF16,F32,RESIDUAL_F32}");
    return VSCB200_ERR_INVALID;
  }
  return vscb200::gemm_bf16(A, W, bias, C, M, N, K, lda, ldw, ldc, epilogue, act, static_cast<cudaStream_t>(stream),
                            nullptr, 0, false, 0, nullptr, nullptr, nullptr, nullptr);
}

// Split-bf16 (fp32-equivalent) form: operands as hi / lo bf16 planes; a bf16 output leaves as two planes too.
extern "C" int vscb200_gemm_split(const void* A_hi, const void* A_lo, const void* W_hi, const void* W_lo, const float* bias,
                                  void* C, void* C_lo, int64_t M, int N, int K, int64_t lda, int64_t ldw, int64_t ldc,
                                  int epilogue, int act, void* stream) {
  if (epilogue < 0 || epilogue > 2) {
    vscb200::set_last_error("vscb200_gemm_split: epilogue must be VSCB200_EPI_{BF16,F32,RESIDUAL_F32}");
    return VSCB200_ERR_INVALID;
  }
  return vscb200::gemm_bf16(A_hi, W_hi, bias, C, M, N, K, lda, ldw, ldc, epilogue, act, static_cast<cudaStream_t>(stream),
                            nullptr, 0, false, 0, nullptr, A_lo, W_lo, C_lo);
}
