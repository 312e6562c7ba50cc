// tcgen05 multi-head self-attention for the ViT encoder (T <= 256 tokens, head_dim 64).
//
// One CTA per (frame, head, 128-query tile); two CTAs are resident per SM (256 of the 512 TMEM columns
// each), so one CTA's softmax overlaps the other's tensor work:
//
//   warp 4       TMEM allocator; its lane 0 issues the TMA loads (Q tile, K, V of the (frame, head) straight
//                out of the packed qkv activation) and every MMA:
//                  S[128, Tpad] = Q K^T       tcgen05.mma SS  (A = Q, B = K, both K-major, 4 K-steps)
//                  O[128, 64]   = P V         tcgen05.mma TS  (A = P from TMEM, B = V as an MN-major operand)
//   warps 0-3    softmax: thread = query row.  Pass 1 reads S from TMEM for the row maximum, pass 2 re-reads
//                it, p = 2^((s - max) * scale * log2 e), accumulates the row sum in fp32 and overwrites S in
//                place with P as packed bf16 pairs (tcgen05.st) -- neither S nor P ever leaves the SM.
//                Epilogue: O from TMEM, * 1/sum, bf16, 128 contiguous bytes per row.
//
// Keys >= T (rows of the next frame inside the TMA box, or zero fill past the end of the tensor) are
// masked to p = 0.  Reference: nn.MultiheadAttention at D/train/train_vid_score/video/clip.py:45 (unfused
// bmm + softmax + bmm in torch 1.11).
#include <stdlib.h>

#include "host_util.h"
#include "kernels.h"
#include "ptx.cuh"

namespace vscb200 {

constexpr int kAtcThreads = 160;
constexpr int kAtcTmemCols = 256;
constexpr int kAtcOCol = 128;        // O accumulator columns [128, 192): S is dead when P.V is issued

struct AtcParams {
  __nv_bfloat16* out;
  int T, Tpad, heads, W, mtiles, reverse, dbg;
  int extra;        // tokens past 256 (T = 257 of a ViT-L/14: one class token on a 16 x 16 grid), handled outside the MMAs
  float scale_log2e;
};


// fp32 pair -> packed bf16x2 on the integer ALU (round-half-up on the magnitude: +0x8000, keep the high halves).
// cvt.rn.bf16x2.f32 (F2FP) issues on the XU pipe, which this kernel saturates with its exponentials; finite
// inputs only (probabilities in [0, 256], normalised outputs).
__device__ __forceinline__ uint32_t pack_bf16x2_alu(float lo, float hi) {
  return __byte_perm(__float_as_uint(lo) + 0x8000u, __float_as_uint(hi) + 0x8000u, 0x7632);
}

// Softmax of this thread's query row over S (fp32, TMEM columns [0, Tpad) of `tlane`), P written back in place
// as packed bf16 pairs (columns [0, Tpad/2)).  Returns 1 / row sum.  Warp-collective (tcgen05.ld/st).
// `sx[0..nx)`: raw scores of extra keys that are not in TMEM (computed by the caller); they take part in the max and the
// sum and their probabilities come back in `px`.
__device__ __forceinline__ float softmax_row_tmem(uint32_t tlane, int T, int Tpad, float scale_log2e, int dbg = 0,
                                                  const float* sx = nullptr, int nx = 0, float* px = nullptr) {
  const int nfull = Tpad >> 5;
  const bool tail16 = (Tpad & 31) != 0;
  // ---- pass 1: row maximum over the valid keys (four independent chains; the next chunk's TMEM load is in
  //      flight while the current one is reduced)
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll 1
  for (int c = 0; c < (dbg ? 1 : nfull); ++c) {
    uint32_t v[32];
    tmem_ld_32x32(tlane + c * 32, v);
    tmem_ld_wait();
    const int lim = T - c * 32;
    if (lim >= 32) {                               // warp-uniform fast path: no masking
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        m0 = fmaxf(m0, fmaxf(__uint_as_float(v[j]), __uint_as_float(v[j + 1])));
        m1 = fmaxf(m1, fmaxf(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])));
        m2 = fmaxf(m2, fmaxf(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5])));
        m3 = fmaxf(m3, fmaxf(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7])));
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < lim) m0 = fmaxf(m0, __uint_as_float(v[j]));
    }
  }
  if (tail16) {
    uint32_t v[16];
    tmem_ld_32x16(tlane + nfull * 32, v);
    tmem_ld_wait();
    const int lim = T - nfull * 32;
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (j < lim) m2 = fmaxf(m2, __uint_as_float(v[j]));
  }
  for (int e = 0; e < nx; ++e) m3 = fmaxf(m3, sx[e]);
  const float mxs = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * scale_log2e;
  // ---- pass 2: p = 2^(s*scale - max*scale), row sum (four chains), P -> TMEM (bf16 pairs, in place over S)
  float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
  auto exp_chunk = [&](const uint32_t (&cur)[32], int c) {
    uint32_t pk[16];
    const int lim = T - c * 32;
    if (lim >= 32) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float p0 = ex2_approx(fmaf(__uint_as_float(cur[j]), scale_log2e, -mxs));
        const float p1 = ex2_approx(fmaf(__uint_as_float(cur[j + 1]), scale_log2e, -mxs));
        const float p2 = ex2_approx(fmaf(__uint_as_float(cur[j + 2]), scale_log2e, -mxs));
        const float p3 = ex2_approx(fmaf(__uint_as_float(cur[j + 3]), scale_log2e, -mxs));
        l0 += p0; l1 += p1; l2 += p2; l3 += p3;
        pk[j >> 1] = pack_bf16x2_alu(p0, p1);
        pk[(j >> 1) + 1] = pack_bf16x2_alu(p2, p3);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        float p0 = ex2_approx(fmaf(__uint_as_float(cur[j]), scale_log2e, -mxs));
        float p1 = ex2_approx(fmaf(__uint_as_float(cur[j + 1]), scale_log2e, -mxs));
        if (j >= lim) p0 = 0.f;
        if (j + 1 >= lim) p1 = 0.f;
        l0 += p0; l1 += p1;
        pk[j >> 1] = pack_bf16x2_alu(p0, p1);
      }
    }
    // P of chunk c lands on columns [16c, 16c+16): inside S chunks <= c, all consumed; chunk c+1 (in flight)
    // starts at column 32(c+1) > 16c+16
    tmem_st_32x16(tlane + c * 16, pk);
  };
#pragma unroll 1
  for (int c = 0; c < nfull; ++c) {
    uint32_t v[32];
    tmem_ld_32x32(tlane + c * 32, v);
    tmem_ld_wait();
    exp_chunk(v, c);
  }
  if (tail16) {
    uint32_t v[16], pk[8];
    tmem_ld_32x16(tlane + nfull * 32, v);
    tmem_ld_wait();
    const int lim = T - nfull * 32;
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      float p0 = ex2_approx(fmaf(__uint_as_float(v[j]), scale_log2e, -mxs));
      float p1 = ex2_approx(fmaf(__uint_as_float(v[j + 1]), scale_log2e, -mxs));
      if (j >= lim) p0 = 0.f;
      if (j + 1 >= lim) p1 = 0.f;
      l2 += p0; l3 += p1;
      pk[j >> 1] = pack_bf16x2_alu(p0, p1);
    }
    tmem_st_32x8(tlane + nfull * 16, pk);
  }
  tmem_st_wait();
  for (int e = 0; e < nx; ++e) {
    px[e] = ex2_approx(fmaf(sx[e], scale_log2e, -mxs));
    l3 += px[e];
  }
  return 1.0f / ((l0 + l1) + (l2 + l3));
}

// 32 fp32 accumulator columns * inv_l -> 32 bf16 (64 contiguous bytes of the output row)
__device__ __forceinline__ void store_o_half(__nv_bfloat16* dst, const uint32_t (&v)[32], float inv_l) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 o;
    o.x = pack_bf16x2_alu(__uint_as_float(v[8 * q]) * inv_l, __uint_as_float(v[8 * q + 1]) * inv_l);
    o.y = pack_bf16x2_alu(__uint_as_float(v[8 * q + 2]) * inv_l, __uint_as_float(v[8 * q + 3]) * inv_l);
    o.z = pack_bf16x2_alu(__uint_as_float(v[8 * q + 4]) * inv_l, __uint_as_float(v[8 * q + 5]) * inv_l);
    o.w = pack_bf16x2_alu(__uint_as_float(v[8 * q + 6]) * inv_l, __uint_as_float(v[8 * q + 7]) * inv_l);
    *reinterpret_cast<uint4*>(dst + q * 8) = o;
  }
}

__global__ void __launch_bounds__(kAtcThreads)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, AtcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const int kv_bytes = p.Tpad * 128;
  uint8_t* sQ = smem;
  uint8_t* sK = smem + 16384;
  uint8_t* sV = sK + kv_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kv_bytes);
  uint64_t *bar_qk = bars, *bar_v = bars + 1, *bar_s = bars + 2, *bar_p = bars + 3, *bar_o = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x % p.mtiles, head = blockIdx.x / p.mtiles;
  const int frame = p.reverse ? gridDim.y - 1 - blockIdx.y : blockIdx.y;
  const int row0 = frame * p.T;

  if (warp == 4) {
    if (lane == 0) {
      prefetch_tmap(&tmQ);
      prefetch_tmap(&tmKV);
      mbar_init(bar_qk, 1); mbar_init(bar_v, 1); mbar_init(bar_s, 1); mbar_init(bar_p, 4); mbar_init(bar_o, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<kAtcTmemCols>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      mbar_expect_tx(bar_qk, 16384 + kv_bytes);
      tma_load_2d(sQ, &tmQ, bar_qk, head * 64, row0 + mt * 128, kEvictFirst);
      tma_load_2d(sK, &tmKV, bar_qk, p.W + head * 64, row0, kEvictNormal);
      mbar_expect_tx(bar_v, kv_bytes);
      tma_load_2d(sV, &tmKV, bar_v, 2 * p.W + head * 64, row0, kEvictNormal);
      // S = Q K^T
      mbar_wait(bar_qk, 0);
      tc_fence_after();
      const uint32_t idesc_s = make_idesc_bf16_f32(128, p.Tpad);
      const uint64_t qd = make_desc_k_sw128(smem_u32(sQ)), kd = make_desc_k_sw128(smem_u32(sK));
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base, qd + 2 * k, kd + 2 * k, idesc_s, k ? 1u : 0u);
      umma_commit(bar_s);
      // O = P V
      mbar_wait(bar_v, 0);
      mbar_wait(bar_p, 0);
      tc_fence_after();
      constexpr uint32_t idesc_o = make_idesc_bf16_f32_bmn(128, 64);
      const uint64_t vd = make_desc_mn_sw128(smem_u32(sV));
      const int ksteps = p.Tpad >> 4;
      for (int i = 0; i < ksteps; ++i)
        umma_bf16_ts(tmem_base + kAtcOCol, tmem_base + i * 8, vd + static_cast<uint64_t>(i) * 128, idesc_o, i ? 1u : 0u);
      umma_commit(bar_o);
    }
  } else {
    const int rbase = mt * 128 + warp * 32;             // first query row of this warp inside the frame
    const bool warp_valid = rbase < p.T;                // warp-uniform
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    float inv_l = 0.f;
    mbar_wait(bar_s, 0);
    tc_fence_after();
    if (warp_valid) inv_l = softmax_row_tmem(tlane, p.T, p.Tpad, p.scale_log2e);
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_p);
    // ---- epilogue: O / l -> bf16
    mbar_wait(bar_o, 0);
    tc_fence_after();
    if (warp_valid) {
      const int row = rbase + lane;
      __nv_bfloat16* orow = p.out + (static_cast<int64_t>(row0) + row) * p.W + head * 64;
#pragma unroll
      for (int hc = 0; hc < 2; ++hc) {
        uint32_t v[32];
        tmem_ld_32x32(tlane + kAtcOCol + hc * 32, v);
        tmem_ld_wait();
        if (row < p.T) store_o_half(orow + hc * 32, v, inv_l);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<kAtcTmemCols>(tmem_base);
}

// ---- SIMT helpers for the tokens past 256 (attention_pp_kernel, extra > 0): 64-wide bf16 rows of a 128B-swizzled tile
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
}
// row r of a tile whose rows are 128 bytes (64 bf16), 16-byte chunk c stored at c ^ (r & 7)
__device__ __forceinline__ uint4 tile_chunk(const uint8_t* tile, int r, int c) {
  return *reinterpret_cast<const uint4*>(tile + r * 128 + ((c ^ (r & 7)) << 4));
}
__device__ __forceinline__ float dot64(const uint8_t* ta, int ra, const uint8_t* tb, int rb) {
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float a[8], b[8];
    unpack8(tile_chunk(ta, ra, c), a);
    unpack8(tile_chunk(tb, rb, c), b);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc = fmaf(a[i], b[i], acc);
  }
  return acc;
}

// ------------------------------------------------------------------ persistent ping-pong variant (128 < T <= 264)
// One CTA per SM walks (frame, head) units.  Both 128-query tiles of a unit are in flight at once, each owned
// by one softmax warpgroup with its own 256 TMEM columns, so one warpgroup's exponentials overlap the other's
// MMAs and epilogue; Q/K/V of the NEXT unit are prefetched by a TMA producer warp into the second smem stage.
//
//   warps 0-3 / 4-7   softmax + epilogue warpgroups (tile 0 / tile 1; thread = query row)
//   warp 8            TMA producer: Q [256 x 64], K, V [Tpad x 64] of unit i+1 while unit i computes
//   warp 9            TMEM allocator (512 columns); lane 0 issues every MMA in the order
//                     S0(u) S1(u) PV0(u) PV1(u) S0(u+1) ... each gated by the mbarrier of the data it needs
constexpr int kAppThreads = 320;

template <int kExtra>     // tokens past 256: 0, or 1 (T = 257)
__global__ void __launch_bounds__(kAppThreads, 1)
attention_pp_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                    const __grid_constant__ CUtensorMap tmX, AtcParams p, int n_units) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const int kv_bytes = p.Tpad * 128;
  // T = 256 + extra (ViT-L/14: 257): the 256 x 256 block runs on the tensor cores as usual; the extra KEYS are folded in by
  // the softmax threads (one 64-long dot product per row and key, a rank-1 update of O) and the extra QUERY rows are
  // computed whole by the otherwise idle lanes of the producer warp.  Their Q / K / V rows ride in three 8-row boxes.
  constexpr int extra = kExtra;
  const int stage_bytes = 32768 + 2 * kv_bytes + (extra ? 3072 : 0);
  const int x_off = 32768 + 2 * kv_bytes;     // Qx | Kx | Vx, 1 KB each
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * stage_bytes);
  uint64_t* qk_full = bars;          // [2] per smem stage: Q + K landed
  uint64_t* qk_empty = bars + 2;     //     both S MMAs of the unit have read them
  uint64_t* v_full = bars + 4;       //     V landed
  uint64_t* v_empty = bars + 6;      //     both P.V MMAs of the unit have read it
  uint64_t* s_full = bars + 8;       // [2] per warpgroup
  uint64_t* p_full = bars + 10;
  uint64_t* o_full = bars + 12;
  uint64_t* o_empty = bars + 14;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // uniform for the compiler
  pdl_launch_dependents();
  if (warp == 9) {
    if (lane == 0) {
      prefetch_tmap(&tmQ);
      prefetch_tmap(&tmKV);
      for (int i = 0; i < 2; ++i) {
        // extra tokens: the 8 softmax warps and the producer warp read the staged rows too and release them themselves
        mbar_init(&qk_full[i], 1); mbar_init(&qk_empty[i], extra ? 10 : 1);
        mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], extra ? 10 : 1);
        mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4); mbar_init(&o_full[i], 1); mbar_init(&o_empty[i], 4);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_wait();

  auto unit_frame_head = [&](int unit, int& frame, int& head) {
    const int f = unit / p.heads;
    head = unit % p.heads;
    frame = p.reverse ? (n_units / p.heads - 1 - f) : f;
  };

  if (warp == 8) {
    // extra query rows of one unit, whole warp: scores against the 256 staged keys + the extra keys, softmax, P.V
    auto extra_rows = [&](int unit, int it) {
      const int st = it & 1;
      int frame, head;
      unit_frame_head(unit, frame, head);
      const uint8_t* sQ = smem + st * stage_bytes;
      const uint8_t *sK = sQ + 32768, *sV = sK + kv_bytes, *sQx = sQ + x_off, *sKx = sQx + 1024, *sVx = sKx + 1024;
      mbar_wait(&qk_full[st], (it >> 1) & 1);
      mbar_wait(&v_full[st], (it >> 1) & 1);
      {
        constexpr int e = 0;
        float sc[9];
#pragma unroll
        for (int i = 0; i < 8; ++i) sc[i] = dot64(sQx, e, sK, lane + 32 * i);
        sc[8] = lane < extra ? dot64(sQx, e, sKx, lane) : -INFINITY;
        float mx = sc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) mx = fmaxf(mx, sc[i]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float l = 0.f;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
          sc[i] = ex2_approx((sc[i] - mx) * p.scale_log2e);
          l += sc[i];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
        float a0 = 0.f, a1 = 0.f;                     // output columns 2*lane, 2*lane + 1
        const uint8_t* vcol = sV + (lane & 3) * 4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll 4
          for (int jj = 0; jj < 32; ++jj) {
            const float pj = __shfl_sync(0xffffffffu, sc[i], jj);
            const int j = 32 * i + jj;
            const uint32_t vv = *reinterpret_cast<const uint32_t*>(vcol + j * 128 + (((lane >> 2) ^ (j & 7)) << 4));
            a0 = fmaf(pj, __uint_as_float(vv << 16), a0);
            a1 = fmaf(pj, __uint_as_float(vv & 0xFFFF0000u), a1);
          }
        }
        {
          const float pj = __shfl_sync(0xffffffffu, sc[8], 0);
          const uint32_t vv = *reinterpret_cast<const uint32_t*>(sVx + ((lane >> 2) << 4) + (lane & 3) * 4);
          a0 = fmaf(pj, __uint_as_float(vv << 16), a0);
          a1 = fmaf(pj, __uint_as_float(vv & 0xFFFF0000u), a1);
        }
        const float inv = 1.0f / l;
        __nv_bfloat16* orow = p.out + (static_cast<int64_t>(frame) * p.T + 256 + e) * p.W + head * 64;
        *reinterpret_cast<uint32_t*>(orow + 2 * lane) = pack_bf16x2(a0 * inv, a1 * inv);
      }
      __syncwarp();
      if (lane == 0) { mbar_arrive(&qk_empty[st]); mbar_arrive(&v_empty[st]); }
    };
    int it = 0, prev_unit = -1;
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++it) {
      if (lane == 0) {
        const int st = it & 1;
        int frame, head;
        unit_frame_head(unit, frame, head);
        const int row0 = frame * p.T;
        uint8_t* sQ = smem + st * stage_bytes;
        const uint32_t par = (it >> 1) & 1;
        mbar_wait(&qk_empty[st], par ^ 1);
        mbar_expect_tx(&qk_full[st], 32768 + kv_bytes + (extra ? 2048 : 0));
        tma_load_2d(sQ, &tmQ, &qk_full[st], head * 64, row0, kEvictFirst);
        tma_load_2d(sQ + 32768, &tmKV, &qk_full[st], p.W + head * 64, row0, kEvictFirst);
        if (extra) {
          tma_load_2d(sQ + x_off, &tmX, &qk_full[st], head * 64, row0 + 256, kEvictFirst);
          tma_load_2d(sQ + x_off + 1024, &tmX, &qk_full[st], p.W + head * 64, row0 + 256, kEvictFirst);
        }
        mbar_wait(&v_empty[st], par ^ 1);
        mbar_expect_tx(&v_full[st], kv_bytes + (extra ? 1024 : 0));
        tma_load_2d(sQ + 32768 + kv_bytes, &tmKV, &v_full[st], 2 * p.W + head * 64, row0, kEvictFirst);
        if (extra) tma_load_2d(sQ + x_off + 2048, &tmX, &v_full[st], 2 * p.W + head * 64, row0 + 256, kEvictFirst);
      }
      __syncwarp();
      if (extra && prev_unit >= 0) extra_rows(prev_unit, it - 1);     // while this unit's loads are in flight
      prev_unit = unit;
    }
    if (extra && prev_unit >= 0) extra_rows(prev_unit, it - 1);
  } else if (warp == 9) {
    {   // whole warp, warp-uniform operands, one elected lane issues (umma_*_warp)
      const uint32_t idesc_s = make_idesc_bf16_f32(128, p.Tpad);
      constexpr uint32_t idesc_o = make_idesc_bf16_f32_bmn(128, 64);
      const int ksteps = p.Tpad >> 4;
      const int n_it = (n_units - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
      auto issue_s = [&](int w, int it) {
        const int st = it & 1;
        const uint32_t sQ = smem_u32(smem + st * stage_bytes);
        mbar_wait(&qk_full[st], (it >> 1) & 1);
        mbar_wait(&o_empty[w], (it & 1) ^ 1);           // the warpgroup has drained O of its previous unit
        tc_fence_after();
        const uint64_t qd = make_desc_k_sw128(sQ + w * 16384), kd = make_desc_k_sw128(sQ + 32768);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss_warp(tmem_base + w * 256, qd + 2 * k, kd + 2 * k, idesc_s, k ? 1u : 0u);
        umma_commit_warp(&s_full[w]);
      };
      auto issue_pv = [&](int w, int it) {
        const int st = it & 1;
        const uint64_t vd = make_desc_mn_sw128(smem_u32(smem + st * stage_bytes) + 32768 + kv_bytes);
        mbar_wait(&v_full[st], (it >> 1) & 1);
        mbar_wait(&p_full[w], it & 1);
        tc_fence_after();
        for (int i = 0; i < ksteps; ++i)
          umma_bf16_ts_warp(tmem_base + w * 256 + kAtcOCol, tmem_base + w * 256 + i * 8, vd + static_cast<uint64_t>(i) * 128,
                       idesc_o, i ? 1u : 0u);
        umma_commit_warp(&o_full[w]);
      };
      // The two warpgroups run half a period apart: while one is in its exponentials (XU-bound) the other one's
      // MMAs and epilogue run, so they do not queue on the same XU pipe at the same time.
      if (n_it > 0) issue_s(0, 0);
      for (int it = 0; it < n_it; ++it) {
        if (it > 0) {
          issue_pv(1, it - 1);
          umma_commit_warp(&v_empty[(it - 1) & 1]);
        }
        issue_s(1, it);
        umma_commit_warp(&qk_empty[it & 1]);
        issue_pv(0, it);
        if (it + 1 < n_it) issue_s(0, it + 1);
      }
      if (n_it > 0) {
        issue_pv(1, n_it - 1);
        umma_commit_warp(&v_empty[(n_it - 1) & 1]);
      }
    }
  } else {
    const int w = warp >> 2, quad = warp & 3;
    const int rbase = w * 128 + quad * 32;
    const bool warp_valid = rbase < p.T;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + w * 256;
    int it = 0;
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      int frame, head;
      unit_frame_head(unit, frame, head);
      float inv_l = 0.f;
      float sx[1] = {-INFINITY}, px[1] = {0.f};
      const int st = it & 1;
      const uint8_t* sQ = smem + st * stage_bytes;
      if (extra) {                                       // this row's score against the key past 256
        mbar_wait(&qk_full[st], (it >> 1) & 1);
        sx[0] = dot64(sQ, rbase + lane, sQ + x_off + 1024, 0);
        __syncwarp();
        if (lane == 0) mbar_arrive(&qk_empty[st]);
      }
      mbar_wait(&s_full[w], ph);
      tc_fence_after();
      if (warp_valid) inv_l = softmax_row_tmem(tlane, p.T < 256 ? p.T : 256, p.Tpad, p.scale_log2e, p.dbg, sx, extra, px);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[w]);
      mbar_wait(&o_full[w], ph);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      if (warp_valid) {
        tmem_ld_32x32(tlane + kAtcOCol, v0);
        tmem_ld_32x32(tlane + kAtcOCol + 32, v1);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[w]);          // TMEM is free for S of the next unit before the stores
      if (extra) {                                       // O += p_x * v_x for the keys past 256 (fp32, rank-1 per key)
        mbar_wait(&v_full[st], (it >> 1) & 1);
        const uint8_t* sVx = sQ + x_off + 2048;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float vf[8];
          unpack8(tile_chunk(sVx, 0, c), vf);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (c < 4) v0[8 * c + i] = __float_as_uint(fmaf(px[0], vf[i], __uint_as_float(v0[8 * c + i])));
            else v1[8 * (c - 4) + i] = __float_as_uint(fmaf(px[0], vf[i], __uint_as_float(v1[8 * (c - 4) + i])));
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&v_empty[st]);
      }
      const int row = rbase + lane;
      if (warp_valid && row < p.T) {
        __nv_bfloat16* orow = p.out + (static_cast<int64_t>(frame) * p.T + row) * p.W + head * 64;
        store_o_half(orow, v0, inv_l);
        store_o_half(orow + 32, v1, inv_l);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<512>(tmem_base);
}

bool attention_tc_supported(int T, int head_dim) {
  static const int off = [] { const char* e = getenv("VSCB200_ATTN_MMA_SYNC"); return e ? atoi(e) : 0; }();
  return !off && head_dim == 64 && (((T + 15) & ~15) <= 256 || T == 257);
}

int attention_tc(const void* qkv, void* out, int n_frames, int T, int heads, cudaStream_t stream, bool reverse) {
  const int extra = T > 256 ? T - 256 : 0;            // 257 = 256 + class token: the tokens past 256 bypass the MMAs
  const int Tpad = extra ? 256 : (T + 15) & ~15;
  const int W = heads * 64;
  const int64_t M = static_cast<int64_t>(n_frames) * T;
  CUtensorMap tmQ, tmKV;
  int rc = make_tmap_2d(&tmQ, qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, 3 * W, 3 * W, 128, 64, true);
  if (rc) return rc;
  if ((rc = make_tmap_2d(&tmKV, qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, 3 * W, 3 * W, Tpad, 64, true))) return rc;
  AtcParams p;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.T = T; p.Tpad = Tpad; p.heads = heads; p.W = W; p.mtiles = extra ? 2 : (T + 127) / 128; p.reverse = reverse ? 1 : 0;
  p.extra = extra;
  { const char* e = getenv("VSCB200_ATTN_DBG"); p.dbg = e ? atoi(e) : 0; }
  p.scale_log2e = (1.0f / sqrtf(64.0f)) * 1.4426950408889634f;
  static const int no_pp = [] { const char* e = getenv("VSCB200_ATTN_NO_PINGPONG"); return e ? atoi(e) : 0; }();
  if (p.mtiles == 2 && !no_pp) {
    CUtensorMap tmQ2;
    if ((rc = make_tmap_2d(&tmQ2, qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, 3 * W, 3 * W, 256, 64, true))) return rc;
    CUtensorMap tmX = tmKV;
    if (extra && (rc = make_tmap_2d(&tmX, qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, 3 * W, 3 * W, 8, 64, true))) return rc;
    const int n_units = n_frames * heads;
    const int smem_pp = 2 * (32768 + 2 * Tpad * 128 + (extra ? 3072 : 0)) + 256 + 1024;
    VSCB_CUDA_OK(cudaFuncSetAttribute(attention_pp_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_pp));
    VSCB_CUDA_OK(cudaFuncSetAttribute(attention_pp_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_pp));
    const int grid_pp = n_units < device_sm_count() ? n_units : device_sm_count();
    ProfScope prof(kProfAttention, stream, 4.0 * n_frames * heads * static_cast<double>(T) * T * 64);
    if (extra) VSCB_CUDA_OK(launch_pdl(attention_pp_kernel<1>, dim3(grid_pp), dim3(kAppThreads), smem_pp, stream, tmQ2, tmKV, tmX, p, n_units));
    else VSCB_CUDA_OK(launch_pdl(attention_pp_kernel<0>, dim3(grid_pp), dim3(kAppThreads), smem_pp, stream, tmQ2, tmKV, tmX, p, n_units));
    count_launch();
    VSCB_CUDA_OK(cudaGetLastError());
    return VSCB200_OK;
  }
  // >= 80 KB per CTA keeps residency at two CTAs per SM (each owns 256 of the 512 TMEM columns)
  int smem = 16384 + 2 * Tpad * 128 + 64 + 1024;
  if (smem < 80 * 1024) smem = 80 * 1024;
  VSCB_CUDA_OK(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid(heads * p.mtiles, n_frames);
  ProfScope prof(kProfAttention, stream, 4.0 * n_frames * heads * static_cast<double>(T) * T * 64);
  attention_tc_kernel<<<grid, kAtcThreads, smem, stream>>>(tmQ, tmKV, p);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

}  // namespace vscb200
