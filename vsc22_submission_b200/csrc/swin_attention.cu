// Swin-V2 window attention on tcgen05 (head_dim 32, windows of N = ws*ws tokens, ws in {4, 8, 16}).
//
// Reference: WindowAttention.forward, swinv2.py:147-185 -- cosine attention
//     softmax( normalize(q).normalize(k)^T * exp(min(logit_scale, log 100)) + 16*sigmoid(cpb_mlp(.))[index] + mask ) v
// plus the (shifted) window partition of SwinTransformerBlock.forward :273-296.  Here:
//   * q and k arrive already L2-normalised per head, q pre-multiplied by the head's logit scale (fused into
//     the QKV GEMM epilogue, gemm.cu), rows already in window order (swin_kernels.cu gather);
//   * the relative-position bias is NOT materialised as [nH, N, N]: the (2ws-1)^2-entry table of the head
//     (computed once per weight load) sits in shared memory and is indexed per score;
//   * the shifted-window mask (-100 between different regions of an edge window, :232-255) is recomputed from
//     the token coordinates.
// Same persistent ping-pong structure as attention_pp_kernel (attention_tc.cu): one CTA per SM walks units =
// (window, head PAIR) -- a pair because a 64-column, 128-byte-swizzled TMA box of the packed qkv activation
// holds two 32-wide heads; the second head's operands are addressed by a +64-byte start offset inside the
// swizzle atom.  Per unit each softmax warpgroup (one per 128-query tile) runs the two heads back to back.
#include <stdlib.h>

#include "host_util.h"
#include "kernels.h"
#include "ptx.cuh"

namespace vscb200 {

constexpr int kSwThreads = 320;      // warps 0-3 / 4-7 softmax warpgroups, warp 8 TMA, warp 9 MMA + TMEM
constexpr int kSwOCol = 128;         // O accumulator columns [128, 160) of a warpgroup's 256-column region

struct SwinAttnParams {
  __nv_bfloat16* out;        // [M, C] window order
  const float* tables;       // [heads][(2ws-1)^2]  16*sigmoid(cpb)
  int C, heads, N, n_windows_total, nWx, nW_per_frame, shift, ws;
};

__device__ __forceinline__ uint32_t pack_bf16x2_alu2(float lo, float hi) {
  return __byte_perm(__float_as_uint(lo) + 0x8000u, __float_as_uint(hi) + 0x8000u, 0x7632);
}

// z[jj] = S[jj] + bias(i, j) (+ mask) for the NCOL columns j = 32c + jj of chunk c (natural-log domain).
//   bias: tab[(yi - yj + WS-1) * TS + (xi - xj + WS-1)] = tp[-(yj_local * TS + xj)] with tp = tab + base_i - c * rows_per_chunk * TS:
//         one LDS with a compile-time offset per score;
//   mask: the shift is WS / 2, so a column's x-region is a compile-time function of jj and its y-region is uniform over
//         the chunk (WS >= 8) or compile-time (WS = 4): pen[ry][rx] (-100 where the region differs from the row's) is
//         precomputed per thread and only edge windows of shifted blocks take this variant.
template <int WS, int NCOL, bool kMasked>
__device__ __forceinline__ void add_bias_mask(float (&z)[32], const uint32_t (&v)[32], const float* tab, int base_i, int c,
                                              const float (&pen)[2][2]) {
  constexpr int TS = 2 * WS - 1;
  constexpr int kRowsPerChunk = 32 / WS > 0 ? 32 / WS : 1;      // window rows covered by a 32-column chunk
  const float* tp = tab + base_i - c * kRowsPerChunk * TS;
  const bool ry_chunk = c * kRowsPerChunk >= WS / 2;             // WS >= 8: all rows of the chunk are on one side
  const float pen_c0 = ry_chunk ? pen[1][0] : pen[0][0], pen_c1 = ry_chunk ? pen[1][1] : pen[0][1];
#pragma unroll
  for (int jj = 0; jj < NCOL; ++jj) {
    const int yl = jj / WS, xj = jj % WS;                        // compile-time
    float zz = __uint_as_float(v[jj]) + tp[-(yl * TS + xj)];
    if (kMasked) {
      const bool rx = xj >= WS / 2;
      if (WS >= 8) zz += rx ? pen_c1 : pen_c0;
      else zz += pen[yl >= WS / 2 ? 1 : 0][rx ? 1 : 0];
    }
    z[jj] = zz;
  }
}

// Softmax of one query row of a window over S (TMEM, fp32) + bias + mask; P written back in place as bf16 pairs.
template <int WS, bool kMasked>
__device__ __forceinline__ float softmax_row(uint32_t tlane, const float* tab, int base_i, const float (&pen)[2][2]) {
  constexpr int N = WS * WS;
  constexpr int kChunks = N >= 32 ? N / 32 : 1;
  constexpr int kCols = N >= 32 ? 32 : 16;
  // ---- pass 1: row maximum of z = s + bias + mask
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll 1
  for (int c = 0; c < kChunks; ++c) {
    uint32_t v[32];
    float z[32];
    if (kCols == 32) tmem_ld_32x32(tlane + c * 32, v);
    else tmem_ld_32x16(tlane, reinterpret_cast<uint32_t(&)[16]>(v));
    tmem_ld_wait();
    add_bias_mask<WS, kCols, kMasked>(z, v, tab, base_i, c, pen);
#pragma unroll
    for (int j = 0; j < kCols; j += 4) {
      m0 = fmaxf(m0, z[j]); m1 = fmaxf(m1, z[j + 1]); m2 = fmaxf(m2, z[j + 2]); m3 = fmaxf(m3, z[j + 3]);
    }
  }
  const float mxs = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * 1.4426950408889634f;
  // ---- pass 2: p = exp(z - max), row sum, P -> TMEM as bf16 pairs (in place over S)
  float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll 1
  for (int c = 0; c < kChunks; ++c) {
    uint32_t v[32], pk[16];
    float z[32];
    if (kCols == 32) tmem_ld_32x32(tlane + c * 32, v);
    else tmem_ld_32x16(tlane, reinterpret_cast<uint32_t(&)[16]>(v));
    tmem_ld_wait();
    add_bias_mask<WS, kCols, kMasked>(z, v, tab, base_i, c, pen);
#pragma unroll
    for (int j = 0; j < kCols; j += 4) {
      const float p0 = ex2_approx(fmaf(z[j], 1.4426950408889634f, -mxs));
      const float p1 = ex2_approx(fmaf(z[j + 1], 1.4426950408889634f, -mxs));
      const float p2 = ex2_approx(fmaf(z[j + 2], 1.4426950408889634f, -mxs));
      const float p3 = ex2_approx(fmaf(z[j + 3], 1.4426950408889634f, -mxs));
      l0 += p0; l1 += p1; l2 += p2; l3 += p3;
      pk[j >> 1] = pack_bf16x2_alu2(p0, p1);
      pk[(j >> 1) + 1] = pack_bf16x2_alu2(p2, p3);
    }
    if (kCols == 32) tmem_st_32x16(tlane + c * 16, pk);
    else tmem_st_32x8(tlane, reinterpret_cast<uint32_t(&)[8]>(pk));
  }
  tmem_st_wait();
  return 1.0f / ((l0 + l1) + (l2 + l3));
}

template <int WS>
__global__ void __launch_bounds__(kSwThreads, 1)
swin_attention_kernel(const __grid_constant__ CUtensorMap tmQKV, SwinAttnParams p, int n_units) {
  constexpr int N = WS * WS;                   // tokens per window: 16 / 64 / 256
  constexpr int TS = 2 * WS - 1;
  constexpr int kTile = N * 128;               // N rows x 64 bf16 (two heads)
  constexpr int kStage = 3 * kTile;
  constexpr int kMT = N > 128 ? 2 : 1;         // 128-query tiles per window
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  constexpr int kStageAl = (kStage + 1023) & ~1023;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kStageAl);
  uint64_t* qk_full = bars;          // [2] per smem stage
  uint64_t* qk_empty = bars + 2;
  uint64_t* v_full = bars + 4;
  uint64_t* v_empty = bars + 6;
  uint64_t* s_full = bars + 8;       // [2] per warpgroup
  uint64_t* p_full = bars + 10;
  uint64_t* o_full = bars + 12;
  uint64_t* o_empty = bars + 14;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  float* tabs = reinterpret_cast<float*>(bars + 18);    // [2 warpgroups][2 heads][TS*TS]

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // uniform for the compiler
  if (warp == 9) {
    if (lane == 0) {
      prefetch_tmap(&tmQKV);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&qk_full[i], 1); mbar_init(&qk_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
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
  const int hpairs = p.heads >> 1;
  const int n_it = (n_units - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (warp == 8) {
    if (lane == 0) {
      int it = 0;
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++it) {
        const int st = it & 1;
        const int win = unit / hpairs, hp = unit % hpairs;
        uint8_t* sQ = smem + st * kStageAl;
        const uint32_t par = (it >> 1) & 1;
        mbar_wait(&qk_empty[st], par ^ 1);
        mbar_expect_tx(&qk_full[st], 2 * kTile);
        tma_load_2d(sQ, &tmQKV, &qk_full[st], hp * 64, win * N, kEvictFirst);
        tma_load_2d(sQ + kTile, &tmQKV, &qk_full[st], p.C + hp * 64, win * N, kEvictFirst);
        mbar_wait(&v_empty[st], par ^ 1);
        mbar_expect_tx(&v_full[st], kTile);
        tma_load_2d(sQ + 2 * kTile, &tmQKV, &v_full[st], 2 * p.C + hp * 64, win * N, kEvictFirst);
      }
    }
  } else if (warp == 9) {
    {   // whole warp, warp-uniform operands, one elected lane issues (umma_*_warp)
      // items = (unit, head of the pair); both warpgroups work on the same item sequence, half a period apart
      constexpr uint32_t idesc_s = make_idesc_bf16_f32(128, N);
      constexpr uint32_t idesc_o = make_idesc_bf16_f32_bmn(128, 32);
      const int n_items = 2 * n_it;
      auto issue_s = [&](int w, int item) {
        const int it = item >> 1, hh = item & 1, st = it & 1;
        const uint32_t sQ = smem_u32(smem + st * kStageAl);
        if (hh == 0) mbar_wait(&qk_full[st], (it >> 1) & 1);
        mbar_wait(&o_empty[w], (item & 1) ^ 1);
        tc_fence_after();
        const uint64_t qd = make_desc_k_sw128(sQ + w * 16384 + hh * 64), kd = make_desc_k_sw128(sQ + kTile + hh * 64);
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_bf16_ss_warp(tmem_base + w * 256, qd + 2 * k, kd + 2 * k, idesc_s, k ? 1u : 0u);
        umma_commit_warp(&s_full[w]);
      };
      auto issue_pv = [&](int w, int item) {
        const int it = item >> 1, hh = item & 1, st = it & 1;
        const uint64_t vd = make_desc_mn_sw128(smem_u32(smem + st * kStageAl) + 2 * kTile + hh * 64);
        if (hh == 0) mbar_wait(&v_full[st], (it >> 1) & 1);
        mbar_wait(&p_full[w], item & 1);
        tc_fence_after();
#pragma unroll 1
        for (int i = 0; i < N / 16; ++i)
          umma_bf16_ts_warp(tmem_base + w * 256 + kSwOCol, tmem_base + w * 256 + i * 8, vd + static_cast<uint64_t>(i) * 128,
                       idesc_o, i ? 1u : 0u);
        umma_commit_warp(&o_full[w]);
      };
      if (kMT == 2) {
        if (n_items > 0) issue_s(0, 0);
        for (int item = 0; item < n_items; ++item) {
          if (item > 0) {
            issue_pv(1, item - 1);
            if ((item - 1) & 1) umma_commit_warp(&v_empty[((item - 1) >> 1) & 1]);
          }
          issue_s(1, item);
          if (item & 1) umma_commit_warp(&qk_empty[(item >> 1) & 1]);
          issue_pv(0, item);
          if (item + 1 < n_items) issue_s(0, item + 1);
        }
        if (n_items > 0) {
          issue_pv(1, n_items - 1);
          umma_commit_warp(&v_empty[((n_items - 1) >> 1) & 1]);
        }
      } else {
        for (int item = 0; item < n_items; ++item) {
          issue_s(0, item);
          if (item & 1) umma_commit_warp(&qk_empty[(item >> 1) & 1]);
          issue_pv(0, item);
          if (item & 1) umma_commit_warp(&v_empty[(item >> 1) & 1]);
        }
      }
    }
  } else {
    const int w = warp >> 2, quad = warp & 3;
    if (w < kMT) {
      const int i_tok = w * 128 + quad * 32 + lane;            // this thread's query token inside the window
      const bool warp_valid = w * 128 + quad * 32 < N;
      const bool row_ok = i_tok < N;
      const int yi = (i_tok / WS) % WS, xi = i_tok % WS;
      const int base_i = (yi + WS - 1) * TS + xi + WS - 1;
      const uint32_t tlane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + w * 256;
      float* my_tabs = tabs + w * 2 * TS * TS;
      const int wg_tid = threadIdx.x & 127;
      int loaded_hp = -1;
      int it = 0;
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++it) {
        const int win = unit / hpairs, hp = unit % hpairs;
        if (hp != loaded_hp) {                                 // bias tables of the two heads -> this warpgroup's smem
          named_bar_sync(1 + w, 128);                          // nobody still reads the old tables
          for (int e = wg_tid; e < 2 * TS * TS; e += 128) my_tabs[e] = p.tables[hp * 2 * TS * TS + e];
          named_bar_sync(1 + w, 128);
          loaded_hp = hp;
        }
        // shifted-window regions: only the last window row / column of a frame mixes regions
        const int wf = win % p.nW_per_frame;
        const bool edge_y = p.shift > 0 && (wf / p.nWx) == (p.nW_per_frame / p.nWx) - 1;
        const bool edge_x = p.shift > 0 && (wf % p.nWx) == p.nWx - 1;
        const bool masked = edge_y || edge_x;
        const int ry_i = (edge_y && yi >= WS / 2) ? 1 : 0, rx_i = (edge_x && xi >= WS / 2) ? 1 : 0;
        float pen[2][2];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int b = 0; b < 2; ++b)
            pen[a][b] = ((edge_y && a != ry_i) || (edge_x && b != rx_i)) ? -100.0f : 0.0f;
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) {
          const uint32_t ph = (2 * it + hh) & 1;
          const float* tab = my_tabs + hh * TS * TS;
          float inv_l = 0.f;
          mbar_wait(&s_full[w], ph);
          tc_fence_after();
          if (warp_valid) {
            if (masked) inv_l = softmax_row<WS, true>(tlane, tab, row_ok ? base_i : (WS - 1) * TS + WS - 1, pen);
            else inv_l = softmax_row<WS, false>(tlane, tab, row_ok ? base_i : (WS - 1) * TS + WS - 1, pen);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[w]);
          mbar_wait(&o_full[w], ph);
          tc_fence_after();
          uint32_t o[32];
          if (warp_valid) {
            tmem_ld_32x32(tlane + kSwOCol, o);
            tmem_ld_wait();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&o_empty[w]);
          if (warp_valid && row_ok) {
            __nv_bfloat16* orow = p.out + (static_cast<int64_t>(win) * N + i_tok) * p.C + (hp * 2 + hh) * 32;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 u;
              u.x = pack_bf16x2_alu2(__uint_as_float(o[8 * q]) * inv_l, __uint_as_float(o[8 * q + 1]) * inv_l);
              u.y = pack_bf16x2_alu2(__uint_as_float(o[8 * q + 2]) * inv_l, __uint_as_float(o[8 * q + 3]) * inv_l);
              u.z = pack_bf16x2_alu2(__uint_as_float(o[8 * q + 4]) * inv_l, __uint_as_float(o[8 * q + 5]) * inv_l);
              u.w = pack_bf16x2_alu2(__uint_as_float(o[8 * q + 6]) * inv_l, __uint_as_float(o[8 * q + 7]) * inv_l);
              *reinterpret_cast<uint4*>(orow + q * 8) = u;
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<512>(tmem_base);
}

template <int WS>
static int launch_swin_attention(const CUtensorMap& tm, const SwinAttnParams& p, int n_units, cudaStream_t stream) {
  constexpr int N = WS * WS;
  constexpr int kStageAl = (3 * N * 128 + 1023) & ~1023;
  // The S MMA always reads 128 query rows (16 KB) from the Q tile base; for windows of fewer tokens the rows past
  // the tile are other staged bytes whose results are discarded -- keep them inside the allocation.
  const int smem = 2 * kStageAl + 18 * 8 + 4 * (2 * WS - 1) * (2 * WS - 1) * 4 + 1024 + (N < 128 ? 16384 : 0);
  VSCB_CUDA_OK(cudaFuncSetAttribute(swin_attention_kernel<WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int grid = n_units < device_sm_count() ? n_units : device_sm_count();
  swin_attention_kernel<WS><<<grid, kSwThreads, smem, stream>>>(tm, p, n_units);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

// qkv: [M, 3C] bf16, rows in window order (M = n_windows_total * ws*ws), q/k cosine-normalised, q scaled.
// tables: [heads][(2ws-1)^2].  out: [M, C] bf16 (window order).
int swin_attention(const void* qkv, void* out, const float* tables, int64_t n_windows_total, int nW_per_frame, int nWx, int ws,
                   int shift, int heads, cudaStream_t stream) {
  VSCB_REQUIRE(ws == 4 || ws == 8 || ws == 16, "swin_attention: window side must be 4, 8 or 16");
  VSCB_REQUIRE(heads >= 2 && heads % 2 == 0, "swin_attention: head count must be even (head_dim 32, processed in pairs)");
  const int C = heads * 32, N = ws * ws;
  const int64_t M = n_windows_total * N;
  VSCB_REQUIRE(M < (1ll << 31) && n_windows_total * (heads / 2) < (1ll << 31), "swin_attention: problem too large");
  CUtensorMap tm;
  int rc = make_tmap_2d(&tm, qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, 3 * C, 3 * C, N, 64, true);
  if (rc) return rc;
  SwinAttnParams p;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.tables = tables; p.C = C; p.heads = heads; p.N = N; p.n_windows_total = static_cast<int>(n_windows_total);
  p.nWx = nWx; p.nW_per_frame = nW_per_frame; p.shift = shift; p.ws = ws;
  const int n_units = static_cast<int>(n_windows_total) * (heads / 2);
  ProfScope prof(kProfAttention, stream, 4.0 * static_cast<double>(M) * N * C);
  if (ws == 16) return launch_swin_attention<16>(tm, p, n_units, stream);
  if (ws == 8) return launch_swin_attention<8>(tm, p, n_units, stream);
  return launch_swin_attention<4>(tm, p, n_units, stream);
}

}  // namespace vscb200
