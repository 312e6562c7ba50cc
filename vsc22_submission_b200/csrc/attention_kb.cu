// K-blocked tcgen05 attention for segments of up to 640 tokens: the ViT-L/16 @ 384 frames (T = 577, head_dim 64) and the
// 24 x 24 / 12 x 12 windows of SwinV2-L @ 384 (576 / 144 tokens, head_dim 32) of BASELINE configs[3].  S of a whole
// segment does not fit the 512 TMEM columns, so the keys are walked in blocks of 128 with an online softmax.
//
// One persistent CTA per SM walks units = (segment, head) -- head PAIRS for head_dim 32: a 64-column, 128-byte-swizzled
// TMA box of the packed qkv activation holds two 32-wide heads, the second head's operands are addressed by a +64-byte
// start offset inside the swizzle atom (as in swin_attention.cu).  K and V of the unit stay resident in shared memory
// (<= 5 blocks of 128 rows each), the 128-query tiles are processed two at a time, one per softmax warpgroup:
//
//   warp 8   TMA producer: K / V blocks of the unit, then the Q tiles of each tile pair
//   warp 9   TMEM allocator; lane 0 issues every MMA.  Per warpgroup w and key block b:
//              S_w(b)  = Q_w K_b^T          tcgen05.mma SS -> TMEM columns [256w, 256w + 128)
//              O_w    += P_w(b) V_b         tcgen05.mma TS (A = P from TMEM, B = V block as an MN-major operand)
//                                           -> TMEM columns [256w + 128, 256w + 128 + head_dim)
//            issued in the order S_w(b+1) right behind PV_w(b): the tensor pipe executes one thread's MMAs in order, so
//            "S_w(b+1) complete" implies "PV_w(b) complete" and the warpgroup may rescale O while it handles block b+1.
//   warps 0-3 / 4-7   softmax warpgroups, thread = query row.  Per block: pass A reads S from TMEM, forms
//            z = s * scale (ViT) or s + bias + mask (Swin: bias from the head's (2ws-1)^2 table in shared memory, the
//            shifted-window mask recomputed from token coordinates), takes the block maximum; the running shift m only
//            moves when the maximum exceeds it by more than 2^8 (then l and O are rescaled through tcgen05.ld / st --
//            after the first block this is rare); pass B: p = 2^(z - m), row sum, P written back over S as bf16 pairs.
//            After the last block: O * 1 / l -> bf16 rows of the output.
//
// Reference: nn.MultiheadAttention at D/train/train_vid_score/video/clip.py:45 (ViT-L/16 @ 384: clip.py:85-163 with
// input_resolution 384); WindowAttention.forward, swinv2.py:147-185, window partition / mask :232-296.
#include <stdlib.h>

#include "host_util.h"
#include "kernels.h"
#include "ptx.cuh"

namespace vscb200 {

constexpr int kKbThreads = 320;
constexpr int kKbMaxBlocks = 5;          // key blocks of 128 rows resident in shared memory
constexpr int kKbTile = 128 * 128;       // one 128-row x 64-column bf16 box
constexpr int kKbOCol = 128;             // O accumulator columns of a warpgroup's 256-column TMEM region
constexpr float kKbLog2e = 1.4426950408889634f;

struct KbParams {
  __nv_bfloat16* out;        // [M, C]
  const float* tables;       // Swin: [heads][(2ws-1)^2]  16*sigmoid(cpb); nullptr for the ViT
  int C, heads, N;           // N tokens per segment
  int n_segs, hgroups;       // head groups per segment: heads (head_dim 64) or heads / 2 (head_dim 32)
  int nb, ntiles;            // key blocks, 128-query tiles per segment
  float scale_log2e;         // ViT: head_dim^-0.5 * log2 e
  int ws, shift, nWx, nW_per_frame;
};

__device__ __forceinline__ uint32_t kb_pack(float lo, float hi) {
  return __byte_perm(__float_as_uint(lo) + 0x8000u, __float_as_uint(hi) + 0x8000u, 0x7632);
}

// z (log2 domain) of the 32 columns [kt0, kt0 + 32) of this thread's row; columns >= N read -inf.
//   ViT: z = s * scale_log2e.
//   Swin: z = (s + tab[base_i - kterm(key)] + pen[ry(key)][rx(key)]) * log2 e with kterm = yk * TS + xk; along the chunk
//   kterm grows by 1 per key and by another ws - 1 whenever xk wraps (all warp-uniform integer bookkeeping).
template <bool kSwin, bool kMasked>
__device__ __forceinline__ void kb_scores(float (&z)[32], const uint32_t (&v)[32], int kt0, int N, float scale_log2e,
                                          const float* tab, int base_i, int ws, int shift, const float (&pen)[2][2]) {
  if (!kSwin) {
    const int lim = N - kt0;
#pragma unroll
    for (int j = 0; j < 32; ++j) z[j] = j < lim ? __uint_as_float(v[j]) * scale_log2e : -INFINITY;
  } else {
    const int TS = 2 * ws - 1;
    int yk = kt0 / ws, xk = kt0 - yk * ws;
    const float* tp = tab + base_i - (yk * TS + xk);
    const int lim = N - kt0;
    const int edge = ws - shift;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float zz = __uint_as_float(v[j]) + tp[-j];
      if (kMasked) {                                                  // selects, not indexing: pen stays in registers
        const float pe0 = xk >= edge ? pen[0][1] : pen[0][0], pe1 = xk >= edge ? pen[1][1] : pen[1][0];
        zz += yk >= edge ? pe1 : pe0;
      }
      z[j] = j < lim ? zz * kKbLog2e : -INFINITY;
      if (++xk == ws) { xk = 0; ++yk; tp -= ws - 1; }
    }
  }
}

// One key block of this thread's row: S in TMEM columns [tS, tS + 128) -> P (bf16 pairs, in place); running (m, l);
// O (TMEM columns [tO, tO + HD)) rescaled when the shift moves.  Warp-collective.
template <int HD, bool kSwin, bool kMasked>
__device__ __forceinline__ void kb_softmax_block(uint32_t tS, uint32_t tO, int kt_blk, bool first, float& m, float& l,
                                                 const KbParams& p, const float* tab, int base_i, const float (&pen)[2][2]) {
  // ---- pass A: block maximum of z
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
  const int nchunks = min(4, (p.N - kt_blk + 31) >> 5);               // chunks with at least one valid key (warp-uniform)
#pragma unroll 1
  for (int c = 0; c < nchunks; ++c) {
    uint32_t v[32];
    float z[32];
    tmem_ld_32x32(tS + c * 32, v);
    tmem_ld_wait();
    kb_scores<kSwin, kMasked>(z, v, kt_blk + c * 32, p.N, p.scale_log2e, tab, base_i, p.ws, p.shift, pen);
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      m0 = fmaxf(m0, z[j]); m1 = fmaxf(m1, z[j + 1]); m2 = fmaxf(m2, z[j + 2]); m3 = fmaxf(m3, z[j + 3]);
    }
    if (kSwin) {                                                      // keep z: pass B does not redo the bias lookups
      uint32_t zb[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) zb[j] = __float_as_uint(z[j]);
      tmem_st_32x16(tS + c * 32, reinterpret_cast<const uint32_t(&)[16]>(zb[0]));
      tmem_st_32x16(tS + c * 32 + 16, reinterpret_cast<const uint32_t(&)[16]>(zb[16]));
    }
  }
  if (kSwin) tmem_st_wait();
  const float bmax = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
  // ---- the shift only moves when the block maximum exceeds it by more than 2^8 (p stays below 2^8 otherwise)
  const bool move = first || bmax > m + 8.0f;
  if (!first && __any_sync(0xffffffffu, move)) {
    const float alpha = move ? ex2_approx(m - bmax) : 1.0f;
    l *= alpha;
#pragma unroll
    for (int h = 0; h < HD / 32; ++h) {
      uint32_t o[32];
      tmem_ld_32x32(tO + h * 32, o);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
      tmem_st_32x16(tO + h * 32, reinterpret_cast<const uint32_t(&)[16]>(o[0]));
      tmem_st_32x16(tO + h * 32 + 16, reinterpret_cast<const uint32_t(&)[16]>(o[16]));
    }
    tmem_st_wait();
  }
  if (move) m = bmax;
  // ---- pass B: p = 2^(z - m), row sum, P -> TMEM (bf16 pairs over the consumed S columns)
  float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    uint32_t pk[16];
    if (c < nchunks) {
      uint32_t v[32];
      tmem_ld_32x32(tS + c * 32, v);
      tmem_ld_wait();
      const int lim = p.N - (kt_blk + c * 32);
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float q0, q1, q2, q3;
        if (kSwin) {
          q0 = ex2_approx(__uint_as_float(v[j]) - m);
          q1 = ex2_approx(__uint_as_float(v[j + 1]) - m);
          q2 = ex2_approx(__uint_as_float(v[j + 2]) - m);
          q3 = ex2_approx(__uint_as_float(v[j + 3]) - m);
        } else {
          q0 = j < lim ? ex2_approx(fmaf(__uint_as_float(v[j]), p.scale_log2e, -m)) : 0.f;
          q1 = j + 1 < lim ? ex2_approx(fmaf(__uint_as_float(v[j + 1]), p.scale_log2e, -m)) : 0.f;
          q2 = j + 2 < lim ? ex2_approx(fmaf(__uint_as_float(v[j + 2]), p.scale_log2e, -m)) : 0.f;
          q3 = j + 3 < lim ? ex2_approx(fmaf(__uint_as_float(v[j + 3]), p.scale_log2e, -m)) : 0.f;
        }
        l0 += q0; l1 += q1; l2 += q2; l3 += q3;
        pk[j >> 1] = kb_pack(q0, q1);
        pk[(j >> 1) + 1] = kb_pack(q2, q3);
      }
    } else {                                                          // keys past the segment: P = 0
#pragma unroll
      for (int j = 0; j < 16; ++j) pk[j] = 0u;
    }
    tmem_st_32x16(tS + c * 16, pk);
  }
  tmem_st_wait();
  l += (l0 + l1) + (l2 + l3);
}

template <int HD, bool kSwin>
__global__ void __launch_bounds__(kKbThreads, 1)
attention_kb_kernel(const __grid_constant__ CUtensorMap tm, KbParams p, int n_units) {
  constexpr int HH = 64 / HD;                    // heads per 64-column box
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sK = smem;
  uint8_t* sV = sK + p.nb * kKbTile;
  uint8_t* sQ = sV + p.nb * kKbTile;             // two 128-row tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(sQ + 2 * kKbTile);
  uint64_t* kv_full = bars;          // K and V of the unit landed
  uint64_t* kv_empty = bars + 1;     // every MMA of the unit has read them
  uint64_t* q_full = bars + 2;
  uint64_t* q_empty = bars + 3;
  uint64_t* s_full = bars + 4;       // [2] per warpgroup
  uint64_t* p_full = bars + 6;
  uint64_t* o_full = bars + 8;
  uint64_t* o_empty = bars + 10;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  float* tabs = reinterpret_cast<float*>(bars + 14);      // Swin: [2 heads][(2ws-1)^2]

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // uniform for the compiler
  if (warp == 9) {
    if (lane == 0) {
      prefetch_tmap(&tm);
      mbar_init(kv_full, 1); mbar_init(kv_empty, 1); mbar_init(q_full, 1); mbar_init(q_empty, 1);
      for (int i = 0; i < 2; ++i) {
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
  const int nqp = (p.ntiles + 1) >> 1;           // query tile pairs per segment
  // units are head-group major: a CTA keeps a head group's bias tables over consecutive units
  auto unit_seg_hg = [&](int unit, int& seg, int& hg) { hg = unit / p.n_segs; seg = unit - hg * p.n_segs; };

  if (warp == 8) {
    if (lane == 0) {
      uint32_t n_kv = 0, n_q = 0;
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        int seg, hg;
        unit_seg_hg(unit, seg, hg);
        const int row0 = seg * p.N;
        mbar_wait(kv_empty, (n_kv & 1) ^ 1);
        mbar_expect_tx(kv_full, 2 * p.nb * kKbTile);
        for (int b = 0; b < p.nb; ++b) tma_load_2d(sK + b * kKbTile, &tm, kv_full, p.C + hg * 64, row0 + b * 128, kEvictFirst);
        for (int b = 0; b < p.nb; ++b) tma_load_2d(sV + b * kKbTile, &tm, kv_full, 2 * p.C + hg * 64, row0 + b * 128, kEvictFirst);
        ++n_kv;
        for (int qp = 0; qp < nqp; ++qp) {
          const int nt = min(2, p.ntiles - 2 * qp);
          mbar_wait(q_empty, (n_q & 1) ^ 1);
          mbar_expect_tx(q_full, nt * kKbTile);
          for (int w = 0; w < nt; ++w) tma_load_2d(sQ + w * kKbTile, &tm, q_full, hg * 64, row0 + (2 * qp + w) * 128, kEvictFirst);
          ++n_q;
        }
      }
    }
  } else if (warp == 9) {
    {                                                      // whole warp, warp-uniform operands, one elected lane issues
      constexpr uint32_t idesc_s = make_idesc_bf16_f32(128, 128);
      constexpr uint32_t idesc_o = make_idesc_bf16_f32_bmn(128, HD);
      uint32_t n_kv = 0, n_q = 0, n_p[2] = {0, 0}, n_oe[2] = {0, 0};
      const uint32_t sK_u = smem_u32(sK), sV_u = smem_u32(sV), sQ_u = smem_u32(sQ);
      auto issue_s = [&](int w, int hh, int b) {
        const uint64_t qd = make_desc_k_sw128(sQ_u + w * kKbTile + hh * 64), kd = make_desc_k_sw128(sK_u + b * kKbTile + hh * 64);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16_ss_warp(tmem_base + w * 256, qd + 2 * k, kd + 2 * k, idesc_s, k ? 1u : 0u);
        umma_commit_warp(&s_full[w]);
      };
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        mbar_wait(kv_full, n_kv & 1);
        ++n_kv;
        for (int qp = 0; qp < nqp; ++qp) {
          const int nt = min(2, p.ntiles - 2 * qp);
          mbar_wait(q_full, n_q & 1);
          ++n_q;
          for (int hh = 0; hh < HH; ++hh) {
            for (int w = 0; w < nt; ++w) {                 // the warpgroup has drained O of its previous item
              mbar_wait(&o_empty[w], (n_oe[w] & 1) ^ 1);
              ++n_oe[w];
              tc_fence_after();
              issue_s(w, hh, 0);
            }
            for (int b = 0; b < p.nb; ++b) {
              for (int w = 0; w < nt; ++w) {
                mbar_wait(&p_full[w], n_p[w] & 1);
                ++n_p[w];
                tc_fence_after();
                const uint64_t vd = make_desc_mn_sw128(sV_u + b * kKbTile + hh * 64);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  umma_bf16_ts_warp(tmem_base + w * 256 + kKbOCol, tmem_base + w * 256 + i * 8, vd + static_cast<uint64_t>(i) * 128,
                               idesc_o, (b | i) ? 1u : 0u);
                if (b + 1 < p.nb) issue_s(w, hh, b + 1);
                else umma_commit_warp(&o_full[w]);
              }
            }
          }
          umma_commit_warp(q_empty);                            // every S MMA of the tile pair has read Q
        }
        umma_commit_warp(kv_empty);
      }
    }
  } else {
    const int w = warp >> 2, quad = warp & 3;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + w * 256;
    const int TS = 2 * p.ws - 1;
    uint32_t n_s = 0, n_o = 0;
    int loaded_hg = -1;
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
      int seg, hg;
      unit_seg_hg(unit, seg, hg);
      if (kSwin && hg != loaded_hg) {                    // bias tables of the head pair -> shared memory (both warpgroups)
        named_bar_sync(1, 256);                          // nobody still reads the old tables
        for (int e = threadIdx.x; e < 2 * TS * TS; e += 256) tabs[e] = p.tables[static_cast<int64_t>(hg) * 2 * TS * TS + e];
        named_bar_sync(1, 256);
        loaded_hg = hg;
      }
      bool edge_y = false, edge_x = false;
      if (kSwin && p.shift > 0) {                        // shifted-window regions: only the last window row / column of a frame
        const int wf = seg % p.nW_per_frame;
        edge_y = (wf / p.nWx) == (p.nW_per_frame / p.nWx) - 1;
        edge_x = (wf % p.nWx) == p.nWx - 1;
      }
      const bool masked = edge_y || edge_x;
      for (int qp = 0; qp < nqp; ++qp) {
        const int tile = 2 * qp + w;
        if (tile >= p.ntiles) continue;                  // warpgroup-uniform: the issuer skips this warpgroup too
        const int i_tok = tile * 128 + quad * 32 + lane;
        const bool row_ok = i_tok < p.N;
        const bool warp_valid = tile * 128 + quad * 32 < p.N;
        int base_i = 0;
        float pen[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
        if (kSwin) {
          const int it = row_ok ? i_tok : 0;
          const int yi = it / p.ws, xi = it - yi * p.ws;
          base_i = (yi + p.ws - 1) * TS + xi + p.ws - 1;
          const int ry_i = (edge_y && yi >= p.ws - p.shift) ? 1 : 0, rx_i = (edge_x && xi >= p.ws - p.shift) ? 1 : 0;
#pragma unroll
          for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) pen[a][b] = ((edge_y && a != ry_i) || (edge_x && b != rx_i)) ? -100.0f : 0.0f;
        }
#pragma unroll 1
        for (int hh = 0; hh < HH; ++hh) {
          const float* tab = tabs + hh * TS * TS;
          float m = -INFINITY, l = 0.f;
#pragma unroll 1
          for (int b = 0; b < p.nb; ++b) {
            mbar_wait(&s_full[w], n_s & 1);
            ++n_s;
            tc_fence_after();
            if (warp_valid) {
              if (masked) kb_softmax_block<HD, kSwin, kSwin>(tlane, tlane + kKbOCol, b * 128, b == 0, m, l, p, tab, base_i, pen);
              else kb_softmax_block<HD, kSwin, false>(tlane, tlane + kKbOCol, b * 128, b == 0, m, l, p, tab, base_i, pen);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[w]);
          }
          mbar_wait(&o_full[w], n_o & 1);
          ++n_o;
          tc_fence_after();
          uint32_t o[HD];
          if (warp_valid) {
#pragma unroll
            for (int h = 0; h < HD / 32; ++h) tmem_ld_32x32(tlane + kKbOCol + h * 32, reinterpret_cast<uint32_t(&)[32]>(o[h * 32]));
            tmem_ld_wait();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&o_empty[w]);      // TMEM is free for the next item's S before the stores
          if (warp_valid && row_ok) {
            const float inv_l = 1.0f / l;
            __nv_bfloat16* orow = p.out + (static_cast<int64_t>(seg) * p.N + i_tok) * p.C + hg * 64 + hh * HD;
#pragma unroll
            for (int q = 0; q < HD / 8; ++q) {
              uint4 u;
              u.x = kb_pack(__uint_as_float(o[8 * q]) * inv_l, __uint_as_float(o[8 * q + 1]) * inv_l);
              u.y = kb_pack(__uint_as_float(o[8 * q + 2]) * inv_l, __uint_as_float(o[8 * q + 3]) * inv_l);
              u.z = kb_pack(__uint_as_float(o[8 * q + 4]) * inv_l, __uint_as_float(o[8 * q + 5]) * inv_l);
              u.w = kb_pack(__uint_as_float(o[8 * q + 6]) * inv_l, __uint_as_float(o[8 * q + 7]) * inv_l);
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

bool attention_kb_supported(int N, int head_dim) {
  static const int off = [] { const char* e = getenv("VSCB200_ATTN_NO_KB"); return e ? atoi(e) : 0; }();
  return !off && (head_dim == 64 || head_dim == 32) && N > 128 && N <= 128 * kKbMaxBlocks;
}

// qkv: [n_segs * N, 3C] bf16 (q | k | v, heads head_dim-wide contiguous); out: [n_segs * N, C] bf16.
// head_dim 64: softmax(q.k * scale) v.  head_dim 32 (Swin-V2): q, k arrive cosine-normalised with the logit scale in q;
// tables [heads][(2ws-1)^2]; rows of a segment are the tokens of one ws x ws window in row-major order.
int attention_kb(const void* qkv, void* out, int64_t n_segs, int N, int heads, int head_dim, float scale, const float* tables,
                 int ws, int shift, int nWx, int nW_per_frame, cudaStream_t stream) {
  VSCB_REQUIRE(attention_kb_supported(N, head_dim), "attention_kb: unsupported segment length / head_dim");
  const bool swin = tables != nullptr;
  VSCB_REQUIRE(swin == (head_dim == 32), "attention_kb: head_dim 32 is the Swin-V2 form (bias tables), 64 the ViT form");
  VSCB_REQUIRE(!swin || (heads % 2 == 0 && ws * ws == N), "attention_kb: Swin form needs an even head count and N = ws * ws");
  const int C = heads * head_dim;
  const int64_t M = n_segs * N;
  const int hgroups = head_dim == 64 ? heads : heads / 2;
  VSCB_REQUIRE(M < (1ll << 31) && n_segs * hgroups < (1ll << 31), "attention_kb: problem too large");
  CUtensorMap tm;
  int rc = make_tmap_2d(&tm, qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, 3 * C, 3 * C, 128, 64, true);
  if (rc) return rc;
  KbParams p;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.tables = tables; p.C = C; p.heads = heads; p.N = N; p.n_segs = static_cast<int>(n_segs); p.hgroups = hgroups;
  p.nb = (N + 127) / 128; p.ntiles = (N + 127) / 128;
  p.scale_log2e = scale * kKbLog2e;
  p.ws = swin ? ws : 1; p.shift = shift; p.nWx = nWx > 0 ? nWx : 1; p.nW_per_frame = nW_per_frame > 0 ? nW_per_frame : 1;
  const int n_units = static_cast<int>(n_segs) * hgroups;
  const int TS = 2 * p.ws - 1;
  const int smem = (2 * p.nb + 2) * kKbTile + 14 * 8 + (swin ? 2 * TS * TS * 4 : 0) + 64 + 1024;
  VSCB_REQUIRE(smem <= 232448, "attention_kb: shared memory");
  const int grid = n_units < device_sm_count() ? n_units : device_sm_count();
  ProfScope prof(kProfAttention, stream, 4.0 * static_cast<double>(M) * N * C);
#define VSCB_KB_LAUNCH(HD, SW)                                                                                                   \
  do {                                                                                                                           \
    VSCB_CUDA_OK(cudaFuncSetAttribute(attention_kb_kernel<HD, SW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));          \
    attention_kb_kernel<HD, SW><<<grid, kKbThreads, smem, stream>>>(tm, p, n_units);                                             \
  } while (0)
  if (swin) VSCB_KB_LAUNCH(32, true);
  else VSCB_KB_LAUNCH(64, false);
#undef VSCB_KB_LAUNCH
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

}  // namespace vscb200
