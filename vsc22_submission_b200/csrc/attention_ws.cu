// Streaming tcgen05 attention for the ViT encoders (head_dim 64, 129 .. 640 tokens per frame) and the large Swin-V2 windows
// (head_dim 32, 144 / 576 tokens): the NW softmax warpgroups of a CTA run INDEPENDENT streams of (segment, head, 128-query
// tile) items, each with its own MMA-issuing warp, keys in blocks of KB with D S buffers per warpgroup in TMEM, so a
// warpgroup never waits for an MMA round trip between blocks.  Shapes (template <HD, kSwin, NW, KB, D>): ViT 3 warpgroups x
// 32-key blocks x 2 buffers (16 warps), Swin 2 x 64 x 2 (11 warps); VSCB200_ATTN_WS_SHAPE selects the others.
//
//   warps 0 .. 4 NW-1     softmax warpgroup w = warp / 4, thread = query row.  Per key block (KB / 32 chunks of 32 columns):
//                     wait S(b); per chunk ONE pass over registers: chunk maximum (FMNMX3 tree), p = 2^(s * scale - m)
//                     (predicated off for the padding keys of a segment's last chunk), P over S in place as bf16 pairs
//                     (truncation: one PRMT per pair); arrive P(b).  The running shift m only moves when a chunk maximum
//                     exceeds it by more than 2^8; then O, the row sum and the P chunks already written are rescaled
//                     through tcgen05.ld / st -- after the first chunks of an item this is rare.  The ROW SUM is
//                     accumulated by the tensor core from the very P the output uses, so normalisation cancels the
//                     truncation bias and the softmax threads carry no adds: in the ViT form by the SAME MMA as O (an
//                     N = 80 MN-major operand whose second 64-column block -- the descriptor's LBO -- is a tile of bf16
//                     ones: O | l land in 80 adjacent TMEM columns), in the Swin form by a second N = 16 MMA.  The
//                     epilogue of an item (O * 1 / l -> bf16 rows) is DEFERRED until after the softmax of the next item's
//                     first block: by then the last P.V of the item has long completed.
//   warps 4 NW .. 5 NW-1  MMA issuer of warpgroup w: the whole warp runs the loop on warp-uniform values, one elected lane
//                     issues (the first issuer warp also owns the TMEM allocation).  Rolling order S(g), S(g+1), [P(g)]
//                     PV(g), S(g+2), [P(g+1)] PV(g+1), S(g+3) ... over the warpgroup's flattened block sequence g (blocks
//                     of consecutive items follow each other without a drain): S runs D blocks ahead of the softmax.
//                     TMEM per warpgroup: D x KB columns of S, 64 of O, 16 of row sums.
//   warp 5 NW             TMA producer: K / V of a (segment, head group) unit into one of two shared-memory stages (one
//                     stage for long segments), the Q tile of every item into its warpgroup's Q buffers.
//
// The tensor pipe executes one thread's MMAs in issue order, so S(g+D) -- issued right behind PV(g) -- cannot overwrite
// the buffer P(g) is read from.  PV(g-1) may still be accumulating into O when the softmax of block g wants to rescale
// O: that (rare) path first waits on a barrier committed behind every PV.  What bounds the kernel (exponentials and the
// per-block barrier chain, about equally) is measured in profiles/r02_attention_ws_experiments.txt.
//
// Reference: nn.MultiheadAttention at D/train/train_vid_score/video/clip.py:45 (unfused bmm + softmax + bmm in
// torch 1.11); frames of 197 (ViT-B/16 @ 224, BASELINE configs[1]), 145 and 577 tokens; window attention of
// D/train/train_v106/vsc/baseline/model_factory/backbones/swinv2.py:72-185.
#include <stdlib.h>

#include "host_util.h"
#include "kernels.h"
#include "ptx.cuh"

namespace vscb200 {

#ifdef WS_PROF
#define WS_T(x) const long long x = clock64()
#else
#define WS_T(x)
#endif

// Two shapes of the same kernel: NW = 2 softmax warpgroups with 64-key blocks (11 warps; TMEM per warpgroup S0 S1 O L =
// 64 + 64 + 64 + 16 columns) and NW = 3 warpgroups with 32-key blocks (16 warps; 32 + 32 + 64 + 16 columns each): a third
// warp per scheduler fills the issue slots the other two leave while they wait on TMEM loads and barriers.
constexpr int kWsQTile = 128 * 128;
constexpr int kWsOnes = 2048;                // 16 keys x 128 B of bf16 ones: B operand of the row-sum MMA
constexpr int kWsMaxKeys = 640;
constexpr int kWsBars = 4 + 11 * 3;          // mbarriers of the widest shape

struct WsParams {
  __nv_bfloat16* out;        // [M, C]
  long long* prof;           // WS_PROF builds: per-CTA cycle counters (nullptr otherwise)
  const float* tables;       // Swin form: [heads][(2ws-1)^2]  16*sigmoid(cpb); nullptr for the ViT form
  int T, C;                  // tokens per segment (frame / window), row width of out
  int hgroups, n_segs;       // 64-column head groups per segment (heads, or head pairs for head_dim 32), segments
  int nb, ntiles, ipu;       // key blocks, 128-query tiles per segment, items per unit (= tiles x heads per group)
  int n_units;               // segments * head groups
  int kv_stages, q_bufs;     // 2 / 2 when they fit the shared memory, else 1 / 1
  int reverse;
  float scale_log2e;         // ViT: head_dim^-0.5 * log2 e; Swin: log2 e (q carries the logit scale)
  int ws, shift, nWx, nW_per_frame;
};

__device__ __forceinline__ uint32_t ws_pack(float lo, float hi) {
  return __byte_perm(__float_as_uint(lo) + 0x8000u, __float_as_uint(hi) + 0x8000u, 0x7632);
}

// Cursor over the CTA's flattened item sequence (units blockIdx.x, + gridDim.x, ...; inside a unit the items r =
// tile * HH + head-of-the-group).  The single-thread roles (producer, issuers) advance it incrementally: no division
// per item.  ViT form: unit = frame * heads + head; Swin form: unit = head pair * n_segs + window (a CTA keeps a pair's
// bias tables over consecutive units).
struct WsItem {
  int k;        // CTA-local unit index
  int r;        // item inside the unit
  int seg, hg;
  bool ok;
};
template <bool kSwin>
__device__ __forceinline__ void ws_unit(WsItem& it, const WsParams& p) {
  const int unit = static_cast<int>(blockIdx.x) + it.k * static_cast<int>(gridDim.x);
  it.ok = unit < p.n_units;
  if (it.ok) {
    if (kSwin) {
      it.hg = unit / p.n_segs;
      it.seg = unit - it.hg * p.n_segs;
    } else {
      const int f = unit / p.hgroups;
      it.hg = unit - f * p.hgroups;
      it.seg = p.reverse ? (p.n_segs - 1 - f) : f;
    }
  }
}
template <bool kSwin>
__device__ __forceinline__ void ws_first(WsItem& it, int r0, const WsParams& p) {
  it.k = 0;
  it.r = r0;
  while (it.r >= p.ipu) { it.r -= p.ipu; ++it.k; }
  ws_unit<kSwin>(it, p);
}
template <bool kSwin>
__device__ __forceinline__ void ws_next(WsItem& it, int step, const WsParams& p) {
  it.r += step;
  if (it.r >= p.ipu) {
    do { it.r -= p.ipu; ++it.k; } while (it.r >= p.ipu);
    ws_unit<kSwin>(it, p);
  }
}
// x mod n and x div n for n in {1, 2} (stage / buffer counts)
__device__ __forceinline__ int ws_mod(int x, int n) { return x & (n - 1); }
__device__ __forceinline__ int ws_div(int x, int n) { return x >> (n - 1); }

// Swin form, per item: this row's position in the bias table and its penalties against the four mask regions
struct WsSwinRow {
  const float* tab;          // the head's (2ws-1)^2 table (shared memory)
  int base_i;                // (yi + ws - 1) * (2ws - 1) + xi + ws - 1
  float pen[2][2];           // [key y-region][key x-region]: -100 where it differs from the row's region
  bool masked;               // the window mixes regions (warp-uniform)
};

// One key block of this thread's row: S in TMEM columns [tS, tS + KB) -> P (bf16 pairs over [tS, tS + KB / 2)); running
// shift m (log2 domain); accumulators O [tO, tO + HD) and row sum [tL, tL + 16) rescaled when the shift moves.  kt: first
// key of the block.  Warp-collective.
template <int HD, bool kSwin, int KB>
__device__ __forceinline__ void ws_softmax_block(uint32_t tS, uint32_t tO, uint32_t tL, int kt, bool first, float& m, int T,
                                                 float scale_log2e, uint64_t* pv_done, uint32_t pv_parity, const WsSwinRow& sw,
                                                 int ws, int shift) {
  const int nchunks = min(KB / 32, (T - kt + 31) >> 5);               // chunks with a valid key (warp-uniform, >= 1)
#pragma unroll 1
  for (int c = 0; c < KB / 32; ++c) {
    uint32_t pk[16];
    if (c < nchunks) {
      uint32_t v[32];
      tmem_ld_32x32(tS + c * 32, v);
      tmem_ld_wait();
      const int lim = T - (kt + c * 32);
      if (kSwin) {
        // z = s + bias + mask: table index base_i - (yk * TS + xk); along the chunk the key term grows by 1 per key and
        // by another ws - 1 whenever xk wraps (warp-uniform integer bookkeeping)
        const int TS = 2 * ws - 1, edge = ws - shift;
        int yk = (kt + c * 32) / ws, xk = (kt + c * 32) - yk * ws;
        const float* tp = sw.tab + sw.base_i - (yk * TS + xk);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float zz = __uint_as_float(v[j]) + tp[-j];
          if (sw.masked) {                                            // selects, not indexing: pen stays in registers
            const float pe0 = xk >= edge ? sw.pen[0][1] : sw.pen[0][0], pe1 = xk >= edge ? sw.pen[1][1] : sw.pen[1][0];
            zz += yk >= edge ? pe1 : pe0;
          }
          v[j] = __float_as_uint(zz);
          if (++xk == ws) { xk = 0; ++yk; tp -= ws - 1; }
        }
      }
      if (lim < 32) {                                                 // warp-uniform: the segment's last chunk
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j >= lim) v[j] = 0xFF800000u;                           // -inf: p = 0
      }
      // chunk maximum on registers
      float m0 = fmaxf(fmaxf(__uint_as_float(v[0]), __uint_as_float(v[1])), __uint_as_float(v[2]));
      float m1 = fmaxf(fmaxf(__uint_as_float(v[3]), __uint_as_float(v[4])), __uint_as_float(v[5]));
#pragma unroll
      for (int j = 6; j < 30; j += 4) {
        m0 = fmaxf(fmaxf(m0, __uint_as_float(v[j])), __uint_as_float(v[j + 1]));
        m1 = fmaxf(fmaxf(m1, __uint_as_float(v[j + 2])), __uint_as_float(v[j + 3]));
      }
      const float cmax = fmaxf(fmaxf(m0, m1), fmaxf(__uint_as_float(v[30]), __uint_as_float(v[31]))) * scale_log2e;
      // the shift only moves when the chunk maximum exceeds it by more than 2^8 (p stays below 2^8 otherwise)
      const bool move = cmax > m + 8.0f;                              // m = -inf at the start of an item
      if (__any_sync(0xffffffffu, move)) {
        const float alpha = move ? ex2_approx(m - cmax) : 1.0f;       // 0 for the first chunk of an item
        if (!first) {                                                 // O and the row sum of the blocks before this one
          mbar_wait(pv_done, pv_parity);                              // the previous block's P.V has left them alone
          tc_fence_after();
#pragma unroll 1
          for (int h = 0; h < HD / 16 + 1; ++h) {
            const uint32_t ta = h < HD / 16 ? tO + h * 16 : tL;
            uint32_t o[16];
            tmem_ld_32x16(ta, o);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
            tmem_st_32x16(ta, o);
          }
        }
        for (int cc = 0; cc < c; ++cc) {                              // P chunks of this block already written
          uint32_t q[16];
          tmem_ld_32x16(tS + cc * 16, q);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j)
            q[j] = ws_pack(__uint_as_float(q[j] << 16) * alpha, __uint_as_float(q[j] & 0xFFFF0000u) * alpha);
          tmem_st_32x16(tS + cc * 16, q);
        }
        tmem_st_wait();
        if (move) m = cmax;
      }
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        float q0 = 0.f, q1 = 0.f;
        if (j < lim) {                                                // warp-uniform: no exponentials for the padding keys
          q0 = ex2_approx(fmaf(__uint_as_float(v[j]), scale_log2e, -m));
          q1 = ex2_approx(fmaf(__uint_as_float(v[j + 1]), scale_log2e, -m));
        }
        pk[j >> 1] = __byte_perm(__float_as_uint(q0), __float_as_uint(q1), 0x7632);     // truncation to bf16 pairs
      }
    } else {                                                          // keys past the segment: P = 0 (the MMA reads the whole block)
#pragma unroll
      for (int j = 0; j < 16; ++j) pk[j] = 0u;
    }
    // P of chunk c lands on columns [16c, 16c + 16) of the buffer: inside S chunk 0, consumed
    tmem_st_32x16(tS + c * 16, pk);
  }
  tmem_st_wait();
}

template <int HD, bool kSwin, int NW, int KB, int D>
__global__ void __launch_bounds__((5 * NW + 1) * 32, 1)
attention_ws_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, WsParams p) {
  constexpr int HH = 64 / HD;                             // heads per 64-column group
  constexpr int kWsThreads = (5 * NW + 1) * 32;           // 4 NW softmax warps, NW issuer warps, one producer warp
  constexpr int kWsKB = KB;                               // keys per block
  constexpr int kWsKTile = KB * 128;                      // one K or V block: KB rows x 64 bf16, 128-byte swizzled
  constexpr int kTStride = D * KB + 80;                   // TMEM columns of a warpgroup: D S buffers | O (64) | row sums (16)
  constexpr int kWsOCol = D * KB, kWsLCol = D * KB + 64;
  static_assert(4 + NW * (7 + 2 * D) <= kWsBars, "attention_ws: barrier block");
  constexpr int kIssuer0 = 4 * NW, kProducer = 5 * NW;
#ifdef WS_NO_FUSE_L
  constexpr bool kFuseL = false;
#else
  constexpr bool kFuseL = HD == 64;                       // the row sums ride in the P.V MMA (ViT form)
#endif
  static_assert(NW * kTStride <= 512, "attention_ws: TMEM budget");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const int kv_bytes = p.nb * kWsKTile;                   // K (or V) of one unit
  uint8_t* sKV = smem;                                    // [kv_stages][K | V]
  uint8_t* sQ = sKV + p.kv_stages * 2 * kv_bytes;         // [NW warpgroups][q_bufs] tiles
  uint8_t* sOnes = sQ + NW * p.q_bufs * kWsQTile;          // bf16 ones: the row-sum operand
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + kWsOnes);
  uint64_t* kv_full = bars;                 // [2] per stage
  uint64_t* kv_empty = bars + 2;            // [2]       the issuers with items in the unit have issued their last MMA of it
  uint64_t* q_full = bars + 4;              // [NW][2] per warpgroup and buffer
  uint64_t* q_empty = q_full + 2 * NW;      // [NW][2]
  uint64_t* s_full = q_empty + 2 * NW;      // [NW][D] per warpgroup and S buffer
  uint64_t* p_full = s_full + D * NW;       // [NW][D]
  uint64_t* o_full = p_full + D * NW;       // [NW]
  uint64_t* o_empty = o_full + NW;          // [NW]
  uint64_t* pv_done = o_empty + NW;         // [NW]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kWsBars);
  float* tabs = reinterpret_cast<float*>(bars + kWsBars + 2);      // Swin form: [2 heads][(2ws-1)^2]

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // uniform for the compiler
  if (warp == kIssuer0) {
    if (lane == 0) {
      prefetch_tmap(&tmQ);
      prefetch_tmap(&tmKV);
      for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], min(p.ipu, NW)); }
      for (int i = 0; i < NW; ++i) { mbar_init(&o_full[i], 1); mbar_init(&o_empty[i], 4); mbar_init(&pv_done[i], 1); }
      for (int i = 0; i < 2 * NW; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); }
      for (int i = 0; i < D * NW; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  for (int i = threadIdx.x; i < kWsOnes / 4; i += kWsThreads) reinterpret_cast<uint32_t*>(sOnes)[i] = 0x3F803F80u;
  fence_proxy_async_smem();                                // generic-proxy writes -> visible to the MMA's operand reads
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == kProducer) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int cur_k = -1;
      int n_q[NW] = {};
      WsItem it;
      int w = 0;
      for (ws_first<kSwin>(it, 0, p); it.ok; ws_next<kSwin>(it, 1, p), w = (w + 1 == NW ? 0 : w + 1)) {
        const int row0 = it.seg * p.T;
        if (it.k != cur_k) {                               // new unit: K and V into the next stage
          cur_k = it.k;
          const int st = ws_mod(it.k, p.kv_stages);
          uint8_t* sK = sKV + st * 2 * kv_bytes;
          mbar_wait(&kv_empty[st], (ws_div(it.k, p.kv_stages) & 1) ^ 1);
          mbar_expect_tx(&kv_full[st], 2 * kv_bytes);
          for (int b = 0; b < p.nb; ++b)
            tma_load_2d(sK + b * kWsKTile, &tmKV, &kv_full[st], p.C + it.hg * 64, row0 + b * kWsKB, kEvictFirst);
          for (int b = 0; b < p.nb; ++b)
            tma_load_2d(sK + kv_bytes + b * kWsKTile, &tmKV, &kv_full[st], 2 * p.C + it.hg * 64, row0 + b * kWsKB, kEvictFirst);
        }
        const int qb = ws_mod(n_q[w], p.q_bufs);
        const int use = ws_div(n_q[w], p.q_bufs);
        ++n_q[w];
        mbar_wait(&q_empty[w * 2 + qb], (use & 1) ^ 1);
        mbar_expect_tx(&q_full[w * 2 + qb], kWsQTile);
        tma_load_2d(sQ + (w * p.q_bufs + qb) * kWsQTile, &tmQ, &q_full[w * 2 + qb], it.hg * 64, row0 + (it.r / HH) * 128, kEvictFirst);
      }
    }
  } else if (warp >= kIssuer0 && warp < kProducer) {
    // ------------------------------------------------------------------ MMA issuer of warpgroup w: the WHOLE warp runs the
    // loop on warp-uniform values (uniform registers), one elected lane issues inside the umma_*_warp wrappers
    {
      const int w = warp - kIssuer0;
      constexpr uint32_t idesc_s = make_idesc_bf16_f32(128, kWsKB);
      constexpr uint32_t idesc_o = make_idesc_bf16_f32_bmn(128, HD);
      constexpr uint32_t idesc_l = make_idesc_bf16_f32_bmn(128, 16);
      constexpr uint32_t idesc_ol = make_idesc_bf16_f32_bmn(128, 80);
      const uint32_t ones_u = smem_u32(sOnes);
      const uint64_t ones_d = make_desc_mn_sw128(ones_u);
      const uint32_t tw = tmem_base + w * kTStride;
      const uint32_t sKV_u = smem_u32(sKV), sQ_u = smem_u32(sQ);
      // two cursors over the warpgroup's flattened block sequence: S runs up to D blocks ahead of PV
      int sb = 0, s_item = 0, s_k = -1;                    // next S: block, warpgroup-local item index, unit seen
      int pb = 0, p_item = 0;
      uint32_t sg = 0, pg = 0;                             // global block counters
      WsItem si, pi;
      ws_first<kSwin>(si, w, p);
      ws_first<kSwin>(pi, w, p);
#ifdef WS_PROF
      long long i_kv = 0, i_q = 0, i_p = 0, i_oe = 0;
      const long long i_begin = clock64();
#endif
      auto issue_s = [&]() {
        WS_T(u0);
        if (si.k != s_k) {                                 // first S of a unit: K / V landed
          s_k = si.k;
          mbar_wait(&kv_full[ws_mod(si.k, p.kv_stages)], ws_div(si.k, p.kv_stages) & 1);
        }
        WS_T(u1);
        const int qb = ws_mod(s_item, p.q_bufs);
        if (sb == 0) mbar_wait(&q_full[w * 2 + qb], ws_div(s_item, p.q_bufs) & 1);
#ifdef WS_PROF
        i_kv += u1 - u0; i_q += clock64() - u1;
#endif
        tc_fence_after();
        const int s_hh = (si.r % HH) * 64;                 // byte offset of the head inside the 64-column box
        const uint64_t qd = make_desc_k_sw128(sQ_u + (w * p.q_bufs + qb) * kWsQTile + s_hh);
        const uint64_t kd = make_desc_k_sw128(sKV_u + ws_mod(si.k, p.kv_stages) * 2 * kv_bytes + sb * kWsKTile + s_hh);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16_ss_warp(tw + (sg % D) * kWsKB, qd + 2 * k, kd + 2 * k, idesc_s, k ? 1u : 0u);
        umma_commit_warp(&s_full[w * D + (sg % D)]);
        ++sg;
        if (++sb == p.nb) {                                // last S of the item: its Q buffer is free once these retire
          umma_commit_warp(&q_empty[w * 2 + qb]);
          sb = 0;
          ++s_item;
          ws_next<kSwin>(si, NW, p);
        }
      };
      // S runs up to D blocks ahead of PV, but never into a unit whose K / V stage is still held by a unit this
      // warpgroup has not finished (the producer refills a stage only after BOTH issuers released it)
      auto fill_s = [&]() {
        while (si.ok && sg - pg < static_cast<uint32_t>(D) && si.k < pi.k + p.kv_stages) issue_s();
      };
      while (pi.ok) {
        fill_s();
        WS_T(u2);
        mbar_wait(&p_full[w * D + (pg % D)], (pg / D) & 1);
        WS_T(u3);
        if (pb == 0 && p_item > 0) mbar_wait(&o_empty[w], (p_item - 1) & 1);   // the warpgroup has read the previous item's O
#ifdef WS_PROF
        i_p += u3 - u2; i_oe += clock64() - u3;
#endif
        tc_fence_after();
        const uint32_t v_addr = sKV_u + ws_mod(pi.k, p.kv_stages) * 2 * kv_bytes + kv_bytes + pb * kWsKTile + (pi.r % HH) * 64;
        if constexpr (kFuseL) {
          // O and the row sums in ONE MMA per 16 keys: N = 80, whose second 64-column block (LBO) is the tile of ones
#pragma unroll
          for (int i = 0; i < kWsKB / 16; ++i) {
            const uint32_t va = v_addr + i * 2048;
            umma_bf16_ts_warp(tw + kWsOCol, tw + (pg % D) * kWsKB + i * 8, make_desc_mn_sw128_lbo(va, ones_u - va), idesc_ol,
                              (pb | i) ? 1u : 0u);
          }
        } else {
          const uint64_t vd = make_desc_mn_sw128(v_addr);
#pragma unroll
          for (int i = 0; i < kWsKB / 16; ++i) {
            umma_bf16_ts_warp(tw + kWsOCol, tw + (pg % D) * kWsKB + i * 8, vd + static_cast<uint64_t>(i) * 128, idesc_o, (pb | i) ? 1u : 0u);
            umma_bf16_ts_warp(tw + kWsLCol, tw + (pg % D) * kWsKB + i * 8, ones_d, idesc_l, (pb | i) ? 1u : 0u);      // row sums
          }
        }
        umma_commit_warp(&pv_done[w]);
        ++pg;
        const int prev_k = pi.k;
        if (++pb == p.nb) {
          umma_commit_warp(&o_full[w]);
          pb = 0;
          ++p_item;
          ws_next<kSwin>(pi, NW, p);
          if (!pi.ok || pi.k != prev_k) umma_commit_warp(&kv_empty[ws_mod(prev_k, p.kv_stages)]);   // this warpgroup is done with the unit
        }
      }
#ifdef WS_PROF
      if (p.prof && lane == 0 && w < 2) {
        long long* o = p.prof + (static_cast<long long>(blockIdx.x) * 8 + 3 + 4 * w) * 5;   // slots of the mostly idle warps 3 / 7
        o[0] = i_kv; o[1] = i_q; o[2] = i_p; o[3] = i_oe; o[4] = clock64() - i_begin;
      }
#endif
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    const int w = warp >> 2, quad = warp & 3;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + w * kTStride;
    uint32_t g = 0;
    int n_item = 0;
    // deferred epilogue of the previous item
    bool pend = false, pend_valid = false;
    __nv_bfloat16* pend_row = nullptr;
    auto epilogue = [&]() {                                // O of the previous item -> registers, accumulator released
      mbar_wait(&o_full[w], (n_item - 1) & 1);
      tc_fence_after();
      uint32_t o[HD], ls[16];
#pragma unroll
      for (int h = 0; h < HD / 32; ++h) tmem_ld_32x32(tlane + kWsOCol + h * 32, reinterpret_cast<uint32_t(&)[32]>(o[h * 32]));
      tmem_ld_32x16(tlane + kWsLCol, ls);
      tmem_ld_wait();
      const float pend_inv_l = 1.0f / __uint_as_float(ls[0]);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[w]);
      if (pend_valid) {
#pragma unroll
        for (int q = 0; q < HD / 8; ++q) {
          uint4 u;
          // cvt.rn.bf16x2 (one F2FP per pair): the XU pipe it issues on is idle during the epilogue
          u.x = pack_bf16x2(__uint_as_float(o[8 * q]) * pend_inv_l, __uint_as_float(o[8 * q + 1]) * pend_inv_l);
          u.y = pack_bf16x2(__uint_as_float(o[8 * q + 2]) * pend_inv_l, __uint_as_float(o[8 * q + 3]) * pend_inv_l);
          u.z = pack_bf16x2(__uint_as_float(o[8 * q + 4]) * pend_inv_l, __uint_as_float(o[8 * q + 5]) * pend_inv_l);
          u.w = pack_bf16x2(__uint_as_float(o[8 * q + 6]) * pend_inv_l, __uint_as_float(o[8 * q + 7]) * pend_inv_l);
          *reinterpret_cast<uint4*>(pend_row + q * 8) = u;
        }
      }
      pend = false;
    };
    WsItem it;
#ifdef WS_PROF
    long long c_wait = 0, c_soft = 0, c_epi = 0, c_arr = 0;
    const long long c_begin = clock64();
#endif
    int loaded_hg = -1;
    const int TS = 2 * p.ws - 1;
    for (ws_first<kSwin>(it, w, p); it.ok; ws_next<kSwin>(it, NW, p)) {
      const int tile = it.r / HH, hh = it.r - tile * HH;
      const int i_tok = tile * 128 + quad * 32 + lane;
      const bool warp_valid = tile * 128 + quad * 32 < p.T;
      WsSwinRow sw;
      sw.tab = tabs + hh * TS * TS;
      sw.base_i = 0;
      sw.masked = false;
      sw.pen[0][0] = sw.pen[0][1] = sw.pen[1][0] = sw.pen[1][1] = 0.f;
      if (kSwin) {
        if (it.hg != loaded_hg) {                          // bias tables of the head pair -> shared memory.  Both warpgroups
          named_bar_sync(1, 128 * NW);                     // walk the units in the same order: whoever gets here first waits
          for (int e = threadIdx.x; e < 2 * TS * TS; e += 128 * NW) tabs[e] = p.tables[static_cast<int64_t>(it.hg) * 2 * TS * TS + e];
          named_bar_sync(1, 128 * NW);                     // until the others have finished the previous pair's items
          loaded_hg = it.hg;
        }
        bool edge_y = false, edge_x = false;
        if (p.shift > 0) {                                 // shifted-window regions: only the last window row / column of a frame
          const int wf = it.seg % p.nW_per_frame;
          edge_y = (wf / p.nWx) == (p.nW_per_frame / p.nWx) - 1;
          edge_x = (wf % p.nWx) == p.nWx - 1;
        }
        sw.masked = edge_y || edge_x;
        const int tok = i_tok < p.T ? i_tok : 0;
        const int yi = tok / p.ws, xi = tok - yi * p.ws;
        sw.base_i = (yi + p.ws - 1) * TS + xi + p.ws - 1;
        const int ry_i = (edge_y && yi >= p.ws - p.shift) ? 1 : 0, rx_i = (edge_x && xi >= p.ws - p.shift) ? 1 : 0;
#pragma unroll
        for (int a2 = 0; a2 < 2; ++a2)
#pragma unroll
          for (int b2 = 0; b2 < 2; ++b2) sw.pen[a2][b2] = ((edge_y && a2 != ry_i) || (edge_x && b2 != rx_i)) ? -100.0f : 0.0f;
      }
      float m = -INFINITY;
#pragma unroll 1
      for (int b = 0; b < p.nb; ++b) {
        WS_T(t0);
        mbar_wait(&s_full[w * D + (g % D)], (g / D) & 1);
        tc_fence_after();
        WS_T(t1);
        // warps whose 32 rows all lie past the frame skip the block: the MMA reads stale TMEM for them, and whatever it
        // computes stays in rows nobody stores
        if (warp_valid)
          ws_softmax_block<HD, kSwin, KB>(tlane + (g % D) * kWsKB, tlane + kWsOCol, tlane + kWsLCol, b * kWsKB, b == 0, m, p.T,
                                      p.scale_log2e, &pv_done[w], (g - 1) & 1, sw, p.ws, p.shift);
        WS_T(t2);
        if (b == 0 && pend) epilogue();                    // before P.V of this item's first block may overwrite O
        WS_T(t3);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[w * D + (g % D)]);
        ++g;
#ifdef WS_PROF
        const long long t4 = clock64();
        c_wait += t1 - t0; c_soft += t2 - t1; c_epi += t3 - t2; c_arr += t4 - t3;
#endif
      }
      ++n_item;
      pend = true;
      pend_valid = i_tok < p.T;
      pend_row = p.out + (static_cast<int64_t>(it.seg) * p.T + (pend_valid ? i_tok : 0)) * p.C + it.hg * 64 + hh * HD;
    }
    if (pend) epilogue();
#ifdef WS_PROF
    if (p.prof && lane == 0 && quad != 3 && warp < 8) {
      long long* o = p.prof + (static_cast<long long>(blockIdx.x) * 8 + warp) * 5;
      o[0] = c_wait; o[1] = c_soft; o[2] = c_epi; o[3] = c_arr; o[4] = clock64() - c_begin;
    }
#endif
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kIssuer0) tmem_dealloc<512>(tmem_base);
}

bool attention_ws_supported(int T, int head_dim) {
  static const int off = [] { const char* e = getenv("VSCB200_ATTN_NO_WS"); return e ? atoi(e) : 0; }();
  return !off && ((head_dim == 64 && T > 128) || (head_dim == 32 && T >= 64)) && T <= kWsMaxKeys;
}

// The two forms share one launcher.  qkv: [n_segs * T, 3C] bf16 (q | k | v, heads head_dim-wide contiguous); out: [n_segs * T, C].
template <int HD, bool kSwin, int NW, int KB, int D>
static int attention_ws_launch_shape(const void* qkv, void* out, int64_t n_segs, int T, int heads, float scale, const float* tables,
                                     int ws, int shift, int nWx, int nW_per_frame, cudaStream_t stream, bool reverse) {
  constexpr int kKTile = KB * 128;
  constexpr int kThreads = (5 * NW + 1) * 32;
  const int C = heads * HD;
  const int64_t M = n_segs * T;
  const int hgroups = C / 64;
  CUtensorMap tmQ, tmKV;
  int rc = make_tmap_2d(&tmQ, qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, 3 * C, 3 * C, 128, 64, true);
  if (rc) return rc;
  if ((rc = make_tmap_2d(&tmKV, qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, 3 * C, 3 * C, KB, 64, true))) return rc;
  WsParams p;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.tables = tables;
  p.T = T; p.C = C; p.hgroups = hgroups; p.n_segs = static_cast<int>(n_segs);
  p.nb = (T + KB - 1) / KB;
  p.ntiles = (T + 127) / 128;
  p.ipu = p.ntiles * (64 / HD);
  p.n_units = static_cast<int>(n_segs) * hgroups;
  p.reverse = reverse ? 1 : 0;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.ws = kSwin ? ws : 1; p.shift = shift; p.nWx = nWx > 0 ? nWx : 1; p.nW_per_frame = nW_per_frame > 0 ? nW_per_frame : 1;
  const int kv_bytes = p.nb * kKTile;
  const int TS = 2 * p.ws - 1;
  const int fixed = kWsOnes + (kWsBars + 2) * 8 + (kSwin ? 2 * TS * TS * 4 : 0) + 64 + 1024;
  p.kv_stages = 2; p.q_bufs = 2;
  if (2 * 2 * kv_bytes + NW * 2 * kWsQTile + fixed > 232448) p.q_bufs = 1;
  if (2 * 2 * kv_bytes + NW * p.q_bufs * kWsQTile + fixed > 232448) { p.kv_stages = 1; p.q_bufs = 2; }
  if (p.kv_stages * 2 * kv_bytes + NW * p.q_bufs * kWsQTile + fixed > 232448) p.q_bufs = 1;
  const int smem = p.kv_stages * 2 * kv_bytes + NW * p.q_bufs * kWsQTile + fixed;
  VSCB_REQUIRE(smem <= 232448, "attention_ws: shared memory");
  const int grid = p.n_units < device_sm_count() ? p.n_units : device_sm_count();
  ProfScope prof(kProfAttention, stream, 4.0 * static_cast<double>(M) * T * C);
  p.prof = nullptr;
#ifdef WS_PROF
  static long long* prof_buf = nullptr;
  if (!prof_buf) cudaMalloc(&prof_buf, 148 * 8 * 5 * sizeof(long long));
  p.prof = prof_buf;
#endif
  VSCB_CUDA_OK(cudaFuncSetAttribute(attention_ws_kernel<HD, kSwin, NW, KB, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  attention_ws_kernel<HD, kSwin, NW, KB, D><<<grid, kThreads, smem, stream>>>(tmQ, tmKV, p);
#ifdef WS_PROF
  {
    static long long h[148 * 8 * 5];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, prof_buf, sizeof(h), cudaMemcpyDeviceToHost);
    static int calls = 0;
    if (++calls == 5)
      for (int cta : {0, 77})
        for (int wp = 0; wp < 8; ++wp) {
          const long long* o = h + (cta * 8 + wp) * 5;
          printf((wp & 3) == 3 ? "WS_PROF cta %d issuer %d: wait kv_full %lld q_full %lld p_full %lld o_empty %lld total %lld\n" : "WS_PROF cta %d warp %d: wait_s %lld softmax %lld epilogue %lld arrive %lld total %lld\n", cta, wp, o[0], o[1], o[2], o[3], o[4]);
        }
  }
#endif
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

// VSCB200_ATTN_WS_SHAPE (A/B switch, read once): 2 = two warpgroups / 64-key blocks, 3 = three warpgroups / 32-key blocks,
// 4..6 = two warpgroups / 32-key blocks with 4, 3, 2 S buffers.  Default: 3 for the ViT form (99 vs 103.5 us at 256 x 12 x 197,
// 220 vs 223 us at 64 x 16 x 577 -- a third warp per scheduler; deeper S buffering measured no gain, profiles/README.md),
// 2 for the Swin form.
static int attention_ws_shape() {
  static const int shape = [] { const char* e = getenv("VSCB200_ATTN_WS_SHAPE"); return e ? atoi(e) : 0; }();
  return shape;
}

static int attention_ws_launch(const void* qkv, void* out, int64_t n_segs, int T, int heads, int head_dim, float scale,
                               const float* tables, int ws, int shift, int nWx, int nW_per_frame, cudaStream_t stream, bool reverse) {
  VSCB_REQUIRE(attention_ws_supported(T, head_dim), "attention_ws: unsupported segment length / head_dim");
  const bool swin = head_dim == 32;
  VSCB_REQUIRE(swin == (tables != nullptr), "attention_ws: head_dim 32 is the Swin-V2 form (bias tables), 64 the ViT form");
  VSCB_REQUIRE(!swin || (heads % 2 == 0 && ws * ws == T), "attention_ws: Swin form needs an even head count and T = ws * ws");
  const int C = heads * head_dim;
  VSCB_REQUIRE(n_segs * T < (1ll << 31) && n_segs * (C / 64) < (1ll << 30), "attention_ws: problem too large");
  const int shape = attention_ws_shape();
  if (swin) {
    if (shape == 3) return attention_ws_launch_shape<32, true, 3, 32, 2>(qkv, out, n_segs, T, heads, scale, tables, ws, shift, nWx, nW_per_frame, stream, reverse);
    return attention_ws_launch_shape<32, true, 2, 64, 2>(qkv, out, n_segs, T, heads, scale, tables, ws, shift, nWx, nW_per_frame, stream, reverse);
  }
#define WS_ARGS qkv, out, n_segs, T, heads, scale, tables, ws, shift, nWx, nW_per_frame, stream, reverse
  if (shape == 0 || shape == 3) return attention_ws_launch_shape<64, false, 3, 32, 2>(WS_ARGS);
  if (shape == 4) return attention_ws_launch_shape<64, false, 2, 32, 4>(WS_ARGS);
  if (shape == 5) return attention_ws_launch_shape<64, false, 2, 32, 3>(WS_ARGS);
  if (shape == 6) return attention_ws_launch_shape<64, false, 2, 32, 2>(WS_ARGS);
  return attention_ws_launch_shape<64, false, 2, 64, 2>(qkv, out, n_segs, T, heads, scale, tables, ws, shift, nWx, nW_per_frame, stream, reverse);
}

// ViT form: softmax(q.k / sqrt(64)) v per (frame, head)
int attention_ws(const void* qkv, void* out, int n_frames, int T, int heads, cudaStream_t stream, bool reverse) {
  return attention_ws_launch(qkv, out, n_frames, T, heads, 64, 1.0f / sqrtf(64.0f), nullptr, 0, 0, 0, 0, stream, reverse);
}
// Swin-V2 form (head_dim 32): q, k arrive cosine-normalised with the logit scale in q; tables [heads][(2ws-1)^2]; the rows
// of a segment are the tokens of one ws x ws window in row-major order (swin_attention.cu's conventions)
int attention_ws_swin(const void* qkv, void* out, int64_t n_windows_total, int ws, int heads, const float* tables, int shift,
                      int nWx, int nW_per_frame, cudaStream_t stream) {
  return attention_ws_launch(qkv, out, n_windows_total, ws * ws, heads, 32, 1.0f, tables, ws, shift, nWx, nW_per_frame, stream, false);
}

}  // namespace vscb200
