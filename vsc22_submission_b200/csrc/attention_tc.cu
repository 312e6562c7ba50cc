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
  int T, Tpad, heads, W, mtiles, reverse;
  float scale_log2e;
};

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
    const int nfull = p.Tpad >> 5;
    const bool tail16 = (p.Tpad & 31) != 0;
    float inv_l = 0.f;
    mbar_wait(bar_s, 0);
    tc_fence_after();
    if (warp_valid) {
      // ---- pass 1: row maximum over the valid keys
      float mx = -INFINITY;
      for (int c = 0; c < nfull; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tlane + c * 32, v);
        tmem_ld_wait();
        const int lim = p.T - c * 32;
        if (lim >= 32) {                                 // warp-uniform fast path: no masking
#pragma unroll
          for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < lim) mx = fmaxf(mx, __uint_as_float(v[j]));
        }
      }
      if (tail16) {
        uint32_t v[16];
        tmem_ld_32x16(tlane + nfull * 32, v);
        tmem_ld_wait();
        const int lim = p.T - nfull * 32;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j < lim) mx = fmaxf(mx, __uint_as_float(v[j]));
      }
      const float mxs = mx * p.scale_log2e;
      // ---- pass 2: p = 2^(s*scale - max*scale), row sum, P -> TMEM (bf16 pairs, in place over S)
      float l = 0.f;
      for (int c = 0; c < nfull; ++c) {
        uint32_t v[32], pk[16];
        tmem_ld_32x32(tlane + c * 32, v);
        tmem_ld_wait();
        const int lim = p.T - c * 32;
        if (lim >= 32) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float p0 = ex2_approx(fmaf(__uint_as_float(v[j]), p.scale_log2e, -mxs));
            const float p1 = ex2_approx(fmaf(__uint_as_float(v[j + 1]), p.scale_log2e, -mxs));
            l += p0 + p1;
            pk[j >> 1] = pack_bf16x2(p0, p1);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float p0 = ex2_approx(fmaf(__uint_as_float(v[j]), p.scale_log2e, -mxs));
            float p1 = ex2_approx(fmaf(__uint_as_float(v[j + 1]), p.scale_log2e, -mxs));
            if (j >= lim) p0 = 0.f;
            if (j + 1 >= lim) p1 = 0.f;
            l += p0 + p1;
            pk[j >> 1] = pack_bf16x2(p0, p1);
          }
        }
        tmem_st_32x16(tlane + c * 16, pk);
      }
      if (tail16) {
        uint32_t v[16], pk[8];
        tmem_ld_32x16(tlane + nfull * 32, v);
        tmem_ld_wait();
        const int lim = p.T - nfull * 32;
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          float p0 = ex2_approx(fmaf(__uint_as_float(v[j]), p.scale_log2e, -mxs));
          float p1 = ex2_approx(fmaf(__uint_as_float(v[j + 1]), p.scale_log2e, -mxs));
          if (j >= lim) p0 = 0.f;
          if (j + 1 >= lim) p1 = 0.f;
          l += p0 + p1;
          pk[j >> 1] = pack_bf16x2(p0, p1);
        }
        tmem_st_32x8(tlane + nfull * 16, pk);
      }
      tmem_st_wait();
      inv_l = 1.0f / l;
    }
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
        if (row < p.T) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(v[8 * q]) * inv_l, __uint_as_float(v[8 * q + 1]) * inv_l);
            o.y = pack_bf16x2(__uint_as_float(v[8 * q + 2]) * inv_l, __uint_as_float(v[8 * q + 3]) * inv_l);
            o.z = pack_bf16x2(__uint_as_float(v[8 * q + 4]) * inv_l, __uint_as_float(v[8 * q + 5]) * inv_l);
            o.w = pack_bf16x2(__uint_as_float(v[8 * q + 6]) * inv_l, __uint_as_float(v[8 * q + 7]) * inv_l);
            *reinterpret_cast<uint4*>(orow + hc * 32 + q * 8) = o;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<kAtcTmemCols>(tmem_base);
}

bool attention_tc_supported(int T, int head_dim) {
  static const int off = [] { const char* e = getenv("VSCB200_ATTN_MMA_SYNC"); return e ? atoi(e) : 0; }();
  return !off && head_dim == 64 && ((T + 15) & ~15) <= 256;
}

int attention_tc(const void* qkv, void* out, int n_frames, int T, int heads, cudaStream_t stream, bool reverse) {
  const int Tpad = (T + 15) & ~15;
  const int W = heads * 64;
  const int64_t M = static_cast<int64_t>(n_frames) * T;
  CUtensorMap tmQ, tmKV;
  int rc = make_tmap_2d(&tmQ, qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, 3 * W, 3 * W, 128, 64, true);
  if (rc) return rc;
  if ((rc = make_tmap_2d(&tmKV, qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, 3 * W, 3 * W, Tpad, 64, true))) return rc;
  AtcParams p;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.T = T; p.Tpad = Tpad; p.heads = heads; p.W = W; p.mtiles = (T + 127) / 128; p.reverse = reverse ? 1 : 0;
  p.scale_log2e = (1.0f / sqrtf(64.0f)) * 1.4426950408889634f;
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
