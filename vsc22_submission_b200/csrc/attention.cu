// Fused multi-head self-attention for the ViT encoder: out = softmax(Q K^T / sqrt(d)) V per (frame, head).
//
// Reference: nn.MultiheadAttention called at D/train/train_vid_score/video/clip.py:45 (unfused
// bmm + softmax + bmm in torch 1.11).  Here: one CTA per (head, frame) whose warps walk the
// 16-row query tiles; K and V of the (frame, head) are staged ONCE (cp.async) in XOR-swizzled shared
// memory, S = QK^T and O = PV run
// on tensor cores (mma.sync m16n8k16 bf16, fp32 accumulate) with an online softmax in registers, so
// neither S nor P ever touches HBM.  head_dim is fixed at 64 (every ViT on the reference's path:
// 768/12, 1024/16).  Attention is 4 % of the encoder FLOPs; the tcgen05 projections carry the rest.
#include "host_util.h"
#include "kernels.h"
#include "ptx.cuh"

namespace vscb200 {

constexpr int kAttThreads = 224;   // 7 warps; T = 197 -> 13 query tiles of 16 rows, two rounds (14 slots)

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void cp_async_16(uint32_t smem_addr, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// qkv: [n*T, 3W] bf16 with W = heads*64, columns [q | k | v]; out: [n*T, W] bf16.
__global__ void __launch_bounds__(kAttThreads, 2)
attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int T, int Tpad, int heads,
                 float scale_log2e) {
  extern __shared__ __align__(16) uint8_t att_smem[];
  uint8_t* sK = att_smem;                       // [Tpad][128 B], 16B chunk c of row r at c ^ (r & 7)
  uint8_t* sV = att_smem + static_cast<size_t>(Tpad) * 128;

  const int frame = blockIdx.y, head = blockIdx.x;
  const int W = heads * 64;
  const int64_t ld = 3 * static_cast<int64_t>(W);
  const __nv_bfloat16* base = qkv + static_cast<int64_t>(frame) * T * ld + head * 64;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- stage K and V (zero-filled beyond T), asynchronously
  const uint32_t sK_u = smem_u32(sK), sV_u = smem_u32(sV);
  for (int i = tid; i < Tpad * 8; i += kAttThreads) {
    const int r = i >> 3, c = i & 7;
    const int off = r * 128 + ((c ^ (r & 7)) << 4);
    if (r < T) {
      const __nv_bfloat16* src = base + r * ld + c * 8;
      cp_async_16(sK_u + off, src + W);
      cp_async_16(sV_u + off, src + 2 * W);
    } else {
      *reinterpret_cast<uint4*>(sK + off) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(sV + off) = make_uint4(0, 0, 0, 0);
    }
  }
  cp_async_wait_all();
  __syncthreads();

  const int g = lane >> 2, t = lane & 3;
  const int lm = lane >> 3, lr = lane & 7;       // ldmatrix lane -> (matrix id, row in matrix)
  const int ntile_q = (T + 15) >> 4;
  for (int qt = warp; qt < ntile_q; qt += kAttThreads / 32) {
  const int q0 = qt * 16;
  // ---- Q fragments for this warp's 16 rows (A operand, 4 k-steps over d = 64)
  uint32_t qa[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const int row = q0 + g + ((h & 1) ? 8 : 0);
      const int d = ks * 16 + 2 * t + ((h & 2) ? 8 : 0);
      qa[ks][h] = (row < T) ? *reinterpret_cast<const uint32_t*>(base + row * ld + d) : 0u;
    }
  }

  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};

  {
    for (int k0 = 0; k0 < Tpad; k0 += 64) {
      const int ntiles = min(8, (Tpad - k0) >> 3);   // 8-key n-tiles in this key block (even)
      float s[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      // S = Q K^T
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < ntiles) {
          const int key = k0 + j * 8 + lr;
#pragma unroll
          for (int kp = 0; kp < 2; ++kp) {       // two k-steps per ldmatrix.x4
            uint32_t b[4];
            const int chunk = kp * 4 + lm;       // d chunk (8 elements) 0..7
            ldmatrix_x4(b, sK_u + key * 128 + ((chunk ^ (key & 7)) << 4));
            mma_bf16_16816(s[j], qa[kp * 2], b[0], b[1]);
            mma_bf16_16816(s[j], qa[kp * 2 + 1], b[2], b[3]);
          }
        }
      }
      // mask padded keys, online softmax (rows g and g+8; a row is spread over the 4 lanes of a quad)
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < ntiles) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int key = k0 + j * 8 + 2 * t + (e & 1);
            if (key >= T) s[j][e] = -INFINITY;
            mx[e >> 1] = fmaxf(mx[e >> 1], s[j][e]);
          }
        }
      }
      float corr[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
        const float m_new = fmaxf(m_run[r], mx[r]);       // finite: every key block holds >= 1 valid key
        corr[r] = exp2f((m_run[r] - m_new) * scale_log2e);
        m_run[r] = m_new;
        l_run[r] *= corr[r];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o[j][0] *= corr[0]; o[j][1] *= corr[0]; o[j][2] *= corr[1]; o[j][3] *= corr[1];
      }
      uint32_t pa[4][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float p[4] = {0.f, 0.f, 0.f, 0.f};
        if (j < ntiles) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            p[e] = exp2f((s[j][e] - m_run[e >> 1]) * scale_log2e);
            l_run[e >> 1] += p[e];
          }
        }
        // C fragments of n-tiles (2i, 2i+1) form the A fragment of k-step i
        pa[j >> 1][(j & 1) * 2 + 0] = pack_bf16x2(p[0], p[1]);
        pa[j >> 1][(j & 1) * 2 + 1] = pack_bf16x2(p[2], p[3]);
      }
      // O += P V
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        if (ks * 2 < ntiles) {
          const int key = k0 + ks * 16 + (lm & 1) * 8 + lr;
#pragma unroll
          for (int dp = 0; dp < 4; ++dp) {       // two 8-wide d tiles per ldmatrix.x4.trans
            uint32_t b[4];
            const int chunk = dp * 2 + (lm >> 1);
            ldmatrix_x4_trans(b, sV_u + key * 128 + ((chunk ^ (key & 7)) << 4));
            mma_bf16_16816(o[dp * 2], pa[ks], b[0], b[1]);
            mma_bf16_16816(o[dp * 2 + 1], pa[ks], b[2], b[3]);
          }
        }
      }
    }
    // ---- normalise and store
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
      l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    __nv_bfloat16* obase = out + static_cast<int64_t>(frame) * T * W + head * 64;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int row = q0 + g + r * 8;
      if (row < T) {
        const float inv = 1.0f / l_run[r];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          *reinterpret_cast<uint32_t*>(obase + static_cast<int64_t>(row) * W + j * 8 + 2 * t) =
              pack_bf16x2(o[j][r * 2] * inv, o[j][r * 2 + 1] * inv);
        }
      }
    }
  }
  }   // query-tile loop
}

bool attention_tc_supported(int T, int head_dim);
int attention_tc(const void* qkv, void* out, int n_frames, int T, int heads, cudaStream_t stream, bool reverse);

int attention(const void* qkv, void* out, int n_frames, int T, int heads, int head_dim, cudaStream_t stream,
              bool reverse) {
  VSCB_REQUIRE(head_dim == 64, "attention: head_dim must be 64");
  VSCB_REQUIRE(n_frames > 0 && T > 0 && heads > 0, "attention: empty problem");
  // 128 < T <= 640: streaming tcgen05 kernel (attention_ws.cu): decoupled warpgroups, keys in blocks of 32, S double-buffered
  if (attention_ws_supported(T, head_dim) && T != 257 && static_cast<int64_t>(n_frames) * T < (1ll << 31))   // 257 = 256 + 1: attention_tc.cu folds the odd token in
    return attention_ws(qkv, out, n_frames, T, heads, stream, reverse);
  // T <= 257: tcgen05 kernel with S resident in TMEM (attention_tc.cu)
  if (attention_tc_supported(T, head_dim) && static_cast<int64_t>(n_frames) * T < (1ll << 31))
    return attention_tc(qkv, out, n_frames, T, heads, stream, reverse);
  // 257 < T <= 640 (ViT-L/16 @ 384: T = 577): K-blocked tcgen05 kernel with an online softmax (attention_kb.cu)
  if (attention_kb_supported(T, head_dim) && static_cast<int64_t>(n_frames) * T < (1ll << 31))
    return attention_kb(qkv, out, n_frames, T, heads, head_dim, 1.0f / sqrtf(static_cast<float>(head_dim)), nullptr, 0, 0, 0, 0,
                        stream);
  const int Tpad = (T + 15) & ~15;
  const size_t smem = static_cast<size_t>(Tpad) * 128 * 2;
  VSCB_REQUIRE(smem <= 200 * 1024, "attention: sequence too long for the single-pass K/V staging");
  VSCB_CUDA_OK(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  dim3 grid(heads, n_frames);
  const float scale_log2e = (1.0f / sqrtf(static_cast<float>(head_dim))) * 1.4426950408889634f;
  ProfScope prof(kProfAttention, stream, 4.0 * n_frames * heads * static_cast<double>(T) * T * head_dim);
  attention_kernel<<<grid, kAttThreads, smem, stream>>>(reinterpret_cast<const __nv_bfloat16*>(qkv),
                                                        reinterpret_cast<__nv_bfloat16*>(out), T, Tpad, heads,
                                                        scale_log2e);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

}  // namespace vscb200

extern "C" int vscb200_attention(const void* qkv_bf16, void* out_bf16, int n_frames, int T, int heads, int head_dim,
                                 void* stream) {
  return vscb200::attention(qkv_bf16, out_bf16, n_frames, T, heads, head_dim, static_cast<cudaStream_t>(stream), false);
}
