// Multi-GPU similarity exchange: every rank searches its bank shard with GLOBAL row ids; the partial
// [nq, k] results are packed into ONE 64-bit key per entry, exchanged with a single all-gather (NCCL, or peer copies
// inside one process -- faiss_compat.index_cpu_to_all_gpus) and merged k-way on the device.
//
//   key = (order-preserving score bits << 32) | ~id      larger key = better; equal scores -> the lower id wins
//         (faiss result order, vsc/index.py:174); 0 = padding (id -1).  Ids must be < 2^32.
//
// Reference: the reference replicates the bank on every GPU instead (vsc/exhaustive_search.py:229-234, co.shard = False).
#include <float.h>

#include "host_util.h"
#include "kernels.h"

namespace vscb200 {

__device__ __forceinline__ uint32_t mk_okey(float f, bool keep_max) {
  const uint32_t u = __float_as_uint(f);
  const uint32_t k = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return keep_max ? k : ~k;
}
__device__ __forceinline__ float mk_okey_inv(uint32_t key, bool keep_max) {
  const uint32_t k = keep_max ? key : ~key;
  return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}

__global__ void topk_pack_kernel(const float* __restrict__ D, const int64_t* __restrict__ I, int64_t n, int keep_max,
                                 unsigned long long* __restrict__ keys) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t id = I[i];
  keys[i] = id < 0 ? 0ull
                   : (static_cast<unsigned long long>(mk_okey(D[i], keep_max != 0)) << 32) |
                         static_cast<uint32_t>(~static_cast<uint32_t>(id));
}

// the same into columns [col0, col0 + k) of rows of ld keys (several partial results side by side: no copy afterwards)
__global__ void topk_pack_cols_kernel(const float* __restrict__ D, const int64_t* __restrict__ I, int64_t nq, int k, int keep_max,
                                      unsigned long long* __restrict__ keys, int ld, int col0) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= nq * k) return;
  const int64_t row = i / k;
  const int c = static_cast<int>(i - row * k);
  const int64_t id = I[i];
  keys[row * ld + col0 + c] = id < 0 ? 0ull
                                     : (static_cast<unsigned long long>(mk_okey(D[i], keep_max != 0)) << 32) |
                                           static_cast<uint32_t>(~static_cast<uint32_t>(id));
}

__device__ __forceinline__ unsigned long long mg_warp_max(unsigned long long v) {
  const uint32_t hi = static_cast<uint32_t>(v >> 32);
  const uint32_t mh = __reduce_max_sync(0xffffffffu, hi);
  const uint32_t lo = hi == mh ? static_cast<uint32_t>(v) : 0u;
  const uint32_t ml = __reduce_max_sync(0xffffffffu, lo);
  return (static_cast<unsigned long long>(mh) << 32) | ml;
}

__device__ __forceinline__ void mg_emit(unsigned long long key, int keep_max, float* D, int64_t* I, float bias = 0.f) {
  if (key != 0ull) {
    *D = mk_okey_inv(static_cast<uint32_t>(key >> 32), keep_max != 0) + bias;
    *I = static_cast<int64_t>(~static_cast<uint32_t>(key & 0xFFFFFFFFull));
  } else {
    *D = keep_max ? -FLT_MAX : FLT_MAX;
    *I = -1;
  }
}

// keys: [parts][nq][ld] (one block per rank, as the all-gather lays them out); the merged result uses columns [col0,
// col0 + kin) of every row (several results may travel in one gather); bias (optional): added to every emitted score of
// the row.  Small merges (parts * kin <= 1024): one warp per query row, kout rounds of "best key below the previous one"
// (keys of a row are distinct: one per bank row).
__global__ void __launch_bounds__(256)
topk_merge_warp_kernel(const unsigned long long* __restrict__ keys, int parts, int64_t nq, int ld, int col0, int kin, int kout,
                       int keep_max, float* __restrict__ D, int64_t* __restrict__ I, const float* __restrict__ bias) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= nq) return;
  const int total = parts * kin;
  unsigned long long prev = ~0ull;
  for (int r = 0; r < kout; ++r) {
    unsigned long long best = 0ull;
    for (int c = lane; c < total; c += 32) {
      const unsigned long long key = keys[(static_cast<int64_t>(c / kin) * nq + row) * ld + col0 + c % kin];
      if (key < prev && key > best) best = key;
    }
    best = mg_warp_max(best);
    if (lane == 0) mg_emit(best, keep_max, D + row * kout + r, I + row * kout + r, bias ? bias[row] : 0.f);
    prev = best;           // 0 once the row is exhausted: every later round emits padding
  }
}

// Large merges: one CTA per query row, bitonic sort (descending) of the padded key list in shared memory.
__global__ void __launch_bounds__(512)
topk_merge_sort_kernel(const unsigned long long* __restrict__ keys, int parts, int64_t nq, int ld, int col0, int kin, int kout,
                       int npad, int keep_max, float* __restrict__ D, int64_t* __restrict__ I, const float* __restrict__ bias) {
  extern __shared__ unsigned long long mg_smem[];
  const int64_t row = blockIdx.x;
  const int total = parts * kin;
  for (int c = threadIdx.x; c < npad; c += blockDim.x)
    mg_smem[c] = c < total ? keys[(static_cast<int64_t>(c / kin) * nq + row) * ld + col0 + c % kin] : 0ull;
  __syncthreads();
  for (int size = 2; size <= npad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < (npad >> 1); i += blockDim.x) {
        const int lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const unsigned long long a = mg_smem[lo], b = mg_smem[hi];
        if ((a < b) == desc) { mg_smem[lo] = b; mg_smem[hi] = a; }
      }
      __syncthreads();
    }
  }
  for (int r = threadIdx.x; r < kout; r += blockDim.x)
    mg_emit(r < npad ? mg_smem[r] : 0ull, keep_max, D + row * kout + r, I + row * kout + r, bias ? bias[row] : 0.f);
}

}  // namespace vscb200

extern "C" {

int vscb200_topk_pack(const float* D_dev, const int64_t* I_dev, int64_t nq, int k, int keep_max, uint64_t* keys_dev, void* stream) {
  using namespace vscb200;
  VSCB_REQUIRE(nq >= 0 && k >= 1 && (nq == 0 || (D_dev && I_dev && keys_dev)), "topk_pack: bad argument");
  const int64_t n = nq * k;
  if (n == 0) return VSCB200_OK;
  topk_pack_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      D_dev, I_dev, n, keep_max, reinterpret_cast<unsigned long long*>(keys_dev));
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

int vscb200_topk_pack_cols(const float* D_dev, const int64_t* I_dev, int64_t nq, int k, int keep_max, uint64_t* keys_dev, int ld,
                           int col0, void* stream) {
  using namespace vscb200;
  VSCB_REQUIRE(nq >= 0 && k >= 1 && col0 >= 0 && col0 + k <= ld && (nq == 0 || (D_dev && I_dev && keys_dev)),
               "topk_pack_cols: bad argument");
  const int64_t n = nq * k;
  if (n == 0) return VSCB200_OK;
  topk_pack_cols_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      D_dev, I_dev, nq, k, keep_max, reinterpret_cast<unsigned long long*>(keys_dev), ld, col0);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

int vscb200_topk_merge_cols(const uint64_t* keys_dev, int parts, int64_t nq, int ld, int col0, int kin, int kout, int keep_max,
                            const float* bias_dev, float* D_dev, int64_t* I_dev, void* stream_v) {
  using namespace vscb200;
  VSCB_REQUIRE(parts >= 1 && nq >= 0 && kin >= 1 && kout >= 1 && col0 >= 0 && col0 + kin <= ld &&
                   (nq == 0 || (keys_dev && D_dev && I_dev)), "topk_merge: bad argument");
  if (nq == 0) return VSCB200_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const unsigned long long* keys = reinterpret_cast<const unsigned long long*>(keys_dev);
  const int64_t total = static_cast<int64_t>(parts) * kin;
  if (total <= 1024 && kout <= 32) {
    topk_merge_warp_kernel<<<static_cast<unsigned>((nq + 7) / 8), 256, 0, stream>>>(keys, parts, nq, ld, col0, kin, kout, keep_max,
                                                                                   D_dev, I_dev, bias_dev);
  } else {
    int npad = 2;
    while (npad < total) npad <<= 1;
    VSCB_REQUIRE(npad <= 16384, "topk_merge: more than 16384 partial results per query");
    const size_t smem = static_cast<size_t>(npad) * sizeof(unsigned long long);
    VSCB_CUDA_OK(cudaFuncSetAttribute(topk_merge_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    topk_merge_sort_kernel<<<static_cast<unsigned>(nq), 512, smem, stream>>>(keys, parts, nq, ld, col0, kin, kout, npad, keep_max,
                                                                            D_dev, I_dev, bias_dev);
  }
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

int vscb200_topk_merge(const uint64_t* keys_dev, int parts, int64_t nq, int kin, int kout, int keep_max, float* D_dev,
                       int64_t* I_dev, void* stream_v) {
  return vscb200_topk_merge_cols(keys_dev, parts, nq, kin, 0, kin, kout, keep_max, nullptr, D_dev, I_dev, stream_v);
}

}  // extern "C"
