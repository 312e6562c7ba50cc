// Stable LSD radix sort of (uint64 key, uint64 value) pairs, 8 bits per pass, over a chosen bit range.
// Used by the global (cross-query) top-K candidate search to order the surviving frame pairs the way the
// reference's Python `sort(key=score, reverse=True)` does (vsc/index.py:162, vsc/candidates.py:39,
// M/infer/infer_matching.py:255): best score first, ties in (query row, bank row) order.
//
// Per pass: (1) per-tile digit histograms, written bucket-major; (2) one CTA per bucket scans its row of
// tile counts; (3) each tile re-reads its elements, ranks them stably inside the tile (warp match + per-warp
// counters, element order = warp-major) and scatters.  HBM-bound: 2 reads + 1 write of 16 B per element per pass.
#include <algorithm>

#include "host_util.h"
#include "kernels.h"

namespace vscb200 {
namespace {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortItems = 16;                          // elements per thread
constexpr int kSortTile = kSortThreads * kSortItems;    // 4096 elements per CTA
constexpr int kWarpSpan = 32 * kSortItems;              // contiguous elements owned by one warp

__global__ void __launch_bounds__(kSortThreads)
sort_hist_kernel(const uint64_t* __restrict__ keys, int64_t n, int shift, uint32_t* __restrict__ tile_hist, int ntiles) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = static_cast<int64_t>(blockIdx.x) * kSortTile;
#pragma unroll 4
  for (int it = 0; it < kSortItems; ++it) {
    const int64_t i = base + it * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  tile_hist[static_cast<size_t>(threadIdx.x) * ntiles + blockIdx.x] = h[threadIdx.x];
}

// CTA b: exclusive scan of tile_hist[b][0..ntiles) in place; totals[b] = the row sum
__global__ void __launch_bounds__(1024)
sort_scan_kernel(uint32_t* __restrict__ tile_hist, int ntiles, uint32_t* __restrict__ totals) {
  __shared__ uint32_t wsum[32];
  __shared__ uint32_t carry_s;
  uint32_t* row = tile_hist + static_cast<size_t>(blockIdx.x) * ntiles;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < ntiles; base += 1024) {
    const int i = base + threadIdx.x;
    const uint32_t v = i < ntiles ? row[i] : 0u;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = wsum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      wsum[lane] = w;   // inclusive over warps
    }
    __syncthreads();
    const uint32_t carry = carry_s;
    const uint32_t before = carry + (warp ? wsum[warp - 1] : 0u) + x - v;
    if (i < ntiles) row[i] = before;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + wsum[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) totals[blockIdx.x] = carry_s;
}

__global__ void __launch_bounds__(kSortThreads)
sort_scatter_kernel(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ vals, uint64_t* __restrict__ keys_out,
                    uint64_t* __restrict__ vals_out, int64_t n, int shift, const uint32_t* __restrict__ tile_hist,
                    const uint32_t* __restrict__ totals, int ntiles) {
  __shared__ uint32_t cnt[kSortWarps][256];
  __shared__ uint32_t tot[256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int w = 0; w < kSortWarps; ++w) cnt[w][threadIdx.x] = 0;
  tot[threadIdx.x] = totals[threadIdx.x];
  __syncthreads();

  const int64_t base = static_cast<int64_t>(blockIdx.x) * kSortTile + warp * kWarpSpan + lane;
  uint64_t k[kSortItems];
  uint32_t rank[kSortItems];
#pragma unroll
  for (int it = 0; it < kSortItems; ++it) {
    const int64_t i = base + it * 32;
    const bool valid = i < n;
    k[it] = valid ? keys[i] : 0ull;
    const uint32_t b = static_cast<uint32_t>(k[it] >> shift) & 255u;
    const uint32_t peers = __match_any_sync(0xffffffffu, valid ? b : 256u + lane);
    const uint32_t lower = __popc(peers & ((1u << lane) - 1u));
    rank[it] = valid ? cnt[warp][b] + lower : 0u;
    __syncwarp();
    if (valid && lower == 0) cnt[warp][b] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  {
    // bucket t: global start = sum of totals of lower buckets + this tile's offset inside the bucket
    const int t = threadIdx.x;
    uint32_t run = tile_hist[static_cast<size_t>(t) * ntiles + blockIdx.x];
    for (int j = 0; j < t; ++j) run += tot[j];
    for (int w = 0; w < kSortWarps; ++w) {
      const uint32_t c = cnt[w][t];
      cnt[w][t] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < kSortItems; ++it) {
    const int64_t i = base + it * 32;
    if (i < n) {
      const uint32_t b = static_cast<uint32_t>(k[it] >> shift) & 255u;
      const uint32_t dst = cnt[warp][b] + rank[it];
      keys_out[dst] = k[it];
      vals_out[dst] = vals[i];
    }
  }
}

}  // namespace

size_t radix_sort_scratch_bytes(int64_t n) {
  const int64_t ntiles = (n + kSortTile - 1) / kSortTile;
  return (static_cast<size_t>(std::max<int64_t>(ntiles, 1)) * 256 + 256) * sizeof(uint32_t);
}

// Sorts ascending by bits [bit_lo, bit_hi) of the key, stably.  (k0, v0) is the input; (k1, v1) the alternate
// buffers.  *result_in_alt tells which pair holds the output (the passes ping-pong).
int radix_sort_pairs(uint64_t* k0, uint64_t* v0, uint64_t* k1, uint64_t* v1, int64_t n, int bit_lo, int bit_hi,
                     void* scratch, int* result_in_alt, cudaStream_t stream) {
  *result_in_alt = 0;
  if (n <= 1 || bit_hi <= bit_lo) return VSCB200_OK;
  VSCB_REQUIRE(n < (1ll << 31), "radix_sort_pairs: more than 2^31 elements");
  const int ntiles = static_cast<int>((n + kSortTile - 1) / kSortTile);
  uint32_t* tile_hist = static_cast<uint32_t*>(scratch);
  uint32_t* totals = tile_hist + static_cast<size_t>(ntiles) * 256;
  uint64_t *ka = k0, *va = v0, *kb = k1, *vb = v1;
  for (int shift = bit_lo; shift < bit_hi; shift += 8) {
    sort_hist_kernel<<<ntiles, kSortThreads, 0, stream>>>(ka, n, shift, tile_hist, ntiles);
    sort_scan_kernel<<<256, 1024, 0, stream>>>(tile_hist, ntiles, totals);
    sort_scatter_kernel<<<ntiles, kSortThreads, 0, stream>>>(ka, va, kb, vb, n, shift, tile_hist, totals, ntiles);
    count_launch(3);
    std::swap(ka, kb);
    std::swap(va, vb);
    *result_in_alt ^= 1;
  }
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

}  // namespace vscb200
