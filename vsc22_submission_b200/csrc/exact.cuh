// Exact fp32 rescoring of (query, bank row) pairs by one warp -- the ONE summation order every search path
// reports scores in, so a pair's score does not depend on which kernel produced it (fused top-k, dense
// top-k, streaming search, range search).  d % 4 == 0: lane L owns elements 4L..4L+3 of every 128-element
// block (128-bit loads), fmaf chain in element order, then a butterfly reduction; otherwise lane-strided scalars.
#pragma once
#include <cuda_runtime.h>

namespace vscb200 {

// kRows independent rows at once (their loads overlap).  q: shared memory (16-byte aligned when d % 4 == 0).
template <int kRows>
__device__ __forceinline__ void exact_rows_warp(const float* __restrict__ q, const float* const (&r)[kRows], int d, int lane,
                                                bool l2, float (&acc)[kRows]) {
#pragma unroll
  for (int u = 0; u < kRows; ++u) acc[u] = 0.f;
  if ((d & 3) == 0) {
    for (int j = lane * 4; j < d; j += 128) {
      const float4 qv = *reinterpret_cast<const float4*>(q + j);
      float4 rv[kRows];
#pragma unroll
      for (int u = 0; u < kRows; ++u) rv[u] = __ldg(reinterpret_cast<const float4*>(r[u] + j));
#pragma unroll
      for (int u = 0; u < kRows; ++u) {
        if (l2) {
          const float a = qv.x - rv[u].x, b = qv.y - rv[u].y, c = qv.z - rv[u].z, e = qv.w - rv[u].w;
          acc[u] = fmaf(a, a, acc[u]); acc[u] = fmaf(b, b, acc[u]); acc[u] = fmaf(c, c, acc[u]); acc[u] = fmaf(e, e, acc[u]);
        } else {
          acc[u] = fmaf(qv.x, rv[u].x, acc[u]); acc[u] = fmaf(qv.y, rv[u].y, acc[u]);
          acc[u] = fmaf(qv.z, rv[u].z, acc[u]); acc[u] = fmaf(qv.w, rv[u].w, acc[u]);
        }
      }
    }
  } else {
    for (int j = lane; j < d; j += 32) {
      const float qj = q[j];
#pragma unroll
      for (int u = 0; u < kRows; ++u) {
        const float rv = __ldg(r[u] + j);
        if (l2) { const float df = qj - rv; acc[u] = fmaf(df, df, acc[u]); }
        else acc[u] = fmaf(qj, rv, acc[u]);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < kRows; ++u) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
  }
}

}  // namespace vscb200
