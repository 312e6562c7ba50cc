// Dense similarity block S[nq, nr] = Q . R^T (inner product) or squared L2, fp32, plus the score
// normalisation prologue kernels.  Reference: faiss IndexFlat scoring behind vsc/index.py:174,
// vsc/exhaustive_search.py:62,74, score_normalization.py:71-103; per-pair sim matrices of
// vsc/baseline/localization.py:32-35.
//
// scores_simt: exact-fp32 FFMA tile kernel (64x64 tile, 4x4 per thread) -- the shape-generic path
// (any d, including the d=3 vectors of the reference's tests).  The tcgen05 split-bf16 kernel in
// sim_tc.cu takes over for d % 8 == 0.
#include "host_util.h"
#include "kernels.h"
#include "ptx.cuh"

namespace vscb200 {

constexpr int kST = 64;    // tile edge
constexpr int kSK = 16;    // k step

template <bool kL2>
__global__ void __launch_bounds__(256)
scores_simt_kernel(const float* __restrict__ Q, const float* __restrict__ R, float* __restrict__ S, int64_t nq,
                   int64_t nr, int d, int64_t ldS, const float* __restrict__ qn, const float* __restrict__ rn) {
  __shared__ float sQ[kSK][kST + 4];
  __shared__ float sR[kSK][kST + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t q0 = static_cast<int64_t>(blockIdx.y) * kST, r0 = static_cast<int64_t>(blockIdx.x) * kST;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < d; k0 += kSK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256;
      const int rr = idx >> 4, kk = idx & 15;
      const bool kin = k0 + kk < d;
      sQ[kk][rr] = (kin && q0 + rr < nq) ? Q[(q0 + rr) * d + k0 + kk] : 0.f;
      sR[kk][rr] = (kin && r0 + rr < nr) ? R[(r0 + rr) * d + k0 + kk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kSK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&sQ[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&sR[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t q = q0 + ty * 4 + i;
    if (q >= nq) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t r = r0 + tx * 4 + j;
      if (r >= nr) continue;
      float v = acc[i][j];
      if (kL2) v = fmaxf(qn[q] + rn[r] - 2.0f * v, 0.f);
      S[q * ldS + r] = v;
    }
  }
}

int scores_simt(const float* Q, const float* R, float* S, int64_t nq, int64_t nr, int d, int64_t ldS, bool l2,
                const float* qn, const float* rn, cudaStream_t stream) {
  if (nq == 0 || nr == 0) return VSCB200_OK;
  dim3 grid(static_cast<unsigned>((nr + kST - 1) / kST), static_cast<unsigned>((nq + kST - 1) / kST));
  VSCB_REQUIRE(grid.y <= 65535, "scores: too many query rows in one block");
  ProfScope prof(kProfScores, stream, 2.0 * static_cast<double>(nq) * nr * d);
  if (l2)
    scores_simt_kernel<true><<<grid, 256, 0, stream>>>(Q, R, S, nq, nr, d, ldS, qn, rn);
  else
    scores_simt_kernel<false><<<grid, 256, 0, stream>>>(Q, R, S, nq, nr, d, ldS, qn, rn);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

// ------------------------------------------------------------------ squared row norms (warp per row)
__global__ void row_sqnorm_kernel(const float* __restrict__ x, int64_t n, int d, float* __restrict__ out) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  float s = 0.f;
  for (int c = lane; c < d; c += 32) {
    const float v = x[row * d + c];
    s = fmaf(v, v, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[row] = s;
}

int row_sqnorm(const float* x, int64_t n, int d, float* out, cudaStream_t stream) {
  if (n == 0) return VSCB200_OK;
  row_sqnorm_kernel<<<static_cast<unsigned>((n * 32 + 255) / 256), 256, 0, stream>>>(x, n, d, out);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

// ------------------------------------------------------------------ score-normalisation prologue
// out[row] = [ l2norm(x[row] without column drop_dim) , last ]   (score_normalization.py:73-83,96-101)
// sklearn.normalize semantics: zero rows stay zero.  Norm accumulated in fp32 like sklearn (float32 in).
// drop_dim_dev != nullptr: the dropped column is read from device memory (written by var_argmin_kernel earlier in the
// stream), so the whole score-normalisation chain is enqueued without a host round trip.
__global__ void sn_transform_kernel(const float* __restrict__ x, int64_t n, int d, int drop_dim, int l2_normalize,
                                    float fill, const float* __restrict__ bias, float* __restrict__ out,
                                    const int* __restrict__ drop_dim_dev) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  if (drop_dim_dev) drop_dim = *drop_dim_dev;
  const float* xr = x + row * d;
  float s = 0.f;
  for (int c = lane; c < d; c += 32) {
    if (c == drop_dim) continue;
    const float v = xr[c];
    s = fmaf(v, v, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  float denom = 1.0f;
  if (l2_normalize) {
    denom = sqrtf(s);
    if (denom == 0.f) denom = 1.0f;
  }
  // output layout: kept columns in order, then the extra dimension.  With drop_dim < 0 nothing is
  // dropped and the output has d+1 columns.
  const int dout = (drop_dim >= 0 && drop_dim < d) ? d : d + 1;
  float* orow = out + row * dout;
  for (int c = lane; c < d; c += 32) {
    if (c == drop_dim) continue;
    const int oc = (drop_dim >= 0 && c > drop_dim) ? c - 1 : c;
    orow[oc] = xr[c] / denom;
  }
  if (lane == 0) orow[dout - 1] = bias ? bias[row] : fill;
}

int sn_transform(const float* x, int64_t n, int d, int drop_dim, int l2_normalize, float fill, const float* bias,
                 float* out, cudaStream_t stream, const int* drop_dim_dev) {
  if (n == 0) return VSCB200_OK;
  VSCB_REQUIRE(drop_dim < d, "sn_transform: drop_dim out of range");
  sn_transform_kernel<<<static_cast<unsigned>((n * 32 + 255) / 256), 256, 0, stream>>>(x, n, d, drop_dim, l2_normalize,
                                                                                      fill, bias, out, drop_dim_dev);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

// The same transform written STRAIGHT into an index's storage (index.cu vscb200_index_add_sn): the fp32 bank row, both
// bf16 operand planes, the squared norm of the transformed row and the bank's running norm maxima in one kernel -- what
// sn_transform + the device copy of add() + the plane pass did in three (107 -> 45 us per 40k x 512 rows).  Arithmetic
// and summation orders are those of sn_transform_kernel and q_hi_norm_kernel, so every stored value is bit-identical to
// the unfused path.
__global__ void __launch_bounds__(256)
sn_add_rows_kernel(const float* __restrict__ x, int64_t n, int d, int drop_dim, int l2_normalize, float fill,
                   const float* __restrict__ bias, const int* __restrict__ drop_dim_dev, float* __restrict__ bank,
                   __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, float* __restrict__ rnorm,
                   unsigned int* __restrict__ max_bits, int dout, int dp) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  if (drop_dim_dev) drop_dim = *drop_dim_dev;
  const float* xr = x + row * d;
  float s = 0.f;
  for (int c = lane; c < d; c += 32) {
    if (c == drop_dim) continue;
    const float v = xr[c];
    s = fmaf(v, v, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  float denom = 1.0f;
  if (l2_normalize) {
    denom = sqrtf(s);
    if (denom == 0.f) denom = 1.0f;
  }
  const float last = bias ? bias[row] : fill;
  float s2 = 0.f, sl = 0.f;
  for (int oc = lane; oc < dp; oc += 32) {                 // output columns, in q_hi_norm_kernel's order
    float v = 0.f;
    if (oc < dout - 1) v = xr[(drop_dim >= 0 && oc >= drop_dim) ? oc + 1 : oc] / denom;
    else if (oc == dout - 1) v = last;
    if (oc < dout) bank[row * dout + oc] = v;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const float e = v - __bfloat162float(h);
    s2 = fmaf(v, v, s2);
    sl = fmaf(e, e, sl);
    hi[row * dp + oc] = h;
    lo[row * dp + oc] = __float2bfloat16_rn(e);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    sl += __shfl_xor_sync(0xffffffffu, sl, o);
  }
  if (lane == 0) {
    rnorm[row] = s2;
    if (__float_as_uint(s2) > *reinterpret_cast<volatile unsigned int*>(max_bits)) atomicMax(max_bits, __float_as_uint(s2));
    if (__float_as_uint(sl) > *reinterpret_cast<volatile unsigned int*>(max_bits + 1)) atomicMax(max_bits + 1, __float_as_uint(sl));
  }
}

int sn_add_rows(const float* x, int64_t n, int d, int drop_dim, int l2_normalize, float fill, const float* bias,
                const int* drop_dim_dev, float* bank, void* hi, void* lo, float* rnorm, unsigned int* max_bits, int dout, int dp,
                cudaStream_t stream) {
  if (n == 0) return VSCB200_OK;
  VSCB_REQUIRE(drop_dim < d, "sn_add_rows: drop_dim out of range");
  sn_add_rows_kernel<<<static_cast<unsigned>((n * 32 + 255) / 256), 256, 0, stream>>>(
      x, n, d, drop_dim, l2_normalize, fill, bias, drop_dim_dev, bank, reinterpret_cast<__nv_bfloat16*>(hi),
      reinterpret_cast<__nv_bfloat16*>(lo), rnorm, max_bits, dout, dp);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

// bias[row] = -beta * mean(D[row, :nk])
__global__ void sn_bias_kernel(const float* __restrict__ D, int64_t nq, int k, int nk, float beta,
                               float* __restrict__ bias) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (row >= nq) return;
  float s = 0.f;
  for (int j = 0; j < nk; ++j) s += D[row * k + j];
  bias[row] = -beta * (s / nk);
}

// Column statistics for `sn_features.var(axis=0).argmin()` (score_normalization.py:72), two passes in double with a
// FIXED reduction order (no atomics): pass 1 column sums, pass 2 sums of squared deviations from the mean.  Row-sharded
// banks all-reduce the [d] vector between the passes (sharding.py).  partial[slab][c]: grid.y row slabs; block 32 x 8.
__global__ void col_partial_kernel(const float* __restrict__ x, int64_t n, int d, const double* __restrict__ sum_in,
                                   double inv_n, double* __restrict__ partial) {
  __shared__ double red[8][32];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int64_t rows_per = (n + gridDim.y - 1) / gridDim.y;
  const int64_t r0 = blockIdx.y * rows_per, r1 = min(n, r0 + rows_per);
  double s = 0.0;
  if (c < d) {
    const double mean = sum_in ? sum_in[c] * inv_n : 0.0;
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) {
      const double v = static_cast<double>(x[r * d + c]) - mean;
      s += sum_in ? v * v : v;
    }
  }
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < d) {
    double t = red[0][threadIdx.x];
#pragma unroll
    for (int i = 1; i < 8; ++i) t += red[i][threadIdx.x];
    partial[static_cast<int64_t>(blockIdx.y) * d + c] = t;
  }
}

__global__ void col_finish_kernel(const double* __restrict__ partial, int slabs, int d, double* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  double t = 0.0;
  for (int sl = 0; sl < slabs; ++sl) t += partial[static_cast<int64_t>(sl) * d + c];
  out[c] = t;
}

// out[c] = sum_r x[r, c]                         (sum_in == nullptr)
//        = sum_r (x[r, c] - sum_in[c] * inv_n)^2 (sum_in != nullptr)
int col_sums(const float* x, int64_t n, int d, const double* sum_in, double inv_n, double* out, cudaStream_t stream) {
  // row slabs of ~256 rows: enough CTAs to keep the HBM pipes full (2048-row slabs: 304 CTAs, 73 us per pass over 40k x 512)
  const int slabs = static_cast<int>(n < 4096 ? 1 : (n / 256 > 1024 ? 1024 : n / 256));
  double* partial = nullptr;
  int rc = pool_alloc(reinterpret_cast<void**>(&partial), sizeof(double) * static_cast<size_t>(slabs) * d, stream);
  if (rc) return rc;
  if (n > 0) {
    col_partial_kernel<<<dim3((d + 31) / 32, slabs), dim3(32, 8), 0, stream>>>(x, n, d, sum_in, inv_n, partial);
    col_finish_kernel<<<(d + 127) / 128, 128, 0, stream>>>(partial, slabs, d, out);
    count_launch(2);
  } else {
    cudaMemsetAsync(out, 0, sizeof(double) * d, stream);
  }
  pool_free(partial, stream);
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

// first minimum (numpy argmin) of the column variances; the common factor 1/n does not change it.  One warp.
__global__ void var_argmin_kernel(const double* __restrict__ ss, int d, int* __restrict__ out) {
  const int lane = threadIdx.x;
  int best = -1;
  double best_var = 0;
  for (int c = lane; c < d; c += 32) {
    const double var = ss[c];
    if (best < 0 || var < best_var) { best = c; best_var = var; }      // strict: keeps the first of equal values
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, best_var, o);
    const int ob = __shfl_xor_sync(0xffffffffu, best, o);
    if (ob >= 0 && (best < 0 || ov < best_var || (ov == best_var && ob < best))) { best = ob; best_var = ov; }
  }
  if (lane == 0) *out = best < 0 ? 0 : best;
}

// both passes + argmin for a bank that lives on one device; scratch: 2 d doubles
int low_var_dim_local(const float* x, int64_t n, int d, double* scratch, int* dim_dev, cudaStream_t stream) {
  int rc = col_sums(x, n, d, nullptr, 0.0, scratch, stream);
  if (rc) return rc;
  if ((rc = col_sums(x, n, d, scratch, 1.0 / static_cast<double>(n), scratch + d, stream))) return rc;
  var_argmin_kernel<<<1, 32, 0, stream>>>(scratch + d, d, dim_dev);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

}  // namespace vscb200

extern "C" {
int vscb200_sn_transform(const float* x_dev, int64_t n, int d, int drop_dim, int l2_normalize, float fill,
                         const float* bias_dev, float* out_dev, void* stream) {
  return vscb200::sn_transform(x_dev, n, d, drop_dim, l2_normalize, fill, bias_dev, out_dev,
                               static_cast<cudaStream_t>(stream), nullptr);
}

int vscb200_sn_transform_dev(const float* x_dev, int64_t n, int d, const int* drop_dim_dev, int l2_normalize, float fill,
                             const float* bias_dev, float* out_dev, void* stream) {
  using namespace vscb200;
  VSCB_REQUIRE(drop_dim_dev, "sn_transform_dev: null drop_dim_dev");
  return vscb200::sn_transform(x_dev, n, d, 0, l2_normalize, fill, bias_dev, out_dev, static_cast<cudaStream_t>(stream),
                               drop_dim_dev);
}

int vscb200_low_var_dim_dev(const float* x_dev, int64_t n, int d, int* dim_dev, void* stream_v) {
  using namespace vscb200;
  VSCB_REQUIRE(n > 0 && d > 0 && dim_dev, "low_var_dim_dev: empty input");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  double* scratch = nullptr;
  int rc = pool_alloc(reinterpret_cast<void**>(&scratch), sizeof(double) * 2 * d, stream);
  if (rc) return rc;
  rc = low_var_dim_local(x_dev, n, d, scratch, dim_dev, stream);
  pool_free(scratch, stream);
  return rc;
}

int vscb200_col_sums(const float* x_dev, int64_t n, int d, const double* sum_in_dev, double inv_n, double* out_dev,
                     void* stream) {
  using namespace vscb200;
  VSCB_REQUIRE(n >= 0 && d > 0 && out_dev && (n == 0 || x_dev), "col_sums: bad argument");
  return col_sums(x_dev, n, d, sum_in_dev, inv_n, out_dev, static_cast<cudaStream_t>(stream));
}

int vscb200_var_argmin_dev(const double* ss_dev, int d, int* dim_dev, void* stream) {
  using namespace vscb200;
  VSCB_REQUIRE(ss_dev && d > 0 && dim_dev, "var_argmin_dev: bad argument");
  var_argmin_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(ss_dev, d, dim_dev);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

namespace vscb200 {
// Row-sharded banks, ONE collective: every shard contributes (S = sum x, M2 = sum (x - S / n)^2, Q = S^2 / n) per column;
// the sums over the shards combine exactly (Chan et al.): M2_total = sum M2 + sum Q - (sum S)^2 / N.
__global__ void col_moment_pack_kernel(const double* __restrict__ sums, double n, int d, double* __restrict__ out3) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  out3[c] = sums[c];
  out3[2 * d + c] = n > 0 ? sums[c] * sums[c] / n : 0.0;
}
__global__ void var_argmin_moments_kernel(const double* __restrict__ m3, double n_total, int d, int* __restrict__ out) {
  const int lane = threadIdx.x;
  int best = -1;
  double best_var = 0;
  for (int c = lane; c < d; c += 32) {
    const double var = m3[d + c] + m3[2 * d + c] - m3[c] * m3[c] / n_total;
    if (best < 0 || var < best_var) { best = c; best_var = var; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, best_var, o);
    const int ob = __shfl_xor_sync(0xffffffffu, best, o);
    if (ob >= 0 && (best < 0 || ov < best_var || (ov == best_var && ob < best))) { best = ob; best_var = ov; }
  }
  if (lane == 0) *out = best < 0 ? 0 : best;
}
}  // namespace vscb200

int vscb200_col_moments_local(const float* x_dev, int64_t n, int d, double* out3_dev, void* stream_v) {
  using namespace vscb200;
  VSCB_REQUIRE(n >= 0 && d > 0 && out3_dev && (n == 0 || x_dev), "col_moments_local: bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream_v);
  double* sums = nullptr;
  int rc = pool_alloc(reinterpret_cast<void**>(&sums), sizeof(double) * d, s);
  if (rc) return rc;
  rc = col_sums(x_dev, n, d, nullptr, 0.0, sums, s);
  if (rc == VSCB200_OK) rc = col_sums(x_dev, n, d, sums, n > 0 ? 1.0 / static_cast<double>(n) : 0.0, out3_dev + d, s);
  if (rc == VSCB200_OK) {
    col_moment_pack_kernel<<<(d + 127) / 128, 128, 0, s>>>(sums, static_cast<double>(n), d, out3_dev);
    count_launch();
  }
  pool_free(sums, s);
  if (rc) return rc;
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

int vscb200_var_argmin_moments(const double* m3_dev, double n_total, int d, int* dim_dev, void* stream) {
  using namespace vscb200;
  VSCB_REQUIRE(m3_dev && d > 0 && n_total > 0 && dim_dev, "var_argmin_moments: bad argument");
  var_argmin_moments_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(m3_dev, n_total, d, dim_dev);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

namespace vscb200 {
// score_normalizev2 (M/vsc/baseline/score_normalization.py:141-153): out[row] = l2_normalize(x[row] - beta * mean_k z[ids[row, k]]).
// One warp per row; the mean accumulates the nk rows in order in fp32 (numpy's reduction over the middle axis), the norm is
// sklearn's normalize: x / max(||x||, eps -> the row is left as is when its norm is 0).
__global__ void __launch_bounds__(256)
sn2_adapt_kernel(const float* __restrict__ x, const float* __restrict__ z, const int64_t* __restrict__ ids, int64_t n, int d,
                 int nk, float beta, int l2_normalize, float* __restrict__ out) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  float ss = 0.f;
  for (int c = lane; c < d; c += 32) {
    float acc = 0.f;
    for (int k = 0; k < nk; ++k) {
      const int64_t id = ids[row * nk + k];
      acc += id >= 0 ? z[id * d + c] : 0.f;
    }
    const float v = x[row * d + c] - (acc / static_cast<float>(nk)) * beta;
    out[row * d + c] = v;
    ss = fmaf(v, v, ss);
  }
  if (!l2_normalize) return;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float nrm = sqrtf(ss);
  if (nrm > 0.f)
    for (int c = lane; c < d; c += 32) out[row * d + c] = out[row * d + c] / nrm;
}
}  // namespace vscb200

int vscb200_sn2_adapt(const float* x_dev, const float* z_dev, const int64_t* ids_dev, int64_t n, int d, int nk, float beta,
                      int l2_normalize, float* out_dev, void* stream) {
  using namespace vscb200;
  VSCB_REQUIRE(d > 0 && nk >= 1 && (n == 0 || (x_dev && z_dev && ids_dev && out_dev)), "sn2_adapt: bad argument");
  if (n == 0) return VSCB200_OK;
  sn2_adapt_kernel<<<static_cast<unsigned>((n * 32 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x_dev, z_dev, ids_dev, n, d, nk, beta, l2_normalize, out_dev);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

int vscb200_sn_bias(const float* D_dev, int64_t nq, int k, int nk, float beta, float* bias_dev, void* stream) {
  using namespace vscb200;
  VSCB_REQUIRE(nk >= 1 && nk <= k, "sn_bias: need 1 <= nk <= k");
  if (nq == 0) return VSCB200_OK;
  sn_bias_kernel<<<static_cast<unsigned>((nq + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      D_dev, nq, k, nk, beta, bias_dev);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

int vscb200_low_var_dim(const float* x_dev, int64_t n, int d, int* dim_host, void* stream_v) {
  using namespace vscb200;
  VSCB_REQUIRE(n > 0 && d > 0 && dim_host, "low_var_dim: empty input");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  // scratch from the caching pool (no cudaMalloc / cudaFree -- each is a device-wide synchronisation -- per call);
  // the argmin runs on the device, 4 bytes come back through a page-locked word
  double* sums = nullptr;
  int rc = pool_alloc(reinterpret_cast<void**>(&sums), sizeof(double) * 2 * d + sizeof(int), stream);
  if (rc) return rc;
  int* best_dev = reinterpret_cast<int*>(sums + 2 * d);
  static thread_local int* best_pinned = nullptr;
  if (!best_pinned && cudaMallocHost(&best_pinned, sizeof(int)) != cudaSuccess) {
    cudaGetLastError();
    pool_free(sums, stream);
    set_last_error("low_var_dim: cudaMallocHost failed");
    return VSCB200_ERR_NOMEM;
  }
  cudaError_t e = cudaSuccess;
  if ((rc = low_var_dim_local(x_dev, n, d, sums, best_dev, stream))) {
    pool_free(sums, stream);
    return rc;
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(best_pinned, best_dev, sizeof(int), cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  pool_free(sums, stream);
  if (e != cudaSuccess) {
    set_last_error(std::string("low_var_dim: ") + cudaGetErrorString(e));
    return VSCB200_ERR_CUDA;
  }
  *dim_host = *best_pinned;
  return VSCB200_OK;
}
}
