// Error bound of a single bf16 tensor-core pass Qh.Rh against the exact fp32 score q.r, shared by the single-pass top-k
// search (sim_tc1.cu) and the single-pass candidate emission (sim_tc.cu emit mode / global_topk.cu).
//   q.r - qh.rh = ql.r + qh.rl   =>   |error| <= |ql| Rmax + |qh| RLmax  (Cauchy-Schwarz), Rmax / RLmax = the largest
// |r| / |r - bf16(r)| over the bank (maintained at add(), index.cu), |ql| computed per query row; plus the fp32
// accumulation of d products.  In key units: L2 keys |q|^2 + |r|^2 - 2 q.r carry twice the product's error.
#pragma once

namespace vscb200 {

__device__ __forceinline__ float sim1_eps(float qn2, float qlo2, const unsigned int* __restrict__ bank_max_bits, int d, int l2) {
  const float qnorm = sqrtf(qn2), qlo = sqrtf(qlo2);
  const float rmax2 = __uint_as_float(bank_max_bits[0]);
  const float rmax = sqrtf(rmax2), rlmax = sqrtf(__uint_as_float(bank_max_bits[1]));
  float eps = (qlo * rmax + (qnorm + qlo) * rlmax + static_cast<float>(d) * 1.1920929e-7f * qnorm * rmax) * 1.001f;
  if (l2) eps = 2.0f * eps + 1e-6f * (qn2 + rmax2);
  return eps;
}

}  // namespace vscb200
