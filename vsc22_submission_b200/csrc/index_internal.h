// The flat index object behind the opaque `vscb200_index*` handle, shared between index.cu (search / range search)
// and global_topk.cu (cross-query top-K candidate search).
#pragma once
#include <stdint.h>

#include <algorithm>
#include <vector>

#include "host_util.h"

struct vscb200_index {
  int d = 0, metric = 0;
  int device = 0;            // CUDA device the index lives on (current device at create)
  int64_t ntotal = 0;        // rows resident on the device
  int64_t capacity = 0;
  float* bank = nullptr;     // [capacity, d]
  float* rnorm = nullptr;    // [capacity] squared norms (L2 metric only)
  int dp = 0;                // d rounded up to 8: row length of the bf16 operand planes
  uint16_t* bank_hi = nullptr;   // [capacity, dp] bf16(x)
  uint16_t* bank_lo = nullptr;   // [capacity, dp] bf16(x - hi)
  uint16_t* q_planes = nullptr; size_t q_planes_bytes = 0;   // hi | lo planes of the current query block
  cudaEvent_t order_ev = nullptr;   // orders this index's work when consecutive calls arrive on different streams
  bool has_stream = false;
  int64_t id_offset = 0;
  float* ws = nullptr;       // score workspace
  size_t ws_bytes = 0;
  // staging for the host-buffer API
  float* q_stage = nullptr; size_t q_stage_bytes = 0;
  float* D_stage = nullptr; size_t D_stage_bytes = 0;
  int64_t* I_stage = nullptr; size_t I_stage_bytes = 0;
  float* qnorm = nullptr; size_t qnorm_bytes = 0;
  unsigned long long* counts = nullptr; size_t counts_bytes = 0;
  float* Dtmp = nullptr; size_t Dtmp_bytes = 0;       // survivors of the tensor-core pass (k + slack per row)
  int64_t* Itmp = nullptr; size_t Itmp_bytes = 0;
  float* cand_d = nullptr; size_t cand_d_bytes = 0;   // fused top-k epilogue candidates [nq, slabs*2*kFK]
  int32_t* cand_i = nullptr; size_t cand_i_bytes = 0;
  int no_fused = 0;
  int no_stream = 0;
  int sim_passes = 1;        // 1: single bf16 pass + margin + exact rescoring (sim_tc1.cu); 3: split-bf16 fused top-k (sim_tc.cu)
  unsigned int* rmax2_bits = nullptr;   // device [2]: max |r|^2 and max |r - bf16(r)|^2 as float bits (single-pass search margin)
  bool rmax2_reset = true;
  int* flags = nullptr; size_t flags_bytes = 0;   // sim_tc1.cu scratch: [nq] overflow flags + [1] their count + [nq] shared thresholds + [nq] list lengths
  int* last_flag_count = nullptr;                 // -> the count of the last single-pass search (diagnostics)
  float* gmax = nullptr; size_t gmax_bytes = 0;       // streaming search: per (32-row group, query) maxima
  unsigned long long* gr_keys = nullptr; size_t gr_keys_bytes = 0;   // streaming search: rescored keys per (query, group, row)
  int* gr_count = nullptr;                            // [128] arrival counters of group_rescore_few (zero between calls)
  // results of the last global (cross-query) candidate search, kept until the next one (global_topk.cu)
  float* g_score = nullptr; size_t g_score_bytes = 0;   // [g_n] exact scores, best first
  int64_t* g_q = nullptr; size_t g_q_bytes = 0;         // [g_n] query rows
  int64_t* g_r = nullptr; size_t g_r_bytes = 0;         // [g_n] bank rows (+ id_offset)
  int64_t g_n = 0;
  float* vp_score = nullptr; size_t vp_score_bytes = 0; // [vp_n] best score per (query video, ref video)
  int64_t* vp_q = nullptr; size_t vp_q_bytes = 0;
  int64_t* vp_r = nullptr; size_t vp_r_bytes = 0;
  int64_t vp_n = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t last_stream = nullptr;   // stream of the most recent call (orders the final frees)
  int force_simt = 0;
};

namespace vscb200 {

template <typename Tp>
int grow(Tp** p, size_t* have, size_t need, cudaStream_t s) {
  if (*have >= need) return VSCB200_OK;
  if (*p) pool_free(*p, s);
  *p = nullptr;
  *have = 0;
  size_t want = std::max(need, static_cast<size_t>(256));
  int rc = pool_alloc(reinterpret_cast<void**>(p), want, s);
  if (rc) return rc;
  *have = want;
  return VSCB200_OK;
}

size_t ws_budget_bytes();
int flush_pending(vscb200_index* ix, cudaStream_t s);
int64_t block_rows(const vscb200_index* ix, int64_t nq);
// S[0:nq, 0:ntotal] for one query block (also fills ix->qnorm[0:nq])
int score_block(vscb200_index* ix, const float* q, int64_t nq, float* S, int64_t ldS, cudaStream_t s);
int own_stream(vscb200_index* ix, cudaStream_t* s);

}  // namespace vscb200
