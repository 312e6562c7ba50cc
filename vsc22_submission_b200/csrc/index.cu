// Flat similarity index (C ABI of seam B): owns the device copy of the descriptor bank, a score
// workspace and staging buffers; search = dense score block (tensor-core or SIMT) + row select.
// Reference: faiss IndexFlat behind vsc/index.py:74-177, vsc/exhaustive_search.py:52-92,206-292,
// vsc/baseline/score_normalization.py:87-98, M/infer/infer_matching.py:217-247.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "host_util.h"
#include "index_internal.h"
#include "kernels.h"

using namespace vscb200;

namespace vscb200 {
// sim_tc.cu: fp32-equivalent tensor-core scoring (split-bf16 tcgen05) on pre-split operand planes
int split_planes(const float* x, void* hi, void* lo, int64_t n, int d, int dp, cudaStream_t stream);
int fused_topk_slabs(int64_t nq, int64_t nr);
int fused_topk_list_len();
int fused_topk_group_rows();
int topk_tc_fused(const void* Qh, const void* Ql, const void* Rh, const void* Rl, int64_t nq, int64_t nr, int dp, bool l2,
                  const float* qn, const float* rn, int slabs, float* cand_d, int32_t* cand_i, cudaStream_t stream);
}  // namespace vscb200

namespace vscb200 {

size_t ws_budget_bytes() {
  const char* e = getenv("VSCB200_WS_MB");     // read per call: tests shrink it to force many query blocks
  size_t v = (e ? static_cast<size_t>(atoll(e)) : 2048) << 20;
  if (v < (1u << 20)) v = 1u << 20;
  return v;
}

int ensure_capacity(vscb200_index* ix, int64_t rows, cudaStream_t s) {
  if (rows <= ix->capacity) return VSCB200_OK;
  int64_t cap = std::max<int64_t>(rows, ix->capacity + ix->capacity / 2);
  cap = std::max<int64_t>(cap, 1024);
  float* nb = nullptr;
  int rc0 = pool_alloc(reinterpret_cast<void**>(&nb), static_cast<size_t>(cap) * ix->d * sizeof(float), s);
  if (rc0) return rc0;
  if (ix->ntotal) {
    VSCB_CUDA_OK(cudaMemcpyAsync(nb, ix->bank, static_cast<size_t>(ix->ntotal) * ix->d * sizeof(float),
                                 cudaMemcpyDeviceToDevice, s));
  }
  float* nn = nullptr;
  {
    if ((rc0 = pool_alloc(reinterpret_cast<void**>(&nn), static_cast<size_t>(cap) * sizeof(float), s))) return rc0;
    if (ix->ntotal)
      VSCB_CUDA_OK(cudaMemcpyAsync(nn, ix->rnorm, static_cast<size_t>(ix->ntotal) * sizeof(float),
                                   cudaMemcpyDeviceToDevice, s));
  }
  uint16_t *nh = nullptr, *nl = nullptr;
  if (!ix->force_simt) {
    const size_t pb = static_cast<size_t>(cap) * ix->dp * sizeof(uint16_t);
    if ((rc0 = pool_alloc(reinterpret_cast<void**>(&nh), pb, s))) return rc0;
    if ((rc0 = pool_alloc(reinterpret_cast<void**>(&nl), pb, s))) return rc0;
    if (ix->ntotal) {
      const size_t ob = static_cast<size_t>(ix->ntotal) * ix->dp * sizeof(uint16_t);
      VSCB_CUDA_OK(cudaMemcpyAsync(nh, ix->bank_hi, ob, cudaMemcpyDeviceToDevice, s));
      VSCB_CUDA_OK(cudaMemcpyAsync(nl, ix->bank_lo, ob, cudaMemcpyDeviceToDevice, s));
    }
  }
  pool_free(ix->bank, s);      // stream-ordered after the copies above
  pool_free(ix->rnorm, s);
  pool_free(ix->bank_hi, s);
  pool_free(ix->bank_lo, s);
  ix->bank = nb;
  ix->rnorm = nn;
  ix->bank_hi = nh;
  ix->bank_lo = nl;
  ix->capacity = cap;
  return VSCB200_OK;
}

// the bank's running norm maxima (margins of the one-pass searches): allocated at the first add, zeroed after a reset
int norm_maxima_ready(vscb200_index* ix, cudaStream_t s) {
  int rc;
  if (!ix->rmax2_bits && (rc = pool_alloc(reinterpret_cast<void**>(&ix->rmax2_bits), 2 * sizeof(unsigned int), s))) return rc;
  if (ix->rmax2_reset) {
    VSCB_CUDA_OK(cudaMemsetAsync(ix->rmax2_bits, 0, 2 * sizeof(unsigned int), s));
    ix->rmax2_reset = false;
  }
  return VSCB200_OK;
}

int append_rows(vscb200_index* ix, const float* x, int64_t n, cudaMemcpyKind kind, cudaStream_t s) {
  if (n == 0) return VSCB200_OK;
  int rc = ensure_capacity(ix, ix->ntotal + n, s);
  if (rc) return rc;
  float* dst = ix->bank + ix->ntotal * ix->d;
  VSCB_CUDA_OK(cudaMemcpyAsync(dst, x, static_cast<size_t>(n) * ix->d * sizeof(float), kind, s));
  if ((rc = norm_maxima_ready(ix, s))) return rc;
  if (ix->force_simt) {
    rc = row_sqnorm(dst, n, ix->d, ix->rnorm + ix->ntotal, s);   // L2 transform + range-search error margins
  } else {
    // one pass over the new rows: squared norms (the same summation order as row_sqnorm), both bf16 operand planes, and
    // the running maxima of |r|^2 and |r - bf16(r)|^2 (margin of the one-pass searches)
    rc = q_hi_norm(dst, ix->bank_hi + ix->ntotal * ix->dp, ix->rnorm + ix->ntotal, nullptr, n, ix->d, ix->dp, s,
                   ix->bank_lo + ix->ntotal * ix->dp, ix->rmax2_bits);
  }
  if (rc) return rc;
  ix->ntotal += n;
  return VSCB200_OK;
}

// ---- host rows on their way to the device ------------------------------------------------------------------------
// The reference adds one video (~30 rows) at a time (vsc/index.py:87-94).  Rows are packed into one of two process-wide
// page-locked 8 MB slots; a full slot (or the next search / add on this index) is uploaded asynchronously on the index's
// own stream while the host fills the other slot -- no growing host vector, no pageable copies, no launch per video.
namespace {
constexpr int kStageMaxDevices = 16;
struct StageSlot {
  float* p = nullptr;
  cudaEvent_t ev[kStageMaxDevices] = {};     // per device: events belong to the device of the stream they are recorded on
  int last_dev = -1;                         // device whose upload used the slot last
};
constexpr size_t kStageBytes = 8u << 20;
std::mutex g_stage_mu;
StageSlot g_stage[2];
int g_stage_cur = 0;
vscb200_index* g_stage_owner = nullptr;      // the index whose rows sit in the current slot
int64_t g_stage_rows = 0;

int stage_init() {
  for (StageSlot& sl : g_stage)
    if (!sl.p) VSCB_CUDA_OK(cudaHostAlloc(reinterpret_cast<void**>(&sl.p), kStageBytes, cudaHostAllocPortable));
  return VSCB200_OK;
}

// all work on an index is ordered, whatever stream each call arrives on
int order_stream(vscb200_index* ix, cudaStream_t s) {
  if (ix->has_stream && ix->last_stream != s) {
    if (!ix->order_ev) VSCB_CUDA_OK(cudaEventCreateWithFlags(&ix->order_ev, cudaEventDisableTiming));
    VSCB_CUDA_OK(cudaEventRecord(ix->order_ev, ix->last_stream));
    VSCB_CUDA_OK(cudaStreamWaitEvent(s, ix->order_ev, 0));
  }
  ix->last_stream = s;
  ix->has_stream = true;
  return VSCB200_OK;
}

// g_stage_mu held: upload the rows staged for `ix` on its own stream and move on to the other slot
int stage_submit(vscb200_index* ix) {
  if (g_stage_owner != ix || g_stage_rows == 0) return VSCB200_OK;
  DeviceGuard guard(ix->device);
  cudaStream_t s;
  int rc = own_stream(ix, &s);
  if (rc) return rc;
  if ((rc = order_stream(ix, s))) return rc;
  StageSlot& sl = g_stage[g_stage_cur];
  const int64_t rows = g_stage_rows;
  g_stage_rows = 0;
  g_stage_owner = nullptr;
  g_stage_cur ^= 1;
  if ((rc = append_rows(ix, sl.p, rows, cudaMemcpyHostToDevice, s))) return rc;
  const int dev = ix->device;
  if (!sl.ev[dev]) VSCB_CUDA_OK(cudaEventCreateWithFlags(&sl.ev[dev], cudaEventDisableTiming));
  VSCB_CUDA_OK(cudaEventRecord(sl.ev[dev], s));   // the slot is free again once this copy has run
  sl.last_dev = dev;
  return VSCB200_OK;
}
}  // namespace

int flush_pending(vscb200_index* ix, cudaStream_t s) {
  {
    std::lock_guard<std::mutex> lk(g_stage_mu);
    int rc = stage_submit(ix);
    if (rc) return rc;
  }
  return order_stream(ix, s);
}

int64_t block_rows(const vscb200_index* ix, int64_t nq) {
  const size_t per_row = static_cast<size_t>(std::max<int64_t>(ix->ntotal, 1)) * sizeof(float);
  int64_t rows = static_cast<int64_t>(ws_budget_bytes() / per_row);
  rows = std::max<int64_t>(rows, 1);
  rows = std::min<int64_t>(rows, 65535 * 64);
  return std::min(rows, std::max<int64_t>(nq, 1));
}

// S[0:nq, 0:ntotal] for one query block
int score_block(vscb200_index* ix, const float* q, int64_t nq, float* S, int64_t ldS, cudaStream_t s) {
  const bool l2 = ix->metric == VSCB200_METRIC_L2;
  const float* qn = nullptr;
  {
    int rc = grow(&ix->qnorm, &ix->qnorm_bytes, static_cast<size_t>(nq) * sizeof(float), s);
    if (rc) return rc;
    rc = row_sqnorm(q, nq, ix->d, ix->qnorm, s);
    if (rc) return rc;
    qn = ix->qnorm;
  }
  if (!ix->force_simt) {
    const size_t plane = static_cast<size_t>(nq) * ix->dp;
    int rc = grow(&ix->q_planes, &ix->q_planes_bytes, 2 * plane * sizeof(uint16_t), s);
    if (rc) return rc;
    if ((rc = split_planes(q, ix->q_planes, ix->q_planes + plane, nq, ix->d, ix->dp, s))) return rc;
    return scores_tc_planes(ix->q_planes, ix->q_planes + plane, ix->bank_hi, ix->bank_lo, S, nq, ix->ntotal, ix->dp,
                            ldS, l2, qn, ix->rnorm, s);
  }
  return scores_simt(q, ix->bank, S, nq, ix->ntotal, ix->d, ldS, l2, qn, ix->rnorm, s);
}

int own_stream(vscb200_index* ix, cudaStream_t* s) {
  if (!ix->own_stream) VSCB_CUDA_OK(cudaStreamCreateWithFlags(&ix->own_stream, cudaStreamNonBlocking));
  *s = ix->own_stream;
  return VSCB200_OK;
}

}  // namespace vscb200

extern "C" {

int vscb200_index_create(int d, int metric, vscb200_index** out) {
  VSCB_REQUIRE(out != nullptr, "index_create: null out");
  VSCB_REQUIRE(d > 0, "index_create: d must be positive");
  VSCB_REQUIRE(metric == VSCB200_METRIC_INNER_PRODUCT || metric == VSCB200_METRIC_L2,
               "index_create: metric must be METRIC_INNER_PRODUCT or METRIC_L2");
  vscb200_index* ix = new vscb200_index();
  ix->d = d;
  cudaGetDevice(&ix->device);
  VSCB_REQUIRE(ix->device >= 0 && ix->device < 16, "index_create: device ordinal out of range");
  ix->dp = (d + 7) & ~7;
  ix->metric = metric;
  const char* e = getenv("VSCB200_FORCE_SIMT");
  ix->force_simt = (e && atoi(e)) ? 1 : 0;
  const char* nf = getenv("VSCB200_NO_FUSED_TOPK");
  ix->no_fused = (nf && atoi(nf)) ? 1 : 0;
  const char* ns = getenv("VSCB200_NO_STREAM_SEARCH");
  ix->no_stream = (ns && atoi(ns)) ? 1 : 0;
  const char* sp = getenv("VSCB200_SIM_PASSES");     // 3: the split-bf16 fused top-k (cross-validation / A-B timing)
  ix->sim_passes = (sp && atoi(sp) == 3) ? 3 : 1;
  *out = ix;
  return VSCB200_OK;
}

void vscb200_index_destroy(vscb200_index* ix) {
  if (!ix) return;
  DeviceGuard guard(ix->device);          // may run from a finaliser while another device is current
  {
    std::lock_guard<std::mutex> lk(g_stage_mu);
    if (g_stage_owner == ix) { g_stage_owner = nullptr; g_stage_rows = 0; }      // rows never searched: dropped
  }
  // the blocks go back to the library pool, ordered after the last stream this index worked on
  cudaStream_t s = ix->last_stream;
  void* blocks[] = {ix->rmax2_bits, ix->flags, ix->bank, ix->rnorm, ix->ws, ix->q_stage, ix->D_stage, ix->I_stage, ix->qnorm, ix->counts,
                    ix->bank_hi, ix->bank_lo, ix->q_planes, ix->Dtmp, ix->Itmp, ix->cand_d, ix->cand_i, ix->gmax, ix->gr_keys, ix->gr_count,
                    ix->g_score, ix->g_q, ix->g_r, ix->vp_score, ix->vp_q, ix->vp_r};
  if (ix->own_stream && s == ix->own_stream) {
    cudaStreamSynchronize(s);     // own stream is destroyed below: drain it, then free un-ordered
    s = nullptr;
  }
  for (void* b : blocks) pool_free(b, s);
  if (ix->own_stream) cudaStreamDestroy(ix->own_stream);
  if (ix->order_ev) cudaEventDestroy(ix->order_ev);
  delete ix;
}

int vscb200_index_add(vscb200_index* ix, const float* x_dev, int64_t n, void* stream) {
  VSCB_REQUIRE(ix && (n == 0 || x_dev), "index_add: null argument");
  VSCB_REQUIRE(n >= 0, "index_add: negative row count");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc = flush_pending(ix, s);
  if (rc) return rc;
  return append_rows(ix, x_dev, n, cudaMemcpyDeviceToDevice, s);
}

/* index.add(sn_transform(x)) in one pass: the score-normalisation transform of raw rows (score_normalization.py:73-83,
 * 96-101: drop a column, L2-normalise, append `fill` or bias[row]) written straight into the index's storage.  Stored values
 * are bit-identical to vscb200_sn_transform followed by vscb200_index_add.  d_in: columns of x; the index dimension must be
 * d_in when a column is dropped (drop_dim_dev != NULL, or 0 <= drop_dim < d_in), else d_in + 1. */
int vscb200_index_add_sn(vscb200_index* ix, const float* x_dev, int64_t n, int d_in, int drop_dim, const int* drop_dim_dev,
                         int l2_normalize, float fill, const float* bias_dev, void* stream) {
  VSCB_REQUIRE(ix && (n == 0 || x_dev), "index_add_sn: null argument");
  VSCB_REQUIRE(n >= 0 && d_in > 0 && drop_dim < d_in, "index_add_sn: bad shape");
  const bool drops = drop_dim_dev != nullptr || drop_dim >= 0;
  VSCB_REQUIRE(ix->d == (drops ? d_in : d_in + 1), "index_add_sn: index dimension does not match the transformed rows");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc = flush_pending(ix, s);
  if (rc || n == 0) return rc;
  if ((rc = ensure_capacity(ix, ix->ntotal + n, s))) return rc;
  float* dst = ix->bank + ix->ntotal * ix->d;
  if (ix->force_simt) {
    if ((rc = sn_transform(x_dev, n, d_in, drop_dim, l2_normalize, fill, bias_dev, dst, s, drop_dim_dev))) return rc;
    rc = row_sqnorm(dst, n, ix->d, ix->rnorm + ix->ntotal, s);
  } else {
    if ((rc = norm_maxima_ready(ix, s))) return rc;
    rc = sn_add_rows(x_dev, n, d_in, drop_dim, l2_normalize, fill, bias_dev, drop_dim_dev, dst, ix->bank_hi + ix->ntotal * ix->dp,
                     ix->bank_lo + ix->ntotal * ix->dp, ix->rnorm + ix->ntotal, ix->rmax2_bits, ix->d, ix->dp, s);
  }
  if (rc) return rc;
  ix->ntotal += n;
  return VSCB200_OK;
}

int vscb200_index_add_host(vscb200_index* ix, const float* x_host, int64_t n) {
  VSCB_REQUIRE(ix && (n == 0 || x_host), "index_add_host: null argument");
  VSCB_REQUIRE(n >= 0, "index_add_host: negative row count");
  const size_t row_bytes = static_cast<size_t>(ix->d) * sizeof(float);
  VSCB_REQUIRE(row_bytes <= kStageBytes, "index_add_host: descriptor dimension too large");
  std::lock_guard<std::mutex> lk(g_stage_mu);
  int rc = stage_init();
  if (rc) return rc;
  if (g_stage_owner && g_stage_owner != ix && (rc = stage_submit(g_stage_owner))) return rc;
  const int64_t cap_rows = static_cast<int64_t>(kStageBytes / row_bytes);
  while (n > 0) {
    StageSlot& sl = g_stage[g_stage_cur];
    if (g_stage_rows == 0) {
      if (sl.last_dev >= 0) VSCB_CUDA_OK(cudaEventSynchronize(sl.ev[sl.last_dev]));   // its previous upload has left the host buffer
      g_stage_owner = ix;
    }
    const int64_t take = std::min(n, cap_rows - g_stage_rows);
    memcpy(sl.p + g_stage_rows * ix->d, x_host, static_cast<size_t>(take) * row_bytes);
    g_stage_rows += take;
    x_host += take * ix->d;
    n -= take;
    if (g_stage_rows == cap_rows && (rc = stage_submit(ix))) return rc;
  }
  return VSCB200_OK;
}

int vscb200_index_reset(vscb200_index* ix) {
  VSCB_REQUIRE(ix, "index_reset: null index");
  std::lock_guard<std::mutex> lk(g_stage_mu);
  if (g_stage_owner == ix) { g_stage_owner = nullptr; g_stage_rows = 0; }
  ix->ntotal = 0;
  ix->rmax2_reset = true;
  return VSCB200_OK;
}

int64_t vscb200_index_ntotal(const vscb200_index* ix) {
  if (!ix) return 0;
  std::lock_guard<std::mutex> lk(g_stage_mu);
  return ix->ntotal + (g_stage_owner == ix ? g_stage_rows : 0);
}
int vscb200_index_dim(const vscb200_index* ix) { return ix ? ix->d : 0; }
int vscb200_index_metric(const vscb200_index* ix) { return ix ? ix->metric : 0; }
int vscb200_index_set_id_offset(vscb200_index* ix, int64_t id_offset) {
  VSCB_REQUIRE(ix, "index_set_id_offset: null index");
  ix->id_offset = id_offset;
  return VSCB200_OK;
}

int vscb200_index_search(vscb200_index* ix, const float* q, int64_t nq, int k, float* D, int64_t* I, void* stream) {
  VSCB_REQUIRE(ix && (nq == 0 || (q && D && I)), "index_search: null argument");
  VSCB_REQUIRE(nq >= 0, "index_search: negative query count");
  VSCB_REQUIRE(k >= 1 && k <= 2048, "index_search: k must be in [1, 2048]");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc = flush_pending(ix, s);
  if (rc) return rc;
  const bool keep_max = ix->metric == VSCB200_METRIC_INNER_PRODUCT;
  // A few query rows against a large bank (the reference's per-video call pattern): the bank is streamed from
  // HBM once, per-group maxima select k + slack groups of 32 rows per query, which are rescored exactly.
  const int kStreamSlack = 6;
  if (!ix->force_simt && !ix->no_stream && nq >= 1 && nq <= 128 && k + kStreamSlack <= 32 && ix->ntotal >= 32768 &&
      ix->ntotal < (1ll << 31) - 256) {
    const int Npad = static_cast<int>((nq + 15) & ~15ll);
    const int64_t G = (ix->ntotal + 31) / 32;
    const int kg = static_cast<int>(std::min<int64_t>(k + kStreamSlack, G));
    if ((rc = grow(&ix->gmax, &ix->gmax_bytes, static_cast<size_t>(G) * Npad * sizeof(float), s))) return rc;
    if ((rc = grow(&ix->qnorm, &ix->qnorm_bytes, static_cast<size_t>(nq) * sizeof(float), s))) return rc;
    const size_t plane = static_cast<size_t>(nq) * ix->dp;
    if ((rc = grow(&ix->q_planes, &ix->q_planes_bytes, 2 * plane * sizeof(uint16_t), s))) return rc;
    // norms and both operand planes of the few query rows in ONE launch (the call is launch-latency sensitive)
    if ((rc = q_hi_norm(q, ix->q_planes, ix->qnorm, nullptr, nq, ix->d, ix->dp, s, ix->q_planes + plane))) return rc;
    if ((rc = sim_stream_groupmax(ix->q_planes, ix->q_planes + plane, ix->bank_hi, ix->bank_lo, nq, ix->ntotal, ix->dp,
                                  !keep_max, ix->qnorm, ix->rnorm, ix->gmax, Npad, s))) return rc;
    const int chunks = group_topk_chunks(G);
    if (chunks <= 256) {
      // per-chunk top groups (one CTA per 256 groups), merged inside the rescoring kernel
      const size_t ncg = static_cast<size_t>(chunks) * kg;
      if ((rc = grow(&ix->cand_d, &ix->cand_d_bytes, static_cast<size_t>(nq) * ncg * sizeof(float), s))) return rc;
      if ((rc = grow(&ix->cand_i, &ix->cand_i_bytes, static_cast<size_t>(nq) * ncg * sizeof(int32_t), s))) return rc;
      if ((rc = group_topk(ix->gmax, G, Npad, static_cast<int>(nq), kg, chunks, ix->cand_d, ix->cand_i, s))) return rc;
      if (!ix->gr_count) {
        if ((rc = pool_alloc(reinterpret_cast<void**>(&ix->gr_count), 128 * sizeof(int), s))) return rc;
        VSCB_CUDA_OK(cudaMemsetAsync(ix->gr_count, 0, 128 * sizeof(int), s));
      }
      if ((rc = grow(&ix->gr_keys, &ix->gr_keys_bytes, group_rescore_few_key_bytes(nq, kg, 32), s))) return rc;
      return group_rescore_few(q, ix->bank, ix->d, !keep_max, ix->ntotal, ix->cand_d, ix->cand_i, static_cast<int>(ncg), kg, nq, k,
                               D, I, ix->id_offset, ix->gr_keys, ix->gr_count, s);
    }
    if ((rc = grow(&ix->Dtmp, &ix->Dtmp_bytes, static_cast<size_t>(nq) * kg * sizeof(float), s))) return rc;
    if ((rc = grow(&ix->Itmp, &ix->Itmp_bytes, static_cast<size_t>(nq) * kg * sizeof(int64_t), s))) return rc;
    if ((rc = topk_rows(ix->gmax, 1, nq, G, kg, true, ix->Dtmp, ix->Itmp, 0, s, Npad))) return rc;
    return group_rescore(q, ix->bank, ix->d, !keep_max, ix->ntotal, ix->Itmp, nullptr, nullptr, 0, kg, nq, k, D, I,
                         ix->id_offset, s);
  }
  // Small k on a large bank, single pass: approximate scores from ONE bf16 MMA per product select, per query, every
  // bank row within a proven error margin of the k-th best; the survivors are rescored in exact fp32 (sim_tc1.cu).
  if (!ix->force_simt && !ix->no_fused && ix->sim_passes == 1 && k <= sim1_max_k() && ix->ntotal >= 2048 &&
      ix->ntotal < (1ll << 31) && ix->d % 4 == 0) {
    const int64_t qblk = 1 << 18;
    for (int64_t q0 = 0; q0 < nq; q0 += qblk) {
      const int64_t nb = std::min<int64_t>(qblk, nq - q0);
      if ((rc = grow(&ix->cand_d, &ix->cand_d_bytes, static_cast<size_t>(nb) * sim1_list_cap() * 8, s))) return rc;
      if ((rc = grow(&ix->qnorm, &ix->qnorm_bytes, static_cast<size_t>(2 * nb) * sizeof(float), s))) return rc;
      const size_t scratch_bytes = static_cast<size_t>(sim1_scratch_ints(nb)) * sizeof(int);
      if ((rc = grow(&ix->flags, &ix->flags_bytes, scratch_bytes, s))) return rc;
      if ((rc = grow(&ix->cand_i, &ix->cand_i_bytes, sim1_partial_bytes(), s))) return rc;
      const float* qb = q + q0 * ix->d;
      const size_t plane = static_cast<size_t>(nb) * ix->dp;
      if ((rc = grow(&ix->q_planes, &ix->q_planes_bytes, 2 * plane * sizeof(uint16_t), s))) return rc;
      if ((rc = q_hi_norm(qb, ix->q_planes, ix->qnorm, ix->qnorm + nb, nb, ix->d, ix->dp, s))) return rc;
      VSCB_CUDA_OK(cudaMemsetAsync(ix->flags, 0, scratch_bytes, s));
      // Threshold bootstrap (inner product): one-pass scores against every stride-th bank row (<= 1024 of them, a strided
      // tensor map: nothing is gathered), the k-th best of which bounds each query's k-th best from below
      // (sim1_boot_tau).  Measured at 10k x 40k: k = 10 0.617 -> 0.607 ms, k = 1 0.306 -> 0.323 ms (the two launches cost
      // more than a k = 1 list saves), hence k >= 4 only.  VSCB200_SIM1_BOOT=0 starts every list cold.
      static const int boot_on = [] { const char* e = getenv("VSCB200_SIM1_BOOT"); return e ? atoi(e) : 1; }();
      const int64_t stride = std::max<int64_t>(64, (ix->ntotal + 1023) / 1024);
      const int64_t ncs = std::min<int64_t>((ix->ntotal / stride) & ~3ll, 1024);
      if (boot_on && keep_max && k >= 4 && ncs >= 64 && ncs >= 4 * k) {
        if ((rc = grow(&ix->ws, &ix->ws_bytes, static_cast<size_t>(nb) * ncs * sizeof(float), s))) return rc;
        if ((rc = scores_tc_planes(ix->q_planes, nullptr, ix->bank_hi, nullptr, ix->ws, nb, ncs, ix->dp, ncs, false, nullptr, nullptr,
                                   s, 1, stride))) return rc;
        if ((rc = sim1_boot_tau(ix->ws, nb, static_cast<int>(ncs), k, ix->qnorm, ix->qnorm + nb, ix->rmax2_bits, ix->d, ix->flags,
                                s))) return rc;
      }
      if ((rc = sim1_topk(ix->q_planes, ix->bank_hi, nb, ix->ntotal, ix->d, ix->dp, !keep_max, k, ix->qnorm, ix->qnorm + nb, ix->rnorm,
                          ix->rmax2_bits, ix->cand_d, ix->flags, s))) return rc;
      if ((rc = sim1_rescore(qb, ix->bank, ix->d, !keep_max, nb, ix->ntotal, ix->cand_d, k, ix->qnorm, ix->qnorm + nb,
                             ix->rmax2_bits, D + q0 * k, I + q0 * k, ix->id_offset, ix->flags, ix->cand_i, s))) return rc;
      ix->last_flag_count = ix->flags + nb;
    }
    return VSCB200_OK;
  }
  // The same with fp32-equivalent scores (three MMAs per product), kept for cross-validation (VSCB200_SIM_PASSES=3).
  const int kFusedSlack = 6;
  if (!ix->force_simt && !ix->no_fused && k + kFusedSlack <= fused_topk_list_len() && ix->ntotal >= 2048 &&
      ix->ntotal < (1ll << 31)) {
    const int fk = fused_topk_list_len();
    for (int64_t q0 = 0; q0 < nq; q0 += (1 << 20)) {
      const int64_t nb = std::min<int64_t>(1 << 20, nq - q0);
      const int slabs = fused_topk_slabs(nb, ix->ntotal);
      const int ncand = slabs * 2 * fk;
      if ((rc = grow(&ix->cand_d, &ix->cand_d_bytes, static_cast<size_t>(nb) * ncand * sizeof(float), s))) return rc;
      if ((rc = grow(&ix->cand_i, &ix->cand_i_bytes, static_cast<size_t>(nb) * ncand * sizeof(int32_t), s))) return rc;
      if ((rc = grow(&ix->qnorm, &ix->qnorm_bytes, static_cast<size_t>(nb) * sizeof(float), s))) return rc;
      const float* qb = q + q0 * ix->d;
      if ((rc = row_sqnorm(qb, nb, ix->d, ix->qnorm, s))) return rc;
      const size_t plane = static_cast<size_t>(nb) * ix->dp;
      if ((rc = grow(&ix->q_planes, &ix->q_planes_bytes, 2 * plane * sizeof(uint16_t), s))) return rc;
      if ((rc = split_planes(qb, ix->q_planes, ix->q_planes + plane, nb, ix->d, ix->dp, s))) return rc;
      if ((rc = topk_tc_fused(ix->q_planes, ix->q_planes + plane, ix->bank_hi, ix->bank_lo, nb, ix->ntotal, ix->dp,
                              !keep_max, ix->qnorm, ix->rnorm, slabs, ix->cand_d, ix->cand_i, s))) return rc;
      if ((rc = group_rescore(qb, ix->bank, ix->d, !keep_max, ix->ntotal, nullptr, ix->cand_d, ix->cand_i, ncand,
                              k + kFusedSlack, nb, k, D + q0 * k, I + q0 * k, ix->id_offset, s,
                              fused_topk_group_rows()))) return rc;
    }
    return VSCB200_OK;
  }
  const int64_t blk = block_rows(ix, nq);
  const int64_t ldS = (ix->ntotal + 3) & ~3ll;
  rc = grow(&ix->ws, &ix->ws_bytes, static_cast<size_t>(blk) * std::max<int64_t>(ldS, 4) * sizeof(float), s);
  if (rc) return rc;
  // Tensor-core scores are fp32-equivalent to ~1e-6: keep k + slack survivors, rescore them exactly.
  const int kRescoreSlack = 8;
  const bool rescore = !ix->force_simt && ix->ntotal > 0;
  const int kin = rescore ? static_cast<int>(std::min<int64_t>(std::min<int64_t>(ix->ntotal, 2048), k + kRescoreSlack)) : k;
  if (rescore) {
    if ((rc = grow(&ix->Dtmp, &ix->Dtmp_bytes, static_cast<size_t>(blk) * kin * sizeof(float), s))) return rc;
    if ((rc = grow(&ix->Itmp, &ix->Itmp_bytes, static_cast<size_t>(blk) * kin * sizeof(int64_t), s))) return rc;
  }
  for (int64_t q0 = 0; q0 < nq; q0 += blk) {
    const int64_t nb = std::min(blk, nq - q0);
    if (ix->ntotal > 0) {
      rc = score_block(ix, q + q0 * ix->d, nb, ix->ws, ldS, s);
      if (rc) return rc;
    }
    if (rescore) {
      if ((rc = topk_rows(ix->ws, ldS, nb, ix->ntotal, kin, keep_max, ix->Dtmp, ix->Itmp, 0, s))) return rc;
      if ((rc = rescore_sort(q + q0 * ix->d, ix->bank, ix->d, !keep_max, ix->Itmp, kin, nb, k, D + q0 * k, I + q0 * k,
                             ix->id_offset, s))) return rc;
    } else {
      rc = topk_rows(ix->ws, ldS, nb, ix->ntotal, k, keep_max, D + q0 * k, I + q0 * k, ix->id_offset, s);
      if (rc) return rc;
    }
  }
  return VSCB200_OK;
}

/* index.reconstruct_n(i0, n): rows [i0, i0 + n) of the bank as float32.  `out` may be host memory or memory of ANY device
 * (unified addressing): this is how faiss_compat.index_cpu_to_all_gpus distributes a bank over the GPUs of one process.
 * The copy runs on the index's own stream (on the index's device) and has completed when the call returns. */
int vscb200_index_reconstruct_n(vscb200_index* ix, int64_t i0, int64_t n, float* out) {
  VSCB_REQUIRE(ix && (n == 0 || out), "index_reconstruct_n: null argument");
  DeviceGuard guard(ix->device);
  cudaStream_t s;
  int rc = own_stream(ix, &s);
  if (rc) return rc;
  if ((rc = flush_pending(ix, s))) return rc;
  VSCB_REQUIRE(i0 >= 0 && n >= 0 && i0 + n <= ix->ntotal, "index_reconstruct_n: row range out of bounds");
  if (n == 0) return VSCB200_OK;
  VSCB_CUDA_OK(cudaMemcpyAsync(out, ix->bank + i0 * ix->d, static_cast<size_t>(n) * ix->d * sizeof(float), cudaMemcpyDefault, s));
  VSCB_CUDA_OK(cudaStreamSynchronize(s));
  return VSCB200_OK;
}

int64_t vscb200_index_last_fallbacks(vscb200_index* ix) {
  // diagnostic: queries of the last single-pass search (its last 2^20-row block) that took the exhaustive fallback
  if (!ix || !ix->last_flag_count) return -1;
  int n = 0;
  if (cudaStreamSynchronize(ix->last_stream) != cudaSuccess ||
      cudaMemcpy(&n, ix->last_flag_count, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  return n;
}

int vscb200_index_search_host(vscb200_index* ix, const float* q_host, int64_t nq, int k, float* D_host,
                              int64_t* I_host) {
  VSCB_REQUIRE(ix && (nq == 0 || (q_host && D_host && I_host)), "index_search_host: null argument");
  if (nq == 0) return VSCB200_OK;
  DeviceGuard guard(ix->device);
  cudaStream_t s;
  int rc = own_stream(ix, &s);
  if (rc) return rc;
  if ((rc = grow(&ix->q_stage, &ix->q_stage_bytes, static_cast<size_t>(nq) * ix->d * sizeof(float), s))) return rc;
  if ((rc = grow(&ix->D_stage, &ix->D_stage_bytes, static_cast<size_t>(nq) * k * sizeof(float), s))) return rc;
  if ((rc = grow(&ix->I_stage, &ix->I_stage_bytes, static_cast<size_t>(nq) * k * sizeof(int64_t), s))) return rc;
  VSCB_CUDA_OK(cudaMemcpyAsync(ix->q_stage, q_host, static_cast<size_t>(nq) * ix->d * sizeof(float),
                               cudaMemcpyHostToDevice, s));
  rc = vscb200_index_search(ix, ix->q_stage, nq, k, ix->D_stage, ix->I_stage, s);
  if (rc) return rc;
  VSCB_CUDA_OK(cudaMemcpyAsync(D_host, ix->D_stage, static_cast<size_t>(nq) * k * sizeof(float),
                               cudaMemcpyDeviceToHost, s));
  VSCB_CUDA_OK(cudaMemcpyAsync(I_host, ix->I_stage, static_cast<size_t>(nq) * k * sizeof(int64_t),
                               cudaMemcpyDeviceToHost, s));
  VSCB_CUDA_OK(cudaStreamSynchronize(s));
  return VSCB200_OK;
}

int vscb200_index_scores(vscb200_index* ix, const float* q, int64_t nq, float* S, int64_t ldS, void* stream) {
  VSCB_REQUIRE(ix && (nq == 0 || (q && S)), "index_scores: null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc = flush_pending(ix, s);
  if (rc) return rc;
  VSCB_REQUIRE(ldS >= ix->ntotal, "index_scores: ldS < ntotal");
  if (nq == 0 || ix->ntotal == 0) return VSCB200_OK;
  for (int64_t q0 = 0; q0 < nq; q0 += 65535ll * 64) {
    const int64_t nb = std::min<int64_t>(65535ll * 64, nq - q0);
    rc = score_block(ix, q + q0 * ix->d, nb, S + q0 * ldS, ldS, s);
    if (rc) return rc;
  }
  return VSCB200_OK;
}

int vscb200_index_range_search_host(vscb200_index* ix, const float* q_host, int64_t nq, float thresh,
                                    uint64_t* lims_host, float** D_out, int64_t** I_out) {
  VSCB_REQUIRE(ix && lims_host && D_out && I_out && (nq == 0 || q_host), "index_range_search_host: null argument");
  *D_out = nullptr;
  *I_out = nullptr;
  lims_host[0] = 0;
  DeviceGuard guard(ix->device);
  cudaStream_t s;
  int rc = own_stream(ix, &s);
  if (rc) return rc;
  if ((rc = flush_pending(ix, s))) return rc;
  const bool keep_max = ix->metric == VSCB200_METRIC_INNER_PRODUCT;
  const int64_t blk = block_rows(ix, nq);
  const int64_t ldS = (ix->ntotal + 3) & ~3ll;
  if ((rc = grow(&ix->ws, &ix->ws_bytes, static_cast<size_t>(blk) * std::max<int64_t>(ldS, 4) * sizeof(float), s))) return rc;
  if ((rc = grow(&ix->q_stage, &ix->q_stage_bytes, static_cast<size_t>(std::max<int64_t>(blk, 1)) * ix->d * sizeof(float), s))) return rc;
  if ((rc = grow(&ix->counts, &ix->counts_bytes, static_cast<size_t>(std::max<int64_t>(blk, 1)) * sizeof(unsigned long long), s))) return rc;
  std::vector<unsigned long long> cnt(static_cast<size_t>(std::max<int64_t>(blk, 1)));
  size_t total = 0, cap = 0;
  float* Dh = nullptr;
  int64_t* Ih = nullptr;
  float* Dd = nullptr; size_t Dd_bytes = 0;
  int64_t* Id = nullptr; size_t Id_bytes = 0;
  auto fail = [&](int code) {
    free(Dh); free(Ih); pool_free(Dd, s); pool_free(Id, s);
    return code;
  };
  for (int64_t q0 = 0; q0 < nq; q0 += blk) {
    const int64_t nb = std::min(blk, nq - q0);
    if (cudaMemcpyAsync(ix->q_stage, q_host + q0 * ix->d, static_cast<size_t>(nb) * ix->d * sizeof(float),
                        cudaMemcpyHostToDevice, s) != cudaSuccess) { set_last_error("range_search: H2D failed"); return fail(VSCB200_ERR_CUDA); }
    if (ix->ntotal > 0 && (rc = score_block(ix, ix->q_stage, nb, ix->ws, ldS, s))) return fail(rc);
    const bool exact = !ix->force_simt && ix->ntotal > 0;
    const float* Qx = exact ? ix->q_stage : nullptr;
    if ((rc = range_count(ix->ws, ldS, nb, ix->ntotal, thresh, keep_max, ix->counts, s, Qx, ix->bank, ix->d, ix->qnorm,
                          ix->rnorm))) return fail(rc);
    if (cudaMemcpyAsync(cnt.data(), ix->counts, static_cast<size_t>(nb) * sizeof(unsigned long long),
                        cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess) { set_last_error(std::string("range_search: count failed: ") + cudaGetErrorString(cudaGetLastError())); return fail(VSCB200_ERR_CUDA); }
    // exclusive scan on the host (block of <= a few thousand rows), offsets back to the device
    size_t blk_total = 0;
    for (int64_t i = 0; i < nb; ++i) {
      const unsigned long long c = cnt[i];
      cnt[i] = blk_total;
      blk_total += c;
      lims_host[q0 + i + 1] = total + blk_total;
    }
    if (blk_total == 0) continue;
    if (total + blk_total > cap) {
      cap = std::max(total + blk_total, cap * 2);
      float* nD = static_cast<float*>(realloc(Dh, cap * sizeof(float)));
      int64_t* nI = static_cast<int64_t*>(realloc(Ih, cap * sizeof(int64_t)));
      if (!nD || !nI) { Dh = nD ? nD : Dh; Ih = nI ? nI : Ih; set_last_error("range_search: host realloc failed"); return fail(VSCB200_ERR_NOMEM); }
      Dh = nD; Ih = nI;
    }
    if ((rc = grow(&Dd, &Dd_bytes, blk_total * sizeof(float), s))) return fail(rc);
    if ((rc = grow(&Id, &Id_bytes, blk_total * sizeof(int64_t), s))) return fail(rc);
    if (cudaMemcpyAsync(ix->counts, cnt.data(), static_cast<size_t>(nb) * sizeof(unsigned long long),
                        cudaMemcpyHostToDevice, s) != cudaSuccess) { set_last_error("range_search: H2D offsets failed"); return fail(VSCB200_ERR_CUDA); }
    if ((rc = range_fill(ix->ws, ldS, nb, ix->ntotal, thresh, keep_max, ix->counts, Dd, Id, ix->id_offset, s, Qx, ix->bank,
                         ix->d, ix->qnorm, ix->rnorm))) return fail(rc);
    if (cudaMemcpyAsync(Dh + total, Dd, blk_total * sizeof(float), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaMemcpyAsync(Ih + total, Id, blk_total * sizeof(int64_t), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess) { set_last_error(std::string("range_search: fill failed: ") + cudaGetErrorString(cudaGetLastError())); return fail(VSCB200_ERR_CUDA); }
    total += blk_total;
  }
  pool_free(Dd, s);
  pool_free(Id, s);
  if (!Dh) { Dh = static_cast<float*>(malloc(sizeof(float))); Ih = static_cast<int64_t*>(malloc(sizeof(int64_t))); }
  *D_out = Dh;
  *I_out = Ih;
  return VSCB200_OK;
}

}  // extern "C"
