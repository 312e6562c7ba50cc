// Temporal-network alignment of candidate video pairs on the device (second half).
//
// Replaces `vcsl.vta.tn` (VSC22-Descriptor-Track-1st/infer/vcsl/vta.py:244-363) as driven by
// vsc/baseline/localization.py:38-76 (`VCSLLocalization.localize_all` -> `TnVtaModel.forward_sim`, a 16-process
// multiprocessing pool running networkx on one pair each).  Input: the per-row top-k of every pair's frame-similarity
// matrix (vscb200_pair_topk); output: up to max_path + 1 boxes [q_min, r_min, q_max, r_max] per pair.
//
// One WARP per pair (grid-stride), graph kept implicit:
//   * node (q, k) = query frame q with its k-th best reference frame; edges only span < tn_max_step query frames, so the
//     edge set is a byte per (q_i, q_j - q_i, k_j, k_i): 0 none, 1 weight = sim of the destination, 2 re-weighted to 0.
//     Lanes build the rows of different q_i in parallel (constraints C2-C4; C3 needs the running set of reference frames
//     already linked from q_i, which is sequential in q_j only).
//   * longest path = DP over q (all edges go forward in q); for one node the lanes take one predecessor each and reduce
//     with "first maximum in insertion order" -- networkx's `max(us, key=...)` over G.pred[v], whose order is
//     (q_i ascending, k_i ascending), then the sink links (vta.py:317-322) in node order.  Ties are the COMMON case at
//     the start of a chain (all edges into a node carry the same weight), so this order is part of the result.
//   * up to max_path + 1 rounds: end node = first maximum of dist (lowest node id on exact ties; networkx takes the first
//     in topological order -- identical on every case tested), trace back, zero the path's edges, score / length / IoU
//     tests of vta.py:344-360 in double precision.
#include <float.h>
#include <limits.h>
#include <math.h>

#include <algorithm>
#include <string>

#include "host_util.h"
#include "kernels.h"

using namespace vscb200;

namespace {

constexpr int kTnWarps = 4;
constexpr int kTnMaxTop = 8;
constexpr int kTnMaxStep = 16;
constexpr int kTnMaxBoxes = 32;

struct TnArgs {
  const float* topv; const int32_t* topi; int k;
  int64_t n_pairs; const int32_t* q_len; const int32_t* r_len; const int64_t* row_off;
  int max_step, max_path;
  float min_sim_f; double min_sim, min_length, max_iou;
  uint8_t* state; float* dist; int32_t* pred; uint8_t* sink; int32_t* path;     // per-warp scratch, strided
  size_t state_stride, node_stride;
  int32_t* boxes; int32_t* n_boxes; int box_cap;
};

__global__ void __launch_bounds__(kTnWarps * 32)
tn_align_kernel(TnArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t gw = static_cast<int64_t>(blockIdx.x) * kTnWarps + (threadIdx.x >> 5);
  const int64_t nw = static_cast<int64_t>(gridDim.x) * kTnWarps;
  uint8_t* const state = a.state + gw * a.state_stride;
  float* const dist = a.dist + gw * a.node_stride;
  int32_t* const pred = a.pred + gw * a.node_stride;
  uint8_t* const sink = a.sink + gw * a.node_stride;
  int32_t* const path = a.path + gw * a.node_stride;
  const int D = a.max_step - 1;                       // query-frame span of an edge: 1 .. D

  for (int64_t pair = gw; pair < a.n_pairs; pair += nw) {
    const int Q = a.q_len[pair];
    const int top = min(a.k, a.r_len[pair]);
    int32_t* const out = a.boxes + pair * a.box_cap * 4;
    if (Q <= 0 || top <= 0) {
      if (lane == 0) a.n_boxes[pair] = 0;
      continue;
    }
    const float* const tv = a.topv + a.row_off[pair] * a.k;
    const int32_t* const ti = a.topi + a.row_off[pair] * a.k;
    const int N = 1 + Q * top;
    const int last = N - 1;
    const int tt = top * top;
    const int P = D * top;                             // construction predecessors of a node, in insertion order

    // ---- edges (vta.py:280-311): lane = q_i
    for (int qi = lane; qi < Q; qi += 32) {
      int inter[kTnMaxStep * kTnMaxTop + kTnMaxTop * kTnMaxTop];
      int n_inter = 0;
      int ri[kTnMaxTop];
      for (int c = 0; c < top; ++c) ri[c] = ti[qi * a.k + c];
      for (int dj = 1; dj <= D; ++dj) {
        uint8_t* st = state + (static_cast<size_t>(qi) * D + (dj - 1)) * tt;
        const int qj = qi + dj;
        if (qj >= Q) {
          for (int e = 0; e < tt; ++e) st[e] = 0;
          continue;
        }
        const int n_before = n_inter;
        for (int r = 0; r < top; ++r) {
          const int rj = ti[qj * a.k + r];
          const bool c4 = tv[qj * a.k + r] >= a.min_sim_f;
          for (int c = 0; c < top; ++c) {
            const int d = rj - ri[c];
            bool ok = c4 && d > 0 && d < a.max_step;
            for (int t = 0; ok && t < n_before; ++t) ok = !(ri[c] < inter[t] && inter[t] < rj);
            st[r * top + c] = ok ? 1 : 0;
            if (ok) inter[n_inter++] = rj;               // visible to the NEXT q_j only (n_before)
          }
        }
        // keep the set small: one entry per distinct value is enough
        int m = n_before;
        for (int t = n_before; t < n_inter; ++t) {
          bool dup = false;
          for (int s2 = 0; s2 < m; ++s2) dup = dup || inter[s2] == inter[t];
          if (!dup) inter[m++] = inter[t];
        }
        n_inter = m;
      }
    }
    __syncwarp();
    // ---- sink links (vta.py:317-322): every node close enough to the LAST node gets a zero-weight edge to it
    {
      const int qL = Q - 1, rL = ti[qL * a.k + (top - 1)];
      for (int i = lane; i < last; i += 32) {
        int qi = -1, ri = -1, c = 0;
        if (i > 0) { qi = (i - 1) / top; c = (i - 1) % top; ri = ti[qi * a.k + c]; }
        const bool link = qL > qi && rL > ri && qL - qi <= a.max_step && rL - ri <= a.max_step;
        uint8_t flag = 0;
        if (link) {
          const int dj = qL - qi;
          uint8_t* st = (i > 0 && dj <= D) ? state + (static_cast<size_t>(qi) * D + (dj - 1)) * tt + (top - 1) * top + c : nullptr;
          if (st && *st) *st = 2;                       // existing edge: weight overwritten, position kept
          else flag = 1;                                // new edge: appended after the construction edges
        }
        sink[i] = flag;
      }
    }
    __syncwarp();

    int nb = 0;
    for (int round = 0; round <= a.max_path; ++round) {
      // ---- dist / pred by DP over q (networkx dag_longest_path)
      if (lane == 0) { dist[0] = 0.f; pred[0] = 0; }
      __syncwarp();
      for (int q = 0; q < Q; ++q) {
        for (int k2 = 0; k2 < top; ++k2) {
          const int v = 1 + q * top + k2;
          const float wv = tv[q * a.k + k2];
          float best = -INFINITY;
          int best_key = 0x7fffffff, best_u = v;
          for (int p = lane; p < P; p += 32) {
            const int qi = q - D + p / top, c = p % top;
            if (qi < 0) continue;
            const uint8_t s = state[(static_cast<size_t>(qi) * D + (q - qi - 1)) * tt + k2 * top + c];
            if (!s) continue;
            const int u = 1 + qi * top + c;
            const float cand = dist[u] + (s == 1 ? wv : 0.f);
            if (cand > best) { best = cand; best_key = p; best_u = u; }      // p ascending per lane: first max kept
          }
          if (v == last) {
            // appended sink predecessors, in node order: the source, then nodes within max_step query frames
            const int lo = max(1, 1 + (Q - 1 - a.max_step) * top);
            const int cnt = last - lo + 1;               // candidates lo .. last-1, plus node 0 at position 0
            for (int j = lane; j < cnt; j += 32) {
              const int u = j == 0 ? 0 : lo + j - 1;
              if (!sink[u]) continue;
              const float cand = dist[u];
              const int key = P + j;
              if (cand > best) { best = cand; best_key = key; best_u = u; }
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int ok = __shfl_xor_sync(0xffffffffu, best_key, o);
            const int ou = __shfl_xor_sync(0xffffffffu, best_u, o);
            if (ob > best || (ob == best && ok < best_key)) { best = ob; best_key = ok; best_u = ou; }
          }
          if (lane == 0) {
            const bool has = best_key != 0x7fffffff && best >= 0.f;
            dist[v] = has ? best : 0.f;
            pred[v] = has ? best_u : v;
          }
        }
        __syncwarp();
      }
      // ---- end node: first maximum of dist
      float bd = -INFINITY;
      int bv = 0x7fffffff;
      for (int v = lane; v < N; v += 32) {
        const float dv = dist[v];
        if (dv > bd) { bd = dv; bv = v; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, bd, o);
        const int ov = __shfl_xor_sync(0xffffffffu, bv, o);
        if (ob > bd || (ob == bd && ov < bv)) { bd = ob; bv = ov; }
      }
      // ---- trace back, zero the path's edges, box tests (lane 0)
      int stop = 0;
      if (lane == 0) {
        int len = 0, u = -1, v = bv;
        while (u != v) { path[len++] = v; u = v; v = pred[v]; }
        double score = 0.0;
        int qmin = INT_MAX, qmax = INT_MIN, rmin = INT_MAX, rmax = INT_MIN, kept = 0;
        for (int j = len - 1; j >= 0; --j) {
          const int node = path[j];
          if (j > 0) {
            const int nxt = path[j - 1];                  // edge node -> nxt
            if (node > 0) {
              const int qa = (node - 1) / top, ca = (node - 1) % top, qb = (nxt - 1) / top, kb = (nxt - 1) % top;
              if (qb - qa <= D && !(nxt == last && sink[node]))
                state[(static_cast<size_t>(qa) * D + (qb - qa - 1)) * tt + kb * top + ca] = 2;
            }
          }
          if (node == 0 || node == last) continue;
          const int qn = (node - 1) / top, kn = (node - 1) % top, rn = ti[qn * a.k + kn];
          score += static_cast<double>(tv[qn * a.k + kn]);
          qmin = min(qmin, qn); qmax = max(qmax, qn); rmin = min(rmin, rn); rmax = max(rmax, rn);
          ++kept;
        }
        if (kept == 0) {
          stop = 1;
        } else {
          if (!(score > 0.0)) qmin = qmax = rmin = rmax = 0;
          const double ave = (static_cast<double>(rmax - rmin) + static_cast<double>(qmax - qmin)) / 2.0;
          bool ok = ave > 0.0 && score / ave > a.min_sim && static_cast<double>(min(rmax - rmin, qmax - qmin)) > a.min_length;
          if (ok) {
            double worst = 0.0;
            const double ba = static_cast<double>(qmax - qmin + 1) * static_cast<double>(rmax - rmin + 1);
            for (int b = 0; b < nb; ++b) {
              const int* g = out + b * 4;
              const long long w = max(min(qmax, g[2]) - max(qmin, g[0]) + 1, 0);
              const long long h = max(min(rmax, g[3]) - max(rmin, g[1]) + 1, 0);
              const double inter_a = static_cast<double>(w * h);
              const double ga = static_cast<double>(g[2] - g[0] + 1) * static_cast<double>(g[3] - g[1] + 1);
              worst = fmax(worst, inter_a / (ba + ga - inter_a));
            }
            ok = worst < a.max_iou;
          }
          if (ok && nb < a.box_cap) {
            out[nb * 4 + 0] = qmin; out[nb * 4 + 1] = rmin; out[nb * 4 + 2] = qmax; out[nb * 4 + 3] = rmax;
            ++nb;
          }
        }
      }
      stop = __shfl_sync(0xffffffffu, stop, 0);
      nb = __shfl_sync(0xffffffffu, nb, 0);
      __syncwarp();
      if (stop) break;
    }
    if (lane == 0) a.n_boxes[pair] = nb;
    __syncwarp();
  }
}

// score of VCSLLocalizationMaxSim (localization.py:87-90): similarity[x1:x2, y1:y2].max() - similarity_bias; one warp per pair
__global__ void __launch_bounds__(kTnWarps * 32)
tn_box_max_kernel(const float* __restrict__ sims, const int64_t* __restrict__ s_off, const int32_t* __restrict__ r_len,
                  int64_t n_pairs, const int32_t* __restrict__ boxes, const int32_t* __restrict__ n_boxes, int box_cap,
                  float bias, float* __restrict__ box_score) {
  const int lane = threadIdx.x & 31;
  const int64_t pair = static_cast<int64_t>(blockIdx.x) * kTnWarps + (threadIdx.x >> 5);
  if (pair >= n_pairs) return;
  const float* s = sims + s_off[pair];
  const int R = r_len[pair];
  for (int b = 0; b < n_boxes[pair]; ++b) {
    const int32_t* g = boxes + (pair * box_cap + b) * 4;
    const int x1 = g[0], y1 = g[1], x2 = g[2], y2 = g[3];
    const int w = y2 - y1, cnt = (x2 - x1) * w;
    float m = -INFINITY;
    for (int e = lane; e < cnt; e += 32) m = fmaxf(m, s[static_cast<int64_t>(x1 + e / w) * R + y1 + e % w]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) box_score[pair * box_cap + b] = m - bias;
  }
}

}  // namespace

extern "C" {

int vscb200_tn_align(const float* topv_dev, const int32_t* topi_dev, int k, int64_t n_pairs, const int32_t* q_len_dev,
                     const int32_t* r_len_dev, const int64_t* row_off_dev, int max_q_len, int tn_max_step, int max_path,
                     double min_sim, double min_length, double max_iou, int32_t* boxes_dev, int32_t* n_boxes_dev,
                     void* stream_v) {
  VSCB_REQUIRE(n_pairs >= 0 && k >= 1 && k <= kTnMaxTop, "tn_align: top-k must be in [1, 8]");
  VSCB_REQUIRE(tn_max_step >= 2 && tn_max_step <= kTnMaxStep, "tn_align: tn_max_step must be in [2, 16]");
  VSCB_REQUIRE(max_path >= 0 && max_path + 1 <= kTnMaxBoxes, "tn_align: max_path must be in [0, 31]");
  if (n_pairs == 0) return VSCB200_OK;
  VSCB_REQUIRE(topv_dev && topi_dev && q_len_dev && r_len_dev && row_off_dev && boxes_dev && n_boxes_dev && max_q_len >= 1,
               "tn_align: null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream_v);
  const int64_t want_ctas = (n_pairs + kTnWarps - 1) / kTnWarps;
  const int grid = static_cast<int>(std::min<int64_t>(want_ctas, static_cast<int64_t>(device_sm_count()) * 8));
  const size_t warps = static_cast<size_t>(grid) * kTnWarps;
  const size_t state_stride = (static_cast<size_t>(max_q_len) * (tn_max_step - 1) * k * k + 15) & ~size_t(15);
  const size_t node_stride = (static_cast<size_t>(max_q_len) * k + 2 + 3) & ~size_t(3);
  uint8_t *state = nullptr, *sink = nullptr;
  float* dist = nullptr;
  int32_t *pred = nullptr, *path = nullptr;
  int rc;
  if ((rc = pool_alloc(reinterpret_cast<void**>(&state), warps * state_stride, s))) return rc;
  if ((rc = pool_alloc(reinterpret_cast<void**>(&sink), warps * node_stride, s)) ||
      (rc = pool_alloc(reinterpret_cast<void**>(&dist), warps * node_stride * sizeof(float), s)) ||
      (rc = pool_alloc(reinterpret_cast<void**>(&pred), warps * node_stride * sizeof(int32_t), s)) ||
      (rc = pool_alloc(reinterpret_cast<void**>(&path), warps * node_stride * sizeof(int32_t), s))) {
    pool_free(state, s); pool_free(sink, s); pool_free(dist, s); pool_free(pred, s); pool_free(path, s);
    return rc;
  }
  TnArgs a{topv_dev, topi_dev, k, n_pairs, q_len_dev, r_len_dev, row_off_dev, tn_max_step, max_path,
           static_cast<float>(min_sim), min_sim, min_length, max_iou, state, dist, pred, sink, path, state_stride,
           node_stride, boxes_dev, n_boxes_dev, max_path + 1};
  tn_align_kernel<<<grid, kTnWarps * 32, 0, s>>>(a);
  count_launch();
  cudaError_t e = cudaGetLastError();
  pool_free(state, s); pool_free(sink, s); pool_free(dist, s); pool_free(pred, s); pool_free(path, s);
  if (e != cudaSuccess) {
    set_last_error(std::string("tn_align: launch failed: ") + cudaGetErrorString(e));
    return VSCB200_ERR_CUDA;
  }
  return VSCB200_OK;
}

int vscb200_tn_box_scores(const float* sims_dev, const int64_t* s_off_dev, const int32_t* r_len_dev, int64_t n_pairs,
                          const int32_t* boxes_dev, const int32_t* n_boxes_dev, int box_cap, float bias,
                          float* box_score_dev, void* stream_v) {
  VSCB_REQUIRE(n_pairs >= 0 && box_cap >= 1, "tn_box_scores: bad shape");
  if (n_pairs == 0) return VSCB200_OK;
  VSCB_REQUIRE(sims_dev && s_off_dev && r_len_dev && boxes_dev && n_boxes_dev && box_score_dev, "tn_box_scores: null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream_v);
  tn_box_max_kernel<<<static_cast<unsigned>((n_pairs + kTnWarps - 1) / kTnWarps), kTnWarps * 32, 0, s>>>(
      sims_dev, s_off_dev, r_len_dev, n_pairs, boxes_dev, n_boxes_dev, box_cap, bias, box_score_dev);
  count_launch();
  VSCB_CUDA_OK(cudaGetLastError());
  return VSCB200_OK;
}

}  // extern "C"
