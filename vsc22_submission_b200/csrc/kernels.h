// Internal (C++) launch entry points shared between translation units.
//
// `reverse`: walk the rows (LayerNorm), M tiles (GEMM) or frames (attention) from the last to the first.
// The encoder's activations (x 155 MB, qkv 232 MB, u 310 MB per 256 frames) exceed the 126 MB L2, so a
// consumer that walks in the producer's order misses on everything (LRU streaming); walking in the
// opposite order consumes the most recently written -- still L2 resident -- rows first.  vit.cu alternates
// the direction kernel by kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vscb200 {

constexpr int VSCB_EPI_PATCH_F32_ID = 3;

int gemm_bf16(const void* A, const void* W, const float* bias, void* C, int64_t M, int N, int K, int64_t lda,
              int64_t ldw, int64_t ldc, int epilogue, int act, cudaStream_t stream, const float* pos, int patch_P,
              bool reverse = false, int qk_norm_cols = 0, const float* qscale = nullptr, const void* A_lo = nullptr,
              const void* W_lo = nullptr, void* C_lo = nullptr);
int attention(const void* qkv, void* out, int n_frames, int T, int heads, int head_dim, cudaStream_t stream,
              bool reverse = false);
// lo_off / y_lo: fp32-equivalent mode -- the bf16 output is written as two planes (hi at y, the rounding residual
// bf16(v - hi) at y + lo_off elements); 0 / nullptr: single plane
int layernorm(const float* x, const float* gamma, const float* beta, void* y, int64_t rows, int width, float eps,
              int out_bf16, cudaStream_t stream, bool reverse = false, int64_t lo_off = 0);
int cast_f32_bf16_padded(const float* x, void* y, int64_t rows, int cols, int ld_out, cudaStream_t stream, void* y_lo = nullptr);
int im2row(const float* frames, void* patches, int64_t n, int img, int patch, int Kp, cudaStream_t stream, int64_t lo_off = 0);
// attention_fp32.cu: fp32 attention over segments of N consecutive rows (hi + lo planes when *_lo_off != 0)
int attention_fp32(const void* qkv, int64_t qkv_lo_off, void* out, int64_t out_lo_off, int64_t n_segs, int N, int heads,
                   int head_dim, float scale, const float* tables, int ws, int res, int shift, int nWx, int nW_per_frame,
                   cudaStream_t stream);
// attention_ws.cu: streaming tcgen05 attention of the ViT encoders (head_dim 64, 129..640 tokens) and large Swin-V2 windows
bool attention_ws_supported(int T, int head_dim);
int attention_ws(const void* qkv, void* out, int n_frames, int T, int heads, cudaStream_t stream, bool reverse);
int attention_ws_swin(const void* qkv, void* out, int64_t n_windows_total, int ws, int heads, const float* tables, int shift,
                      int nWx, int nW_per_frame, cudaStream_t stream);
// attention_kb.cu: K-blocked tcgen05 attention, segments of 129..640 tokens (ViT form: head_dim 64; Swin-V2 form: head_dim 32
// with relative-position bias tables and the shifted-window mask)
bool attention_kb_supported(int N, int head_dim);
int attention_kb(const void* qkv, void* out, int64_t n_segs, int N, int heads, int head_dim, float scale, const float* tables,
                 int ws, int shift, int nWx, int nW_per_frame, cudaStream_t stream);
int cls_rows(const float* cls, const float* pos, float* x, int64_t n, int T, int W, cudaStream_t stream);
int gem_head(const float* y, const float* gamma, const float* beta, const float* head_w, const float* head_b,
             float* out, int64_t n, int T, int C, int out_dim, float eps, float p, bool fuse_ln,
             cudaStream_t stream);

}  // namespace vscb200

namespace vscb200 {
int scores_simt(const float* Q, const float* R, float* S, int64_t nq, int64_t nr, int d, int64_t ldS, bool l2,
                const float* qn, const float* rn, cudaStream_t stream);
int row_sqnorm(const float* x, int64_t n, int d, float* out, cudaStream_t stream);
// x [n, d] fp32 -> hi, lo [n, dp] bf16 planes (sim_tc.cu)
int split_planes(const float* x, void* hi, void* lo, int64_t n, int d, int dp, cudaStream_t stream);
// sim_tc1.cu: single-pass (one bf16 MMA per product) top-k selection with an error margin + exact rescoring
int sim1_pairs(int64_t nq, int64_t nr);
int sim1_list_cap();
// sim.cu: score-normalisation transform of raw rows written straight into an index's storage (bank, planes, norms, maxima)
int sn_add_rows(const float* x, int64_t n, int d, int drop_dim, int l2_normalize, float fill, const float* bias,
                const int* drop_dim_dev, float* bank, void* hi, void* lo, float* rnorm, unsigned int* max_bits, int dout, int dp,
                cudaStream_t stream);
int q_hi_norm(const float* x, void* hi, float* sq, float* sq_lo, int64_t n, int d, int dp, cudaStream_t stream, void* lo = nullptr,
              unsigned int* max_bits = nullptr);
int bank_norm_max(const float* x, int64_t n, int d, unsigned int* max_bits, cudaStream_t stream);
int sim1_max_k();
int sim1_scratch_ints(int64_t nq);
size_t sim1_partial_bytes();
int sim1_topk(const void* Qh, const void* Rh, int64_t nq, int64_t nr, int d, int dp, bool l2, int k, const float* qn,
              const float* qn_lo, const float* rn, const unsigned int* bank_max_bits, void* cand, int* scratch,
              cudaStream_t stream);
// thresholds of the query rows from one-pass scores S [nq, ncs] against a column sample of the bank (ncs <= 1024)
int sim1_boot_tau(const float* S, int64_t nq, int ncs, int k, const float* qn, const float* qn_lo,
                  const unsigned int* bank_max_bits, int d, int* scratch, cudaStream_t stream);
int sim1_rescore(const float* Q, const float* bank, int d, bool l2, int64_t nq, int64_t nr, const void* cand, int k,
                 const float* qn, const float* qn_lo, const unsigned int* bank_max_bits, float* D, int64_t* I, int64_t id_offset,
                 int* scratch, void* partial, cudaStream_t stream);
int sn_transform(const float* x, int64_t n, int d, int drop_dim, int l2_normalize, float fill, const float* bias,
                 float* out, cudaStream_t stream, const int* drop_dim_dev = nullptr);
// row r = S + r*ldS, element i of a row at [i*es] (es = 1: dense rows; es > 1: a column of a row-major matrix)
int topk_rows(const float* S, int64_t ldS, int64_t nq, int64_t n, int k, bool keep_max, float* D, int64_t* I,
              int64_t id_offset, cudaStream_t stream, int64_t es = 1);
// sim_tc.cu: scores on pre-split operand planes.  passes = 3: fp32-equivalent (hi.hi + lo.hi + hi.lo), 1: Qh.Rh only;
// r_stride > 1: bank row j of the call is row j * r_stride of the planes (column sample of the score block)
int scores_tc_planes(const void* Qh, const void* Ql, const void* Rh, const void* Rl, float* S, int64_t nq, int64_t nr,
                     int dp, int64_t ldS, bool l2, const float* qn, const float* rn, cudaStream_t stream, int passes = 3,
                     int64_t r_stride = 1);
int scores_tc_emit(const void* Qh, const void* Rh, int64_t nq, int64_t nr, int dp, bool l2, const float* qn, const float* rn,
                   const float* marg, float radius, bool has_radius, int64_t q0, float* bufv, uint64_t* bufp,
                   unsigned long long* counter, unsigned long long cap, cudaStream_t stream);
// sim_stream.cu: streaming search for a few query rows (group maxima + exact rescoring of the best groups)
int sim_stream_groupmax(const void* Qh, const void* Ql, const void* Rh, const void* Rl, int64_t nq, int64_t nr, int dp,
                        bool l2, const float* qn, const float* rn, float* gmax, int Npad, cudaStream_t stream);
int group_topk_chunks(int64_t G);
int group_topk(const float* gmax, int64_t G, int Npad, int nq, int kg, int chunks, float* cand_v, int32_t* cand_g,
               cudaStream_t stream);
int group_rescore(const float* Q, const float* bank, int d, bool l2, int64_t nr, const int64_t* gsel, const float* cand_v,
                  const int32_t* cand_g, int ncg, int kg, int64_t nq, int k, float* D, int64_t* I, int64_t id_offset,
                  cudaStream_t stream, int gs = 32);
// few query rows: one CTA per (query, group); keys / counters: scratch (counters zero before the first call, left zero)
size_t group_rescore_few_key_bytes(int64_t nq, int kg, int gs);
int group_rescore_few(const float* Q, const float* bank, int d, bool l2, int64_t nr, const float* cand_v, const int32_t* cand_g,
                      int ncg, int kg, int64_t nq, int k, float* D, int64_t* I, int64_t id_offset, void* keys, int* counters,
                      cudaStream_t stream, int gs = 32);
int rescore_sort(const float* Q, const float* bank, int d, bool l2, const int64_t* Iin, int kin, int64_t nq, int k,
                 float* D, int64_t* I, int64_t id_offset, cudaStream_t stream);
// Q != nullptr: S holds tensor-core scores; borderline pairs and reported distances are recomputed in fp32
int range_count(const float* S, int64_t ldS, int64_t nq, int64_t n, float thr, bool keep_max,
                unsigned long long* counts, cudaStream_t stream, const float* Q, const float* bank, int d,
                const float* qn, const float* rn);
int range_fill(const float* S, int64_t ldS, int64_t nq, int64_t n, float thr, bool keep_max,
               const unsigned long long* offsets, float* D, int64_t* I, int64_t id_offset, cudaStream_t stream,
               const float* Q, const float* bank, int d, const float* qn, const float* rn);
}  // namespace vscb200
