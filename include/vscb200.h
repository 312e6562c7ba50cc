/* vscb200 -- C ABI of the B200-native VSC22 hot path (libvscb200.so).
 *
 * Plain C: opaque handles, raw pointers, sizes, an explicit CUDA stream (passed as void*, it is a
 * cudaStream_t / CUstream; NULL = legacy default stream).  Every function returns an int status
 * (0 = ok); no exception crosses this boundary; vscb200_last_error() gives the message of the last
 * failure on the calling thread.  No CUDA call is made at library-load time (the reference forks
 * DataLoader workers before CUDA init -- vsc/baseline/inference.py:1-17).
 *
 * Two seams of the reference are served:
 *
 *  (B) the `faiss` calls of vsc/index.py, vsc/exhaustive_search.py,
 *      vsc/baseline/score_normalization.py and M/infer/infer_matching.py   -> vscb200_index_*
 *  (A) the TorchScript encoder call `model(frames)` of D/infer/src/extractor.py:25,
 *      D/infer/extract_query_feats.py:148,160 and M/infer/infer_matching.py:127 -> vscb200_vit_*
 *
 * `_host` variants take HOST buffers and perform the host<->device copies themselves: they are what
 * the numpy-facing faiss-compatible module binds (the reference passes numpy arrays to faiss).
 */
#ifndef VSCB200_H_
#define VSCB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSCB200_OK 0
#define VSCB200_ERR_INVALID 1 /* bad argument / shape / state */
#define VSCB200_ERR_CUDA 2    /* a CUDA runtime/driver call failed */
#define VSCB200_ERR_NOMEM 3

#define VSCB200_METRIC_INNER_PRODUCT 0 /* faiss.METRIC_INNER_PRODUCT */
#define VSCB200_METRIC_L2 1            /* faiss.METRIC_L2 (squared L2) */

const char* vscb200_last_error(void);
int vscb200_version(void);
/* number of kernel launches issued by this library since load (bench.py's gpu_launches) */
int64_t vscb200_launch_count(void);
/* faiss.get_num_gpus() -- vsc/index.py:169, exhaustive_search.py:28,229; 0 when no CUDA device */
int vscb200_device_count(void);
int vscb200_set_device(int device);
/* return the library's cached device blocks (index workspaces / staging) to the driver */
int vscb200_trim(void);
/* Per-kernel device timing with CUDA events on the launching stream (bench.py's live roofline).
 * kinds: 0 gemm, 1 attention, 2 layernorm, 3 other encoder kernels, 4 similarity scores, 5 select.
 * work = algorithmic FLOPs (gemm/attention/scores) or bytes (others). */
int vscb200_prof_enable(int on);
int vscb200_prof_collect(double* ms_by_kind, int64_t* launches_by_kind, double* work_by_kind, int nkinds);

/* ------------------------------------------------------------------------------------------------
 * (B) flat similarity index.  Replaces faiss.index_factory(d,"Flat",metric) / IndexFlat
 *     (vsc/index.py:81; exhaustive_search.py:26,70,102) and its methods.
 * ---------------------------------------------------------------------------------------------- */
typedef struct vscb200_index vscb200_index;

/* faiss.index_factory(d, "Flat", metric) -- vsc/index.py:81 */
int vscb200_index_create(int d, int metric, vscb200_index** out);
void vscb200_index_destroy(vscb200_index* ix);
/* index.add(x) -- vsc/index.py:94.  x: [n, d] float32, C-contiguous. Rows get ids ntotal..ntotal+n-1. */
int vscb200_index_add(vscb200_index* ix, const float* x_dev, int64_t n, void* stream);
/* index.add(rows transformed by score normalisation) in one pass over the raw rows -- replaces the pair
 * `sn_features = normalize(np.delete(x, low_var_dim, 1)); np.concatenate([.., fill])` (score_normalization.py:73-83,
 * 96-101) + index.add (vsc/index.py:94) without materialising the transformed array.  See vscb200_sn_transform for the
 * arguments; the index dimension is d_in when a column is dropped, else d_in + 1. */
int vscb200_index_add_sn(vscb200_index* ix, const float* x_dev, int64_t n, int d_in, int drop_dim, const int* drop_dim_dev,
                         int l2_normalize, float fill, const float* bias_dev, void* stream);
int vscb200_index_add_host(vscb200_index* ix, const float* x_host, int64_t n);
/* index.reset() -- exhaustive_search.py:39 */
int vscb200_index_reset(vscb200_index* ix);
/* index.ntotal / index.d / index.metric_type */
int64_t vscb200_index_ntotal(const vscb200_index* ix);
int vscb200_index_dim(const vscb200_index* ix);
int vscb200_index_metric(const vscb200_index* ix);
/* Restrict searches to bank rows [row0, row0+rows) and report ids offset by id_offset: bank
 * sharding over ranks.  Default: whole bank, offset 0. */
int vscb200_index_set_id_offset(vscb200_index* ix, int64_t id_offset);

/* index.search(x, k) -> (D, I) -- vsc/index.py:174, score_normalization.py:95,141,
 * exhaustive_search.py:62, infer_matching.py:232.  D: [nq,k] float32 best-first, I: [nq,k] int64;
 * ties -> lower id; k > ntotal pads with (-FLT_MAX | +FLT_MAX, -1).  1 <= k <= 2048. */
int vscb200_index_search(vscb200_index* ix, const float* q_dev, int64_t nq, int k, float* D_dev, int64_t* I_dev,
                         void* stream);
int vscb200_index_search_host(vscb200_index* ix, const float* q_host, int64_t nq, int k, float* D_host,
                              int64_t* I_host);
/* index.reconstruct_n(i0, n) -> rows [i0, i0+n) as float32 into host memory or the memory of any device; complete on return */
int vscb200_index_reconstruct_n(vscb200_index* ix, int64_t i0, int64_t n, float* out);
/* diagnostic: how many queries of the last small-k batch search needed the exhaustive fp32 fallback (-1: none ran) */
int64_t vscb200_index_last_fallbacks(vscb200_index* ix);

/* index.range_search(x, thresh) -> (lims, D, I) -- exhaustive_search.py:74,126,246,
 * infer_matching.py:235.  Strict '>' (IP) / '<' (L2).  lims: [nq+1] uint64 (host); D/I are allocated
 * by the library in HOST memory (ascending id inside each row) and released with vscb200_free. */
int vscb200_index_range_search_host(vscb200_index* ix, const float* q_host, int64_t nq, float thresh,
                                    uint64_t* lims_host, float** D_host, int64_t** I_host);
void vscb200_free(void* p);

/* Dense score block S[nq, ntotal] (float32, device) -- the per-pair similarity matrices of
 * vsc/baseline/localization.py:32-35 and BASELINE config 5. */
int vscb200_index_scores(vscb200_index* ix, const float* q_dev, int64_t nq, float* S_dev, int64_t ldS,
                         void* stream);

/* Global (cross-query) candidate search -- vsc/index.py:142-165 `_global_threshold_knn_search` with
 * vsc/exhaustive_search.py:206-292 `range_search_max_results` (adaptive radius, <= 2K survivors), and, with a
 * threshold, the search loop of M/infer/infer_matching.py:229-247.  Finds the `global_k` best (query row, bank row)
 * pairs over ALL nq query rows (global_k <= 0: no limit), optionally restricted to pairs strictly better than
 * `thresh` (use_thresh != 0; '>' IP / '<' L2).  Scores are exact fp32; order: best first, ties by (query row,
 * bank row).  The result stays inside the index until the next global search; *n_found_host receives its length. */
int vscb200_index_global_search(vscb200_index* ix, const float* q_dev, int64_t nq, int64_t global_k, int use_thresh,
                                float thresh, int64_t* n_found_host, void* stream);
/* copy the retained result: scores [n] float32, query rows [n] int64, bank rows [n] int64 (device or host memory) */
int vscb200_index_global_results(vscb200_index* ix, float* scores, int64_t* qrows, int64_t* brows, void* stream);
/* Reduce the retained frame pairs to (query video, ref video) candidates scored by their best frame pair, sorted
 * best first -- vsc/index.py:119-140 + vsc/candidates.py:24-40 (MaxScoreAggregation + sort), infer_matching.py:248-256.
 * q_offsets_dev [nqv+1] / r_offsets_dev [nrv+1] (int64, device): first row of each video in the query matrix / bank. */
int vscb200_index_global_video_pairs(vscb200_index* ix, const int64_t* q_offsets_dev, int64_t nqv,
                                     const int64_t* r_offsets_dev, int64_t nrv, int64_t* n_pairs_host, void* stream);
int vscb200_index_video_pair_results(vscb200_index* ix, float* scores, int64_t* qvideo, int64_t* rvideo, void* stream);

/* Fused score normalisation prologue (vsc/baseline/score_normalization.py:71-103):
 * out[n, d] = [ l2norm(drop(x, drop_dim)) , last ] where last is `fill` (1.0 for references, :99-101)
 * or, when bias_dev != NULL, bias_dev[row] (queries, :96-97).  x: [n, d] -> out: [n, d]. */
int vscb200_sn_transform(const float* x_dev, int64_t n, int d, int drop_dim, int l2_normalize, float fill,
                         const float* bias_dev, float* out_dev, void* stream);
/* the same with the dropped column read from device memory (*drop_dim_dev written by vscb200_low_var_dim_dev earlier in
 * the stream): the whole score-normalisation chain is enqueued without a host round trip.  Output has d columns. */
int vscb200_sn_transform_dev(const float* x_dev, int64_t n, int d, const int* drop_dim_dev, int l2_normalize, float fill,
                             const float* bias_dev, float* out_dev, void* stream);
/* column variance argmin left in device memory (*dim_dev, int32), no synchronisation */
int vscb200_low_var_dim_dev(const float* x_dev, int64_t n, int d, int* dim_dev, void* stream);
/* column variance argmin of a [n, d] matrix (score_normalization.py:72), result to *dim_host */
int vscb200_low_var_dim(const float* x_dev, int64_t n, int d, int* dim_host, void* stream);
/* Row-sharded banks: the two passes of the column variance separately, so that the [d] vectors can be
 * all-reduced between them.  out[c] = sum_r x[r,c] (sum_in_dev == NULL) or sum_r (x[r,c] - sum_in_dev[c]*inv_n)^2.
 * Fixed reduction order (deterministic).  vscb200_var_argmin_dev: first minimum of ss_dev[0..d) -> *dim_dev. */
int vscb200_col_sums(const float* x_dev, int64_t n, int d, const double* sum_in_dev, double inv_n, double* out_dev,
                     void* stream);
int vscb200_var_argmin_dev(const double* ss_dev, int d, int* dim_dev, void* stream);
/* The same with ONE collective for a row-sharded bank: vscb200_col_moments_local writes out3_dev[0..3d) = per column
 * (sum x | sum (x - local mean)^2 | (sum x)^2 / n) of the local shard; the caller all-reduces (sums) the 3d doubles over
 * the shards; vscb200_var_argmin_moments combines them exactly (M2 = sum M2_r + sum S_r^2 / n_r - (sum S_r)^2 / N) and
 * writes the first column of minimum variance. */
int vscb200_col_moments_local(const float* x_dev, int64_t n, int d, double* out3_dev, void* stream);
int vscb200_var_argmin_moments(const double* m3_dev, double n_total, int d, int* dim_dev, void* stream);
/* Exchange format of partial top-k results between bank shards: one 64-bit key per entry,
 * (order-preserving score bits << 32) | ~id (0 = padding; ids < 2^32), so ONE all-gather moves scores and ids.
 * vscb200_topk_merge: keys_dev [parts][nq][kin] (the all-gather layout) -> the kout best per query, best first, ties to
 * the lower id (faiss order). */
int vscb200_topk_pack(const float* D_dev, const int64_t* I_dev, int64_t nq, int k, int keep_max, uint64_t* keys_dev,
                      void* stream);
int vscb200_topk_merge(const uint64_t* keys_dev, int parts, int64_t nq, int kin, int kout, int keep_max, float* D_dev,
                       int64_t* I_dev, void* stream);
/* vscb200_topk_pack into columns [col0, col0 + k) of rows of ld keys: the layout vscb200_topk_merge_cols reads. */
int vscb200_topk_pack_cols(const float* D_dev, const int64_t* I_dev, int64_t nq, int k, int keep_max, uint64_t* keys_dev, int ld,
                           int col0, void* stream);
/* score_normalizev2 of the matching track (VSC22-Matching-Track-1st/vsc/baseline/score_normalization.py:141-153):
 * out[row] = l2_normalize(x[row] - beta * mean_k z[ids[row, k]]), ids = the nk nearest noise rows of normalize(x[row])
 * (from vscb200_index_search over the normalised noise bank); z is the UN-normalised noise bank. */
int vscb200_sn2_adapt(const float* x_dev, const float* z_dev, const int64_t* ids_dev, int64_t n, int d, int nk, float beta,
                      int l2_normalize, float* out_dev, void* stream);
/* The same over columns [col0, col0 + kin) of rows of ld keys (several partial results travelling in one all-gather), with
 * an optional per-row bias added to the merged scores (the score-normalisation bias of the query, which does not change
 * the ranking of a query's references and can therefore be applied after the exchange). */
int vscb200_topk_merge_cols(const uint64_t* keys_dev, int parts, int64_t nq, int ld, int col0, int kin, int kout, int keep_max,
                            const float* bias_dev, float* D_dev, int64_t* I_dev, void* stream);
/* bias[row] = -beta * mean(D[row, :nk])  (score_normalization.py:96) */
int vscb200_sn_bias(const float* D_dev, int64_t nq, int k, int nk, float beta, float* bias_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (A) ViT frame encoder.  Replaces the TorchScript module called at D/infer/src/extractor.py:25
 *     (architecture: D/train/train_vid_score/video/clip.py:85-163 + backbones/vit.py:42-58 tail,
 *     or timm ViT + sscd.py:30-40,86 head).
 * ---------------------------------------------------------------------------------------------- */
typedef struct vscb200_vit vscb200_vit;

/* Arithmetic of the encoders' matrix products.  BF16: operands rounded to bf16 (fp32 accumulation, residual stream,
 * LayerNorm, softmax) -- the throughput mode, equal to the matched-precision oracle within 1e-3.  FP32: every operand
 * carried as two bf16 planes (hi + lo, 16 mantissa bits), three tcgen05 MMAs per product, attention in fp32 -- equal to
 * the reference's fp32 modules within 1e-3 (the reference runs them without autocast: D/infer/src/extractor.py:25). */
#define VSCB200_PRECISION_BF16 0
#define VSCB200_PRECISION_FP32 1

#define VSCB200_ACT_QUICK_GELU 0 /* clip.py:22-25 */
#define VSCB200_ACT_GELU 1       /* timm Mlp (erf) */
#define VSCB200_TAIL_TOKENS 0          /* [n, T, W]  (clip.py:158; caller slices [:,0]) */
#define VSCB200_TAIL_GEM_LINEAR 1      /* backbones/vit.py:42-58 */
#define VSCB200_TAIL_GEM_CONV_LINEAR 2 /* sscd.py:30-40 + :86 */

typedef struct vscb200_vit_spec {
  int img, patch, width, layers, heads;
  int patch_bias; /* timm: 1, CLIP: 0 (clip.py:105) */
  int pre_norm;   /* CLIP ln_pre (clip.py:152) */
  int act;
  int tail;
  int out_dim;
  int gem_hidden;
  float ln_eps;
  float gem_p;
  int precision; /* VSCB200_PRECISION_* */
} vscb200_vit_spec;

int vscb200_vit_create(const vscb200_vit_spec* spec, int max_frames, vscb200_vit** out);
void vscb200_vit_destroy(vscb200_vit* m);
/* Upload one named fp32 parameter (device pointer); names as in oracle/vit_ref.py / encoder.py:
 * patch_w patch_b cls pos ln_pre_w ln_pre_b l<i>.{ln1_w,ln1_b,qkv_w,qkv_b,proj_w,proj_b,ln2_w,ln2_b,
 * fc1_w,fc1_b,fc2_w,fc2_b} ln_post_w ln_post_b gem_conv_w gem_conv_b head_w head_b */
int vscb200_vit_set_param(vscb200_vit* m, const char* name, const float* w_dev, int64_t count, void* stream);
/* model(frames): frames [n,3,img,img] float32 NCHW contiguous (device) -> out (device) [n,out_dim] or
 * [n,T,W]. Any n >= 1 (internally chunked by max_frames). */
int vscb200_vit_forward(vscb200_vit* m, const float* frames_dev, int64_t n, float* out_dev, void* stream);
/* same with pinned/pageable HOST buffers (copies inside) */
int vscb200_vit_forward_host(vscb200_vit* m, const float* frames_host, int64_t n, float* out_host);
int64_t vscb200_vit_out_elems_per_frame(const vscb200_vit* m);

/* ------------------------------------------------------------------------------------------------
 * (A) Swin-V2 frame encoder.  Replaces the TorchScript module `swinv2_v1xx` called at
 *     D/infer/src/extractor.py:25 / D/infer/extract_query_feats.py:148 (architecture:
 *     D/train/train_v106/vsc/baseline/model_factory/backbones/swinv2.py:502-633, config_v106.py:8-24).
 *     head_dim is 32 in every stage (heads[i] * 32 == embed << i); window sides 4 / 8 / 16 run on the tcgen05 window
 *     attention, any other side dividing the token map (24: SwinV2-L@384) on the fp32 kernel.
 * ---------------------------------------------------------------------------------------------- */
typedef struct vscb200_swin vscb200_swin;

typedef struct vscb200_swin_spec {
  int img, patch, embed, n_stages;
  int depths[4];
  int heads[4];
  int window;
  int pretrained_windows[4]; /* swinv2.py:107-112: normalisation of the log-spaced relative coordinates */
  int out_dim;
  float ln_eps;
  float gem_p;
  int precision; /* VSCB200_PRECISION_* */
} vscb200_swin_spec;

int vscb200_swin_create(const vscb200_swin_spec* spec, int max_frames, vscb200_swin** out);
void vscb200_swin_destroy(vscb200_swin* m);
/* Upload one named fp32 parameter (device pointer); names are the reference's state-dict names
 * (patch_embed.proj.weight ... layers.<i>.blocks.<j>.attn.qkv.weight ... output_proj.bias). */
int vscb200_swin_set_param(vscb200_swin* m, const char* name, const float* w_dev, int64_t count, void* stream);
/* model(frames): frames [n,3,img,img] float32 NCHW contiguous (device) -> out (device) [n,out_dim] */
int vscb200_swin_forward(vscb200_swin* m, const float* frames_dev, int64_t n, float* out_dev, void* stream);
int vscb200_swin_forward_host(vscb200_swin* m, const float* frames_host, int64_t n, float* out_host);
int vscb200_swin_out_dim(const vscb200_swin* m);

/* ------------------------------------------------------------------------------------------------
 * Ensemble tail: per-model row L2 normalisation -> concatenation -> PCA.transform
 * ((X - mean_) @ components_.T), D/infer/concat_pca_sn.py:56-64, D/infer/extract_query_feats.py:169-204,
 * M/infer/infer_matching.py:140-145.  parts_dev: n_parts (<= 8) device pointers [n, dims[i]] fp32 (HOST array of
 * pointers); mean_dev [sum dims]; components_dev [out_dim, sum dims] (sklearn PCA.components_); out_dev [n, out_dim].
 * ---------------------------------------------------------------------------------------------- */
int vscb200_ensemble_pca(const float* const* parts_dev, const int* dims, int n_parts, int64_t n, const float* mean_dev,
                         const float* components_dev, int out_dim, float* out_dev, void* stream);
/* Near-duplicate frame filter of ONE query video (extract_query_feats.py:190-200): keep_dev[i] = 0 for every frame
 * removed by the greedy sweep over frames in descending mean-similarity order (threshold FRAME_THRESHOLD = 0.975 in the
 * reference).  feat_dev: [n, d] float32 (rows are normalised inside), n <= 4096. */
int vscb200_near_dup_keep(const float* feat_dev, int64_t n, int d, double threshold, uint8_t* keep_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Candidate-pair frame-similarity matrices: `np.matmul(a, b.T)` (+ similarity_bias) of
 * vsc/baseline/localization.py:32-35,52-57 for ALL candidate pairs in one launch, and the per-row top-k that
 * opens the temporal-network alignment (vcsl/vta.py:262-265).  q_dev [all query frames, d], r_dev [all ref frames, d];
 * pair p = rows [q_off[p], q_off[p]+q_len[p]) x [r_off[p], r_off[p]+r_len[p]); its row-major [q_len, r_len] block starts at
 * sims_dev + s_off[p].  Top-k rows of pair p start at row_off[p] in topv/topi [sum q_len, k] (-1 / -FLT_MAX padding).
 * All descriptor arrays are device pointers.
 * ---------------------------------------------------------------------------------------------- */
int vscb200_pair_sims(const float* q_dev, const float* r_dev, int d, int64_t n_pairs, const int64_t* q_off_dev,
                      const int32_t* q_len_dev, const int64_t* r_off_dev, const int32_t* r_len_dev, const int64_t* s_off_dev,
                      float bias, float* sims_dev, void* stream);
int vscb200_pair_topk(const float* sims_dev, int64_t n_pairs, const int32_t* q_len_dev, const int32_t* r_len_dev,
                      const int64_t* s_off_dev, const int64_t* row_off_dev, int k, float* topv_dev, int32_t* topi_dev,
                      void* stream);

/* Frame preprocessing (after JPEG decode): Pillow's antialiased bicubic `Image.resize` + torchvision
 * ToTensor + Normalize of D/infer/src/transform.py:20-43, bit-exact.  frames_dev: [n, H, W, 3] uint8 RGB (device);
 * mean3 / std3: host float[3]; mid_scratch_dev: n*H*out_w*3 bytes (may be NULL when W == out_w);
 * out_dev: [n, 3, out_h, out_w] float32 -- the tensor the encoder's forward takes. */
int vscb200_resize_normalize(const uint8_t* frames_dev, int64_t n, int H, int W, int out_h, int out_w, const float* mean3,
                             const float* std3, uint8_t* mid_scratch_dev, float* out_dev, void* stream);

/* Frame decoding (the decode half): the baseline JPEG files of one video's frames -> RGB on the device,
 * bit-identical to PIL.Image.open(...).convert("RGB") (libjpeg-turbo defaults: islow IDCT, fancy upsampling), which the
 * reference calls frame by frame at D/infer/src/dataset.py:137-141.  jpeg_ptrs / jpeg_sizes: the n files in host memory
 * (all of one size and chroma subsampling); rgb_dev: [n, H, W, 3] uint8 (device) -- the input of vscb200_resize_normalize.
 * rgb_dev == NULL: only parse and return the frame size through h_out / w_out.  8-bit YCbCr 4:2:0 / 4:2:2 / 4:4:4 or
 * grey, one interleaved scan, restart intervals; progressive / arithmetic / CMYK files return VSCB200_ERR_INVALID. */
int vscb200_jpeg_decode(const uint8_t* const* jpeg_ptrs, const uint64_t* jpeg_sizes, int64_t n, uint8_t* rgb_dev, int* h_out,
                        int* w_out, void* stream);

/* Matching-track candidate features: M/infer/src/utils.py:18-47 / :50-73 + the zero-padded
 * similarity images of M/infer/src/dataset.py:103-144.  sims_dev: the blocks written by vscb200_pair_sims (whole query
 * video x reference).  seg_len[p] = query_video_len_map[qid]: when the query holds several seg_len-frame copies, the one
 * with the best mean of its 10 largest row maxima is kept.  images_dev [n_pairs, 1 or 2, H, W]: the kept block cropped /
 * zero-padded to H x W and (with_transpose) its transpose (ref x query); info_dev [n_pairs, 4] = (segment, rows kept,
 * h, w).  rowmax_scratch_dev: [sum q_len] floats. */
int vscb200_pair_segment_images(const float* sims_dev, int64_t n_pairs, const int32_t* q_len_dev, const int32_t* r_len_dev,
                                const int64_t* s_off_dev, const int64_t* row_off_dev, const int32_t* seg_len_dev,
                                float* rowmax_scratch_dev, int H, int W, int with_transpose, float* images_dev,
                                int32_t* info_dev, void* stream);

/* Temporal-network alignment of every candidate pair (vcsl/vta.py:244-363 `tn`, run by localization.py:38-76 through a
 * 16-process pool): input = the per-row top-k of vscb200_pair_topk (same k, row_off, lengths); output = up to
 * max_path + 1 boxes [q_min, r_min, q_max, r_max] per pair in boxes_dev [n_pairs, max_path + 1, 4] and their number in
 * n_boxes_dev [n_pairs].  max_q_len >= every q_len (sizes the per-warp scratch).  k <= 8, tn_max_step <= 16. */
int vscb200_tn_align(const float* topv_dev, const int32_t* topi_dev, int k, int64_t n_pairs, const int32_t* q_len_dev,
                     const int32_t* r_len_dev, const int64_t* row_off_dev, int max_q_len, int tn_max_step, int max_path,
                     double min_sim, double min_length, double max_iou, int32_t* boxes_dev, int32_t* n_boxes_dev,
                     void* stream);
/* VCSLLocalizationMaxSim.score (localization.py:87-90): sims[x1:x2, y1:y2].max() - bias for every box */
int vscb200_tn_box_scores(const float* sims_dev, const int64_t* s_off_dev, const int32_t* r_len_dev, int64_t n_pairs,
                          const int32_t* boxes_dev, const int32_t* n_boxes_dev, int box_cap, float bias,
                          float* box_score_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Building blocks exported for unit tests and micro-benchmarks.
 * ---------------------------------------------------------------------------------------------- */
#define VSCB200_EPI_BF16 0          /* C_bf16 = act(A*W^T + bias)            */
#define VSCB200_EPI_F32 1           /* C_f32  = act(A*W^T + bias)            */
#define VSCB200_EPI_RESIDUAL_F32 2  /* C_f32 += A*W^T + bias   (in place)    */
/* C[M,N] = A[M,K](bf16) * W[N,K]^T(bf16) (+bias f32[N]) ; tcgen05 + TMA + TMEM persistent kernel.
 * K % 8 == 0, N % 8 == 0; act: -1 none, VSCB200_ACT_*. */
int vscb200_gemm_bf16(const void* A_bf16, const void* W_bf16, const float* bias, void* C, int64_t M, int N, int K,
                      int64_t lda, int64_t ldw, int64_t ldc, int epilogue, int act, void* stream);
/* the split-bf16 (fp32-equivalent) form: A and W as hi / lo bf16 planes; a bf16 output leaves as two planes too */
int vscb200_gemm_split(const void* A_hi, const void* A_lo, const void* W_hi, const void* W_lo, const float* bias, void* C,
                       void* C_lo, int64_t M, int N, int K, int64_t lda, int64_t ldw, int64_t ldc, int epilogue, int act,
                       void* stream);
int vscb200_layernorm(const float* x, const float* gamma, const float* beta, void* y, int64_t rows, int width,
                      float eps, int out_bf16, void* stream);
/* qkv: [n*T, 3W] bf16 (q|k|v, heads contiguous 64-wide) -> out [n*T, W] bf16; head_dim 64 */
int vscb200_attention(const void* qkv_bf16, void* out_bf16, int n_frames, int T, int heads, int head_dim,
                      void* stream);
int vscb200_cast_f32_bf16(const float* x, void* y_bf16, int64_t count, void* stream);
/* x fp32 [count] -> hi = bf16(x), lo = bf16(x - hi) */
int vscb200_split_f32_bf16(const float* x, void* hi_bf16, void* lo_bf16, int64_t count, void* stream);
/* fp32 attention over n_segs segments of N consecutive rows (ViT frames), head_dim 32 / 64, q.k scaled by head_dim^-0.5;
 * qkv [n_segs*N, 3*heads*head_dim] as hi (+ lo, may be NULL) planes -> out hi (+ lo) [n_segs*N, heads*head_dim] */
int vscb200_attention_fp32(const void* qkv_hi, const void* qkv_lo, void* out_hi, void* out_lo, int64_t n_segs, int N,
                           int heads, int head_dim, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VSCB200_H_ */
