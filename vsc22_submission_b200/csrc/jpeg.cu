// Baseline JPEG decoding of a batch of equally sized frames on the device (the decode half).
//
// Reference: PIL.Image.open(io.BytesIO(jpeg_bytes)) in the frame loop of VSC22-Descriptor-Track-1st/infer/src/dataset.py:
// 137-141 (one zip of ffmpeg-written JPEG frames per video), i.e. libjpeg-turbo at its defaults.  The output is
// BIT-IDENTICAL to it (tests/test_gpu_jpeg.py against Pillow itself): the published IJG algorithms are restated --
//   entropy decoding   T.81 baseline sequential Huffman (jdhuff.c): 9-bit lookahead tables, byte-stuffing, restart intervals
//   inverse DCT        jpeg_idct_islow (jidctint.c): CONST_BITS 13, PASS1_BITS 2, zero-AC column / row shortcuts,
//                      range_limit[x & 1023]
//   chroma upsampling  h2v2_fancy_upsample / h2v1_fancy_upsample (jdsample.c triangle filters; replication when the
//                      downsampled width is <= 2, as jinit_upsampler selects)
//   colour conversion  ycc_rgb_convert (jdcolor.c, 16-bit fixed point)
// Scope = what ffmpeg's mjpeg encoder emits: 8-bit, YCbCr 4:2:0 / 4:2:2 / 4:4:4 or grey, one interleaved scan; anything
// else (progressive, arithmetic, CMYK, 12-bit) fails loudly with VSCB200_ERR_INVALID.
//
// Host: marker parsing of every frame (tables, geometry, the entropy-coded segment; restart intervals are located by
// scanning for RSTn), one descriptor per independent entropy segment, compressed bytes + descriptors staged and uploaded.
// Device:
//   jpeg_unstuff_kernel   one CTA per entropy segment: the zero byte behind every 0xFF is removed by a block-wide
//                         compaction, so that the decoder reads plain aligned 32-bit words
//   jpeg_huffman_kernel   one entropy segment (a whole frame, or one restart interval) per warp, decoded by lane 0 while
//                         the other lanes have loaded the frame's Huffman tables into shared memory; quantised
//                         coefficients (natural order, int16) into the zero-filled coefficient planes.  A segment is a
//                         serial bit stream: the parallelism is over frames (and restart intervals), so throughput
//                         grows with the number of frames per call until every SM holds a few dozen warps
//   jpeg_idct_kernel      one thread per 8 x 8 block: dequantise, two-pass integer IDCT in registers, uint8 samples into the
//                         component planes
//   jpeg_color_kernel     one thread per pixel: Y + fancy-upsampled Cb / Cr -> RGB, [n, H, W, 3] uint8 -- the layout
//                         vscb200_resize_normalize (resize.cu) takes
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "host_util.h"
#include "kernels.h"

namespace vscb200 {

namespace {

constexpr int kJpLook = 9;

struct JpHuff {                    // one DHT table, decoder form
  uint16_t look[1 << kJpLook];     // (code length << 8) | symbol for codes of <= 9 bits, 0 otherwise
  int32_t maxcode[18];             // largest code of each length (-1: none); [17] = sentinel
  int32_t valoff[17];              // vals index of the first code of each length minus that code
  uint8_t vals[256];
};

struct JpGeom {                    // identical for every frame of a batch
  int H, W, ncomp;
  int hs[3], vs[3];                // sampling factors
  int hmax, vmax;
  int mcux, mcuy;
  int bw[3], bh[3];                // blocks per row / column of each component (padded grid)
  int64_t coef_off[3];             // element offsets inside one frame's coefficient block
  int64_t coef_per_frame;
  int64_t plane_off[3];            // byte offsets inside one frame's sample planes
  int64_t plane_per_frame;
};

struct JpSegment {                 // one independently decodable entropy-coded segment
  int frame;
  int mcu0, n_mcu;
  uint32_t byte0, byte1;           // [byte0, byte1) of the batch's byte buffer (as in the file: byte-stuffed)
  uint32_t clean0;                 // offset of the segment's unstuffed bytes in the clean buffer (16-byte aligned)
  int16_t tab_dc[3], tab_ac[3];    // Huffman table indices per component
};

struct JpQuant { uint16_t q[3][64]; };       // per frame, natural order

__constant__ uint8_t c_zigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                     41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                     30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
const uint8_t h_zigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                              41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                              30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// ------------------------------------------------------------------------------------------------ entropy decoding
// Reads the UNSTUFFED bytes of one segment (jpeg_unstuff_kernel: 0xFF00 -> 0xFF, zero padding behind the data) as aligned
// 32-bit words: no per-byte loads and no marker tests on the decoder's critical path.
struct BitReader {
  const uint32_t* p;               // next word (the segment's clean bytes start 16-byte aligned)
  uint64_t acc;
  int n;
  __device__ __forceinline__ void fill() {                 // n <= 32 on entry
    const uint32_t w = __ldg(p++);
    acc = (acc << 32) | __byte_perm(w, 0, 0x0123);         // big-endian bit order
    n += 32;
  }
  __device__ __forceinline__ uint32_t peek(int k) {        // k <= 16
    if (n < k) fill();
    return static_cast<uint32_t>(acc >> (n - k)) & ((1u << k) - 1u);
  }
  __device__ __forceinline__ void skip(int k) { n -= k; }
  __device__ __forceinline__ uint32_t get(int k) {
    if (k == 0) return 0;
    const uint32_t v = peek(k);
    n -= k;
    return v;
  }
};

__device__ __forceinline__ int jp_decode_symbol(BitReader& br, const JpHuff& h) {
  const uint32_t idx = br.peek(kJpLook);
  const uint32_t e = h.look[idx];
  if (e) {
    br.skip(static_cast<int>(e >> 8));
    return static_cast<int>(e & 255u);
  }
  int32_t code = static_cast<int32_t>(idx);
  br.skip(kJpLook);
  int l = kJpLook;
  while (l < 17 && code > h.maxcode[l]) {            // maxcode[17] sentinel ends a corrupt stream
    code = (code << 1) | static_cast<int32_t>(br.get(1));
    ++l;
  }
  if (l > 16) return 0;
  return h.vals[(code + h.valoff[l]) & 255];
}

__device__ __forceinline__ int jp_extend(uint32_t v, int t) {
  return t == 0 ? 0 : (v >= (1u << (t - 1)) ? static_cast<int>(v) : static_cast<int>(v) - (1 << t) + 1);
}

// One CTA per segment: drop the zero byte behind every 0xFF (T.81 B.1.1.5 byte stuffing) with a block-wide compaction, and
// put 16 zero bytes behind the data (the decoder may read ahead; a truncated stream decodes zeros like jdhuff.c).
constexpr int kJpUnThreads = 256, kJpUnPer = 16;
__global__ void __launch_bounds__(kJpUnThreads)
jpeg_unstuff_kernel(const uint8_t* __restrict__ bytes, const JpSegment* __restrict__ segs, uint8_t* __restrict__ clean) {
  __shared__ int s_cnt[kJpUnThreads];
  __shared__ int s_base;
  const JpSegment sg = segs[blockIdx.x];
  const uint8_t* src = bytes + sg.byte0;
  const int len = static_cast<int>(sg.byte1 - sg.byte0);
  uint8_t* dst = clean + sg.clean0;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int t0 = 0; t0 < len; t0 += kJpUnThreads * kJpUnPer) {
    const int i0 = t0 + threadIdx.x * kJpUnPer;
    uint8_t v[kJpUnPer];
    int keep = 0, cnt = 0;
    uint8_t prev = i0 > 0 && i0 - 1 < len ? src[i0 - 1] : 0;
#pragma unroll
    for (int j = 0; j < kJpUnPer; ++j) {
      const int i = i0 + j;
      const uint8_t b = i < len ? src[i] : 0;
      const bool k = i < len && !(b == 0x00 && prev == 0xFF);
      if (k) { keep |= 1 << j; ++cnt; }
      v[j] = b;
      prev = b;
    }
    s_cnt[threadIdx.x] = cnt;
    __syncthreads();
    // exclusive scan of the per-thread counts (Hillis-Steele over 256 entries)
    for (int o = 1; o < kJpUnThreads; o <<= 1) {
      const int add = threadIdx.x >= o ? s_cnt[threadIdx.x - o] : 0;
      __syncthreads();
      s_cnt[threadIdx.x] += add;
      __syncthreads();
    }
    int at = s_base + s_cnt[threadIdx.x] - cnt;
#pragma unroll
    for (int j = 0; j < kJpUnPer; ++j)
      if (keep & (1 << j)) dst[at++] = v[j];
    __syncthreads();
    if (threadIdx.x == kJpUnThreads - 1) s_base += s_cnt[kJpUnThreads - 1];
    __syncthreads();
  }
  if (threadIdx.x < 16) dst[s_base + threadIdx.x] = 0;
}

constexpr int kJpTabsPerSeg = 6;          // DC / AC of up to three components, staged in shared memory per warp

__global__ void __launch_bounds__(32)
jpeg_huffman_kernel(const uint8_t* __restrict__ clean, const JpSegment* __restrict__ segs, const JpHuff* __restrict__ tabs,
                    JpGeom g, int16_t* __restrict__ coef) {
  __shared__ JpHuff s_tab[kJpTabsPerSeg];
  const JpSegment sg = segs[blockIdx.x];
  const int lane = threadIdx.x;
  for (int c = 0; c < g.ncomp; ++c) {               // all lanes copy the tables; lane 0 decodes
    const uint32_t* sd = reinterpret_cast<const uint32_t*>(tabs + sg.tab_dc[c]);
    const uint32_t* sa = reinterpret_cast<const uint32_t*>(tabs + sg.tab_ac[c]);
    uint32_t* dd = reinterpret_cast<uint32_t*>(&s_tab[2 * c]);
    uint32_t* da = reinterpret_cast<uint32_t*>(&s_tab[2 * c + 1]);
    for (int i = lane; i < static_cast<int>(sizeof(JpHuff) / 4); i += 32) { dd[i] = sd[i]; da[i] = sa[i]; }
  }
  __syncwarp();
  if (lane != 0) return;
  BitReader br;
  br.p = reinterpret_cast<const uint32_t*>(clean + sg.clean0); br.acc = 0; br.n = 0;
  int pred[3] = {0, 0, 0};
  int16_t* fc = coef + static_cast<int64_t>(sg.frame) * g.coef_per_frame;
  for (int m = sg.mcu0; m < sg.mcu0 + sg.n_mcu; ++m) {
    const int my = m / g.mcux, mx = m - my * g.mcux;
    for (int c = 0; c < g.ncomp; ++c) {
      const JpHuff& hd = s_tab[2 * c];
      const JpHuff& ha = s_tab[2 * c + 1];
      for (int by = 0; by < g.vs[c]; ++by) {
        for (int bx = 0; bx < g.hs[c]; ++bx) {
          int16_t* blk = fc + g.coef_off[c] + (static_cast<int64_t>(my * g.vs[c] + by) * g.bw[c] + mx * g.hs[c] + bx) * 64;
          const int t = jp_decode_symbol(br, hd);
          pred[c] += jp_extend(br.get(t), t);
          blk[0] = static_cast<int16_t>(pred[c]);
          int k = 1;
          while (k < 64) {
            const int rs = jp_decode_symbol(br, ha);
            const int r = rs >> 4, s = rs & 15;
            if (s == 0) {
              if (r != 15) break;
              k += 16;
              continue;
            }
            k += r;
            if (k > 63) break;                       // corrupt stream: stay inside the block
            blk[c_zigzag[k]] = static_cast<int16_t>(jp_extend(br.get(s), s));
            ++k;
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ jidctint.c
constexpr int kCB = 13, kP1 = 2;
constexpr int F_0_298631336 = 2446, F_0_390180644 = 3196, F_0_541196100 = 4433, F_0_765366865 = 6270, F_0_899976223 = 7373,
              F_1_175875602 = 9633, F_1_501321110 = 12299, F_1_847759065 = 15137, F_1_961570560 = 16069,
              F_2_053119869 = 16819, F_2_562915447 = 20995, F_3_072711026 = 25172;

__device__ __forceinline__ int jp_descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

__device__ __forceinline__ void jp_idct_1d(const int (&d)[8], int (&o)[8], int shift) {
  int z2 = d[2], z3 = d[6];
  int z1 = (z2 + z3) * F_0_541196100;
  const int tmp2 = z1 + z3 * (-F_1_847759065);
  const int tmp3 = z1 + z2 * F_0_765366865;
  const int tmp0 = (d[0] + d[4]) << kCB, tmp1 = (d[0] - d[4]) << kCB;
  const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  int t0 = d[7], t1 = d[5], t2 = d[3], t3 = d[1];
  z1 = t0 + t3; z2 = t1 + t2; z3 = t0 + t2;
  int z4 = t1 + t3;
  const int z5 = (z3 + z4) * F_1_175875602;
  t0 *= F_0_298631336; t1 *= F_2_053119869; t2 *= F_3_072711026; t3 *= F_1_501321110;
  z1 *= -F_0_899976223; z2 *= -F_2_562915447; z3 *= -F_1_961570560; z4 *= -F_0_390180644;
  z3 += z5; z4 += z5;
  t0 += z1 + z3; t1 += z2 + z4; t2 += z2 + z3; t3 += z1 + z4;
  o[0] = jp_descale(tmp10 + t3, shift); o[7] = jp_descale(tmp10 - t3, shift);
  o[1] = jp_descale(tmp11 + t2, shift); o[6] = jp_descale(tmp11 - t2, shift);
  o[2] = jp_descale(tmp12 + t1, shift); o[5] = jp_descale(tmp12 - t1, shift);
  o[3] = jp_descale(tmp13 + t0, shift); o[4] = jp_descale(tmp13 - t0, shift);
}

__device__ __forceinline__ uint8_t jp_range_limit(int x) {          // sample_range_limit + CENTERJSAMPLE at x & RANGE_MASK
  const int idx = x & 1023;
  return static_cast<uint8_t>(idx < 128 ? idx + 128 : (idx < 512 ? 255 : (idx < 896 ? 0 : idx - 896)));
}

__global__ void __launch_bounds__(128)
jpeg_idct_kernel(const int16_t* __restrict__ coef, const JpQuant* __restrict__ quant, JpGeom g, int64_t n_frames,
                 uint8_t* __restrict__ planes) {
  const int64_t blocks_per_frame = g.coef_per_frame / 64;
  const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= n_frames * blocks_per_frame) return;
  const int64_t f = gid / blocks_per_frame;
  int64_t b = gid - f * blocks_per_frame;
  int c = 0;
  while (c + 1 < g.ncomp && b >= g.coef_off[c + 1] / 64) ++c;
  b -= g.coef_off[c] / 64;
  const int by = static_cast<int>(b / g.bw[c]), bx = static_cast<int>(b - static_cast<int64_t>(by) * g.bw[c]);
  const int16_t* src = coef + f * g.coef_per_frame + g.coef_off[c] + b * 64;
  const uint16_t* q = quant[f].q[c];
  int ws[8][8];                                      // [row][col] after the column pass
#pragma unroll
  for (int col = 0; col < 8; ++col) {
    int d[8];
    bool zero_ac = true;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      d[r] = static_cast<int>(src[r * 8 + col]) * static_cast<int>(q[r * 8 + col]);
      if (r > 0 && d[r] != 0) zero_ac = false;
    }
    // jidctint.c tests the COEFFICIENTS for zero; a zero product of a non-zero coefficient does not exist (q >= 1)
    if (zero_ac) {
      const int dc = d[0] << kP1;
#pragma unroll
      for (int r = 0; r < 8; ++r) ws[r][col] = dc;
    } else {
      int o[8];
      jp_idct_1d(d, o, kCB - kP1);
#pragma unroll
      for (int r = 0; r < 8; ++r) ws[r][col] = o[r];
    }
  }
  const int stride = g.bw[c] * 8;
  uint8_t* dst = planes + f * g.plane_per_frame + g.plane_off[c] + static_cast<int64_t>(by * 8) * stride + bx * 8;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    bool zero_row = true;
#pragma unroll
    for (int col = 1; col < 8; ++col)
      if (ws[r][col] != 0) zero_row = false;
    uint8_t px[8];
    if (zero_row) {
      const uint8_t v = jp_range_limit(jp_descale(ws[r][0], kP1 + 3));
#pragma unroll
      for (int col = 0; col < 8; ++col) px[col] = v;
    } else {
      int o[8];
      jp_idct_1d(ws[r], o, kCB + kP1 + 3);
#pragma unroll
      for (int col = 0; col < 8; ++col) px[col] = jp_range_limit(o[col]);
    }
    *reinterpret_cast<uint2*>(dst + static_cast<int64_t>(r) * stride) =
        make_uint2(px[0] | (px[1] << 8) | (px[2] << 16) | (static_cast<uint32_t>(px[3]) << 24),
                   px[4] | (px[5] << 8) | (px[6] << 16) | (static_cast<uint32_t>(px[7]) << 24));
  }
}

// ------------------------------------------------------------------------------------------------ jdsample.c + jdcolor.c
// sample of chroma plane p (real extent ch x cw, row stride `stride`) upsampled to full-resolution pixel (Y, X)
__device__ __forceinline__ int jp_chroma(const uint8_t* __restrict__ p, int stride, int ch, int cw, int fh, int fv, int Y, int X) {
  if (fh == 1 && fv == 1) return p[static_cast<int64_t>(Y) * stride + X];
  if (cw <= 2 || (fh == 1 && fv == 2)) {                            // jinit_upsampler: plain replication
    return p[static_cast<int64_t>(Y / fv) * stride + X / fh];
  }
  const int cx = X >> 1;
  if (fv == 1) {                                                    // h2v1_fancy_upsample
    const uint8_t* row = p + static_cast<int64_t>(Y) * stride;
    const int v = row[cx];
    if ((X & 1) == 0) return cx == 0 ? v : (v * 3 + row[cx - 1] + 1) >> 2;
    return cx == cw - 1 ? v : (v * 3 + row[cx + 1] + 2) >> 2;
  }
  // h2v2_fancy_upsample: 3 * nearer row + farther row, then the same weights horizontally on the column sums
  const int cy = Y >> 1;
  const int far = (Y & 1) ? min(cy + 1, ch - 1) : max(cy - 1, 0);
  const uint8_t* r0 = p + static_cast<int64_t>(cy) * stride;
  const uint8_t* r1 = p + static_cast<int64_t>(far) * stride;
  const int cs = r0[cx] * 3 + r1[cx];
  if ((X & 1) == 0) {
    if (cx == 0) return (cs * 4 + 8) >> 4;
    return (cs * 3 + (r0[cx - 1] * 3 + r1[cx - 1]) + 8) >> 4;
  }
  if (cx == cw - 1) return (cs * 4 + 7) >> 4;
  return (cs * 3 + (r0[cx + 1] * 3 + r1[cx + 1]) + 7) >> 4;
}

__device__ __forceinline__ uint8_t jp_clamp8(int v) { return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v)); }

__global__ void __launch_bounds__(256)
jpeg_color_kernel(const uint8_t* __restrict__ planes, JpGeom g, int64_t n_frames, uint8_t* __restrict__ rgb) {
  const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t per = static_cast<int64_t>(g.H) * g.W;
  if (gid >= n_frames * per) return;
  const int64_t f = gid / per;
  const int64_t r = gid - f * per;
  const int Y = static_cast<int>(r / g.W), X = static_cast<int>(r - static_cast<int64_t>(Y) * g.W);
  const uint8_t* fp = planes + f * g.plane_per_frame;
  const int y = fp[g.plane_off[0] + static_cast<int64_t>(Y) * (g.bw[0] * 8) + X];
  uint8_t* o = rgb + gid * 3;
  if (g.ncomp == 1) { o[0] = o[1] = o[2] = static_cast<uint8_t>(y); return; }
  int cc[2];
#pragma unroll
  for (int c = 1; c < 3; ++c) {
    const int fh = g.hmax / g.hs[c], fv = g.vmax / g.vs[c];
    const int ch = (g.H * g.vs[c] + g.vmax - 1) / g.vmax, cw = (g.W * g.hs[c] + g.hmax - 1) / g.hmax;
    cc[c - 1] = jp_chroma(fp + g.plane_off[c], g.bw[c] * 8, ch, cw, fh, fv, Y, X);
  }
  const int cb = cc[0] - 128, cr = cc[1] - 128;
  // FIX(1.40200) = 91881, FIX(1.77200) = 116130, FIX(0.71414) = 46802, FIX(0.34414) = 22554; ONE_HALF = 32768
  const int R = y + ((91881 * cr + 32768) >> 16);
  const int G = y + ((-22554 * cb + 32768 - 46802 * cr) >> 16);
  const int B = y + ((116130 * cb + 32768) >> 16);
  o[0] = jp_clamp8(R); o[1] = jp_clamp8(G); o[2] = jp_clamp8(B);
}

// ------------------------------------------------------------------------------------------------ host: marker parsing
struct ParsedFrame {
  int H = 0, W = 0, ncomp = 0;
  int hs[3] = {1, 1, 1}, vs[3] = {1, 1, 1}, tq[3] = {0, 0, 0}, id[3] = {0, 0, 0};
  int td[3] = {0, 0, 0}, ta[3] = {0, 0, 0};
  uint16_t qt[4][64];
  bool have_qt[4] = {false, false, false, false};
  std::vector<uint8_t> dht[2][4];          // [class][id]: 16 counts + values
  int dri = 0;
  size_t scan0 = 0, scan1 = 0;             // entropy-coded bytes [scan0, scan1) of the file
};

bool parse_jpeg(const uint8_t* d, size_t n, ParsedFrame* pf, std::string* err) {
  auto fail = [&](const char* m) { *err = m; return false; };
  if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) return fail("not a JPEG file (no SOI marker)");
  size_t i = 2;
  bool have_sof = false, have_sos = false;
  int adobe = -1;
  while (i + 1 < n && !have_sos) {
    if (d[i] != 0xFF) return fail("corrupt JPEG: marker expected");
    while (i < n && d[i] == 0xFF) ++i;
    if (i >= n) break;
    const int m = d[i++];
    if (m == 0xD9) break;
    if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
    if (i + 2 > n) return fail("corrupt JPEG: truncated segment");
    const size_t L = (static_cast<size_t>(d[i]) << 8) | d[i + 1];
    if (L < 2 || i + L > n) return fail("corrupt JPEG: bad segment length");
    const uint8_t* s = d + i + 2;
    const size_t sl = L - 2;
    i += L;
    if (m == 0xDB) {
      size_t j = 0;
      while (j < sl) {
        const int pq = s[j] >> 4, tq = s[j] & 15;
        ++j;
        if (tq > 3 || j + (pq ? 128 : 64) > sl) return fail("corrupt JPEG: bad DQT");
        for (int k = 0; k < 64; ++k) {
          const int v = pq ? ((s[j + 2 * k] << 8) | s[j + 2 * k + 1]) : s[j + k];
          pf->qt[tq][h_zigzag[k]] = static_cast<uint16_t>(v);
        }
        pf->have_qt[tq] = true;
        j += pq ? 128 : 64;
      }
    } else if (m == 0xC4) {
      size_t j = 0;
      while (j < sl) {
        const int tc = s[j] >> 4, th = s[j] & 15;
        if (tc > 1 || th > 3 || j + 17 > sl) return fail("corrupt JPEG: bad DHT");
        int cnt = 0;
        for (int k = 0; k < 16; ++k) cnt += s[j + 1 + k];
        if (cnt > 256 || j + 17 + cnt > sl) return fail("corrupt JPEG: bad DHT");
        pf->dht[tc][th].assign(s + j + 1, s + j + 17 + cnt);
        j += 17 + cnt;
      }
    } else if (m == 0xC0 || m == 0xC1) {
      if (sl < 6 || s[0] != 8) return fail("unsupported JPEG: only 8-bit samples");
      pf->H = (s[1] << 8) | s[2];
      pf->W = (s[3] << 8) | s[4];
      pf->ncomp = s[5];
      if (pf->ncomp != 1 && pf->ncomp != 3) return fail("unsupported JPEG: only grey or YCbCr (1 or 3 components)");
      if (sl < static_cast<size_t>(6 + 3 * pf->ncomp)) return fail("corrupt JPEG: bad SOF");
      for (int c = 0; c < pf->ncomp; ++c) {
        pf->id[c] = s[6 + 3 * c];
        pf->hs[c] = s[7 + 3 * c] >> 4;
        pf->vs[c] = s[7 + 3 * c] & 15;
        pf->tq[c] = s[8 + 3 * c];
        if (pf->tq[c] > 3) return fail("corrupt JPEG: bad quantisation table id");
      }
      have_sof = true;
    } else if (m == 0xC2 || m == 0xC3 || (m >= 0xC5 && m <= 0xC7) || (m >= 0xC9 && m <= 0xCB) || (m >= 0xCD && m <= 0xCF)) {
      return fail("unsupported JPEG: progressive / lossless / arithmetic coding (the ingest path decodes baseline files)");
    } else if (m == 0xDD) {
      if (sl < 2) return fail("corrupt JPEG: bad DRI");
      pf->dri = (s[0] << 8) | s[1];
    } else if (m == 0xEE && sl >= 12 && memcmp(s, "Adobe", 5) == 0) {
      adobe = s[11];
    } else if (m == 0xDA) {
      if (!have_sof) return fail("corrupt JPEG: SOS before SOF");
      if (sl < 1 || s[0] != pf->ncomp || sl < static_cast<size_t>(1 + 2 * pf->ncomp)) return fail("unsupported JPEG: only single-scan (interleaved) files");
      for (int k = 0; k < pf->ncomp; ++k) {
        const int cid = s[1 + 2 * k];
        int c = -1;
        for (int q = 0; q < pf->ncomp; ++q)
          if (pf->id[q] == cid) c = q;
        if (c != k) return fail("unsupported JPEG: scan components out of frame order");
        pf->td[c] = s[2 + 2 * k] >> 4;
        pf->ta[c] = s[2 + 2 * k] & 15;
        if (pf->td[c] > 3 || pf->ta[c] > 3) return fail("corrupt JPEG: bad Huffman table id");
      }
      // end of the entropy-coded segment: the first 0xFF that is neither stuffed (0xFF00) nor a restart marker
      size_t j = i;
      for (;;) {
        const void* hit = j < n ? memchr(d + j, 0xFF, n - j) : nullptr;
        if (!hit) { j = n; break; }
        j = static_cast<size_t>(static_cast<const uint8_t*>(hit) - d);
        if (j + 1 >= n) { j = n; break; }
        if (d[j + 1] != 0x00 && !(d[j + 1] >= 0xD0 && d[j + 1] <= 0xD7)) break;
        j += 2;
      }
      pf->scan0 = i;
      pf->scan1 = j;
      have_sos = true;
    }
  }
  if (!have_sof || !have_sos) return fail("corrupt JPEG: no frame header or no scan");
  if (pf->H <= 0 || pf->W <= 0) return fail("corrupt JPEG: empty frame");
  if (pf->ncomp == 3 && adobe >= 0 && adobe != 1) return fail("unsupported JPEG: Adobe RGB / CMYK colour transform");
  for (int c = 0; c < pf->ncomp; ++c) {
    if (!pf->have_qt[pf->tq[c]] || pf->dht[0][pf->td[c]].empty() || pf->dht[1][pf->ta[c]].empty()) return fail("corrupt JPEG: missing table");
    if (pf->hs[c] < 1 || pf->hs[c] > 2 || pf->vs[c] < 1 || pf->vs[c] > 2) return fail("unsupported JPEG: sampling factors above 2");
  }
  if (pf->ncomp == 3) {
    if (pf->hs[1] != 1 || pf->vs[1] != 1 || pf->hs[2] != 1 || pf->vs[2] != 1) return fail("unsupported JPEG: subsampled luma / oversampled chroma");
    if (pf->hs[0] == 1 && pf->vs[0] == 2) return fail("unsupported JPEG: 4:4:0 subsampling");
  } else {
    pf->hs[0] = pf->vs[0] = 1;                       // a single-component scan is not interleaved: MCU = one block
  }
  return true;
}

void build_huff(const std::vector<uint8_t>& dht, JpHuff* h) {
  memset(h, 0, sizeof(*h));
  int code = 0, k = 0;
  for (int l = 1; l <= 16; ++l) {
    const int cnt = dht[l - 1];
    h->valoff[l] = k - code;
    if (cnt) {
      for (int q = 0; q < cnt; ++q, ++code, ++k) {
        if (l <= kJpLook) {
          const int sym = dht[16 + k];
          const int base = code << (kJpLook - l);
          for (int f = 0; f < (1 << (kJpLook - l)); ++f) h->look[base + f] = static_cast<uint16_t>((l << 8) | sym);
        }
      }
      h->maxcode[l] = code - 1;
    } else {
      h->maxcode[l] = -1;
    }
    code <<= 1;
  }
  h->maxcode[17] = 0x7FFFFFFF;
  for (size_t q = 16; q < dht.size() && q - 16 < 256; ++q) h->vals[q - 16] = dht[q];
}

}  // namespace

}  // namespace vscb200

// jpeg_ptrs[i] / jpeg_sizes[i]: the n files (host memory).  rgb_dev: [n, H, W, 3] uint8 on the device, H x W = the size of
// the FIRST frame (returned through h_out / w_out when rgb_dev is NULL: a size query that decodes nothing).
extern "C" int vscb200_jpeg_decode(const uint8_t* const* jpeg_ptrs, const uint64_t* jpeg_sizes, int64_t n, uint8_t* rgb_dev,
                                   int* h_out, int* w_out, void* stream_v) {
  using namespace vscb200;
  VSCB_REQUIRE(n >= 0 && (n == 0 || (jpeg_ptrs && jpeg_sizes)), "jpeg_decode: null argument");
  if (n == 0) return VSCB200_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream_v);
  std::vector<ParsedFrame> pf(static_cast<size_t>(n));
  std::string err;
  for (int64_t i = 0; i < n; ++i) {
    if (!parse_jpeg(jpeg_ptrs[i], static_cast<size_t>(jpeg_sizes[i]), &pf[i], &err)) {
      set_last_error("jpeg_decode: frame " + std::to_string(i) + ": " + err);
      return VSCB200_ERR_INVALID;
    }
    const ParsedFrame &a = pf[0], &b = pf[i];
    bool same = a.H == b.H && a.W == b.W && a.ncomp == b.ncomp;
    for (int c = 0; same && c < a.ncomp; ++c) same = a.hs[c] == b.hs[c] && a.vs[c] == b.vs[c];
    if (!same) {
      set_last_error("jpeg_decode: frame " + std::to_string(i) + " differs from frame 0 in size or chroma subsampling (one call decodes the equally sized frames of one video)");
      return VSCB200_ERR_INVALID;
    }
  }
  if (h_out) *h_out = pf[0].H;
  if (w_out) *w_out = pf[0].W;
  if (!rgb_dev) return VSCB200_OK;

  JpGeom g = {};
  g.H = pf[0].H; g.W = pf[0].W; g.ncomp = pf[0].ncomp;
  g.hmax = 1; g.vmax = 1;
  for (int c = 0; c < g.ncomp; ++c) {
    g.hs[c] = pf[0].hs[c]; g.vs[c] = pf[0].vs[c];
    g.hmax = std::max(g.hmax, g.hs[c]); g.vmax = std::max(g.vmax, g.vs[c]);
  }
  g.mcux = (g.W + 8 * g.hmax - 1) / (8 * g.hmax);
  g.mcuy = (g.H + 8 * g.vmax - 1) / (8 * g.vmax);
  int64_t co = 0, po = 0;
  for (int c = 0; c < g.ncomp; ++c) {
    g.bw[c] = g.mcux * g.hs[c]; g.bh[c] = g.mcuy * g.vs[c];
    g.coef_off[c] = co; g.plane_off[c] = po;
    co += static_cast<int64_t>(g.bw[c]) * g.bh[c] * 64;
    po += static_cast<int64_t>(g.bw[c]) * g.bh[c] * 64;
  }
  g.coef_per_frame = co; g.plane_per_frame = po;
  VSCB_REQUIRE(static_cast<double>(n) * co < 2.0e10, "jpeg_decode: batch too large");

  // ---- host staging: entropy-coded bytes back to back, one descriptor per independent segment, distinct Huffman tables
  std::vector<JpHuff> tabs;
  std::vector<std::vector<uint8_t>> tab_src;
  auto table_index = [&](const std::vector<uint8_t>& dht) {
    for (size_t t = 0; t < tab_src.size(); ++t)
      if (tab_src[t] == dht) return static_cast<int>(t);
    tab_src.push_back(dht);
    tabs.emplace_back();
    build_huff(dht, &tabs.back());
    return static_cast<int>(tabs.size() - 1);
  };
  std::vector<JpSegment> segs;
  std::vector<JpQuant> quant(static_cast<size_t>(n));
  size_t total_bytes = 0;
  for (int64_t i = 0; i < n; ++i) total_bytes += pf[i].scan1 - pf[i].scan0 + 8;
  VSCB_REQUIRE(total_bytes < (1ull << 32), "jpeg_decode: more than 4 GiB of compressed data in one call");
  const size_t bytes_al = (total_bytes + 255) & ~static_cast<size_t>(255);
  const int total_mcu = g.mcux * g.mcuy;
  // first pass: descriptors (byte offsets are known from the parse)
  size_t off = 0;
  for (int64_t i = 0; i < n; ++i) {
    const ParsedFrame& f = pf[i];
    const uint8_t* d = jpeg_ptrs[i];
    for (int c = 0; c < g.ncomp; ++c) memcpy(quant[i].q[c], f.qt[f.tq[c]], sizeof(uint16_t) * 64);
    JpSegment sg = {};
    sg.frame = static_cast<int>(i);
    for (int c = 0; c < g.ncomp; ++c) {
      sg.tab_dc[c] = static_cast<int16_t>(table_index(f.dht[0][f.td[c]]));
      sg.tab_ac[c] = static_cast<int16_t>(table_index(f.dht[1][f.ta[c]]));
    }
    if (f.dri <= 0) {
      sg.mcu0 = 0; sg.n_mcu = total_mcu;
      sg.byte0 = static_cast<uint32_t>(off); sg.byte1 = static_cast<uint32_t>(off + (f.scan1 - f.scan0));
      segs.push_back(sg);
    } else {
      size_t p0 = f.scan0;
      int mcu = 0;
      while (mcu < total_mcu) {
        size_t p1 = p0;                              // the next RSTn (or the end of the scan)
        for (;;) {
          const void* hit = p1 < f.scan1 ? memchr(d + p1, 0xFF, f.scan1 - p1) : nullptr;
          if (!hit) { p1 = f.scan1; break; }
          p1 = static_cast<size_t>(static_cast<const uint8_t*>(hit) - d);
          if (p1 + 1 >= f.scan1) { p1 = f.scan1; break; }
          if (d[p1 + 1] >= 0xD0 && d[p1 + 1] <= 0xD7) break;
          p1 += 2;
        }
        sg.mcu0 = mcu; sg.n_mcu = std::min(f.dri, total_mcu - mcu);
        sg.byte0 = static_cast<uint32_t>(off + (p0 - f.scan0)); sg.byte1 = static_cast<uint32_t>(off + (p1 - f.scan0));
        segs.push_back(sg);
        mcu += sg.n_mcu;
        p0 = p1 < f.scan1 ? p1 + 2 : f.scan1;
      }
    }
    off += f.scan1 - f.scan0 + 8;
  }
  // unstuffed copies of the segments: 16-byte aligned starts, room for the 16 zero bytes behind each
  size_t clean_total = 0;
  for (JpSegment& sg : segs) {
    sg.clean0 = static_cast<uint32_t>(clean_total);
    clean_total += ((static_cast<size_t>(sg.byte1 - sg.byte0) + 16 + 15) & ~static_cast<size_t>(15));
  }
  clean_total += 256;
  VSCB_REQUIRE(clean_total < (1ull << 32), "jpeg_decode: more than 4 GiB of compressed data in one call");
  // one staging block: bytes | segments | tables | quant
  const size_t seg_bytes = segs.size() * sizeof(JpSegment), tab_bytes = tabs.size() * sizeof(JpHuff), q_bytes = quant.size() * sizeof(JpQuant);
  auto al = [](size_t v) { return (v + 255) & ~static_cast<size_t>(255); };
  const size_t o_seg = bytes_al, o_tab = o_seg + al(seg_bytes), o_q = o_tab + al(tab_bytes), stage_total = o_q + al(q_bytes);
  // page-locked staging block, kept per thread and grown on demand (cudaMallocHost costs milliseconds)
  static thread_local uint8_t* host = nullptr;
  static thread_local size_t host_cap = 0;
  if (host_cap < stage_total) {
    if (host) cudaFreeHost(host);
    host = nullptr;
    host_cap = 0;
    VSCB_CUDA_OK(cudaMallocHost(&host, stage_total + (stage_total >> 2)));
    host_cap = stage_total + (stage_total >> 2);
  }
  off = 0;
  for (int64_t i = 0; i < n; ++i) {
    const size_t len = pf[i].scan1 - pf[i].scan0;
    memcpy(host + off, jpeg_ptrs[i] + pf[i].scan0, len);
    memset(host + off + len, 0, 8);
    off += len + 8;
  }
  memcpy(host + o_seg, segs.data(), seg_bytes);
  memcpy(host + o_tab, tabs.data(), tab_bytes);
  memcpy(host + o_q, quant.data(), q_bytes);

  const size_t coef_bytes = static_cast<size_t>(n) * g.coef_per_frame * sizeof(int16_t);
  const size_t plane_bytes = static_cast<size_t>(n) * g.plane_per_frame;
  uint8_t* dev = nullptr;
  int rc = pool_alloc(reinterpret_cast<void**>(&dev), al(stage_total) + al(clean_total) + coef_bytes + plane_bytes + 512, s);
  if (rc) return rc;
  uint8_t* clean = dev + al(stage_total);
  int16_t* coef = reinterpret_cast<int16_t*>(clean + al(clean_total));
  uint8_t* planes = reinterpret_cast<uint8_t*>(coef) + coef_bytes;
  cudaError_t e = cudaMemcpyAsync(dev, host, stage_total, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemsetAsync(coef, 0, coef_bytes, s);
  if (e == cudaSuccess) {
    jpeg_unstuff_kernel<<<static_cast<unsigned>(segs.size()), kJpUnThreads, 0, s>>>(dev, reinterpret_cast<const JpSegment*>(dev + o_seg), clean);
    jpeg_huffman_kernel<<<static_cast<unsigned>(segs.size()), 32, 0, s>>>(clean, reinterpret_cast<const JpSegment*>(dev + o_seg),
                                                                         reinterpret_cast<const JpHuff*>(dev + o_tab), g, coef);
    const int64_t nblocks = n * (g.coef_per_frame / 64);
    jpeg_idct_kernel<<<static_cast<unsigned>((nblocks + 127) / 128), 128, 0, s>>>(coef, reinterpret_cast<const JpQuant*>(dev + o_q), g, n, planes);
    const int64_t npix = n * static_cast<int64_t>(g.H) * g.W;
    jpeg_color_kernel<<<static_cast<unsigned>((npix + 255) / 256), 256, 0, s>>>(planes, g, n, rgb_dev);
    count_launch(4);
    e = cudaGetLastError();
  }
  // the staging block is reused by the next call: the upload must have left it
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  pool_free(dev, s);
  if (e != cudaSuccess) {
    set_last_error(std::string("jpeg_decode: ") + cudaGetErrorString(e));
    return VSCB200_ERR_CUDA;
  }
  return VSCB200_OK;
}
