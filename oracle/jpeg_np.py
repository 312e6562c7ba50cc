"""ORACLE (test infrastructure, never on the product path): baseline JPEG decoding restated in numpy / plain Python.

What it restates: the decoder the reference's frame ingest calls -- ``PIL.Image.open(io.BytesIO(jpeg_bytes))`` followed by
``Resize / ToTensor`` (VSC22-Descriptor-Track-1st/infer/src/dataset.py:137-141; the frames are the JPEG files ffmpeg wrote
into one zip per video).  Pillow decodes with libjpeg-turbo (not part of /root/reference; Pillow 12.2.0 bundles
libjpeg-turbo 3.x here) at its defaults, i.e. the published algorithms of the IJG library:

* entropy decoding: ITU-T T.81 baseline sequential Huffman (jdhuff.c), restart intervals included;
* dequantisation + inverse DCT: ``jpeg_idct_islow`` (jidctint.c: CONST_BITS 13, PASS1_BITS 2, the 12 FIX_ constants, the
  all-zero-AC column / row shortcuts, ``range_limit[x & RANGE_MASK]``); libjpeg-turbo's SIMD versions are bit-exact with it;
* chroma upsampling: ``h2v2_fancy_upsample`` / ``h2v1_fancy_upsample`` (jdsample.c: triangle filter 3/4 + 1/4 with the
  +8 / +7 and +1 / +2 rounding constants), plain replication for other ratios is NOT needed by ffmpeg's output;
* colour conversion: ``ycc_rgb_convert`` (jdcolor.c: 16-bit fixed-point tables, SCALEBITS 16).

Pinned by tests/test_oracle_jpeg.py: bit-identical to Pillow on synthetic frames over qualities 30..95, 4:2:0 / 4:2:2 / 4:4:4
subsampling, odd sizes, restart intervals and optimised Huffman tables.  Scope = what ffmpeg's mjpeg encoder emits
(8-bit, 3-component YCbCr or 1-component grey, baseline); progressive / arithmetic / CMYK files raise ValueError.
"""
from __future__ import annotations

import numpy as np

ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14,
                   21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53,
                   60, 61, 54, 47, 55, 62, 63], dtype=np.int64)           # zigzag position -> natural (row-major) index


def parse(data: bytes) -> dict:
    """Marker segments of a baseline JPEG -> tables, frame header, the entropy-coded bytes of the (single) scan."""
    if data[:2] != b"\xff\xd8":
        raise ValueError("not a JPEG (no SOI)")
    qt, dc, ac = {}, {}, {}
    frame, scan, dri, adobe = None, None, 0, None
    i = 2
    while i < len(data):
        if data[i] != 0xFF:
            raise ValueError("marker expected")
        while data[i] == 0xFF:
            i += 1
        m = data[i]
        i += 1
        if m == 0xD9:
            break
        if m == 0x01 or 0xD0 <= m <= 0xD7:
            continue
        L = (data[i] << 8) | data[i + 1]
        seg = data[i + 2:i + L]
        i += L
        if m == 0xDB:                                   # DQT
            j = 0
            while j < len(seg):
                pq, tq = seg[j] >> 4, seg[j] & 15
                j += 1
                if pq:
                    vals = [(seg[j + 2 * k] << 8) | seg[j + 2 * k + 1] for k in range(64)]
                    j += 128
                else:
                    vals = list(seg[j:j + 64])
                    j += 64
                q = np.zeros(64, dtype=np.int64)
                q[ZIGZAG] = vals                        # stored in zigzag order -> natural order
                qt[tq] = q
        elif m == 0xC4:                                 # DHT
            j = 0
            while j < len(seg):
                tc, th = seg[j] >> 4, seg[j] & 15
                counts = list(seg[j + 1:j + 17])
                n = sum(counts)
                vals = list(seg[j + 17:j + 17 + n])
                j += 17 + n
                (ac if tc else dc)[th] = (counts, vals)
        elif m in (0xC0, 0xC1):                         # SOF0 / SOF1: baseline / extended sequential Huffman
            if seg[0] != 8:
                raise ValueError("only 8-bit samples")
            h, w, nc = (seg[1] << 8) | seg[2], (seg[3] << 8) | seg[4], seg[5]
            comps = [dict(id=seg[6 + 3 * c], h=seg[7 + 3 * c] >> 4, v=seg[7 + 3 * c] & 15, tq=seg[8 + 3 * c]) for c in range(nc)]
            frame = dict(h=h, w=w, comps=comps)
        elif m in (0xC2, 0xC3, 0xC5, 0xC6, 0xC7, 0xC9, 0xCA, 0xCB, 0xCD, 0xCE, 0xCF):
            raise ValueError("progressive / lossless / arithmetic JPEG is outside the ingest path's scope")
        elif m == 0xDD:
            dri = (seg[0] << 8) | seg[1]
        elif m == 0xEE and seg[:5] == b"Adobe":
            adobe = seg[11]
        elif m == 0xDA:                                 # SOS: the entropy-coded segment follows
            ns = seg[0]
            sel = {seg[1 + 2 * c]: (seg[2 + 2 * c] >> 4, seg[2 + 2 * c] & 15) for c in range(ns)}
            if frame is None or ns != len(frame["comps"]):
                raise ValueError("only single-scan (interleaved) baseline files")
            j = i
            while not (data[j] == 0xFF and data[j + 1] != 0x00 and not (0xD0 <= data[j + 1] <= 0xD7)):
                j += 1
            scan = dict(sel=sel, data=data[i:j])
            i = j
    if frame is None or scan is None:
        raise ValueError("no frame / scan")
    if len(frame["comps"]) not in (1, 3):
        raise ValueError("only grey or YCbCr")
    if adobe is not None and adobe != 1 and len(frame["comps"]) == 3:
        raise ValueError("Adobe RGB / CMYK colour transforms are outside the ingest path's scope")
    return dict(frame=frame, qt=qt, dc=dc, ac=ac, scan=scan, dri=dri)


def _huff_table(counts, vals):
    """T.81 Annex C: code -> (length, value) via (mincode, maxcode, valptr) per length."""
    code, k = 0, 0
    mincode, maxcode, valptr = [0] * 17, [-1] * 17, [0] * 17
    for l in range(1, 17):
        if counts[l - 1]:
            valptr[l], mincode[l] = k, code
            code += counts[l - 1]
            k += counts[l - 1]
            maxcode[l] = code - 1
        code <<= 1
    return mincode, maxcode, valptr, vals


class _Bits:
    def __init__(self, data: bytes):
        self.d, self.i, self.acc, self.n, self.marker = data, 0, 0, 0, False

    def _fill(self):
        while self.n <= 24:
            b = 0
            if not self.marker and self.i < len(self.d):
                b = self.d[self.i]
                self.i += 1
                if b == 0xFF:
                    nxt = self.d[self.i] if self.i < len(self.d) else 0xD9
                    if nxt == 0:
                        self.i += 1                        # stuffed zero byte
                    else:                                  # a marker: stop here and feed zeros (jdhuff.c does the same)
                        self.i -= 1
                        self.marker = True
                        b = 0
            self.acc = ((self.acc << 8) | b) & 0xFFFFFFFFFF
            self.n += 8

    def get(self, k: int) -> int:
        if k == 0:
            return 0
        if self.n < k:
            self._fill()
        self.n -= k
        return (self.acc >> self.n) & ((1 << k) - 1)

    def decode(self, tab) -> int:
        mincode, maxcode, valptr, vals = tab
        code = 0
        for l in range(1, 17):
            code = (code << 1) | self.get(1)
            if maxcode[l] >= 0 and code <= maxcode[l] and code >= mincode[l]:
                return vals[valptr[l] + code - mincode[l]]
        raise ValueError("bad Huffman code")

    def restart(self):
        """Drop the bits left in the accumulator (padding) and skip the RSTn marker."""
        self.acc, self.n = 0, 0
        if not self.marker:                                # the marker was not reached by the prefetch yet
            while self.i < len(self.d) and not (self.d[self.i] == 0xFF and 0xD0 <= self.d[self.i + 1] <= 0xD7):
                self.i += 1
        self.i += 2
        self.marker = False


def _extend(v: int, t: int) -> int:
    return v if t == 0 or v >= (1 << (t - 1)) else v - (1 << t) + 1


def decode_coefficients(p: dict):
    """-> per component int32 [blocks_y, blocks_x, 64] (natural order, quantised), over the PADDED block grid."""
    fr = p["frame"]
    comps = fr["comps"]
    hmax, vmax = max(c["h"] for c in comps), max(c["v"] for c in comps)
    mcux, mcuy = -(-fr["w"] // (8 * hmax)), -(-fr["h"] // (8 * vmax))
    if len(comps) == 1:                                     # a single-component scan is not interleaved: MCU = one block
        comps[0]["h"] = comps[0]["v"] = 1
        hmax = vmax = 1
        mcux, mcuy = -(-fr["w"] // 8), -(-fr["h"] // 8)
    coef = [np.zeros((mcuy * c["v"], mcux * c["h"], 64), dtype=np.int32) for c in comps]
    dct = {k: _huff_table(*v) for k, v in p["dc"].items()}
    act = {k: _huff_table(*v) for k, v in p["ac"].items()}
    bits = _Bits(p["scan"]["data"])
    pred = [0] * len(comps)
    n_mcu = 0
    for my in range(mcuy):
        for mx in range(mcux):
            if p["dri"] and n_mcu and n_mcu % p["dri"] == 0:
                bits.restart()
                pred = [0] * len(comps)
            n_mcu += 1
            for ci, c in enumerate(comps):
                td, ta = p["scan"]["sel"][c["id"]]
                for by in range(c["v"]):
                    for bx in range(c["h"]):
                        blk = coef[ci][my * c["v"] + by, mx * c["h"] + bx]
                        t = bits.decode(dct[td])
                        pred[ci] += _extend(bits.get(t), t)
                        blk[0] = pred[ci]
                        k = 1
                        while k < 64:
                            rs = bits.decode(act[ta])
                            r, s = rs >> 4, rs & 15
                            if s == 0:
                                if r != 15:
                                    break
                                k += 16
                                continue
                            k += r
                            blk[ZIGZAG[k]] = _extend(bits.get(s), s)
                            k += 1
    return coef, (hmax, vmax)


# ---- jidctint.c: jpeg_idct_islow ----------------------------------------------------------------------------------
CONST_BITS, PASS1_BITS = 13, 2
F_0_298631336, F_0_390180644, F_0_541196100, F_0_765366865 = 2446, 3196, 4433, 6270
F_0_899976223, F_1_175875602, F_1_501321110, F_1_847759065 = 7373, 9633, 12299, 15137
F_1_961570560, F_2_053119869, F_2_562915447, F_3_072711026 = 16069, 16819, 20995, 25172


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def _range_limit(x):
    """sample_range_limit + CENTERJSAMPLE indexed by x & RANGE_MASK (1023): the IDCT's output stage."""
    idx = x & 1023
    return np.where(idx < 128, idx + 128, np.where(idx < 512, 255, np.where(idx < 896, 0, idx - 896))).astype(np.uint8)


def _idct_1d(d0, d1, d2, d3, d4, d5, d6, d7, shift):
    z2, z3 = d2, d6
    z1 = (z2 + z3) * F_0_541196100
    tmp2 = z1 + z3 * (-F_1_847759065)
    tmp3 = z1 + z2 * F_0_765366865
    tmp0 = (d0 + d4) << CONST_BITS
    tmp1 = (d0 - d4) << CONST_BITS
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    t0, t1, t2, t3 = d7, d5, d3, d1
    z1, z2, z3, z4 = t0 + t3, t1 + t2, t0 + t2, t1 + t3
    z5 = (z3 + z4) * F_1_175875602
    t0, t1, t2, t3 = t0 * F_0_298631336, t1 * F_2_053119869, t2 * F_3_072711026, t3 * F_1_501321110
    z1, z2, z3, z4 = z1 * (-F_0_899976223), z2 * (-F_2_562915447), z3 * (-F_1_961570560), z4 * (-F_0_390180644)
    z3, z4 = z3 + z5, z4 + z5
    t0, t1, t2, t3 = t0 + z1 + z3, t1 + z2 + z4, t2 + z2 + z3, t3 + z1 + z4
    return [_descale(tmp10 + t3, shift), _descale(tmp11 + t2, shift), _descale(tmp12 + t1, shift), _descale(tmp13 + t0, shift),
            _descale(tmp13 - t0, shift), _descale(tmp12 - t1, shift), _descale(tmp11 - t2, shift), _descale(tmp10 - t3, shift)]


def idct_islow(coef: np.ndarray, q: np.ndarray) -> np.ndarray:
    """coef int [..., 64] (natural order, quantised), q [64] -> uint8 [..., 8, 8].  The all-zero-AC shortcuts of the C code
    give the same values as the full butterflies EXCEPT for their rounding: they are restated explicitly."""
    x = (coef.astype(np.int64) * q).reshape(coef.shape[:-1] + (8, 8))            # [.., row, col]
    # pass 1: columns -> workspace scaled by 2^PASS1_BITS
    cols = [x[..., r, :] for r in range(8)]
    ws = np.stack(_idct_1d(*cols, CONST_BITS - PASS1_BITS), axis=-2)             # [.., row, col]
    zero_ac = np.all(x[..., 1:, :] == 0, axis=-2)                                # per column
    dcval = x[..., 0, :] << PASS1_BITS
    ws = np.where(zero_ac[..., None, :], dcval[..., None, :], ws)
    # pass 2: rows
    rows = [ws[..., :, c] for c in range(8)]
    out = np.stack(_idct_1d(*rows, CONST_BITS + PASS1_BITS + 3), axis=-1)        # [.., row, col]
    zero_row = np.all(ws[..., :, 1:] == 0, axis=-1)                              # per row
    dcrow = _descale(ws[..., :, 0], PASS1_BITS + 3)
    out = np.where(zero_row[..., None], dcrow[..., None], out)
    return _range_limit(out)


# ---- jdsample.c: fancy (triangle) upsampling --------------------------------------------------------------------------
def _h2v1_fancy_rows(p: np.ndarray) -> np.ndarray:
    """[rows, w] -> [rows, 2w] with libjpeg's horizontal 3/4 + 1/4 filter (h2v1_fancy_upsample)."""
    p = p.astype(np.int32)
    rows, w = p.shape
    out = np.zeros((rows, 2 * w), dtype=np.int32)
    if w == 1:
        out[:, 0] = out[:, 1] = p[:, 0]
        return out.astype(np.uint8)
    left = np.concatenate([p[:, :1], p[:, :-1]], axis=1)
    right = np.concatenate([p[:, 1:], p[:, -1:]], axis=1)
    out[:, 0::2] = (p * 3 + left + 1) >> 2
    out[:, 1::2] = (p * 3 + right + 2) >> 2
    out[:, 0] = p[:, 0]
    out[:, -1] = p[:, -1]
    return out.astype(np.uint8)


def _h2v2_fancy(p: np.ndarray) -> np.ndarray:
    """[h, w] -> [2h, 2w] (h2v2_fancy_upsample): vertically 3 * nearer row + farther row, horizontally the same weights on
    those column sums, (x * 4 + 8) >> 4 at the left edge, (x * 4 + 7) >> 4 at the right edge."""
    p = p.astype(np.int32)
    h, w = p.shape
    up = np.concatenate([p[:1], p[:-1]], axis=0)            # row above (replicated at the top)
    dn = np.concatenate([p[1:], p[-1:]], axis=0)
    out = np.zeros((2 * h, 2 * w), dtype=np.int32)
    for v, other in ((0, up), (1, dn)):
        cs = p * 3 + other                                  # column sums of the two contributing rows
        if w == 1:
            out[v::2, 0] = (cs[:, 0] * 4 + 8) >> 4
            out[v::2, 1] = (cs[:, 0] * 4 + 7) >> 4
            continue
        last = np.concatenate([cs[:, :1], cs[:, :-1]], axis=1)
        nxt = np.concatenate([cs[:, 1:], cs[:, -1:]], axis=1)
        even = (cs * 3 + last + 8) >> 4
        odd = (cs * 3 + nxt + 7) >> 4
        even[:, 0] = (cs[:, 0] * 4 + 8) >> 4
        odd[:, -1] = (cs[:, -1] * 4 + 7) >> 4
        out[v::2, 0::2] = even
        out[v::2, 1::2] = odd
    return out.astype(np.uint8)


# ---- jdcolor.c: ycc_rgb_convert -------------------------------------------------------------------------------------
def _ycc_tables():
    x = np.arange(256, dtype=np.int64) - 128
    fix = lambda v: int(v * 65536 + 0.5)
    cr_r = (fix(1.40200) * x + 32768) >> 16
    cb_b = (fix(1.77200) * x + 32768) >> 16
    cr_g = -fix(0.71414) * x
    cb_g = -fix(0.34414) * x + 32768
    return cr_r, cb_b, cr_g, cb_g


def ycc_to_rgb(y, cb, cr):
    cr_r, cb_b, cr_g, cb_g = _ycc_tables()
    y = y.astype(np.int64)
    r = np.clip(y + cr_r[cr], 0, 255)
    g = np.clip(y + ((cb_g[cb] + cr_g[cr]) >> 16), 0, 255)
    b = np.clip(y + cb_b[cb], 0, 255)
    return np.stack([r, g, b], axis=-1).astype(np.uint8)


def decode(data: bytes) -> np.ndarray:
    """JPEG bytes -> uint8 [H, W, 3] (RGB; grey files are replicated, as ``Image.convert('RGB')`` does)."""
    p = parse(data)
    coef, (hmax, vmax) = decode_coefficients(p)
    fr = p["frame"]
    H, W = fr["h"], fr["w"]
    planes = []
    for c, cf in zip(fr["comps"], coef):
        blocks = idct_islow(cf, p["qt"][c["tq"]])                                 # [by, bx, 8, 8]
        by, bx = blocks.shape[:2]
        plane = blocks.transpose(0, 2, 1, 3).reshape(by * 8, bx * 8)
        # the real (unpadded) extent of this component: ceil(image * samp / max_samp)
        ch, cw = -(-H * c["v"] // vmax), -(-W * c["h"] // hmax)
        plane = plane[:ch, :cw]
        fh, fv = hmax // c["h"], vmax // c["v"]
        if (fh, fv) == (1, 1):
            full = plane
        elif cw <= 2 and (fh, fv) in ((2, 1), (2, 2)):      # jinit_upsampler: fancy only when downsampled_width > 2
            full = np.repeat(np.repeat(plane, fv, axis=0), fh, axis=1)
        elif (fh, fv) == (2, 1):
            full = _h2v1_fancy_rows(plane)
        elif (fh, fv) == (2, 2):
            full = _h2v2_fancy(plane)
        else:
            raise ValueError(f"chroma subsampling {fh}x{fv} is outside the ingest path's scope")
        planes.append(full[:H, :W])
    if len(planes) == 1:
        return np.repeat(planes[0][..., None], 3, axis=-1)
    return ycc_to_rgb(*planes)
