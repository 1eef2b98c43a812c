"""TEST INFRASTRUCTURE — CPU restatement of the JPEG ENCODE at the far end of the reference's frame loop
(`cv2.imwrite(output_path, bgr_img)` in reference combine_detect.py:176-180, i.e. OpenCV's bundled libjpeg-turbo with
OpenCV's defaults: quality 95, 4:2:0, baseline, the standard Huffman tables, no restart markers, JFIF header).

SURVEY §8f row N1 ("frame I/O either side of the path"): for the device encoder (csrc/jpeg_enc.cu) to be a drop-in, the
FILE it produces must be byte-identical to what cv2.imwrite produces from the same pixels.  This module is the checker:
the published integer arithmetic of libjpeg restated with numpy —
  jccolor.c  rgb_ycc_convert (16-bit fixed point, the 0.5-epsilon rounding fudge of the chroma rows),
  jcsample.c h2v2_downsample (bias 1,2,1,2,...), with the edge replication of jcprepct.c / expand_right_edge,
  jfdctint.c jpeg_fdct_islow (CONST_BITS 13, PASS1_BITS 2) on level-shifted samples,
  jcdctmgr.c quantisation (round half away from zero against 8 * q), jcparam.c jpeg_set_quality / the Annex K tables,
  jccoefct.c dummy blocks at the right / bottom edge (zero AC, DC of the previous block),
  jchuff.c   encode_one_block + byte stuffing + 1-padding, jcmarker.c header layout.
Pinned (tests/test_jpeg_enc_oracle.py): byte-for-byte against `cv2.imencode('.jpg', ...)` of this image's OpenCV build for
many sizes (odd, tiny, MCU-edge cases, 1080p), qualities and the three sampling modes.  Only tests may import this module.
"""
import numpy as np

from .jpeg_oracle import ZIGZAG

# ITU T.81 Annex K.1 / K.2 (natural order) and K.3 - K.6 (BITS, HUFFVAL)
STD_LUMA_Q = np.array([16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56,
                       14, 17, 22, 29, 51, 87, 80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92,
                       49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99], np.int32)
STD_CHROMA_Q = np.array([17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99,
                         47, 66, 99, 99, 99, 99, 99, 99] + [99] * 32, np.int32)
DC_LUMA_BITS = [0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0]
DC_CHROMA_BITS = [0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0]
DC_VALS = list(range(12))
AC_LUMA_BITS = [0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7D]
AC_LUMA_VALS = [
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32,
    0x81, 0x91, 0xA1, 0x08, 0x23, 0x42, 0xB1, 0xC1, 0x15, 0x52, 0xD1, 0xF0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0A, 0x16,
    0x17, 0x18, 0x19, 0x1A, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2A, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3A, 0x43, 0x44, 0x45,
    0x46, 0x47, 0x48, 0x49, 0x4A, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5A, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69,
    0x6A, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7A, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8A, 0x92, 0x93, 0x94,
    0x95, 0x96, 0x97, 0x98, 0x99, 0x9A, 0xA2, 0xA3, 0xA4, 0xA5, 0xA6, 0xA7, 0xA8, 0xA9, 0xAA, 0xB2, 0xB3, 0xB4, 0xB5, 0xB6,
    0xB7, 0xB8, 0xB9, 0xBA, 0xC2, 0xC3, 0xC4, 0xC5, 0xC6, 0xC7, 0xC8, 0xC9, 0xCA, 0xD2, 0xD3, 0xD4, 0xD5, 0xD6, 0xD7, 0xD8,
    0xD9, 0xDA, 0xE1, 0xE2, 0xE3, 0xE4, 0xE5, 0xE6, 0xE7, 0xE8, 0xE9, 0xEA, 0xF1, 0xF2, 0xF3, 0xF4, 0xF5, 0xF6, 0xF7, 0xF8,
    0xF9, 0xFA]
AC_CHROMA_BITS = [0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77]
AC_CHROMA_VALS = [
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81,
    0x08, 0x14, 0x42, 0x91, 0xA1, 0xB1, 0xC1, 0x09, 0x23, 0x33, 0x52, 0xF0, 0x15, 0x62, 0x72, 0xD1, 0x0A, 0x16, 0x24, 0x34,
    0xE1, 0x25, 0xF1, 0x17, 0x18, 0x19, 0x1A, 0x26, 0x27, 0x28, 0x29, 0x2A, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3A, 0x43, 0x44,
    0x45, 0x46, 0x47, 0x48, 0x49, 0x4A, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5A, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68,
    0x69, 0x6A, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7A, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8A, 0x92,
    0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9A, 0xA2, 0xA3, 0xA4, 0xA5, 0xA6, 0xA7, 0xA8, 0xA9, 0xAA, 0xB2, 0xB3, 0xB4,
    0xB5, 0xB6, 0xB7, 0xB8, 0xB9, 0xBA, 0xC2, 0xC3, 0xC4, 0xC5, 0xC6, 0xC7, 0xC8, 0xC9, 0xCA, 0xD2, 0xD3, 0xD4, 0xD5, 0xD6,
    0xD7, 0xD8, 0xD9, 0xDA, 0xE2, 0xE3, 0xE4, 0xE5, 0xE6, 0xE7, 0xE8, 0xE9, 0xEA, 0xF2, 0xF3, 0xF4, 0xF5, 0xF6, 0xF7, 0xF8,
    0xF9, 0xFA]


def quant_table(std, quality):
    """jcparam.c jpeg_quality_scaling + jpeg_add_quant_table (force_baseline): natural order, int32"""
    quality = min(max(int(quality), 1), 100)
    scale = 5000 // quality if quality < 50 else 200 - 2 * quality
    return np.clip((std.astype(np.int64) * scale + 50) // 100, 1, 255).astype(np.int32)


def code_table(bits, vals):
    """jchuff.c jpeg_make_c_derived_tbl: symbol -> (code, length)"""
    ehufco, ehufsi = {}, {}
    code, k = 0, 0
    for length in range(1, 17):
        for _ in range(bits[length - 1]):
            ehufco[vals[k]], ehufsi[vals[k]] = code, length
            code += 1
            k += 1
        code <<= 1
    return ehufco, ehufsi


# ------------------------------------------------------------------------------------------------------------------
# jccolor.c rgb_ycc_convert
# ------------------------------------------------------------------------------------------------------------------
def _fix(x):
    return int(x * 65536 + 0.5)


def rgb_to_ycc(rgb):
    r, g, b = (rgb[..., i].astype(np.int64) for i in range(3))
    half, off = 1 << 15, 128 << 16
    y = (_fix(0.29900) * r + _fix(0.58700) * g + _fix(0.11400) * b + half) >> 16
    cb = (-_fix(0.16874) * r - _fix(0.33126) * g + _fix(0.50000) * b + off + half - 1) >> 16
    cr = (_fix(0.50000) * r - _fix(0.41869) * g - _fix(0.08131) * b + off + half - 1) >> 16
    return y.astype(np.uint8), cb.astype(np.uint8), cr.astype(np.uint8)


# ------------------------------------------------------------------------------------------------------------------
# jcprepct.c / jcsample.c: edge replication + down-sampling to whole blocks of each component
# ------------------------------------------------------------------------------------------------------------------
def _pad(plane, rows, cols):
    """replicate the last row / column out to (rows, cols)"""
    p = plane
    if p.shape[0] < rows:
        p = np.concatenate([p, np.repeat(p[-1:], rows - p.shape[0], 0)], 0)
    if p.shape[1] < cols:
        p = np.concatenate([p, np.repeat(p[:, -1:], cols - p.shape[1], 1)], 1)
    return p


def downsample(plane, hs, vs, out_rows, out_cols):
    """hs, vs in {1, 2}: pixels per output sample; output (out_rows, out_cols) covers whole blocks of the component.
    Columns: the FULL-RESOLUTION row is extended to out_cols * hs by its last pixel (jcsample.c expand_right_edge) before
    it is averaged.  Rows: the full-resolution image is only extended to a whole row group (vs rows, jcprepct.c
    expand_bottom_edge on the colour buffer); the rows still missing are copies of the last DOWN-SAMPLED row
    (jcprepct.c pre_process_data, 'pad the output to a full iMCU height')."""
    real_rows = -(-plane.shape[0] // vs)
    p = _pad(plane, real_rows * vs, out_cols * hs).astype(np.int32)
    if hs == 1 and vs == 1:
        d = p
    elif hs == 2 and vs == 1:                     # h2v1_downsample: bias 0,1,0,1,...
        bias = np.tile(np.array([0, 1], np.int32), out_cols // 2 + 1)[:out_cols]
        d = (p[:, 0::2] + p[:, 1::2] + bias) >> 1
    elif hs == 2 and vs == 2:                     # h2v2_downsample: bias 1,2,1,2,...
        bias = np.tile(np.array([1, 2], np.int32), out_cols // 2 + 1)[:out_cols]
        d = (p[0::2, 0::2] + p[0::2, 1::2] + p[1::2, 0::2] + p[1::2, 1::2] + bias) >> 2
    else:
        raise NotImplementedError((hs, vs))
    return _pad(d, out_rows, out_cols).astype(np.uint8)


# ------------------------------------------------------------------------------------------------------------------
# jfdctint.c jpeg_fdct_islow, vectorised over blocks; input (..., 8, 8) level-shifted samples, output scaled by 8
# ------------------------------------------------------------------------------------------------------------------
F_0_298, F_0_390, F_0_541, F_0_765, F_0_899, F_1_175 = 2446, 3196, 4433, 6270, 7373, 9633
F_1_501, F_1_847, F_1_961, F_2_053, F_2_562, F_3_072 = 12299, 15137, 16069, 16819, 20995, 25172


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def _fdct_1d(d, first):
    """one pass along the last axis (8 values); `first`: row pass (results scaled up by 4), else the column pass"""
    tmp0, tmp7 = d[..., 0] + d[..., 7], d[..., 0] - d[..., 7]
    tmp1, tmp6 = d[..., 1] + d[..., 6], d[..., 1] - d[..., 6]
    tmp2, tmp5 = d[..., 2] + d[..., 5], d[..., 2] - d[..., 5]
    tmp3, tmp4 = d[..., 3] + d[..., 4], d[..., 3] - d[..., 4]
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    sh = 13 - 2 if first else 13 + 2
    if first:
        o0, o4 = (tmp10 + tmp11) << 2, (tmp10 - tmp11) << 2
    else:
        o0, o4 = _descale(tmp10 + tmp11, 2), _descale(tmp10 - tmp11, 2)
    z1 = (tmp12 + tmp13) * F_0_541
    o2 = _descale(z1 + tmp13 * F_0_765, sh)
    o6 = _descale(z1 + tmp12 * (-F_1_847), sh)
    z1, z2, z3, z4 = tmp4 + tmp7, tmp5 + tmp6, tmp4 + tmp6, tmp5 + tmp7
    z5 = (z3 + z4) * F_1_175
    t4, t5, t6, t7 = tmp4 * F_0_298, tmp5 * F_2_053, tmp6 * F_3_072, tmp7 * F_1_501
    z1, z2, z3, z4 = z1 * (-F_0_899), z2 * (-F_2_562), z3 * (-F_1_961) + z5, z4 * (-F_0_390) + z5
    o7, o5, o3, o1 = _descale(t4 + z1 + z3, sh), _descale(t5 + z2 + z4, sh), _descale(t6 + z2 + z3, sh), _descale(t7 + z1 + z4, sh)
    return np.stack([o0, o1, o2, o3, o4, o5, o6, o7], -1)


def fdct_islow(blocks_u8):
    d = blocks_u8.astype(np.int64) - 128
    rows = _fdct_1d(d, True)                                   # pass 1: along each row
    return _fdct_1d(rows.swapaxes(-1, -2), False).swapaxes(-1, -2)   # pass 2: along each column


def quantize(coef, q):
    """jcdctmgr.c: divisor = 8 * q, round half away from zero.  coef (..., 8, 8), q (64,) natural order"""
    div = (q.astype(np.int64) << 3).reshape(8, 8)
    a = np.abs(coef)
    v = (a + (div >> 1)) // div
    return np.where(coef < 0, -v, v).astype(np.int32)


def component_coefficients(plane_u8, blocks_y, blocks_x, q):
    """(blocks_y*8, blocks_x*8) samples -> (blocks_y, blocks_x, 64) quantised coefficients, natural order"""
    b = plane_u8.reshape(blocks_y, 8, blocks_x, 8).transpose(0, 2, 1, 3)
    return quantize(fdct_islow(b), q).reshape(blocks_y, blocks_x, 64)


# ------------------------------------------------------------------------------------------------------------------
# container + entropy coding
# ------------------------------------------------------------------------------------------------------------------
class _BitWriter:
    def __init__(self):
        self.out, self.acc, self.n = bytearray(), 0, 0

    def put(self, code, size):
        if size == 0:
            return
        self.acc = (self.acc << size) | (code & ((1 << size) - 1))
        self.n += size
        while self.n >= 8:
            b = (self.acc >> (self.n - 8)) & 0xFF
            self.out.append(b)
            if b == 0xFF:
                self.out.append(0)
            self.n -= 8
        self.acc &= (1 << self.n) - 1

    def flush(self):
        if self.n:
            self.put((1 << (8 - self.n)) - 1, 8 - self.n)      # pad with one-bits


def _nbits(v):
    return int(v).bit_length()


def _encode_block(bw, blk_zz, last_dc, dc_tab, ac_tab):
    """jchuff.c encode_one_block; blk_zz: 64 coefficients in zig-zag order"""
    diff = int(blk_zz[0]) - last_dc
    t, t2 = (diff, diff) if diff >= 0 else (-diff, diff - 1)
    n = _nbits(t)
    bw.put(dc_tab[0][n], dc_tab[1][n])
    bw.put(t2, n)
    r = 0
    nz = np.nonzero(blk_zz[1:])[0]
    last = int(nz[-1]) + 1 if len(nz) else 0
    for k in range(1, last + 1):
        v = int(blk_zz[k])
        if v == 0:
            r += 1
            continue
        while r > 15:
            bw.put(ac_tab[0][0xF0], ac_tab[1][0xF0])
            r -= 16
        t, t2 = (v, v) if v >= 0 else (-v, v - 1)
        n = _nbits(t)
        bw.put(ac_tab[0][(r << 4) + n], ac_tab[1][(r << 4) + n])
        bw.put(t2, n)
        r = 0
    if last < 63:
        bw.put(ac_tab[0][0], ac_tab[1][0])


def _seg(marker, payload):
    return bytes([0xFF, marker]) + (len(payload) + 2).to_bytes(2, "big") + bytes(payload)


def header(width, height, sampling, qt):
    """jcmarker.c write_file_header + write_frame_header + write_scan_header, as libjpeg emits them for a 3-component
    YCbCr image: SOI, APP0 (JFIF 1.01, density 1:1, no units), DQT x2, SOF0, DHT x4 (DC0, AC0, DC1, AC1), SOS"""
    hs, vs = sampling
    out = bytearray(b"\xff\xd8")
    out += _seg(0xE0, b"JFIF\x00\x01\x01\x00\x00\x01\x00\x01\x00\x00")
    for i in range(2):
        out += _seg(0xDB, bytes([i]) + bytes(int(v) for v in qt[i][ZIGZAG]))
    sof = bytes([8]) + height.to_bytes(2, "big") + width.to_bytes(2, "big") + bytes([3])
    sof += bytes([1, (hs << 4) | vs, 0, 2, 0x11, 1, 3, 0x11, 1])
    out += _seg(0xC0, sof)
    for tc_th, bits, vals in ((0x00, DC_LUMA_BITS, DC_VALS), (0x10, AC_LUMA_BITS, AC_LUMA_VALS),
                              (0x01, DC_CHROMA_BITS, DC_VALS), (0x11, AC_CHROMA_BITS, AC_CHROMA_VALS)):
        out += _seg(0xC4, bytes([tc_th]) + bytes(bits) + bytes(vals))
    out += _seg(0xDA, bytes([3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0]))
    return bytes(out)


def image_coefficients(rgb, quality=95, sampling=(2, 2)):
    """-> (qt pair, [Y, Cb, Cr] coefficient arrays (mcu_rows*v, mcu_cols*h, 64) INCLUDING the dummy blocks of jccoefct.c)"""
    H, W = rgb.shape[:2]
    hs, vs = sampling
    qt = (quant_table(STD_LUMA_Q, quality), quant_table(STD_CHROMA_Q, quality))
    y, cb, cr = rgb_to_ycc(rgb)
    mcux, mcuy = -(-W // (8 * hs)), -(-H // (8 * vs))
    out = []
    for plane, (ch, cv), q in ((y, (hs, vs), qt[0]), (cb, (1, 1), qt[1]), (cr, (1, 1), qt[1])):
        # component size in samples (jcmaster.c: ceil(image * samp / max_samp)), real blocks, MCU-padded blocks
        cw, chh = -(-W * ch // hs), -(-H * cv // vs)
        bx, by = -(-cw // 8), -(-chh // 8)
        ds = downsample(plane, hs // ch, vs // cv, by * 8, bx * 8)
        real = component_coefficients(ds, by, bx, q)
        full = np.zeros((mcuy * cv, mcux * ch, 64), np.int32)
        full[:by, :bx] = real
        # jccoefct.c: dummy blocks right of the last real block of a row: zero AC, DC of the previous block
        for c in range(bx, mcux * ch):
            full[:by, c, 0] = full[:by, c - 1, 0]
        # dummy block rows below: DC of the previous block IN MCU ORDER (MCU_buffer[blkn - 1]): for the first block of
        # such a row that is the LAST block of the row above within the same MCU
        for r in range(by, mcuy * cv):
            for m in range(mcux):
                prev = full[r - 1, m * ch + ch - 1, 0]
                for c in range(ch):
                    full[r, m * ch + c, 0] = prev
                    prev = full[r, m * ch + c, 0]
        out.append(full)
    return qt, out


def encode_rgb(rgb, quality=95, sampling=(2, 2)):
    """(H, W, 3) uint8 RGB -> bytes of the JPEG file cv2.imwrite / cv2.imencode('.jpg') produce from the same pixels
    (cv2 takes BGR; the pixel values are what matters)"""
    H, W = rgb.shape[:2]
    hs, vs = sampling
    qt, coefs = image_coefficients(rgb, quality, sampling)
    dc = (code_table(DC_LUMA_BITS, DC_VALS), code_table(DC_CHROMA_BITS, DC_VALS))
    ac = (code_table(AC_LUMA_BITS, AC_LUMA_VALS), code_table(AC_CHROMA_BITS, AC_CHROMA_VALS))
    bw = _BitWriter()
    mcux, mcuy = -(-W // (8 * hs)), -(-H // (8 * vs))
    last = [0, 0, 0]
    zz = [c[..., ZIGZAG] for c in coefs]
    for my in range(mcuy):
        for mx in range(mcux):
            for ci, (ch, cv) in enumerate(((hs, vs), (1, 1), (1, 1))):
                t = 0 if ci == 0 else 1
                for by in range(cv):
                    for bx in range(ch):
                        blk = zz[ci][my * cv + by, mx * ch + bx]
                        _encode_block(bw, blk, last[ci], dc[t], ac[t])
                        last[ci] = int(blk[0])
    bw.flush()
    return header(W, H, sampling, qt) + bytes(bw.out) + b"\xff\xd9"
