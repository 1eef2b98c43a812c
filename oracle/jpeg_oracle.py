"""TEST INFRASTRUCTURE — CPU restatement of the JPEG decode that feeds the reference's frame loop
(`cv2.imread` in reference combine_detect.py:167-180, i.e. OpenCV's bundled libjpeg-turbo 3.x with its defaults:
baseline Huffman, `JDCT_ISLOW`, fancy up-sampling, YCbCr -> RGB by the integer tables of jdcolor.c).

SURVEY §8f row N1 ("frames arrive device-resident") needs a decoder whose pixels are IDENTICAL to cv2.imread's, otherwise
every downstream parity statement (detections, blurred bytes) would start from different frames.  This module is the
checker for that future device path: entropy decoding is plain Python (small images only), everything after it —
de-quantisation, the 8x8 inverse DCT, chroma up-sampling, colour conversion — is the published integer arithmetic of
libjpeg (jidctint.c `jpeg_idct_islow`, jdsample.c `h2v1/h2v2_fancy_upsample`, jdcolor.c `ycc_rgb_convert`), restated with
numpy.  libjpeg-turbo's SIMD paths are bit-exact with that C code for these methods, which is what makes the pin possible.

Pinned (tests/test_jpeg_oracle.py): bit-exact against `cv2.imdecode` for 4:2:0 / 4:2:2 / 4:4:4 / greyscale files, odd
sizes, several qualities and restart intervals.  Only tests may import this module.
"""
import numpy as np

ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14,
                   21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46,
                   53, 60, 61, 54, 47, 55, 62, 63], np.int32)      # natural index of the k-th zig-zag coefficient


# ------------------------------------------------------------------------------------------------------------------
# container: markers and tables
# ------------------------------------------------------------------------------------------------------------------
class Component:
    def __init__(self, cid, h, v, tq):
        self.cid, self.h, self.v, self.tq = cid, h, v, tq
        self.td = self.ta = 0


class Huff:
    """canonical Huffman table (ITU T.81 Annex C): code length -> (min code, max code, first symbol index)"""

    def __init__(self, counts, symbols):
        self.symbols = symbols
        self.mincode, self.maxcode, self.valptr = [0] * 17, [-1] * 17, [0] * 17
        code = k = 0
        for length in range(1, 17):
            self.valptr[length] = k
            self.mincode[length] = code
            code += counts[length - 1]
            k += counts[length - 1]
            self.maxcode[length] = code - 1 if counts[length - 1] else -1
            code <<= 1
        # 9-bit look-ahead table: (symbol, length) for every prefix that resolves within 9 bits
        self.fast = [None] * 512
        code = k = 0
        for length in range(1, 10):
            for _ in range(counts[length - 1]):
                base = code << (9 - length)
                for fill in range(1 << (9 - length)):
                    self.fast[base + fill] = (symbols[k], length)
                code += 1
                k += 1
            code <<= 1


def parse(data: bytes):
    """-> dict(width, height, comps, qt {id: (64,) int32 natural order}, dc/ac {id: Huff}, restart, scan bytes)"""
    assert data[:2] == b"\xff\xd8", "not a JPEG"
    pos = 2
    out = {"qt": {}, "dc": {}, "ac": {}, "restart": 0}
    while True:
        assert data[pos] == 0xFF
        while data[pos + 1] == 0xFF:
            pos += 1
        marker = data[pos + 1]
        pos += 2
        if marker == 0xD9:
            raise ValueError("EOI before SOS")
        length = (data[pos] << 8) | data[pos + 1]
        seg = data[pos + 2:pos + length]
        pos += length
        if marker == 0xDB:                                   # DQT
            i = 0
            while i < len(seg):
                pq, tq = seg[i] >> 4, seg[i] & 15
                i += 1
                if pq:
                    vals = [(seg[i + 2 * k] << 8) | seg[i + 2 * k + 1] for k in range(64)]
                    i += 128
                else:
                    vals = list(seg[i:i + 64])
                    i += 64
                q = np.zeros(64, np.int32)
                q[ZIGZAG] = vals
                out["qt"][tq] = q
        elif marker in (0xC0, 0xC1):                         # SOF0 / SOF1: sequential Huffman, 8-bit
            assert seg[0] == 8, "only 8-bit samples"
            out["height"], out["width"] = (seg[1] << 8) | seg[2], (seg[3] << 8) | seg[4]
            out["comps"] = [Component(seg[6 + 3 * c], seg[7 + 3 * c] >> 4, seg[7 + 3 * c] & 15, seg[8 + 3 * c])
                            for c in range(seg[5])]
        elif marker == 0xC2:
            raise NotImplementedError("progressive JPEG (cv2.imwrite never writes it by default)")
        elif marker == 0xC4:                                 # DHT
            i = 0
            while i < len(seg):
                tc, th = seg[i] >> 4, seg[i] & 15
                counts = list(seg[i + 1:i + 17])
                n = sum(counts)
                out["ac" if tc else "dc"][th] = Huff(counts, list(seg[i + 17:i + 17 + n]))
                i += 17 + n
        elif marker == 0xDD:                                 # DRI
            out["restart"] = (seg[0] << 8) | seg[1]
        elif marker == 0xDA:                                 # SOS: one interleaved scan of all components
            ns = seg[0]
            assert ns == len(out["comps"]), "non-interleaved scans are not written by cv2"
            for s in range(ns):
                comp = next(c for c in out["comps"] if c.cid == seg[1 + 2 * s])
                comp.td, comp.ta = seg[2 + 2 * s] >> 4, seg[2 + 2 * s] & 15
            out["scan"] = data[pos:]
            return out


# ------------------------------------------------------------------------------------------------------------------
# entropy decoding (T.81 F.2.2) -> quantised coefficients in natural order
# ------------------------------------------------------------------------------------------------------------------
class Bits:
    def __init__(self, data):
        self.d, self.p, self.acc, self.n = data, 0, 0, 0

    def fill(self, need):
        while self.n < need:
            if self.p >= len(self.d):
                b = 0
            else:
                b = self.d[self.p]
                if b == 0xFF:
                    nxt = self.d[self.p + 1] if self.p + 1 < len(self.d) else 0xD9
                    if nxt == 0:
                        self.p += 2                      # stuffed zero
                    else:
                        b = 0                            # a marker: feed zeros, do not advance (libjpeg does the same)
                else:
                    self.p += 1
            self.acc = ((self.acc << 8) | b) & 0xFFFFFFFFFF
            self.n += 8

    def peek(self, k):
        self.fill(k)
        return (self.acc >> (self.n - k)) & ((1 << k) - 1)

    def skip(self, k):
        self.n -= k

    def get(self, k):
        if k == 0:
            return 0
        v = self.peek(k)
        self.n -= k
        return v

    def restart(self):
        """byte-align, expect RSTn"""
        self.acc = self.n = 0
        while not (self.d[self.p] == 0xFF and 0xD0 <= self.d[self.p + 1] <= 0xD7):
            self.p += 1
        self.p += 2


def _decode_symbol(bits, h):
    e = h.fast[bits.peek(9)]
    if e is not None:
        bits.skip(e[1])
        return e[0]
    code = bits.get(9)
    length = 9
    while length <= 16 and (h.maxcode[length] < 0 or code > h.maxcode[length]):
        code = (code << 1) | bits.get(1)
        length += 1
    assert length <= 16, "corrupt Huffman code"
    return h.symbols[h.valptr[length] + code - h.mincode[length]]


def _extend(v, t):
    return v if v >= (1 << (t - 1)) else v - (1 << t) + 1


def decode_coefficients(j):
    """-> list (per component) of int32 arrays (blocks_y, blocks_x, 64), natural order, still quantised; the block grid
    covers whole MCUs"""
    comps = j["comps"]
    hmax, vmax = max(c.h for c in comps), max(c.v for c in comps)
    mcux, mcuy = -(-j["width"] // (8 * hmax)), -(-j["height"] // (8 * vmax))
    coefs = [np.zeros((mcuy * c.v, mcux * c.h, 64), np.int32) for c in comps]
    bits = Bits(j["scan"])
    pred = [0] * len(comps)
    count = 0
    for my in range(mcuy):
        for mx in range(mcux):
            if j["restart"] and count and count % j["restart"] == 0:
                bits.restart()
                pred = [0] * len(comps)
            count += 1
            for ci, c in enumerate(comps):
                dc, ac = j["dc"][c.td], j["ac"][c.ta]
                for by in range(c.v):
                    for bx in range(c.h):
                        blk = coefs[ci][my * c.v + by, mx * c.h + bx]
                        t = _decode_symbol(bits, dc)
                        pred[ci] += _extend(bits.get(t), t) if t else 0
                        blk[0] = pred[ci]
                        k = 1
                        while k < 64:
                            rs = _decode_symbol(bits, ac)
                            r, s = rs >> 4, rs & 15
                            if s == 0:
                                if r != 15:
                                    break
                                k += 16
                                continue
                            k += r
                            blk[ZIGZAG[k]] = _extend(bits.get(s), s)
                            k += 1
    return coefs


# ------------------------------------------------------------------------------------------------------------------
# jidctint.c: jpeg_idct_islow (CONST_BITS 13, PASS1_BITS 2), vectorised over blocks
# ------------------------------------------------------------------------------------------------------------------
F_0_298, F_0_390, F_0_541, F_0_765, F_0_899, F_1_175 = 2446, 3196, 4433, 6270, 7373, 9633
F_1_501, F_1_847, F_1_961, F_2_053, F_2_562, F_3_072 = 12299, 15137, 16069, 16819, 20995, 25172


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def _idct_1d(v, shift):
    """v: (..., 8) int64 along the last axis; one pass of the LL&M butterfly with the final descale by `shift`"""
    z2, z3 = v[..., 2], v[..., 6]
    z1 = (z2 + z3) * F_0_541
    tmp2 = z1 - z3 * F_1_847
    tmp3 = z1 + z2 * F_0_765
    tmp0 = (v[..., 0] + v[..., 4]) << 13
    tmp1 = (v[..., 0] - v[..., 4]) << 13
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    t0, t1, t2, t3 = v[..., 7], v[..., 5], v[..., 3], v[..., 1]
    z1, z2, z3, z4 = t0 + t3, t1 + t2, t0 + t2, t1 + t3
    z5 = (z3 + z4) * F_1_175
    t0, t1, t2, t3 = t0 * F_0_298, t1 * F_2_053, t2 * F_3_072, t3 * F_1_501
    z1, z2, z3, z4 = -z1 * F_0_899, -z2 * F_2_562, -z3 * F_1_961 + z5, -z4 * F_0_390 + z5
    t0, t1, t2, t3 = t0 + z1 + z3, t1 + z2 + z4, t2 + z2 + z3, t3 + z1 + z4
    out = np.stack([tmp10 + t3, tmp11 + t2, tmp12 + t1, tmp13 + t0, tmp13 - t0, tmp12 - t1, tmp11 - t2, tmp10 - t3], -1)
    return _descale(out, shift)


def idct_islow(coef, quant):
    """coef (..., 64) quantised, quant (64,) -> (..., 8, 8) uint8 samples.  The zero-AC shortcuts of the C code give the same
    values as the general path (DESCALE of an exact multiple), so they are not restated."""
    x = (coef.astype(np.int64) * quant.astype(np.int64)).reshape(coef.shape[:-1] + (8, 8))
    ws = _idct_1d(x.swapaxes(-1, -2), 13 - 2).swapaxes(-1, -2)       # pass 1 works on columns
    y = _idct_1d(ws, 13 + 2 + 3)                                     # pass 2 on rows
    idx = y & 0x3FF                                                  # range_limit[(x) & RANGE_MASK], table centred at 128
    out = np.where(idx < 128, idx + 128, np.where(idx < 512, 255, np.where(idx < 896, 0, idx - 896)))
    return out.astype(np.uint8)


def blocks_to_plane(samples):
    by, bx = samples.shape[:2]
    return samples.transpose(0, 2, 1, 3).reshape(by * 8, bx * 8)


# ------------------------------------------------------------------------------------------------------------------
# jdsample.c: fancy (triangle) up-sampling; edges use the component's REAL extent, not the MCU padding
# ------------------------------------------------------------------------------------------------------------------
def h2v1_fancy(plane, dw):
    """(rows, >= dw) -> (rows, 2 * dw); jdsample.c h2v1_fancy_upsample"""
    p = plane[:, :dw].astype(np.int32)
    out = np.empty((p.shape[0], 2 * dw), np.int32)
    left = np.concatenate([p[:, :1], p[:, :-1]], 1)
    right = np.concatenate([p[:, 1:], p[:, -1:]], 1)
    out[:, 0::2] = (3 * p + left + 1) >> 2
    out[:, 1::2] = (3 * p + right + 2) >> 2
    out[:, 0] = p[:, 0]
    out[:, -1] = p[:, -1]
    return out.astype(np.uint8)


def h2v2_fancy(plane, dw, dh):
    """(>= dh, >= dw) -> (2 * dh, 2 * dw); jdsample.c h2v2_fancy_upsample, rows above / below clamped to [0, dh - 1]
    (jdmainct.c duplicates the first / last real sample row as context)"""
    p = plane[:dh, :dw].astype(np.int32)
    up = np.concatenate([p[:1], p[:-1]], 0)
    down = np.concatenate([p[1:], p[-1:]], 0)
    out = np.empty((2 * dh, 2 * dw), np.int32)
    for v, other in ((0, up), (1, down)):
        colsum = 3 * p + other                                     # thiscolsum of every column
        last = np.concatenate([colsum[:, :1], colsum[:, :-1]], 1)
        nxt = np.concatenate([colsum[:, 1:], colsum[:, -1:]], 1)
        even = (3 * colsum + last + 8) >> 4
        odd = (3 * colsum + nxt + 7) >> 4
        even[:, 0] = (4 * colsum[:, 0] + 8) >> 4
        odd[:, -1] = (4 * colsum[:, -1] + 7) >> 4
        out[v::2, 0::2] = even
        out[v::2, 1::2] = odd
    return out.astype(np.uint8)


# ------------------------------------------------------------------------------------------------------------------
# jdcolor.c: build_ycc_rgb_table + ycc_rgb_convert (SCALEBITS 16)
# ------------------------------------------------------------------------------------------------------------------
def _fix(x):
    return int(x * 65536 + 0.5)


_X = np.arange(256, dtype=np.int64) - 128
CR_R = ((_fix(1.40200) * _X + 32768) >> 16).astype(np.int32)
CB_B = ((_fix(1.77200) * _X + 32768) >> 16).astype(np.int32)
CR_G = (-_fix(0.71414) * _X).astype(np.int64)
CB_G = (-_fix(0.34414) * _X + 32768).astype(np.int64)


def ycc_to_rgb(y, cb, cr):
    yy = y.astype(np.int32)
    r = np.clip(yy + CR_R[cr], 0, 255)
    g = np.clip(yy + ((CB_G[cb] + CR_G[cr]) >> 16).astype(np.int32), 0, 255)
    b = np.clip(yy + CB_B[cb], 0, 255)
    return np.stack([r, g, b], -1).astype(np.uint8)


# ------------------------------------------------------------------------------------------------------------------
def decode_planes(data: bytes):
    """-> (header dict, list of full-MCU sample planes per component before up-sampling)"""
    j = parse(data)
    coefs = decode_coefficients(j)
    planes = [blocks_to_plane(idct_islow(cf, j["qt"][c.tq])) for cf, c in zip(coefs, j["comps"])]
    return j, planes


def decode_rgb(data: bytes):
    """bytes of a baseline JPEG -> (H, W, 3) uint8 RGB, the array `cv2.cvtColor(cv2.imread(path), COLOR_BGR2RGB)` holds
    (reference combine_detect.py:172); greyscale files give three equal channels like cv2's IMREAD_COLOR"""
    j, planes = decode_planes(data)
    H, W, comps = j["height"], j["width"], j["comps"]
    if len(comps) == 1:
        g = planes[0][:H, :W]
        return np.stack([g, g, g], -1)
    hmax, vmax = max(c.h for c in comps), max(c.v for c in comps)
    full = []
    for c, p in zip(comps, planes):
        dw, dh = -(-W * c.h // hmax), -(-H * c.v // vmax)            # downsampled_width / _height (jdmaster.c)
        if c.h == hmax and c.v == vmax:
            f = p
        elif c.h * 2 == hmax and c.v * 2 == vmax:
            # jdsample.c jinit_upsampler: the triangle filter needs more than two columns, else plain replication
            f = h2v2_fancy(p, dw, dh) if dw > 2 else np.repeat(np.repeat(p[:dh, :dw], 2, 0), 2, 1)
        elif c.h * 2 == hmax and c.v == vmax:
            f = h2v1_fancy(p, dw) if dw > 2 else np.repeat(p[:, :dw], 2, 1)
        else:
            raise NotImplementedError(f"sampling {c.h}x{c.v} of {hmax}x{vmax}")
        full.append(f[:H, :W])
    return ycc_to_rgb(*full)
