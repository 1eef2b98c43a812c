// jpeg_host.cu — host half of a device-resident JPEG ingest: marker parsing and baseline Huffman
// entropy decoding to quantised DCT coefficients.  HOST code only (it lives in a .cu so that the one Makefile rule
// builds it); the device half — de-quantisation, jpeg_idct_islow, fancy up-sampling, YCbCr -> RGB, bit-exact with the
// libjpeg-turbo that cv2.imread uses (reference combine_detect.py:167-180) — is not written yet; its checker is
// oracle/jpeg_oracle.py.  What this file produces is pinned against that oracle's coefficients (tests/test_jpeg_host.py).
//
// Entropy decoding is the serial ~half of a JPEG decode; everything after it is independent per 8x8 block / per pixel
// and belongs on the GPU.  Output layout: per component, blocks in raster order [blocks_y][blocks_x][64], coefficients in
// NATURAL (row-major) order, int16, still quantised: 3 bytes per pixel for a 4:2:0 file, the volume of the decoded RGB
// frame, in three contiguous arrays ready for H2D copies.
#include "common.cuh"

namespace {

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct HuffTable {
  bool present = false;
  uint8_t symbols[256];
  int32_t mincode[18], maxcode[18], valptr[18];
  uint16_t fast[512];            // (length << 8) | symbol for codes of at most 9 bits, 0 = longer code
  // AC tables: code AND its value bits resolved by one 9-bit look-ahead when both fit (most small coefficients):
  // (value << 8) | (run << 4) | total bits, 0 = take the general path
  int16_t fast_ac[512];
};

// false if the code lengths over-subscribe the code space (a corrupt table would index past the look-up tables)
bool build_huff(HuffTable& t, const uint8_t* counts, const uint8_t* syms, int nsym) {
  for (int len = 1, code = 0; len <= 16; ++len) {
    code += counts[len - 1];
    if (code > (1 << len)) return false;
    code <<= 1;
  }
  t.present = true;
  memcpy(t.symbols, syms, static_cast<size_t>(nsym));
  memset(t.fast, 0, sizeof(t.fast));
  int code = 0, k = 0;
  for (int len = 1; len <= 16; ++len) {
    t.valptr[len] = k;
    t.mincode[len] = code;
    for (int i = 0; i < counts[len - 1]; ++i, ++k, ++code) {
      if (len <= 9) {
        const int base = code << (9 - len);
        for (int f = 0; f < (1 << (9 - len)); ++f) t.fast[base + f] = static_cast<uint16_t>((len << 8) | syms[k]);
      }
    }
    t.maxcode[len] = counts[len - 1] ? code - 1 : -1;
    code <<= 1;
  }
  t.maxcode[17] = 0x7FFFFFFF;
  for (int i = 0; i < 512; ++i) {
    t.fast_ac[i] = 0;
    const uint16_t e = t.fast[i];
    if (!e) continue;
    const int len = e >> 8, run = (e >> 4) & 15, mag = e & 15;
    if (mag && len + mag <= 9) {
      int k = ((i << len) & 511) >> (9 - mag);                 // the value bits that follow the code
      if (k < (1 << (mag - 1))) k -= (1 << mag) - 1;           // T.81 EXTEND
      if (k >= -128 && k <= 127) t.fast_ac[i] = static_cast<int16_t>(k * 256 + run * 16 + len + mag);
    }
  }
  return true;
}

struct Parsed {
  vd_jpeg_info_t info;
  int tq[3], td[3], ta[3];
  uint16_t qt[4][64];
  bool have_qt[4];
  HuffTable dc[4], ac[4];
  const uint8_t* scan;
  size_t scan_size;
};

inline int rd16(const uint8_t* p) { return (p[0] << 8) | p[1]; }

// returns VD_OK or an error code with the message set
int parse(const uint8_t* d, size_t n, Parsed& P) {
  // P is a thread_local scratch object reused across calls: EVERY piece of per-file state is reset here, so a file
  // without a DHT segment (abbreviated MJPEG frames) can never be decoded with the previous file's Huffman tables
  memset(&P.info, 0, sizeof(P.info));
  memset(P.have_qt, 0, sizeof(P.have_qt));
  for (int i = 0; i < 4; ++i) P.dc[i].present = P.ac[i].present = false;
  for (int i = 0; i < 3; ++i) P.tq[i] = P.td[i] = P.ta[i] = 0;
  P.scan = nullptr;
  P.scan_size = 0;
  VD_REQUIRE(d && n >= 4 && d[0] == 0xFF && d[1] == 0xD8, "vd_jpeg: not a JPEG stream");
  size_t pos = 2;
  bool have_sof = false;
  int cid[3] = {0, 0, 0};
  while (true) {
    VD_REQUIRE(pos + 4 <= n && d[pos] == 0xFF, "vd_jpeg: marker expected at byte %zu", pos);
    while (pos + 1 < n && d[pos + 1] == 0xFF) ++pos;
    const int marker = d[pos + 1];
    pos += 2;
    VD_REQUIRE(marker != 0xD9, "vd_jpeg: end of image before a scan");
    VD_REQUIRE(pos + 2 <= n, "vd_jpeg: truncated segment");
    const int len = rd16(d + pos);
    VD_REQUIRE(len >= 2 && pos + static_cast<size_t>(len) <= n, "vd_jpeg: truncated segment");
    const uint8_t* seg = d + pos + 2;
    const int sl = len - 2;
    pos += static_cast<size_t>(len);
    if (marker == 0xDB) {
      int i = 0;
      while (i < sl) {
        const int pq = seg[i] >> 4, tq = seg[i] & 15;
        ++i;
        VD_REQUIRE(tq < 4 && i + (pq ? 128 : 64) <= sl, "vd_jpeg: bad quantisation table");
        for (int k = 0; k < 64; ++k) P.qt[tq][kZigzag[k]] = static_cast<uint16_t>(pq ? rd16(seg + i + 2 * k) : seg[i + k]);
        P.have_qt[tq] = true;
        i += pq ? 128 : 64;
      }
    } else if (marker == 0xC0 || marker == 0xC1) {
      VD_REQUIRE(sl >= 6 && seg[0] == 8, "vd_jpeg: only 8-bit samples");
      P.info.height = rd16(seg + 1);
      P.info.width = rd16(seg + 3);
      P.info.ncomp = seg[5];
      if (!(P.info.ncomp == 1 || P.info.ncomp == 3) || sl < 6 + 3 * P.info.ncomp) {
        vd_set_error("vd_jpeg: %d components not supported", P.info.ncomp);
        return VD_ERR_UNSUPPORTED;
      }
      VD_REQUIRE(P.info.width > 0 && P.info.height > 0, "vd_jpeg: empty image");
      for (int c = 0; c < P.info.ncomp; ++c) {
        cid[c] = seg[6 + 3 * c];
        P.info.h[c] = seg[7 + 3 * c] >> 4;
        P.info.v[c] = seg[7 + 3 * c] & 15;
        P.tq[c] = seg[8 + 3 * c];
        if (P.info.h[c] < 1 || P.info.h[c] > 2 || P.info.v[c] < 1 || P.info.v[c] > 2 || P.tq[c] > 3) {
          vd_set_error("vd_jpeg: sampling factors %dx%d not supported", P.info.h[c], P.info.v[c]);
          return VD_ERR_UNSUPPORTED;
        }
      }
      have_sof = true;
    } else if (marker == 0xC2 || (marker >= 0xC3 && marker <= 0xCF && marker != 0xC4 && marker != 0xC8 && marker != 0xCC)) {
      vd_set_error("vd_jpeg: only baseline / sequential Huffman JPEG (SOF marker 0x%02X)", marker);
      return VD_ERR_UNSUPPORTED;
    } else if (marker == 0xC4) {
      int i = 0;
      while (i < sl) {
        VD_REQUIRE(i + 17 <= sl, "vd_jpeg: bad Huffman table");
        const int tc = seg[i] >> 4, th = seg[i] & 15;
        int nsym = 0;
        for (int k = 0; k < 16; ++k) nsym += seg[i + 1 + k];
        VD_REQUIRE(tc < 2 && th < 4 && nsym <= 256 && i + 17 + nsym <= sl, "vd_jpeg: bad Huffman table");
        VD_REQUIRE(build_huff(tc ? P.ac[th] : P.dc[th], seg + i + 1, seg + i + 17, nsym),
                   "vd_jpeg: Huffman table over-subscribes its code space");
        i += 17 + nsym;
      }
    } else if (marker == 0xDD) {
      VD_REQUIRE(sl >= 2, "vd_jpeg: bad DRI");
      P.info.restart_interval = rd16(seg);
    } else if (marker == 0xDA) {
      VD_REQUIRE(have_sof, "vd_jpeg: scan before frame header");
      VD_REQUIRE(sl >= 1 + 2 * P.info.ncomp + 3, "vd_jpeg: bad scan header");
      if (seg[0] != P.info.ncomp) {
        vd_set_error("vd_jpeg: non-interleaved scans are not supported");
        return VD_ERR_UNSUPPORTED;
      }
      bool named[3] = {false, false, false};
      for (int s = 0; s < P.info.ncomp; ++s) {
        int c = -1;
        for (int q = 0; q < P.info.ncomp; ++q)
          if (cid[q] == seg[1 + 2 * s]) c = q;
        VD_REQUIRE(c >= 0, "vd_jpeg: scan names an unknown component");
        VD_REQUIRE(!named[c], "vd_jpeg: scan names a component twice");
        named[c] = true;
        P.td[c] = seg[2 + 2 * s] >> 4;
        P.ta[c] = seg[2 + 2 * s] & 15;
        VD_REQUIRE(P.td[c] < 4 && P.ta[c] < 4 && P.dc[P.td[c]].present && P.ac[P.ta[c]].present && P.have_qt[P.tq[c]],
                   "vd_jpeg: scan refers to a missing table");
      }
      P.scan = d + pos;
      P.scan_size = n - pos;
      break;
    }
  }
  int hmax = 1, vmax = 1;
  for (int c = 0; c < P.info.ncomp; ++c) {
    hmax = P.info.h[c] > hmax ? P.info.h[c] : hmax;
    vmax = P.info.v[c] > vmax ? P.info.v[c] : vmax;
  }
  if (P.info.ncomp == 1) P.info.h[0] = P.info.v[0] = hmax = vmax = 1;   // a single-component scan is never interleaved
  const int mcux = vd_cdiv(P.info.width, 8 * hmax), mcuy = vd_cdiv(P.info.height, 8 * vmax);
  for (int c = 0; c < P.info.ncomp; ++c) {
    P.info.blocks_x[c] = mcux * P.info.h[c];
    P.info.blocks_y[c] = mcuy * P.info.v[c];
    memcpy(P.info.quant[c], P.qt[P.tq[c]], sizeof(P.info.quant[c]));
  }
  return VD_OK;
}

struct BitReader {
  const uint8_t* d;
  size_t n, p = 0;
  uint64_t acc = 0;
  int bits = 0;
  // refill to more than 56 bits; markers and the end of data feed zeros (the reader stays in front of a marker)
  inline void fill() {
    if (bits <= 32 && p + 4 <= n) {         // fast path: four bytes without 0xFF
      const uint32_t w = (static_cast<uint32_t>(d[p]) << 24) | (static_cast<uint32_t>(d[p + 1]) << 16) |
                         (static_cast<uint32_t>(d[p + 2]) << 8) | d[p + 3];
      // any byte == 0xFF  <=>  any byte of ~w == 0
      const uint32_t m = ~w;
      if (!((m - 0x01010101u) & ~m & 0x80808080u)) {
        acc = (acc << 32) | w;
        bits += 32;
        p += 4;
        if (bits > 56) return;
      }
    }
    while (bits <= 56) {
      uint32_t b = 0;
      if (p < n) {
        b = d[p];
        if (b == 0xFF) {
          const uint32_t nx = (p + 1 < n) ? d[p + 1] : 0xD9;
          if (nx == 0) p += 2;
          else b = 0;                       // marker: stay in front of it
        } else {
          ++p;
        }
      }
      acc = (acc << 8) | b;
      bits += 8;
    }
  }
  inline void need32() {                    // one Huffman code (<= 16 bits) + its value bits (<= 15) fit
    if (bits < 32) fill();
  }
  inline uint32_t peek(int k) const { return static_cast<uint32_t>((acc >> (bits - k)) & ((1u << k) - 1u)); }
  inline void skip(int k) { bits -= k; }
  inline uint32_t get(int k) {
    const uint32_t v = peek(k);
    bits -= k;
    return v;
  }
  bool restart() {                          // byte-align and step over the RSTn marker
    acc = 0;
    bits = 0;
    while (p + 1 < n && !(d[p] == 0xFF && d[p + 1] >= 0xD0 && d[p + 1] <= 0xD7)) ++p;
    if (p + 1 >= n) return false;
    p += 2;
    return true;
  }
};

// caller guarantees >= 16 bits in the reader
inline int decode_symbol(BitReader& br, const HuffTable& t) {
  const uint16_t e = t.fast[br.peek(9)];
  if (e) {
    br.skip(e >> 8);
    return e & 0xFF;
  }
  int code = static_cast<int>(br.get(9)), len = 9;
  while (len < 17 && (t.maxcode[len] < 0 || code > t.maxcode[len])) {
    code = (code << 1) | static_cast<int>(br.get(1));
    ++len;
  }
  if (len > 16) return -1;
  return t.symbols[(t.valptr[len] + code - t.mincode[len]) & 255];
}

// T.81 F.2.2.1 EXTEND without a (coin-flip) branch: top bit clear -> v - (2^t - 1)
inline int extend(uint32_t v, int t) {
  const int m = static_cast<int>(v >> (t - 1)) - 1;            // 0 if the top bit is set, -1 otherwise
  return static_cast<int>(v) + (m & (1 - (1 << t)));
}

}  // namespace

extern "C" int vd_jpeg_read_header(const uint8_t* data, size_t size, vd_jpeg_info_t* info) {
  VD_REQUIRE(data && info, "vd_jpeg_read_header: null argument");
  static thread_local Parsed P;
  const int rc = parse(data, size, P);
  if (rc != VD_OK) return rc;
  *info = P.info;
  return VD_OK;
}

// the record the device decoder reads (csrc/jpeg_dec.cu): geometry, per-file quantisation and Huffman tables, scan extent
extern "C" int vd_jpeg_prepare_device(const uint8_t* data, size_t size, vd_jpeg_dev_hdr_t* hdr) {
  VD_REQUIRE(data && hdr, "vd_jpeg_prepare_device: null argument");
  static thread_local Parsed P;
  const int rc = parse(data, size, P);
  if (rc != VD_OK) return rc;
  const vd_jpeg_info_t& I = P.info;
  memset(hdr, 0, sizeof(*hdr));
  if (I.restart_interval != 0) {
    vd_set_error("vd_jpeg_prepare_device: restart markers are decoded on the CPU path");
    return VD_ERR_UNSUPPORTED;
  }
  for (int c = 0; c < I.ncomp; ++c) {
    const bool ok = c == 0 ? true : (I.h[c] == 1 && I.v[c] == 1);
    if (!ok || P.td[c] > 1 || P.ta[c] > 1 || (c == 0 && I.ncomp == 3 && I.h[0] == 1 && I.v[0] == 2)) {
      vd_set_error("vd_jpeg_prepare_device: sampling / table layout not taken by the device path");
      return VD_ERR_UNSUPPORTED;
    }
  }
  hdr->width = I.width; hdr->height = I.height; hdr->ncomp = I.ncomp;
  int nb = 0;
  for (int c = 0; c < I.ncomp; ++c) {
    hdr->h[c] = I.h[c]; hdr->v[c] = I.v[c];
    hdr->blocks_x[c] = I.blocks_x[c]; hdr->blocks_y[c] = I.blocks_y[c];
    hdr->td[c] = static_cast<uint8_t>(P.td[c]); hdr->ta[c] = static_cast<uint8_t>(P.ta[c]);
    memcpy(hdr->quant[c], I.quant[c], sizeof(hdr->quant[c]));
    for (int by = 0; by < I.v[c]; ++by)
      for (int bx = 0; bx < I.h[c]; ++bx, ++nb) {
        hdr->blk_comp[nb] = static_cast<uint8_t>(c);
        hdr->blk_bx[nb] = static_cast<uint8_t>(bx);
        hdr->blk_by[nb] = static_cast<uint8_t>(by);
      }
  }
  hdr->blocks_per_mcu = nb;
  hdr->mcux = I.blocks_x[0] / I.h[0];
  hdr->mcuy = I.blocks_y[0] / I.v[0];
  for (int t = 0; t < 2; ++t)
    for (int kind = 0; kind < 2; ++kind) {
      const HuffTable& src = kind ? P.ac[t] : P.dc[t];
      vd_jpeg_huff_dev_t& dst = kind ? hdr->ac[t] : hdr->dc[t];
      if (!src.present) continue;
      memcpy(dst.fast, src.fast, sizeof(dst.fast));
      memcpy(dst.symbols, src.symbols, sizeof(dst.symbols));
      for (int len = 0; len < 18; ++len) {
        dst.maxcode[len] = (len >= 1 && len <= 16) ? src.maxcode[len] : (len == 17 ? 0x7FFFFFFF : -1);
        dst.valoff[len] = (len >= 1 && len <= 16) ? src.valptr[len] - src.mincode[len] : 0;
      }
    }
  // extent of the entropy-coded segment: up to the first marker (0xFF followed by anything but a stuffed zero)
  const uint8_t* s = P.scan;
  const uint8_t* end = P.scan + P.scan_size;
  const uint8_t* q = s;
  while (q < end) {
    q = static_cast<const uint8_t*>(memchr(q, 0xFF, static_cast<size_t>(end - q)));
    if (!q) { q = end; break; }
    if (q + 1 >= end || q[1] != 0x00) break;
    q += 2;
  }
  hdr->scan_offset = static_cast<uint32_t>(s - data);
  hdr->scan_bytes = static_cast<uint32_t>(q - s);
  VD_REQUIRE(hdr->scan_bytes > 0, "vd_jpeg_prepare_device: empty scan");
  return VD_OK;
}

extern "C" int vd_jpeg_decode_coefficients(const uint8_t* data, size_t size, int16_t* coef0, int16_t* coef1,
                                           int16_t* coef2) {
  VD_REQUIRE(data && coef0, "vd_jpeg_decode_coefficients: null argument");
  static thread_local Parsed P;
  const int rc = parse(data, size, P);
  if (rc != VD_OK) return rc;
  const vd_jpeg_info_t& I = P.info;
  VD_REQUIRE(I.ncomp == 1 || (coef1 && coef2), "vd_jpeg_decode_coefficients: three coefficient buffers needed");
  int16_t* out[3] = {coef0, coef1, coef2};
  for (int c = 0; c < I.ncomp; ++c)
    memset(out[c], 0, sizeof(int16_t) * 64 * static_cast<size_t>(I.blocks_x[c]) * I.blocks_y[c]);
  const int mcux = I.blocks_x[0] / I.h[0], mcuy = I.blocks_y[0] / I.v[0];
  BitReader br{P.scan, P.scan_size};
  int pred[3] = {0, 0, 0};
  long count = 0;
  for (int my = 0; my < mcuy; ++my) {
    for (int mx = 0; mx < mcux; ++mx, ++count) {
      if (I.restart_interval && count && count % I.restart_interval == 0) {
        VD_REQUIRE(br.restart(), "vd_jpeg: restart marker missing");
        pred[0] = pred[1] = pred[2] = 0;
      }
      for (int c = 0; c < I.ncomp; ++c) {
        const HuffTable& dc = P.dc[P.td[c]];
        const HuffTable& ac = P.ac[P.ta[c]];
        for (int by = 0; by < I.v[c]; ++by) {
          for (int bx = 0; bx < I.h[c]; ++bx) {
            int16_t* blk = out[c] + 64 * (static_cast<size_t>(my * I.v[c] + by) * I.blocks_x[c] + mx * I.h[c] + bx);
            br.need32();
            const int t = decode_symbol(br, dc);
            VD_REQUIRE(t >= 0 && t <= 15, "vd_jpeg: corrupt DC code");
            if (t) pred[c] += extend(br.get(t), t);
            blk[0] = static_cast<int16_t>(pred[c]);
            for (int k = 1; k < 64;) {
              br.need32();
              const int f = ac.fast_ac[br.peek(9)];
              if (f) {                                   // run, size and value from one look-up
                k += (f >> 4) & 15;
                VD_REQUIRE(k < 64, "vd_jpeg: coefficient index out of range");
                br.skip(f & 15);
                blk[kZigzag[k++]] = static_cast<int16_t>(f >> 8);
                continue;
              }
              const int rs = decode_symbol(br, ac);
              VD_REQUIRE(rs >= 0, "vd_jpeg: corrupt AC code");
              const int r = rs >> 4, s = rs & 15;
              if (s == 0) {
                if (r != 15) break;
                k += 16;
                continue;
              }
              k += r;
              VD_REQUIRE(k < 64, "vd_jpeg: coefficient index out of range");
              blk[kZigzag[k]] = static_cast<int16_t>(extend(br.get(s), s));
              ++k;
            }
          }
        }
      }
    }
  }
  return VD_OK;
}
