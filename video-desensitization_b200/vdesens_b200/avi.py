"""Motion-JPEG in AVI (RIFF) without a codec library: the container side of SURVEY §8f row N3.

The reference turns a video into a directory of JPEG files with an ffmpeg subprocess (combine_detect.py:279-476),
runs the frame loop over the files and has cv2.VideoWriter re-encode the processed files into the output container
(:479-595; `.avi` -> XVID / MJPG, :505).  A Motion-JPEG AVI already IS a sequence of JPEG files: `AviReader` hands the
frames' bytes to the device decoder (vdesens_b200.jpeg.JpegDecoder) straight from the mapped file, `AviWriter` stores
what the device encoder produced — no pixel ever crosses PCIe and nothing is written to a scratch directory.

Format notes (AVI 1.0 + the OpenDML extension, both as ffmpeg / VirtualDub write them):
  RIFF 'AVI ' { LIST hdrl { avih, LIST strl { strh, strf, indx }, LIST odml { dmlh } }, LIST movi { 00dc ..., ix00 }, idx1 }
  RIFF 'AVIX' { LIST movi { 00dc ..., ix00 } }        one more per ~1 GB, so a file may pass the 32-bit RIFF size
Chunks are padded to even sizes.  `idx1` offsets count from the 'movi' fourcc; `ix00` entries point at the chunk DATA
relative to the index's base offset; the super index `indx` lists the `ix00` chunks with their absolute file offsets.
"""
import mmap
import os
import struct
from fractions import Fraction

AVIF_HASINDEX = 0x10
AVIF_ISINTERLEAVED = 0x100
AVIIF_KEYFRAME = 0x10
_MASTER_ENTRIES = 256               # super-index slots reserved in the header (ffmpeg reserves the same number)
_MJPG_HANDLERS = (b"MJPG", b"mjpg", b"MJPA", b"mjpa", b"JPEG", b"jpeg", b"AVRn", b"AVDJ", b"dmb1", b"IJPG")


# The four "typical" Huffman tables of ITU-T T.81 Annex K.3 as one DHT segment (DC / AC luminance, DC / AC chrominance):
# what libjpeg writes when it does not optimise its tables, and what an ABBREVIATED Motion-JPEG frame (no DHT segment of
# its own: capture hardware, the Windows MJPG codec, "AVI1" streams) implies.
STD_DHT = bytes.fromhex(
    "ffc401a20000010501010101010100000000000000000102030405060708090a0b100002010303020403050504040000017d0102030004110512"
    "2131410613516107227114328191a1082342b1c11552d1f02433627282090a161718191a25262728292a3435363738393a434445464748494a53"
    "5455565758595a636465666768696a737475767778797a838485868788898a92939495969798999aa2a3a4a5a6a7a8a9aab2b3b4b5b6b7b8b9ba"
    "c2c3c4c5c6c7c8c9cad2d3d4d5d6d7d8d9dae1e2e3e4e5e6e7e8e9eaf1f2f3f4f5f6f7f8f9fa0100030101010101010101010000000000000102"
    "030405060708090a0b1100020102040403040705040400010277000102031104052131061241510761711322328108144291a1b1c109233352f0"
    "156272d10a162434e125f11718191a262728292a35363738393a434445464748494a535455565758595a636465666768696a737475767778797a"
    "82838485868788898a92939495969798999aa2a3a4a5a6a7a8a9aab2b3b4b5b6b7b8b9bac2c3c4c5c6c7c8c9cad2d3d4d5d6d7d8d9dae2e3e4e5"
    "e6e7e8e9eaf2f3f4f5f6f7f8f9fa")


class AviError(ValueError):
    pass


def ensure_dht(blob):
    """A Motion-JPEG frame that carries no Huffman tables gets the standard ones spliced in before its scan, which makes
    it a complete JPEG file for libjpeg and for the device decoder (ffmpeg's mjpeg decoder assumes the same tables; its
    `mjpeg2jpeg` bitstream filter does this splice).  Frames that define tables — or are not JPEG — come back unchanged."""
    n = len(blob)
    if n < 4 or blob[0] != 0xFF or blob[1] != 0xD8:
        return blob
    pos = 2
    while pos + 4 <= n:
        if blob[pos] != 0xFF:
            return blob                           # not a marker where one has to be: leave it to the decoder
        m = blob[pos + 1]
        if m == 0xFF:                             # fill byte
            pos += 1
        elif m == 0xC4:
            return blob
        elif m == 0xDA:
            return blob[:pos] + STD_DHT + blob[pos:]
        elif m == 0x01 or 0xD0 <= m <= 0xD8:      # markers without a length
            pos += 2
        else:
            pos += 2 + ((blob[pos + 2] << 8) | blob[pos + 3])
    return blob


def _u32(buf, pos):
    return struct.unpack_from("<I", buf, pos)[0]


class AviReader:
    """Frames of the first video stream of an AVI file as byte strings, in presentation order.

    Attributes: width, height, fps (float), rate / scale (the exact ratio), handler (strh fccHandler), compression
    (strf biCompression), is_mjpeg, and len() = number of frames found in the file (a zero-length chunk — AVI's "same
    picture again" — repeats the previous frame's bytes).  The 'movi' lists are walked directly (every RIFF segment,
    'rec ' groups included), so a file whose index is missing or truncated still reads."""

    def __init__(self, path):
        self.path = path
        self._f = open(path, "rb")
        size = os.fstat(self._f.fileno()).st_size
        if size < 12:
            self._f.close()
            raise AviError(f"{path}: not an AVI file")
        self._m = mmap.mmap(self._f.fileno(), 0, access=mmap.ACCESS_READ)
        m = self._m
        if m[0:4] != b"RIFF" or m[8:12] != b"AVI ":
            self.close()
            raise AviError(f"{path}: not an AVI file")
        self.width = self.height = 0
        self.rate, self.scale = 0, 1
        self.handler = self.compression = b""
        self._usec = 0
        self._stream = None                      # index of the first video stream
        self._nstreams = 0
        self._frames = []                        # (offset, size)
        pos = 0
        while pos + 12 <= size and m[pos:pos + 4] == b"RIFF":
            rsize = _u32(m, pos + 4)
            end = min(size, pos + 8 + rsize) if rsize else size
            if m[pos + 8:pos + 12] in (b"AVI ", b"AVIX"):
                self._walk(pos + 12, end, top=True)
            pos = pos + 8 + rsize + (rsize & 1)
            if rsize == 0:
                break
        if self._stream is None:
            self.close()
            raise AviError(f"{path}: no video stream")
        if not self.rate or not self.scale:
            self.rate, self.scale = (1000000, self._usec) if self._usec else (25, 1)
        self.fps = self.rate / self.scale
        self.is_mjpeg = self.compression in _MJPG_HANDLERS or self.handler in _MJPG_HANDLERS

    def _walk(self, pos, end, top=False):
        m = self._m
        while pos + 8 <= end:
            cid, csz = m[pos:pos + 4], _u32(m, pos + 4)
            body = pos + 8
            stop = min(end, body + csz)
            if cid == b"LIST":
                kind = m[body:body + 4]
                if kind == b"hdrl":
                    self._headers(body + 4, stop)
                elif kind in (b"movi", b"rec "):
                    self._walk(body + 4, stop)
            elif not top and self._stream is not None and len(cid) == 4 and cid[2:4] in (b"dc", b"db") \
                    and cid[:2] == b"%02d" % self._stream:
                if body + csz > end:
                    return                       # truncated file: stop at the last complete frame
                if csz == 0 and self._frames:
                    self._frames.append(self._frames[-1])
                elif csz:
                    self._frames.append((body, csz))
            pos = body + csz + (csz & 1)

    def _headers(self, pos, end):
        m = self._m
        while pos + 8 <= end:
            cid, csz = m[pos:pos + 4], _u32(m, pos + 4)
            body = pos + 8
            if cid == b"avih" and csz >= 40:
                self._usec = _u32(m, body)
                self.width, self.height = _u32(m, body + 32), _u32(m, body + 36)
            elif cid == b"LIST" and m[body:body + 4] == b"strl":
                self._strl(body + 4, min(end, body + csz))
            pos = body + csz + (csz & 1)

    def _strl(self, pos, end):
        m = self._m
        index = self._nstreams
        self._nstreams += 1
        video = False
        while pos + 8 <= end:
            cid, csz = m[pos:pos + 4], _u32(m, pos + 4)
            body = pos + 8
            if cid == b"strh" and csz >= 36:
                video = m[body:body + 4] == b"vids" and self._stream is None
                if video:
                    self._stream = index
                    self.handler = bytes(m[body + 4:body + 8])
                    self.scale, self.rate = _u32(m, body + 20), _u32(m, body + 24)
            elif cid == b"strf" and video and csz >= 20:
                w, h = struct.unpack_from("<ii", m, body + 4)
                self.width, self.height = abs(w), abs(h)
                self.compression = bytes(m[body + 16:body + 20])
            pos = body + csz + (csz & 1)

    def __len__(self):
        return len(self._frames)

    def frame(self, i):
        """bytes of frame i — one complete JPEG file when is_mjpeg (an abbreviated frame gets the standard Huffman
        tables, `ensure_dht`)"""
        off, n = self._frames[i]
        blob = self._m[off:off + n]
        return ensure_dht(blob) if self.is_mjpeg else blob

    def frames(self, indices):
        return [self.frame(i) for i in indices]

    def __iter__(self):
        for i in range(len(self._frames)):
            yield self.frame(i)

    def close(self):
        if getattr(self, "_m", None) is not None:
            self._m.close()
            self._m = None
        if getattr(self, "_f", None) is not None:
            self._f.close()
            self._f = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def fps_ratio(fps):
    """frames per second as the (rate, scale) pair of the stream header: 60 -> (60, 1), 29.97 -> (2997, 100)"""
    fr = Fraction(str(float(fps))).limit_denominator(100000)
    if fr <= 0:
        raise ValueError(f"fps must be positive, got {fps!r}")
    return fr.numerator, fr.denominator


class AviWriter:
    """Writes a Motion-JPEG AVI: `write(jpeg_bytes)` per frame, `close()` patches the headers.  Every frame must be a
    complete JPEG file of width x height (what vdesens_b200.jpeg.JpegEncoder returns, or cv2.imencode('.jpg', ...)).
    A new RIFF 'AVIX' segment starts when the current one would pass `riff_limit` bytes (OpenDML; 1 GB like ffmpeg),
    each segment carrying its own `ix00` index, so outputs larger than 4 GB stay well-formed."""

    def __init__(self, path, width, height, fps, fourcc=b"MJPG", riff_limit=1 << 30):
        if isinstance(fourcc, str):
            fourcc = fourcc.encode("ascii")
        if len(fourcc) != 4:
            raise ValueError("fourcc must have four characters")
        self.path, self.width, self.height = path, int(width), int(height)
        self.rate, self.scale = fps_ratio(fps)
        self.fourcc = bytes(fourcc)
        self.riff_limit = int(riff_limit)
        self.count = 0
        self._max_chunk = 0
        self._bytes = 0
        self._f = open(path, "wb")
        self._segments = []                      # per closed segment: (ix00 offset, ix00 size with header, frames)
        self._seg_frames = []                    # (data offset, size) of the frames of the open segment
        self._first_index = []                   # idx1 entries of the first segment
        self._riff_start = 0
        self._movi_start = 0
        self._first_frames = 0
        self._write_header()
        self._open_movi()

    # ---- header ------------------------------------------------------------------------------------------------
    def _header_bytes(self, total, first_riff_frames):
        usec = int(round(1e6 * self.scale / self.rate))
        max_bps = int(self._bytes * self.rate / (self.scale * max(total, 1))) if total else 0
        avih = struct.pack("<14I", usec, min(max_bps, 0xFFFFFFFF), 0, AVIF_HASINDEX | AVIF_ISINTERLEAVED,
                           first_riff_frames, 0, 1, self._max_chunk, self.width, self.height, 0, 0, 0, 0)
        strh = struct.pack("<4s4sIHHIIIIIIIi4h", b"vids", self.fourcc, 0, 0, 0, 0, self.scale, self.rate, 0, total,
                           self._max_chunk, 0xFFFFFFFF, 0, 0, 0, self.width, self.height)
        strf = struct.pack("<IiiHH4sIiiII", 40, self.width, self.height, 1, 24, self.fourcc,
                           self.width * self.height * 3, 0, 0, 0, 0)
        segs = self._segments[:_MASTER_ENTRIES]
        indx = struct.pack("<HBBI4sIII", 4, 0, 0, len(segs), b"00dc", 0, 0, 0)
        for off, size, frames in segs:
            indx += struct.pack("<QII", off, size, frames)
        indx += bytes(16 * (_MASTER_ENTRIES - len(segs)))
        strl = b"strl" + self._chunk(b"strh", strh) + self._chunk(b"strf", strf) + self._chunk(b"indx", indx)
        odml = b"odml" + self._chunk(b"dmlh", struct.pack("<I", total) + bytes(244))
        return self._chunk(b"LIST", b"hdrl" + self._chunk(b"avih", avih) + self._chunk(b"LIST", strl)
                           + self._chunk(b"LIST", odml))

    @staticmethod
    def _chunk(cid, body):
        return cid + struct.pack("<I", len(body)) + body + (b"\0" if len(body) & 1 else b"")

    def _write_header(self):
        self._f.write(b"RIFF\0\0\0\0AVI ")
        self._f.write(self._header_bytes(0, 0))

    # ---- segments ----------------------------------------------------------------------------------------------
    def _open_movi(self):
        self._movi_start = self._f.tell()
        self._f.write(b"LIST\0\0\0\0movi")
        self._seg_frames = []

    def _close_segment(self):
        """ix00 at the end of the movi list, then (first segment) idx1; patches the LIST and RIFF sizes"""
        f = self._f
        base = self._movi_start + 8               # the 'movi' fourcc: entries are relative to it
        ix_off = f.tell()
        ix = struct.pack("<HBBI4sQI", 2, 0, 1, len(self._seg_frames), b"00dc", base, 0)
        ix += b"".join(struct.pack("<II", off - base, size) for off, size in self._seg_frames)
        f.write(self._chunk(b"ix00", ix))
        movi_end = f.tell()
        self._segments.append((ix_off, movi_end - ix_off, len(self._seg_frames)))
        if not self._riff_start:                   # first segment: legacy index for AVI 1.0 readers
            idx = b"".join(struct.pack("<4sIII", b"00dc", AVIIF_KEYFRAME, off - 8 - base, size)
                           for off, size in self._seg_frames)
            f.write(self._chunk(b"idx1", idx))
            self._first_frames = len(self._seg_frames)
        end = f.tell()
        f.seek(self._movi_start + 4)
        f.write(struct.pack("<I", movi_end - self._movi_start - 8))
        f.seek(self._riff_start + 4)
        f.write(struct.pack("<I", end - self._riff_start - 8))
        f.seek(end)

    def write(self, blob):
        if self._f is None:
            raise ValueError("AviWriter is closed")
        n = len(blob)
        if n == 0:
            raise ValueError("empty frame")
        f = self._f
        # 8 bytes per ix00 entry, 16 per idx1 entry, plus the index headers, still have to fit behind this frame
        projected = f.tell() + 8 + n + 1 + 24 * (len(self._seg_frames) + 1) + 64 - self._riff_start
        if self._seg_frames and projected > self.riff_limit:
            self._close_segment()
            self._riff_start = f.tell()
            f.write(b"RIFF\0\0\0\0AVIX")
            self._open_movi()
        f.write(b"00dc" + struct.pack("<I", n))
        self._seg_frames.append((f.tell(), n))
        f.write(blob)
        if n & 1:
            f.write(b"\0")
        self.count += 1
        self._bytes += n
        self._max_chunk = max(self._max_chunk, n)

    def close(self):
        if self._f is None:
            return
        self._close_segment()
        hdr = self._header_bytes(self.count, self._first_frames)
        self._f.seek(12)
        self._f.write(hdr)
        self._f.close()
        self._f = None

    def __enter__(self):
        return self

    def __exit__(self, exc_type, *exc):
        self.close()
