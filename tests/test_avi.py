"""CPU: the Motion-JPEG AVI container code (vdesens_b200/avi.py) against OpenCV's ffmpeg backend —
the library behind the reference's own video stage (combine_detect.py:479-595 cv2.VideoWriter; :349-359 ffmpeg) — and
the file-level drop-ins convert_video_to_frames / create_video (reference :279-476, :479-595).  No GPU needed."""
import os

import cv2
import numpy as np
import pytest

from vdesens_b200 import avi, synth


def _jpeg(rgb, quality=95):
    ok, buf = cv2.imencode(".jpg", cv2.cvtColor(rgb, cv2.COLOR_RGB2BGR), [cv2.IMWRITE_JPEG_QUALITY, quality])
    assert ok
    return buf.tobytes()


def _decode(blob):
    return cv2.cvtColor(cv2.imdecode(np.frombuffer(blob, np.uint8), cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)


def _frames(n, h=120, w=160, seed=0):
    return [synth.synthetic_frame(h, w, seed + i) for i in range(n)]


def _closest(picture, candidates, tol=8.0):
    """index of the candidate this picture shows (ffmpeg's and libjpeg's decodes of one JPEG differ by their chroma
    up-sampling: a few grey levels on these noisy frames), -1 if none is within `tol` mean absolute difference"""
    d = [np.abs(picture.astype(np.int32) - c).mean() for c in candidates]
    i = int(np.argmin(d))
    return i if d[i] < tol and all(d[j] > 2 * d[i] for j in range(len(d)) if j != i) else -1


def _capture_all(path):
    cap = cv2.VideoCapture(str(path))
    assert cap.isOpened()
    meta = (int(cap.get(cv2.CAP_PROP_FRAME_COUNT)), cap.get(cv2.CAP_PROP_FPS), int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)),
            int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT)))
    out = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        out.append(cv2.cvtColor(f, cv2.COLOR_BGR2RGB))
    cap.release()
    return meta, out


@pytest.mark.parametrize("riff_limit", [1 << 30, 60000, 20000])
@pytest.mark.parametrize("fps", [60, 29.97, 12.5])
def test_writer_reader_round_trip_and_opencv_reads_the_file(tmp_path, riff_limit, fps):
    """what AviWriter stores comes back byte for byte from AviReader (one RIFF segment or several OpenDML 'AVIX' ones,
    odd-sized chunks padded), and ffmpeg (through cv2.VideoCapture) sees the same stream: frame count, rate, size, and
    pictures equal to libjpeg's decode of the same bytes up to the two decoders' chroma up-sampling difference"""
    frames = _frames(23)
    blobs = [_jpeg(f, 95 - (i % 3)) for i, f in enumerate(frames)]
    assert any(len(b) & 1 for b in blobs) and any(not len(b) & 1 for b in blobs)
    path = tmp_path / "out.avi"
    with avi.AviWriter(str(path), 160, 120, fps, riff_limit=riff_limit) as wr:
        for b in blobs:
            wr.write(b)
    assert wr.count == 23 and (len(wr._segments) > 1) == (riff_limit < (1 << 30))
    with avi.AviReader(str(path)) as rd:
        assert (len(rd), rd.width, rd.height, rd.is_mjpeg, rd.compression) == (23, 160, 120, True, b"MJPG")
        assert abs(rd.fps - fps) < 1e-9
        assert [rd.frame(i) for i in range(23)] == blobs and list(rd) == blobs and rd.frames([3, 1]) == [blobs[3], blobs[1]]
    (count, cfps, w, h), got = _capture_all(path)
    assert (count, w, h) == (23, 160, 120) and abs(cfps - fps) < 1e-3 and len(got) == 23
    want = [_decode(b) for b in blobs]
    for i, g in enumerate(got):
        assert _closest(g, want) == i
    # random access through the indexes (idx1 in the first segment, ix00 / indx beyond it)
    cap = cv2.VideoCapture(str(path))
    for i in (17, 4, 22):
        cap.set(cv2.CAP_PROP_POS_FRAMES, i)
        ok, f = cap.read()
        assert ok and _closest(cv2.cvtColor(f, cv2.COLOR_BGR2RGB), want) == i
    cap.release()


def test_index_chunks_point_at_the_frames(tmp_path):
    """idx1 offsets count from the 'movi' fourcc to the chunk header; ix00 entries from the base offset to the DATA; the
    super index lists every ix00 chunk with its absolute position and frame count"""
    import struct
    blobs = [_jpeg(f) for f in _frames(9, 48, 64)]
    path = tmp_path / "i.avi"
    with avi.AviWriter(str(path), 64, 48, 30, riff_limit=12000) as wr:
        for b in blobs:
            wr.write(b)
    data = path.read_bytes()
    assert struct.unpack_from("<I", data, 4)[0] + 8 <= len(data) and data[8:12] == b"AVI "
    movi = data.index(b"movi")
    idx1 = data.index(b"idx1")
    n1 = struct.unpack_from("<I", data, idx1 + 4)[0] // 16
    for k in range(n1):
        cid, flags, off, size = struct.unpack_from("<4sIII", data, idx1 + 8 + 16 * k)
        assert cid == b"00dc" and flags == avi.AVIIF_KEYFRAME and size == len(blobs[k])
        assert data[movi + off:movi + off + 4] == b"00dc" and data[movi + off + 8:movi + off + 8 + size] == blobs[k]
    indx = data.index(b"indx")
    longs, sub, typ, used, cid = struct.unpack_from("<HBBI4s", data, indx + 8)
    assert (longs, sub, typ, cid) == (4, 0, 0, b"00dc") and used == len(wr._segments) > 1
    seen = 0
    for s in range(used):
        off, size, dur = struct.unpack_from("<QII", data, indx + 8 + 24 + 16 * s)
        assert data[off:off + 4] == b"ix00" and struct.unpack_from("<I", data, off + 4)[0] + 8 == size
        l2, sub2, typ2, n, cid2, base, _ = struct.unpack_from("<HBBI4sQI", data, off + 8)
        assert (l2, sub2, typ2, cid2, n) == (2, 0, 1, b"00dc", dur)
        for k in range(n):
            o, sz = struct.unpack_from("<II", data, off + 8 + 24 + 8 * k)
            assert data[base + o:base + o + sz] == blobs[seen] and data[base + o - 8:base + o - 4] == b"00dc"
            seen += 1
    assert seen == 9
    assert struct.unpack_from("<I", data, data.index(b"dmlh") + 8)[0] == 9


def test_reader_takes_what_opencv_writes(tmp_path):
    """a Motion-JPEG AVI written by ffmpeg (cv2.VideoWriter 'MJPG', the reference's own writer, :505): every frame is a
    complete baseline JPEG the device decoder's host parser accepts, and libjpeg's decode of it is the picture ffmpeg
    shows; an XVID file is an AVI but not Motion-JPEG"""
    from vdesens_b200 import jpeg as vjpeg
    frames = _frames(7, 96, 128, seed=40)
    path = tmp_path / "cv.avi"
    wr = cv2.VideoWriter(str(path), cv2.VideoWriter_fourcc(*"MJPG"), 25, (128, 96))
    assert wr.isOpened()
    for f in frames:
        wr.write(cv2.cvtColor(f, cv2.COLOR_RGB2BGR))
    wr.release()
    _, shown = _capture_all(path)
    with avi.AviReader(str(path)) as rd:
        assert (len(rd), rd.width, rd.height, rd.is_mjpeg) == (7, 128, 96, True) and abs(rd.fps - 25) < 1e-9
        for i in range(7):
            blob = rd.frame(i)
            assert blob[:2] == b"\xff\xd8" and blob[-2:] == b"\xff\xd9"
            hdr = vjpeg.prepare_device(blob)
            assert (hdr.width, hdr.height) == (128, 96)
            assert _closest(_decode(blob), shown) == i
    xvid = tmp_path / "x.avi"
    wr = cv2.VideoWriter(str(xvid), cv2.VideoWriter_fourcc(*"XVID"), 25, (128, 96))
    if wr.isOpened():
        for f in frames:
            wr.write(cv2.cvtColor(f, cv2.COLOR_RGB2BGR))
        wr.release()
        with avi.AviReader(str(xvid)) as rd:
            assert not rd.is_mjpeg and len(rd) == 7


def test_reader_edge_cases(tmp_path):
    blobs = [_jpeg(f) for f in _frames(5, 48, 64)]
    path = tmp_path / "t.avi"
    with avi.AviWriter(str(path), 64, 48, 30) as wr:
        for b in blobs:
            wr.write(b)
    data = path.read_bytes()
    # truncated in the middle of the fourth frame (a recording that was cut off): the complete frames still read
    cut = data.index(blobs[3]) + len(blobs[3]) // 2
    (tmp_path / "cut.avi").write_bytes(data[:cut])
    with avi.AviReader(str(tmp_path / "cut.avi")) as rd:
        assert list(rd) == blobs[:3]
    # a zero-length chunk repeats the previous picture; 'rec ' groups and foreign chunks inside movi are walked / skipped
    import struct

    def chunk(cid, body):
        return cid + struct.pack("<I", len(body)) + body + (b"\0" if len(body) & 1 else b"")
    movi = b"movi" + chunk(b"00dc", blobs[0]) + chunk(b"01wb", b"audio") + chunk(b"00dc", b"") + \
        chunk(b"LIST", b"rec " + chunk(b"00dc", blobs[1]) + chunk(b"JUNK", b"xyz")) + chunk(b"00db", blobs[2])
    hdrl = data[12:data.index(b"LIST", data.index(b"odml") - 8 + 260)]          # the writer's header list as is
    body = b"AVI " + hdrl + chunk(b"LIST", movi)
    (tmp_path / "hand.avi").write_bytes(b"RIFF" + struct.pack("<I", len(body)) + body)
    with avi.AviReader(str(tmp_path / "hand.avi")) as rd:
        assert list(rd) == [blobs[0], blobs[0], blobs[1], blobs[2]]
    # not an AVI / no video stream / empty
    (tmp_path / "no.avi").write_bytes(b"RIFF\x04\0\0\0WAVE")
    (tmp_path / "short.avi").write_bytes(b"RIFF")
    for name in ("no.avi", "short.avi"):
        with pytest.raises(avi.AviError):
            avi.AviReader(str(tmp_path / name))
    with pytest.raises(OSError):
        avi.AviReader(str(tmp_path / "missing.avi"))
    with pytest.raises(ValueError):
        avi.AviWriter(str(tmp_path / "z.avi"), 64, 48, 0)
    wr = avi.AviWriter(str(tmp_path / "empty.avi"), 64, 48, 30)
    with pytest.raises(ValueError):
        wr.write(b"")
    wr.close()
    wr.close()                                   # idempotent
    with pytest.raises(ValueError):
        wr.write(blobs[0])
    with avi.AviReader(str(tmp_path / "empty.avi")) as rd:
        assert len(rd) == 0 and rd.is_mjpeg
    assert avi.fps_ratio(60) == (60, 1) and avi.fps_ratio(29.97) == (2997, 100) and avi.fps_ratio(23.976) == (2997, 125)


def test_convert_video_to_frames_and_create_video(tmp_path):
    """reference :279-476 / :479-595 as file-level drop-ins.  Motion-JPEG: the frame files ARE the container's chunks and
    the re-assembled video holds the files bit for bit (no transcode either way); other codecs go through OpenCV"""
    import combine_detect as cd
    blobs = [_jpeg(f) for f in _frames(11, 96, 128, seed=7)]
    src = tmp_path / "in.avi"
    with avi.AviWriter(str(src), 128, 96, 30) as wr:
        for b in blobs:
            wr.write(b)
    d = tmp_path / "frames"
    (d).mkdir()
    (d / "stale.jpg").write_bytes(b"x")          # the reference clears old frames first (:290-291)
    assert cd.convert_video_to_frames(str(src), str(d)) == 11
    assert sorted(os.listdir(d)) == ["frame_%06d.jpg" % i for i in range(1, 12)]
    assert [(d / ("frame_%06d.jpg" % (i + 1))).read_bytes() for i in range(11)] == blobs
    assert cd.convert_video_to_frames(str(src), str(tmp_path / "third"), interval=3) == 4
    assert [(tmp_path / "third" / ("frame_%06d.jpg" % (i + 1))).read_bytes() for i in range(4)] == blobs[0::3]
    out = tmp_path / "re" / "again.avi"
    assert cd.create_video(str(d), str(out), fps=60) is True
    with avi.AviReader(str(out)) as rd:
        assert list(rd) == blobs and rd.fps == 60
    assert cd.create_video(str(tmp_path / "nothing"), str(tmp_path / "n.avi")) is False
    # a non-MJPEG source decodes through OpenCV; an .mp4 target encodes through it (reference codec list)
    mp4 = tmp_path / "v.mp4"
    assert cd.create_video(str(d), str(mp4), fps=30) is True
    (count, fps, w, h), shown = _capture_all(mp4)
    assert (count, w, h) == (11, 128, 96) and abs(fps - 30) < 1e-3
    assert cd.convert_video_to_frames(str(mp4), str(tmp_path / "frommp4")) == 11
    back = cv2.cvtColor(cv2.imread(str(tmp_path / "frommp4" / "frame_000001.jpg")), cv2.COLOR_BGR2RGB)
    assert back.shape == (96, 128, 3) and np.abs(back.astype(np.int32) - _decode(blobs[0])).mean() < 12.0
    assert cd.convert_video_to_frames(str(tmp_path / "missing.mp4"), str(tmp_path / "m")) == 0
    # frames of different sizes cannot be muxed as they are: OpenCV path (resizes to the first frame, reference :580-581)
    mixed = tmp_path / "mixed"
    mixed.mkdir()
    (mixed / "a.jpg").write_bytes(blobs[0])
    (mixed / "b.jpg").write_bytes(_jpeg(_frames(1, 48, 64)[0]))
    assert cd.create_video(str(mixed), str(tmp_path / "mixed.avi"), fps=30) is True
    (count, _, w, h), _ = _capture_all(tmp_path / "mixed.avi")
    assert (count, w, h) == (2, 128, 96)


def test_video_pipeline_argument_checks(tmp_path):
    import combine_detect as cd
    with pytest.raises(ValueError):
        cd.process_video_pipeline("a.avi", "b.avi", None, None, mode="blur")
    with pytest.raises(ValueError):
        cd.process_video_pipeline("a.avi", "b.avi", None, None, on_error="ignore")
    assert cd.process_single_video(str(tmp_path / "noext"), str(tmp_path), None, None) is False
    assert cd.process_single_video(str(tmp_path / "missing.avi"), str(tmp_path), None, None) is False


def test_abbreviated_mjpeg_frames_get_the_standard_huffman_tables(tmp_path):
    """capture hardware and the Windows MJPG codec store frames WITHOUT a DHT segment (the T.81 Annex K tables are implied):
    classic libjpeg refuses such a file, ffmpeg and newer libjpeg-turbo assume the tables.  AviReader splices them in, after which the frame is the
    complete file libjpeg wrote (cv2 uses exactly these tables), decodes to the same pixels, and is what the device
    decoder's host parser takes"""
    import struct
    from vdesens_b200 import jpeg as vjpeg

    def strip_dht(blob):
        out, pos = bytearray(blob[:2]), 2
        while True:
            m, L = blob[pos + 1], struct.unpack(">H", blob[pos + 2:pos + 4])[0]
            if m == 0xDA:
                return bytes(out) + blob[pos:]
            if m != 0xC4:
                out += blob[pos:pos + 2 + L]
            pos += 2 + L

    full = [_jpeg(f) for f in _frames(4, 96, 128, seed=90)]
    short = [strip_dht(b) for b in full]
    assert all(len(s) == len(b) - 4 * 4 - 416 for s, b in zip(short, full))          # four DHT segments gone
    raw = cv2.imdecode(np.frombuffer(short[0], np.uint8), cv2.IMREAD_COLOR)    # classic libjpeg: "table not defined" (None);
    assert raw is None or np.array_equal(raw[..., ::-1], _decode(full[0]))       # libjpeg-turbo >= 2.1 assumes Annex K too
    with pytest.raises(RuntimeError):                                            # the device decoder's parser wants the tables
        vjpeg.prepare_device(short[0])
    fixed = avi.ensure_dht(short[0])
    assert np.array_equal(_decode(fixed), _decode(full[0])) and len(fixed) == len(short[0]) + len(avi.STD_DHT)
    assert avi.ensure_dht(full[0]) is full[0] and avi.ensure_dht(b"RIFF") == b"RIFF" and avi.ensure_dht(b"") == b""
    assert avi.ensure_dht(b"\xff\xd8\xff\xd9") == b"\xff\xd8\xff\xd9"
    hdr = vjpeg.prepare_device(fixed)
    assert (hdr.width, hdr.height) == (128, 96)
    path = tmp_path / "abbr.avi"
    with avi.AviWriter(str(path), 128, 96, 30) as wr:
        for b in short:
            wr.write(b)
    with avi.AviReader(str(path)) as rd:
        got = list(rd)
    assert all(np.array_equal(_decode(g), _decode(f)) for g, f in zip(got, full))
    _, shown = _capture_all(path)                                   # ffmpeg plays the abbreviated stream as well
    assert len(shown) == 4 and all(_closest(s, [_decode(f) for f in full]) == i for i, s in enumerate(shown))
