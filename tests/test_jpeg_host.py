"""libvdesens' host-side JPEG entropy decoder (csrc/jpeg_host.cu, SURVEY §8f row N1 — host half only, nothing in the
frame loop uses it yet) against oracle/jpeg_oracle.py, which is itself pinned bit-exact against cv2.imdecode.  CPU tests:
the functions are host code and need no GPU."""
import numpy as np
import pytest
import cv2

from oracle import jpeg_oracle as jo
from vdesens_b200 import jpeg, synth

S = cv2.IMWRITE_JPEG_SAMPLING_FACTOR
SF = {"444": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, "422": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422,
      "420": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420}


def _encode(img, params):
    ok, buf = cv2.imencode(".jpg", img, params)
    assert ok
    return buf


@pytest.mark.parametrize("sampling", ["420", "422", "444"])
@pytest.mark.parametrize("hw,quality,rst", [((48, 64), 95, 0), ((45, 77), 95, 3), ((135, 241), 60, 0), ((33, 47), 100, 1),
                                            ((8, 3), 95, 0), ((270, 480), 30, 0)])
def test_coefficients_match_the_oracle(sampling, hw, quality, rst):
    rng = np.random.default_rng(hw[0] + quality)
    for img in (synth.synthetic_frame(hw[0], hw[1], quality), rng.integers(0, 256, (hw[0], hw[1], 3), dtype=np.uint8)):
        params = [S, SF[sampling], cv2.IMWRITE_JPEG_QUALITY, quality] + ([cv2.IMWRITE_JPEG_RST_INTERVAL, rst] if rst else [])
        data = _encode(img, params).tobytes()
        info, coefs, quant = jpeg.decode_coefficients(data)
        j = jo.parse(data)
        want = jo.decode_coefficients(j)
        assert (info.width, info.height, info.ncomp) == (hw[1], hw[0], 3) and info.restart_interval == j["restart"]
        for c, comp in enumerate(j["comps"]):
            assert (info.h[c], info.v[c]) == (comp.h, comp.v)
            assert coefs[c].shape == want[c].shape and np.array_equal(coefs[c].astype(np.int32), want[c])
            assert np.array_equal(quant[c].astype(np.int32), j["qt"][comp.tq])


def test_host_coefficients_through_the_oracle_stages_equal_cv2():
    """entropy decoding by the library, everything after it by the oracle's libjpeg restatement == cv2.imdecode; also a
    greyscale file and one full-HD frame (the size the frame loop reads)"""
    for img, params in [(synth.synthetic_frame(1080, 1920, 1000), []), (synth.synthetic_frame(64, 80, 6)[..., 0].copy(), [])]:
        buf = _encode(img, params)
        data = buf.tobytes()
        info, coefs, quant = jpeg.decode_coefficients(data)
        planes = [jo.blocks_to_plane(jo.idct_islow(cf.astype(np.int32), q.astype(np.int32))) for cf, q in zip(coefs, quant)]
        H, W = info.height, info.width
        if info.ncomp == 1:
            got = np.stack([planes[0][:H, :W]] * 3, -1)
        else:
            cb, cr = (jo.h2v2_fancy(p, -(-W // 2), -(-H // 2))[:H, :W] for p in planes[1:])
            got = jo.ycc_to_rgb(planes[0][:H, :W], cb, cr)
        want = cv2.cvtColor(cv2.imdecode(buf, cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)
        assert np.array_equal(got, want)


def test_unsupported_and_broken_streams_raise_instead_of_crashing():
    img = synth.synthetic_frame(64, 96, 1)
    prog = _encode(img, [cv2.IMWRITE_JPEG_PROGRESSIVE, 1]).tobytes()
    with pytest.raises(RuntimeError, match="baseline"):
        jpeg.decode_coefficients(prog)
    with pytest.raises(RuntimeError, match="not a JPEG"):
        jpeg.read_header(b"\x89PNG\r\n\x1a\n" + bytes(32))
    good = _encode(img, []).tobytes()
    with pytest.raises(RuntimeError):
        jpeg.read_header(good[:40])                              # header cut inside a segment
    # a scan cut short decodes (zeros are fed past the end, like libjpeg's warning path) or raises — it must not crash
    cut = good[:len(good) // 2]
    try:
        info, coefs, _ = jpeg.decode_coefficients(cut)
        assert coefs[0].shape == (8, 12, 64)
    except RuntimeError:
        pass
    # random garbage after a valid header
    rng = np.random.default_rng(0)
    sos = good.index(b"\xff\xda")
    for _ in range(20):
        junk = good[:sos + 14] + rng.integers(0, 256, 600, dtype=np.uint8).tobytes()
        try:
            jpeg.decode_coefficients(junk)
        except RuntimeError:
            pass


def test_corrupt_huffman_table_is_rejected():
    good = _encode(synth.synthetic_frame(32, 32, 2), []).tobytes()
    i = good.index(b"\xff\xc4")
    bad = bytearray(good)
    bad[i + 5] = 200                       # 200 codes of length 1: over-subscribed
    with pytest.raises(RuntimeError):
        jpeg.decode_coefficients(bytes(bad))
    # fuzz the table segments: never a crash
    rng = np.random.default_rng(1)
    for _ in range(200):
        b = bytearray(good)
        for _ in range(4):
            b[int(rng.integers(i, i + 420))] = int(rng.integers(0, 256))
        try:
            jpeg.decode_coefficients(bytes(b))
        except RuntimeError:
            pass


def test_tables_of_the_previous_file_are_never_reused():
    """parse() runs on a thread_local scratch object.  A file WITHOUT a DHT segment (abbreviated MJPEG
    frame) decoded right after a normal file on the same thread must fail with "missing table", not be entropy-decoded
    with the previous file's Huffman tables; same for a scan that names one component twice."""
    img = synth.synthetic_frame(48, 64, 3)
    good = _encode(img, []).tobytes()
    jpeg.decode_coefficients(good)

    def strip(data, marker):
        out, pos = bytearray(data[:2]), 2
        while data[pos + 1] != 0xDA:
            ln = (data[pos + 2] << 8) | data[pos + 3]
            if data[pos + 1] != marker:
                out += data[pos:pos + 2 + ln]
            pos += 2 + ln
        return bytes(out + data[pos:])

    no_dht = strip(good, 0xC4)
    assert len(no_dht) < len(good) - 400
    with pytest.raises(RuntimeError, match="missing table"):
        jpeg.decode_coefficients(no_dht)
    jpeg.decode_coefficients(good)
    with pytest.raises(RuntimeError, match="missing table"):
        jpeg.decode_coefficients(strip(good, 0xDB))
    # duplicate component id in the scan header
    b = bytearray(good)
    sos = b.index(b"\xff\xda")
    assert b[sos + 4] == 3
    b[sos + 7] = b[sos + 5]                 # second scan component := first component's id
    jpeg.decode_coefficients(good)
    with pytest.raises(RuntimeError, match="twice"):
        jpeg.decode_coefficients(bytes(b))
