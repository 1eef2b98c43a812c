"""GPU: the video stage (reference combine_detect.py:597-646 process_video_pipeline) with the frames
resident in HBM between the container's two ends.  The oracle for a Motion-JPEG output is the reference's own chain on
the same pixels: cv2.imdecode of the source frame -> detect + composite (batch_process_frames, itself pinned against the
oracle in test_gpu_kernels.py) -> cv2.imencode at quality 95 (= the `processed_*.jpg` the reference's loop writes and
its create_video reads back): every frame of the output file must equal those bytes."""
import warnings

import cv2
import numpy as np
import pytest

from vdesens_b200 import avi, synth

pytestmark = pytest.mark.gpu


def _jpeg(rgb, quality=95):
    ok, buf = cv2.imencode(".jpg", cv2.cvtColor(rgb, cv2.COLOR_RGB2BGR), [cv2.IMWRITE_JPEG_QUALITY, quality])
    assert ok
    return buf.tobytes()


def _decode(blob):
    return cv2.cvtColor(cv2.imdecode(np.frombuffer(blob, np.uint8), cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)


def _detector(H, W, batch):
    from detect_face.face import Retinaface
    sd = synth.synthetic_state_dict(0, face_rate="r01")
    return Retinaface(state_dict=sd, input_shape=[H, W, 3], confidence=0.5, nms_iou=0.4, letterbox_image=True, cuda=True,
                      batch_size=batch)


def _plates(batch, verbose=False, conf=0.5):
    assert all(isinstance(b, np.ndarray) and b.ndim == 3 for b in batch)
    return [(None, [[10 + k, 12, 60, 40]]) for k, _ in enumerate(batch)]


def _write_avi(path, blobs, w, h, fps=30):
    with avi.AviWriter(str(path), w, h, fps) as wr:
        for b in blobs:
            wr.write(b)


def _expected(det, frames, plates, batch, **kw):
    import combine_detect
    out, nf, npl = combine_detect.batch_process_frames(list(frames), det, plates, batch_size=batch, **kw)
    return [_jpeg(f) for f in out], nf, npl


@pytest.mark.parametrize("frame_hw,input_hw", [((160, 224), (160, 224)), ((96, 128), (160, 160))])
def test_mjpeg_avi_in_mjpeg_avi_out_equals_the_reference_chain(tmp_path, frame_hw, input_hw):
    """Motion-JPEG AVI -> device decode -> detect + blur -> device encode -> Motion-JPEG AVI: counts and every output
    frame's bytes equal the reference chain's; 10 frames in batches of 4 (4 + 4 + 2); frames of the network's own size and
    frames the device letterboxes (reference face.py:73-80)"""
    import combine_detect
    (h, w), (H, W) = frame_hw, input_hw
    det = _detector(H, W, 4)
    blobs = [_jpeg(synth.synthetic_frame(h, w, 5200 + i)) for i in range(10)]
    src, dst = tmp_path / "in.avi", tmp_path / "out" / "res.avi"
    _write_avi(src, blobs, w, h, fps=25)
    want, want_faces, want_plates = _expected(det, [_decode(b) for b in blobs], _plates, 4, mode="gauss", ksize=15)
    ok, n, nf, npl = combine_detect.process_video_pipeline(str(src), str(dst), det, _plates, temp_dir=str(tmp_path / "tmp"),
                                                           fps=60, batch_size=4, mode="gauss", ksize=15)
    assert (ok, n, nf, npl) == (True, 10, want_faces, want_plates) and nf > 0 and npl == 10
    assert not (tmp_path / "tmp").exists()                         # nothing goes through a scratch directory
    with avi.AviReader(str(dst)) as rd:
        assert (len(rd), rd.width, rd.height, rd.fps, rd.is_mjpeg) == (10, w, h, 60, True)
        for i in range(10):
            assert rd.frame(i) == want[i], f"frame {i} differs from the reference chain's file"
    cap = cv2.VideoCapture(str(dst))                               # and a stock reader plays it
    assert cap.isOpened() and int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == 10
    cap.release()
    # fps=None keeps the source's rate; mosaic mode; no plate provider
    want, want_faces, _ = _expected(det, [_decode(b) for b in blobs], None, 4, mode="mosaic", mosaic_level=6)
    ok, n, nf, npl = combine_detect.process_video_pipeline(str(src), str(tmp_path / "m.avi"), det, None, fps=None, batch_size=4,
                                                           mode="mosaic", mosaic_level=6)
    assert (ok, n, nf, npl) == (True, 10, want_faces, 0)
    with avi.AviReader(str(tmp_path / "m.avi")) as rd:
        assert rd.fps == 25 and list(rd) == want


def test_frames_the_device_decoder_does_not_take_and_broken_frames(tmp_path):
    """a progressive JPEG inside the stream is decoded by the host codec for that frame only (same pixels, so the same
    output); a frame that no decoder can read fails its batch: on_error='raise' raises, on_error='skip' drops exactly that
    batch (reference :226-228) and the remaining batches keep their order"""
    import combine_detect
    H, W = 160, 224
    det = _detector(H, W, 4)
    frames = [synth.synthetic_frame(H, W, 6100 + i) for i in range(12)]
    blobs = [_jpeg(f) for f in frames]
    ok, prog = cv2.imencode(".jpg", cv2.cvtColor(frames[5], cv2.COLOR_RGB2BGR), [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    blobs[5] = prog.tobytes()
    src = tmp_path / "p.avi"
    _write_avi(src, blobs, W, H)
    want, want_faces, _ = _expected(det, [_decode(b) for b in blobs], None, 4, mode="gauss", ksize=15)
    assert combine_detect.process_video_pipeline(str(src), str(tmp_path / "o.avi"), det, None, batch_size=4, mode="gauss",
                                                 ksize=15) == (True, 12, want_faces, 0)
    with avi.AviReader(str(tmp_path / "o.avi")) as rd:
        assert list(rd) == want
    bad = list(blobs)
    bad[6] = b"\xff\xd8\xff\xd9"                                   # batch 1 (frames 4..7) cannot be decoded at all
    _write_avi(tmp_path / "bad.avi", bad, W, H)
    with pytest.raises(Exception):
        combine_detect.process_video_pipeline(str(tmp_path / "bad.avi"), str(tmp_path / "b1.avi"), det, None, batch_size=4,
                                              mode="gauss", ksize=15)
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        ok, n, _, _ = combine_detect.process_video_pipeline(str(tmp_path / "bad.avi"), str(tmp_path / "b2.avi"), det, None,
                                                            batch_size=4, mode="gauss", ksize=15, on_error="skip")
    assert ok and n == 8 and any("dropped" in str(c.message) for c in caught)
    with avi.AviReader(str(tmp_path / "b2.avi")) as rd:
        assert list(rd) == want[:4] + want[8:]


def test_other_containers_go_through_opencv_at_either_end(tmp_path):
    """an input that is not Motion-JPEG is decoded by cv2.VideoCapture and uploaded; an output that is not .avi is encoded
    by cv2.VideoWriter with the reference's codec list (:505-512).  mp4v in -> MJPG out is still exact w.r.t. the frames
    OpenCV decodes; MJPG in -> mp4 out is compared as pictures (the codec is lossy)"""
    import combine_detect
    H, W = 160, 224
    det = _detector(H, W, 4)
    frames = [synth.synthetic_frame(H, W, 7300 + i) for i in range(9)]
    mp4 = tmp_path / "in.mp4"
    wr = cv2.VideoWriter(str(mp4), cv2.VideoWriter_fourcc(*"mp4v"), 30, (W, H))
    assert wr.isOpened()
    for f in frames:
        wr.write(cv2.cvtColor(f, cv2.COLOR_RGB2BGR))
    wr.release()
    cap = cv2.VideoCapture(str(mp4))
    shown = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        shown.append(cv2.cvtColor(f, cv2.COLOR_BGR2RGB))
    cap.release()
    assert len(shown) == 9
    want, want_faces, want_plates = _expected(det, shown, _plates, 4, mode="gauss", ksize=15)
    assert combine_detect.process_video_pipeline(str(mp4), str(tmp_path / "o.avi"), det, _plates, batch_size=4, mode="gauss",
                                                 ksize=15) == (True, 9, want_faces, want_plates)
    with avi.AviReader(str(tmp_path / "o.avi")) as rd:
        assert list(rd) == want
    # process_single_video names the output like the reference (:656-658) and keeps the container
    blobs = [_jpeg(f) for f in frames]
    _write_avi(tmp_path / "clip.avi", blobs, W, H)
    outdir = tmp_path / "videos"
    outdir.mkdir()
    assert combine_detect.process_single_video(str(tmp_path / "clip.avi"), str(outdir), det, None, str(tmp_path / "t"),
                                               batch_size=4, mode="gauss", ksize=15) is True
    want2, _, _ = _expected(det, [_decode(b) for b in blobs], None, 4, mode="gauss", ksize=15)
    with avi.AviReader(str(outdir / "clip_processed.avi")) as rd:
        assert list(rd) == want2 and rd.fps == 60
    # MJPG in -> .mp4 out
    ok, n, nf, _ = combine_detect.process_video_pipeline(str(tmp_path / "clip.avi"), str(tmp_path / "o.mp4"), det, None,
                                                         batch_size=4, mode="gauss", ksize=15)
    assert ok and n == 9
    cap = cv2.VideoCapture(str(tmp_path / "o.mp4"))
    assert int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == 9 and abs(cap.get(cv2.CAP_PROP_FPS) - 60) < 1e-3
    for i in range(9):
        ok, f = cap.read()
        assert ok
        d = np.abs(cv2.cvtColor(f, cv2.COLOR_BGR2RGB).astype(np.int32) - _decode(want2[i])).mean()
        assert d < 12.0, f"frame {i}: mean abs difference {d}"
    cap.release()
    # a file OpenCV cannot open: (False, 0, 0, 0) like the reference's failed extraction (:614-616)
    (tmp_path / "junk.mkv").write_bytes(b"not a video")
    assert combine_detect.process_video_pipeline(str(tmp_path / "junk.mkv"), str(tmp_path / "j.avi"), det, None) == (False, 0, 0, 0)


def test_abbreviated_mjpeg_stream_is_decoded_on_the_device(tmp_path):
    """frames without Huffman tables of their own (capture hardware, the Windows MJPG codec): AviReader splices the standard
    tables in, the device decoder takes every frame (no host fallback), and the output equals the reference chain's"""
    import struct
    import combine_detect
    from vdesens_b200.jpeg import JpegDecoder
    H, W = 160, 224
    det = _detector(H, W, 4)

    def strip_dht(blob):
        out, pos = bytearray(blob[:2]), 2
        while True:
            m, L = blob[pos + 1], struct.unpack(">H", blob[pos + 2:pos + 4])[0]
            if m == 0xDA:
                return bytes(out) + blob[pos:]
            if m != 0xC4:
                out += blob[pos:pos + 2 + L]
            pos += 2 + L

    full = [_jpeg(synth.synthetic_frame(H, W, 8800 + i)) for i in range(6)]
    _write_avi(tmp_path / "abbr.avi", [strip_dht(b) for b in full], W, H)
    with avi.AviReader(str(tmp_path / "abbr.avi")) as rd:
        dec = JpegDecoder(4, device="cuda:0")
        frames, on_cpu = dec.decode(rd.frames(range(4)))
        assert on_cpu == [] and all(np.array_equal(frames[i].cpu().numpy(), _decode(full[i])) for i in range(4))
    want, want_faces, _ = _expected(det, [_decode(b) for b in full], None, 4, mode="gauss", ksize=15)
    assert combine_detect.process_video_pipeline(str(tmp_path / "abbr.avi"), str(tmp_path / "o.avi"), det, None, batch_size=4,
                                                 mode="gauss", ksize=15) == (True, 6, want_faces, 0)
    with avi.AviReader(str(tmp_path / "o.avi")) as rd:
        assert list(rd) == want
