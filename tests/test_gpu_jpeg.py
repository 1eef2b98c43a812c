"""GPU: the device JPEG encoder (csrc/jpeg_enc.cu, SURVEY §8f row N1, reference combine_detect.py:176-180 cv2.imwrite)
through the C-ABI, byte for byte against cv2.imencode and against oracle/jpeg_enc_oracle.py (itself pinned on cv2)."""
import numpy as np
import pytest
import torch
import cv2

from oracle import jpeg_enc_oracle as je
from vdesens_b200 import synth
from vdesens_b200.jpeg import JpegEncoder

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _cv2(rgb, quality=95):
    ok, buf = cv2.imencode(".jpg", cv2.cvtColor(rgb, cv2.COLOR_RGB2BGR), [cv2.IMWRITE_JPEG_QUALITY, quality])
    assert ok
    return buf.tobytes()


def _image(h, w, kind, seed):
    if kind == "smooth":
        return synth.synthetic_frame(max(h, 8), max(w, 8), seed)[:h, :w].copy()
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def _first_diff(a, b):
    n = min(len(a), len(b))
    return next((i for i in range(n) if a[i] != b[i]), n)


@pytest.mark.parametrize("hw", [(16, 16), (8, 8), (17, 23), (48, 64), (45, 77), (135, 241), (1, 1), (9, 33), (24, 40), (40, 24),
                                (100, 100), (270, 480), (2, 300), (300, 2), (136, 130)])
@pytest.mark.parametrize("kind", ["smooth", "noise"])
def test_encoder_is_byte_identical_to_cv2(hw, kind):
    h, w = hw
    imgs = [_image(h, w, kind, h * 7 + w + i) for i in range(3)]
    enc = JpegEncoder(3, h, w, device=DEV, max_bytes_per_frame=h * w * 4 + 4096)
    got = enc.encode([torch.from_numpy(im).to(DEV) for im in imgs])
    for im, g in zip(imgs, got):
        want = _cv2(im)
        assert g == want, f"{hw} {kind}: first difference at byte {_first_diff(g, want)} of {len(g)}/{len(want)}"
        assert g == je.encode_rgb(im)


@pytest.mark.parametrize("quality", [1, 30, 50, 75, 100])
def test_encoder_other_qualities(quality):
    imgs = [_image(45, 77, k, quality) for k in ("smooth", "noise")]
    enc = JpegEncoder(2, 45, 77, device=DEV, quality=quality, max_bytes_per_frame=45 * 77 * 6 + 4096)
    for im, g in zip(imgs, enc.encode([torch.from_numpy(im).to(DEV) for im in imgs])):
        want = _cv2(im, quality)
        assert g == want, f"q={quality}: first difference at byte {_first_diff(g, want)} of {len(g)}/{len(want)}"


def test_encoder_full_hd_batch_and_views():
    """the frame loop's size: a 1080p batch (68th MCU row = dummy luma blocks), frames given as views of one batch tensor;
    decoding the device's file gives exactly the pixels decoding cv2's file gives (same bytes) and the file sizes match"""
    B, H, W = 4, 1080, 1920
    frames = np.stack([synth.synthetic_frame(H, W, 2000 + i) for i in range(B)])
    dev = torch.from_numpy(frames).to(DEV)
    enc = JpegEncoder(B, H, W, device=DEV)
    got = enc.encode(dev)
    for i in range(B):
        want = _cv2(frames[i])
        assert got[i] == want, f"frame {i}: first difference at byte {_first_diff(got[i], want)} of {len(got[i])}/{len(want)}"
    assert np.array_equal(cv2.imdecode(np.frombuffer(got[0], np.uint8), cv2.IMREAD_COLOR),
                          cv2.imdecode(np.frombuffer(_cv2(frames[0]), np.uint8), cv2.IMREAD_COLOR))
    # a second call on the same encoder (workspace re-zeroed) and a short batch
    again = enc.encode(dev[:2])
    assert again == got[:2]


def test_encoder_reports_a_frame_that_does_not_fit():
    im = _image(64, 64, "noise", 5)
    enc = JpegEncoder(1, 64, 64, device=DEV, max_bytes_per_frame=1024)
    with pytest.raises(RuntimeError, match="did not fit"):
        enc.encode([torch.from_numpy(im).to(DEV)])
    enc2 = JpegEncoder(1, 64, 64, device=DEV, max_bytes_per_frame=64 * 64 * 4)
    assert enc2.encode([torch.from_numpy(im).to(DEV)])[0] == _cv2(im)


# ------------------------------------------------------------------------------------------------
# device decoder (csrc/jpeg_dec.cu; reference combine_detect.py:167-173 cv2.imread + BGR->RGB)
# ------------------------------------------------------------------------------------------------
S = cv2.IMWRITE_JPEG_SAMPLING_FACTOR
SF = {"444": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, "422": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422,
      "420": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420}


def _file(rgb, params=()):
    src = rgb if rgb.ndim == 2 else cv2.cvtColor(rgb, cv2.COLOR_RGB2BGR)
    ok, buf = cv2.imencode(".jpg", src, list(params))
    assert ok
    return buf.tobytes()


def _cv2_decode(data):
    return cv2.cvtColor(cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)


@pytest.mark.parametrize("sampling", ["420", "422", "444"])
@pytest.mark.parametrize("hw", [(48, 64), (45, 77), (135, 241), (8, 3), (1, 1), (17, 5), (270, 480), (33, 47)])
def test_device_reconstruction_from_host_coefficients_equals_cv2(sampling, hw):
    """de-quantisation, jpeg_idct_islow, fancy up-sampling (edge columns / rows, tiny planes) and colour conversion on
    the device, entropy decoding still on the host: pixels identical to cv2.imread's"""
    from vdesens_b200 import jpeg
    for kind, q in (("smooth", 95), ("noise", 60)):
        data = _file(_image(hw[0], hw[1], kind, hw[0] + 3 * hw[1]), [S, SF[sampling], cv2.IMWRITE_JPEG_QUALITY, q])
        got = jpeg.reconstruct_from_coefficients(data, DEV)
        want = _cv2_decode(data)
        assert np.array_equal(got, want), f"{sampling} {hw} {kind}: {(got != want).sum()} bytes differ"
    grey = _file(_image(hw[0], hw[1], "smooth", 7)[..., 0].copy())
    assert np.array_equal(jpeg.reconstruct_from_coefficients(grey, DEV), _cv2_decode(grey))


@pytest.mark.parametrize("sub_bytes", [512, 1024, 4096])
@pytest.mark.parametrize("hw,params", [((135, 241), ()), ((270, 480), (cv2.IMWRITE_JPEG_OPTIMIZE, 1)), ((48, 64), (cv2.IMWRITE_JPEG_QUALITY, 30)),
                                       ((45, 77), (S, SF["444"])), ((50, 70), (S, SF["422"])), ((8, 3), ()), ((540, 960), (cv2.IMWRITE_JPEG_QUALITY, 100))])
def test_device_decoder_equals_cv2(hw, params, sub_bytes):
    """the whole device path, parallel Huffman decoding included (custom 'optimised' tables like ffmpeg's mjpeg encoder
    writes, several subsequence sizes): pixels identical to cv2.imread's, no file falls back to the CPU.
    The one exception is by construction: white noise at quality 100 has 63 non-zero AC coefficients in almost EVERY
    block, so hardly any block ends with an EOB code and a decoder that starts with a wrong zig-zag index keeps that
    offset for tens of kilobytes: with 512-byte subsequences the 32 rounds are not enough (they are with 1 KB).  The
    decoder must notice (result flag) and the file takes the CPU path; the pixels are right either way."""
    from vdesens_b200.jpeg import JpegDecoder
    files = [_file(_image(hw[0], hw[1], kind, 11 * i + hw[1]), params) for i, kind in enumerate(("smooth", "noise", "smooth"))]
    dec = JpegDecoder(3, device=DEV, max_file_bytes=max(len(f) for f in files) + 1024, sub_bytes=sub_bytes)
    frames, on_cpu = dec.decode(files)
    never_syncs = cv2.IMWRITE_JPEG_QUALITY in params and params[-1] == 100
    assert on_cpu == [] or (never_syncs and on_cpu == [1]), (on_cpu, [hex(v) for v in dec.result_host.tolist()])
    for i, f in enumerate(files):
        got, want = frames[i].cpu().numpy(), _cv2_decode(f)
        assert np.array_equal(got, want), f"file {i}: {(got != want).sum()} bytes differ"
    # a short batch on the same decoder, and a grey file (its own geometry)
    frames, on_cpu = dec.decode(files[:1])
    assert on_cpu == [] and np.array_equal(frames[0].cpu().numpy(), _cv2_decode(files[0]))
    print("sync rounds used (first round in which no exit state changed):", dec.sync_rounds())


def test_device_decoder_full_hd_and_fallbacks():
    """1080p files (the frame loop's size; ~0.9 MB of entropy data each -> ~1800 decoder threads per file); a progressive
    file and a file with restart markers in the same batch take the CPU fallback and still come out right"""
    from vdesens_b200.jpeg import JpegDecoder
    H, W = 1080, 1920
    imgs = [synth.synthetic_frame(H, W, 3000 + i) for i in range(4)]
    files = [_file(imgs[0]), _file(imgs[1], (cv2.IMWRITE_JPEG_OPTIMIZE, 1)), _file(imgs[2], (cv2.IMWRITE_JPEG_PROGRESSIVE, 1)),
             _file(imgs[3], (cv2.IMWRITE_JPEG_RST_INTERVAL, 7))]
    dec = JpegDecoder(4, device=DEV)
    frames, on_cpu = dec.decode(files)
    assert on_cpu == [2, 3]
    for i, f in enumerate(files):
        assert np.array_equal(frames[i].cpu().numpy(), _cv2_decode(f)), i
    assert dec.stats == {"files": 4, "device": 2, "fallback": 2}


def test_decode_detect_encode_round_trip_matches_the_cv2_loop(tmp_path):
    """files in -> device decode -> (no boxes) -> device encode -> files out == cv2.imread -> cv2.imwrite on the host"""
    from vdesens_b200.jpeg import JpegDecoder
    H, W = 270, 480
    files = [_file(synth.synthetic_frame(H, W, 60 + i)) for i in range(2)]
    dec, enc = JpegDecoder(2, device=DEV), JpegEncoder(2, H, W, device=DEV)
    frames, _ = dec.decode(files)
    out = enc.encode(frames)
    for f, o in zip(files, out):
        assert o == _cv2(_cv2_decode(f))


def test_directory_loop_with_the_codec_on_the_device_equals_the_host_loop(tmp_path):
    """combine_detect.batch_process_images on a directory of JPEG frames (what the reference's ffmpeg stage leaves behind,
    combine_detect.py:279-476): io="device" (decode + encode on the GPU) writes the SAME files, byte for byte, as
    io="host" (cv2.imread / cv2.imwrite), and returns the same counts; plates come from a host provider"""
    import os
    from detect_face.face import Retinaface
    import combine_detect
    H, W = 160, 224
    sd = synth.synthetic_state_dict(0, face_rate="r01")
    det = Retinaface(state_dict=sd, input_shape=[H, W, 3], confidence=0.5, nms_iou=0.4, letterbox_image=True, cuda=True,
                     batch_size=4)
    src = tmp_path / "in"
    src.mkdir()
    for i in range(10):                                    # batches of 4: 4 + 4 + 2
        cv2.imwrite(str(src / f"frame_{i:06d}.jpg"), cv2.cvtColor(synth.synthetic_frame(H, W, 4200 + i), cv2.COLOR_RGB2BGR))

    def plates(batch, verbose=False, conf=0.5):
        assert all(isinstance(b, np.ndarray) and b.shape == (H, W, 3) for b in batch)
        return [(None, [[10 + k, 20, 90, 70]]) for k, _ in enumerate(batch)]

    outs = {}
    for io in ("host", "device"):
        dst = tmp_path / io
        outs[io] = combine_detect.batch_process_images(str(src), str(dst), det, plates, batch_size=4, mode="gauss", ksize=15, io=io)
        assert sorted(os.listdir(dst)) == [f"processed_frame_{i:06d}.jpg" for i in range(10)]
    assert outs["host"] == outs["device"] and outs["host"][0] == 10 and outs["host"][2] == 10 and outs["host"][1] > 0
    for i in range(10):
        a = (tmp_path / "host" / f"processed_frame_{i:06d}.jpg").read_bytes()
        b = (tmp_path / "device" / f"processed_frame_{i:06d}.jpg").read_bytes()
        assert a == b, f"frame {i}: files differ ({len(a)} vs {len(b)} bytes)"
    # "auto" picks the device codec for such a directory; a PNG directory keeps the host codec
    assert combine_detect.batch_process_images(str(src), str(tmp_path / "auto"), det, None, batch_size=4, mode="gauss", ksize=15) == \
        combine_detect.batch_process_images(str(src), str(tmp_path / "host2"), det, None, batch_size=4, mode="gauss", ksize=15, io="host")


def test_device_decoder_survives_corrupt_entropy_data():
    """random damage inside the entropy-coded segment (bit flips, zeroed runs, a truncated tail): the decoder kernels
    must neither fault nor hang — they either flag the file (CPU fallback) or produce SOME frame of the right shape — and
    the decoder object keeps working for the next, intact batch"""
    from vdesens_b200.jpeg import JpegDecoder, prepare_device
    rng = np.random.default_rng(99)
    img = _image(135, 241, "smooth", 1)
    good = _file(img)
    h = prepare_device(good)
    s0, s1 = h.scan_offset, h.scan_offset + h.scan_bytes
    dec = JpegDecoder(4, device=DEV, max_file_bytes=len(good) + 4096,
                      fallback=lambda data: np.zeros((135, 241, 3), np.uint8))         # never hand corrupt data to cv2 here
    for trial in range(6):
        files = []
        for i in range(4):
            b = bytearray(good)
            if trial % 3 == 0:
                for _ in range(1 + 5 * i):
                    b[int(rng.integers(s0, s1))] ^= 1 << int(rng.integers(0, 8))
            elif trial % 3 == 1:
                a = int(rng.integers(s0, s1 - 64))
                b[a:a + 16 * (i + 1)] = bytes(16 * (i + 1))
            else:
                del b[s1 - 50 * (i + 1):s1]
            files.append(bytes(b))
        frames, on_cpu = dec.decode(files)
        torch.cuda.synchronize()
        assert tuple(frames.shape) == (4, 135, 241, 3)
    frames, on_cpu = dec.decode([good] * 4)
    assert on_cpu == [] and np.array_equal(frames[0].cpu().numpy(), _cv2_decode(good))
