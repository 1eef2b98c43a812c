"""CPU tests of the host side: C-ABI library loads and exports every declared symbol (no compute calls),
header/binding consistency, graph/weights bookkeeping, sharding + ordered gather over gloo (world 2)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from vdesens_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "vdesens.h")).read()
    declared = set(re.findall(r"\b(vd_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"vd_frame", "vd_status"}
    lib = _lib.lib()                       # raises if the .so is missing: there is no fallback
    lib16 = _lib.lib("fp16")
    for name in sorted(declared):
        assert hasattr(lib, name), f"libvdesens.so does not export {name}"
        assert hasattr(lib16, name), f"libvdesens_fp16.so does not export {name}"
    assert lib.vd_act_is_fp16() == 0 and lib16.vd_act_is_fp16() == 1
    assert set(_lib.EXPORTED) == declared, (set(_lib.EXPORTED) ^ declared)
    assert lib.vd_abi_version() == 1
    assert b"sm_100a" in lib.vd_build_info()


def test_header_constants_match_the_binding():
    """every numeric #define of include/vdesens.h that the ctypes binding mirrors has the same value there"""
    from vdesens_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "vdesens.h")).read()
    defs = {m.group(1): int(m.group(2), 0) for m in re.finditer(r"#define\s+(VD_[A-Z0-9_]+)\s+(0x[0-9a-fA-F]+|\d+)u?\b", hdr)}
    mirrored = [n for n in dir(_lib) if n.startswith("VD_") and isinstance(getattr(_lib, n), int)]
    assert len(mirrored) >= 15
    for n in mirrored:
        assert n in defs, f"{n} is not defined in vdesens.h"
        assert defs[n] == getattr(_lib, n), (n, defs[n], getattr(_lib, n))
    flags = [v for n, v in defs.items() if n.startswith("VD_CONV_")]
    assert len(flags) == len(set(flags)) and all(v & (v - 1) == 0 for v in flags)      # distinct single bits


def test_host_only_entry_points():
    from vdesens_b200 import _lib
    from oracle import cv_oracle as co
    lib = _lib.lib()
    assert lib.vd_x0_pitch(1920) == 1936 and lib.vd_x0_pitch(7) == 16 and lib.vd_x0_pitch(640) % 16 == 0
    for k, s in [(1, 0), (3, 0), (7, 0), (9, 0), (31, 0), (99, 0), (15, 2.5), (255, 0)]:
        assert _lib.gaussian_taps_q8(k, s) == co.gaussian_taps_q8(k, s).tolist()
    assert lib.vd_gaussian_taps_q8(4, 0.0, (ctypes.c_uint16 * 256)()) != 0      # even ksize rejected
    assert "ksize" in _lib.last_error()
    assert lib.vd_nms_workspace_bytes(16, 4096) > 16 * 4096 * 64 * 8
    with pytest.raises(RuntimeError):
        _lib.check(lib.vd_conv_plan_create(None, None), "vd_conv_plan_create")


def test_graph_bookkeeping():
    from vdesens_b200 import arch, synth, weights
    convs, ops, bufs = arch.build_graph(fuse_downsample=False)
    assert len(convs) == 73 and sum(1 for o in ops if o.kind == "heads") == 3        # 73 + 9 head convs = 82
    convs, ops, bufs = arch.build_graph()                                             # 4 projections fused away
    assert len(convs) == 69 and sum(1 for c in convs if c.wkey2) == 4
    spec = arch.state_dict_spec()
    assert len(spec) == 456
    assert sum(int(np.prod(s)) for _, s, k in spec if k not in ("bn_mean", "bn_var", "bn_count")) == 27293600
    # every buffer is written before it is read
    written = {"x0"}
    for o in ops:
        assert o.src in written, o
        for extra in (o.residual, o.upsample_add, o.src2):
            assert not extra or extra in written
        written.add(o.dst)
    flops = 0
    for o in ops:
        if o.kind == "conv":
            h, w = arch.feature_hw(1080, 1920, next(b.stride for b in bufs if b.name == o.dst))
            flops += 2 * (o.conv.cin * o.conv.k ** 2 + o.conv.cin2) * o.conv.cout * h * w
    for lvl, s in enumerate((8, 16, 32)):
        h, w = arch.feature_hw(1080, 1920, s)
        flops += 2 * 256 * 32 * h * w
    assert abs(flops / 1e9 - 449.71) < 0.2
    sd = synth.synthetic_state_dict(0)
    packed = weights.pack_all(sd)
    assert packed["body.conv1"][0].shape == (64, 224) and packed["heads.1"][0].shape == (32, 256)
    w0 = packed["body.conv1"][0].float().view(64, 7, 8, 4)
    assert float(w0[:, :, 7].abs().max()) == 0 and float(w0[..., 3:].abs().max()) == 0
    # BN fold: folded conv == conv + BN on a random input (fp32 check of the algebra, not of the device)
    c = next(c for c in convs if c.name == "body.layer2.0.conv2")
    w, b = weights._fold(weights.to_numpy_sd(sd), c)
    x = torch.randn(1, c.cin, 9, 9)
    y = torch.nn.functional.conv2d(x, torch.from_numpy(sd[c.wkey]), None, c.stride, c.pad)
    y = torch.nn.functional.batch_norm(y, torch.from_numpy(sd[c.bnkey + ".running_mean"]),
                                       torch.from_numpy(sd[c.bnkey + ".running_var"]),
                                       torch.from_numpy(sd[c.bnkey + ".weight"]), torch.from_numpy(sd[c.bnkey + ".bias"]),
                                       False, 0.0, 1e-5)
    y2 = torch.nn.functional.conv2d(x, torch.from_numpy(w), torch.from_numpy(b), c.stride, c.pad)
    assert torch.allclose(y, y2, atol=1e-4, rtol=1e-4)


def test_anchor_table_matches_oracle():
    from detect_face.utils.anchors import Anchors
    from detect_face.utils.config import cfg_re50
    from oracle import retinaface_oracle as ro
    for h, w in [(96, 128), (135, 241), (640, 640)]:
        assert torch.equal(Anchors(cfg_re50, image_size=[h, w]).get_anchors(), ro.anchors(h, w))


def test_shard_ranges_cover_stream():
    from vdesens_b200 import shard
    for n, world, batch in [(8192, 8, 16), (100, 3, 16), (5, 4, 2), (0, 2, 4), (17, 2, 1)]:
        seen = []
        for r in range(world):
            s, e = shard.shard_range(n, r, world, batch)
            assert s <= e and (s % batch == 0 or s == n)
            seen += list(range(s, e))
        assert seen == list(range(n))


def test_no_cpu_fallback_without_device():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from detect_face.face import Retinaface
    with pytest.raises(RuntimeError):
        Retinaface(state_dict={}, input_shape=[64, 64, 3])
    import combine_detect
    with pytest.raises(RuntimeError):
        combine_detect.mosaic_rectangle_region_single(np.zeros((8, 8, 3), np.uint8), 0, 0, 4, 4)


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, os.path.join(sys.argv[1], "video-desensitization_b200"))
import torch.distributed as dist
from vdesens_b200 import shard
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = 37
s, e = shard.shard_range(n, rank, world, batch=4)
local = [{"frame": i, "boxes": [[i, i, i + 1, i + 1]] * (i % 3)} for i in range(s, e)]
out = shard.gather_ordered(local, s)
if rank == 0:
    assert [o["frame"] for o in out] == list(range(n)), out
    print("GATHER_OK", len(out))
else:
    assert out is None
dist.destroy_process_group()
'''


def test_ordered_gather_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(script), ROOT],
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "GATHER_OK 37" in r.stdout


def test_load_image_rgb_equals_the_references_two_steps(tmp_path):
    """combine_detect.load_image_rgb == cv2.cvtColor(cv2.imread(p), COLOR_BGR2RGB) (reference combine_detect.py:170-172)
    for JPEG, PNG, greyscale and alpha files"""
    import cv2
    import combine_detect as cd
    from vdesens_b200 import synth
    rgb = synth.synthetic_frame(135, 241, 9)
    bgr = cv2.cvtColor(rgb, cv2.COLOR_RGB2BGR)
    files = {"a.jpg": bgr, "b.png": bgr, "c.png": bgr[..., 0].copy(), "d.png": np.dstack([bgr, bgr[..., :1]]),
             "e.jpg": bgr[..., 1].copy()}
    for name, img in files.items():
        p = str(tmp_path / name)
        assert cv2.imwrite(p, img)
        want = cv2.cvtColor(cv2.imread(p), cv2.COLOR_BGR2RGB)
        got = cd.load_image_rgb(p)
        assert got.dtype == np.uint8 and got.shape == want.shape and np.array_equal(got, want), name
    with pytest.raises(ValueError):
        cd.load_image_rgb(str(tmp_path / "missing.jpg"))


def test_decode_ahead_yields_groups_in_order(tmp_path):
    """the frame loop's decode-ahead stage: every group comes back complete and in order, whatever the pool size"""
    import cv2
    import combine_detect as cd
    from vdesens_b200 import synth
    paths = []
    for i in range(7):
        p = str(tmp_path / f"f{i:02d}.png")
        cv2.imwrite(p, cv2.cvtColor(synth.synthetic_frame(24, 32, 100 + i), cv2.COLOR_RGB2BGR))
        paths.append(p)
    groups = [paths[0:3], paths[3:6], paths[6:7]]
    assert 6 <= cd._io_workers() <= 32
    for workers in (None, 1, 4):
        got = list(cd._decoded_batches(groups, workers=workers))
        assert [g[0] for g in got] == groups
        for files, frames in got:
            assert len(frames) == len(files)
            for f, fr in zip(files, frames):
                assert np.array_equal(fr, synth.synthetic_frame(24, 32, 100 + int(os.path.basename(f)[1:3])))
    assert list(cd._decoded_batches([])) == []


def test_plate_boxes_accepts_tuples_and_ultralytics_results():
    """reference combine_detect.py:239 honours only tuples — which drops every plate of the Ultralytics model the
    reference itself builds (:868-874).  Here a Results-like object (`.boxes.xyxy`) is read, and any other type warns."""
    import combine_detect as cd

    class _Boxes:
        def __init__(self, a):
            self.xyxy = torch.tensor(a, dtype=torch.float32)

    class _Results:
        def __init__(self, a):
            self.boxes = _Boxes(a)

    res = [(None, [[1, 2, 3, 4]]), _Results([[5.5, 6.5, 70.25, 80.0], [0, 0, 9, 9]]), _Results(np.zeros((0, 4))), None]
    assert cd._plate_boxes(res, 0) == [[1, 2, 3, 4]]
    assert cd._plate_boxes(res, 1) == [[5.5, 6.5, 70.25, 80.0], [0.0, 0.0, 9.0, 9.0]]
    assert cd._plate_boxes(res, 2) == [] and cd._plate_boxes(res, 3) == [] and cd._plate_boxes(None, 0) == []
    cd._warned_plate_type = False
    with pytest.warns(RuntimeWarning, match="NOT applied"):
        assert cd._plate_boxes([["not", "a", "tuple"]], 0) == []
    # the integer boxes the compositing pass receives: int() truncation toward zero (reference :243)
    assert cd._plates_of(lambda frames, verbose, conf: [_Results([[-5.5, 6.9, 70.25, 80.0]])], [None]) == [[(-5, 6, 70, 80)]]
    assert cd._plates_of(None, [None, None]) == [[], []]


def test_mode_and_switch_validation():
    from vdesens_b200 import _lib
    import combine_detect as cd
    assert _lib.mode_id("mosaic") == _lib.VD_MODE_MOSAIC and _lib.mode_id("gauss") == _lib.VD_MODE_GAUSS
    for bad in ("gaus", "Mosaic", "", None):
        with pytest.raises(ValueError):
            _lib.mode_id(bad)
    with pytest.raises(ValueError):
        cd.composite_frames([np.zeros((8, 8, 3), np.uint8)], [[(0, 0, 4, 4)]], mode="blur")
    with pytest.raises(ValueError):
        cd.batch_process_images("/nonexistent", "/nonexistent", None, None, on_error="ignore")


def test_engine_cache_reuses_the_larger_replica(monkeypatch):
    """every new list length used to build a full DetectorEngine replica.  A shorter batch now reuses the
    engine already built, a longer one replaces it, and at most _MAX_ENGINES are alive."""
    from detect_face import face as ff
    built = []

    class FakeEngine:
        def __init__(self, sd, b, *a, **k):
            self.B = b
            built.append(b)

    monkeypatch.setattr(ff, "DetectorEngine", FakeEngine)
    r = ff.Retinaface.__new__(ff.Retinaface)
    r.__dict__.update(ff.Retinaface._defaults)
    r.batch_size, r.cand_cap, r.max_det, r.precision, r._sd, r.device, r._engines = None, 4096, 2048, "bf16", {}, "cuda:0", {}
    monkeypatch.setattr(torch.cuda, "empty_cache", lambda: None)
    assert r._engine_for(16).B == 16
    assert r._engine_for(5).B == 16 and r._engine_for(16).B == 16 and r._engine_for(1).B == 16
    assert built == [16]
    assert r._engine_for(1, exact=True).B == 1 and built == [16, 1]
    assert r._engine_for(32).B == 32 and built == [16, 1, 32] and len(r._engines) == ff.Retinaface._MAX_ENGINES
    assert sorted(r._engines) == [1, 32]                 # the least recently used (16) was freed
    assert r._engine_for(7).B == 32 and built == [16, 1, 32]
    r.batch_size = 8                                     # a configured batch wins over the list length
    assert r._engine_for(3).B == 32                      # still served by the larger replica
    r._engines.clear()
    assert r._engine_for(3).B == 8
