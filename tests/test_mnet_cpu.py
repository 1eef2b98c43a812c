"""CPU: the MobileNet0.25 variant — oracle restatement against the committed outputs of the
unmodified reference run with backbone='mobilenet', and the device op list + weight packing (padded SSH, identity
pass-through, two-input head GEMM) emulated on the CPU against that oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import retinaface_oracle as ro
from vdesens_b200 import arch_mnet as am, synth, weights


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_detect_mnet_golden.npz"))


@pytest.fixture(scope="module")
def sd_np():
    return synth.synthetic_state_dict(0, face_rate="r01", backbone="mobilenet")


def _frames(g, name):
    return [synth.synthetic_frame(int(h), int(w), int(s)) for h, w, s in g[name + ".frames"]]


def test_state_dict_spec_is_the_reference_layout(sd_np):
    spec = am.state_dict_spec()
    assert len(spec) == 300 and len({k for k, _, _ in spec}) == 300          # SURVEY: 300 entries for cfg_mnet
    assert sum(int(np.prod(s)) for _, s, kind in spec if kind != "bn_count") == 433343 - 47
    assert set(sd_np) == {k for k, _, _ in spec}


@pytest.mark.needs_reference
def test_state_dict_keys_match_the_reference_model():
    from oracle import ref_shims
    rf, _ = ref_shims.load_reference()
    cfg = ref_shims._loaded["mods"]["detect_face.utils.config"].cfg_mnet
    ref_sd = rf.RetinaFace(cfg=cfg, mode="eval", pretrained=False).state_dict()
    mine = {k: tuple(s) for k, s, _ in am.state_dict_spec()}
    assert {k: tuple(v.shape) for k, v in ref_sd.items()} == mine


@pytest.mark.parametrize("case", ["native", "letterbox", "odd", "vga"])
def test_oracle_heads_and_rows_match_reference(gold, sd_np, case):
    sd = ro.to_torch_sd(sd_np)
    frames = _frames(gold, case)
    H, W = (int(v) for v in gold[case + ".input_shape"][:2])
    x, shapes = ro.preprocess(frames, (H, W))
    loc, conf, landms = ro.forward_mnet(sd, x)
    np.testing.assert_allclose(loc.numpy(), gold[case + ".loc"], atol=2e-4, rtol=0)
    np.testing.assert_allclose(conf.numpy(), gold[case + ".conf"], atol=2e-4, rtol=0)
    np.testing.assert_allclose(landms.numpy(), gold[case + ".landms"], atol=2e-4, rtol=0)
    # post-processing on the reference's own head tensors -> the reference's rows
    rows = ro.postprocess(torch.from_numpy(gold[case + ".loc"]), torch.from_numpy(gold[case + ".conf"]),
                          torch.from_numpy(gold[case + ".landms"]), shapes, (H, W), 0.5, 0.4)
    for i, r in enumerate(rows):
        want = gold[f"{case}.rows{i}"]
        assert r.shape[0] == want.shape[0]
        if len(want):
            np.testing.assert_allclose(r.numpy(), want, atol=1e-6, rtol=0)


@pytest.mark.parametrize("case", ["letterbox", "odd"])
def test_device_op_list_emulated_tracks_the_fp32_oracle(gold, sd_np, case):
    """graph + packing check without a GPU: fp32 'storage' makes the emulation exact up to BN folding, bf16 storage
    stays within the stated reduced-precision bounds."""
    sd = ro.to_torch_sd(sd_np)
    frames = _frames(gold, case)
    H, W = (int(v) for v in gold[case + ".input_shape"][:2])
    x, _ = ro.preprocess(frames, (H, W))
    cap = {}
    loc, conf, ldm = ro.forward_mnet(sd, x, capture=cap, softmax=False)
    ecap = {}
    eloc, ecls, eldm = ro.forward_mnet_emulated(weights.pack_mnet(sd_np, dtype=torch.float32), x, capture=ecap,
                                                dtype=torch.float32)
    assert float((eloc - loc).abs().max()) < 2e-4 and float((ecls - conf).abs().max()) < 2e-4
    assert float((eldm - ldm).abs().max()) < 2e-4
    # SSH re-expression: S1[0:32] | S3[0:32] is the reference's relu(cat(conv3X3, conv5X5, conv7X7))
    for lvl in (1, 2, 3):
        got = torch.cat([ecap[f"ssh{lvl}.S1"][:, :32], ecap[f"ssh{lvl}.S3"][:, :32]], dim=1)
        assert float((got - cap[f"ssh{lvl}.out"]).abs().max()) < 1e-4
    # reduced-precision storage on this UNTRAINED network (13 depthwise + BN stages, each re-normalising): stated
    # bounds bf16 |err| <= 0.35 (rms ~0.04 of values with rms ~0.5), fp16 |err| <= 0.06
    bloc, bcls, _ = ro.forward_mnet_emulated(weights.pack_mnet(sd_np), x)
    assert float((bloc - loc).abs().max()) < 0.35 and float((bcls - conf).abs().max()) < 0.35
    hloc, hcls, _ = ro.forward_mnet_emulated(weights.pack_mnet(sd_np, torch.float16), x, dtype=torch.float16)
    assert float((hloc - loc).abs().max()) < 0.06 and float((hcls - conf).abs().max()) < 0.06
