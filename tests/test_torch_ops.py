"""`torch.ops.vdesens.*` — the C-ABI exposed as torch custom ops (north_star).  Registration, schemas, fake
(shape) kernels and the absence of a CPU kernel are checked here without a GPU; the GPU tests call the hot path
end to end THROUGH the dispatcher and compare with the oracle / golden fixtures like tests/test_gpu_kernels.py."""
import os

import numpy as np
import pytest
import torch

from oracle import cv_oracle as co
from oracle import retinaface_oracle as ro
from vdesens_b200 import synth

DEV = "cuda:0"


@pytest.fixture(scope="module")
def T():
    import vdesens_b200.torch_ops as t
    return t


def test_ops_are_registered_with_cuda_only_kernels(T):
    assert T.registered() == T.OP_NAMES
    ns = torch.ops.vdesens
    assert "Tensor(a0!)[] frames" in str(ns.composite_.default._schema)       # in-place on the frames, returns nothing
    assert str(ns.composite_.default._schema).endswith("-> ()")
    for name in T.OP_NAMES:
        op = getattr(ns, name).default
        assert torch._C._dispatch_has_kernel_for_dispatch_key(op.name(), "CUDA"), name
        assert not torch._C._dispatch_has_kernel_for_dispatch_key(op.name(), "CPU"), name


def test_cpu_tensors_are_refused_not_computed(T):
    with pytest.raises(NotImplementedError):
        torch.ops.vdesens.maxpool3x3s2(torch.zeros(1, 4, 4, 64, dtype=torch.bfloat16))
    with pytest.raises(NotImplementedError):
        torch.ops.vdesens.ingest([torch.zeros(8, 8, 3, dtype=torch.uint8)], 8, 8)


def test_fake_kernels_give_the_output_shapes(T):
    m = dict(device="meta")
    x = torch.empty(2, 135, 241, 64, dtype=torch.bfloat16, **m)
    w = torch.empty(128, 9 * 64, dtype=torch.bfloat16, **m)
    b = torch.empty(128, **m)
    assert torch.ops.vdesens.conv2d(x, w, b, 3, 2, 1, True, None).shape == (2, 68, 121, 128)
    assert torch.ops.vdesens.maxpool3x3s2(x).shape == (2, 68, 121, 64)
    fr = [torch.empty(270, 480, 3, dtype=torch.uint8, **m)] * 3
    assert torch.ops.vdesens.ingest_nchw_f32(fr, 320, 320).shape == (3, 3, 320, 320)
    A = 2 * (40 * 40 + 20 * 20 + 10 * 10)
    c, n, s = torch.ops.vdesens.decode_threshold(torch.empty(3, A, 4, **m), torch.empty(3, A, 2, **m),
                                                 torch.empty(3, A, 10, **m), 320, 320, 0.5, 512)
    assert c.shape == (3, 512, 16) and n.shape == (3,) and s.shape == (1,) and n.dtype == torch.int32
    r, oc, s2 = torch.ops.vdesens.nms(c, n, 0.4, torch.empty(3, 4, **m), 64)
    assert r.shape == (3, 64, 16) and oc.shape == (3,)


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_ingest_op_exact(T):
    frames = [synth.synthetic_frame(150, 200, 7), synth.synthetic_frame(260, 180, 8)]
    want, _ = ro.preprocess(frames, (128, 128))
    dfr = [torch.from_numpy(f).to(DEV) for f in frames]
    got = torch.ops.vdesens.ingest_nchw_f32(dfr, 128, 128)
    assert torch.equal(got.cpu(), want)
    x0 = torch.ops.vdesens.ingest(dfr, 128, 128)
    assert x0.dtype == torch.bfloat16
    assert torch.equal(x0.cpu().float()[:, :, 3:3 + 128, :3].permute(0, 3, 1, 2), want)


@pytest.mark.gpu
def test_conv_and_pool_ops_vs_torch(T):
    from vdesens_b200 import weights
    g = torch.Generator().manual_seed(5)
    B, H, W, Cin, Cout = 2, 33, 47, 64, 128
    x = (torch.randn(B, H, W, Cin, generator=g) * 0.5).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * 0.05).to(torch.bfloat16)
    bias = torch.randn(Cout, generator=g) * 0.1
    wp = w.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous()                  # K order (r, s, cin)
    y = torch.ops.vdesens.conv2d(x.to(DEV), wp.to(DEV), bias.to(DEV), 3, 2, 1, True, None)
    ref = torch.relu(torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, stride=2, padding=1))
    ref = ref.permute(0, 2, 3, 1)
    got = y.cpu().float()
    assert got.shape == ref.shape
    # bf16 output rounding of an fp32 accumulation: half an ulp of bf16 relative + accumulation-order slack
    assert float((got - ref).abs().max()) <= 2 ** -8 * float(ref.abs().max()) + 1e-3
    p = torch.ops.vdesens.maxpool3x3s2(y)
    pref = torch.nn.functional.max_pool2d(y.float().permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
    assert torch.equal(p.float().cpu(), pref.cpu())                            # exact
    assert weights is not None


@pytest.mark.gpu
def test_decode_nms_ops_same_kept_sets_as_reference(T, golden_dir):
    from vdesens_b200 import ops
    g = np.load(os.path.join(golden_dir, "ref_detect_golden.npz"))
    case = "letterbox"
    fr = g[case + ".frames"]
    H, W = (int(v) for v in g[case + ".input_shape"][:2])
    loc, conf, ldm = (torch.from_numpy(g[f"{case}.{k}"]).to(DEV) for k in ("loc", "conf", "landms"))
    cand, count, st = torch.ops.vdesens.decode_threshold(loc, conf, ldm, H, W, 0.5, 1024)
    unlb = torch.from_numpy(ops.unletterbox_params((H, W), [(int(h), int(w)) for h, w, _ in fr]))
    rows, oc, st2 = torch.ops.vdesens.nms(cand, count, 0.4, unlb, 256)
    assert int(st.item()) == 0 and int(st2.item()) == 0
    for i in range(len(fr)):
        want = g[f"{case}.rows{i}"]                 # rows of the executed reference (tests/golden/make script)
        n = int(oc[i])
        assert n == want.shape[0]
        np.testing.assert_allclose(rows[i, :n, :15].cpu().numpy(), want, atol=1e-5, rtol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("mode,ksize", [(0, 1), (1, 31)])
def test_composite_op_in_place_bit_exact(T, mode, ksize):
    frames = [synth.synthetic_frame(270, 480, 900 + i) for i in range(2)]
    boxes = [synth.synthetic_boxes(12, 270, 480, seed=70 + i, min_size=1, max_size=150) for i in range(2)]
    boxes[0] = boxes[0][:5]                                                      # ragged counts
    want = [co.composite(f, b, "mosaic" if mode == 0 else "gauss", 8, ksize) for f, b in zip(frames, boxes)]
    E = max(len(b) for b in boxes)
    eb = np.zeros((2, E, 4), np.int32)
    for i, bl in enumerate(boxes):
        eb[i, :len(bl)] = np.asarray(bl, np.int32)
    ec = np.array([len(b) for b in boxes], np.int32)
    dfr = [torch.from_numpy(f.copy()).to(DEV) for f in frames]
    out = torch.ops.vdesens.composite_(dfr, torch.from_numpy(eb), torch.from_numpy(ec), mode, 8, ksize, 0.0, None, None)
    assert out is None
    for i in range(2):
        assert np.array_equal(dfr[i].cpu().numpy(), want[i])
