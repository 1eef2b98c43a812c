"""RetinaFace-MobileNet0.25 (`Retinaface(backbone="mobilenet")`) described as data.

Follows the reference topology
  * backbone MobileNetV1 x0.25 through IntermediateLayerGetter stage1/2/3
                                           -- reference detect_face/nets/mobilenet025.py:3-49, retinaface.py:60-73
  * FPN with 64 channels, LeakyReLU(0.1) because out_channel <= 64  -- detect_face/nets/layers.py:68-114 (:71)
  * SSH(64, 64) x3, LeakyReLU(0.1) on the two inner conv_bn        -- layers.py:37-66 (:41)
  * Class/Bbox/Landmark heads on 64 channels                      -- retinaface.py:13-51
Every BN is eval-mode and folded at pack time.

How the ops map onto kernels (vdesens_b200/engine.py executes this list):
  "stem"  3x3/2 3->8 on the ingest layout              -> vd_mnet_stem         (CUDA cores, HBM-bound)
  "dw"    depthwise 3x3 (+BN+leaky)                    -> vd_mnet_dw3x3        (CUDA cores, HBM-bound)
  "pw"    pointwise 1x1 with Cin < 64                  -> vd_mnet_pw           (CUDA cores, HBM-bound)
  "dwpw"  a whole conv_dw block with Cin <= 32         -> vd_mnet_dwpw         (dw result stays in shared memory;
          build_graph(fuse_dwpw=True), what the engine runs; bit-identical to "dw" + "pw")
  "conv"  everything with Cin % 64 == 0                -> conv_igemm_kernel    (tcgen05), activation by a per-channel
          negative-side slope vector (VD_CONV_SLOPE: 0 = ReLU, 0.1 = LeakyReLU, 1 = identity)
  "heads" fused N = 32 1x1 GEMM over TWO inputs (VD_CONV_DUAL)

SSH on 64-channel-granular tensor-core tiles.  The reference concatenates 32 + 16 + 16 channels produced by five
3x3 convolutions with 16/32-channel tensors in between; the implicit-GEMM kernel reads and writes 64-channel lines,
so the module is re-expressed as THREE 64->64 3x3 convolutions over zero-padded weights (extra tensor-core work on
zeros is free next to the HBM traffic it saves):
  S1 = conv(x):   out[ 0:32] = conv3X3 (ReLU: the concat's F.relu)   out[32:48] = conv5X5_1 (leaky)   out[48:64] = 0
  S2 = conv(S1):  reads in[32:48];  out[0:16] = conv5X5_2 (ReLU)     out[16:32] = conv7X7_2 (leaky)   rest 0
  S3 = conv(S2):  out[0:16] = in[0:16] through an identity centre tap (exact: 1.0 * v in fp32, ReLU of a ReLU)
                  out[16:32] = conv7x7_3 of in[16:32] (ReLU)                                          rest 0
  heads = 1x1 over [S1 | S3] with the head weights scattered to the live channels (K = 128).
"""
from dataclasses import dataclass
from typing import List

from .arch import BufSpec, feature_hw, num_anchors, level_offsets, HEAD_ROWS, BN_EPS  # noqa: F401 (re-exported)

# cfg_mnet of the reference (detect_face/utils/config.py:1-16), only the keys the path reads
CFG_MNET = {
    "name": "mobilenet0.25",
    "min_sizes": [[16, 32], [64, 128], [256, 512]],
    "steps": [8, 16, 32],
    "variance": [0.1, 0.2],
    "clip": False,
    "return_layers": {"stage1": 1, "stage2": 2, "stage3": 3},
    "in_channel": 32,
    "out_channel": 64,
}
LEAKY = 0.1

# (stage, index inside the stage's nn.Sequential, cin, cout, stride) of every conv_dw block, mobilenet025.py:24-49
DW_BLOCKS = [
    ("stage1", 1, 8, 16, 1), ("stage1", 2, 16, 32, 2), ("stage1", 3, 32, 32, 1), ("stage1", 4, 32, 64, 2),
    ("stage1", 5, 64, 64, 1),
    ("stage2", 0, 64, 128, 2), ("stage2", 1, 128, 128, 1), ("stage2", 2, 128, 128, 1), ("stage2", 3, 128, 128, 1),
    ("stage2", 4, 128, 128, 1), ("stage2", 5, 128, 128, 1),
    ("stage3", 0, 128, 256, 2), ("stage3", 1, 256, 256, 1),
]


@dataclass
class MOp:
    kind: str                 # "stem" | "dw" | "pw" | "dwpw" | "conv" | "heads"
    name: str                 # key into weights.pack_mnet()
    src: str
    dst: str
    cin: int = 0
    cout: int = 0
    k: int = 1
    stride: int = 1
    pad: int = 0
    src2: str = ""            # heads: second input
    upsample_add: str = ""    # FPN top-down add (after the activation, layers.py:102-111)
    level: int = -1


def state_dict_spec():
    """[(key, shape, kind)] of the reference state_dict for cfg_mnet (300 entries)."""
    spec = []

    def conv(key, shape, kind="conv_neck"):
        spec.append((key + ".weight", shape, kind))

    def bn(prefix, c):
        spec.append((prefix + ".weight", (c,), "bn_weight"))
        spec.append((prefix + ".bias", (c,), "bn_bias"))
        spec.append((prefix + ".running_mean", (c,), "bn_mean"))
        spec.append((prefix + ".running_var", (c,), "bn_var"))
        spec.append((prefix + ".num_batches_tracked", (), "bn_count"))

    conv("body.stage1.0.0", (8, 3, 3, 3))
    bn("body.stage1.0.1", 8)
    for st, i, cin, cout, _ in DW_BLOCKS:
        p = f"body.{st}.{i}"
        conv(p + ".0", (cin, 1, 3, 3), "conv_dw")
        bn(p + ".1", cin)
        conv(p + ".3", (cout, cin, 1, 1))
        bn(p + ".4", cout)
    for name, cin, k in (("output1", 64, 1), ("output2", 128, 1), ("output3", 256, 1), ("merge1", 64, 3), ("merge2", 64, 3)):
        conv(f"fpn.{name}.0", (64, cin, k, k))
        bn(f"fpn.{name}.1", 64)
    for s in ("ssh1", "ssh2", "ssh3"):
        for name, cin, cout in (("conv3X3", 64, 32), ("conv5X5_1", 64, 16), ("conv5X5_2", 16, 16),
                                ("conv7X7_2", 16, 16), ("conv7x7_3", 16, 16)):
            conv(f"{s}.{name}.0", (cout, cin, 3, 3))
            bn(f"{s}.{name}.1", cout)
    for head, n in (("ClassHead", 4), ("BboxHead", 8), ("LandmarkHead", 20)):
        for lvl in range(3):
            spec.append((f"{head}.{lvl}.conv1x1.weight", (n, 64, 1, 1), "head_weight"))
            spec.append((f"{head}.{lvl}.conv1x1.bias", (n,), "head_bias"))
    return spec


DWPW_PAIRS = ((8, 16), (16, 32), (32, 32), (32, 64))     # (cin, cout) instantiated by vd_mnet_dwpw


def build_graph(fuse_dwpw: bool = False):
    """Returns (ops, bufs).  Buffer 'x0' is the ingested frame.  fuse_dwpw: conv_dw blocks with (cin, cout) in
    DWPW_PAIRS become one "dwpw" op (weights <block>.dw and <block>.pw) and their depthwise buffer disappears."""
    ops: List[MOp] = []
    bufs: List[BufSpec] = [BufSpec("x0", 3, 1)]

    def buf(name, ch, stride):
        bufs.append(BufSpec(name, ch, stride))
        return name

    cur, cur_stride = buf("m.stem", 8, 2), 2
    ops.append(MOp("stem", "body.stage1.0", "x0", cur, 3, 8, 3, 2, 1))
    feats = {}
    for st, i, cin, cout, s in DW_BLOCKS:
        p = f"body.{st}.{i}"
        if fuse_dwpw and (cin, cout) in DWPW_PAIRS:
            cur_stride *= s
            o = buf(p + ".pw", cout, cur_stride)
            ops.append(MOp("dwpw", p, cur, o, cin, cout, 3, s, 1))
            cur = o
            feats[st] = cur
            continue
        d = buf(p + ".dw", cin, cur_stride * s)
        ops.append(MOp("dw", p + ".dw", cur, d, cin, cin, 3, s, 1))
        cur_stride *= s
        o = buf(p + ".pw", cout, cur_stride)
        ops.append(MOp("conv" if cin % 64 == 0 else "pw", p + ".pw", d, o, cin, cout, 1, 1, 0))
        cur = o
        feats[st] = cur
    c3, c4, c5 = feats["stage1"], feats["stage2"], feats["stage3"]
    buf("fpn.o3", 64, 32)
    ops.append(MOp("conv", "fpn.output3", c5, "fpn.o3", 256, 64, 1, 1, 0))
    buf("fpn.o2s", 64, 16)
    ops.append(MOp("conv", "fpn.output2", c4, "fpn.o2s", 128, 64, 1, 1, 0, upsample_add="fpn.o3"))
    buf("fpn.o2", 64, 16)
    ops.append(MOp("conv", "fpn.merge2", "fpn.o2s", "fpn.o2", 64, 64, 3, 1, 1))
    buf("fpn.o1s", 64, 8)
    ops.append(MOp("conv", "fpn.output1", c3, "fpn.o1s", 64, 64, 1, 1, 0, upsample_add="fpn.o2"))
    buf("fpn.o1", 64, 8)
    ops.append(MOp("conv", "fpn.merge1", "fpn.o1s", "fpn.o1", 64, 64, 3, 1, 1))
    for lvl, (src, stride) in enumerate([("fpn.o1", 8), ("fpn.o2", 16), ("fpn.o3", 32)]):
        s = f"ssh{lvl + 1}"
        for j in (1, 2, 3):
            buf(f"{s}.S{j}", 64, stride)
        ops.append(MOp("conv", f"{s}.S1", src, f"{s}.S1", 64, 64, 3, 1, 1))
        ops.append(MOp("conv", f"{s}.S2", f"{s}.S1", f"{s}.S2", 64, 64, 3, 1, 1))
        ops.append(MOp("conv", f"{s}.S3", f"{s}.S2", f"{s}.S3", 64, 64, 3, 1, 1))
        ops.append(MOp("heads", f"heads.{lvl}", f"{s}.S1", "heads", 64, HEAD_ROWS, 1, 1, 0, src2=f"{s}.S3", level=lvl))
    return ops, bufs
