"""RetinaFace-ResNet50 described as data: every convolution of the hot path with the state_dict
keys it consumes, and the op list the device engine executes.

Follows the reference network topology
  * backbone: torchvision resnet50 (v1.5: stride on the 3x3) through IntermediateLayerGetter
    layer2/3/4                         -- reference detect_face/retinaface.py:71-78,121
  * FPN (3 lateral 1x1, 2 top-down merges) -- reference detect_face/nets/layers.py:68-114
  * SSH x3                             -- reference detect_face/nets/layers.py:37-66
  * Class/Bbox/Landmark heads, 2 anchors per cell -- reference detect_face/retinaface.py:13-51,140-147
Every BN is eval-mode (folded at pack time); every activation is plain ReLU for the ResNet-50
config (leaky == 0 because out_channel 256 > 64, layers.py:41,71).
"""
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

# cfg_re50 of the reference (detect_face/utils/config.py:18-29), only the keys the path reads
CFG_RE50 = {
    "name": "Resnet50",
    "min_sizes": [[16, 32], [64, 128], [256, 512]],
    "steps": [8, 16, 32],
    "variance": [0.1, 0.2],
    "clip": False,
    "return_layers": {"layer2": 1, "layer3": 2, "layer4": 3},
    "in_channel": 256,
    "out_channel": 256,
}

BN_EPS = 1e-5


@dataclass
class ConvSpec:
    name: str                 # unique layer name
    wkey: str                 # state_dict key of the conv weight (Cout,Cin,k,k)
    bnkey: Optional[str]      # prefix of the BN that follows (".weight/.bias/.running_mean/.running_var")
    biaskey: Optional[str]    # conv bias key (heads only)
    cin: int
    cout: int
    k: int
    stride: int
    pad: int
    relu: bool                # ReLU applied in the epilogue (after residual add when there is one)
    # fused second K segment (bottleneck block 0: conv3 + downsample projection share one accumulator)
    wkey2: Optional[str] = None
    bnkey2: Optional[str] = None
    cin2: int = 0
    stride2: int = 1


@dataclass
class Op:
    kind: str                 # "conv" | "maxpool" | "heads"
    conv: Optional[ConvSpec] = None
    src: str = ""             # input buffer name
    dst: str = ""             # output buffer name
    dst_coff: int = 0         # channel offset inside dst (SSH concat slices)
    src2: str = ""            # input of the fused 1x1 projection (conv.wkey2), read with conv.stride2
    residual: str = ""        # buffer added before ReLU (bottleneck identity), same shape as dst
    upsample_add: str = ""    # buffer (half resolution) nearest-upsampled and added AFTER ReLU (FPN top-down)
    pool: bool = False        # conv1 only: the 3x3/2 max-pool runs inside the convolution kernel, dst is the pooled tensor
    level: int = -1           # heads: pyramid level 0..2
    sibling: Optional[ConvSpec] = None # VD_CONV_TWO_DST: a second layer reading the same tensor with the same geometry, run
    sibling_dst: str = ""              # as extra output channels of this GEMM; its own destination buffer / slice
    sibling_coff: int = 0
    chain: Optional[ConvSpec] = None   # VD_CONV_CHAIN: the NEXT 1x1 convolution, computed by the same kernel from the
    chain_dst: str = ""                # output tile while it is still in shared memory; chain_dst is its output buffer


@dataclass
class BufSpec:
    name: str
    channels: int
    stride: int               # spatial reduction factor w.r.t. the network input (2,4,8,16,32)


def _cdiv(a, b):
    return -(-a // b)


def feature_hw(h: int, w: int, stride: int) -> Tuple[int, int]:
    """Every strided op in the stack yields ceil(n/2), hence ceil(n/stride) overall;
    the reference's Anchors uses the same ceil (detect_face/utils/anchors.py:20)."""
    hh, ww = h, w
    s = stride
    while s > 1:
        hh, ww = _cdiv(hh, 2), _cdiv(ww, 2)
        s //= 2
    return hh, ww


def num_anchors(h: int, w: int) -> int:
    return sum(2 * feature_hw(h, w, s)[0] * feature_hw(h, w, s)[1] for s in CFG_RE50["steps"])


def level_offsets(h: int, w: int) -> List[int]:
    offs, a = [], 0
    for s in CFG_RE50["steps"]:
        offs.append(a)
        fh, fw = feature_hw(h, w, s)
        a += 2 * fh * fw
    return offs


def build_graph(fuse_downsample: bool = True, fuse_stem_pool: bool = False, fuse_chain: bool = False,
                fuse_siblings: bool = False):
    """Returns (convs, ops, bufs).  Buffer 'x0' is the ingested frame (3 real channels).
    With fuse_stem_pool conv1+bn1+relu+maxpool (retinaface.py:71-73) is one op writing 'p1' directly; the
    half-resolution 64-channel tensor 'c1' (4x the size of either neighbour) is never materialised.
    With fuse_downsample the projection shortcut of each stage's first bottleneck (torchvision
    Bottleneck.downsample) is not a separate convolution: its 1x1 GEMM accumulates into conv3's
    accumulator (same output pixels), which removes one write and one read of the widest tensors.
    With fuse_chain a 256-channel conv3 also computes the 1x1 convolution that follows it (the next bottleneck's conv1,
    64 or 128 outputs): layer1's three block outputs are written once and read back only as residuals.
    With fuse_siblings two consecutive 3x3 convolutions of an SSH module that read the SAME tensor (conv3X3 + conv5X5_1
    on the FPN output, conv5X5_2 + conv7X7_2 on t5; reference nets/layers.py:54-62) run as one GEMM over the
    concatenated filters (N = 192 / 128 instead of 128 + 64 / 64 + 64): the input is fetched once and the narrow
    N = 64 GEMMs, which cannot feed the tensor pipe, disappear."""
    convs: List[ConvSpec] = []
    ops: List[Op] = []
    bufs: List[BufSpec] = []

    def buf(name, ch, stride):
        bufs.append(BufSpec(name, ch, stride))
        return name

    def conv(name, wkey, bnkey, cin, cout, k, stride, pad, relu, src, dst, biaskey=None,
             dst_coff=0, residual="", upsample_add="", second=None):
        c = ConvSpec(name, wkey, bnkey, biaskey, cin, cout, k, stride, pad, relu)
        src2 = ""
        if second is not None:
            c.wkey2, c.bnkey2, c.cin2, c.stride2, src2 = second
        convs.append(c)
        ops.append(Op("conv", c, src, dst, dst_coff, src2, residual, upsample_add))

    buf("x0", 3, 1)
    if fuse_stem_pool:
        buf("p1", 64, 4)
        conv("body.conv1", "body.conv1.weight", "body.bn1", 3, 64, 7, 2, 3, True, "x0", "p1")
        ops[-1].pool = True
    else:
        buf("c1", 64, 2)
        conv("body.conv1", "body.conv1.weight", "body.bn1", 3, 64, 7, 2, 3, True, "x0", "c1")
        buf("p1", 64, 4)
        ops.append(Op("maxpool", None, "c1", "p1"))

    cur, cur_ch, cur_stride = "p1", 64, 4
    layer_cfg = [("layer1", 64, 3, 1), ("layer2", 128, 4, 2), ("layer3", 256, 6, 2), ("layer4", 512, 3, 2)]
    feats = {}
    for lname, planes, nblocks, lstride in layer_cfg:
        for b in range(nblocks):
            s = lstride if b == 0 else 1
            out_stride = cur_stride * s
            p = f"body.{lname}.{b}"
            t1 = buf(f"{lname}.{b}.t1", planes, cur_stride)
            t2 = buf(f"{lname}.{b}.t2", planes, out_stride)
            out = buf(f"{lname}.{b}.out", planes * 4, out_stride)
            conv(f"{p}.conv1", f"{p}.conv1.weight", f"{p}.bn1", cur_ch, planes, 1, 1, 0, True, cur, t1)
            conv(f"{p}.conv2", f"{p}.conv2.weight", f"{p}.bn2", planes, planes, 3, s, 1, True, t1, t2)
            if b == 0 and fuse_downsample:
                conv(f"{p}.conv3", f"{p}.conv3.weight", f"{p}.bn3", planes, planes * 4, 1, 1, 0, True, t2, out,
                     second=(f"{p}.downsample.0.weight", f"{p}.downsample.1", cur_ch, s, cur))
            else:
                if b == 0:
                    idn = buf(f"{lname}.{b}.ds", planes * 4, out_stride)
                    conv(f"{p}.downsample", f"{p}.downsample.0.weight", f"{p}.downsample.1", cur_ch,
                         planes * 4, 1, s, 0, False, cur, idn)
                else:
                    idn = cur
                conv(f"{p}.conv3", f"{p}.conv3.weight", f"{p}.bn3", planes, planes * 4, 1, 1, 0, True,
                     t2, out, residual=idn)
            cur, cur_ch, cur_stride = out, planes * 4, out_stride
        feats[lname] = cur

    c3, c4, c5 = feats["layer2"], feats["layer3"], feats["layer4"]
    oc = CFG_RE50["out_channel"]
    # FPN: laterals; the top-down "nearest upsample + add" rides in the lateral's epilogue
    buf("fpn.o3", oc, 32)
    conv("fpn.output3", "fpn.output3.0.weight", "fpn.output3.1", 2048, oc, 1, 1, 0, True, c5, "fpn.o3")
    buf("fpn.o2s", oc, 16)
    conv("fpn.output2", "fpn.output2.0.weight", "fpn.output2.1", 1024, oc, 1, 1, 0, True, c4, "fpn.o2s",
         upsample_add="fpn.o3")
    buf("fpn.o2", oc, 16)
    conv("fpn.merge2", "fpn.merge2.0.weight", "fpn.merge2.1", oc, oc, 3, 1, 1, True, "fpn.o2s", "fpn.o2")
    buf("fpn.o1s", oc, 8)
    conv("fpn.output1", "fpn.output1.0.weight", "fpn.output1.1", 512, oc, 1, 1, 0, True, c3, "fpn.o1s",
         upsample_add="fpn.o2")
    buf("fpn.o1", oc, 8)
    conv("fpn.merge1", "fpn.merge1.0.weight", "fpn.merge1.1", oc, oc, 3, 1, 1, True, "fpn.o1s", "fpn.o1")

    # SSH: concat = channel slices of one buffer; the trailing ReLU over the concat is applied per slice
    for lvl, (src, stride) in enumerate([("fpn.o1", 8), ("fpn.o2", 16), ("fpn.o3", 32)]):
        s = f"ssh{lvl + 1}"
        f_out = buf(f"{s}.out", oc, stride)
        t5 = buf(f"{s}.t5", oc // 4, stride)
        t7 = buf(f"{s}.t7", oc // 4, stride)
        conv(f"{s}.conv3X3", f"{s}.conv3X3.0.weight", f"{s}.conv3X3.1", oc, oc // 2, 3, 1, 1, True, src,
             f_out, dst_coff=0)
        conv(f"{s}.conv5X5_1", f"{s}.conv5X5_1.0.weight", f"{s}.conv5X5_1.1", oc, oc // 4, 3, 1, 1, True,
             src, t5)
        conv(f"{s}.conv5X5_2", f"{s}.conv5X5_2.0.weight", f"{s}.conv5X5_2.1", oc // 4, oc // 4, 3, 1, 1,
             True, t5, f_out, dst_coff=oc // 2)
        conv(f"{s}.conv7X7_2", f"{s}.conv7X7_2.0.weight", f"{s}.conv7X7_2.1", oc // 4, oc // 4, 3, 1, 1,
             True, t5, t7)
        conv(f"{s}.conv7x7_3", f"{s}.conv7x7_3.0.weight", f"{s}.conv7x7_3.1", oc // 4, oc // 4, 3, 1, 1,
             True, t7, f_out, dst_coff=3 * oc // 4)
        ops.append(Op("heads", None, f_out, "heads", level=lvl))
    if fuse_siblings:
        fused, i = [], 0
        while i < len(ops):
            op = ops[i]
            nxt = ops[i + 1] if i + 1 < len(ops) else None
            if (op.kind == "conv" and nxt is not None and nxt.kind == "conv" and nxt.src == op.src and op.conv.k == 3
                    and (nxt.conv.k, nxt.conv.stride, nxt.conv.pad, nxt.conv.relu, nxt.conv.cin) ==
                    (op.conv.k, op.conv.stride, op.conv.pad, op.conv.relu, op.conv.cin)
                    and not (op.residual or nxt.residual or op.upsample_add or nxt.upsample_add or op.src2 or nxt.src2)
                    and op.conv.cout % 64 == 0 and nxt.conv.cout % 64 == 0 and op.conv.cout + nxt.conv.cout in (128, 192)
                    and op.conv.name.startswith("ssh")):
                op.sibling, op.sibling_dst, op.sibling_coff = nxt.conv, nxt.dst, nxt.dst_coff
                i += 1
            fused.append(op)
            i += 1
        ops = fused
    if fuse_chain:
        fused, i = [], 0
        while i < len(ops):
            op = ops[i]
            nxt = ops[i + 1] if i + 1 < len(ops) else None
            if (op.kind == "conv" and op.conv.k == 1 and op.conv.cout == 256 and not op.upsample_add and nxt is not None
                    and nxt.kind == "conv" and nxt.src == op.dst and nxt.conv.k == 1 and nxt.conv.stride == 1
                    and nxt.conv.cout in (64, 128) and nxt.conv.relu and not nxt.residual and not nxt.src2
                    and not nxt.upsample_add and nxt.dst_coff == 0 and op.conv.name.startswith("body.")):
                op.chain, op.chain_dst = nxt.conv, nxt.dst
                i += 1
            fused.append(op)
            i += 1
        ops = fused
    return convs, ops, bufs


# head layout of the fused N=32 1x1 GEMM per level (rows of the packed head weight):
#   [0:8)   BboxHead   (anchor0: dx,dy,dw,dh; anchor1: ...)     retinaface.py:28-37
#   [8:12)  ClassHead  (anchor0: bg,face; anchor1: bg,face)     retinaface.py:13-23
#   [12:32) LandmarkHead (anchor0: 10 values; anchor1: 10 values) retinaface.py:42-51
HEAD_ROWS = 32


def head_keys(level: int):
    return (f"BboxHead.{level}.conv1x1", f"ClassHead.{level}.conv1x1", f"LandmarkHead.{level}.conv1x1")


def state_dict_spec():
    """[(key, shape, kind)] for every tensor of the reference state_dict (456 entries, 27,293,600
    parameters).  kind in {conv_backbone, conv_neck, bn_weight, bn_bias, bn_mean, bn_var,
    bn_count, head_weight, head_bias}."""
    convs, _, _ = build_graph(fuse_downsample=False)
    spec = []
    for c in convs:
        kind = "conv_backbone" if c.name.startswith("body.") else "conv_neck"
        spec.append((c.wkey, (c.cout, c.cin, c.k, c.k), kind))
        if c.bnkey:
            spec.append((c.bnkey + ".weight", (c.cout,), "bn_weight"))
            spec.append((c.bnkey + ".bias", (c.cout,), "bn_bias"))
            spec.append((c.bnkey + ".running_mean", (c.cout,), "bn_mean"))
            spec.append((c.bnkey + ".running_var", (c.cout,), "bn_var"))
            spec.append((c.bnkey + ".num_batches_tracked", (), "bn_count"))
    for lvl in range(3):
        for hk, n in zip(head_keys(lvl), (8, 4, 20)):
            spec.append((hk + ".weight", (n, 256, 1, 1), "head_weight"))
            spec.append((hk + ".bias", (n,), "head_bias"))
    return spec
