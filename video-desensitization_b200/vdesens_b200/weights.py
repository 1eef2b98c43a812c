"""Weight packer: reference state_dict (.pth format of detect_face/face.py:52-56) -> BN-folded bf16
K-major GEMM operands for conv_tcgen05.cu.

Eval-mode BN (eps 1e-5) folds into the convolution:  w' = w * g/sqrt(v+eps),  b' = beta - mean * g/sqrt(v+eps)
(SURVEY §8a epilogue kinds).  K order of the packed matrix is (r, s, cin), matching the TMA tap loop.
"""
import numpy as np
import torch

from . import arch


def _fold(sd, c: arch.ConvSpec, second: bool = False):
    wkey, bnkey = (c.wkey2, c.bnkey2) if second else (c.wkey, c.bnkey)
    w = np.asarray(sd[wkey], np.float64)
    if bnkey:
        g = np.asarray(sd[bnkey + ".weight"], np.float64)
        beta = np.asarray(sd[bnkey + ".bias"], np.float64)
        mean = np.asarray(sd[bnkey + ".running_mean"], np.float64)
        var = np.asarray(sd[bnkey + ".running_var"], np.float64)
        scale = g / np.sqrt(var + arch.BN_EPS)
        w = w * scale[:, None, None, None]
        b = beta - mean * scale
    else:
        b = np.zeros(w.shape[0], np.float64)
    if c.biaskey and not second:
        b = b + np.asarray(sd[c.biaskey], np.float64)
    return w.astype(np.float32), b.astype(np.float32)


def to_numpy_sd(sd):
    out = {}
    for k, v in sd.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k[7:] if k.startswith("module.") else k] = v
    return out


def _split(mat, dtype):
    """fp32 matrix -> (hi, lo) in `dtype` with hi + lo ~= mat to ~2x the mantissa bits"""
    t = torch.from_numpy(np.ascontiguousarray(mat))
    hi = t.to(dtype)
    lo = (t - hi.float()).to(dtype)
    return hi, lo


class Packed(dict):
    """{layer name: (w, bias)} plus `acc_scale` {layer name: float}: split mode packs w * 2^k (see _split_scale)"""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.acc_scale = {}


def _split_scale(mat):
    """power of two 2^k with max|mat| * 2^k in [8192, 16384): BN-folded weights are O(1e-2), so the lo half of an
    unscaled fp16 hi/lo pair is a SUBNORMAL (absolute resolution 6e-8, i.e. ~2^-18 of the weight, measured: heads 20x
    further from fp32 than cuDNN's own fp32); scaled, it is a normal number and the pair carries 22 bits.  Exact (a
    power of two), undone on the fp32 accumulator by acc_scale = 2^-k."""
    m = float(np.abs(mat).max())
    if not np.isfinite(m) or m <= 0.0:
        return 1.0
    return float(2.0 ** np.floor(np.log2(16384.0 / m)))


def pack_all(sd, dtype=torch.bfloat16, split=False):
    """split=True (fp32-class mode, VD_CONV_SPLIT): activations are stored as planes [hi | lo | hi], so every tap's
    weight slice becomes [w_hi | w_hi | w_lo] (x_hi*w_hi + x_lo*w_hi + x_hi*w_lo); the stem, whose input is exact,
    gets [w_hi | w_lo] per filter row.
    Returns {layer name: (w_packed bf16 CPU tensor [Cout, K], bias fp32 CPU tensor [Cout])} for every
    conv of the graph plus 'heads.{0,1,2}' (the three 1x1 heads of a level fused to one N=32 GEMM)."""
    sd = to_numpy_sd(sd)
    convs, _, _ = arch.build_graph(fuse_downsample=not split)
    packed = Packed()
    for c in convs:
        w, b = _fold(sd, c)
        if c.name == "body.conv1":
            # conv1 on the ingest layout: K = 7 filter rows x (8 taps x 4 channels); tap 7 and
            # channel 3 are zero padding
            wp = np.zeros((c.cout, 7, 8, 4), np.float32)
            wp[:, :, :7, :3] = np.transpose(w, (0, 2, 3, 1))
            mat = wp.reshape(c.cout, -1)
        else:
            mat = np.transpose(w, (0, 2, 3, 1)).reshape(c.cout, -1)
        if c.wkey2:      # fused projection: [conv3 K | downsample K], biases add
            assert not split
            w2, b2 = _fold(sd, c, second=True)
            mat = np.concatenate([mat, w2.reshape(c.cout, -1)], axis=1)
            b = b + b2
        if split:
            sc = _split_scale(mat)
            packed.acc_scale[c.name] = 1.0 / sc
            mat = mat * np.float32(sc)
            if c.name == "body.conv1":
                hi, lo = _split(mat.reshape(c.cout, 7, 32), dtype)
                wt = torch.cat([hi, lo], dim=2).reshape(c.cout, -1).contiguous()
            else:
                hi, lo = _split(mat.reshape(c.cout, c.k * c.k, c.cin), dtype)
                wt = torch.cat([hi, hi, lo], dim=2).reshape(c.cout, -1).contiguous()
            packed[c.name] = (wt, torch.from_numpy(b.copy()))
            continue
        packed[c.name] = (torch.from_numpy(np.ascontiguousarray(mat)).to(dtype),
                          torch.from_numpy(b.copy()))
    for lvl in range(3):
        ws, bs = [], []
        for hk in arch.head_keys(lvl):
            ws.append(np.asarray(sd[hk + ".weight"], np.float32).reshape(-1, 256))
            bs.append(np.asarray(sd[hk + ".bias"], np.float32))
        mat = np.concatenate(ws, 0)
        assert mat.shape == (arch.HEAD_ROWS, 256)
        if split:
            sc = _split_scale(mat)
            packed.acc_scale[f"heads.{lvl}"] = 1.0 / sc
            hi, lo = _split(mat * np.float32(sc), dtype)
            packed[f"heads.{lvl}"] = (torch.cat([hi, hi, lo], dim=1).contiguous(), torch.from_numpy(np.concatenate(bs)))
        else:
            packed[f"heads.{lvl}"] = (torch.from_numpy(mat).to(dtype), torch.from_numpy(np.concatenate(bs)))
    return packed


# ------------------------------------------------------------------------------------------------
# MobileNet0.25 variant (arch_mnet.py)
# ------------------------------------------------------------------------------------------------
def _fold_keys(sd, wkey, bnkey):
    w = np.asarray(sd[wkey], np.float64)
    g = np.asarray(sd[bnkey + ".weight"], np.float64)
    beta = np.asarray(sd[bnkey + ".bias"], np.float64)
    mean = np.asarray(sd[bnkey + ".running_mean"], np.float64)
    var = np.asarray(sd[bnkey + ".running_var"], np.float64)
    scale = g / np.sqrt(var + arch.BN_EPS)
    return (w * scale[:, None, None, None]).astype(np.float32), (beta - mean * scale).astype(np.float32)


def pack_mnet(sd, dtype=torch.bfloat16):
    """{op name: (w, bias fp32, slope fp32)} for every op of arch_mnet.build_graph().  Layouts:
    stem  w [(r,s,c4)=36][8]         dw  w [9][C]         pw  w [Cin][Cout]
    conv  w [Cout][(r,s,cin)] (K-major GEMM operand of conv_tcgen05.cu)      heads  w [32][128] over [S1 | S3].
    slope = negative-side factor per output channel: 0 ReLU, 0.1 LeakyReLU(0.1), 1 identity."""
    from . import arch_mnet as am
    sd = to_numpy_sd(sd)
    L = am.LEAKY
    packed = {}

    def t(a):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))

    def put(name, w, b, slope):
        packed[name] = (t(w).to(dtype), t(b), t(np.broadcast_to(np.asarray(slope, np.float32), b.shape)))

    w, b = _fold_keys(sd, "body.stage1.0.0.weight", "body.stage1.0.1")          # (8,3,3,3)
    ws = np.zeros((3, 3, 4, 8), np.float32)
    ws[:, :, :3, :] = np.transpose(w, (2, 3, 1, 0))
    put("body.stage1.0", ws.reshape(36, 8), b, L)
    for st, i, cin, cout, _ in am.DW_BLOCKS:
        p = f"body.{st}.{i}"
        w, b = _fold_keys(sd, p + ".0.weight", p + ".1")                        # (cin,1,3,3)
        put(p + ".dw", np.transpose(w[:, 0], (1, 2, 0)).reshape(9, cin), b, L)
        w, b = _fold_keys(sd, p + ".3.weight", p + ".4")                        # (cout,cin,1,1)
        if cin % 64 == 0:
            put(p + ".pw", w.reshape(cout, cin), b, L)
        else:
            put(p + ".pw", w.reshape(cout, cin).T, b, L)
    for name in ("output1", "output2", "output3", "merge1", "merge2"):
        w, b = _fold_keys(sd, f"fpn.{name}.0.weight", f"fpn.{name}.1")
        put(f"fpn.{name}", np.transpose(w, (0, 2, 3, 1)).reshape(w.shape[0], -1), b, L)
    for lvl in range(3):
        s = f"ssh{lvl + 1}"
        f = {n: _fold_keys(sd, f"{s}.{n}.0.weight", f"{s}.{n}.1")
             for n in ("conv3X3", "conv5X5_1", "conv5X5_2", "conv7X7_2", "conv7x7_3")}
        # S1: [conv3X3 | conv5X5_1 | 0]
        W = np.zeros((64, 3, 3, 64), np.float32)
        B = np.zeros((64,), np.float32)
        W[0:32] = np.transpose(f["conv3X3"][0], (0, 2, 3, 1))
        W[32:48] = np.transpose(f["conv5X5_1"][0], (0, 2, 3, 1))
        B[0:32], B[32:48] = f["conv3X3"][1], f["conv5X5_1"][1]
        put(f"{s}.S1", W.reshape(64, -1), B, np.r_[np.zeros(32), np.full(16, L), np.zeros(16)])
        # S2: reads channels 32:48 of S1 -> [conv5X5_2 | conv7X7_2 | 0]
        W = np.zeros((64, 3, 3, 64), np.float32)
        B = np.zeros((64,), np.float32)
        W[0:16, :, :, 32:48] = np.transpose(f["conv5X5_2"][0], (0, 2, 3, 1))
        W[16:32, :, :, 32:48] = np.transpose(f["conv7X7_2"][0], (0, 2, 3, 1))
        B[0:16], B[16:32] = f["conv5X5_2"][1], f["conv7X7_2"][1]
        put(f"{s}.S2", W.reshape(64, -1), B, np.r_[np.zeros(16), np.full(16, L), np.zeros(32)])
        # S3: channels 0:16 pass through an identity centre tap, [16:32] = conv7x7_3 of channels 16:32
        W = np.zeros((64, 3, 3, 64), np.float32)
        B = np.zeros((64,), np.float32)
        for c in range(16):
            W[c, 1, 1, c] = 1.0
        W[16:32, :, :, 16:32] = np.transpose(f["conv7x7_3"][0], (0, 2, 3, 1))
        B[16:32] = f["conv7x7_3"][1]
        put(f"{s}.S3", W.reshape(64, -1), B, np.zeros(64))
        # heads over [S1 (64) | S3 (64)]: concat channel j of the reference = S1[j] (j < 32), S3[j - 32] (j >= 32)
        ws, bs = [], []
        for hk in ("BboxHead", "ClassHead", "LandmarkHead"):
            ws.append(np.asarray(sd[f"{hk}.{lvl}.conv1x1.weight"], np.float32).reshape(-1, 64))
            bs.append(np.asarray(sd[f"{hk}.{lvl}.conv1x1.bias"], np.float32))
        hw = np.concatenate(ws, 0)
        assert hw.shape == (am.HEAD_ROWS, 64)
        mat = np.zeros((am.HEAD_ROWS, 128), np.float32)
        mat[:, 0:32] = hw[:, 0:32]
        mat[:, 64:96] = hw[:, 32:64]
        put(f"heads.{lvl}", mat, np.concatenate(bs), 1.0)
    return packed
