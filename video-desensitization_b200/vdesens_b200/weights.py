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


def pack_all(sd, dtype=torch.bfloat16, split=False):
    """split=True (fp32-class mode, VD_CONV_SPLIT): activations are stored as planes [hi | lo | hi], so every tap's
    weight slice becomes [w_hi | w_hi | w_lo] (x_hi*w_hi + x_lo*w_hi + x_hi*w_lo); the stem, whose input is exact,
    gets [w_hi | w_lo] per filter row.
    Returns {layer name: (w_packed bf16 CPU tensor [Cout, K], bias fp32 CPU tensor [Cout])} for every
    conv of the graph plus 'heads.{0,1,2}' (the three 1x1 heads of a level fused to one N=32 GEMM)."""
    sd = to_numpy_sd(sd)
    convs, _, _ = arch.build_graph(fuse_downsample=not split)
    packed = {}
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
            hi, lo = _split(mat, dtype)
            packed[f"heads.{lvl}"] = (torch.cat([hi, hi, lo], dim=1).contiguous(), torch.from_numpy(np.concatenate(bs)))
        else:
            packed[f"heads.{lvl}"] = (torch.from_numpy(mat).to(dtype), torch.from_numpy(np.concatenate(bs)))
    return packed
