"""PriorBox table with the reference's interface (detect_face/utils/anchors.py:7-41), vectorised.
The device decode kernel recomputes priors analytically; this table exists for callers that read
`Retinaface.anchors`."""
from math import ceil

import numpy as np
import torch


class Anchors(object):
    def __init__(self, cfg, image_size=None):
        self.min_sizes = cfg["min_sizes"]
        self.steps = cfg["steps"]
        self.clip = cfg["clip"]
        self.image_size = image_size
        self.feature_maps = [[ceil(self.image_size[0] / step), ceil(self.image_size[1] / step)] for step in self.steps]

    def get_anchors(self):
        h, w = self.image_size[0], self.image_size[1]
        levels = []
        for k, (fh, fw) in enumerate(self.feature_maps):
            cy = (np.arange(fh, dtype=np.float64) + 0.5) * self.steps[k] / h
            cx = (np.arange(fw, dtype=np.float64) + 0.5) * self.steps[k] / w
            t = np.empty((fh, fw, len(self.min_sizes[k]), 4), np.float64)
            t[..., 0] = cx[None, :, None]
            t[..., 1] = cy[:, None, None]
            for a, ms in enumerate(self.min_sizes[k]):
                t[:, :, a, 2] = ms / w
                t[:, :, a, 3] = ms / h
            levels.append(t.reshape(-1, 4))
        out = torch.from_numpy(np.concatenate(levels, 0).astype(np.float32))
        if self.clip:
            out.clamp_(max=1, min=0)
        return out
