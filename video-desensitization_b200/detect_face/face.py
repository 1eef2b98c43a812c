"""Drop-in replacement for the reference's `detect_face.face.Retinaface` (detect_face/face.py:14-231)
backed by libvdesens (sm_100a CUDA kernels behind the C-ABI of include/vdesens.h).

Same constructor kwargs/defaults, attributes and methods; same return types; same error behaviour
(Python exceptions).  Differences that are deliberate:
  * `cuda=False` raises: there is no CPU fallback in this build;
  * `backbone="mobilenet"` (cfg_mnet) runs the MobileNet0.25 variant: sub-64-channel layers as streaming kernels,
    the rest on the same tcgen05 convolution kernel with LeakyReLU(0.1) epilogues (bf16 / fp16 storage only);
  * multi-GPU is one process per GPU (vdesens_b200.shard), not nn.DataParallel (face.py:56);
  * extra optional kwargs: `state_dict` (weights given in memory instead of `model_path`),
    `batch_size` (engine batch, default = first batch seen), `cand_cap`, `max_det`, `device_index`,
    `precision` ("bf16" default; "fp16" = same kernels with half storage: boxes agree with the fp32
    reference to IoU >= 0.99 on every anchor; "fp32x" = split hi/lo planes, fp32-class results at ~1/3 speed).
"""
import time

import numpy as np
import torch

from vdesens_b200 import ops as _ops
from vdesens_b200.engine import DetectorEngine
from .utils.anchors import Anchors
from .utils.config import cfg_mnet, cfg_re50


class _NetHandle(object):
    """What callers see as `Retinaface.net`: callable on the tensor `preprocess` returns
    ((B,3,H,W) fp32, mean-subtracted) -> (loc, softmax(conf), landms), like RetinaFace.forward in eval
    mode (reference detect_face/retinaface.py:114-148).  `.to()/.eval()/.cuda()` are accepted no-ops
    (reference detect_face/detect.py:38-39 calls them)."""

    def __init__(self, owner):
        self._owner = owner

    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    def cuda(self, *a, **k):
        return self

    def __call__(self, x):
        o = self._owner
        B, _, H, W = x.shape
        if (H, W) != (o.input_shape[0], o.input_shape[1]):
            raise ValueError(f"network input {H}x{W} differs from input_shape {o.input_shape[:2]}")
        eng = o._engine_for(B)
        if B > eng.B:
            raise ValueError(f"batch {B} exceeds the engine batch {eng.B} (batch_size=...)")
        x0 = eng.buf["x0"]
        with torch.cuda.stream(eng.stream):
            eng.stream.wait_stream(torch.cuda.current_stream(eng.device))
            x0[:B, :, 3:3 + W, :3] = x.to(eng.device).permute(0, 2, 3, 1).to(eng.act_dtype)
            eng.enqueue_network()              # slots >= B keep whatever they held; their outputs are dropped below
        loc, cls, ldm = eng.head_tensors()
        return loc[:B], torch.softmax(cls[:B], dim=-1), ldm[:B]


class Retinaface(object):
    _defaults = {
        "model_path": '/home/liuwj/Retinaface_resnet50.pth',
        "backbone": 'resnet50',
        "confidence": 0.5,
        "nms_iou": 0.45,
        "input_shape": [1280, 1280, 3],
        "letterbox_image": True,
        "cuda": True,
    }

    @classmethod
    def get_defaults(cls, n):
        return cls._defaults.get(n, f"Unrecognized attribute name '{n}'")

    def __init__(self, **kwargs):
        self.__dict__.update(self._defaults)
        self.state_dict = None
        self.batch_size = None
        self.cand_cap = 4096
        self.max_det = 2048
        self.device_index = None
        self.precision = "bf16"
        for name, value in kwargs.items():
            setattr(self, name, value)
        # reference face.py:35: anything but "mobilenet" selects cfg_re50
        self.cfg = cfg_mnet if self.backbone == "mobilenet" else cfg_re50
        if not (self.cuda and torch.cuda.is_available()):
            raise RuntimeError("vdesens Retinaface needs a CUDA (sm_100a) device; there is no CPU fallback")
        idx = torch.cuda.current_device() if self.device_index is None else int(self.device_index)
        self.device = torch.device("cuda", idx)
        if self.letterbox_image:
            self.anchors = Anchors(self.cfg, image_size=[self.input_shape[0], self.input_shape[1]]).get_anchors()
            self.anchors = self.anchors.to(self.device)
        else:
            self.anchors = None
        self._engines = {}
        self.generate()

    def generate(self):
        sd = self.state_dict
        if sd is None:
            sd = torch.load(self.model_path, map_location="cpu")
        self._sd = sd
        self.net = _NetHandle(self)
        print(f'Model {self.model_path if self.state_dict is None else "<state_dict>"} loaded.')

    _MAX_ENGINES = 2          # a replica holds the packed weights and every activation buffer (GBs at 16 x 1280^2)

    def _engine_for(self, batch, exact=False):
        """The engine that runs `batch` frames.  An engine is built for a fixed batch (`batch_size` if given, else the
        first batch seen) and REUSED for every shorter batch — the callers pad (detect_rows) or slice (net) — so the
        tail batch of a directory does not build a second replica.  Only a LARGER batch (or exact=True: get_FPS times
        a one-image forward) builds a new engine; at most _MAX_ENGINES are kept, the least recently used is freed."""
        want = int(self.batch_size or batch)
        fit = [b for b in self._engines if (b == want if exact else b >= want)]
        if fit:
            b = min(fit)
        else:
            while len(self._engines) >= self._MAX_ENGINES:
                old = self._engines.pop(next(iter(self._engines)))       # dicts keep insertion (= LRU, see below) order
                del old
                torch.cuda.empty_cache()
            b = want
            self._engines[b] = DetectorEngine(self._sd, b, self.input_shape[0], self.input_shape[1], device=self.device,
                                              confidence=self.confidence, nms_iou=self.nms_iou, cand_cap=self.cand_cap,
                                              max_det=self.max_det, precision=self.precision,
                                              backbone="mobilenet" if self.backbone == "mobilenet" else "resnet50")
        eng = self._engines.pop(b)
        self._engines[b] = eng               # most recently used last
        eng.confidence, eng.nms_iou = float(self.confidence), float(self.nms_iou)
        return eng

    # ---------------------------------------------------#
    def _upload(self, images):
        return [torch.from_numpy(np.ascontiguousarray(img)).to(self.device, non_blocking=False) for img in images]

    def preprocess(self, images):
        """list of np.ndarray (H,W,3) -> (tensor (B,3,H,W) fp32 on device, shapes (B,2) [h,w])."""
        if not self.letterbox_image:
            raise ValueError("Batch inference requires letterbox_image=True for shape alignment.")
        frames = self._upload(images)
        x = _ops.ingest_nchw_f32(frames, self.input_shape[0], self.input_shape[1])
        shapes = torch.tensor([[f.shape[0], f.shape[1]] for f in frames], dtype=torch.float32).to(self.device)
        return x, shapes

    def postprocess(self, loc, conf, landms, image_shapes, batch_size):
        """decode + threshold + NMS + un-letterbox -> list of (M_i,15) device tensors."""
        H, W = self.input_shape[0], self.input_shape[1]
        cand, count, st = _ops.decode_threshold_lcl(loc.float(), conf.float(), landms.float(), H, W, self.confidence,
                                                    cap=self.cand_cap)
        shapes = [(float(h), float(w)) for h, w in image_shapes.cpu().numpy()] if self.letterbox_image else \
            [(float(H), float(W))] * batch_size
        rows, oc, st2 = _ops.nms(cand, count, self.nms_iou, _ops.unletterbox_params((H, W), shapes), max_det=self.max_det)
        oc_h = oc.cpu().numpy()
        if int(st.item()) or int(st2.item()):
            raise RuntimeError("candidate or detection capacity exceeded (raise cand_cap / max_det)")
        return [rows[i, :int(oc_h[i]), :15].clone() if oc_h[i] > 0 else torch.empty((0, 15), device=rows.device)
                for i in range(batch_size)]

    # ---------------------------------------------------#
    def detect_rows(self, images):
        """device path of detect_images: returns (rows (n,max_det,16) numpy normalised to each image,
        counts (n,)) for the n input images."""
        n = len(images)
        eng = self._engine_for(n)
        rows_all = np.zeros((n, eng.max_det, 16), np.float32)
        counts_all = np.zeros((n,), np.int32)
        for s in range(0, n, eng.B):
            chunk = list(images[s:s + eng.B])
            real = len(chunk)
            chunk += [chunk[-1]] * (eng.B - real)          # pad the tail batch
            eng.set_frames(self._upload(chunk))
            eng.forward()
            rows, counts = eng.fetch_results()
            rows_all[s:s + real] = rows[:real]
            counts_all[s:s + real] = counts[:real]
        return rows_all, counts_all

    def detect_images(self, images):
        """list of np.ndarray (H,W,3) -> list of (image copy, [[x1,y1,x2,y2],...]) in original pixels,
        descending score, not clipped (reference face.py:120-150)."""
        if not isinstance(images, list):
            images = [images]
        if not self.letterbox_image:
            raise ValueError("Batch inference requires letterbox_image=True for shape alignment.")
        old_images = [img.copy() for img in images]
        rows, counts = self.detect_rows(images)
        out = []
        for i, img in enumerate(old_images):
            boxes = []
            if counts[i] > 0:
                h, w = np.float32(img.shape[0]), np.float32(img.shape[1])
                r = rows[i, :counts[i], :4].copy()
                r *= np.array([w, h, w, h])
                boxes = r.tolist()
            out.append((img, boxes))
        return out

    # ---------------------------------------------------#
    def get_FPS(self, image, test_interval=100):
        """seconds per network forward on one image (reference face.py:155-180), device-synchronised."""
        image = np.asarray(image)
        x, _ = self.preprocess([image.astype(np.uint8)])
        eng = self._engine_for(1, exact=self.batch_size is None)      # a one-image forward, as the reference times it
        for _ in range(10):
            self.net(x)
        eng.stream.synchronize()
        t1 = time.time()
        for _ in range(test_interval):
            with torch.cuda.stream(eng.stream):
                eng.enqueue_network()
        eng.stream.synchronize()
        return (time.time() - t1) / test_interval

    def get_map_txt(self, image):
        """(M,15) numpy in pixels of the original image (reference face.py:185-230).
        Deviation: the reference letterboxes the float32 image here (face.py:196-203, cv2.resize on float data);
        this path casts to uint8 first and uses the same fixed-point uint8 resize as detect_images.  For uint8
        input at the network's own size the two agree exactly; with a real down-scale they differ by < 1 grey level
        per pixel (float bilinear vs cv2's 11-bit fixed point), which moves scores in the 4th decimal."""
        image = np.asarray(image)
        ih, iw = image.shape[:2]
        rows, counts = self.detect_rows([image.astype(np.uint8)])
        if counts[0] == 0:
            return np.array([])
        d = rows[0, :counts[0], :15].copy()
        d[:, :4] *= np.array([iw, ih, iw, ih])
        d[:, 5:] *= np.tile([iw, ih], 5)
        return d
