"""The 'kernel to beat' of SURVEY §8(d): the reference's own GPU execution model — PyTorch + cuDNN (TF32 allowed,
torch default) for the network, torchvision NMS, Python loops with per-frame syncs for the post-processing, cv2 on the
CPU for the blur — run on this B200 through the oracle port (the same torch / torchvision / cv2 calls the reference
makes; /root/reference itself is not present on the GPU box).  TEST/PROFILING TOOL, not part of the product or bench.
usage: python tools/kernel_to_beat.py [B] [steps]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "video-desensitization_b200"))
import numpy as np, torch, cv2
from torchvision.ops import batched_nms
from oracle import retinaface_oracle as ro
from vdesens_b200 import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
H, W, K = 1080, 1920, 31
dev = "cuda:0"
sd = {k: v.to(dev) for k, v in ro.to_torch_sd(synth.synthetic_state_dict(0, face_rate="r001")).items()}
frames = [synth.synthetic_frame(H, W, 1000 + i) for i in range(B)]
pri = ro.anchors(H, W).to(dev).unsqueeze(0)
torch.backends.cudnn.allow_tf32 = True
torch.backends.cudnn.benchmark = True
def step():
    t = {}
    t0 = time.perf_counter()
    x, shapes = ro.preprocess(frames, (H, W))              # numpy/cv2 on the host, as face.py:65-88
    x = x.to(dev); torch.cuda.synchronize(); t["preprocess+h2d"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    loc, conf, ldm = ro.forward(sd, x); torch.cuda.synchronize(); t["network_cudnn_tf32"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    boxes = ro.decode(loc, pri); lm = ro.decode_landm(ldm, pri)
    det = torch.cat([boxes, conf[:, :, 1:2], lm], -1)
    out = []
    for b in range(B):                                     # utils_bbox.py:114-127, one sync per frame
        d = det[b]; d = d[d[:, 4] >= 0.5]
        if len(d) == 0: out.append(d); continue
        keep = batched_nms(d[:, :4], d[:, 4], torch.zeros(len(d), device=dev), 0.4)
        out.append(d[keep])
    res = [o.cpu().numpy() for o in out]; t["postprocess"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    for i, r in enumerate(res):                            # CPU compositing, combine_detect.py:241-249 idiom
        img = frames[i].copy()
        for bx in (r[:, :4] * np.array([W, H, W, H], np.float32)).tolist():
            x1, y1, x2, y2 = (int(v) for v in bx)
            x1, y1, x2, y2 = max(0, x1), max(0, y1), min(W, x2), min(H, y2)
            if x2 > x1 and y2 > y1:
                img[y1:y2, x1:x2] = cv2.GaussianBlur(img[y1:y2, x1:x2], (K, K), 0)
    t["cpu_gaussian_blur"] = time.perf_counter() - t0
    return t
step(); step()
acc = {}
for _ in range(steps):
    for k, v in step().items(): acc[k] = acc.get(k, 0.0) + v
tot = sum(acc.values()) / steps
print(json.dumps({"what": "oracle port of the reference's GPU execution model on this B200 (torch %s, cudnn %s)" % (torch.__version__, torch.backends.cudnn.version()),
                  "batch": B, "ms_per_batch": {k: v / steps * 1e3 for k, v in acc.items()}, "frames_per_s_total": B / tot,
                  "frames_per_s_network_only": B / (acc["network_cudnn_tf32"] / steps)}))
