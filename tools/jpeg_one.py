"""One device JPEG decode + encode of a 16 x 1080p batch (for ncu).  usage: python tools/jpeg_one.py [sub_bytes]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-desensitization_b200"))
import cv2, numpy as np, torch
from vdesens_b200 import synth
from vdesens_b200.jpeg import JpegDecoder, JpegEncoder
sub = int(sys.argv[1]) if len(sys.argv) > 1 else 512
B, H, W = 16, 1080, 1920
files = [cv2.imencode(".jpg", cv2.cvtColor(synth.synthetic_frame(H, W, 8000 + i), cv2.COLOR_RGB2BGR))[1].tobytes() for i in range(B)]
dec = JpegDecoder(B, sub_bytes=sub)
for _ in range(2):
    frames, cpu = dec.decode(files)
enc = JpegEncoder(B, H, W)
out = enc.encode(frames)
print("ok", cpu, dec.sync_rounds(), len(out[0]))
