#!/usr/bin/env python
"""bench.py — 1080p detect + Gaussian-blur frames/s (BASELINE.json metric), device-timed.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = one pass of the hot path over one batch of 16 synthetic 1080p frames (BASELINE config[1]):
frame copy (the reference's img.copy(), combine_detect.py:247) -> ingest -> RetinaFace-ResNet50 bf16 on
tcgen05 -> decode/threshold -> NMS/un-letterbox -> Gaussian blur (k=31) over the detected faces and the
supplied plate boxes.  `value`: frames resident in HBM, CUDA events on the engine stream, max over ranks.
`e2e`: the same step through the public drop-in API (combine_detect pipeline) from HOST buffers, with the
H2D of the frames and the D2H of the processed frames + boxes inside the timed region.
N > 1 (torchrun): frames shard across ranks with no collective on the data path (weak scaling).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(ROOT, "video-desensitization_b200")
for p in (ROOT, SRC):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np
import torch

H, W, BATCH = 1080, 1920, 16
KSIZE = 31
N_PLATES = 4
CONV_GFLOP_PER_FRAME = 449.71          # 82 convs at 1080p, 2*MAC
# average dram__bytes_read.sum + dram__bytes_write.sum per convolution launch of one 16-frame step, from the ncu capture
# summarised in profiles/r01_ncu_summary.md §2 (28.54 GB over the 69 launches of a step).  Algorithmic activation
# traffic: 2.085 GB/frame unfused - 0.133 GB/frame (conv1's output stays on chip in the fused
# conv1+maxpool kernel) - 0.199 GB/frame (the three 256-channel layer1 outputs are not read back by the next conv1,
# VD_CONV_CHAIN) = 1.75 GB/frame * 16 = 28.0 GB per step, i.e. no wasted re-reads.  The figure belongs to the
# configuration it was captured on: CONV_DRAM_LAUNCHES convolution launches per step.
CONV_DRAM_BYTES_PER_LAUNCH = 4.136e8
CONV_DRAM_LAUNCHES = 69
METRIC = "1080p detect+blur frames/sec"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p["bf16_tflops_sustained"]), "measured"
    except Exception:
        return 6650.0, 1400.0, "fallback"


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons; started before the warm-up (nvidia-smi needs a moment to come
    up) and evaluated over the wall-clock windows of the timed regions"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.proc, self.windows = index, [], None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def window(self, t0, t1):
        self.windows.append((t0, t1))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        sm, mx, reasons, n_in = [], None, set(), 0
        for ts, s in self.samples:
            inside = any(t0 - 0.05 <= ts <= t1 + 0.1 for t0, t1 in self.windows)
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                clk, mx = float(f[0]), float(f[1])
            except ValueError:
                continue
            if not inside:
                continue
            n_in += 1
            sm.append(clk)
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": n_in, "samples_total": len(self.samples),
                "windows": "device-timed steps + end-to-end steps"}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path (torch CPU fp32 detector + cv2.GaussianBlur idiom)
# --------------------------------------------------------------------------------------------------
def cpu_detect_blur(sd_t, frames, plates, threads):
    from oracle import retinaface_oracle as ro
    import cv2
    torch.set_num_threads(threads)
    cv2.setNumThreads(threads)
    res = ro.detect_images(sd_t, frames, (H, W), 0.5, 0.4)
    out = []
    for (img, boxes), pl in zip(res, plates):
        img = img.copy()
        allb = [tuple(int(v) for v in b) for b in boxes] + list(pl)
        for (x1, y1, x2, y2) in allb:
            x1, y1, x2, y2 = max(0, x1), max(0, y1), min(W, x2), min(H, y2)
            if x2 > x1 and y2 > y1:
                img[y1:y2, x1:x2] = cv2.GaussianBlur(img[y1:y2, x1:x2], (KSIZE, KSIZE), 0)
        out.append(img)
    return out


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port: /root/reference does
    not exist on the GPU box and is pure Python + torch/cv2 calls, which the port restates 1:1)."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    from oracle import retinaface_oracle as ro
    from vdesens_b200 import synth
    threads = os.cpu_count() or 1
    sample = 2
    sd_t = ro.to_torch_sd(synth.synthetic_state_dict(0, face_rate="r001"))
    frames = [synth.synthetic_frame(H, W, 1000 + i) for i in range(sample)]
    plates = [synth.synthetic_boxes(N_PLATES, H, W, seed=50 + i) for i in range(sample)]
    for _ in range(args.warmup):
        cpu_detect_blur(sd_t, frames, plates, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_detect_blur(sd_t, frames, plates, threads)
    dt = time.perf_counter() - t0
    fps = sample * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                             "sample": f"{sample} frames of the 16-frame 1080p batch per step"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(n_gpus):
    return {"workload": "BASELINE configs[1]: 1080p synthetic frames, batch 16 per GPU, RetinaFace-ResNet50 bf16 detect "
                        f"+ Gaussian blur k={KSIZE} over detected faces + {N_PLATES} plate boxes/frame",
            "batch_per_gpu": BATCH, "frame": [H, W, 3], "input_shape": [H, W, 3], "confidence": 0.5, "nms_iou": 0.4,
            "weights": "seeded random init, BN calibrated, ~0.1% of anchors above threshold (~90 faces/frame)",
            "l2": "inputs cycle through 3 distinct 100 MB batches; each step streams > 10 GB of activations (>> 126 MB L2)",
            "parallelism": f"frame-parallel x{n_gpus}, no collective"}


# --------------------------------------------------------------------------------------------------
def run_ours(args):
    rank, world, local = dist_env()
    import torch.distributed as dist
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    from vdesens_b200 import shard, synth
    # one rank per GPU: keep the rank and its pinned staging buffers on the GPU's NUMA node (host plumbing, N > 1 only)
    numa_cpus = shard.bind_to_gpu_numa_node(local) if (world > 1 and os.environ.get("VD_NUMA_BIND", "1") != "0") else None
    from vdesens_b200.engine import DetectorEngine
    from vdesens_b200.pipeline import Anonymizer

    hbm_peak, tf_peak, peak_src = load_peaks()
    sd = synth.synthetic_state_dict(0, face_rate="r001")
    precision = os.environ.get("VD_BENCH_PRECISION", "bf16")        # bf16 = BASELINE's dtype; fp16 / fp32x for comparison
    eng = DetectorEngine(sd, BATCH, H, W, device=dev, confidence=0.5, nms_iou=0.4, precision=precision)
    NB = 3                                   # distinct resident batches, cycled
    base = rank * 100000                     # every rank owns different frames (sharded stream)
    host_batches = [[synth.synthetic_frame(H, W, 1000 + base + b * BATCH + i) for i in range(BATCH)] for b in range(NB)]
    pristine = [torch.from_numpy(np.stack(hb)).to(dev) for hb in host_batches]
    plates = [[synth.synthetic_boxes(N_PLATES, H, W, seed=50 + base + b * BATCH + i) for i in range(BATCH)]
              for b in range(NB)]
    work = torch.empty_like(pristine[0])
    eng.set_frames(work)
    eng.set_ext_boxes(plates[0])
    st = eng.stream

    def step(i, ev=None):
        with torch.cuda.stream(st):
            work.copy_(pristine[i % NB], non_blocking=True)      # the reference's img.copy() (combine_detect.py:247)
            if ev is not None:
                ev[4].record(st)
            eng.enqueue_ingest()
            if ev is not None:
                ev[0].record(st)
            eng.enqueue_network()
            if ev is not None:
                ev[1].record(st)
            eng.enqueue_postprocess()
            if ev is not None:
                ev[2].record(st)
            eng.enqueue_composite("gauss", ksize=KSIZE)
            if ev is not None:
                ev[3].record(st)

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    for i in range(args.warmup):
        step(i)
    st.synchronize()
    eng.fetch_results()                      # raises on any device status bit

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    barrier()
    tw0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    seg = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(args.steps)]
    e0.record(st)
    for i in range(args.steps):
        step(i, seg[i])
    e1.record(st)
    barrier()
    ms_total = e0.elapsed_time(e1)
    if sampler:
        sampler.window(tw0, time.time())
    rows, counts = eng.fetch_results()
    net_ms = sum(s[0].elapsed_time(s[1]) for s in seg) / args.steps
    post_ms = sum(s[1].elapsed_time(s[2]) for s in seg) / args.steps
    comp_ms = sum(s[2].elapsed_time(s[3]) for s in seg) / args.steps
    ingest_ms = sum(s[4].elapsed_time(s[0]) for s in seg) / args.steps
    copy_ms = ms_total / args.steps - net_ms - post_ms - comp_ms - ingest_ms
    # algorithmic bytes of the streaming stages: each ROI pixel read once and written once for the blur;
    # u8 frame in + 4-channel bf16 x0 out for the ingest; the fp32 head rows (16 values per anchor) for the decode
    roi_bytes = 0                            # boxes of the last timed step (its detections + the fixed plate boxes)
    for f in range(BATCH):
        bx = [[float(v) for v in r[:4] * np.array([W, H, W, H], np.float32)] for r in rows[f, :int(counts[f])]]
        for (x1, y1, x2, y2) in [tuple(int(v) for v in b) for b in bx] + list(plates[0][f]):
            x1, y1, x2, y2 = max(0, x1), max(0, y1), min(W, x2), min(H, y2)
            if x2 > x1 and y2 > y1:
                roi_bytes += 6 * (x2 - x1) * (y2 - y1)
    n_anchor = 2 * sum(-(-H // s) * -(-W // s) for s in (8, 16, 32))
    x0_pitch = int(eng.buf["x0"].shape[2])

    def hbm_stage(nbytes, ms, note):
        gbs = nbytes / ms / 1e6
        return {"bytes": int(nbytes), "ms": ms, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / hbm_peak, "note": note}

    stage_hbm = {
        "frame_copy": hbm_stage(2 * BATCH * H * W * 3, copy_ms, "device copy of the batch (the reference's img.copy()), torch copy kernel"),
        "ingest": hbm_stage(BATCH * (3 * H * W + 8 * H * x0_pitch), ingest_ms, "u8 HWC read + 4-channel bf16 x0 written"),
        "decode_threshold_nms": hbm_stage(BATCH * n_anchor * 16 * 4, post_ms,
                                          "fp32 head rows read once; the NMS part is compare-bound (no bandwidth roofline)"),
        "composite": hbm_stage(roi_bytes, comp_ms, "2*3*w*h per clipped box; sequential box chains make this stage latency-bound"),
    }

    # ---- e2e: public API, host buffers, copies inside the timed region ----
    # StreamingAnonymizer.run() is the call a user makes for a stream of batches: every step copies its
    # 16 frames from pinned host memory to the device, runs detect + blur, and copies the processed frames
    # and the boxes back to pinned host memory; copies of neighbouring batches overlap the compute.
    from vdesens_b200.pipeline import StreamingAnonymizer
    sa = StreamingAnonymizer(eng, "gauss", ksize=KSIZE, depth=int(os.environ.get("VD_E2E_DEPTH", "2")))
    pinned = [torch.from_numpy(np.stack(hb)).pin_memory() for hb in host_batches]
    for _ in sa.run([pinned[i % NB] for i in range(3)], [plates[i % NB] for i in range(3)]):
        pass
    barrier()
    e2e_steps = args.steps
    n_faces, checksum = 0, 0
    t0 = time.perf_counter()
    for out, boxes in sa.run((pinned[i % NB] for i in range(e2e_steps)), (plates[i % NB] for i in range(e2e_steps))):
        n_faces += sum(len(b) for b in boxes)
        checksum += int(out[0, 0, 0, 0])                # the result is host-visible here
    torch.cuda.synchronize(dev)
    e2e_ms = (time.perf_counter() - t0) * 1e3
    if sampler:
        sampler.window(time.time() - e2e_ms / 1e3, time.time())
    clocks = sampler.stop() if sampler else None
    e2e_host = {k: (v * 1e3 / max(1, sa.stats["batches"]) if k != "batches" else v) for k, v in sa.stats.items()}

    t = torch.tensor([ms_total, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # order-preserving host gather of the per-frame box counts (sequence = rank-major)
        gathered = [None] * world
        dist.all_gather_object(gathered, counts.tolist())
    ms_total, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    frames_total = BATCH * args.steps * world
    value = frames_total / (ms_total / 1e3)
    e2e_value = BATCH * e2e_steps * world / (e2e_ms / 1e3)
    launches_per_step = eng.launches_per_forward + 2          # + composite plan/run
    tflops = CONV_GFLOP_PER_FRAME * BATCH / net_ms            # GFLOP / ms = TFLOP/s
    n_conv = sum(1 for kind, _, _ in eng.plans if kind == "conv")
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": {"bf16": "bf16", "fp16": "f16", "fp32x": "f16x3 (split hi/lo, fp32-class)"}[precision],
        "data": "synthetic", "config": workload_config(world),
        "host_affinity": (f"each rank bound to the CPUs local to its GPU (NVML; rank 0: {len(numa_cpus)} CPUs)" if numa_cpus else "unbound"),
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": BATCH * H * W * 3,
                "d2h_bytes_per_step": BATCH * H * W * 3 + int(rows.nbytes) + int(counts.nbytes) + 4,
                "steps": e2e_steps, "host_ms_per_batch": e2e_host, "timing": "wall clock around StreamingAnonymizer.run over the steps, pinned H2D + D2H of frames and boxes every step (double-buffered), max over ranks"},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "conv_igemm_kernel + stem_pool_kernel (%d launches per step: 82 convolutions, 4 projections fused into conv3, 9 heads fused to 3, maxpool fused into conv1, %d next-block conv1 chained into the preceding conv3)" % (n_conv, getattr(eng, "n_chained", 0)),
                     "achieved": tflops, "peak": tf_peak, "unit": "TFLOP/s", "frac": tflops / tf_peak,
                     "peak_source": peak_src + " bf16_tflops_sustained",
                     "traffic": CONV_DRAM_BYTES_PER_LAUNCH if (precision == "bf16" and n_conv == CONV_DRAM_LAUNCHES) else None,
                     "traffic_note": "avg DRAM bytes per conv launch (ncu, profiles/r01_ncu_summary.md)",
                     "launches_per_step": n_conv, "avg_launch_ms": net_ms / n_conv,
                     "algorithmic_flops_per_launch": CONV_GFLOP_PER_FRAME * 1e9 * BATCH / n_conv,
                     "algorithmic_flops_per_step": CONV_GFLOP_PER_FRAME * 1e9 * BATCH, "segment_ms": net_ms},
        "segments_ms": {"network": net_ms, "decode_nms": post_ms, "composite": comp_ms,
                        "ingest": ingest_ms, "frame_copy": copy_ms},
        "stage_hbm": stage_hbm, "hbm_peak_gbs": hbm_peak,
        "faces_per_frame": float(np.mean(counts)),
    }
    if world == 1:
        # CPU baseline: the oracle port on this box's host cores, bounded sample
        from oracle import retinaface_oracle as ro
        threads = os.cpu_count() or 1
        sample = 2
        sd_t = ro.to_torch_sd(sd)
        fr = host_batches[0][:sample]
        pl = plates[0][:sample]
        cpu_detect_blur(sd_t, fr[:1], pl[:1], threads)       # warm-up
        t0 = time.perf_counter()
        cpu_detect_blur(sd_t, fr, pl, threads)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": sample / dt, "unit": "frames/s", "cores": threads, "kind": "port",
                                "sample": f"{sample} of the 16 frames of one 1080p batch, torch CPU fp32 + cv2.GaussianBlur"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
        run_ours(args)


if __name__ == "__main__":
    main()
