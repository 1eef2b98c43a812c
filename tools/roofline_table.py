"""Per-launch roofline of one bench step (no GPU needed): joins profiles/<tag>_conv_step_dram.csv (ncu: duration, DRAM
bytes, tensor-pipe activity of every convolution launch of one 16-frame 1080p step) with the algorithmic FLOPs of the
layer(s) each launch computes (vdesens_b200/arch.py) and bounds every launch by
    t_bound = max(DRAM bytes / measured copy peak, FLOPs / measured bf16 peak)        (MEASURED_PEAKS.json).
usage: python tools/roofline_table.py [tag] > profiles/<tag>_roofline_table.md"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-desensitization_b200"))
from vdesens_b200 import arch

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
B, H, W = 16, 1080, 1920
try:
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    hbm, tf, src = float(pk["hbm_gbs"]) * 1e9, float(pk["bf16_tflops_sustained"]) * 1e12, "MEASURED_PEAKS.json"
except Exception:
    hbm, tf, src = 6547.8e9, 1375.9e12, "SURVEY 8(d) copy of MEASURED_PEAKS.json"
convs, ops, bufs = arch.build_graph(fuse_stem_pool=True, fuse_chain=True, fuse_siblings=(tag != "r01"))
flops_px = {c.name: 2.0 * c.cout * (c.cin * c.k * c.k + c.cin2) for c in convs}
hw = {b.name: arch.feature_hw(H, W, b.stride) for b in bufs}
launches = []                                  # (label, flops) in plan order = launch order
for op in ops:
    if op.kind == "conv":
        h, w = hw[op.dst]
        if op.pool:                            # dst holds the pooled tensor; conv1 itself runs at stride 2
            h, w = arch.feature_hw(H, W, 2)
        f = flops_px[op.conv.name] * h * w * B
        label = op.conv.name
        if op.chain is not None:
            f += flops_px[op.chain.name] * h * w * B
            label += " + " + op.chain.name
        if op.sibling is not None:
            f += flops_px[op.sibling.name] * h * w * B
            label += " + " + op.sibling.name.split(".")[-1]
        launches.append((label, f))
    elif op.kind == "heads":
        h, w = hw[op.src]
        launches.append((f"heads.{op.level}", 2.0 * 256 * 32 * h * w * B))
rows = list(csv.DictReader(l for l in open(os.path.join(ROOT, "profiles", f"{tag}_conv_step_dram.csv")) if not l.startswith("#")))
assert len(rows) == len(launches), (len(rows), len(launches))
print(f"# Per-launch roofline of one bench step ({B} x {H}x{W}), peaks from {src}: HBM {hbm / 1e12:.3f} TB/s, bf16 {tf / 1e12:.1f} TFLOP/s\n")
print("ncu durations are serialised and cold-cache (`profiles/%s_ncu_summary.md` §2); `bound` = max(DRAM-time, tensor-time), `eff` = bound / measured.\n" % tag)
print("| # | layer(s) | us | DRAM MB | GFLOP | DRAM-bound us | tensor-bound us | bound | eff |")
print("|---|---|---:|---:|---:|---:|---:|---|---:|")
tot_t = tot_b = tot_hbm = tot_tc = 0.0
gaps = []
cls = {"hbm": [0.0, 0.0], "tensor": [0.0, 0.0]}
for i, (r, (label, f)) in enumerate(zip(rows, launches)):
    t = float(r["us"])
    by = float(r["dram_read_bytes"]) + float(r["dram_write_bytes"])
    th, tt = by / hbm * 1e6, f / tf * 1e6
    bound = max(th, tt)
    kind = "hbm" if th >= tt else "tensor"
    cls[kind][0] += t
    cls[kind][1] += bound
    tot_t += t; tot_b += bound; tot_hbm += th; tot_tc += tt
    gaps.append((t - bound, label.replace("body.", ""), t, kind, bound / t))
    print(f"| {i} | {label.replace('body.', '')} | {t:.1f} | {by / 1e6:.0f} | {f / 1e9:.1f} | {th:.1f} | {tt:.1f} | {kind} | {bound / t:.2f} |")
print(f"\nSum over the {len(rows)} launches: measured {tot_t / 1e3:.2f} ms; sum of per-launch bounds {tot_b / 1e3:.2f} ms "
      f"(= {tot_b / tot_t:.0%} of measured); all-DRAM {tot_hbm / 1e3:.2f} ms, all-tensor {tot_tc / 1e3:.2f} ms.")
for k, (t, b) in cls.items():
    print(f"* launches bounded by {k}: {t / 1e3:.2f} ms measured, {b / 1e3:.2f} ms at the bound ({b / t:.0%}).")
print("\n## Where the gap is (measured - bound, largest first)\n")
print("| gap us | layer(s) | us | bound | eff |")
print("|---:|---|---:|---|---:|")
for g, label, t, kind, eff in sorted(gaps, reverse=True)[:16]:
    print(f"| {g:.0f} | {label} | {t:.1f} | {kind} | {eff:.2f} |")
print(f"\nTotal gap {sum(g[0] for g in gaps) / 1e3:.2f} ms.  conv1's bound counts its 147 real taps x channels; the kernel issues K = 224 (8 taps x 4 "
      "channels per filter row) at N = 64, so its structural floor is about three times that bound.")
