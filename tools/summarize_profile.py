"""Turns the CSVs written by tools/gpu_profile.sh (gpurun_out/) into profiles/<tag>_*.{csv,md} sections 1-3.
usage: python tools/summarize_profile.py r01    (the hand-written findings section is kept in the .md)"""
import csv, os, re, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def short(k):
    k = re.sub(r"^void\s+", "", k)
    k = re.sub(r"<unnamed>::|\(anonymous namespace\)::", "", k)
    k = re.sub(r"\(int\)|\(bool\)", "", k)
    if not k.startswith("at::"):
        k = re.sub(r"\(.*$", "", k)          # drop the argument list
    return k[:80]


def rows_of(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    return list(csv.DictReader(lines))


def val(r):
    return float(r["Metric Value"].replace(",", ""))


PREC = os.environ.get("VD_BENCH_PRECISION", "fp16")
out = [f"# ncu evidence {tag} (B200, driver 580, CUDA 12.9) — final state of the round, precision {PREC}\n",
       "Command profiled (each ncu pass only after the same command exited 0 without ncu): `python bench.py --no-extras --steps 2 --warmup 3`.",
       "Recipe: `tools/gpu_profile.sh`; this file: `tools/summarize_profile.py`.  Numbers printed by runs under ncu are never bench values.\n"]

# ---- 1. launch list ----
rows = rows_of(os.path.join(G, "launches.csv"))
agg = collections.OrderedDict()
for r in rows:
    k = short(r["Kernel Name"])
    n, t = agg.get(k, (0, 0.0))
    agg[k] = (n + 1, t + val(r) / 1e3)
tot = sum(t for _, t in agg.values())
out.append("## 1. Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none -c 700`)\n")
out.append("Per-launch times are cold-cache and serialised: compare SHARES with bench.py's live CUDA-event segments "
           f"(`profiles/{tag}_bench_final.json`, `segments_ms`).\n")
out.append("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k}` | {n} | {t:.1f} | {100 * t / tot:.1f} % |")
conv = sum(t for k, (n, t) in agg.items() if "conv_igemm" in k or "stem_pool" in k or "conv3x3_rows" in k or "maxpool" in k)
# warm-up + timed + end-to-end steps all run the same kernels, so shares over the whole list are shares of a step
comp = sum(t for k, (n, t) in agg.items() if "composite" in k)
out.append(f"\nconvolution (conv_igemm + stem_pool + conv3x3_rows) share: {100 * conv / tot:.1f} %; compositing share: {100 * comp / tot:.1f} %.")
out.append(f"Raw list: `profiles/{tag}_launches_bench_steps2.csv`.\n")
with open(os.path.join(P, f"{tag}_launches_bench_steps2.csv"), "w") as f:
    f.write("id,kernel,grid,block,us\n")
    for r in rows:
        f.write(f'{r["ID"]},"{short(r["Kernel Name"])}","{r["Grid Size"]}","{r["Block Size"]}",{val(r) / 1e3:.2f}\n')

# ---- 2. conv launches of one step ----
rows = rows_of(os.path.join(G, "conv_step.csv"))
by = collections.OrderedDict()
for r in rows:
    d = by.setdefault(r["ID"], {"kernel": short(r["Kernel Name"])})
    d[r["Metric Name"]] = val(r) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ns": 1.0, "us": 1e3, "ms": 1e6, "%": 1.0}.get(r["Metric Unit"], 1.0)
L = list(by.values())
t_us = sum(d["gpu__time_duration.sum"] for d in L) / 1e3
rd = sum(d["dram__bytes_read.sum"] for d in L)
wr = sum(d["dram__bytes_write.sum"] for d in L)
tp = sum(d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"] * d["gpu__time_duration.sum"] for d in L) / (t_us * 1e3)
out.append(f"## 2. All {len(L)} convolution launches of one timed step (`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active…`, one pass)\n")
out.append(f"* {len(L)} launches, {t_us / 1e3:.2f} ms in total (bench `segments_ms.network` for 16 frames is the live figure)")
out.append(f"* DRAM traffic: {rd / 1e9:.2f} GB read + {wr / 1e9:.2f} GB written = {(rd + wr) / 1e9:.2f} GB per step = {(rd + wr) / 16e9:.3f} GB per frame; "
           f"average {(rd + wr) / len(L) / 1e6:.0f} MB per launch (this is `roofline.traffic` in bench.py)")
out.append(f"* average DRAM rate {(rd + wr) / (t_us * 1e-6) / 1e12:.2f} TB/s; time-weighted tensor-pipe active {tp:.1f} %\n")
out.append("| # | kernel | us | DRAM read MB | DRAM write MB | tensor pipe active % |\n|---|---|---:|---:|---:|---:|")
for i, d in enumerate(L):
    out.append(f"| {i} | `{d['kernel']}` | {d['gpu__time_duration.sum'] / 1e3:.1f} | {d['dram__bytes_read.sum'] / 1e6:.1f} | "
               f"{d['dram__bytes_write.sum'] / 1e6:.1f} | {d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']:.1f} |")
out.append(f"\nRaw: `profiles/{tag}_conv_step_dram.csv`.\n")
with open(os.path.join(P, f"{tag}_conv_step_dram.csv"), "w") as f:
    f.write(f"# ncu capture of the {len(L)} convolution launches of one timed step of `python bench.py --no-extras --steps 2 --warmup 3`, precision={PREC}\n")
    f.write("launch,kernel,us,dram_read_bytes,dram_write_bytes,tensor_pipe_active_pct\n")
    for i, d in enumerate(L):
        f.write(f'{i},"{d["kernel"]}",{d["gpu__time_duration.sum"] / 1e3:.2f},{d["dram__bytes_read.sum"]:.0f},{d["dram__bytes_write.sum"]:.0f},'
                f'{d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]:.2f}\n')
print(f"traffic per launch: {(rd + wr) / len(L):.4g} bytes; per step {(rd + wr) / 1e9:.2f} GB; launches {len(L)}")

# ---- 3. --set full key metrics ----
rep = os.path.join(G, "prof_conv.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader([l for l in raw.splitlines() if l.startswith('"')]))
    hdr, units, body = rr[0], rr[1], rr[2:]
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
            "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "smsp__warps_active.avg.per_cycle_active"]
    idx = [hdr.index(k) for k in keys if k in hdr]
    kn = hdr.index("Kernel Name")
    with open(os.path.join(P, f"{tag}_conv_igemm_ncu_full_key_metrics.csv"), "w") as f:
        f.write("launch,kernel," + ",".join(hdr[i] for i in idx) + "\n")
        for n, b in enumerate(body):
            f.write(f'{n},"{short(b[kn])}",' + ",".join(f"{b[i]} {units[i]}" for i in idx) + "\n")
    out.append(f"## 3. `ncu --set full --import-source on` on {len(body)} consecutive convolution launches\n")
    out.append(f"Key metrics per launch: `profiles/{tag}_conv_igemm_ncu_full_key_metrics.csv` (the .ncu-rep itself is not committed).  "
               "The capture starts at convolution launch 0 of the first timed step (`-s 3*NCONV` skips the warm-up steps), so launch "
               "n here is row n of section 2: stem, layer1 (three chained conv3 launches), layer2, the first bottlenecks of layer3.\n")

md = os.path.join(P, f"{tag}_ncu_summary.md")
keep = ""
if os.path.exists(md):
    old = open(md).read()
    m = re.search(r"^## 4\..*", old, re.S | re.M)
    keep = m.group(0) if m else ""
open(md, "w").write("\n".join(out) + "\n" + keep)
print("wrote", md)
