"""Brief of one-kernel ncu report: key utilisation metrics + the top stall locations of the source page.
usage: python tools/ncu_brief.py <report.ncu-rep> [n_top]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 22
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__mem_tensor_reads_op_ldt.sum.pct_of_peak_sustained_elapsed"]
for h, u, v in zip(hdr, units, vals):
    if h in keys:
        print(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; ix = {k: i for i, k in enumerate(h)}; data = rows[2:]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
print("total samples", tot, "warp instructions", sum(int(r[ix["Instructions Executed"]] or 0) for r in data))
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:ntop]:
    st = {k: int(r[ix[k]] or 0) for k in ix if k.startswith("stall_") and "Not Issued" not in k}
    st = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(r[ix["Address"]][-5:], r[ix["# Samples"]], r[ix["Instructions Executed"]], r[ix["Source"]][:72], st)
