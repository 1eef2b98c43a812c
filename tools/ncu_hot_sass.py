"""Top sampled SASS instructions of one kernel in an .ncu-rep (needs --set full --import-source on):
   python tools/ncu_hot_sass.py <report.ncu-rep> <kernel-id regex:N> [top]"""
import csv, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", kid], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
hdr = rows[hi]
body = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0] != "Address"]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
tot = sum(int(r[isamp]) for r in body)
print(rows[0][1][:100] if rows and len(rows[0]) > 1 else "", "| samples", tot, "| SASS instructions", len(body))
order = {id(r): i for i, r in enumerate(body)}
for r in sorted(body, key=lambda r: -int(r[isamp]))[:top_n]:
    print(f"{int(r[isamp]) * 100 / tot:5.1f}%  #{order[id(r)]:5d}  exec {r[iex]:>9s}  {r[isrc].strip()[:100]}")
