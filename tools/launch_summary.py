"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (measurement tooling).
usage: python tools/launch_summary.py launches.csv "<command that was profiled>" """
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rows:
    if r is hdr or r[ik] == "Kernel Name":
        continue
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}.get(r[iu], 1e-6)
    name = re.sub(r"\(.*", "", r[ik])[:110]
    tot[name] += v * scale
    cnt[name] += 1
total = sum(tot.values())
print(f"ncu launch list of `{sys.argv[2] if len(sys.argv) > 2 else '?'}` (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)")
print(f"total kernel time {total:.1f} ms over {sum(cnt.values())} launches")
ours = sum(v for k, v in tot.items() if "frk::" in k or "frk_" in k or "<unnamed>::" in k and "at::" not in k)
print(f"kernels of this repository: {ours:.1f} ms = {100 * ours / total:.1f}% of all kernel time\n")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:28]:
    print(f"{v:10.3f} ms  {100 * v / total:5.1f}%  x{cnt[k]:<5d} {k}")
