"""profiles/k1_traffic_r02.json from an `ncu --set full` capture of the bench launch of K1 (measurement tooling):
  ncu --set full --clock-control none --import-source on -k regex:mrts_predict_kernel -s 3 -c 1 -o gpurun_out/k1_r02_bench \
      python bench.py --steps 1 --warmup 3 --no-fit
  python tools/k1_traffic_capture.py gpurun_out/k1_r02_bench.ncu-rep
The JSON records the commit and a hash of the kernel sources; bench.py reports `traffic` only while they still match."""
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bench import K1_SOURCES, WORKLOAD, sources_sha256  # noqa: E402

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, r = rows[0], rows[1], rows[2]


def val(name):
    i = hdr.index(name)
    v, u = float(r[i].replace(",", "")), units[i]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(u, 1.0)
    return v * scale, u


rd, _ = val("dram__bytes_read.sum")
wr, _ = val("dram__bytes_write.sum")
dur, du = val("gpu__time_duration.sum")
w = WORKLOAD
out = {
    "kernel": r[hdr.index("Kernel Name")],
    "launch": f"bench.py --steps 1 --warmup 3 --no-fit ({w['rows_per_gpu']} rows x {w['knots']} knots, K={w['K']}), 4th launch",
    "dram_bytes_read": rd, "dram_bytes_write": wr, "traffic_bytes": rd + wr,
    "algorithmic_bytes": 8 * w["rows_per_gpu"] * (w["d"] + w["K"]),
    "duration": [dur, du],
    "dmma_pipe_active_pct": float(r[hdr.index("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active")]),
    "commit": subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip(),
    "k1_sources": list(K1_SOURCES), "k1_sources_sha256": sources_sha256(K1_SOURCES),
    "source": rep,
}
(ROOT / "profiles" / "k1_traffic_r02.json").write_text(json.dumps(out, indent=1))
print(json.dumps(out, indent=1))
