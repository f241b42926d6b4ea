"""Summarise an .ncu-rep (raw + source pages) into a small text file for profiles/.  Tooling."""
import csv
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warp_latency_per_inst_issued.ratio"]
for r in rows[2:]:
    print("=" * 100)
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:95s} {r[i]} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
# one block per profiled launch: a "Kernel Name" row, a header row, then the SASS lines
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1] if len(r) > 1 else "", "h": None, "data": []}
        blocks.append(cur)
    elif cur is not None and cur["h"] is None:
        cur["h"] = r
    elif cur is not None and len(r) == len(cur["h"]):
        cur["data"].append(r)
for bi, b in enumerate(blocks):
    h, data = b["h"], b["data"]
    idx = {n: i for i, n in enumerate(h)}
    tot = sum(int(r[idx["# Samples"]]) for r in data) or 1
    print(f"\nsource page, launch {bi}: {b['name'][:90]}\n{len(data)} SASS instructions, {tot} warp-state samples; top {topn} by samples")
    stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    agg = {c: sum(int(r[idx[c]] or 0) for r in data) for c in stall_cols}
    print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
    if "L1 Wavefronts Shared Excessive" in idx:
        exc = [(int(r[idx["L1 Wavefronts Shared Excessive"]] or 0), r[1].strip()[:70]) for r in data]
        exc = sorted([e for e in exc if e[0] > 0], reverse=True)[:8]
        print("shared-memory excessive wavefronts (bank conflicts) by instruction:", exc if exc else "none")
    for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:topn]:
        st = {c: int(r[idx[c]]) for c in stall_cols if r[idx[c]] not in ("", "0")}
        st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:2])
        print(f"{int(r[idx['# Samples']]):7d} {100.0 * int(r[idx['# Samples']]) / tot:5.1f}%  {r[1].strip()[:64]:64s} {st}")
