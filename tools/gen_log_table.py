"""Regenerates autofrk_b200/csrc/tps_log_table.h (128-entry reciprocal / log table, 60-digit mpmath)."""
import mpmath as mp

mp.mp.dps = 60
rows = []
for j in range(128):
    c = mp.mpf(1) + (mp.mpf(j) + mp.mpf("0.5")) / 128
    rc = float(1 / c)
    lc = float(-mp.log(mp.mpf(rc)))
    rows.append("    {%s, %s}," % (rc.hex(), lc.hex()))
print("\n".join(rows))
