#!/usr/bin/env python
"""Per-SASS-instruction view of an ncu report: executed count, average active threads, stall samples.

    python tools/ncu_sass.py report.ncu-rep [--min-exec N] > profiles/xxx_sass.txt
"""
import csv
import io
import subprocess
import sys

path = sys.argv[1]
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
tot_exec = sum(int(r[col["Instructions Executed"]]) for r in rows[2:] if len(r) > 5)
tot_samp = sum(int(r[col["# Samples"]]) for r in rows[2:] if len(r) > 5)
print(f"# {rows[0][1]}")
print(f"# total warp-instructions {tot_exec}, stall samples {tot_samp}")
print(f"{'idx':>5s} {'exec':>10s} {'%exec':>6s} {'thr':>5s} {'samp':>7s} {'%samp':>6s} {'long_sb':>7s} {'short':>6s} {'math':>5s} {'wait':>5s} {'notsel':>6s}  sass")
for i, r in enumerate(rows[2:]):
    if len(r) < 6:
        continue
    ex = int(r[col["Instructions Executed"]])
    sm = int(r[col["# Samples"]])
    g = lambda k: r[col[k]] if k in col else ""
    print(f"{i:5d} {ex:10d} {100 * ex / max(tot_exec, 1):6.2f} {g('Avg. Threads Executed'):>5s} {sm:7d} {100 * sm / max(tot_samp, 1):6.2f} "
          f"{g('stall_long_sb'):>7s} {g('stall_short_sb'):>6s} {g('stall_math'):>5s} {g('stall_wait'):>5s} {g('stall_not_selected'):>6s}  {r[col['Source']].strip()}")
