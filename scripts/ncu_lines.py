#!/usr/bin/env python
"""Per-source-line executed warp instructions of a kernel from an `ncu --set full --import-source on`
report:  python scripts/ncu_lines.py report.ncu-rep [units]  (units = what to divide by, e.g. centres)."""
import csv
import io
import subprocess
import sys
from collections import defaultdict

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
counts, text, current, column, stall = defaultdict(int), {}, None, None, defaultdict(int)
for row in csv.reader(io.StringIO(raw)):
    if not row:
        continue
    if row[0] == "File Path":
        current = row[1].split("/")[-1]
    elif row[0] == "Line No":
        column = row.index("Instructions Executed")
        samples = row.index("# Samples")
    elif row[0].isdigit() and column is not None and len(row) > column:
        key = (current, int(row[0]))
        text[key] = row[1].strip()
        try:
            counts[key] += int(row[column])
            stall[key] += int(row[samples])
        except ValueError:
            pass
total = sum(counts.values())
all_samples = max(1, sum(stall.values()))
print(f"total {total / 1e9:.2f} G warp instructions, {total / units:.1f} per unit")
for key, n in sorted(counts.items(), key=lambda kv: -kv[1])[:70]:
    print(f"{n / units:9.1f} {100.0 * n / total:5.1f}%  stall {100.0 * stall[key] / all_samples:5.1f}%  {key[0]}:{key[1]:<4d} {text[key][:110]}")
