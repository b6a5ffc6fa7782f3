"""Aggregate an ncu source page per CUDA-C line: python tools/ncu_lines.py report.ncu-rep [top]
Prints the lines with the most stall samples / executed warp instructions (needs -lineinfo and --import-source on)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
data = []
fname = ""
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    if r and r[0] == "Line No":
        hdr = r
        csamp, cinst, cthr = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
        continue
    if hdr is None or not r or not r[0].isdigit():
        continue
    try:
        data.append((int(r[csamp]), int(r[cinst]), int(r[cthr]), fname, int(r[0]), r[1].strip()))
    except ValueError:
        pass
tot_s = sum(d[0] for d in data)
tot_i = sum(d[1] for d in data)
print(f"total samples {tot_s}, warp instructions {tot_i}")
for s, i, t, fn, ln, src in sorted(data, reverse=True)[:top]:
    print(f"{100.0 * s / max(tot_s, 1):5.1f}% samp {100.0 * i / max(tot_i, 1):5.1f}% inst thr/inst {t / max(i, 1):4.1f} {fn}:{ln}: {src[:100]}")
