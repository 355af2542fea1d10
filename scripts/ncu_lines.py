"""Aggregate the warp-stall samples of an .ncu-rep (captured with --import-source on, built with -lineinfo) by CUDA source
line: `python scripts/ncu_lines.py file.ncu-rep [top]`.  Prints samples and executed instructions per (file, line)."""
import collections
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = collections.defaultdict(lambda: [0, 0])
cur_file, cur = None, None
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
    elif r[0].isdigit():
        cur = (cur_file, int(r[0]))
    elif r[0] == '' and len(r) > 7 and r[2].startswith('0x'):
        try:
            agg[cur][0] += int(r[4])
            agg[cur][1] += int(r[7])
        except ValueError:
            pass
tot = sum(v[0] for v in agg.values())
toti = sum(v[1] for v in agg.values())
print(f"total samples {tot}, warp instructions {toti}")
for k, v in sorted(agg.items(), key=lambda kv: (kv[0][0], kv[0][1])):
    if v[0] * 400 >= tot or v[1] * 400 >= toti:
        print(f"{k[0]:>22s}:{k[1]:<5d} samples {v[0]:6d} ({100 * v[0] / tot:5.1f} %)  instr {v[1]:10d} ({100 * v[1] / max(toti, 1):5.1f} %)")
