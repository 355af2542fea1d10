"""Opcode counts of the shipped library (cuobjdump -sass): which kernels hold tensor-core / TMEM / TMA / packed-FP32 / FP64
instructions.  `python scripts/sass_summary.py > profiles/sass_summary.txt` (no GPU needed)."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "amp-sparc-spatialmodulation_b200", "csrc", "libampsm_b200.so")
WATCH = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKPF", "FFMA2", "FMUL2", "DFMA", "CREDUX", "REDUX", "SYNCS",
         "LDGSTS", "UTCATOMSWS", "MUFU", "SHFL"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
tot = collections.Counter()
per = collections.defaultdict(collections.Counter)
fn = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and fn:
        op = m.group(1)
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                tot[w] += 1
                per[fn][w] += 1
print("SASS opcode counts of libampsm_b200.so (cuobjdump -sass, sm_100a), round 2b\n")
for w in WATCH:
    print(f"{w:12s} {tot[w]}")
print("\nkernels holding tensor-core / TMEM / TMA / packed-FP32 / FP64 instructions:")
for f in sorted(per):
    c = per[f]
    keys = [w for w in ("UTCHMMA", "LDTM", "UTCBAR", "UTMALDG", "UBLKCP", "FFMA2", "DFMA", "CREDUX") if c[w]]
    if keys:
        name = re.sub(r"^_ZN5ampsm\d*", "", f)
        print("  " + name[:110] + ": " + ", ".join(f"{w} {c[w]}" for w in keys))
