"""Aggregate an ncu SASS source page by CUDA source line.

usage: ncu_by_line.py <report.ncu-rep> <lib.so> <mangled-kernel-name-substring> [demangled-name substring selecting the launch, e.g. "<(int)3"]

`ncu --page source --csv` gives per-SASS-instruction counters without the CUDA
line; nvdisasm -g on the same binary gives the line of every instruction.  The
two listings are joined by position.
"""
import csv
import re
import subprocess
import sys
import tempfile
from collections import defaultdict
from pathlib import Path

rep, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
tmp = Path(tempfile.mkdtemp())
subprocess.run(["cuobjdump", "-xelf", "all", str(Path(lib).resolve())], cwd=tmp, check=True, capture_output=True)
cubin = next(tmp.glob("*.cubin"))
sass = subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], capture_output=True, text=True).stdout.split("\n")
start = [k for k, l in enumerate(sass) if l.startswith(".text.") and kern in l and l.rstrip().endswith(":")][0]
cur = None
seq = []
for l in sass[start + 1:]:
    if l.startswith("//---------------------"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        seq.append((int(m.group(1), 16), cur, m.group(2)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
# one section per profiled launch, each headed by a "Kernel Name" row; keep the selected one
heads = [k for k, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
pick = sys.argv[4] if len(sys.argv) > 4 else ""
sel = [k for k in range(len(heads) - 1) if pick in rows[heads[k]][1]][0]
rows = rows[heads[sel]:heads[sel + 1]]
hdr = rows[1]
ie, te, sm = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
prof = [(r[1], float(r[ie]), float(r[te]), float(r[sm] or 0)) for r in rows[2:] if len(r) > ie]
assert len(prof) == len(seq), (len(prof), len(seq))
by = defaultdict(lambda: [0.0, 0.0, 0.0])
tot = sum(p[1] for p in prof)
tots = sum(p[3] for p in prof)
for (a, loc, txt), p in zip(seq, prof):
    key = loc if loc else ("?", 0)
    by[key][0] += p[1]
    by[key][1] += p[2]
    by[key][2] += p[3]
byfile = defaultdict(lambda: [0.0, 0.0, 0.0])
for k, v in by.items():
    for q in range(3):
        byfile[k[0]][q] += v[q]
print("SASS instructions", len(seq), "total warp inst", tot, "samples", tots)
for k, v in sorted(byfile.items(), key=lambda x: -x[1][0]):
    print(f"{k:20s} inst {v[0] / tot * 100:5.1f}%  thr/inst {v[1] / max(v[0], 1):5.1f}  samples {v[2] / tots * 100:5.1f}%")
print()
for k, v in sorted(by.items(), key=lambda x: -x[1][2])[:50]:
    print(f"{k[0]:16s}:{k[1]:4d} inst {v[0] / tot * 100:5.2f}% thr/inst {v[1] / max(v[0], 1):5.1f} samples {v[2] / tots * 100:5.2f}%")
op = defaultdict(float)
for (a, loc, txt), p in zip(seq, prof):
    t = txt.split()
    o = t[1] if t[0].startswith("@") else t[0]
    op[o.split(".")[0]] += p[1]
print()
print("opcode mix (% of warp instructions):")
print(", ".join(f"{k} {v / tot * 100:.1f}" for k, v in sorted(op.items(), key=lambda x: -x[1])[:28]))

if len(sys.argv) > 5:  # top SASS instructions by samples
    print("\ntop SASS instructions by samples:")
    order = sorted(range(len(seq)), key=lambda k: -prof[k][3])[: int(sys.argv[5])]
    for k in order:
        a, loc, txt = seq[k]
        print(f"{a:6x} {str(loc):28s} samples {prof[k][3] / tots * 100:5.2f}% exec {prof[k][1]:12.0f} thr/inst {prof[k][2] / max(prof[k][1], 1):5.1f}  {txt}")
