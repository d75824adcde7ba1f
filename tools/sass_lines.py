"""Attribute an ncu SASS-page CSV (ncu -i X.ncu-rep --page source --csv --print-source=sass -k ...) to CUDA source lines
using nvdisasm -g line info of the same cubin.  Usage: sass_lines.py <sass.csv> <nvdisasm.txt> <kernel-mangled-substring>"""
import csv, re, sys
from collections import defaultdict
sass_csv, dis, key = sys.argv[1:4]
# nvdisasm: find the function section
lines = open(dis).read().split('\n')
start = None
for i, l in enumerate(lines):
    if l.startswith('.text.') and key in l:
        start = i
        break
assert start is not None, "kernel not found"
cur = None
inst_line = []   # per instruction in order: (file:line)
for l in lines[start + 1:]:
    if l.startswith('.text.') or l.startswith('.section'):
        if inst_line:
            break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if re.match(r'\s*/\*[0-9a-f]{4,}\*/', l):
        inst_line.append(cur)
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
i_s = hdr.index('# Samples'); i_ie = hdr.index('Instructions Executed')
acc = defaultdict(lambda: [0, 0])
k = 0
tot_s = tot_i = 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        s = int(r[i_s]); ie = int(r[i_ie])
    except ValueError:
        continue
    loc = inst_line[k] if k < len(inst_line) else None
    k += 1
    acc[loc][0] += s; acc[loc][1] += ie
    tot_s += s; tot_i += ie
print("instructions in csv", k, "in disasm", len(inst_line), "samples", tot_s, "warp-inst", tot_i)
for loc, (s, ie) in sorted(acc.items(), key=lambda x: -x[1][0])[:int(sys.argv[4]) if len(sys.argv) > 4 else 40]:
    print("%-28s samples %5.1f%%  inst %5.1f%%" % ("%s:%s" % loc if loc else "?", 100.0 * s / tot_s, 100.0 * ie / tot_i))
