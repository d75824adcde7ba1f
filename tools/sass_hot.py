"""Hot source lines per kernel from an ncu SASS-page CSV holding several kernels
(ncu -i X.ncu-rep --page source --csv --print-source=sass --kernel-name regex:...), attributed to CUDA source lines
with the `nvdisasm -g -c` listing of the SAME cubin (cuobjdump -xelf all lib.so; nvdisasm -g -c x.cubin > dis.txt).
usage: sass_hot.py <sass.csv> <dis.txt> [kernel-substring] [top-N] [focus-file:lo-hi]
Lines are the innermost frame; with a focus (and a `nvdisasm -gi` listing) an instruction is attributed to the frame of its
inline chain that lies in focus-file within lines lo..hi (e.g. glcm_body.h:226-415 = the body of glcm_angle)."""
import csv, re, subprocess, sys
from collections import defaultdict

sass_csv, dis = sys.argv[1:3]
want = sys.argv[3] if len(sys.argv) > 3 else ""
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 30
focus = None
if len(sys.argv) > 5:
    ff, rr = sys.argv[5].split(":")
    focus = (ff, int(rr.split("-")[0]), int(rr.split("-")[1]))

# ---- nvdisasm: per function, per instruction (in order) the innermost "file:line" and the inline chain's outermost line
funcs = {}
cur_fn, chain, chain_open = None, [], False
for l in open(dis):
    if l.startswith(".text."):
        cur_fn = l.strip()[len(".text."):].rstrip(":")
        funcs[cur_fn] = []
        chain, chain_open = [], False
        continue
    if cur_fn is None:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        fr = (m.group(1).split("/")[-1], int(m.group(2)))
        if chain_open:
            chain.append(fr)
        else:
            chain = [fr]
            chain_open = True
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
        chain_open = False
        loc = chain[0] if chain else None
        if focus and chain:
            hit = [f for f in chain if f[0] == focus[0] and focus[1] <= f[1] <= focus[2]]
            loc = hit[-1] if hit else ("other", 0)
        funcs[cur_fn].append("%s:%s" % loc if loc else None)
names = list(funcs)
dem = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.split("\n")
by_dem = {}
for n, d in zip(names, dem):
    by_dem[n] = d.replace("void ", "").split("<")[0].split("(")[0].strip()

# ---- the CSV: tables separated by "Kernel Name" rows
rows = list(csv.reader(open(sass_csv)))
tables, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        tables.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
seen = set()
for t in tables:
    key = re.sub(r"\s+", "", t["name"].split("(const")[0].split("(nyx::")[0].replace("void ", ""))
    short = t["name"].split("(")[0].replace("void ", "")
    if want not in t["name"] or not t["rows"]:
        continue
    hdr = t["hdr"]
    i_s, i_ie, i_src = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
    tot_i = sum(int(r[i_ie] or 0) for r in t["rows"] if len(r) > i_ie)
    if (key, tot_i) in seen:
        continue
    seen.add((key, tot_i))
    # match the disassembly function by demangled prefix (template args spelled differently: compare without them)
    base = t["name"].replace("void ", "").split("<")[0].split("(")[0].strip()
    cand = [n for n, d in by_dem.items() if d == base and len(funcs[n]) == len(t["rows"])]
    locs = funcs[cand[0]] if cand else None
    acc = defaultdict(lambda: [0, 0])
    ops = defaultdict(lambda: [0, 0])
    tot_s = 0
    for k, r in enumerate(t["rows"]):
        try:
            s, ie = int(r[i_s] or 0), int(r[i_ie] or 0)
        except ValueError:
            continue
        loc = locs[k] if locs and k < len(locs) else "?"
        acc[loc][0] += s
        acc[loc][1] += ie
        op = r[i_src].split()[1] if r[i_src].startswith("@") and len(r[i_src].split()) > 1 else r[i_src].split()[0] if r[i_src].split() else "?"
        op = op.split(".")[0]
        ops[op][0] += s
        ops[op][1] += ie
        tot_s += s
    print("== %s  (%d SASS instructions, matched %s)  samples %d  warp-inst %d" % (t["name"][:90], len(t["rows"]), bool(locs), tot_s, tot_i))
    for loc, (s, ie) in sorted(acc.items(), key=lambda x: -x[1][0])[:topn]:
        print("   %-26s samples %5.1f%%  inst %5.1f%%" % (loc, 100.0 * s / max(tot_s, 1), 100.0 * ie / max(tot_i, 1)))
    print("   -- opcodes: " + "  ".join("%s %.1f%%/%.1f%%" % (o, 100.0 * v[0] / max(tot_s, 1), 100.0 * v[1] / max(tot_i, 1))
                                          for o, v in sorted(ops.items(), key=lambda x: -x[1][1])[:14]))
