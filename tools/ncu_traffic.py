"""DRAM traffic per pipeline stage from an `ncu --set full` raw-page CSV of ONE resident configs[1] step
(tools/gpu_full.sh): dram__bytes_read.sum + dram__bytes_write.sum, summed over the launches of the stage.
usage: ncu_traffic.py <raw.csv> > profiles/rN_traffic.json"""
import csv, json, sys
import re
FAMS = {1: "glcm", 2: "glrlm", 3: "gldzm", 4: "glszm", 5: "gldm", 6: "ngldm", 7: "ngtdm"}
STAGE = [("k_scan_stats", "scan"), ("k_pack_tiles", "pack"), ("k_pack_cloud", "pack"), ("k_roi_intensity", "intensity"),
         ("k_select_columns", "select"), ("k_compact", "compact"), ("k_fill_class_lists", "compact"), ("k_init_stats", "init")]


def stage_of(name):
    m = re.search(r"k_roi_family(?:_w)?<(?:\(int\))?(\d)", name)
    if m:
        return FAMS.get(int(m.group(1)))
    for pat, st in STAGE:
        if pat in name:
            return st
    return None


rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ki, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out, launches = {}, {}
for r in rows[2:]:
    st = stage_of(r[ki])
    if st:
        b = float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
        out[st] = out.get(st, 0.0) + b
        launches[st] = launches.get(st, 0) + 1
print(json.dumps({"source": sys.argv[1].split("/")[-1] + " (ncu --set full --clock-control none, one resident configs[1] step, roi_batches=1, families serialised)",
                  "bytes_per_step": {k: int(v) for k, v in out.items()}, "launches_per_step": launches}, indent=1))
