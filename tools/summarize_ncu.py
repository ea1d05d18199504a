"""Condense an `ncu --csv` metrics log (one row per kernel x metric) into a per-launch table and a per-kernel summary.

    python tools/summarize_ncu.py <ncu.csv> <out_prefix>      -> <out_prefix>_launches.csv, <out_prefix>_summary.json"""
import csv
import json
import re
import sys
from collections import OrderedDict, defaultdict


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = name.replace("uavsal::", "")
    m = re.match(r"([\w:]+)(<[^>]*>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else name[:60]


def main():
    src, prefix = sys.argv[1], sys.argv[2]
    rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ik, im, iv, iu, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
    ig, ib = hdr.index("Grid Size"), hdr.index("Block Size")
    launches = OrderedDict()
    for r in rows[1:]:
        d = launches.setdefault(r[ii], {"kernel": short(r[ik]), "grid": r[ig], "block": r[ib]})
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        unit = r[iu]
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        d[r[im]] = v * scale
    with open(prefix + "_launches.csv", "w") as fh:
        fh.write("id,kernel,grid,block,duration_us,dram_read_bytes,dram_write_bytes\n")
        for i, d in launches.items():
            fh.write("%s,\"%s\",\"%s\",\"%s\",%.2f,%.0f,%.0f\n" % (i, d["kernel"], d["grid"], d["block"], d.get("gpu__time_duration.sum", 0),
                                                              d.get("dram__bytes_read.sum", 0), d.get("dram__bytes_write.sum", 0)))
    agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for d in launches.values():
        a = agg[d["kernel"]]
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0)
        a[2] += d.get("dram__bytes_read.sum", 0)
        a[3] += d.get("dram__bytes_write.sum", 0)
    tot = sum(a[1] for a in agg.values()) or 1.0
    out = {"source": src, "total_us": round(tot, 1), "launches": len(launches), "kernels": []}
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out["kernels"].append({"kernel": k, "launches": a[0], "us": round(a[1], 1), "share": round(a[1] / tot, 4),
                               "dram_read_MB": round(a[2] / 1e6, 1), "dram_write_MB": round(a[3] / 1e6, 1),
                               "dram_GBps": round((a[2] + a[3]) / a[1] / 1e3, 1) if a[1] else None})
    json.dump(out, open(prefix + "_summary.json", "w"), indent=1)
    for k in out["kernels"][:12]:
        print(k)


if __name__ == "__main__":
    main()
