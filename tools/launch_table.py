#!/usr/bin/env python
"""Per-kernel table of one training step from an `ncu --metrics gpu__time_duration.sum --csv` launch list.

usage: python tools/launch_table.py gpurun_out/launches_bench_tf32.csv [out.json]
One step = the launches between the last two `nchw_to_nhwc_pad` kernels (one per forward)."""
import collections
import csv
import json
import re
import sys


def main(path, out=None):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    names = [r[idx["Kernel Name"]] for r in data]
    unit = data[0][idx["Metric Unit"]]
    sc = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1e-3)
    vals = [float(r[idx["Metric Value"]].replace(",", "")) * sc for r in data]
    marks = [i for i, n in enumerate(names) if "nchw_to_nhwc_pad" in n]
    a, b = marks[-2], marks[-1]
    agg = collections.OrderedDict()
    for n, v in zip(names[a:b], vals[a:b]):
        k = re.sub(r"\(.*", "", n).replace("void ", "")[:70]
        d = agg.setdefault(k, [0, 0.0])
        d[0] += 1
        d[1] += v
    tot = sum(v for _, v in agg.values())
    table = [{"kernel": k, "launches_per_step": c, "us_per_step": round(v, 1), "us_per_launch": round(v / c, 1),
              "share": round(v / tot, 4)} for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])]
    res = {"source": path, "note": "ncu per-launch times are serialised and cold-cache: shares, not absolutes",
           "us_per_step": round(tot, 1), "launches_per_step": b - a, "kernels": table}
    if out:
        json.dump(res, open(out, "w"), indent=1)
    print("total us %.1f  launches %d" % (tot, b - a))
    for t in table[:40]:
        print("%-70s %4d %9.1f %8.1f %5.1f%%" % (t["kernel"], t["launches_per_step"], t["us_per_step"], t["us_per_launch"],
                                               100 * t["share"]))


if __name__ == "__main__":
    main(*sys.argv[1:3])
