#!/usr/bin/env python
"""Turns an ncu capture into an entry of profiles/r02_traffic.json (what bench.py prints as roofline.traffic).

  python scripts/ncu_traffic.py --key trace_headline --csv gpurun_out/x.csv --kernel 'trace_g2_kernel' --units 16777216 [--note ...]
  python scripts/ncu_traffic.py --key render_cfg5_extend --csv gpurun_out/y.csv --kernel 'extend_' --units <rays of the captured frame>

--csv is the `--csv --log-file` output of an ncu run that collected dram__bytes_read.sum and dram__bytes_write.sum (e.g.
`--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct`, or the raw page
of a --set full report: `ncu -i rep --page raw --csv`).  All launches whose kernel name matches are summed; `units` is
what they processed together (rays).  The entry records the sha256 of the kernel sources (bench.kernel_source_hash): bench.py
refuses to print a traffic figure whose sources have changed since."""
from __future__ import annotations

import argparse
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def parse(path, kernel_re):
    rows = list(csv.reader(l for l in open(path, errors="replace") if l.startswith('"')))
    hdr = rows[0]
    tot = {"dram__bytes_read.sum": 0.0, "dram__bytes_write.sum": 0.0, "gpu__time_duration.sum": 0.0}
    launches = 0
    hits = []
    if "Metric Name" in hdr:  # long format (--metrics ... --csv)
        ki, mi, ui, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value"), hdr.index("ID")
        seen = set()
        for r in rows[1:]:
            if not re.search(kernel_re, r[ki]):
                continue
            v = float(r[vi].replace(",", ""))
            u = r[ui].lower()
            scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "usecond": 1e-3, "nsecond": 1e-6, "msecond": 1.0}.get(u, 1)
            if r[mi] in tot:
                tot[r[mi]] += v * scale
            if r[mi] == "lts__t_sector_hit_rate.pct":
                hits.append(v)
            if r[ii] not in seen:
                seen.add(r[ii]); launches += 1
    else:  # wide format (--page raw --csv): second row holds the units
        units = rows[1]
        ki = hdr.index("Kernel Name")
        for r in rows[2:]:
            if not re.search(kernel_re, r[ki]):
                continue
            launches += 1
            for m in tot:
                if m in hdr:
                    c = hdr.index(m)
                    u = units[c].lower()
                    scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "usecond": 1e-3, "nsecond": 1e-6, "msecond": 1.0}.get(u, 1)
                    tot[m] += float(r[c].replace(",", "")) * scale
            if "lts__t_sector_hit_rate.pct" in hdr:
                hits.append(float(r[hdr.index("lts__t_sector_hit_rate.pct")]))
    return tot, launches, hits


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--key", required=True)
    ap.add_argument("--csv", required=True)
    ap.add_argument("--kernel", required=True)
    ap.add_argument("--units", type=float, required=True)
    ap.add_argument("--note", default="")
    args = ap.parse_args()
    import bench
    tot, launches, hits = parse(args.csv, args.kernel)
    if launches == 0:
        raise SystemExit("no launch matches " + args.kernel)
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    db = json.load(open(p)) if os.path.exists(p) else {}
    db[args.key] = {"kernel": args.kernel, "launches": launches, "units_per_launch": args.units, "dram_bytes_read": tot["dram__bytes_read.sum"],
                    "dram_bytes_write": tot["dram__bytes_write.sum"], "gpu_time_ms_under_ncu": tot["gpu__time_duration.sum"],
                    "dram_bytes_per_unit": (tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"]) / args.units,
                    "lts_hit_rate_pct_mean": (sum(hits) / len(hits)) if hits else None,
                    "source": f"{os.path.basename(args.csv)} (ncu --clock-control none; launches summed: {launches}) {args.note}".strip(),
                    "source_sha256": bench.kernel_source_hash(args.key)}
    json.dump(db, open(p, "w"), indent=1)
    print(json.dumps(db[args.key], indent=1))


if __name__ == "__main__":
    main()
