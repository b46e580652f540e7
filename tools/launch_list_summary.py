"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`) into the JSON kept
under profiles/: launches, total time and share of the profiled time per kernel.
Usage: python tools/launch_list_summary.py in.csv out.json "<command that was profiled>" """
import csv
import json
import re
import sys


def main():
    src, dst, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("=="))]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3, "s": 1e3}
    agg, order = {}, []
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("exb::", "").replace("signed char", "int8").strip()
        ms = float(r[vi].replace(",", "")) * scale[r[ui]]
        if name not in agg:
            agg[name] = [0, 0.0]
            order.append(name)
        agg[name][0] += 1
        agg[name][1] += ms
    total = sum(v[1] for v in agg.values())
    out = {
        "command": cmd,
        "note": "per-launch times under ncu are cold-cache and serialised: the SHARES are what must agree with the "
                "bench's event-timed stages",
        "launches": sum(v[0] for v in agg.values()),
        "total_ms": total,
        "kernels": [{"kernel": k, "launches": agg[k][0], "total_ms": agg[k][1], "share": agg[k][1] / total} for k in order],
    }
    json.dump(out, open(dst, "w"), indent=1)
    for k in out["kernels"]:
        print("%-40s %5d %10.3f ms  %.4f" % (k["kernel"], k["launches"], k["total_ms"], k["share"]))


if __name__ == "__main__":
    main()
