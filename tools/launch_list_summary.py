"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv ...`)
into per-kernel launch counts, time and share of the captured time:   python tools/launch_list_summary.py X.csv out.json "<command>" ["note"]"""
import csv, json, re, sys


def main(path, out, command, note=""):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    name, val, unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    kernels, total = {}, 0.0
    for r in rows[1:]:
        try:
            t = float(r[val].replace(",", ""))
        except ValueError:
            continue
        ms = t * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[unit], 1e-6)
        k = re.sub(r"\(.*$", "", r[name]).strip()
        e = kernels.setdefault(k, {"launches": 0, "ms": 0.0})
        e["launches"] += 1
        e["ms"] += ms
        total += ms
    for e in kernels.values():
        e["share"] = e["ms"] / total
    kernels = dict(sorted(kernels.items(), key=lambda kv: -kv[1]["ms"]))
    json.dump({"command": command, "note": note, "launches_captured": sum(e["launches"] for e in kernels.values()), "total_ms": total,
               "kernels": kernels}, open(out, "w"), indent=1)
    for k, e in list(kernels.items())[:6]:
        print("%-90s %5d %12.3f ms %.6f" % (k[:90], e["launches"], e["ms"], e["share"]))


if __name__ == "__main__":
    main(*sys.argv[1:5])
