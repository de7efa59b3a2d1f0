"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file L.csv
<command>`) into per-kernel launch counts, mean durations and shares of the summed kernel time:
    python examples/launch_shares.py L.csv out.json "<command>" [bench_line.json]
The optional bench line is the JSON line the same command printed WITHOUT ncu (its CUDA-event shares are stored beside
the launch-list shares; numbers printed under ncu are never bench values)."""
import csv, gzip, json, re, sys

src, dst, command = sys.argv[1], sys.argv[2], sys.argv[3]
op = gzip.open if src.endswith(".gz") else open
rows = [r for r in csv.reader(op(src, "rt")) if r]
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}
agg, total = {}, 0.0
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    m = re.match(r"(?:void )?(?:ax1::)?([A-Za-z0-9_]+)(<[^>]*>)?", r[kn])
    k = (m.group(1) + (m.group(2) or "")) if m else r[kn]
    us = float(r[mv].replace(",", "")) * scale.get(r[mu], 1.0)
    e = agg.setdefault(k, [0, 0.0])
    e[0] += 1
    e[1] += us
    total += us
out = {"command": command, "note": "per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes",
       "total_us": total, "launches": sum(e[0] for e in agg.values()),
       "kernels": {k: {"launches": e[0], "avg_us": e[1] / e[0], "share": e[1] / total}
                   for k, e in sorted(agg.items(), key=lambda kv: -kv[1][1])}}
if len(sys.argv) > 4:
    line = json.loads([l for l in open(sys.argv[4]) if l.startswith("{")][-1])
    out["bench_line_of_the_plain_run"] = {"ms_per_step": line["ms_per_step"], "clocks": line.get("clocks"),
                                          "share_of_step": {k: v["share_of_step"] for k, v in line["roofline"]["kernels"].items()}}
json.dump(out, open(dst, "w"), indent=1)
for k, v in list(out["kernels"].items())[:12]:
    print("%-36s %6d x %10.1f us  %5.1f %%" % (k, v["launches"], v["avg_us"], 100 * v["share"]))
