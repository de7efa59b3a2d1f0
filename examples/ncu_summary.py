"""Summarise an `ncu --page raw --csv` export (one row per profiled launch) into profiles/*.json:
    python examples/ncu_summary.py gpurun_out/r2_prof_raw.csv profiles/r02_top_kernels_ncu.json profiles/r02_traffic.json
Per kernel (mean over its launches in the capture): duration, DRAM bytes read / written, DRAM throughput % of peak,
registers, active warps %, IPC -- the evidence behind roofline.traffic in the bench line."""
import csv, json, re, sys

src, dst, dst_traffic = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(src)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {n: i for i, n in enumerate(hdr)}
want = {"gpu__time_duration.sum": "duration", "dram__bytes_read.sum": "dram_bytes_read", "dram__bytes_write.sum": "dram_bytes_write",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct", "launch__registers_per_thread": "registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct", "sm__inst_executed.avg.per_cycle_active": "ipc_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1tex_throughput_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
        "l1tex__t_sector_hit_rate.pct": "l1_hit_rate_pct", "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
        "launch__block_size": "block", "launch__grid_size": "grid", "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct"}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def val(r, name):
    if name not in col:
        return None
    try:
        v = float(r[col[name]].replace(",", ""))
    except ValueError:
        return None
    return v * scale.get(units[col[name]], 1.0)


def short(name):
    m = re.match(r"(?:void )?(?:ax1::)?([A-Za-z0-9_]+)(<[^>]*>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else name


agg = {}
for r in data:
    k = short(r[col["Kernel Name"]])
    e = agg.setdefault(k, {"launches": 0})
    e["launches"] += 1
    for n, key in want.items():
        v = val(r, n)
        if v is not None:
            e[key] = e.get(key, 0.0) + v
out = {}
for k, e in agg.items():
    n = e.pop("launches")
    o = {"launches_in_capture": n}
    for key, v in e.items():
        o[key + ("_ms" if key == "duration" else "")] = v / n
    if "dram_bytes_read" in o and "duration_ms" in o:
        o["dram_GBps"] = (o["dram_bytes_read"] + o["dram_bytes_write"]) / (o["duration_ms"] * 1e-3) / 1e9
    out[k] = o
meta = {"capture": "ncu --set full --clock-control none, python examples/prof_kernels.py 93184 128 (named workload, 16.87 M vertices, ILU0 slab 128), round 2 final kernels",
        "note": "means over the launches of each kernel in the capture; durations are under the profiler (cold cache, serialised) -- bench.py times with CUDA events"}
json.dump({**meta, "kernels": out}, open(dst, "w"), indent=1)
# what bench.py reads for roofline.traffic: bench names -> kernels
names = {"ilu_apply": "ilu_apply_kernel<2, 0>", "spmv_dot1": "spmv_compact_kernel<1>", "spmv_dot2": "spmv_compact_kernel<2>", "spmv": "spmv_kernel<0>",
         "compact": "compact_rows_kernel",
         "ilu_factor": "ilu_factor_kernel", "jacobian": "jacobian_kernel", "residual": "residual_kernel", "axpy": "xr_update_kernel"}
tr = {}
for b, k in names.items():
    cand = [kk for kk in out if kk.replace(" ", "") == k.replace(" ", "")]
    if cand:
        e = out[cand[0]]
        tr[b] = {"kernel": cand[0], "dram_bytes_read": e.get("dram_bytes_read"), "dram_bytes_write": e.get("dram_bytes_write"),
                 "duration_ms_under_ncu": e.get("duration_ms")}
json.dump({"capture": "profiles/r02_top_kernels_ncu.json", "kernels": tr}, open(dst_traffic, "w"), indent=1)
print(json.dumps(out, indent=1)[:3000])
