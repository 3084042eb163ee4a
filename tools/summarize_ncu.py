"""Turn gpurun_out ncu artefacts into the committed summaries under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches_r1.csv profiles/launches_r1.md
    python tools/summarize_ncu.py report   gpurun_out/prof_attn_r1.ncu-rep profiles/prof_attn_r1.md
    python tools/summarize_ncu.py traffic  gpurun_out/prof_cross_r1.ncu-rep profiles/traffic_r1.json cross_attn_kernel large-v3 32
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "launch__occupancy_limit_shared_mem",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]


def short(name: str) -> str:
    name = re.sub(r"\(.*", "", name)
    name = name.replace("wl::", "")
    return name[:70]


def launches(src, dst):
    rows = []
    with open(src, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(io.StringIO("".join(lines)))
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        rows.append((r["Kernel Name"], ns))
    agg = defaultdict(lambda: [0, 0.0])
    for k, ns in rows:
        agg[short(k)][0] += 1
        agg[short(k)][1] += ns
    total = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` -- per-launch times are cold-cache and "
                "serialised: compare SHARES, not absolutes.\n\n")
        f.write(f"{len(rows)} launches, {total / 1e6:.2f} ms summed kernel time\n\n| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {ns / 1e6:.3f} | {100 * ns / total:.1f}% | {ns / n / 1e3:.1f} |\n")
    print(open(dst).read())


def traffic(src, dst, kernel, model, streams):
    """Per-launch DRAM bytes (read + write) of `kernel` from a --set full capture -> JSON read by bench.py."""
    import json
    import os
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rd[0], rd[1], rd[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    vals, times = [], []
    for row in data:
        if kernel not in row[idx["Kernel Name"]]:
            continue
        rb = float(row[idx["dram__bytes_read.sum"]].replace(",", "")) * mult.get(units[idx["dram__bytes_read.sum"]], 1)
        wb = float(row[idx["dram__bytes_write.sum"]].replace(",", "")) * mult.get(units[idx["dram__bytes_write.sum"]], 1)
        vals.append(rb + wb)
        t = float(row[idx["gpu__time_duration.sum"]].replace(",", ""))
        times.append(t * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[idx["gpu__time_duration.sum"]], 1e-3))
    doc = {}
    if os.path.exists(dst):
        doc = json.load(open(dst))
    doc[kernel] = {"dram_bytes_per_launch": sum(vals) / len(vals), "launches": len(vals), "profiled_launch_us": sum(times) / len(times),
                   "model": model, "streams": int(streams), "source": os.path.basename(src),
                   "how": "ncu --set full --clock-control none: dram__bytes_read.sum + dram__bytes_write.sum, mean over the captured launches"}
    json.dump(doc, open(dst, "w"), indent=1)
    print(json.dumps(doc[kernel]))


def report(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rd[0], rd[1], rd[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\n")
        for row in data:
            f.write(f"## `{short(row[idx['Kernel Name']])}`  (id {row[idx['ID']]})\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in idx:
                    f.write(f"| {k} | {row[idx[k]]} | {units[idx[k]]} |\n")
            try:
                t = float(row[idx["gpu__time_duration.sum"]].replace(",", ""))
                tu = units[idx["gpu__time_duration.sum"]]
                t_s = t * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(tu, 1e-9)
                rb = float(row[idx["dram__bytes_read.sum"]].replace(",", ""))
                wb = float(row[idx["dram__bytes_write.sum"]].replace(",", ""))
                mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                rb *= mult.get(units[idx["dram__bytes_read.sum"]], 1)
                wb *= mult.get(units[idx["dram__bytes_write.sum"]], 1)
                f.write(f"| **DRAM traffic** | {(rb + wb) / 1e6:.2f} | MB |\n| **DRAM GB/s (this profiled launch)** | {(rb + wb) / t_s / 1e9:.1f} | GB/s |\n")
            except Exception as e:  # noqa
                f.write(f"| (derived metrics unavailable: {e}) | | |\n")
            f.write("\n")
    print(open(dst).read()[:6000])


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "traffic":
    traffic(*sys.argv[2:7])
    sys.exit(0)

if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2], sys.argv[3])
