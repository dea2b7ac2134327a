"""Turns ncu outputs into the small markdown summaries kept under profiles/.

  python scripts/summarize_ncu.py launches gpurun_out/launches_r1.csv          > profiles/r01_launches.md
  python scripts/summarize_ncu.py report   gpurun_out/prof_gemm_r1b.ncu-rep    > profiles/r01_gemm_ncu.md
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    total = 0.0
    n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")[:80]
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1.0)
        d = agg.setdefault(name, [0, 0.0])
        d[0] += 1
        d[1] += v
        total += v
        n += 1
    print(f"ncu launch list `{path}`: {n} launches, {total/1e3:.2f} ms of kernel time "
          "(cold-cache, serialised: compare shares, not absolutes)\n")
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if t / total < 0.0005:
            continue
        print(f"| `{k}` | {c} | {t:.1f} | {100*t/total:.1f}% |")


def report(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"### `{name[:100]}`  (from `{path}`)\n\n| metric | value | unit |\n|---|---:|---|")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"| {k} | {vals[i]} | {units[i]} |")
        print()


def traffic(path):
    """DRAM bytes per launch of the mlp_fused_kernel launches in a capture -> JSON for bench.py's roofline.traffic
    (python scripts/summarize_ncu.py traffic gpurun_out/X.ncu-rep profiles/X.md > profiles/fused_traffic.json)."""
    import json
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    per = []
    for vals in rows[2:]:
        if "mlp_fused_kernel" not in vals[hdr.index("Kernel Name")]:
            continue
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(k)
            tot += float(vals[i].replace(",", "")) * scale[units[i]]
        per.append(tot)
    print(json.dumps({"dram_bytes_per_launch": round(sum(per) / len(per)), "launches": len(per),
                      "dram_bytes": [round(x) for x in per],
                      "source": sys.argv[3] if len(sys.argv) > 3 else path}, indent=1))


if __name__ == "__main__":
    {"launches": launches, "report": report, "traffic": traffic}[sys.argv[1]](sys.argv[2])
