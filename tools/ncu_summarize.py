"""Turns ncu CSV exports into the text summaries kept under profiles/.

    python tools/ncu_summarize.py launches <launches.csv> "<command line>"      > profiles/rN_launches_*.txt
    python tools/ncu_summarize.py raw <raw.csv> "<title>"                        > profiles/rN_ncu_kernels.txt

launches.csv: `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv <cmd>`
raw.csv:      `ncu -i report.ncu-rep --page raw --csv > raw.csv`
"""
import collections
import csv
import sys

RAW_METRICS = [
    "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__cluster_size", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__shared_mem_per_block_dynamic",
]


def launches(path, command):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", "")) / 1000.0
    ours = sum(t for k, (c, t) in agg.items() if "coattn::" in k)
    print(f"# ncu launch list of `{command}`")
    print("# gpu__time_duration.sum, --clock-control none; cold-cache, serialised launches: compare SHARES, not absolutes")
    print("# kernel | launches | total us | avg us | share of coattn kernels")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        share = f"{100 * t / ours:.1f}%" if "coattn::" in k else "-"
        print(f"{k[:90]} | {c} | {t:.1f} | {t / c:.1f} | {share}")


def raw(path, title):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    print(f"# {title}")
    seen = set()
    for r in rows[2:]:
        name = r[ki].split("(")[0]
        if name in seen:
            continue
        seen.add(name)
        print(f"\n## {name}")
        for m in RAW_METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"{m} = {r[i]} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2], sys.argv[3])
