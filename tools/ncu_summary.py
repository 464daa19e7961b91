"""Turn ncu reports / launch lists from gpurun_out/ into the committed summaries
under profiles/.

    python tools/ncu_summary.py <tag> [report.ncu-rep ...] [--launches launches.csv]

Writes profiles/<tag>_kernels.csv (one row per captured launch: duration, DRAM
bytes, DRAM %, SM %, issue-slot %, registers, occupancy, warp instructions),
profiles/<tag>_launches.csv (kernel name, launches, total / mean device time and
share, from the `--metrics gpu__time_duration.sum` pass) and updates
profiles/traffic.json (DRAM bytes per launch per kernel, read by bench.py).
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")

METRICS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__inst_executed.sum",
    "smsp__thread_inst_executed.sum",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu.sum",
    "lts__t_sector_hit_rate.pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed_op_shared_atom.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum",
]

# bench.py's stage names for the kernels of the path
BENCH_NAME = [
    (r"d8_kernel", "d8_kernel"),
    (r"acc_fused", "acc_fused_kernel"),
    (r"acc_local_w", "acc_local_w_kernel"),
    (r"acc_final_w", "acc_final_w_kernel"),
    (r"coarse_resolve_kernel<.*i128", "coarse_resolve_kernel<i128>"),
    (r"weight_max", "weight_max_kernel"),
    (r"acc_local", "acc_local_cnt_kernel"),
    (r"super_tile_kernel<\(int\)1>|super_tile_kernel<1>", "super_tile_kernel<1>"),
    (r"super_tile_kernel<\(int\)3>|super_tile_kernel<3>", "super_tile_kernel<3>"),
    (r"coarse_resolve", "coarse_resolve_kernel"),
    (r"acc_final", "acc_final_cnt_kernel"),
    (r"terrain_kernel", "terrain_kernel"),
]


def short(name):
    m = re.search(r"(\w+)(<[^(]*>)?\(", name)
    base = re.sub(r"^void\s+", "", name)
    base = re.sub(r"<unnamed>::", "", base)
    return base.split("(")[0] if m else base


def to_bytes(v, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(v) * scale.get(unit, 1)


def to_ms(v, unit):
    scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3,
             "second": 1e3, "nsecond": 1e-6}
    return float(v) * scale.get(unit, 1)


def read_report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv", "--metrics",
                          ",".join(METRICS)], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(head)}
    res = []
    for r in rows[2:]:
        if len(r) < len(head):
            continue
        d = {"kernel": short(r[col["Kernel Name"]]), "grid": r[col["Grid Size"]],
             "block": r[col["Block Size"]]}
        for m in METRICS:
            if m not in col or r[col[m]] == "":
                continue
            v, u = r[col[m]].replace(",", ""), units[col[m]]
            if m.startswith("dram__bytes"):
                d[m] = to_bytes(v, u)
            elif m == "gpu__time_duration.sum":
                d[m] = to_ms(v, u)
            else:
                try:
                    d[m] = float(v)
                except ValueError:
                    d[m] = v
        res.append(d)
    return res


def summarise_launches(path, tag):
    """ncu --csv --log-file output of the gpu__time_duration pass."""
    text = open(path, errors="replace").read().splitlines()
    start = next(i for i, l in enumerate(text) if l.startswith('"ID"'))
    rows = list(csv.DictReader(io.StringIO("\n".join(text[start:]))))
    agg = collections.OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = short(r["Kernel Name"])
        ms = to_ms(r["Metric Value"].replace(",", ""), r["Metric Unit"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values()) or 1.0
    dst = os.path.join(PROF, f"{tag}_launches.csv")
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_ms", "mean_ms", "share_of_listed_time"])
        for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, n, f"{ms:.4f}", f"{ms / n:.4f}", f"{ms / tot:.4f}"])
    print("wrote", dst)


def main():
    args = sys.argv[1:]
    tag = args.pop(0)
    launches = None
    if "--launches" in args:
        i = args.index("--launches")
        launches = args[i + 1]
        del args[i:i + 2]
    os.makedirs(PROF, exist_ok=True)
    rows = []
    for rep in args:
        for d in read_report(rep):
            d["report"] = os.path.basename(rep)
            rows.append(d)
    if rows:
        dst = os.path.join(PROF, f"{tag}_kernels.csv")
        cols = ["report", "kernel", "grid", "block"] + METRICS
        with open(dst, "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(cols)
            for d in rows:
                w.writerow([d.get(c, "") for c in cols])
        print("wrote", dst)
        tpath = os.path.join(PROF, "traffic.json")
        try:
            traffic = json.load(open(tpath))
        except Exception:
            traffic = {}
        for d in rows:
            for pat, name in BENCH_NAME:
                if re.search(pat, d["kernel"]):
                    traffic[name] = d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
                    traffic[name + "@source"] = f"{d['report']} ({tag}), grid {d['grid']}"
                    break
        json.dump(traffic, open(tpath, "w"), indent=1, sort_keys=True)
        print("wrote", tpath)
        for d in rows:
            print(f"{d['kernel'][:60]:60s} {d.get('gpu__time_duration.sum', 0):8.3f} ms  "
                  f"dram {(d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)) / 1e6:9.1f} MB  "
                  f"dram% {d.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 0):5.1f}  "
                  f"sm% {d.get('sm__throughput.avg.pct_of_peak_sustained_elapsed', 0):5.1f}  "
                  f"winst {d.get('smsp__inst_executed.sum', 0) / 1e6:8.1f} M  "
                  f"regs {d.get('launch__registers_per_thread', 0):.0f}")
    if launches:
        summarise_launches(launches, tag)


if __name__ == "__main__":
    main()
