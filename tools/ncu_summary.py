"""Summaries of ncu output for profiles/.

    python tools/ncu_summary.py launches gpurun_out/TAG_launches.csv profiles/TAG_launches_summary.csv
        per-kernel launch count, total time and share of the GPU time of an `ncu --metrics gpu__time_duration.sum --csv` launch list
    python tools/ncu_summary.py full gpurun_out/TAG_full.ncu-rep profiles/TAG_ncu_full
        writes _raw.csv (ncu --page raw --csv), _summary.json (the metrics DESIGN.md / bench.py quote, per launch) and
        traffic per kernel name (dram bytes read + written per launch)
"""
import csv
import io
import json
import re
import subprocess
import sys
from collections import OrderedDict

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "sm__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__cycles_active.avg", "smsp__cycles_active.min", "smsp__cycles_active.max", "sm__cycles_elapsed.max", "sm__cycles_elapsed.avg",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.min", "smsp__inst_executed.max", "smsp__inst_executed.avg",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum",
]


def short(name):
    m = re.match(r"(?:void )?([A-Za-z0-9_]+)", name)
    return m.group(1) if m else name


def launches(src, dst):
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    hdr = rows[0]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows[1:]:
        k = short(r[kn])
        n, t = agg.get(k, (0, 0.0))
        agg[k] = (n + 1, t + float(r[mv].replace(",", "")))
    tot = sum(t for _, t in agg.values())
    with open(dst, "w") as f:
        f.write("kernel,launches,total_ns,share_of_gpu_time\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k},{n},{int(t)},{t / tot:.4f}\n")
    print(open(dst).read())


def full(rep, prefix):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    open(prefix + "_raw.csv", "w").write(raw)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    out, traffic = [], {}
    for r in rows[2:]:
        d = {"kernel": short(r[kn]), "grid": r[hdr.index("Grid Size")], "block": r[hdr.index("Block Size")]}
        for i, h in enumerate(hdr):
            if h in KEEP:
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                u = units[i]
                if h.startswith("dram__bytes") or h.endswith("t_bytes.sum"):
                    v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                    u = "byte"
                if h == "gpu__time_duration.sum":
                    v *= {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1)
                    u = "ns"
                d[h] = v
                d.setdefault("_units", {})[h] = u
        d["dram_bytes"] = d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
        out.append(d)
    json.dump(out, open(prefix + "_summary.json", "w"), indent=1)
    for d in out:
        print(f"{d['kernel']:16s} grid {d['grid']:>14s} {d.get('gpu__time_duration.sum', 0) / 1e6:8.4f} ms  dram {d['dram_bytes'] / 1e6:8.1f} MB  "
              f"inst {d.get('smsp__inst_executed.sum', 0) / 1e6:8.1f} M  issue(act) {d.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0):5.1f} %  "
              f"smsp active min/avg/max {d.get('smsp__cycles_active.min', 0) / 1e6:.3f}/{d.get('smsp__cycles_active.avg', 0) / 1e6:.3f}/{d.get('smsp__cycles_active.max', 0) / 1e6:.3f} M  "
              f"elapsed {d.get('sm__cycles_elapsed.max', 0) / 1e6:.3f} M  smem wavefronts {d.get('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 0) / 1e6:.1f} M")


def traffic(summary_json, dst):
    """profiles/rNN_traffic.json for bench.py: DRAM bytes (read + written) per launch of the workload's main launches, from a _summary.json"""
    rows = json.load(open(summary_json))
    big = {}
    for d in rows:                                   # the longest launch of every kernel = the one on the whole recording / the MSC set
        k = d["kernel"]
        if k not in big or d.get("gpu__time_duration.sum", 0) > big[k].get("gpu__time_duration.sum", 0):
            big[k] = dict(d)
    for k, b in big.items():                         # a launch split in two over the same grid (k_vit_trace: DAB / data and DAB+ halves): their sum
        same = [d for d in rows if d["kernel"] == k and d["grid"] == b["grid"]]
        if k == "k_vit_trace" and len(same) == 2:
            b["dram_bytes"] = sum(d["dram_bytes"] for d in same)
            b["gpu__time_duration.sum"] = sum(d.get("gpu__time_duration.sum", 0) for d in same)
    name = {"k_viterbi": "k_viterbi(msc)", "k_gather": "k_gather(msc)", "k_vit_trace": "k_vit_trace(msc)"}
    out = {"source": f"ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch ({summary_json}); k_ofdm_demod: the launch on frames 16.. (the first 16 frames are a second, small launch)",
           "bytes_per_launch": {name.get(k, k): int(d["dram_bytes"]) for k, d in big.items()},
           "ms_under_ncu": {name.get(k, k): round(d.get("gpu__time_duration.sum", 0) / 1e6, 4) for k, d in big.items()}}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])
