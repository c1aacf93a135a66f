"""Helpers to summarise ncu outputs (launch list CSV and .ncu-rep raw pages) into the markdown kept under profiles/."""
import collections
import csv
import re
import subprocess
import sys


def launch_table(path, min_launches=0):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        k = re.sub(r"\(.*", "", row["Kernel Name"])[:90]
        agg[k][0] += 1
        agg[k][1] += v
    return {k: v for k, v in agg.items() if v[0] >= min_launches}


def raw_metrics(rep, names):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    kcol = hdr.index("Kernel Name")
    res = []
    for r in rows[2:]:
        d = {"kernel": re.sub(r"\(.*", "", r[kcol])[:60]}
        for n in names:
            if n in hdr:
                d[n] = r[hdr.index(n)] + " " + rows[1][hdr.index(n)]
        res.append(d)
    return res


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        t = launch_table(sys.argv[2])
        for k, v in sorted(t.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 20]:
            print(f"{v[1]:10.1f} us  n={v[0]:4d}  avg={v[1]/v[0]:8.2f}  {k}")
    else:
        names = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                 "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                 "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
                 "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
                 "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "launch__grid_size",
                 "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                 "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
                 "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio" ]
        for d in raw_metrics(sys.argv[2], names):
            print("----", d.pop("kernel"))
            for k, v in d.items():
                print(f"   {k:75s} {v}")
