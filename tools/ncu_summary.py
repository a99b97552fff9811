"""Extract the judged metrics from .ncu-rep files (`ncu --set full` captures) into one JSON list:
    python tools/ncu_summary.py gpurun_out/r02_*.ncu-rep > profiles/r02_ncu_full_summary.json
One entry per profiled launch: kernel, duration, DRAM bytes read / written, DRAM %, tensor-pipe %, L2 hit rate,
achieved occupancy, registers, grid."""
import csv
import io
import json
import os
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
           "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
           "launch__cluster_size", "lts__t_bytes.sum"]


def summarize(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(io.StringIO(out.stdout)))
    hdr = next((i for i, r in enumerate(rows) if "Kernel Name" in r), None)
    if hdr is None:
        return []
    names, units = rows[hdr], rows[hdr + 1]
    res = []
    for r in rows[hdr + 2:]:
        if len(r) != len(names):
            continue
        d = dict(zip(names, r))
        e = {"report": os.path.basename(path).replace(".ncu-rep", ""), "kernel": d["Kernel Name"][:90]}
        for m in METRICS:
            if m in d and d[m] != "":
                u = units[names.index(m)]
                e[m] = f"{d[m]} {u}".strip()
        res.append(e)
    return res


if __name__ == "__main__":
    allr = []
    for p in sys.argv[1:]:
        allr += summarize(p)
    json.dump(allr, sys.stdout, indent=1)
