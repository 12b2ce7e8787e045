"""Refresh one entry of profiles/traffic.json from an `ncu --set full` capture of the dominant kernel.
usage: python profiles/update_traffic.py <report.ncu-rep> <key, e.g. verletlist/dp/128> <summary file the numbers are kept in> [git commit the capture was taken at]
The summary (profiles/summarize.py raw) is written to <summary file>; the entry gets dram__bytes_read.sum + dram__bytes_write.sum
of the FIRST kernel in the report."""
import csv
import json
import os
import subprocess
import sys

rep, key, summ = sys.argv[1:4]
commit = sys.argv[4] if len(sys.argv) > 4 else None
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, r = rows[0], rows[1], rows[2]


def val(name):
    i = hdr.index(name)
    v, u = float(r[i].replace(",", "")), units[i].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]


rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
kernel = r[hdr.index("Kernel Name")].split("(")[0]
with open(summ, "w") as f:
    f.write(subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "summarize.py"), "raw", rep], capture_output=True, text=True).stdout)
path = os.path.join(ROOT, "profiles", "traffic.json")
tj = json.load(open(path))
def num(name):
    return float(r[hdr.index(name)].replace(",", "")) if name in hdr else None


dur_i = hdr.index("gpu__time_duration.sum")
dur = float(r[dur_i].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[dur_i].lower(), 1.0)
tj[key] = {"bytes": int(rd + wr), "kernel": kernel,
           "source": "%s (%.3f GB read + %.2f MB write)" % (os.path.relpath(summ, ROOT), rd / 1e9, wr / 1e6),
           "ncu": {"duration_ms": dur,
                   "fp64_pipe_pct": num("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
                   "fma_pipe_pct": num("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                   "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                   "l1tex_throughput_pct": num("l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
                   "dram_throughput_pct": num("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                   "fma_pipe_cycles_active_pct": num("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                   "l1_hit_pct": num("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": num("lts__t_sector_hit_rate.pct"),
                   "registers_per_thread": num("launch__registers_per_thread")},
           "commit": commit}
json.dump(tj, open(path, "w"), indent=1)
print(key, tj[key])
