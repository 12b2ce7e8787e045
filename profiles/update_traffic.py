"""Refresh one entry of profiles/traffic.json from an `ncu --set full` capture of the dominant kernel.
usage: python profiles/update_traffic.py <report.ncu-rep> <key, e.g. verletlist/dp/128> <summary file the numbers are kept in>
The summary (profiles/summarize.py raw) is written to <summary file>; the entry gets dram__bytes_read.sum + dram__bytes_write.sum
of the FIRST kernel in the report."""
import csv
import json
import os
import subprocess
import sys

rep, key, summ = sys.argv[1:4]
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
tj[key] = {"bytes": int(rd + wr), "kernel": kernel,
           "source": "%s (%.3f GB read + %.2f MB write)" % (os.path.relpath(summ, ROOT), rd / 1e9, wr / 1e6)}
json.dump(tj, open(path, "w"), indent=1)
print(key, tj[key])
