"""l1tex sub-unit utilisation of one ncu --set full report: python profiles/ncu_l1.py <report.ncu-rep> [substr ...]"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
pats = sys.argv[2:] or ["l1tex__data_pipe", "l1tex__lsu_writeback", "l1tex__lsuin", "l1tex__t_output", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests",
                        "l1tex__data_bank", "l1tex__m_xbar2l1tex_read_sectors.sum", "l1tex__f_", "l1tex__t_set", "l1tex__throughput", "breakdown"]
for r in rows[2:]:
    print("----", r[hdr.index("Kernel Name")].split("(")[0])
    for h, u, v in zip(hdr, units, r):
        if any(p in h for p in pats) and v not in ("", "n/a") and not any(x in h for x in (".min", ".max", "per_second", "peak_sustained_active")):
            print("  %-90s %s %s" % (h, v, u))
