#!/bin/bash
# quick counter comparison of force-kernel variants at 128^3 (one launch each, step 30): usage r2_ncu_quick.sh name "opts" regex ...
cd "$(dirname "$0")/.."
M=gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_bank_conflicts_pipe_lsu.sum,l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,dram__bytes_read.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,l1tex__m_xbar2l1tex_read_sectors.sum
while [ $# -ge 3 ]; do
  name=$1; opts=$2; rx=$3; shift 3
  timeout 200 ncu --clock-control none -k regex:$rx -s 30 -c 1 --metrics $M --csv --log-file gpurun_out/q_$name.csv python profiles/profile_case.py --nx 128 --steps 40 $opts > gpurun_out/q_$name.log 2>&1
  python - "$name" <<'PY'
import csv,sys
rows=list(csv.reader(open('gpurun_out/q_%s.csv'%sys.argv[1])))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; mn=h.index('Metric Name'); mv=h.index('Metric Value'); mu=h.index('Metric Unit')
v={r[mn]:float(r[mv].replace(',','')) for r in rows[hi+1:]}
u={r[mn]:r[mu] for r in rows[hi+1:]}
t=v['gpu__time_duration.sum']; t=t/1e6 if u['gpu__time_duration.sum']=='ns' else (t/1e3 if u['gpu__time_duration.sum']=='us' else t)
pw=8388608/32*74.9
print("%-14s %.3f ms | per pair-warp: D %.2f (conflicts %.2f) T %.2f sectors %.1f req %.2f | L1 hit %.1f%% L2 hit %.1f%% L2->L1 sectors/pw %.2f | D %.1f%% fp64 %.1f%% issue %.1f%% warps %.1f%% lsb %.1f | inst %.0fM dram rd %.2f GB" % (
 sys.argv[1], t, v['l1tex__data_pipe_lsu_wavefronts.sum']/pw, v['l1tex__data_bank_conflicts_pipe_lsu.sum']/pw, v['l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum']/pw,
 v['l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum']/pw, v['l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum']/pw, v['l1tex__t_sector_hit_rate.pct'], v['lts__t_sector_hit_rate.pct'], v['l1tex__m_xbar2l1tex_read_sectors.sum']/pw,
 v['l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'], v['sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'], v['smsp__issue_active.avg.pct_of_peak_sustained_active'],
 v['sm__warps_active.avg.pct_of_peak_sustained_active'], v['smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio'], v['smsp__inst_executed.sum']/1e6, v['dram__bytes_read.sum']/(1e9 if u['dram__bytes_read.sum']=='byte' else 1)))
PY
done
