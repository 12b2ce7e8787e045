#!/bin/bash
# Round-1 session-4 closing measurements on one B200 (run under gpurun): ncu --set full of the two dominant kernels at
# 128^3 (-> profiles/traffic.json on the box, copies in gpurun_out/), full GPU test suite, default bench line, ncu launch
# list of the bench command.
cd "$(dirname "$0")/.."
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 200 python -m pytest tests -x -q -m gpu > $O/pytest_s4_final.log 2>&1; tail -3 $O/pytest_s4_final.log
timeout 100 $NCU -k regex:k_force_lj_full_fi -s 30 -c 1 -o $O/prof_r1_s4_vlforce128 python profiles/profile_case.py --nx 128 --steps 45 > $O/ncu_s4_vl.log 2>&1
python profiles/update_traffic.py $O/prof_r1_s4_vlforce128.ncu-rep verletlist/dp/128 profiles/r1_s4_vlforce128_raw.txt
cp profiles/traffic.json profiles/r1_s4_vlforce128_raw.txt $O/
timeout 240 python bench.py > $O/bench_s4_n1.json 2> $O/bench_s4_n1.err; tail -c 600 $O/bench_s4_n1.json; tail -2 $O/bench_s4_n1.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_r1_s4_bench128.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-secondary > $O/ncu_s4_l.log 2>&1
python profiles/summarize.py launches $O/launches_r1_s4_bench128.csv | head -12
timeout 100 $NCU -k regex:k_cp_force_lj_sp_packed -s 30 -c 1 -o $O/prof_r1_s4_cpforce128 python profiles/cp_case.py --nx 128 --steps 45 --timing 0 > $O/ncu_s4_cp.log 2>&1
python profiles/update_traffic.py $O/prof_r1_s4_cpforce128.ncu-rep clusterpair/sp/128 profiles/r1_s4_cpforce128_raw.txt
cp profiles/traffic.json profiles/r1_s4_cpforce128_raw.txt $O/
