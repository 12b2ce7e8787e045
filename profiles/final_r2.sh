#!/bin/bash
# Round-2 closing measurements on one B200 (run under gpurun): full GPU test suite, default bench line, ncu launch list of the
# bench command, ncu --set full of the dominant kernels at 128^3 (reports come back in gpurun_out/; traffic.json is refreshed
# from them on the build box with profiles/update_traffic.py <report> <key> <summary> <commit>).
cd "$(dirname "$0")/.."
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 python -m pytest tests -x -q -m gpu > $O/r2_final_pytest.log 2>&1; tail -3 $O/r2_final_pytest.log
timeout 600 python bench.py > $O/r2_final_bench_n1.json 2> $O/r2_final_bench_n1.err; echo "bench rc=$?"; tail -c 400 $O/r2_final_bench_n1.json; tail -2 $O/r2_final_bench_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_r2_bench128.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-secondary --no-parity > $O/r2_final_ncu_l.log 2>&1
python profiles/summarize.py launches $O/launches_r2_bench128.csv | head -12
timeout 150 $NCU -k regex:k_force_lj_full_fi -s 30 -c 1 -o $O/prof_r2_vlforce128 python profiles/profile_case.py --nx 128 --steps 45 > $O/r2_final_ncu_vl.log 2>&1
timeout 150 $NCU -k regex:k_cp_force_lj_sp_duo -s 30 -c 1 -o $O/prof_r2_cpforce128 python profiles/cp_case.py --nx 128 --steps 45 --timing 0 > $O/r2_final_ncu_cp.log 2>&1
timeout 150 $NCU -k regex:k_build_neighbor_v6 -s 1 -c 1 -o $O/prof_r2_neigh128 python profiles/profile_case.py --nx 128 --steps 25 > $O/r2_final_ncu_nb.log 2>&1
ls -la $O/*.ncu-rep
