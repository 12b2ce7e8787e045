#!/bin/bash
# Round-2 session-3 closing measurements on one B200 (run under gpurun): full GPU test suite, smoke(), default bench line, ncu launch
# list of the bench command, ncu --set full of the dominant kernels at 128^3 (reports come back in gpurun_out/; traffic.json is
# refreshed from them on the build box with profiles/update_traffic.py <report> <key> <summary> <commit>).
cd "$(dirname "$0")/.."
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 400 python -m pytest tests -q -m gpu > $O/r2s3_final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2s3_final_pytest.log
python __graft_entry__.py --smoke > $O/r2s3_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r2s3_final_smoke.log
timeout 900 python bench.py > $O/r2s3_final_bench_n1.json 2> $O/r2s3_final_bench_n1.err; echo "bench rc=$?"; tail -c 400 $O/r2s3_final_bench_n1.json; tail -2 $O/r2s3_final_bench_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_r2s3_bench128.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-secondary --no-parity > $O/r2s3_final_ncu_l.log 2>&1
python profiles/summarize.py launches $O/launches_r2s3_bench128.csv | head -12
timeout 150 $NCU -k regex:k_force_lj_full_fi -s 30 -c 1 -o $O/prof_r2s3_vlforce128 python profiles/profile_case.py --nx 128 --steps 45 > $O/r2s3_final_ncu_vl.log 2>&1
timeout 150 $NCU -k regex:k_cp_force_lj_sp_duo -s 30 -c 1 -o $O/prof_r2s3_cpforce128 python profiles/cp_case.py --nx 128 --steps 45 --timing 0 > $O/r2s3_final_ncu_cp.log 2>&1
timeout 150 $NCU -k regex:k_build_neighbor_v6 -s 1 -c 1 -o $O/prof_r2s3_neigh128 python profiles/profile_case.py --nx 128 --steps 25 > $O/r2s3_final_ncu_nb.log 2>&1
timeout 150 $NCU -k regex:k_cp_build_neighbor -s 1 -c 1 -o $O/prof_r2s3_cpneigh128 python profiles/cp_case.py --nx 128 --steps 25 --timing 0 > $O/r2s3_final_ncu_cpnb.log 2>&1
ls -la $O/*r2s3*.ncu-rep
