#!/bin/bash
# Round-end evidence run on one B200 (gpurun): GPU tests, bench (own arm + reference arm), ncu captures of the dominant
# kernels and the launch lists.  Every command runs plain first; a number printed under ncu is never a bench value.
O=gpurun_out
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -45 > $O/r02_pytest_gpu_final.log
timeout 400 python bench.py > $O/r02_bench_final.json 2> $O/bench_final.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_bench_reference_arm.json 2> $O/bench_ref.err
timeout 100 python scripts/gpu_berk_once.py > /dev/null 2>&1 && \
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"igemm_tf32_gdn|nhwc_split" --launch-skip 4 -c 4 \
    -o $O/r02_berk_final2 -f python scripts/gpu_berk_once.py > $O/ncu_berk2.log 2>&1
timeout 100 python scripts/gpu_lift_tc_once.py > /dev/null 2>&1 && \
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:lift_step_tc --launch-skip 2 -c 1 \
    -o $O/r02_lift_tc_final2 -f python scripts/gpu_lift_tc_once.py > $O/ncu_lift2.log 2>&1
timeout 100 python scripts/gpu_cond2zt_once.py > /dev/null 2>&1 && \
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_cond2zt_launches_final2.csv \
    python scripts/gpu_cond2zt_once.py > $O/ncu_c2b.log 2>&1
timeout 200 python bench.py --steps 1 --warmup 1 --blocks none > $O/bench_for_list2.json 2> $O/bench_for_list2.err && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1800 --csv --log-file $O/r02_launches_headline_final2.csv \
    python bench.py --steps 1 --warmup 1 --blocks none > $O/ncu_list2.log 2>&1
tail -3 $O/r02_pytest_gpu_final.log
cut -c1-300 $O/r02_bench_final.json
cut -c1-300 $O/r02_bench_reference_arm.json
