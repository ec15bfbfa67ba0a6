#!/bin/bash
# Second half of the round-end evidence run: GPU tests of the committed state, the step's kernel shares without ncu
# (torch.profiler) and the ncu launch list long enough to hold one complete timed step.
O=gpurun_out
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -45 > $O/r02_pytest_gpu_final.log
timeout 200 python scripts/gpu_profile_step.py > $O/r02_step_profile.txt 2>&1
timeout 200 python bench.py --steps 1 --warmup 1 --blocks none > $O/bench_for_list2.json 2> $O/bench_for_list2.err && \
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file $O/r02_launches_headline_final2.csv \
    python bench.py --steps 1 --warmup 1 --blocks none > $O/ncu_list2.log 2>&1
tail -3 $O/r02_pytest_gpu_final.log
head -12 $O/r02_step_profile.txt
