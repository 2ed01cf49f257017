#!/bin/bash
# Measurement on one B200 (run under gpurun): plain bench, then the ncu launch list of the same command, then one
# `--set full` capture of a sweep launch.  Numbers printed under ncu are never bench values.
# usage: tools/profile_c2.sh <tag>      (afterwards, here: python tools/ncu_summary.py <tag> to refresh profiles/)
set -u
T=${1:-r2}
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2_$T.json 2> gpurun_out/bench_c2_$T.err || exit 1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_c2_$T.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2_$T.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l_$T.log 2>&1
# sweep_kernel launches: 0 = set_state, then one per iteration (segmented sweep: no tail launch): skip 2 = iteration 2
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 2 -c 1 -f \
    -o gpurun_out/prof_c2_sweep_$T python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f_$T.log 2>&1
# the table pass (second kernel of a step by time): one capture of an annealed-iteration launch
ncu --set full --clock-control none --import-source on -k regex:tables_kernel -s 2 -c 1 -f \
    -o gpurun_out/prof_c2_tables_$T python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_t_$T.log 2>&1
ls -la gpurun_out/*_$T.*
cat gpurun_out/bench_c2_$T.json | cut -c1-600
