#!/bin/bash
# Round 2 measurement pass on ONE B200 (under gpurun), in three parts (a gpurun call returns at most 64 MiB and an
# `ncu --set full` report of this kernel is ~24 MB):
#   lines    fp64 peaks with a clock record; one bench line per config that fits one GPU (C1, C3, C4, the one-GPU rehearsal
#            of C5, C4 with 5 % missing responses)
#   c2       the C2 bench line + profile set (launch list, `--set full` of the sweep and of the table pass)
#   others   one `--set full` capture each of the clustered sweep (C5 shape) and of the missing-response sweep (C4 shape)
# usage: tools/gpu_r2_measure.sh <part> <tag>
PART=${1:-lines}
T=${2:-r2}
mkdir -p gpurun_out
B="python bench.py"
if [ $PART = lines ]; then
  bash tools/microbench/run_fp64_peaks.sh $T > /dev/null 2>&1
  timeout 900 $B --config C1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C1_$T.json 2> gpurun_out/bench_C1_$T.err
  timeout 900 $B --config C4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C4_$T.json 2> gpurun_out/bench_C4_$T.err
  timeout 900 $B --config C4 --steps 10 --warmup 3 --na-frac 0.05 > gpurun_out/bench_C4na_$T.json 2> gpurun_out/bench_C4na_$T.err
  timeout 1500 $B --config C3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C3_$T.json 2> gpurun_out/bench_C3_$T.err
  AQ_BENCH_PACKED_ABOVE_GB=0.1 timeout 900 $B --config C5slab --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C5slab_$T.json 2> gpurun_out/bench_C5slab_$T.err
  for f in C1 C4 C4na C3 C5slab; do echo "== $f"; cut -c1-260 gpurun_out/bench_${f}_$T.json; tail -1 gpurun_out/bench_${f}_$T.err; done
  tail -12 gpurun_out/fp64_peaks_$T.txt
elif [ $PART = c2 ]; then
  bash tools/profile_c2.sh $T > gpurun_out/profile_c2_$T.log 2>&1
  cut -c1-300 gpurun_out/bench_c2_$T.json
else
  # launch 0 = set_state, 2 = second sweep
  AQ_BENCH_PACKED_ABOVE_GB=0.1 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 2 -c 1 -f \
      -o gpurun_out/prof_c5slab_sweep_$T python bench.py --config C5slab --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c5_$T.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 2 -c 1 -f \
      -o gpurun_out/prof_c4na_sweep_$T python bench.py --config C4 --na-frac 0.05 --steps 1 --warmup 3 > gpurun_out/ncu_c4na_$T.log 2>&1
fi
ls -la gpurun_out/*$T*
