#!/bin/bash
# Round 2, GPU call A (1 GPU): box facts, the whole GPU suite, C2 bench with / without the segmented sweep, and the
# 8-GPU slab shape (q_local = 2500) on one GPU with / without segments.
mkdir -p gpurun_out
{ nproc; free -g | head -2; nvidia-smi -L; } > gpurun_out/r2a_box.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2a_tests.log
tail -15 gpurun_out/r2a_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench_c2_seg.json 2> gpurun_out/r2a_bench_c2_seg.err; echo "bench seg rc=$?"
AQ_NO_SEG=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_c2_noseg.json 2> gpurun_out/r2a_bench_c2_noseg.err; echo "bench noseg rc=$?"
timeout 300 python tools/prof_sweep.py 1000 50000 2500 4 > gpurun_out/r2a_slab2500_seg.log 2>&1; echo "slab seg rc=$?"
AQ_NO_SEG=1 timeout 300 python tools/prof_sweep.py 1000 50000 2500 4 > gpurun_out/r2a_slab2500_noseg.log 2>&1; echo "slab noseg rc=$?"
timeout 300 python tools/prof_sweep.py 1000 50000 5000 3 > gpurun_out/r2a_slab5000_seg.log 2>&1
AQ_NO_SEG=1 timeout 300 python tools/prof_sweep.py 1000 50000 5000 3 > gpurun_out/r2a_slab5000_noseg.log 2>&1
cat gpurun_out/r2a_box.txt gpurun_out/r2a_slab*.log
cat gpurun_out/r2a_bench_c2_seg.json gpurun_out/r2a_bench_c2_noseg.json | cut -c1-1800
tail -3 gpurun_out/r2a_bench_c2_seg.err
