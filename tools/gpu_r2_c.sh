#!/bin/bash
# Round 2, GPU call C (1 GPU): the whole GPU suite, a one-GPU rehearsal of the C5 path, then the C2 profile set.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r2c_tests.log
AQ_BENCH_PACKED_ABOVE_GB=0.1 timeout 600 python bench.py --config C5slab --steps 3 --warmup 1 > gpurun_out/r2c_bench_c5slab.json 2> gpurun_out/r2c_bench_c5slab.err; echo "c5slab rc=$?"
tail -5 gpurun_out/r2c_bench_c5slab.err; cut -c1-1500 gpurun_out/r2c_bench_c5slab.json
bash tools/profile_c2.sh r2a
