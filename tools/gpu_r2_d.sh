#!/bin/bash
# Round 2, GPU call D (1 GPU): A/B of the pipelined S-phase finish (three builds), cluster sizes for n = 3000 / 5000.
mkdir -p gpurun_out
L=atlasqtl_b200
for v in "" _sfm1 _sf2; do
  echo "== lib$v: 1000 x 50000 x 2500 (8-GPU slab of C2)"; AQ_LIB=$PWD/$L/libatlasqtl_b200$v.so timeout 300 python tools/prof_sweep.py 1000 50000 2500 4 | tail -2
done
timeout 300 python -m pytest tests/test_gpu_sweep.py tests/test_gpu_segments.py -m gpu -q 2>&1 | tail -3
for n in 3000 5000; do for c in 4 5 6 7 8; do
  q=1500; [ $n = 5000 ] && q=2500
  echo "== n=$n cluster=$c"; AQ_FORCE_CLUSTER=$c timeout 300 python tools/prof_sweep.py $n 8000 $q 2 2>&1 | tail -1
done; done
