#!/bin/bash
# A/B on one box: each variant measured twice, interleaved
L=$PWD/atlasqtl_b200
for rep in 1 2; do
for v in libatlasqtl_b200 libab_cfwd; do
  echo "== $v rep $rep"
  AQ_LIB=$L/$v.so timeout 300 python tools/prof_sweep.py 5000 8000 2500 3 | tail -1
  AQ_LIB=$L/$v.so timeout 300 python tools/prof_sweep.py 3000 8000 1500 3 | tail -1
  AQ_LIB=$L/$v.so timeout 300 python tools/prof_sweep.py 1500 8000 2400 3 | tail -1
done; done
AQ_LIB=$L/libab_cfwd.so timeout 600 python -m pytest tests/test_gpu_edge.py tests/test_gpu_missing.py -m gpu -q -k "ragged or missing" 2>&1 | tail -3
