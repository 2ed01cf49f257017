#!/bin/bash
# Round 2, 8-GPU call: C2 at 8 GPUs, C5 (n=5000, p=500k, q=20k) at 8 GPUs [and at 4 with a third argument], NCCL parity (2 ranks).
T=${1:-r2n8}
mkdir -p gpurun_out
{ nproc; free -g | head -2; nvidia-smi -L | wc -l; } > gpurun_out/${T}_box.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/${T}_bench_c2_g8.json 2> gpurun_out/${T}_bench_c2_g8.err; echo "c2 g8 rc=$?"
tail -2 gpurun_out/${T}_bench_c2_g8.err; cut -c1-330 gpurun_out/${T}_bench_c2_g8.json
timeout 900 $TR --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 --config C5 --steps 2 --warmup 1 > gpurun_out/${T}_bench_c5_g8.json 2> gpurun_out/${T}_bench_c5_g8.err; echo "c5 g8 rc=$?"
tail -2 gpurun_out/${T}_bench_c5_g8.err; cut -c1-330 gpurun_out/${T}_bench_c5_g8.json
timeout 300 python -m pytest tests/test_gpu_multi_slab.py -m gpu -q -k nccl > gpurun_out/${T}_nccl_test.log 2>&1; echo "nccl test rc=$?"; tail -3 gpurun_out/${T}_nccl_test.log
if [ -n "$2" ]; then
timeout 1500 $TR --nproc-per-node 4 --master-port 29514 bench.py --gpus 4 --config C5 --steps 2 --warmup 1 > gpurun_out/${T}_bench_c5_g4.json 2> gpurun_out/${T}_bench_c5_g4.err; echo "c5 g4 rc=$?"
tail -2 gpurun_out/${T}_bench_c5_g4.err; cut -c1-330 gpurun_out/${T}_bench_c5_g4.json
fi
