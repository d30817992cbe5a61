#!/bin/bash
# 8-GPU evidence run (one box): distributed parity test, bench line at N=8 (with the concatenated-batch parity check),
# BASELINE configs[1] end to end (scripts/config2_step.py).  Outputs under gpurun_out/.
mkdir -p gpurun_out
N=${N:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q -s > gpurun_out/dist_n$N.log 2>&1; echo "dist rc=$?"; tail -2 gpurun_out/dist_n$N.log | cut -c1-200
timeout 400 $TR --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_n$N.log | cut -c1-300
timeout 500 $TR --master-port 29541 scripts/config2_step.py --steps 10 --warmup 3 > gpurun_out/config2_n$N.log 2>&1; echo "config2 rc=$?"; tail -1 gpurun_out/config2_n$N.log | cut -c1-600
