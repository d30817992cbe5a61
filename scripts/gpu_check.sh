#!/bin/bash
# First-contact GPU run: each group in its own process so that one trapped kernel does not hide the rest.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "== $name rc=$?"; tail -n ${TAILN:-12} gpurun_out/$name.log; }
run t_umma 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "umma"
run t_stream 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "upsample or unce or unkd"
run t_prep 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "label_downsample or prep_integer"
run t_con 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "contrastive_loss or compat or whole_hot or full_size"
run smoke 300 python __graft_entry__.py --smoke
TAILN=3 run bench 900 python bench.py --steps 5 --warmup 3
