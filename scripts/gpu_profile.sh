#!/bin/bash
# ncu evidence: launch list of one bench command + full-set capture of the hot kernels (1 GPU, short run)
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/launches_run.log 2>&1
echo "launches rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"con_sweep|upsample|unce_|unkd_|prep_pack" -s ${SKIP:-11} -c ${COUNT:-11} -f -o gpurun_out/prof $B > gpurun_out/prof_run.log 2>&1
echo "prof rc=$?"; ls -la gpurun_out | tail -8
