#!/bin/bash
# One bench line per BASELINE.json config (N=1), tracked under profiles/ (see DESIGN.md section 5).
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python bench.py --steps 10 --warmup 3 "$@" > gpurun_out/$name.log 2>&1; echo "== $name rc=$?"; tail -1 gpurun_out/$name.log | cut -c1-${CUT:-260}; }
run bench_voc
run bench_voc_b3 --batch 3
run bench_ade --workload ade
run bench_city --workload city
for px in 8192 32768 131072 262144; do run bench_sweep_$px --workload sweep:$px --no-cpu-baseline; done
