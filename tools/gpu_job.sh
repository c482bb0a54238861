#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python bench.py 2> gpurun_out/bench_full5.err | tail -n 1 > gpurun_out/bench_full5.json
cut -c1-200 gpurun_out/bench_full5.json; tail -n 3 gpurun_out/bench_full5.err
(timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2h_launches_cfg2_bench.csv python bench.py --configs none --no-e2e --no-cpu --steps 2 --warmup 1 > gpurun_out/ncu_a.log 2>&1)
awk -F'","' 'NR>2{print $5, $NF}' gpurun_out/r2h_launches_cfg2_bench.csv | sed -n 3,14p
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -n 1 | cut -c1-400
