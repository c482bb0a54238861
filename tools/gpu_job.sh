#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -n 8
(timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2e_launches_cfg2_bench.csv python bench.py --configs none --no-e2e --no-cpu --steps 2 --warmup 1 > gpurun_out/ncu_a.log 2>&1)
(timeout 600 ncu --set full --clock-control none -c 12 -o gpurun_out/r2e_cfg2_full -f python bench.py --configs none --no-e2e --no-cpu --steps 1 --warmup 0 > gpurun_out/ncu_b.log 2>&1)
(timeout 400 ncu --set full --clock-control none -k regex:k_preview -c 4 -o gpurun_out/r2e_preview -f python bench.py --configs none --no-e2e --no-cpu --steps 1 --warmup 0 > gpurun_out/ncu_c.log 2>&1)
(timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2e_launches_cfg5.csv python bench.py --config cfg5 --configs none --no-e2e --no-cpu --steps 1 --warmup 1 > gpurun_out/ncu_d.log 2>&1)
ls -la gpurun_out/*.ncu-rep gpurun_out/r2e_*.csv
tail -n 2 gpurun_out/ncu_b.log | cut -c1-200
