#!/bin/bash
mkdir -p gpurun_out
(timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_cfg5_512.csv python bench.py --config cfg5 --configs none --no-e2e --no-cpu --steps 1 --warmup 1 > gpurun_out/ncu_l5.log 2>&1)
(timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_ph_sync -c 2 -o gpurun_out/r2c_cfg5_sync -f python bench.py --config cfg5 --batch 64 --configs none --no-e2e --no-cpu --steps 1 --warmup 0 > gpurun_out/ncu_r2c.log 2>&1)
tail -n 2 gpurun_out/ncu_l5.log
