#!/bin/bash
# GPU-box job of the moment (experiment helper, not product).  Everything it writes goes to gpurun_out/.
mkdir -p gpurun_out
(timeout 500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu4.txt 2>&1; echo rc=$? >> gpurun_out/pytest_gpu4.txt)
(timeout 300 python bench.py --config cfg5 --configs none --no-e2e --steps 5 --warmup 2 > gpurun_out/bench_cfg5_b.txt 2> gpurun_out/bench_cfg5_b.err)
(timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_ph_sync -c 2 -o gpurun_out/r2b_cfg5_sync -f python bench.py --config cfg5 --batch 64 --configs none --no-e2e --no-cpu --steps 1 --warmup 0 > gpurun_out/ncu_r2b.log 2>&1)
(timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_huff_lane -c 1 -o gpurun_out/r2b_cfg5_vseg -f python bench.py --config cfg5 --batch 64 --configs none --no-e2e --no-cpu --steps 1 --warmup 0 > gpurun_out/ncu_r2b2.log 2>&1)
(timeout 300 python bench.py --config cfg2 --configs none --no-e2e --no-cpu --steps 10 --warmup 3 --idct-kernel 0 > gpurun_out/bench_cfg2_b.txt 2>&1)
grep -E "passed|failed" gpurun_out/pytest_gpu4.txt | tail -n 3
