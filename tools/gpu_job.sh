#!/bin/bash
# GPU-box job of the moment (experiment helper, not product).  Everything it writes goes to gpurun_out/.
mkdir -p gpurun_out
(timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu2.txt 2>&1; echo rc=$? >> gpurun_out/pytest_gpu2.txt)
(timeout 700 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full2.txt 2> gpurun_out/bench_full2.err)
(timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_cfg5.csv python bench.py --config cfg5 --batch 64 --configs none --no-e2e --no-cpu --steps 1 --warmup 1 > gpurun_out/ncu_cfg5.log 2>&1)
(timeout 120 tools/cu/viaddmin_test > gpurun_out/viaddmin.txt 2>&1)
for v in v0 v1 v2 v3 v4 v5 v6 v7; do
  (JSGPU_LIB=$PWD/jpegsnoop_b200/variants/libjsgpu_$v.so timeout 200 python bench.py --configs none --no-e2e --no-cpu --steps 10 --warmup 3 > gpurun_out/var_$v.txt 2> gpurun_out/var_$v.err)
done
(JSGPU_LIB=$PWD/jpegsnoop_b200/variants/libjsgpu_v1.so timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batch_matches or random_corpus" > gpurun_out/pytest_v1.txt 2>&1)
(JSGPU_LIB=$PWD/jpegsnoop_b200/variants/libjsgpu_v2.so timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batch_matches or random_corpus" > gpurun_out/pytest_v2.txt 2>&1)
tail -n 2 gpurun_out/pytest_gpu2.txt gpurun_out/viaddmin.txt gpurun_out/pytest_v1.txt
