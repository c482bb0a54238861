#!/bin/bash
# GPU-box job of the moment (experiment helper, not product).  Everything it writes goes to gpurun_out/.
mkdir -p gpurun_out
(timeout 500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu5.txt 2>&1; echo rc=$? >> gpurun_out/pytest_gpu5.txt)
(timeout 300 python bench.py --config cfg5 --configs none --no-e2e --no-cpu --steps 5 --warmup 2 > gpurun_out/bench_cfg5_c.txt 2> gpurun_out/bench_cfg5_c.err)
for v in h1 h2; do
  (JSGPU_LIB=$PWD/jpegsnoop_b200/variants/libjsgpu_$v.so timeout 200 python bench.py --configs none --no-e2e --no-cpu --steps 10 --warmup 3 > gpurun_out/var_$v.txt 2> gpurun_out/var_$v.err)
  (JSGPU_LIB=$PWD/jpegsnoop_b200/variants/libjsgpu_$v.so timeout 200 python bench.py --config cfg5 --configs none --no-e2e --no-cpu --steps 5 --warmup 2 > gpurun_out/var5_$v.txt 2> gpurun_out/var5_$v.err)
done
(timeout 200 python bench.py --config cfg3shard --configs none --no-e2e --no-cpu --steps 5 --warmup 2 > gpurun_out/bench_cfg3_c.txt 2>&1)
grep -E "passed|failed" gpurun_out/pytest_gpu5.txt | tail -n 3
