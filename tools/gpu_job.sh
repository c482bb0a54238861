#!/bin/bash
# GPU-box job of the moment (experiment helper, not product).  Everything it writes goes to gpurun_out/.
mkdir -p gpurun_out
(timeout 400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu3.txt 2>&1; echo rc=$? >> gpurun_out/pytest_gpu3.txt)
(timeout 400 ncu --set full --import-source on --clock-control none -k regex:"k_ph_sync|k_huff_lane" -s 15 -c 4 -o gpurun_out/r2a_cfg5_huff -f python bench.py --config cfg5 --batch 64 --configs none --no-e2e --no-cpu --steps 1 --warmup 1 > gpurun_out/ncu_r2a.log 2>&1)
for v in w0 w1 w2 w3 w4 w5 w6 w7; do
  (JSGPU_LIB=$PWD/jpegsnoop_b200/variants/libjsgpu_$v.so timeout 200 python bench.py --configs none --no-e2e --no-cpu --steps 10 --warmup 3 > gpurun_out/var_$v.txt 2> gpurun_out/var_$v.err)
done
tail -n 4 gpurun_out/pytest_gpu3.txt
