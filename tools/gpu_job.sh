#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tiff_export.py -x -q -m gpu 2>&1 | tail -n 6
timeout 1500 python bench.py 2> gpurun_out/bench_full3.err | tail -n 1 > gpurun_out/bench_full3.json
cut -c1-1500 gpurun_out/bench_full3.json; tail -n 5 gpurun_out/bench_full3.err
