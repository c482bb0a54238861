#!/bin/bash
mkdir -p gpurun_out
which compute-sanitizer || ls /usr/local/cuda/bin | grep -i sanit
(timeout 900 compute-sanitizer --tool memcheck --error-exitcode 77 --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "single_image_dropin and auto and idct_fixed or unusual and auto or damaged_scans and auto_selfsync" > gpurun_out/san_a.log 2>&1; echo "rc_a=$?" >> gpurun_out/san_a.log)
grep -E "ERROR SUMMARY|rc_a|passed|failed|Invalid|out of bounds" gpurun_out/san_a.log | head -20
(timeout 600 compute-sanitizer --tool memcheck --error-exitcode 77 --print-limit 20 python -m pytest tests/test_gpu_preview.py tests/test_tiff_export.py -q -x -m gpu > gpurun_out/san_b.log 2>&1; echo "rc_b=$?" >> gpurun_out/san_b.log)
grep -E "ERROR SUMMARY|rc_b|passed|failed|Invalid|out of bounds" gpurun_out/san_b.log | head -20
