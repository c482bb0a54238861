#!/bin/bash
mkdir -p gpurun_out
(timeout 800 compute-sanitizer --tool memcheck --error-exitcode 77 --print-limit 20 python -m pytest tests/test_gpu_parity.py tests/test_gpu_detail.py -q -x -m gpu -k "auto or detailed or host_marker or unsupported or dc_only" > gpurun_out/san_c.log 2>&1; echo "rc_c=$?" >> gpurun_out/san_c.log)
grep -E "ERROR SUMMARY|rc_c|passed|failed|Invalid|out of bounds|deselected" gpurun_out/san_c.log | head -20
