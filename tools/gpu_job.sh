#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_preview.py -x -q 2>&1 | tail -n 30
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_preview.py 2>&1 | tail -n 4
