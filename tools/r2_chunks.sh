#!/bin/bash
mkdir -p gpurun_out/r2
python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | tail -8
