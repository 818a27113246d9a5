#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "cluster or tile_width or small_operand" > gpurun_out/r_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r_tests.log
timeout 600 python tools/cluster_bench.py > gpurun_out/r_cluster.log 2>&1; echo "cluster bench rc=$?"; cat gpurun_out/r_cluster.log
timeout 300 python tools/tc_timeline.py 1024 2>&1 | head -8
