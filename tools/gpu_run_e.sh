#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "tile_width or small_operand" > gpurun_out/e_tests_new.log 2>&1; echo "new tests rc=$?" > gpurun_out/e_summary.txt
timeout 400 python tools/round2_sweep.py > gpurun_out/e_sweep.txt 2>&1; echo "sweep rc=$?" >> gpurun_out/e_summary.txt
cat gpurun_out/e_summary.txt; tail -3 gpurun_out/e_tests_new.log; cat gpurun_out/e_sweep.txt
