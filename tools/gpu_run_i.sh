#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/i_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/i_summary.txt
timeout 400 python tools/round2_sweep.py > gpurun_out/i_sweep.txt 2>&1; echo "sweep rc=$?" >> gpurun_out/i_summary.txt
timeout 200 python tools/tc_timeline.py 1024 > gpurun_out/i_timeline.txt 2>&1
cat gpurun_out/i_summary.txt; tail -4 gpurun_out/i_tests.log; cat gpurun_out/i_sweep.txt; head -8 gpurun_out/i_timeline.txt
