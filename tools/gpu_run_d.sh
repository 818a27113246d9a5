#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -x > gpurun_out/d_tests_new.log 2>&1; echo "new tests rc=$?" > gpurun_out/d_summary.txt
timeout 600 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_round2.py > gpurun_out/d_tests_old.log 2>&1; echo "old tests rc=$?" >> gpurun_out/d_summary.txt
timeout 300 python tools/round2_sweep.py > gpurun_out/d_sweep.txt 2>&1; echo "sweep rc=$?" >> gpurun_out/d_summary.txt
timeout 200 python tools/tc_timeline.py 1024 > gpurun_out/d_timeline.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --quick --no-cpu-baseline > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; echo "bench rc=$?" >> gpurun_out/d_summary.txt
cat gpurun_out/d_summary.txt; tail -3 gpurun_out/d_tests_new.log; tail -3 gpurun_out/d_tests_old.log; cat gpurun_out/d_sweep.txt
