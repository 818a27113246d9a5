#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/f_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/f_summary.txt
timeout 400 python tools/round2_sweep.py > gpurun_out/f_sweep.txt 2>&1; echo "sweep rc=$?" >> gpurun_out/f_summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 --quick --no-cpu-baseline > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?" >> gpurun_out/f_summary.txt
cat gpurun_out/f_summary.txt; tail -5 gpurun_out/f_tests.log; cat gpurun_out/f_sweep.txt; tail -c 300 gpurun_out/f_bench.err
