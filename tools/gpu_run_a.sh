#!/bin/bash
# first GPU pass of round 2: regression tests, new tests, smoke, bench (each in its own process)
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_round2.py > gpurun_out/a_tests_old.log 2>&1; echo "old tests rc=$?" >> gpurun_out/a_summary.txt
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q > gpurun_out/a_tests_new.log 2>&1; echo "new tests rc=$?" >> gpurun_out/a_summary.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/a_summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?" >> gpurun_out/a_summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 --matmul-small 0 --quick --no-cpu-baseline > gpurun_out/a_bench_full.json 2> gpurun_out/a_bench_full.err; echo "bench(full engine) rc=$?" >> gpurun_out/a_summary.txt
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/a_bench_ref.json 2> gpurun_out/a_bench_ref.err; echo "bench(ref) rc=$?" >> gpurun_out/a_summary.txt
cat gpurun_out/a_summary.txt
tail -5 gpurun_out/a_tests_old.log; tail -30 gpurun_out/a_tests_new.log; tail -3 gpurun_out/a_smoke.log; tail -c 600 gpurun_out/a_bench.err
