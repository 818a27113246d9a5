#!/bin/bash
# final single-GPU validation: every GPU test, smoke, the default bench line, the reference arm, the one-process multi-handle mode
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/z_tests.log 2>&1; echo "gpu tests rc=$?" > gpurun_out/z_summary.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/z_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/z_summary.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/z_bench_n1.json 2> gpurun_out/z_bench_n1.err; echo "bench rc=$?" >> gpurun_out/z_summary.txt
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/z_bench_ref.json 2> gpurun_out/z_bench_ref.err; echo "bench(ref) rc=$?" >> gpurun_out/z_summary.txt
timeout 600 python bench.py --multi-handle --gpus 2 --steps 5 --warmup 3 > gpurun_out/z_bench_multi2_on1.json 2> gpurun_out/z_bench_multi.err; echo "bench(multi-handle, 2 handles on 1 GPU) rc=$?" >> gpurun_out/z_summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 --matmul-small 0 --quick --no-cpu-baseline > gpurun_out/z_bench_fullwidth.json 2> gpurun_out/z_bench_fullwidth.err; echo "bench(full-width engine forced) rc=$?" >> gpurun_out/z_summary.txt
cat gpurun_out/z_summary.txt; tail -4 gpurun_out/z_tests.log; tail -2 gpurun_out/z_smoke.log
