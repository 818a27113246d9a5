#!/bin/bash
# final profiling passes (one ncu invocation per gpurun call): `list` = per-launch device times of an un-graphed step,
# `full` = ncu --set full of the step's kernels + the stand-alone micro-benchmarks
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --quick --no-cpu-baseline --no-e2e --no-graph --sustain-s 0"
if [ "$1" = "list" ]; then
  $CMD > gpurun_out/y_plain.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/y_launches.csv $CMD > gpurun_out/y_ncu1.log 2>&1
  echo "launch list rc=$?"
  # the witness stream's write pattern in isolation (tools/store_pattern.cu, built into variants/ beforehand)
  if [ -x variants/store_pattern ]; then for w in 60 56 64; do timeout 120 variants/store_pattern $w; done > gpurun_out/y_store_pattern.txt 2>&1; fi
  timeout 300 python tools/cluster_bench.py > gpurun_out/y_cluster_bench.txt 2>&1
  timeout 300 python tools/tc_timeline.py 1024 > gpurun_out/y_timeline.txt 2>&1
else
  $CMD > gpurun_out/y_plain2.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"fr_matmul_tc_kernel|tc_split|rescale_|mat_vec_prefix|gamma_powers" -c 11 -o gpurun_out/y_prof $CMD > gpurun_out/y_ncu2.log 2>&1
  echo "full capture rc=$?"
fi
ls -la gpurun_out/y_*
