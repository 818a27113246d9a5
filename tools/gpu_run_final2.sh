#!/bin/bash
# final profiling pass: per-launch device times + ncu --set full of the step's kernels (un-graphed so that every kernel is listed)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --quick --no-cpu-baseline --no-e2e --no-graph --sustain-s 0"
$CMD > gpurun_out/y_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/y_launches.csv $CMD > gpurun_out/y_ncu1.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/y_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"fr_matmul_tc_kernel|tc_split|rescale_kernel|mat_vec_prefix|gamma_powers" -c 11 -o gpurun_out/y_prof $CMD > gpurun_out/y_ncu2.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/y_*
