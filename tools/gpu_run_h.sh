#!/bin/bash
# launch lists (ncu, gpu__time_duration) for the headline step and for one rank's share of the 8-GPU job (128-row slab)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --quick --no-cpu-baseline --no-e2e --no-graph --sustain-s 0"
$CMD > gpurun_out/h_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/h_launches.csv $CMD > gpurun_out/h_ncu1.log 2>&1
echo "launch list rc=$?"
CMD2="python bench.py --steps 2 --warmup 3 --quick --no-cpu-baseline --no-e2e --no-graph --sustain-s 0 --shape 128,1024,1024"
$CMD2 > gpurun_out/h_plain2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/h_launches_slab.csv $CMD2 > gpurun_out/h_ncu2.log 2>&1
echo "slab launch list rc=$?"
tail -2 gpurun_out/h_plain2.log | cut -c1-300
