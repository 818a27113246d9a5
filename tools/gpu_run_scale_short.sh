#!/bin/bash
# 1 -> 8 GPU scaling of the headline step only (the way the driver launches it) + BASELINE configs[4] on 8 GPUs
mkdir -p gpurun_out
rm -f gpurun_out/s_summary.txt
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --quick --no-cpu-baseline > gpurun_out/s_bench_n$N.json 2> gpurun_out/s_bench_n$N.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 20 --warmup 5 --quick --no-cpu-baseline > gpurun_out/s_bench_n$N.json 2> gpurun_out/s_bench_n$N.err
  fi
  echo "N=$N rc=$?" >> gpurun_out/s_summary.txt
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus 8 --steps 10 --warmup 3 --quick --no-cpu-baseline --no-e2e --shape 4096,2048,4096 > gpurun_out/s_cfg4_n8.json 2> gpurun_out/s_cfg4_n8.err
echo "cfg4 N=8 rc=$?" >> gpurun_out/s_summary.txt
if [ -x variants/store_pattern ]; then for w in 60 56 64; do timeout 120 variants/store_pattern $w; done > gpurun_out/y_store_pattern.txt 2>&1; fi
cat gpurun_out/s_summary.txt
python - <<'PY'
import json
def load(f): return json.loads([l for l in open(f) if l.startswith('{')][-1])
for n in (1,2,4,8):
    try:
        d=load(f'gpurun_out/s_bench_n{n}.json'); k=d['roofline']['kernels']
        print(n, round(d['ms_per_step'],4), 'ms', f"{d['value']:.3e}", 'e2e ms', round(d['e2e']['ms_per_step'],2), 'mm', round(k['fr_matmul']['ms'],4), 'rescale', round(k['rescale_kernel']['ms'],4), 'matvec', round(k['mat_vec_prefix']['ms'],4), d['gpu_launches'], d['roofline']['verified']['ranks'])
    except Exception as e: print(n, 'failed', e)
try:
    d=load('gpurun_out/s_cfg4_n8.json'); print('cfg4 n8', round(d['ms_per_step'],3), 'ms', f"{d['value']:.3e}")
except Exception as e: print('cfg4 failed', e)
PY
