#!/bin/bash
# 1 -> 8 GPU scaling of the headline step, the way the driver launches it
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/s_gpus.txt 2>&1
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --quick --no-cpu-baseline > gpurun_out/s_bench_n$N.json 2> gpurun_out/s_bench_n$N.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 20 --warmup 5 --quick --no-cpu-baseline > gpurun_out/s_bench_n$N.json 2> gpurun_out/s_bench_n$N.err
  fi
  echo "N=$N rc=$?" >> gpurun_out/s_summary.txt
done
cat gpurun_out/s_summary.txt
python - <<'PY'
import json
for n in (1,2,4,8):
    try:
        d=json.loads([l for l in open(f'gpurun_out/s_bench_n{n}.json') if l.startswith('{')][-1])
        k=d['roofline']['kernels']
        print(n, round(d['ms_per_step'],4), 'ms', f"{d['value']:.3e}", 'e2e ms', round(d['e2e']['ms_per_step'],2), 'mm', round(k['fr_matmul']['ms'],4), 'rescale', round(k['rescale_kernel']['ms'],4), 'matvec', round(k['mat_vec_prefix']['ms'],4), d['gpu_launches'])
    except Exception as e:
        print(n, 'failed', e)
PY
