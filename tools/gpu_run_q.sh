#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_host_mirror.py -m gpu -q -x -k "tail_split or small_operand or bulk_hand_off or graph" > gpurun_out/q_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/q_tests.log
timeout 300 python tools/tc_timeline.py 1024 2>&1 | head -8
timeout 300 python tools/round2_sweep.py 2>&1 | grep "mat-mul call" 
timeout 600 python bench.py --steps 10 --warmup 3 --quick --no-cpu-baseline > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --steps 5 --warmup 3 --quick --no-cpu-baseline --no-e2e --shape 4096,2048,4096 > gpurun_out/q_cfg4_n1.json 2> gpurun_out/q_cfg4_n1.err; echo "cfg4 1-GPU rc=$?"
python - <<'PY'
import json
def load(f): return json.loads([l for l in open(f) if l.startswith('{')][-1])
for f in ('q_bench','q_cfg4_n1'):
    try:
        d=load(f'gpurun_out/{f}.json'); print(f, round(d['ms_per_step'],4), 'ms', f"{d['value']:.3e}", {k:round(v['ms'],4) for k,v in d['roofline']['kernels'].items() if 'ms' in v})
    except Exception as e: print(f,'failed',e)
PY
