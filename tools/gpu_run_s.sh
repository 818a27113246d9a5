#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "rescale or mul_witness or smoke" > gpurun_out/s_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/s_tests.log
for t in "rescale_store=3,rescale_fast_sums=0" "rescale_store=3,rescale_fast_sums=1" "rescale_store=0,rescale_fast_sums=0" "rescale_store=0,rescale_fast_sums=1"; do TUNE=$t timeout 120 python tools/rescale_variant.py 2>&1 | tail -1; done
for t in "rescale_store=3" "rescale_store=0"; do H2SVD_LIB=variants/libh2svd_storeonly.so TUNE=$t timeout 120 python tools/rescale_variant.py 2>&1 | tail -1; done
timeout 600 python bench.py --steps 10 --warmup 3 --quick --no-cpu-baseline > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
def load(f): return json.loads([l for l in open(f) if l.startswith('{')][-1])
d=load('gpurun_out/s_bench.json'); print(round(d['ms_per_step'],4), 'ms', f"{d['value']:.3e}", {k:round(v['ms'],4) for k,v in d['roofline']['kernels'].items() if 'ms' in v})
PY
