#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/b1_ref.json 2> gpurun_out/b1_ref.err; tail -c 400 gpurun_out/b1_ref.json; echo
( time python bench.py --steps 20 --warmup 5 > gpurun_out/b1_bench.json 2> gpurun_out/b1_bench.err ) 2>&1 | tail -3
python - <<'PY'
import json
d=json.load(open('gpurun_out/b1_bench.json'))
print('ms/step', d['ms_per_step'], 'launches', d['gpu_launches'], 'e2e', d['e2e'].get('ms_per_step'))
print('roofline', {k:d['roofline'][k] for k in ('achieved','frac','avg_launch_ms','traffic','traffic_capture')})
print('extra', json.dumps(d.get('extra'), indent=1)[:3000])
print('scaling_base', d.get('scaling_base'))
PY
tail -5 gpurun_out/b1_bench.err
