#!/bin/bash
# multi-GPU session: real-device peer-slab tests, then bench.py under torchrun at the given rank counts
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NS=${1:-2}
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_peer_slab_gpu.py -q -m gpu -x 2>&1 | tail -4
port=$((29600 + RANDOM % 200))
for n in $NS; do
  port=$((port+1))
  out=gpurun_out/m_bench_n${n}.json
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --steps 10 --warmup 3 > $out 2> ${out%.json}.err ) 2>&1 | grep real
  python - "$out" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(f"n={d['n_gpus']} G={d['config']['grid']}: {d['ms_per_step']:.2f} ms/step {d['value']:.4g} upd/s launches {d['gpu_launches']}")
    print(" parity", d.get('parity'))
    print(" e2e", {k:(v if not isinstance(v,str) else v[:60]) for k,v in d.get('e2e',{}).items()})
    print(" extra", json.dumps(d.get('extra'))[:900])
except Exception as e:
    print("FAILED", sys.argv[1], e); print(open(sys.argv[1][:-5]+'.err').read()[-1500:])
PY
done
if [ -n "$PHASES" ]; then
  for n in $PHASES; do
    port=$((port+1))
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
        tools/slab_phase_times.py 32768 40 > gpurun_out/m_phases_n${n}.log 2>&1
    tail -25 gpurun_out/m_phases_n${n}.log
  done
fi
