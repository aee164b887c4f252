#!/bin/bash
# Strong-scaling runs on one box.  usage: scale_run.sh "<G> <K> <steps>" "1 2 4 8" [extra bench flags]   (run under gpurun --gpus 8)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
set -- $1 "$2"
G=$1; K=$2; S=$3; NS=$4; EXTRA=$5
port=$((29600 + RANDOM % 200))
for n in $NS; do
  port=$((port+1))
  out=gpurun_out/scale_G${G}_K${K}_n${n}.json
  if [ "$n" = "1" ]; then
    python bench.py --grid $G --iters $K --steps $S --warmup 3 --skip-extras > $out 2> ${out%.json}.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --grid $G --iters $K --steps $S --warmup 3 $EXTRA > $out 2> ${out%.json}.err
  fi
  python - "$out" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(f"n={d['n_gpus']} G={d['config']['grid']} K={d['config']['iters']}: {d['ms_per_step']:.2f} ms/step  {d['value']:.4g} upd/s  launches {d['gpu_launches']} clocks {d['clocks']}")
except Exception as e:
    print("FAILED", sys.argv[1], e)
PY
done
