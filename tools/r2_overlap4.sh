#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for rep in 1 2; do
for f in 0 1 2; do
echo "== density solve forked at point $f (0 after the viscosity solves, 1 after the first projection, 2 after advect)"
SF_DEV_DENS_FORK=$f timeout 300 python tools/step_ab.py 8192 40 auto=5:0 chunk256=5:256 chunk512=5:512
done; done > gpurun_out/b4_fork_point.log 2>&1; cat gpurun_out/b4_fork_point.log
