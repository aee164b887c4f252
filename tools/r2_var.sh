#!/bin/bash
cd "$(dirname "$0")/.."
for i in 1 2 3; do python tools/step_ab.py 8192 40 default= overlap_only=16:0 default_again= seq=19:0; done > gpurun_out/b5_variance.log 2>&1; cat gpurun_out/b5_variance.log
