#!/bin/bash
# Build the A/B variants of tools/run5.sh into build/ (git-ignored, travels with the gpurun snapshot).  ~2 min each, run in the build container.
cd "$(dirname "$0")/.."
set -e
b() { out=$1; shift; [ -f build/libsf_$out.so ] && [ build/libsf_$out.so -nt fluidsimulationcuda_b200/csrc/sf_jacobi.cu ] || python -m fluidsimulationcuda_b200.build --out build/libsf_$out.so "$@"; }
b c3     -DSF_PRESSURE_CTAS=3
b il1    -DSF_INNER_LOOP=1
b il1c3  -DSF_INNER_LOOP=1 -DSF_PRESSURE_CTAS=3
b il2    -DSF_INNER_LOOP=2
b il2c3  -DSF_INNER_LOOP=2 -DSF_PRESSURE_CTAS=3
b il2c3e -DSF_INNER_LOOP=2 -DSF_PRESSURE_CTAS=3 -DSF_EDGE_SPLIT=1
b gg     -DSF_GUARDED_GROUP=1
b ggil2  -DSF_GUARDED_GROUP=1 -DSF_INNER_LOOP=2
ls -la build/
