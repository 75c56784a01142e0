#!/bin/bash
# experiment sweep for the small-batch GEMM: one shape, debug switches and tuning knobs via env.
# Every run is wrapped in its own short timeout: an experiment switch that deadlocks must not eat the GPU budget.
run() { echo "== $*"; env "$@" timeout 60 python tools/bench_gemm.py --ms ${MS:-16} --formats ${FMT:-4} --shapes ${SHAPES:-4096x14336} 2>&1 | grep -o '"M": [0-9]*.*"us": [0-9.]*' ; }
for v in "$@"; do run $v; done
