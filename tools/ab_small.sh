#!/bin/bash
# A/B of libquanta_b200.so builds kept under quanta_b200/csrc/build/variants/lib<tag>.so (git-ignored, they travel
# with the gpurun snapshot): for each tag, parity of the small-batch GEMM, then tools/bench_gemm.py.
#   tools/ab_small.sh "0 A B" [formats] [ms] [shapes]
cd "$(dirname "$0")/.."
TAGS=${1:-"0"}; FMT=${2:-4}; MS=${3:-1,8,16}; SH=${4:-4096x14336,14336x4096}
mkdir -p gpurun_out
cp quanta_b200/libquanta_b200.so /tmp/lib_keep.so
for t in $TAGS; do
  cp quanta_b200/csrc/build/variants/lib$t.so quanta_b200/libquanta_b200.so
  echo "=== variant $t" | tee -a gpurun_out/ab_small.log
  timeout 300 python -m pytest tests/test_gpu_gemm.py -x -q -m gpu -k "small or wna16" 2>&1 | tail -2 | tee -a gpurun_out/ab_small.log
  timeout 200 python tools/bench_gemm.py --formats $FMT --ms $MS --shapes $SH 2>&1 | tee -a gpurun_out/ab_small.log
done
cp /tmp/lib_keep.so quanta_b200/libquanta_b200.so
