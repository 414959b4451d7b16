#!/bin/bash
# Round 2, session v (1 GPU): cluster kernel modes (0 relaxed / 1 strict / 2 gathers from L2): parity + it/s
tag=r02v
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_zz_opt_in_features.py tests/test_gpu_variants.py -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -4 $out/${tag}_pytest_gpu.log
for mode in 0 1 2; do
  echo "-- CPPPD_CLUSTER_MODE=$mode"
  CPPPD_CLUSTER_MODE=$mode timeout 300 python tools/small_bench.py 5000 2>&1 | tail -5 | tee -a $out/${tag}_small_mode$mode.jsonl
done
