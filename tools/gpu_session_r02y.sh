#!/bin/bash
# Round 2, session y (1 GPU): full GPU suite with the cluster kernel as the default persistent path (+ locality
# numbering), small-config timings, ncu of the cluster kernel and of the pipelined long-row kernel
tag=r02y
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -x -q --durations=5 > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -4 $out/${tag}_pytest_gpu.log
timeout 300 python tools/small_bench.py 20000 > $out/${tag}_small.jsonl 2> $out/${tag}_small.err
echo "small exit $?"; cat $out/${tag}_small.jsonl | cut -c1-420
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_cluster_iterate --launch-skip 1 -c 1 -f -o $out/${tag}_cluster \
  python tools/small_bench.py 3000 > $out/${tag}_ncu_cluster.log 2>&1
echo "ncu cluster exit $?"
timeout 600 ncu --set full --clock-control none -k regex:'k_long_partial|k_long_finish' --launch-skip 20 -c 2 \
  -f -o $out/${tag}_long python tools/quick_bench.py --kind l1svm --size 50000 --iters 4 --reps 1 > $out/${tag}_ncu_long.log 2>&1
echo "ncu long exit $?"
