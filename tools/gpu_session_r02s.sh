#!/bin/bash
# Round 2, session s (1 GPU): GPU suite, small-config timings with the relaxed cluster barrier (and strict, and the
# cluster kernel forced onto the one-CTA LPs), set-up phases of the reordered layout, ncu of the cluster kernel
tag=r02s
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -x -q --durations=5 > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -4 $out/${tag}_pytest_gpu.log
timeout 300 python tools/small_bench.py > $out/${tag}_small.jsonl 2> $out/${tag}_small.err
echo "small exit $?"; cat $out/${tag}_small.jsonl | cut -c1-420; tail -3 $out/${tag}_small.err
echo "-- strict barrier"
CPPPD_CLUSTER_MODE=1 timeout 300 python tools/small_bench.py 5000 2>&1 | tail -5 | tee $out/${tag}_small_strict.jsonl
echo "-- cluster forced"
CPPPD_FORCE_CLUSTER=1 timeout 300 python tools/small_bench.py 5000 2>&1 | tail -7 | cut -c1-420 | tee $out/${tag}_small_forced.jsonl
echo "-- setup phases (flags 8: reordered layout on one GPU)"
CPPPD_SETUP_TIMING=1 timeout 300 python tools/quick_bench.py --size 4096 --flags 8 --iters 5 --reps 1 2>&1 | grep "cpppd setup" | tee $out/${tag}_setup_phases.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_cluster_iterate --launch-skip 1 -c 1 -f -o $out/${tag}_cluster \
  python tools/small_bench.py 3000 > $out/${tag}_ncu_cluster.log 2>&1
echo "ncu exit $?"
