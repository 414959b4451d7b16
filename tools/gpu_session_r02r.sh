#!/bin/bash
# Round 2, session r (1 GPU): full GPU suite with the cluster kernel + pipelined long rows, small-config timings
tag=r02r
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -x -q --durations=5 > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -4 $out/${tag}_pytest_gpu.log
timeout 300 python tools/small_bench.py > $out/${tag}_small.jsonl 2> $out/${tag}_small.err
echo "small exit $?"; cat $out/${tag}_small.jsonl; tail -3 $out/${tag}_small.err
