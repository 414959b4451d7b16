#!/bin/bash
# Round 2, session p (1 GPU): long-row kernel shapes / segment order / segment length on the L1-SVM LP
tag=r02p
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
timeout 600 python tools/long_sweep.py --size 100000 > $out/${tag}_long_sweep.jsonl 2> $out/${tag}_long_sweep.err
echo "sweep exit $?"
tail -3 $out/${tag}_long_sweep.err
cat $out/${tag}_long_sweep.jsonl | cut -c1-200
