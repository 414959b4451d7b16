#!/bin/bash
# Round 2, session j (1 GPU): grid-stride variants (8, 9), unrolled long-row kernel, e2e after the solver-stream fix
tag=r02j
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
log=$out/${tag}_session.log
echo "== 1. GPU tests: variants, skewed rows" | tee $log
timeout 900 python -m pytest tests/test_gpu_variants.py tests/test_gpu_skewed_rows.py -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a $log
tail -2 $out/${tag}_pytest_gpu.log | tee -a $log
echo "== 2. Potts 4096^2: variants 2, 8, 9 with generic / compressed storage" | tee -a $log
for f in 0 3 11; do for v in 2 8 9; do
  timeout 300 python tools/quick_bench.py --size 4096 --iters 100 --reps 3 --variant $v --flags $f >> $out/${tag}_variants.jsonl 2>> $out/${tag}_variants.err
done; done
echo "== 3. L1-SVM 100 000 / 400 000 samples" | tee -a $log
timeout 600 python tools/quick_bench.py --kind l1svm --size 100000 --iters 10 --reps 3 >> $out/${tag}_l1svm.jsonl 2>> $out/${tag}_l1svm.err
echo "l1svm 100000 exit $?" | tee -a $log
timeout 1200 python tools/quick_bench.py --kind l1svm --size 400000 --iters 6 --reps 2 >> $out/${tag}_l1svm.jsonl 2>> $out/${tag}_l1svm.err
echo "l1svm 400000 exit $?" | tee -a $log
echo "== 4. headline bench with e2e" | tee -a $log
timeout 600 python bench.py --steps 10 --variants 0 --small-configs 0 --secondary '' --no-cpu-baseline > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
echo "bench exit $?" | tee -a $log
echo "== done" | tee -a $log
