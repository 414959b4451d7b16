#!/bin/bash
# Round 2, session i (1 GPU): the full GPU suite on the final library, the default bench line, L1-SVM numbers, ncu evidence
tag=r02i
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
log=$out/${tag}_session.log
echo "== 1. GPU tests" | tee $log
timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a $log
tail -3 $out/${tag}_pytest_gpu.log | tee -a $log
echo "== 2. smoke + default bench (N = 1)" | tee -a $log
timeout 120 python __graft_entry__.py --smoke >> $log 2>&1
timeout 900 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
echo "bench exit $?" | tee -a $log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
echo "reference arm exit $?" | tee -a $log
echo "== 3. L1-SVM (configs[2] family): 100 000 and 200 000 samples x 1 000 features" | tee -a $log
for size in 100000 200000; do
  timeout 900 python tools/quick_bench.py --kind l1svm --size $size --iters 10 --reps 3 >> $out/${tag}_l1svm.jsonl 2>> $out/${tag}_l1svm.err
  echo "l1svm $size exit $?" | tee -a $log
done
timeout 900 python bench.py --workload l1svm --size 100000 --steps 4 --iters-per-step 10 --e2e-steps 1 --e2e-iters 50 --no-cpu-baseline --small-configs 0 \
  > $out/${tag}_bench_l1svm.json 2> $out/${tag}_bench_l1svm.err
echo "bench l1svm exit $?" | tee -a $log
echo "== 3b. compressed storage (flags 3 / 11) on the Potts LP: every kernel variant" | tee -a $log
for f in 3 11; do for v in 1 2 3 4 5 6 7; do
  timeout 300 python tools/quick_bench.py --size 4096 --iters 100 --reps 3 --variant $v --flags $f >> $out/${tag}_variants.jsonl 2>> $out/${tag}_variants.err
done; done
echo "== 4. ncu launch list of the default bench command (headline only)" | tee -a $log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
  python bench.py --steps 1 --warmup 3 --iters-per-step 10 --e2e-steps 0 --variants 0 --no-cpu-baseline --small-configs 0 --secondary '' > $out/${tag}_ncu_bench.log 2>&1
echo "launch list exit $?" | tee -a $log
echo "== 5. ncu --set full: banded window kernels on the random LP, long-row kernels on the L1-SVM LP" | tee -a $log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_primal_band|k_dual_band' --launch-skip 900 -c 10 \
  -f -o $out/${tag}_band python tools/quick_bench.py --kind random --size 20000000 --iters 4 --reps 1 > $out/${tag}_ncu_band.log 2>&1
echo "ncu band exit $?" | tee -a $log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_long_partial|k_long_finish|k_primal|k_dual' --launch-skip 150 -c 8 \
  -f -o $out/${tag}_l1svm python tools/quick_bench.py --kind l1svm --size 50000 --iters 4 --reps 1 > $out/${tag}_ncu_l1svm.log 2>&1
echo "ncu l1svm exit $?" | tee -a $log
echo "== done" | tee -a $log
