#!/bin/bash
# Round 2, final one-GPU session: GPU suite, smoke, the default bench line, L1-SVM at 400 000 samples
tag=r02zz
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -x -q --durations=5 > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -3 $out/${tag}_pytest_gpu.log
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
echo "bench exit $?"
python - <<P
import json
d=json.loads(open("$out/${tag}_bench_n1.json").read().strip().splitlines()[-1])
print(d["value"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["call_seconds"], d["e2e"]["call_phases_rank0"], d["parity"]["status"])
print(d["latency_bound_configs"])
for k,v in (d["secondary_workloads"] or {}).items(): print(k, v.get("value"), v.get("parity",{}).get("status"), v.get("roofline",{}).get("frac"), v.get("e2e",{}).get("value"), v.get("error"))
print({k: (v.get("iterations_per_s"), v.get("frac_of_peak_actual_bytes")) for k,v in d["variants"].items()})
print(d["cpu_baseline"])
P
timeout 600 python tools/quick_bench.py --kind l1svm --size 400000 --iters 10 --reps 3 > $out/${tag}_l1svm_400k.jsonl 2> $out/${tag}_l1svm_400k.err
echo "l1svm 400k exit $?"; cut -c1-330 $out/${tag}_l1svm_400k.jsonl; tail -2 $out/${tag}_l1svm_400k.err
