#!/bin/bash
# session 6: final build -- spot A/B of what changed, ncu --set full of the cluster kernel with stall reasons, final bench lines
OUT=gpurun_out/s6
mkdir -p $OUT
export PYTHONPATH=$PWD:$PWD/tests
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "round2 or compiled or smooth or config3 or config4 or native_library" > $OUT/pytest_new.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_new.log; tail -3 $OUT/pytest_new.log
timeout 400 python tools/ab_plans.py 1234 2053 617 1009 1000000 390625 200000 44100 16384 32768 65536 > $OUT/ab_plans.txt 2>&1; cut -c1-180 $OUT/ab_plans.txt
timeout 300 /usr/local/cuda/bin/ncu --set full --clock-control none --import-source on -k regex:run_cluster -s 4 -c 1 -o /tmp/full_cluster python tools/ab_plans.py 65536 > $OUT/ncu_cluster.log 2>&1
python tools/ncu_summary.py /tmp/full_cluster.ncu-rep > $OUT/ncu_full_cluster_256x256.md 2>&1; grep -E "stall|duration|issue" $OUT/ncu_full_cluster_256x256.md | cut -c1-150
timeout 400 /usr/local/cuda/bin/ncu --set full --clock-control none --import-source on -k regex:run_fused -c 1 -o /tmp/full_fused python bench.py --profile --steps 1 --logs 20 > $OUT/ncu_fused.log 2>&1
python tools/ncu_summary.py /tmp/full_fused.ncu-rep > $OUT/ncu_full_fused_1024x1024.md 2>&1; grep -E "stall|duration|issue" $OUT/ncu_full_fused_1024x1024.md | cut -c1-150
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/s6/bench.json'))
print("value",d["value"],"frac",d["roofline"]["frac"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
print("per_size", [(r["log2n"], r["frac"], r.get("frac_b2b")) for r in d["config"]["per_size"]])
for r in d.get("other_configs") or []:
    r=dict(r); r.pop("per_size",None); print(json.dumps(r)[:900])
PY
