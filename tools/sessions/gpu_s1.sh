#!/bin/bash
# session 1 of this re-entry: state check -- full GPU suite, both bench arms
OUT=gpurun_out/s1
mkdir -p $OUT
export PYTHONPATH=$PWD:$PWD/tests
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.draw,memory.total --format=csv > $OUT/env.txt 2>&1
nproc >> $OUT/env.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/s1/bench.json'))
print("value",d["value"],"frac",d["roofline"]["frac"])
print("per_size", json.dumps(d["config"].get("per_size")))
print("e2e",json.dumps(d["e2e"]))
print("cpu",json.dumps({k:v for k,v in d["cpu_baseline"].items() if k!="per_size"}))
print("other", json.dumps(d.get("other_configs")))
PY
tail -c 600 $OUT/bench_reference.json
