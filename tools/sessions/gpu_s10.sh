#!/bin/bash
# session 10: final build (ring rule for 8 MiB transforms): two-pass parity, A/B line, bench
OUT=gpurun_out/s10
mkdir -p $OUT
export PYTHONPATH=$PWD:$PWD/tests
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config2 or power_of_two or fresh_process or device_path or threads" > $OUT/pytest_two_pass.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_two_pass.log; tail -3 $OUT/pytest_two_pass.log
timeout 300 python tools/ab_two_pass.py 15,16,17,18,19,20 > $OUT/ab_default.txt 2>&1; tail -1 $OUT/ab_default.txt
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/s10/bench.json'))
print("value",d["value"],"frac",d["roofline"]["frac"], "e2e", d["e2e"]["value"])
print("per_size", [(r["log2n"], r["plan"], r["frac"]) for r in d["config"]["per_size"]][-4:])
PY
