#!/bin/bash
OUT=gpurun_out/s12
mkdir -p $OUT
export PYTHONPATH=$PWD:$PWD/tests
nproc > $OUT/host.txt; free -g >> $OUT/host.txt; lscpu | head -20 >> $OUT/host.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "host or device_path or threads or error_behaviour" > $OUT/pytest_host.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_host.log
tail -n 3 $OUT/pytest_host.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
tail -c 1500 $OUT/bench_ref.json
timeout 900 python bench.py --no-extras > $OUT/bench.json 2> $OUT/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s12/bench.json'))
print("value",d["value"],"frac",d["roofline"]["frac"])
print("e2e",json.dumps(d["e2e"]))
print("cpu",json.dumps({k:v for k,v in d["cpu_baseline"].items() if k!="per_size"}))
PY
