#!/bin/bash
# session 14: last check of HEAD -- full GPU suite, smoke(), a short bench line
OUT=gpurun_out/s14
mkdir -p $OUT
export PYTHONPATH=$PWD:$PWD/tests
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('$OUT/bench.json')); print('value', d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'])"
