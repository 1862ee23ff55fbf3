#!/bin/bash
# session 3: cluster single-pass plans and compiled composite tiles -- parity, A/B, one ncu --set full of the cluster kernel, bench
OUT=gpurun_out/s3
mkdir -p $OUT
export PYTHONPATH=$PWD:$PWD/tests
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "round2 or smooth or large_non_power or native_library" > $OUT/pytest_new.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_new.log; tail -4 $OUT/pytest_new.log
AB_PLANS_OUT=$OUT/ab_cluster.json timeout 600 python tools/ab_plans.py 16384 32768 65536 131072 10000 44100 48000 100000 1000000 617 1234 97 > $OUT/ab_plans.txt 2>&1; echo "ab rc=$?"; grep -v complex128 $OUT/ab_plans.txt
B200FFT_SMOOTH_COMPILED=0 timeout 300 python tools/ab_plans.py 10000 44100 48000 100000 1000000 > $OUT/ab_plans_generic.txt 2>&1; grep "complex64.*auto" $OUT/ab_plans_generic.txt
timeout 300 /usr/local/cuda/bin/ncu --set full --clock-control none --import-source on -k regex:run_cluster -s 4 -c 1 -o /tmp/full_cluster python tools/ab_plans.py 65536 > $OUT/ncu_cluster.log 2>&1
python tools/ncu_summary.py /tmp/full_cluster.ncu-rep > $OUT/ncu_full_cluster_256x256.md 2>&1; head -60 $OUT/ncu_full_cluster_256x256.md
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/s3/bench.json'))
print("value",d["value"],"frac",d["roofline"]["frac"])
print("per_size", [(r["log2n"], r["plan"], r["frac"], r.get("frac_b2b")) for r in d["config"]["per_size"]])
print("e2e",json.dumps(d["e2e"])[:400])
for r in d.get("other_configs") or []: print(json.dumps(r)[:600])
PY
