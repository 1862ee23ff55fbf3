#!/bin/bash
# session 12: final evidence of the round on the final build -- full GPU suite, both bench arms, ncu launch list of one whole step with DRAM
# bytes (caches left alone), --set full of the dominant kernel (fused four-step at 2^20)
OUT=gpurun_out/s12
mkdir -p $OUT
export PYTHONPATH=$PWD:$PWD/tests
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.draw,memory.total --format=csv > $OUT/env.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
timeout 600 /usr/local/cuda/bin/ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
   -c 200 --csv --log-file $OUT/launches.csv python bench.py --profile --profile-cold --steps 1 > $OUT/ncu_list.log 2>&1
python tools/launch_list_summary.py $OUT/launches.csv > $OUT/launch_list_summary.md 2>&1; gzip -9 -f $OUT/launches.csv; tail -14 $OUT/launch_list_summary.md | cut -c1-200
timeout 400 /usr/local/cuda/bin/ncu --set full --clock-control none --import-source on -k regex:run_fused -c 1 -o /tmp/full_fused python bench.py --profile --steps 1 --logs 20 > $OUT/ncu_fused.log 2>&1
python tools/ncu_summary.py /tmp/full_fused.ncu-rep > $OUT/ncu_full_fused_1024x1024.md 2>&1
python - <<'PY'
import json
d=json.load(open('gpurun_out/s12/bench.json'))
print("value",d["value"],"frac",d["roofline"]["frac"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], "launches", d["gpu_launches"], d["clocks"])
print("per_size", [(r["log2n"], r["frac"], r.get("frac_b2b")) for r in d["config"]["per_size"]])
for r in d.get("other_configs") or []:
    r=dict(r); r.pop("per_size",None); r.pop("rows",None); print(json.dumps(r)[:400])
PY
