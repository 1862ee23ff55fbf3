#!/bin/bash
# multi-GPU session (gpurun --gpus N): the 2-rank NCCL scatter -> FFT -> gather parity test, then bench.py under torchrun on all N GPUs
# (weak-scaling sweep + e2e with NUMA binding + BASELINE config 5 with scatter / FFT / gather timed separately and checked)
N=${1:-2}
OUT=gpurun_out/multi$N
mkdir -p $OUT
export PYTHONPATH=$PWD:$PWD/tests
nvidia-smi --query-gpu=index,name,pci.bus_id --format=csv > $OUT/env.txt 2>&1
nvidia-smi topo -m >> $OUT/env.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sharded" > $OUT/pytest_sharded.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_sharded.log; tail -3 $OUT/pytest_sharded.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
tail -5 $OUT/bench.err
python - $OUT <<'PY'
import json, sys
d=json.loads([l for l in open(sys.argv[1] + '/bench.json') if l.startswith('{')][-1])
print("n_gpus", d["n_gpus"], "value", d["value"], "frac", d["roofline"]["frac"], "ms", d["ms_per_step"])
print("e2e", json.dumps(d["e2e"])[:600])
for r in d.get("other_configs") or []:
    r=dict(r); r.pop("per_size",None); r.pop("rows",None); print(json.dumps(r)[:900])
PY
