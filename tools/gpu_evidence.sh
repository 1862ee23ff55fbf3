#!/bin/bash
# evidence run: full GPU test-suite, bench lines (ours + reference arm), ncu launch list with DRAM bytes (caches left alone),
# ncu --set full of the top kernels.  gpurun copies back at most 64 MiB: reports are summarised on the box.
set -x
TAG=${1:-r2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONPATH=$PWD:$PWD/tests
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.draw,memory.total --format=csv > $OUT/env.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2>> $OUT/bench.err
# launch list of one whole step (--profile-cold: no warm-up sweep, one timed sweep = every launch of the step),
# caches left alone between launches so the L2 hand-off between the two passes is the real one
# (a whole step is ~9000 launches = 25 minutes under ncu: r2b spent most of its GPU budget here.  -c bounds it; every
#  kernel of the step appears within the first 7000 launches and tools/.. scale per-CTA means to the step)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
   -c 4000 --csv --log-file $OUT/launches.csv python bench.py --profile --profile-cold --steps 1 --logs 10,11,12,13,14,15,16,18,20 > $OUT/ncu_list.log 2>&1
python tools/launch_list_summary.py $OUT/launches.csv > $OUT/launch_list_summary.md 2>&1
gzip -9 $OUT/launches.csv
# --set full: Direct{4096}, then the two TMA passes of 1024x1024 (third chunk pair)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:run_kernel -c 1 -o /tmp/full_direct python bench.py --profile --steps 1 --logs 12 > $OUT/ncu_full_direct4096.log 2>&1
python tools/ncu_summary.py /tmp/full_direct.ncu-rep > $OUT/ncu_full_direct4096.md 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:run_kernel_tma -s 8 -c 2 -o /tmp/full_tma python bench.py --profile --steps 1 --logs 20 > $OUT/ncu_full_tma_1024.log 2>&1
python tools/ncu_summary.py /tmp/full_tma.ncu-rep > $OUT/ncu_full_tma_1024.md 2>&1
ncu -i /tmp/full_tma.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > $OUT/ncu_full_tma_1024_raw.csv.gz
du -sh gpurun_out; ls -la $OUT
