#!/bin/bash
# evidence run: bench lines (ours + reference arm), ncu launch list with DRAM bytes, ncu --set full of the top kernels
set -x
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?" >> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
# one sweep = 13737 launches (run-time check below); profile the second (timed) sweep only
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
   -s 13737 -c 13737 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --profile --steps 1 > gpurun_out/${TAG}_ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:run_ -s 2049 -c 3 -o gpurun_out/${TAG}_full_4096_n20 python bench.py --profile --steps 1 --logs 12,20 > gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:run_ -s 130 -c 3 -o gpurun_out/${TAG}_full_13_16 python bench.py --profile --steps 1 --logs 13,16 > gpurun_out/${TAG}_ncu_full2.log 2>&1
ls -la gpurun_out | tail -12
