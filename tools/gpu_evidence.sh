#!/bin/bash
# evidence run: bench lines (ours + reference arm), ncu launch list with DRAM bytes, ncu --set full of the top kernels.
# gpurun copies back at most 64 MiB: reports are summarised on the box and only text/gz comes home.
set -x
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2>> $OUT/bench.err
# one sweep = 13737 launches; profile the second (timed) sweep only, caches left alone so the L2 hand-off is real
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
   -s 13737 -c 13737 --csv --log-file $OUT/launches.csv python bench.py --profile --steps 1 > $OUT/ncu_list.log 2>&1
python tools/launch_list_summary.py $OUT/launches.csv > $OUT/launch_list_summary.md 2>&1
gzip -9 $OUT/launches.csv
for spec in "4096_n20 12,20 2049" "8192_n16 13,16 130"; do
  set -- $spec
  timeout 600 ncu --set full --clock-control none -k regex:run_ -s $3 -c 3 -o /tmp/full_$1 python bench.py --profile --steps 1 --logs $2 > $OUT/ncu_full_$1.log 2>&1
  python tools/ncu_summary.py /tmp/full_$1.ncu-rep > $OUT/ncu_full_$1.md 2>&1
  ncu -i /tmp/full_$1.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > $OUT/ncu_full_$1_raw.csv.gz
done
du -sh gpurun_out; ls -la $OUT
