#!/bin/bash
# L2 persistence window over the two-pass intermediate: timings and DRAM traffic
set -x
OUT=gpurun_out/r1y
mkdir -p $OUT
export PYTHONPATH=$PWD:$PWD/tests
python - > $OUT/props.txt 2>&1 <<'PY'
from cuda.bindings import runtime as rt
err, p = rt.cudaGetDeviceProperties(0)
print("l2CacheSize", p.l2CacheSize, "persistingL2CacheMaxSize", p.persistingL2CacheMaxSize, "accessPolicyMaxWindowSize", p.accessPolicyMaxWindowSize)
PY
cat $OUT/props.txt
B200FFT_L2_PERSIST_MB=64 timeout 300 python tests/variant_check.py > $OUT/variant.log 2>&1; tail -1 $OUT/variant.log
for v in "B200FFT_L2_PERSIST_MB=32" "B200FFT_L2_PERSIST_MB=64" "B200FFT_L2_PERSIST_MB=96" "B200FFT_L2_PERSIST_MB=64 B200FFT_CHUNK_MB=48" "B200FFT_L2_PERSIST_MB=64 B200FFT_TMA_TILES=1" "B200FFT_FLOW=1 B200FFT_L2_PERSIST_MB=48" "B200FFT_FLOW=1 B200FFT_L2_PERSIST_MB=64" "B200FFT_FLOW=1 B200FFT_L2_PERSIST_MB=96"; do
  env $v timeout 200 python tools/ab_two_pass.py 15,16,17,18,19,20 >> $OUT/ab.log 2>&1
done
grep SUMMARY $OUT/ab.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
B200FFT_FLOW=1 B200FFT_L2_PERSIST_MB=64 timeout 300 ncu --metrics $M --clock-control none --cache-control none -k regex:run_flow -s 2 -c 2 --csv --log-file $OUT/flow_20.csv python tools/ab_two_pass.py 20 > $OUT/flow_20.log 2>&1
B200FFT_L2_PERSIST_MB=64 timeout 600 ncu --metrics $M --clock-control none --cache-control none -k regex:run_kernel -s 1200 -c 600 --csv --log-file $OUT/chunk_20.csv python tools/ab_two_pass.py 20 > $OUT/chunk_20.log 2>&1
python - <<'PY'
import csv,glob,collections
for f in sorted(glob.glob('gpurun_out/r1y/*.csv')):
    rows=[r for r in csv.reader(open(f)) if len(r)>10]
    hdr=rows[0]
    iname=hdr.index('Metric Name'); ival=hdr.index('Metric Value'); ik=hdr.index('Kernel Name'); iid=hdr.index('ID')
    per=collections.defaultdict(lambda: collections.defaultdict(float)); n=set()
    for r in rows[1:]:
        v=float(r[ival].replace(',','')); n.add(r[iid])
        k='A' if 'LoadCols' in r[ik] and 'run_flow' not in r[ik] else ('B' if 'LoadRowsTw' in r[ik] and 'run_flow' not in r[ik] else 'flow')
        per[k][r[iname]]+=v
    print(f, 'launches', len(n))
    for k,d in per.items(): print('   ',k, {m: f"{v:.4g}" for m,v in d.items()})
PY
gzip -9 $OUT/*.csv
