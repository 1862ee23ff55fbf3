#!/bin/bash
# DRAM traffic of the two-pass plans in normal operation (caches NOT flushed between launches)
set -x
OUT=gpurun_out/r1x
mkdir -p $OUT
export PYTHONPATH=$PWD:$PWD/tests
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum
for lg in 20 16; do
B200FFT_FLOW=1 timeout 300 ncu --metrics $M --clock-control none --cache-control none -k regex:run_flow -s 2 -c 2 --csv --log-file $OUT/flow_$lg.csv python tools/ab_two_pass.py $lg > $OUT/flow_$lg.log 2>&1
timeout 600 ncu --metrics $M --clock-control none --cache-control none -k regex:run_kernel -s 1200 -c 600 --csv --log-file $OUT/chunk_$lg.csv python tools/ab_two_pass.py $lg > $OUT/chunk_$lg.log 2>&1
done
python - <<'PY'
import csv,glob,collections
for f in sorted(glob.glob('gpurun_out/r1x/*.csv')):
    rows=[r for r in csv.reader(open(f)) if len(r)>10]
    hdr=rows[0]; 
    iname=hdr.index('Metric Name'); ival=hdr.index('Metric Value'); ik=hdr.index('Kernel Name'); iid=hdr.index('ID')
    tot=collections.defaultdict(float); per=collections.defaultdict(lambda: collections.defaultdict(float)); n=set()
    for r in rows[1:]:
        v=float(r[ival].replace(',',''))
        tot[r[iname]]+=v; n.add(r[iid])
        k='A' if 'LoadCols' in r[ik] else ('B' if 'LoadRowsTw' in r[ik] else 'flow')
        per[k][r[iname]]+=v
    print(f, 'launches', len(n))
    for k,d in per.items(): print('   ',k, {m: f"{v:.4g}" for m,v in d.items()})
PY
gzip -9 $OUT/*.csv
