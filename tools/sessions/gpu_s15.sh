#!/bin/bash
# session 15: ncu --set full of the composite-length kernels (compiled 1000-point passes, one-pass Rader over a smooth inner FFT, one-pass Smooth)
OUT=gpurun_out/s15
mkdir -p $OUT
export PYTHONPATH=$PWD:$PWD/tests
NCU=/usr/local/cuda/bin/ncu
timeout 200 $NCU --set full --clock-control none --import-source on -k regex:run_kernel -s 8 -c 2 -o /tmp/c1 python tools/ab_plans.py 1000000 > $OUT/l1.log 2>&1
python tools/ncu_summary.py /tmp/c1.ncu-rep > $OUT/ncu_full_compiled_tiles_1000x1000.md 2>&1
timeout 200 $NCU --set full --clock-control none --import-source on -k regex:run_kernel_loop -s 3 -c 1 -o /tmp/c2 python tools/ab_plans.py 617 > $OUT/l2.log 2>&1
python tools/ncu_summary.py /tmp/c2.ncu-rep > $OUT/ncu_full_smoothconv_rader617.md 2>&1
timeout 200 $NCU --set full --clock-control none --import-source on -k regex:run_kernel_dyn -s 3 -c 1 -o /tmp/c3 python tools/ab_plans.py 1000 > $OUT/l3.log 2>&1
python tools/ncu_summary.py /tmp/c3.ncu-rep > $OUT/ncu_full_smooth_1000.md 2>&1
grep -h -E "^## |duration|dram %|issue active|achieved occ|stall" $OUT/*.md | cut -c1-150
