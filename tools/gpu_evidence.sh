#!/bin/bash
# evidence run: bench lines (ours + reference arm), ncu launch list with DRAM bytes, ncu --set full of the top kernels.
# gpurun copies back at most 64 MiB: reports are summarised on the box and only text/gz comes home.
set -x
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2>> $OUT/bench.err
# launch list of one whole step (--profile-cold: no warm-up sweep, one timed sweep = every launch of the step),
# caches left alone between launches so the L2 hand-off between the two passes is the real one
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
   --csv --log-file $OUT/launches.csv python bench.py --profile-cold --steps 1 > $OUT/ncu_list.log 2>&1
python tools/launch_list_summary.py $OUT/launches.csv > $OUT/launch_list_summary.md 2>&1
gzip -9 $OUT/launches.csv
for spec in "4096_n20 12,20 2049" "8192_n16 13,16 130"; do
  set -- $spec
  timeout 600 ncu --set full --clock-control none -k regex:run_ -s $3 -c 3 -o /tmp/full_$1 python bench.py --profile --steps 1 --logs $2 > $OUT/ncu_full_$1.log 2>&1
  python tools/ncu_summary.py /tmp/full_$1.ncu-rep > $OUT/ncu_full_$1.md 2>&1
  ncu -i /tmp/full_$1.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > $OUT/ncu_full_$1_raw.csv.gz
done
python - > $OUT/smooth_timing.txt 2>&1 <<'PY'
import numpy as np, torch, sys
sys.path.insert(0, '.')
import rustfft_b200 as rb
pl = rb.FftPlanner(np.complex64)
def t(n, batch):
    f = pl.plan_fft_forward(n)
    x = torch.view_as_complex(torch.rand(batch*n, 2, device='cuda')).contiguous(); y = torch.empty_like(x)
    for _ in range(2): f.process_device(x, out=y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): f.process_device(x, out=y)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/5
    print(f"n={n:9d} batch={batch:6d} {f.describe():50s} {ms:9.4f} ms  {16.0*n*batch/ms/1e6:8.1f} GB/s  frac={16.0*n*batch/ms/1e6/6487.4:.3f}", flush=True)
for n in [100, 360, 1000, 1200, 1536, 2000, 2401, 3000, 3600, 4000, 143, 961, 1131, 997, 1234]:
    t(n, (1<<28)//n//8)
PY
du -sh gpurun_out; ls -la $OUT
