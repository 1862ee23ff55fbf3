#!/bin/bash
# session 16: the SmoothConv launch segmentation on hardware (general Rader / MixedRadix x Rader / Bluestein-over-smooth cases) + smoke()
export PYTHONPATH=$PWD:$PWD/tests
mkdir -p gpurun_out/s16
timeout 140 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "round2 and (rader or bluestein or default_prime or serialisation)" > gpurun_out/s16/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s16/pytest.log; tail -2 gpurun_out/s16/pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s16/smoke.log 2>&1; echo "smoke rc=$?"
