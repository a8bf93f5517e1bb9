#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_wide_capture.py -x -q -m gpu -k "wide or 1024" > gpurun_out/wacq_tests.log 2>&1
tail -3 gpurun_out/wacq_tests.log
timeout 300 python bench.py --nfft 1024 --syms 128 --steps 30 --no-cpu --no-e2e --no-secondary > gpurun_out/wacq.json 2> gpurun_out/wacq.err
python - <<P
import json
d=json.load(open("gpurun_out/wacq.json")); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["acquire_kernel_ms"], d["ber"])
P
