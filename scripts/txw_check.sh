#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tx_resident_kernel" > gpurun_out/txw_tests.log 2>&1
tail -4 gpurun_out/txw_tests.log
for pth in warp spec warp spec; do
OFDM_TX_PATH=$pth timeout 300 python bench.py --workload tx --steps 30 > gpurun_out/txw_$pth.json 2> gpurun_out/txw_$pth.err
python - <<P
import json
d=json.load(open("gpurun_out/txw_$pth.json")); print("$pth", d["ms_per_step"], d["value"], d["roofline"]["frac"], d["frames_match_oracle_on_sample"], d["gpu_launches"], d["clocks"]["sm_mhz"])
P
done
