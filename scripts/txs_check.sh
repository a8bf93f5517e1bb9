#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tx_resident or wideband_1024" > gpurun_out/txs_tests.log 2>&1
tail -4 gpurun_out/txs_tests.log
for pth in resident spec ""; do
OFDM_TX_PATH=$pth timeout 300 python bench.py --workload tx --nfft 1024 --syms 128 --steps 30 > gpurun_out/txs_w$pth.json 2> gpurun_out/txs_w$pth.err
python - <<P
import json
d=json.load(open("gpurun_out/txs_w$pth.json")); print("wide [$pth]", d["ms_per_step"], d["value"], d["roofline"]["frac"], d["frames_match_oracle_on_sample"], d["gpu_launches"], d["roofline"]["kernel"][:40])
P
done
timeout 300 python bench.py --workload tx --steps 30 > gpurun_out/txs_n.json 2> gpurun_out/txs_n.err
python - <<P
import json
d=json.load(open("gpurun_out/txs_n.json")); print("narrow default", d["ms_per_step"], d["value"], d["roofline"]["frac"], d["frames_match_oracle_on_sample"], d["gpu_launches"], d["roofline"]["kernel"][:40])
P
