#!/bin/bash
# wide (nfft 1024) TX: parity tests of the one-pass kernel, then the bench line of its two variants
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "wide_tx_resident" > gpurun_out/wtx_tests.log 2>&1
tail -5 gpurun_out/wtx_tests.log
for db in 0 1; do
OFDM_WTX_DB=$db timeout 300 python bench.py --workload tx --nfft 1024 --syms 128 --steps 20 > gpurun_out/wtx_db$db.json 2> gpurun_out/wtx_db$db.err
python - <<P
import json
d=json.load(open("gpurun_out/wtx_db$db.json")); print("db=$db", d["ms_per_step"], d["value"], d["roofline"]["frac"], d["frames_match_oracle_on_sample"], d["clocks"]["sm_mhz"])
P
done
