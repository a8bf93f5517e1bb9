#!/bin/bash
for f in full skipcp gather; do
OFDM_RX_FEED=$f timeout 300 python bench.py --nfft 1024 --syms 128 --steps 5 --no-cpu --no-secondary > gpurun_out/feedw_$f.json 2> gpurun_out/feedw_$f.err
python - <<P
import json
d=json.load(open("gpurun_out/feedw_$f.json"))["e2e"]; print("wide $f", d["value"], d["ms_per_step"], d["h2d_bytes_per_step"], d["h2d_gb_per_s_per_gpu"], d["matches_device_path"])
P
done
