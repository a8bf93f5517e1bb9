#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "host_feed" > gpurun_out/feed_tests.log 2>&1
tail -15 gpurun_out/feed_tests.log
for f in auto; do
timeout 300 python bench.py --steps 5 --no-cpu --no-secondary > gpurun_out/feed_$f.json 2> gpurun_out/feed_$f.err
python - <<P
import json
d=json.load(open("gpurun_out/feed_$f.json")); print("$f", json.dumps(d["e2e"]))
P
tail -2 gpurun_out/feed_$f.err
done
