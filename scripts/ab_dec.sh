#!/bin/bash
# same-box A/B of engine builds on the headline workload + the decode parity tests for each
for L in "$@"; do
  cp $L ofdm_b200/libofdm_b200.so
  python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q -m gpu -k "not wide and not tx_ and not host_feed" 2>&1 | tail -1
  for i in 1 2; do
  python bench.py --steps 50 --warmup 3 --no-e2e --no-cpu --no-secondary 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$L', d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['acquire_kernel_ms'], d['ms_per_step'], d['ber']['bit_errs'], d['ber']['frames_failed'], d['clocks']['sm_mhz'])"
  done
done
