#!/bin/bash
# Same-box A/B of two builds of the engine: scripts/ab.sh ofdm_b200/lib_prev.so ofdm_b200/lib_new.so
# Alternates the two libraries under bench.py (kernel-only leg) and prints value / ms per step / roofline frac of each run;
# leaves the LAST library in place as ofdm_b200/libofdm_b200.so.
set -e
A=$1; B=$2; N=${3:-3}
for i in $(seq $N); do
  for L in $A $B; do
    cp $L ofdm_b200/libofdm_b200.so
    python bench.py --steps 50 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$L', d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['acquire_kernel_ms'], d['ms_per_step'], d['ber']['bit_errs'], d['clocks']['sm_mhz'])"
  done
done
