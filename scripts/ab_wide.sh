#!/bin/bash
# same-box A/B of two engine builds on the wideband workload: scripts/ab_wide.sh lib_prev.so lib_new.so [rounds]
A=$1; B=$2; N=${3:-3}
for i in $(seq $N); do
  for L in $A $B; do
    cp $L ofdm_b200/libofdm_b200.so
    python bench.py --nfft 1024 --syms 128 --steps 30 --warmup 3 --no-e2e --no-cpu --no-secondary 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$L', d['value'], d['roofline']['kernel_ms'], d['roofline']['acquire_kernel_ms'], d['ms_per_step'], d['ber']['bit_errs'], d['clocks']['sm_mhz'])"
  done
done
