#!/bin/bash
# same-box A/B of engine builds on the TX workload: scripts/ab_tx.sh libA.so libB.so ...
for i in 1 2; do
for L in "$@"; do
  cp $L ofdm_b200/libofdm_b200.so
  python bench.py --workload tx --steps 30 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$L', d['ms_per_step'], d['value'], d['roofline']['frac'], d['frames_match_oracle_on_sample'])"
done
done
