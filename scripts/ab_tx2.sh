#!/bin/bash
run() { cp $1 ofdm_b200/libofdm_b200.so; python bench.py --workload tx --steps 30 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1 pad=$OFDM_TX_SMEM_PAD', d['ms_per_step'], d['roofline']['frac'], d['frames_match_oracle_on_sample'])"; }
run ofdm_b200/lib_4slot.so
OFDM_TX_SMEM_PAD=65536 run ofdm_b200/lib_4slot.so
run ofdm_b200/lib_5slot.so
run ofdm_b200/lib_5slot_late.so
run ofdm_b200/lib_4slot.so
