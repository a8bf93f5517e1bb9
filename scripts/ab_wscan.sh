#!/bin/bash
for L in "$@"; do
  cp $L ofdm_b200/libofdm_b200.so
  python bench.py --workload capture --nfft 1024 --syms 128 --capture-samples 1e9 --steps 10 --no-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$L', d['ms_per_step'], d['value'], d['roofline']['frac'], d['all_offsets_exact'], d['peaks_found'], d['each_frame_found_once'])"
done
