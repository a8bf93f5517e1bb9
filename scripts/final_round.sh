#!/bin/bash
# End-of-round measurement pass on one B200: GPU parity tests, every bench line (refresh_profiles.sh), then one
# `ncu --set full` capture of the two RX kernels on the full 4096-stream workload (after the same command exited 0 plain).
# Usage: gpurun --timeout 1500 -- 'bash scripts/final_round.sh'; then python scripts/collect_profiles.py here.
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/final_pytest_gpu.txt
bash scripts/refresh_profiles.sh
python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $O/final_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -f -k regex:"rx_decode|rx_acquire" -c 2 -o $O/dec_r1final \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $O/final_ncu.log 2>&1
python tools/ncu_summary.py $O/dec_r1final.ncu-rep > $O/r1_ncu_rx_kernels_4096streams.txt 2>&1
python tools/sass_by_line.py $O/dec_r1final.ncu-rep ofdm_b200/libofdm_b200.so rx_decode_kernelILi2ELb1ELb1ELi1ELb0E 8347648 > $O/r1_ncu_decode_by_line.txt 2>&1
cat $O/final_pytest_gpu.txt; head -c 600 $O/r1_bench.json; echo; head -20 $O/r1_ncu_rx_kernels_4096streams.txt
