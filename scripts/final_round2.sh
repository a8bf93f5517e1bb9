#!/bin/bash
# Round-2 measurement pass on one B200: GPU parity tests, every bench line, the launch list and `ncu --set full` captures of the
# RX kernels, the wideband kernels, the capture scan and the TX passes (each only after the same command has exited 0 plain).
# Usage: gpurun --timeout 2400 -- 'bash scripts/final_round2.sh'; then python scripts/collect_profiles.py r2
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/r2_pytest_gpu.txt
python bench.py --steps 20 --warmup 5             2> $O/bench.err  | tail -1 > $O/r2_bench.json
python bench.py --steps 200 --warmup 5 --no-secondary --no-cpu 2> $O/bench200.err | tail -1 > $O/r2_bench_200steps.json
python bench.py --impl reference --steps 20 --warmup 5 2> $O/ref.err | tail -1 > $O/r2_bench_reference_arm.json
python bench.py --nfft 1024 --syms 128 --steps 50 --no-secondary 2> $O/wide.err | tail -1 > $O/r2_bench_wide_n1024.json
python bench.py --workload capture --steps 20     2> $O/cap.err    | tail -1 > $O/r2_bench_capture_1e9.json
python bench.py --workload tx --steps 50          2> $O/tx.err     | tail -1 > $O/r2_bench_tx.json
python bench.py --workload rs --steps 50          2> $O/rs.err     | tail -1 > $O/r2_bench_rs.json
python bench.py --workload ingest                 2> $O/ingest.err | tail -1 > $O/r2_bench_ingest.json
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > $O/launch_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > $O/launch_ncu.log 2>&1
N="ncu --set full --clock-control none --import-source on -f"
python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-secondary > $O/p0.log 2>&1 && \
  $N -k regex:"rx_decode|rx_acquire" -c 2 -o $O/dec python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-secondary > $O/n0.log 2>&1
python tools/ncu_summary.py $O/dec.ncu-rep > $O/r2_ncu_rx_kernels_4096streams.txt 2>&1
python tools/sass_by_line.py $O/dec.ncu-rep ofdm_b200/libofdm_b200.so rx_decode_kernelILi2ELb1ELb1ELi1ELb0E 8347648 > $O/r2_ncu_decode_by_line.txt 2>&1
python bench.py --nfft 1024 --syms 128 --steps 2 --no-e2e --no-cpu --no-secondary > $O/p1.log 2>&1 && \
  $N -k regex:"wide_decode|wide_acquire" -c 2 -o $O/wide python bench.py --nfft 1024 --syms 128 --steps 2 --no-e2e --no-cpu --no-secondary > $O/n1.log 2>&1
python tools/ncu_summary.py $O/wide.ncu-rep > $O/r2_ncu_wide.txt 2>&1
python tools/sass_by_line.py $O/wide.ncu-rep ofdm_b200/libofdm_b200.so wide_decode_kernelILi2ELb1ELb1ELi1ELb0E 524288 > $O/r2_ncu_wide_by_line.txt 2>&1
python bench.py --workload capture --capture-samples 400000000 --steps 2 --no-cpu > $O/p2.log 2>&1 && \
  $N -k regex:sync_scan -c 1 -o $O/scan python bench.py --workload capture --capture-samples 400000000 --steps 2 --no-cpu > $O/n2.log 2>&1
python tools/ncu_summary.py $O/scan.ncu-rep > $O/r2_ncu_scan.txt 2>&1
python bench.py --workload tx --steps 2 > $O/p3.log 2>&1 && \
  $N -k regex:"tx_resident|tx_tile" -c 1 -o $O/tx python bench.py --workload tx --steps 2 > $O/n3.log 2>&1
python tools/ncu_summary.py $O/tx.ncu-rep > $O/r2_ncu_tx.txt 2>&1
python tools/sass_by_line.py $O/tx.ncu-rep ofdm_b200/libofdm_b200.so tx_resident_kernelILi2ELb1ELb1E 8347648 > $O/r2_ncu_tx_by_line.txt 2>&1
rm -f $O/*.ncu-rep
cat $O/r2_pytest_gpu.txt; head -c 700 $O/r2_bench.json; echo; wc -c $O/r2_*.json $O/r2_ncu_*.txt $O/launches.csv
