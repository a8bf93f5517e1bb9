#!/bin/bash
# Round-2 closing measurement pass on one B200 (second half of the round: wideband TX, barrier-free TX, host feed): GPU parity
# tests, the driver's bench command, the reference arm, every per-workload line, the launch list, and `ncu --set full`
# captures of the kernels that changed (each only after the same command has exited 0 plain).
# Usage: gpurun --timeout 2400 -- 'bash scripts/final_round2b.sh'; then python scripts/collect_profiles.py r2
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/r2_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
python bench.py --steps 20 --warmup 5             2> $O/bench.err  | tail -1 > $O/r2_bench.json
python bench.py --impl reference --steps 20 --warmup 5 2> $O/ref.err | tail -1 > $O/r2_bench_reference_arm.json
python bench.py --nfft 1024 --syms 128 --steps 50 --no-secondary 2> $O/wide.err | tail -1 > $O/r2_bench_wide_n1024.json
python bench.py --workload tx --steps 50          2> $O/tx.err     | tail -1 > $O/r2_bench_tx.json
python bench.py --workload tx --nfft 1024 --syms 128 --steps 50 2> $O/txw.err | tail -1 > $O/r2_bench_tx_wide_n1024.json
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > $O/launch_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > $O/launch_ncu.log 2>&1
N="ncu --set full --clock-control none --import-source on -f"
$N -k regex:"tx_spec_kernel" -c 1 -o $O/txw python bench.py --workload tx --steps 2 > $O/n3.log 2>&1
python tools/ncu_summary.py $O/txw.ncu-rep > $O/r2_ncu_tx_spec.txt 2>&1
python tools/sass_by_line.py $O/txw.ncu-rep ofdm_b200/libofdm_b200.so tx_spec_kernelILi2ELb1ELb1E 8347648 > $O/r2_ncu_tx_spec_by_line.txt 2>&1
$N -k regex:"wide_tx_spec" -c 1 -o $O/wtxr python bench.py --workload tx --nfft 1024 --syms 128 --steps 2 > $O/n4.log 2>&1
python tools/ncu_summary.py $O/wtxr.ncu-rep > $O/r2_ncu_wide_tx_spec.txt 2>&1
python tools/sass_by_line.py $O/wtxr.ncu-rep ofdm_b200/libofdm_b200.so wide_tx_spec_kernelILi2ELb1ELb1E 524288 > $O/r2_ncu_wide_tx_spec_by_line.txt 2>&1
rm -f $O/*.ncu-rep
cat $O/r2_pytest_gpu.txt $O/smoke.log; head -c 600 $O/r2_bench.json; echo; wc -c $O/r2_bench*.json $O/r2_ncu_tx_spec.txt $O/r2_ncu_wide_tx_spec.txt $O/launches.csv
