#!/bin/bash
# Runs every bench line of the round on the GPU box and leaves the JSON lines / launch list in gpurun_out/ (scratch);
# scripts/collect_profiles.py copies them into profiles/. Usage: gpurun --timeout 1500 -- 'bash scripts/refresh_profiles.sh'
set -u
O=gpurun_out
mkdir -p $O
python bench.py                                   2> $O/bench.err          | tail -1 > $O/r1_bench.json
python bench.py --impl reference --steps 3 --warmup 1 2> $O/ref.err        | tail -1 > $O/r1_bench_reference_arm.json
python bench.py --nfft 1024 --syms 128 --steps 50 2> $O/wide.err           | tail -1 > $O/r1_bench_wide_n1024.json
python bench.py --workload capture --steps 20     2> $O/cap.err            | tail -1 > $O/r1_bench_capture_1e9.json
python bench.py --workload tx --steps 50          2> $O/tx.err             | tail -1 > $O/r1_bench_tx.json
python bench.py --workload rs --steps 50          2> $O/rs.err             | tail -1 > $O/r1_bench_rs.json
python bench.py --workload ingest                 2> $O/ingest.err         | tail -1 > $O/r1_bench_ingest.json
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > $O/launch_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > $O/launch_ncu.log 2>&1
wc -c $O/r1_bench*.json $O/launches.csv
