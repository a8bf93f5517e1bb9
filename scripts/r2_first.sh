#!/bin/bash
# round-2 first contact: topology probe, GPU parity tests, the default bench line (with secondary legs), reference arm, sanitizer
set -u
O=gpurun_out
mkdir -p $O
{ nvidia-smi topo -m; echo; numactl --hardware 2>&1; echo; lscpu | head -30; echo; nproc; free -g; } > $O/topo.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/r2_pytest_gpu.txt
cat $O/r2_pytest_gpu.txt
python bench.py --steps 20 --warmup 5 2> $O/bench.err | tail -1 > $O/r2_bench_first.json
head -c 1500 $O/r2_bench_first.json; echo; tail -3 $O/bench.err
python bench.py --impl reference --steps 20 --warmup 5 2> $O/ref.err | tail -1 > $O/r2_bench_reference_first.json
head -c 400 $O/r2_bench_reference_first.json; echo
bash scripts/sanitize.sh
