#!/bin/bash
# multi-GPU lines of the round: weak + strong scaling of the streams workload (device-resident and e2e with the measured
# H2D ceiling) and ONE capture split over the ranks. Usage: gpurun --gpus N -- 'bash scripts/r2_multi.sh N'
set -u
N=$1
O=gpurun_out; mkdir -p $O
{ nvidia-smi topo -m; nproc; free -g; } > $O/topo_n$N.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 20 --warmup 5 2> $O/multi_weak.err | tail -1 > $O/r2_bench_n${N}_weak.json
$TR bench.py --gpus $N --steps 20 --warmup 5 --scaling strong 2> $O/multi_strong.err | tail -1 > $O/r2_bench_n${N}_strong.json
$TR bench.py --gpus $N --workload capture --steps 10 2> $O/multi_cap.err | tail -1 > $O/r2_bench_n${N}_capture.json
python - <<EOF
import json
for k in ("weak", "strong", "capture"):
    try:
        d = json.load(open("$O/r2_bench_n${N}_%s.json" % k))
        print(k, d["value"], d["ms_per_step"], d.get("roofline", {}).get("frac"), (d.get("e2e") or {}), d.get("each_frame_found_once"), d.get("peaks_found"), d.get("ber"))
    except Exception as e:
        print(k, "failed", e)
EOF
tail -n 3 $O/multi_weak.err $O/multi_strong.err $O/multi_cap.err
