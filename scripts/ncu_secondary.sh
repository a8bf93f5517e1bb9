#!/bin/bash
# ncu --set full captures of the secondary kernels (wideband decode, capture scan, TX, RS); each only after the same
# command has exited 0 without ncu. Usage: gpurun --timeout 1500 -- 'bash scripts/ncu_secondary.sh'
set -u
O=gpurun_out
N="ncu --set full --clock-control none --import-source on -f"
python bench.py --nfft 1024 --syms 128 --steps 3 --no-e2e --no-cpu > $O/p1.log 2>&1 && $N -k regex:wide_decode -c 1 -o $O/sec_wide python bench.py --nfft 1024 --syms 128 --steps 3 --no-e2e --no-cpu > $O/n1.log 2>&1
python bench.py --workload capture --capture-samples 200000000 --steps 3 > $O/p2.log 2>&1 && $N -k regex:sync_scan -c 1 -o $O/sec_scan python bench.py --workload capture --capture-samples 200000000 --steps 3 > $O/n2.log 2>&1
python bench.py --workload tx --steps 2 > $O/p3.log 2>&1 && $N -k regex:tx_tile -c 2 -o $O/sec_tx python bench.py --workload tx --steps 2 > $O/n3.log 2>&1
python bench.py --workload rs --steps 2 > $O/p4.log 2>&1 && $N -k regex:rs_ -c 2 -o $O/sec_rs python bench.py --workload rs --steps 2 > $O/n4.log 2>&1
ls -la $O/sec_*.ncu-rep
# keep the text summaries only (the reports exceed what gpurun copies back)
for k in wide scan tx rs; do python tools/ncu_summary.py $O/sec_$k.ncu-rep > $O/r1_ncu_$k.txt 2>&1; done
rm -f $O/sec_*.ncu-rep
