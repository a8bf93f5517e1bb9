#!/bin/bash
O=gpurun_out
N="ncu --set full --clock-control none --import-source on -f"
$N -k regex:"wide_tx_resident" -c 1 -o $O/wtxr python bench.py --workload tx --nfft 1024 --syms 128 --steps 2 > $O/wtxr_n.log 2>&1
python tools/ncu_summary.py $O/wtxr.ncu-rep > $O/r2_ncu_wide_tx.txt 2>&1
python tools/sass_by_line.py $O/wtxr.ncu-rep ofdm_b200/libofdm_b200.so wide_tx_resident_kernelILi2ELb1ELb1E 524288 > $O/r2_ncu_wide_tx_by_line.txt 2>&1
head -80 $O/r2_ncu_wide_tx.txt
