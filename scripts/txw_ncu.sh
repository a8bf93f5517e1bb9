#!/bin/bash
O=gpurun_out
N="ncu --set full --clock-control none --import-source on -f"
$N -k regex:"tx_warp" -c 1 -o $O/txw python bench.py --workload tx --steps 2 > $O/txw_n.log 2>&1
python tools/ncu_summary.py $O/txw.ncu-rep > $O/r2_ncu_tx_warp.txt 2>&1
python tools/sass_by_line.py $O/txw.ncu-rep ofdm_b200/libofdm_b200.so tx_warp_kernelILi2ELb1ELb1E 8347648 > $O/r2_ncu_tx_warp_by_line.txt 2>&1
grep -E "dram__bytes|duration|bank_conflicts|wavefronts_mem_shared.sum.pct|inst_executed_pipe|warps_active|stalled_(long|short|wait|math|mio|not_sel|lg|barrier|branch|dispatch)|inst_executed.sum |issue_active" $O/r2_ncu_tx_warp.txt
